#!/bin/bash
# A/B of K1 build variants (ELEM_B200_LIB selects the library): ms per block at 4096 and 131072 voices
mkdir -p gpurun_out
for name in base SIN4 BOTH; do
  lib=elementary_b200/libelem_b200_$name.so; [ $name = base ] && lib=elementary_b200/libelem_b200.so
  for v in 4096 131072; do
    ELEM_B200_LIB=$PWD/$lib python bench.py --steps 100 --warmup 10 --voices $v --no-cpu-baseline > gpurun_out/var_${name}_$v.json 2>/dev/null
    python - <<PY
import json
d = json.load(open("gpurun_out/var_${name}_$v.json"))
print("$name", $v, "ms/step", round(d["ms_per_step"], 4), "k1 ms", round(d["roofline"]["kernel_ms"], 4))
PY
  done
done
