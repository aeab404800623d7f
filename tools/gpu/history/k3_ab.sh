#!/bin/bash
mkdir -p gpurun_out
for name in base s2c5 s2c4 s3c3; do
  lib=elementary_b200/libelem_b200_$name.so; [ $name = base ] && lib=elementary_b200/libelem_b200.so
  ELEM_B200_LIB=$PWD/$lib timeout 300 python bench_configs.py 4 > gpurun_out/k3ab_$name.jsonl 2>/dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/k3ab_$name.jsonl").readline())
print("$name", "ms/block", round(d["ms_per_block"], 4), "k1", round(d["k1_ms"], 4), "k3", round(d["k3_ms"], 4), "frac", round(d["roofline"]["frac"], 3))
PY
done
