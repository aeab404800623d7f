#!/bin/bash
mkdir -p gpurun_out
run() { # name, lib, args...
  name=$1; lib=$2; shift; shift
  ELEM_B200_LIB=$PWD/elementary_b200/$lib python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/y_$name.json 2>gpurun_out/y_$name.err || tail -3 gpurun_out/y_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/y_$name.json"))
    print("$name", "Msamples/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "k1 ms", round(d["roofline"]["kernel_ms"],4))
except Exception as e: print("$name fail", e)
PY
}
for mb in mb5 mb6 mb8; do
  run v131072_$mb libelem_b200_$mb.so --voices 131072
  run v4096_$mb libelem_b200_$mb.so
  run v32768_$mb libelem_b200_$mb.so --voices 32768
done
run v32768_mb4 libelem_b200.so --voices 32768
