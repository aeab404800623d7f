#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { # name, args...
  name=$1; shift
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/x_$name.json 2>gpurun_out/x_$name.err || tail -3 gpurun_out/x_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/x_$name.json"))
    print("$name", "Msamples/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "k1 ms", round(d["roofline"]["kernel_ms"],4))
except Exception as e: print("$name fail", e)
PY
}
run v131072_n8 --voices 131072
run v131072_n4 --voices 131072 --opt niter=4
run v131072_n8_w2 --voices 131072 --opt warps_per_cta=2
run v131072_n4_w2 --voices 131072 --opt niter=4 --opt warps_per_cta=2
run v131072_n4_w1 --voices 131072 --opt niter=4 --opt warps_per_cta=1
run v4096_tw2 --tile-width 2
run v4096_tw4 --tile-width 4
run v4096_tw2_w2 --tile-width 2 --opt warps_per_cta=2
run v4096_tw2_w1 --tile-width 2 --opt warps_per_cta=1
