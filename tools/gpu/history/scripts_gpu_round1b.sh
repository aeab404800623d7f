#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for tw in 1 2 4 8 32; do
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline --tile-width $tw > gpurun_out/b_tw$tw.json 2>gpurun_out/b_tw$tw.err || tail -3 gpurun_out/b_tw$tw.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/b_tw$tw.json"))
    print("4096 voices tile_width", $tw, "Msamples/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "k1 ms", round(d["roofline"]["kernel_ms"],4), "e2e", round(d["e2e"]["value"]))
except Exception as e: print("fail", e)
PY
done
for v in 32768 131072 262144; do
  python bench.py --steps 30 --warmup 5 --voices $v --no-cpu-baseline > gpurun_out/bench_v$v.json 2>gpurun_out/bench_v$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_v$v.json"))
print($v, "Msamples/s", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "k1 ms", round(d["roofline"]["kernel_ms"],3), "frac", round(d["roofline"]["frac"],4), "L", d["config"]["tile_width"], "rt x", round(d["realtime_factor"],1))
PY
done
