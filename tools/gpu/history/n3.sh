#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/n3_pytest.txt
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/n3_bench.json 2> gpurun_out/n3_bench.err; tail -2 gpurun_out/n3_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/n3_bench.json"))
print("4096 voices: Msamples/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "k1 ms", round(d["roofline"]["kernel_ms"], 4), "e2e", round(d["e2e"]["value"]))
PY
python bench.py --steps 30 --warmup 5 --voices 131072 --no-cpu-baseline > gpurun_out/n3_v131072.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/n3_v131072.json"))
print("131072 voices: Msamples/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "k1 ms", round(d["roofline"]["kernel_ms"], 4))
PY
