#!/bin/bash
# Round 2, run F: host delivery of the mix bus (process() polls mapped host memory), fixed split test, bench, then the ncu captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02f_pytest.txt
python bench.py --steps 100 --warmup 10 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -2 gpurun_out/r02f_bench.err
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-t1 --no-configs --opt host_deliver=0 > gpurun_out/r02f_bench_copy_path.json 2>/dev/null
python - <<'PY'
import json
for f in ("r02f_bench", "r02f_bench_copy_path"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "parity", d.get("parity_ok"), d["kernel_ms"])
PY
bash tools/gpu/profile_all.sh r02f
