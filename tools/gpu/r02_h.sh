#!/bin/bash
# Round 2, run H: warp pipeline for one-voice groups (config 5) — parity, then the stage-count A/B; offline path with pinned staging.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "pipelined or offline or many_voice or host_delivery or fuzz" 2>&1 | tail -15 | cut -c1-400 | tee gpurun_out/r02h_pytest.txt
for st in 0 2 3 4; do
  timeout 600 python bench_configs.py 5 --stages $st > gpurun_out/r02h_config5_s$st.json 2> gpurun_out/r02h_config5_s$st.err || tail -3 gpurun_out/r02h_config5_s$st.err
done
python - <<'PY'
import json
for st in (0, 2, 3, 4):
    try:
        d = json.load(open(f"gpurun_out/r02h_config5_s{st}.json"))
        print("stages", st, d["pipeline_stages"], "ms/block", round(d["ms_per_block"], 4), "Msamples/s", round(d["msamples_per_s"], 1), "offline Msamples/s", round(d["offline"]["msamples_per_s"], 1), "parity", d["parity"]["worst_err_over_tol"] if d["parity"] else None)
    except Exception as e:
        print("stages", st, "FAILED", e)
PY
