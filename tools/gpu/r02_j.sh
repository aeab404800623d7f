#!/bin/bash
# Round 2, run J: recurrences of narrow tiles walked by every lane of the voice column (shuffle-fed, no shared memory on the dependent
# path), rand as an LCG jump, mm1p division hoisted: whole GPU suite (both kernels), opcode profile, config 5 stage A/B, headline bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | cut -c1-400 | tee gpurun_out/r02j_pytest.txt
ELEM_B200_SPECIALIZE=1 timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "pole or biquad or mm1p or rand or fuzz or subsynth or env" 2>&1 | tail -5 | cut -c1-400 | tee gpurun_out/r02j_pytest_spec.txt
bash tools/gpu/opprof.sh 2>&1 | grep -v "^$" | head -40
for st in 0 3 4; do
  timeout 600 python bench_configs.py 5 --stages $st > gpurun_out/r02j_config5_s$st.json 2> gpurun_out/r02j_config5_s$st.err || tail -3 gpurun_out/r02j_config5_s$st.err
done
python - <<'PY'
import json
for st in (0, 3, 4):
    try:
        for line in open(f"gpurun_out/r02j_config5_s{st}.json"):
            d = json.loads(line)
            if d["config"].startswith("5"):
                print("stages", st, d["pipeline_stages"], "ms/block", round(d["ms_per_block"], 4), "Msamples/s", round(d["msamples_per_s"], 1), "offline Msamples/s", round(d["offline"]["msamples_per_s"], 1), "parity", d["parity"]["worst_err_over_tol"] if d["parity"] else None)
    except Exception as e:
        print("stages", st, "FAILED", e)
PY
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-configs > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; tail -2 gpurun_out/r02j_bench.err; cut -c1-250 gpurun_out/r02j_bench.json
