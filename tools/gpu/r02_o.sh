#!/bin/bash
# Round 2, run O: recurrence loops with the next sample's operands fetched one step ahead: parity, opcode profile, config 5
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | cut -c1-400 | tee gpurun_out/r02o_pytest.txt
ELEM_B200_SPECIALIZE=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "pole or biquad or mm1p or env or fuzz" 2>&1 | tail -3 | cut -c1-400 | tee gpurun_out/r02o_pytest_spec.txt
ELEM_B200_LIB=$PWD/elementary_b200/libelem_b200_prof.so python tools/opprof.py 1250 0 0 | tee gpurun_out/opprof_o_s0.txt | head -14
for cfg in "0 0" "3 0" "4 0"; do set -- $cfg
  timeout 600 python bench_configs.py 5 --stages $1 --niter $2 > gpurun_out/r02o_config5_s$1_n$2.json 2> gpurun_out/r02o_config5_s$1_n$2.err || tail -3 gpurun_out/r02o_config5_s$1_n$2.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02o_config5_s*_n*.json")):
    try:
        for line in open(f):
            d = json.loads(line)
            if d["config"].startswith("5"):
                print(f.split("/")[-1], d["pipeline_stages"], "ms/block", round(d["ms_per_block"], 4), "Msamples/s", round(d["msamples_per_s"], 1), "offline Msamples/s", round(d["offline"]["msamples_per_s"], 1), "parity", round(d["parity"]["worst_err_over_tol"], 4) if d["parity"] else None)
    except Exception as e:
        print(f, "FAILED", e)
PY
