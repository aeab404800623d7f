#!/bin/bash
# Round 2, run M: A/B of the narrow-tile recurrence form (shuffle-fed vs broadcast reads) and of the interpreter loop (plain vs
# software-pipelined), calibrated pipeline cost model.  v0 = product library (shuffle-fed, plain loop), v1 = + prefetching loop,
# v2 = broadcast reads.  Parity first (all libraries), then config 5 at 0 / 3 / 4 stages.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | cut -c1-400 | tee gpurun_out/r02m_pytest.txt
ELEM_B200_SPECIALIZE=1 timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "pole or biquad or mm1p or rand or fuzz or subsynth or env or delay" 2>&1 | tail -5 | cut -c1-600 | tee gpurun_out/r02m_pytest_spec.txt
for v in v1 v2; do ELEM_B200_LIB=$PWD/elementary_b200/libelem_b200_$v.so timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -x -q -k "not soak and not thread and not registered" 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/r02m_pytest_$v.txt; done
ELEM_B200_LIB=$PWD/elementary_b200/libelem_b200_prof.so python tools/opprof.py 1250 0 0 | tee gpurun_out/opprof_m_s0.txt | head -14
ELEM_B200_LIB=$PWD/elementary_b200/libelem_b200_prof.so python tools/opprof.py 1250 3 0 | tee gpurun_out/opprof_m_s3.txt | head -3
for v in v0 v1 v2; do for st in 0 3 4; do
  lib=$PWD/elementary_b200/libelem_b200_$v.so; [ $v = v0 ] && lib=$PWD/elementary_b200/libelem_b200.so
  ELEM_B200_LIB=$lib timeout 600 python bench_configs.py 5 --stages $st > gpurun_out/r02m_config5_${v}_s$st.json 2> gpurun_out/r02m_config5_${v}_s$st.err || tail -3 gpurun_out/r02m_config5_${v}_s$st.err
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02m_config5_v*_s*.json")):
    try:
        for line in open(f):
            d = json.loads(line)
            if d["config"].startswith("5"):
                print(f.split("/")[-1], d["pipeline_stages"], "ms/block", round(d["ms_per_block"], 4), "Msamples/s", round(d["msamples_per_s"], 1), "offline Msamples/s", round(d["offline"]["msamples_per_s"], 1), "parity", round(d["parity"]["worst_err_over_tol"], 4) if d["parity"] else None)
    except Exception as e:
        print(f, "FAILED", e)
PY
