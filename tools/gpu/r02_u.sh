#!/bin/bash
# Round 2, run U: register budget of the pipeline kernel (launch bounds 8 / 6 / 5 / 4 CTAs per SM = 64 / 80 / 96 / 128 registers) on config 5
mkdir -p gpurun_out
for v in 8 6 5 4; do
  lib=$PWD/elementary_b200/libelem_b200_mb$v.so; [ $v = 8 ] && lib=$PWD/elementary_b200/libelem_b200.so
  for st in 4 3; do
  ELEM_B200_LIB=$lib timeout 600 python bench_configs.py 5 --stages $st > gpurun_out/r02u_config5_mb${v}_s$st.json 2> gpurun_out/r02u_config5_mb${v}_s$st.err || tail -3 gpurun_out/r02u_config5_mb${v}_s$st.err
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02u_config5_mb*_s*.json")):
    try:
        for line in open(f):
            d = json.loads(line)
            if d["config"].startswith("5"):
                print(f.split("/")[-1], d["pipeline_stages"], "ms/block", round(d["ms_per_block"], 4), "Msamples/s", round(d["msamples_per_s"], 1), "offline Msamples/s", round(d["offline"]["msamples_per_s"], 1), "parity", round(d["parity"]["worst_err_over_tol"], 4) if d["parity"] else None)
    except Exception as e:
        print(f, "FAILED", e)
PY
