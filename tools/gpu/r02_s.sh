#!/bin/bash
# Round 2, run S: config 5, warps per CTA of the many-graphs launch (1250 one-warp tiles over 148 SMs: CTA granularity decides how evenly
# the warps spread over the SMs), with and without the 4-stage pipeline
mkdir -p gpurun_out
for cfg in "1 0" "2 0" "4 0" "1 4"; do set -- $cfg
  timeout 600 python bench_configs.py 5 --wpc $1 --stages $2 > gpurun_out/r02s_config5_wpc$1_s$2.json 2> gpurun_out/r02s_config5_wpc$1_s$2.err || tail -3 gpurun_out/r02s_config5_wpc$1_s$2.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02s_config5_wpc*_s*.json")):
    try:
        for line in open(f):
            d = json.loads(line)
            if d["config"].startswith("5"):
                print(f.split("/")[-1], d["pipeline_stages"], "ms/block", round(d["ms_per_block"], 4), "Msamples/s", round(d["msamples_per_s"], 1), "offline Msamples/s", round(d["offline"]["msamples_per_s"], 1), "parity", round(d["parity"]["worst_err_over_tol"], 4) if d["parity"] else None)
    except Exception as e:
        print(f, "FAILED", e)
PY
