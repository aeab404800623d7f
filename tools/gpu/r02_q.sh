#!/bin/bash
# Round 2, run Q: headline kernel geometry A/B on the specialised K1 (4096 voices): L = 2 T = 64 (default), L = 2 T = 128, L = 1 T = 128, L = 1 T = 32
mkdir -p gpurun_out
for cfg in "0 0" "2 8" "1 4" "1 0"; do set -- $cfg
  extra=""; [ $1 != 0 ] && extra="--tile-width $1"; [ $2 != 0 ] && extra="$extra --opt niter=$2"
  python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-t1 --no-configs $extra > gpurun_out/r02q_L$1_n$2.json 2> gpurun_out/r02q_L$1_n$2.err || tail -2 gpurun_out/r02q_L$1_n$2.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02q_L*_n*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "L", d["engine"]["tile_width"], "ms/step", round(d["ms_per_step"], 4), "K1", round(d["roofline"]["kernel_ms"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4), "parity", d.get("parity_ok"), round(d.get("worst_err_over_tol", -1), 4), d["engine"].get("spec", {}).get("spec_regs"), d["engine"].get("spec", {}).get("spec_local_bytes"))
    except Exception as e:
        print(f, "FAILED", e)
PY
