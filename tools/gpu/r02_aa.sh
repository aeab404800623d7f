#!/bin/bash
# Round 2, run AA: mid-range voice counts — specialised narrow geometries compiled for 8 CTAs per SM (64 registers) with twice the warps
# (target_tiles 4096) against the defaults (4 CTAs per SM, ~2048 warps)
mkdir -p gpurun_out
for v in 8192 16384 32768 65536; do
  for cfg in "default" "mb8_t4096" "mb8_t2048"; do
    extra=""
    [ $cfg = mb8_t4096 ] && extra="--opt spec_minblocks=8 --opt target_tiles=4096"
    [ $cfg = mb8_t2048 ] && extra="--opt spec_minblocks=8"
    python bench.py --steps 30 --warmup 5 --voices $v --no-cpu-baseline --no-t1 --no-configs $extra > gpurun_out/r02aa_v${v}_$cfg.json 2>/dev/null
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02aa_v*_*.json"), key=lambda f: (int(f.split("_v")[1].split("_")[0]), f)):
    try:
        d = json.load(open(f))
        sp = d["engine"].get("spec", {})
        print(f.split("/")[-1], "L", d["engine"]["tile_width"], "ms/step", round(d["ms_per_step"], 4), "K1", round(d["roofline"]["kernel_ms"], 4), "regs", sp.get("spec_regs"), "local", sp.get("spec_local_bytes"), "parity", d.get("parity_ok"))
    except Exception as e:
        print(f, "FAILED", e)
PY
