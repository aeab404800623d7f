#!/bin/bash
# Round 2, run C: svf warp scan (L <= 4) — parity on both kernels, then bench A/B over tile widths.
mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /proc/loadavg; nproc
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02c_pytest.txt
ELEM_B200_SPECIALIZE=1 timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "svf or subsynth or fuzz or split or plumbing or soak" 2>&1 | tail -5 | tee gpurun_out/r02c_pytest_spec.txt
python bench_configs.py 1 > gpurun_out/r02c_config1.json 2>&1; cut -c1-900 gpurun_out/r02c_config1.json
for tw in 1 2 4; do
  python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-t1 --tile-width $tw > gpurun_out/r02c_spec_L$tw.json 2> gpurun_out/r02c_spec_L$tw.err
  python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-t1 --tile-width $tw --specialize 0 > gpurun_out/r02c_interp_L$tw.json 2> gpurun_out/r02c_interp_L$tw.err
done
for wpc in 1 2 8; do
  python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-t1 --tile-width 1 --opt warps_per_cta=$wpc > gpurun_out/r02c_spec_L1_wpc$wpc.json 2>/dev/null
  python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-t1 --tile-width 2 --opt warps_per_cta=$wpc > gpurun_out/r02c_spec_L2_wpc$wpc.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c_*_L*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "L", d["engine"]["tile_width"], "ms/step", round(d["ms_per_step"], 4), "K1", round(d["roofline"]["kernel_ms"], 4), "parity", d.get("worst_err_over_tol"), d["engine"].get("spec", {}).get("spec_regs"))
    except Exception as e:
        print(f, "FAILED", e)
PY
