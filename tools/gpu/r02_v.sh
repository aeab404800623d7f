#!/bin/bash
# Round 2, run V: compute-sanitizer memcheck over the kernels added this round (column-walk recurrences, pipeline kernel, K3 epilogue,
# host delivery), the voice-count sweep of the specialised K1, and the default bench line with the e2e loop free of per-kernel events
mkdir -p gpurun_out
SEL="pipelined or many_voice or host_delivery or delay_read_head or pole or biquad or mm1p or rand or offline"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/r02v_memcheck_k1.log 2>&1; echo "memcheck K1/K2 rc=$?"; tail -2 gpurun_out/r02v_memcheck_k1.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_convolve_gpu.py -m gpu -x -q -k "epilogue or 16384 or varying" > gpurun_out/r02v_memcheck_k3.log 2>&1; echo "memcheck K3 rc=$?"; tail -2 gpurun_out/r02v_memcheck_k3.log
grep -h "ERROR SUMMARY" gpurun_out/r02v_memcheck_*.log
for v in 8192 16384 32768 65536 262144; do
  python bench.py --steps 30 --warmup 5 --voices $v --no-cpu-baseline --no-t1 --no-configs > gpurun_out/r02v_v$v.json 2>/dev/null
done
python bench.py --no-cpu-baseline --no-configs > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; tail -2 gpurun_out/r02v_bench.err
python - <<'PY'
import json
print("# voices  L  ms/step  K1_ms  Msamples/s  realtime_x  hbm_frac  k1  parity")
for v in (4096, 8192, 16384, 32768, 65536, 131072, 262144):
    try:
        if v == 131072:
            d = json.load(open("gpurun_out/r02v_bench.json"))["t1_million_voices"]
            print(v, d["tile_width"], round(d["ms_per_block"], 4), round(d["k1_ms"], 4), round(d["value"], 1), round(d["realtime_factor"], 2), round(d["hbm_frac"], 5), d["k1"], d["parity_ok"])
            continue
        f = "gpurun_out/r02v_bench.json" if v == 4096 else f"gpurun_out/r02v_v{v}.json"
        d = json.load(open(f))
        print(v, d["engine"]["tile_width"], round(d["ms_per_step"], 4), round(d["roofline"]["kernel_ms"], 4), round(d["value"], 1), round(d["realtime_factor"], 2), round(d["roofline"]["frac"], 5), d["engine"]["k1"][:11], d.get("parity_ok"))
    except Exception as e:
        print(v, "FAILED", e)
d = json.load(open("gpurun_out/r02v_bench.json"))
print("e2e", d["e2e"], "traffic", d["roofline"]["traffic"], "issue", d["roofline"]["issue_slots"])
PY
