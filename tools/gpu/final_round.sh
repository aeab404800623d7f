#!/bin/bash
# End-of-round measurement set on one B200 (run from the repo root under gpurun): GPU tests, smoke, headline bench + reference arm,
# voice-count sweep, other configs.  Outputs land in gpurun_out/ and are copied into profiles/ by hand.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -2 gpurun_out/final_bench_n1.err; cut -c1-400 gpurun_out/final_bench_n1.json
python bench.py --impl reference --steps 100 --warmup 10 > gpurun_out/final_bench_ref.json 2>&1; cut -c1-300 gpurun_out/final_bench_ref.json
for v in 8192 16384 32768 65536 131072 262144; do
  python bench.py --steps 30 --warmup 5 --voices $v --no-cpu-baseline > gpurun_out/final_v$v.json 2>/dev/null
done
python - <<'PY'
import json
print("# voices  L  ms/step  K1_ms  Msamples/s  realtime_x  hbm_frac")
for v in (4096, 8192, 16384, 32768, 65536, 131072, 262144):
    f = "gpurun_out/final_bench_n1.json" if v == 4096 else f"gpurun_out/final_v{v}.json"
    d = json.load(open(f))
    print(v, d["engine"]["tile_width"], round(d["ms_per_step"], 4), round(d["roofline"]["kernel_ms"], 4), round(d["value"], 1), round(d["realtime_factor"], 2), round(d["roofline"]["frac"], 5))
PY
python bench_configs.py 3 4 5 > gpurun_out/final_configs.jsonl 2>/dev/null; cut -c1-260 gpurun_out/final_configs.jsonl
