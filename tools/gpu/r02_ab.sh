#!/bin/bash
# Round 2, run AB (final): whole GPU suite on the final build (+ specialised slice), smoke, ncu captures, voice sweep with the automatic
# dense mid-range geometry, default bench line and reference arm
mkdir -p gpurun_out
sha256sum elementary_b200/libelem_b200.so | cut -c1-16
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | cut -c1-400 | tee gpurun_out/r02ab_pytest.txt
ELEM_B200_SPECIALIZE=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "subsynth or wide_tiles or fuzz or mm1p" 2>&1 | tail -2 | cut -c1-300 | tee gpurun_out/r02ab_pytest_spec.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/gpu/profile_all.sh r02ab > gpurun_out/r02ab_profile.log 2>&1; tail -7 gpurun_out/r02ab_profile.log
for v in 8192 16384 32768 65536 262144; do
  python bench.py --steps 30 --warmup 5 --voices $v --no-cpu-baseline --no-t1 --no-configs > gpurun_out/r02ab_v$v.json 2>/dev/null
done
python bench.py > gpurun_out/r02ab_bench.json 2> gpurun_out/r02ab_bench.err; tail -2 gpurun_out/r02ab_bench.err
python - <<'PY'
import json
print("# voices  L  ms/step  K1_ms  Msamples/s  realtime_x  hbm_frac  regs  parity")
for v in (4096, 8192, 16384, 32768, 65536, 131072, 262144):
    try:
        if v == 131072:
            d = json.load(open("gpurun_out/r02ab_bench.json"))["t1_million_voices"]
            print(v, d["tile_width"], round(d["ms_per_block"], 4), round(d["k1_ms"], 4), round(d["value"], 1), round(d["realtime_factor"], 2), round(d["hbm_frac"], 5), "-", d["parity_ok"]); continue
        f = "gpurun_out/r02ab_bench.json" if v == 4096 else f"gpurun_out/r02ab_v{v}.json"
        d = json.load(open(f))
        print(v, d["engine"]["tile_width"], round(d["ms_per_step"], 4), round(d["roofline"]["kernel_ms"], 4), round(d["value"], 1), round(d["realtime_factor"], 2), round(d["roofline"]["frac"], 5), d["engine"].get("spec", {}).get("spec_regs"), d.get("parity_ok"))
    except Exception as e:
        print(v, "FAILED", e)
d = json.load(open("gpurun_out/r02ab_bench.json")); print("value", d["value"], "e2e", d["e2e"]["value"], {k: round(v["ms_per_block"], 4) for k, v in d["other_configs"].items()})
PY
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02ab_bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/r02ab_bench_ref.json
