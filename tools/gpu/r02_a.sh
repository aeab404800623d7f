#!/bin/bash
# Round 2, run A: full GPU parity suite (interpreter), then the FIRST hardware run of the NVRTC-specialised K1:
# a parity subset with ELEM_B200_SPECIALIZE=1 (strict) and bench lines interpreter vs specialised at 4096 and 131072 voices.
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02a_pytest.txt
echo "suite seconds: $(( $(date +%s) - t0 ))"
t0=$(date +%s)
ELEM_B200_SPECIALIZE=1 timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "subsynth or additive or svf or delay or phasor or fuzz or plumbing or soak or biquad or taps or table or blep" 2>&1 | tail -6 | tee gpurun_out/r02a_pytest_spec.txt
echo "spec subset seconds: $(( $(date +%s) - t0 ))"
for v in 4096 131072; do
  python bench.py --steps 50 --warmup 10 --voices $v --no-cpu-baseline > gpurun_out/r02a_interp_v$v.json 2> gpurun_out/r02a_interp_v$v.err
  python bench.py --steps 50 --warmup 10 --voices $v --no-cpu-baseline --opt specialize=2 --opt specialize_strict=1 > gpurun_out/r02a_spec_v$v.json 2> gpurun_out/r02a_spec_v$v.err
  tail -3 gpurun_out/r02a_spec_v$v.err
done
python - <<'PY'
import json
for v in (4096, 131072):
    for k in ("interp", "spec"):
        try:
            d = json.load(open(f"gpurun_out/r02a_{k}_v{v}.json"))
            print(k, v, "L", d["engine"]["tile_width"], "ms/step", round(d["ms_per_step"], 4), "K1 ms", round(d["roofline"]["kernel_ms"], 4), "Msamples/s", round(d["value"], 1), d["engine"].get("spec"))
        except Exception as e:
            print(k, v, "FAILED", e)
PY
