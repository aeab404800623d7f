#!/bin/bash
# Round 2, run L: broadcast-read recurrence loops with register outputs, software-pipelined interpreter loop (next header + operand
# words prefetched), chain step prefetch: parity (both kernels), opcode profile, config 5, headline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | cut -c1-400 | tee gpurun_out/r02l_pytest.txt
ELEM_B200_SPECIALIZE=1 timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "pole or biquad or mm1p or rand or fuzz or subsynth or env or delay" 2>&1 | tail -5 | cut -c1-400 | tee gpurun_out/r02l_pytest_spec.txt
for cfg in "0 0" "3 0"; do set -- $cfg; ELEM_B200_LIB=$PWD/elementary_b200/libelem_b200_prof.so python tools/opprof.py 1250 $1 $2 | tee gpurun_out/opprof_s$1_n$2.txt | head -14; done
for cfg in "0 0" "2 0" "3 0" "4 0"; do set -- $cfg
  timeout 600 python bench_configs.py 5 --stages $1 --niter $2 > gpurun_out/r02l_config5_s$1_n$2.json 2> gpurun_out/r02l_config5_s$1_n$2.err || tail -3 gpurun_out/r02l_config5_s$1_n$2.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02l_config5_s*_n*.json")):
    try:
        for line in open(f):
            d = json.loads(line)
            if d["config"].startswith("5"):
                print(f.split("/")[-1], d["pipeline_stages"], "ms/block", round(d["ms_per_block"], 4), "Msamples/s", round(d["msamples_per_s"], 1), "offline Msamples/s", round(d["offline"]["msamples_per_s"], 1), "parity", d["parity"]["worst_err_over_tol"] if d["parity"] else None)
    except Exception as e:
        print(f, "FAILED", e)
PY
python bench_configs.py 3 > gpurun_out/r02l_config3.json 2>/dev/null; cut -c1-200 gpurun_out/r02l_config3.json
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-configs --specialize 0 --no-t1 > gpurun_out/r02l_bench_interp.json 2>/dev/null; cut -c1-250 gpurun_out/r02l_bench_interp.json
