#!/bin/bash
# Round 2, run Z (final after the mm1p change): whole GPU suite + specialised slice, config 5, ncu captures of the final kernels, default bench line
mkdir -p gpurun_out
sha256sum elementary_b200/libelem_b200.so | cut -c1-16
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | cut -c1-400 | tee gpurun_out/r02z_pytest.txt
ELEM_B200_SPECIALIZE=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "mm1p or fuzz or subsynth" 2>&1 | tail -2 | cut -c1-300 | tee gpurun_out/r02z_pytest_spec.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for st in 4 0; do python bench_configs.py 5 --stages $st > gpurun_out/r02z_config5_s$st.json 2>/dev/null; python -c "
import json
for line in open('gpurun_out/r02z_config5_s$st.json'):
    d=json.loads(line)
    if d['config'].startswith('5'): print('config5 stages', d['pipeline_stages'], 'ms/block', round(d['ms_per_block'],4), 'Msamples/s', round(d['msamples_per_s'],1), 'offline', round(d['offline']['msamples_per_s'],1), 'parity', round(d['parity']['worst_err_over_tol'],4))"; done
bash tools/gpu/profile_all.sh r02z > gpurun_out/r02z_profile.log 2>&1; tail -7 gpurun_out/r02z_profile.log
python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; tail -2 gpurun_out/r02z_bench.err; cut -c1-300 gpurun_out/r02z_bench.json
