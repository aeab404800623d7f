#!/bin/bash
# Last check of a commit on one B200: the GPU test-suite, smoke, the headline bench line (no CPU baseline leg)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/sanity_bench.json 2> gpurun_out/sanity_bench.err; echo "stdout lines: $(wc -l < gpurun_out/sanity_bench.json)"; python -c "
import json; d=json.load(open('gpurun_out/sanity_bench.json')); print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])"
