#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_convolve_gpu.py tests/test_golden.py -m gpu -x -q -k "convol" 2>&1 | tail -3
for i in 1 2; do timeout 300 python bench_configs.py 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('ms/block', round(d['ms_per_block'],4), 'k1', round(d['k1_ms'],4), 'k3', round(d['k3_ms'],4), 'frac', round(d['roofline']['frac'],3))"; done
