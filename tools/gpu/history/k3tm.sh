#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_convolve_gpu.py tests/test_golden.py tests/test_parity_gpu.py -m gpu -x -q -k "convol or split_live" 2>&1 | tail -4
for v in tm phases; do
  ELEM_B200_CONV_KERNEL=$v timeout 300 python bench_configs.py 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$v: ms/block', round(d['ms_per_block'],4), 'k1', round(d['k1_ms'],4), 'k3', round(d['k3_ms'],4), 'frac', round(d['roofline']['frac'],3))"
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_convolve_gpu.py -m gpu -x -q -k "16384 or larger_graph or varying or other_block" > gpurun_out/k3tm_memcheck.log 2>&1; echo "memcheck K3 rc=$?"; tail -2 gpurun_out/k3tm_memcheck.log
