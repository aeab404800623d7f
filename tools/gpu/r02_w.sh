#!/bin/bash
# Round 2, run W: K3 grid = whole rounds (256 CTAs x 2 channel pairs instead of 296 CTAs with 1 or 2) — parity, then config 4 A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_convolve_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/r02w_pytest.txt
for v in balanced unbalanced balanced unbalanced; do
  lib=$PWD/elementary_b200/libelem_b200.so; [ $v = unbalanced ] && lib=$PWD/elementary_b200/libelem_b200_convunbal.so
  ELEM_B200_LIB=$lib python bench_configs.py 4 > gpurun_out/r02w_config4_$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r02w_config4_$v.json')); print('$v', 'ms/block', round(d['ms_per_block'],4), 'k3_ms', round(d['k3_ms'],5), 'frac', round(d['roofline']['frac'],4))" | tee -a gpurun_out/r02w_ab.txt
done
