#!/bin/bash
# Round 2, run X: K3 ring depth 5 (2 CTAs per SM still fit: 104 KB each) vs 4, config 4
mkdir -p gpurun_out
for v in s4 s5 s4 s5; do
  lib=$PWD/elementary_b200/libelem_b200.so; [ $v = s5 ] && lib=$PWD/elementary_b200/libelem_b200_convs5.so
  ELEM_B200_LIB=$lib timeout 300 python -m pytest tests/test_convolve_gpu.py -m gpu -x -q -k "16384 or epilogue" 2>&1 | tail -1
  ELEM_B200_LIB=$lib python bench_configs.py 4 > gpurun_out/r02x_config4_$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r02x_config4_$v.json')); print('$v', 'ms/block', round(d['ms_per_block'],4), 'k3_ms', round(d['k3_ms'],5), 'frac', round(d['roofline']['frac'],4))" | tee -a gpurun_out/r02x_ab.txt
done
