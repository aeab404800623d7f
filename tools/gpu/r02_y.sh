#!/bin/bash
# Round 2, run Y: K3 producer fills the first ring stages before the CTA barrier (early) vs behind it (late); ring depth 5 in both
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_convolve_gpu.py tests/test_golden.py tests/test_parity_gpu.py -m gpu -x -q -k "convolve or golden or split_live or heterogeneous" 2>&1 | tail -3 | cut -c1-300 | tee gpurun_out/r02y_pytest.txt
for v in early late early late; do
  lib=$PWD/elementary_b200/libelem_b200.so; [ $v = late ] && lib=$PWD/elementary_b200/libelem_b200_convlate.so
  ELEM_B200_LIB=$lib python bench_configs.py 4 > gpurun_out/r02y_config4_$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r02y_config4_$v.json')); print('$v', 'ms/block', round(d['ms_per_block'],4), 'k3_ms', round(d['k3_ms'],5), 'frac', round(d['roofline']['frac'],4))" | tee -a gpurun_out/r02y_ab.txt
done
