#!/bin/bash
for name in base tm_s3c2 tm_s4c2 tm_s6c1; do
  lib=elementary_b200/libelem_b200_$name.so; [ $name = base ] && lib=elementary_b200/libelem_b200.so
  ELEM_B200_LIB=$PWD/$lib timeout 300 python bench_configs.py 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$name: ms/block', round(d['ms_per_block'],4), 'k1', round(d['k1_ms'],4), 'k3', round(d['k3_ms'],4), 'frac', round(d['roofline']['frac'],3))"
done
