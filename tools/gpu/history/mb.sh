#!/bin/bash
for name in base mb6 mb5; do
  lib=elementary_b200/libelem_b200_$name.so; [ $name = base ] && lib=elementary_b200/libelem_b200.so
  for v in 32768 131072 262144; do
    ELEM_B200_LIB=$PWD/$lib python bench.py --steps 30 --warmup 5 --voices $v --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', $v, 'L', d['config']['tile_width'], 'ms/step', round(d['ms_per_step'],4), 'k1', round(d['roofline']['kernel_ms'],4))"
  done
done
