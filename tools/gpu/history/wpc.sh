#!/bin/bash
for v in 4096 32768 65536; do
  for w in 1 2 4; do
    python bench.py --steps 50 --warmup 5 --voices $v --no-cpu-baseline --opt warps_per_cta=$w 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print($v, 'wpc', $w, 'L', d['config']['tile_width'], 'ms/step', round(d['ms_per_step'],4), 'k1', round(d['roofline']['kernel_ms'],4))"
  done
done
