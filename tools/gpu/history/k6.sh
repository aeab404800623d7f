#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for i in 1 2; do timeout 300 python bench_configs.py 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('config4 ms/block', round(d['ms_per_block'],4), 'k1', round(d['k1_ms'],4), 'k3', round(d['k3_ms'],4), 'launches', d['launches_per_block'], 'rt x', round(d['realtime_x'],1))"; done
python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('4096: ms/step', round(d['ms_per_step'],4), 'k1', round(d['roofline']['kernel_ms'],4), 'value', round(d['value']))"
