#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_golden.py tests/test_events_gpu.py -m gpu -x -q -k "table or fft or scope_two or capture-0 or meter_every" > gpurun_out/k5_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/k5_memcheck.log
python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('4096: ms/step', round(d['ms_per_step'],4), 'k1', round(d['roofline']['kernel_ms'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'value', round(d['value']))"
