#!/bin/bash
# K3 check: convolver GPU tests under a hang guard, then config 4 numbers, then a memcheck/racecheck slice
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_convolve_gpu.py tests/test_golden.py -m gpu -x -q -k "convol" 2>&1 | tail -5 | tee gpurun_out/k3_pytest.txt
timeout 300 python bench_configs.py 4 > gpurun_out/k3_config4.jsonl 2>gpurun_out/k3_config4.err; cut -c1-700 gpurun_out/k3_config4.jsonl; tail -2 gpurun_out/k3_config4.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_convolve_gpu.py -m gpu -x -q -k "16384 or larger_graph or varying or other_block" > gpurun_out/k3_memcheck.log 2>&1; echo "memcheck K3 rc=$?"; tail -2 gpurun_out/k3_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_convolve_gpu.py -m gpu -x -q -k "ir_lengths or larger_graph" > gpurun_out/k3_racecheck.log 2>&1; echo "racecheck K3 rc=$?"; tail -2 gpurun_out/k3_racecheck.log
