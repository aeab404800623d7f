#!/bin/bash
# compute-sanitizer passes over a representative slice of the GPU parity tests (memcheck: out-of-bounds / misaligned;
# racecheck: shared-memory hazards between the lanes/warps of the interpreter and of the FFT convolver).
mkdir -p gpurun_out
SEL="subsynth32_voices or delay_variants or taps_feedback or svf or heterogeneous or split_live or additive"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_memcheck_k1.log 2>&1; echo "memcheck K1 rc=$?"; tail -3 gpurun_out/sanitize_memcheck_k1.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_convolve_gpu.py -m gpu -x -q -k "16384 or larger_graph or varying or other_block" > gpurun_out/sanitize_memcheck_k3.log 2>&1; echo "memcheck K3 rc=$?"; tail -3 gpurun_out/sanitize_memcheck_k3.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "subsynth32_voices or delay_variants or taps_feedback" > gpurun_out/sanitize_racecheck_k1.log 2>&1; echo "racecheck K1 rc=$?"; tail -3 gpurun_out/sanitize_racecheck_k1.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_convolve_gpu.py -m gpu -x -q -k "ir_lengths or larger_graph" > gpurun_out/sanitize_racecheck_k3.log 2>&1; echo "racecheck K3 rc=$?"; tail -3 gpurun_out/sanitize_racecheck_k3.log
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitize_*.log
