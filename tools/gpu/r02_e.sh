#!/bin/bash
# Round 2, run E: the failing split test with its message, then the rest of the suite past it.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "split_keeps_sequencer" --tb=short 2>&1 | tail -60 | cut -c1-1500 | tee gpurun_out/r02e_split.txt
timeout 1200 python -m pytest tests -m gpu -q --deselect "tests/test_parity_gpu.py::test_split_keeps_sequencer_and_analysis_state" 2>&1 | tail -8 | tee gpurun_out/r02e_pytest_rest.txt
