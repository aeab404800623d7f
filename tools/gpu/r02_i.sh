#!/bin/bash
# Round 2, run I: opcode profile of the interpreter on config 5 (cost model calibration), K3 root/mix epilogue parity + config 4.
mkdir -p gpurun_out
bash tools/gpu/opprof.sh 2>&1 | tail -60
timeout 900 python -m pytest tests/test_convolve_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "convolve or offline or heterogeneous" 2>&1 | tail -12 | cut -c1-500 | tee gpurun_out/r02i_pytest.txt
python bench_configs.py 4 > gpurun_out/r02i_config4.json 2> gpurun_out/r02i_config4.err; tail -2 gpurun_out/r02i_config4.err; cut -c1-700 gpurun_out/r02i_config4.json
