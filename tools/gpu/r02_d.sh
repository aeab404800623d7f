#!/bin/bash
# Round 2, run D: full GPU suite with the new tests (custom nodes, N1/N2, offline/steady path), configs 3/4/5, bench line.
mkdir -p gpurun_out
sha256sum elementary_b200/libelem_b200.so elementary_b200/runtime.py bench.py | cut -c1-16,65- 
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02d_pytest.txt
python bench_configs.py 5 > gpurun_out/r02d_config5.json 2> gpurun_out/r02d_config5.err; tail -2 gpurun_out/r02d_config5.err; cut -c1-1200 gpurun_out/r02d_config5.json
python bench_configs.py 3 4 > gpurun_out/r02d_config34.json 2> gpurun_out/r02d_config34.err; tail -2 gpurun_out/r02d_config34.err; cut -c1-700 gpurun_out/r02d_config34.json
python bench.py --steps 100 --warmup 10 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -2 gpurun_out/r02d_bench.err; cut -c1-300 gpurun_out/r02d_bench.json
