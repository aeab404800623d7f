#!/bin/bash
# Quick GPU sanity set (run from the repo root under gpurun): GPU parity tests, smoke, headline bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/check_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 100 --warmup 10 > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; tail -2 gpurun_out/check_bench.err; cut -c1-600 gpurun_out/check_bench.json
