#!/bin/bash
# Round 2, run T (final): whole GPU suite on the final build, smoke, ncu captures, then the default bench line and the reference arm
mkdir -p gpurun_out
sha256sum elementary_b200/libelem_b200.so | cut -c1-16
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | cut -c1-400 | tee gpurun_out/r02t_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/gpu/profile_all.sh r02t > gpurun_out/r02t_profile.log 2>&1; tail -8 gpurun_out/r02t_profile.log
python bench.py > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; tail -2 gpurun_out/r02t_bench.err; cut -c1-400 gpurun_out/r02t_bench.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02t_bench_ref.json 2> gpurun_out/r02t_bench_ref.err; cut -c1-300 gpurun_out/r02t_bench_ref.json
