#!/bin/bash
# ncu captures of the final kernels (one B200): K1 at 4096 voices (L = 2) and 131072 voices (L = 32), K3 (1024 channels), plus the
# launch list of the headline command.  Reports land in gpurun_out/; summaries are written into profiles/ by tools/ncu_summary.py.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:render_block -s 5 -c 1 -f -o gpurun_out/prof_k1_v4096 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_block -s 5 -c 1 -f -o gpurun_out/prof_k1_v131072 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --voices 131072 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:convolve_chunk -s 45 -c 1 -f -o gpurun_out/prof_k3 python bench_configs.py 4 --quick > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none -k regex:mix_reduce -s 5 -c 1 -f -o gpurun_out/prof_k2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
