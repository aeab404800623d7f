#!/bin/bash
# ncu captures of the current kernels (one B200): K1 at 4096 voices (L = 2) and 131072 voices (L = 32), both specialised (the bench default),
# K2, K3 (1024 channels), render_groups_kernel (config 5, 1250 graphs), plus the launch list of the headline command.
# Reports land in gpurun_out/; summaries are written into profiles/ by tools/ncu_summary.py, profiles/ncu_traffic.json by tools/ncu_to_json.py.
mkdir -p gpurun_out
P=${1:-prof}
ncu --set full --clock-control none --import-source on -k regex:render_block -s 5 -c 1 -f -o gpurun_out/${P}_k1_v4096 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-t1 --no-parity --no-configs > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_block -s 5 -c 1 -f -o gpurun_out/${P}_k1_v131072 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-t1 --no-parity --no-configs --voices 131072 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:convolve_chunk -s 45 -c 1 -f -o gpurun_out/${P}_k3 python bench_configs.py 4 --quick > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none -k regex:mix_reduce -s 5 -c 1 -f -o gpurun_out/${P}_k2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-t1 --no-parity --no-configs > gpurun_out/ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_groups -s 8 -c 1 -f -o gpurun_out/${P}_groups python bench_configs.py 5 --quick --graphs 1250 > gpurun_out/ncu5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 40 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-t1 --no-parity --no-configs > gpurun_out/ncu_launch.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/${P}_launches.csv
