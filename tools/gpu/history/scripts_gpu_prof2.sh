#!/bin/bash
mkdir -p gpurun_out
for tw in 1 2 4 8; do
  python bench.py --steps 50 --warmup 5 --no-cpu-baseline --tile-width $tw | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('4096 voices L', d['config']['tile_width'], 'Msamples/s', round(d['value']), 'k1 ms', round(d['roofline']['kernel_ms'],4))"
done
for v in 8192 16384 32768 65536; do
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --voices $v | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['voices_total'], 'L', d['config']['tile_width'], 'Msamples/s', round(d['value']), 'k1 ms', round(d['roofline']['kernel_ms'],4), 'rt x', round(d['realtime_factor'],1))"
done
ncu --set full --clock-control none --import-source on -k regex:render_block -s 5 -c 1 -o gpurun_out/prof_v4096 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_block -s 5 -c 1 -o gpurun_out/prof_v131072 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --voices 131072 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:convolve_chunk -s 45 -c 1 -o gpurun_out/prof_k3 python bench_configs.py 4 --quick > gpurun_out/ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ls gpurun_out/*.ncu-rep
