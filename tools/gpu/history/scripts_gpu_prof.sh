#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:render_block -s 5 -c 1 -o gpurun_out/prof_v4096_tw2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tile-width 2 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_block -s 5 -c 1 -o gpurun_out/prof_v131072 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --voices 131072 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
