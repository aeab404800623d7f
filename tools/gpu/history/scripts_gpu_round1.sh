#!/bin/bash
# GPU session 1: smoke, bench, reference arm, voice-count sweep, ncu launch list + full capture of K1
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; lscpu | grep "Model name" >> gpurun_out/host.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
for v in 16384 65536 131072 262144; do
  python bench.py --steps 30 --warmup 5 --voices $v --no-cpu-baseline > gpurun_out/bench_v$v.json 2>gpurun_out/bench_v$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_v$v.json"))
print($v, "Msamples/s", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "k1 ms", round(d["roofline"]["kernel_ms"],3), "frac", round(d["roofline"]["frac"],4), "L", d["config"]["tile_width"], "rt x", round(d["realtime_factor"],1))
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_block -s 5 -c 2 -o gpurun_out/prof_k1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
