#!/bin/bash
# Round 2, run B: the WHOLE GPU suite on the specialised kernels (strict), the new bench line (parity self-check, T1 line,
# reference arm), launch list + ncu --set full of the specialised K1 at 4096 and 131072 voices.
mkdir -p gpurun_out
t0=$(date +%s)
ELEM_B200_SPECIALIZE=1 timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02b_pytest_spec_full.txt
echo "spec suite seconds: $(( $(date +%s) - t0 ))"
python bench.py --steps 100 --warmup 10 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -3 gpurun_out/r02b_bench.err; cut -c1-1500 gpurun_out/r02b_bench.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err; cut -c1-600 gpurun_out/r02b_bench_ref.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02b_bench_ref2.json 2> /dev/null; cut -c1-200 gpurun_out/r02b_bench_ref2.json
# launch list of the bench command (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-t1 > /dev/null 2>&1
for v in 4096 131072; do
  ncu --set full --clock-control none --import-source on -k regex:render_block -s 6 -c 1 -o gpurun_out/r02b_k1_spec_v$v -f python bench.py --steps 4 --warmup 3 --voices $v --no-cpu-baseline --no-parity --no-t1 > gpurun_out/r02b_ncu_v$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
