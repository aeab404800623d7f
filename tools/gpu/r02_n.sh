#!/bin/bash
# Round 2, run N: (1) K3 radix-4 vs radix-2 transform (config 4), (2) 128-sample one-voice tiles with out-of-line heavy math (config 5)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_convolve_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "convolve or pipelined or delay or fuzz or many_voice" 2>&1 | tail -4 | cut -c1-400 | tee gpurun_out/r02n_pytest.txt
for v in product convr2; do
  lib=$PWD/elementary_b200/libelem_b200.so; [ $v = convr2 ] && lib=$PWD/elementary_b200/libelem_b200_convr2.so
  ELEM_B200_LIB=$lib python bench_configs.py 4 > gpurun_out/r02n_config4_$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r02n_config4_$v.json')); print('$v', 'ms/block', round(d['ms_per_block'],4), 'k3_ms', round(d['k3_ms'],5), 'frac', round(d['roofline']['frac'],4))"
done
ELEM_B200_LIB=$PWD/elementary_b200/libelem_b200_prof.so python tools/opprof.py 1250 0 4 | tee gpurun_out/opprof_n_s0_n4.txt | head -12
for cfg in "0 0" "0 4" "3 4"; do set -- $cfg
  timeout 600 python bench_configs.py 5 --stages $1 --niter $2 > gpurun_out/r02n_config5_s$1_n$2.json 2> gpurun_out/r02n_config5_s$1_n$2.err || tail -3 gpurun_out/r02n_config5_s$1_n$2.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02n_config5_s*_n*.json")):
    try:
        for line in open(f):
            d = json.loads(line)
            if d["config"].startswith("5"):
                print(f.split("/")[-1], d["pipeline_stages"], "ms/block", round(d["ms_per_block"], 4), "Msamples/s", round(d["msamples_per_s"], 1), "offline Msamples/s", round(d["offline"]["msamples_per_s"], 1), "parity", round(d["parity"]["worst_err_over_tol"], 4) if d["parity"] else None)
    except Exception as e:
        print(f, "FAILED", e)
PY
