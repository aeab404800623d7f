#!/bin/bash
# Round 2, N = 8 validation (run with gpurun --gpus 8): K4 bit-exact against NCCL-gathered partials at world 8, then the headline
# line at 8 GPUs (incl. the 8 x 131072 = 1 M voice leg) with the fused exchange.
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/gpu/peer_test.py 2>&1 | grep -v "^W\|^\*\*\*\|^$" | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r02_n${N}_fused.json 2> gpurun_out/r02_n${N}_fused.err || tail -5 gpurun_out/r02_n${N}_fused.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_n${N}_fused.json"))
t1 = d.get("t1_million_voices") or {}
print("N=$N Msamples/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "kernel_ms", d["kernel_ms"], "parity", d.get("parity_ok"), "| T1", t1.get("voices_total"), "voices", round(t1.get("ms_per_block", 0), 4), "ms/block", t1.get("parity_ok"))
PY
