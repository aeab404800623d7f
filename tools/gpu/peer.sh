#!/bin/bash
# K4 on N GPUs (N = number of visible devices): bit-exact check against NCCL-gathered partials, then bench fused vs nccl
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/gpu/peer_test.py 2>&1 | grep -v "^W\|^\*\*\*\|^$" | tail -6
for c in fused nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 200 --warmup 20 --collective $c > gpurun_out/peer_n${N}_$c.json 2> gpurun_out/peer_n${N}_$c.err || tail -5 gpurun_out/peer_n${N}_$c.err
  python - <<PY
import json
d = json.load(open("gpurun_out/peer_n${N}_$c.json"))
print("N=$N $c: Msamples/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "|", d["engine"]["collective"][:70])
PY
done
