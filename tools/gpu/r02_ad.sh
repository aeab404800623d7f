#!/bin/bash
# Round 2, run AD (2 GPUs): bench with the per-kernel event pairs out of the timed value loop: N = 1 default line, then N = 2 fused
mkdir -p gpurun_out
python bench.py --no-configs > gpurun_out/r02ad_bench_n1.json 2> gpurun_out/r02ad_bench_n1.err; tail -2 gpurun_out/r02ad_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r02ad_bench_n2.json 2> gpurun_out/r02ad_bench_n2.err || tail -5 gpurun_out/r02ad_bench_n2.err
python - <<'PY'
import json
for f in ("r02ad_bench_n1", "r02ad_bench_n2"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 4), {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["kernel_ms"].items() if k != "note"}, "parity", d.get("parity_ok"), "traffic", d["roofline"]["traffic"], "t1", round(d["t1_million_voices"]["ms_per_block"], 4))
PY
