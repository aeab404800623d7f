#!/usr/bin/env python
"""Per-opcode cycle profile of the K1 interpreter on BASELINE config 5's graphs (needs the -DEB_OPPROF A/B library:
ELEM_B200_LIB=elementary_b200/libelem_b200_prof.so python tools/opprof.py [n_graphs] [pipeline_stages]).  Prints cycles per dispatch
(one 32-sample tile at L = 1) per opcode — the calibration of the pipeline cost model in graph_host.cpp (costOf)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from elementary_b200 import Runtime, graphs

NAMES = ["END", "SEG", "FILL0", "COPY", "LOADIN", "CHAIN", "PHASOR", "SPHASOR", "COUNTER", "ACCUM", "LATCH", "MAXHOLD", "RAND", "POLE", "ENV", "BIQUAD",
         "PREWARP", "MM1P", "SVF", "SVFSHELF", "Z", "DELAY", "SDELAY", "TABLE", "BLEP", "TAPIN", "TAPOUT", "ROOT", "STOREBUF", "LOADBUF", "PROMOTE"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
stages = int(sys.argv[2]) if len(sys.argv) > 2 else 0
niter = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rt = Runtime(48000.0, 512, n, device=0, pipeline_stages=stages, **({"niter": niter} if niter else {}))
for i in range(n):
    assert rt.apply_instructions(graphs.random_graph(i, 64), voices=(i, i + 1)) == 0
from elementary_b200.runtime import FLAG_MIX
for _ in range(4):
    rt.enqueue_block(0, 1, 512, FLAG_MIX)
rt.synchronize()
rt.debug_opprof(True)
B = 8
for _ in range(B):
    rt.enqueue_block(0, 1, 512, FLAG_MIX)
rt.synchronize()
p = rt.debug_opprof(True).astype(np.float64)
tot_ops = p[:63, 0].sum()
whole = p[63, 0]
print(json.dumps({"graphs": n, "pipeline_stages": stages, "niter": niter, "blocks": B, "warp_cycles_total": whole, "cycles_in_op_bodies": tot_ops, "share_in_bodies": tot_ops / max(1.0, whole),
                  "warps": p[63, 1] / B, "cycles_per_warp_block": whole / max(1.0, p[63, 1])}))
rows = []
for op in range(63):
    if p[op, 1] > 0:
        rows.append((p[op, 0], NAMES[op] if op < len(NAMES) else str(op), p[op, 1]))
for c, name, cnt in sorted(rows, reverse=True):
    print(f"{name:10s} share {100 * c / tot_ops:6.2f}%  cycles/dispatch {c / cnt:9.1f}  dispatches/graph-block {cnt / n / B:7.2f}")
