"""Multi-GPU check of K4 (run under torchrun, one rank per GPU): the fused peer-memory all-reduce of the mix bus must equal
the rank-ordered float sum of the per-rank partial mixes BIT FOR BIT (gathered with NCCL from a second, identical runtime
that does not exchange), on every rank, for every block."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from elementary_b200 import Runtime, graphs                              # noqa: E402
from elementary_b200.distributed import attach_peer_mix                  # noqa: E402
from elementary_b200.runtime import FLAG_MIX, FLAG_ALLREDUCE             # noqa: E402

SR, BS, V = 48000.0, 512, 96


def build(rank, local_rank):
    rt = Runtime(SR, BS, V, device=local_rank)
    assert rt.apply_instructions(graphs.subsynth32()) == 0
    ida, idb = graphs.subsynth32_param_ids()
    f0 = np.array([graphs.subsynth32_f0(rank * V + v) for v in range(V)], dtype=np.float64)
    assert rt.set_property_per_voice(ida, "value", f0) == 0 and rt.set_property_per_voice(idb, "value", f0 * 1.007) == 0
    return rt


def main():
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    a, b = build(rank, local_rank), build(rank, local_rank)
    attach_peer_mix(a)
    mix_a = torch.as_tensor(a.mix_device(2), device=f"cuda:{local_rank}")
    mix_b = torch.as_tensor(b.mix_device(2), device=f"cuda:{local_rank}")
    bad = 0
    for blk in range(12):
        a.enqueue_block(0, 2, BS, FLAG_MIX | FLAG_ALLREDUCE)
        b.enqueue_block(0, 2, BS, FLAG_MIX)
        a.synchronize(); b.synchronize()
        parts = [torch.empty_like(mix_b) for _ in range(world)]
        dist.all_gather(parts, mix_b.clone())
        want = torch.zeros_like(mix_b)
        for p in parts:                       # rank order, float32, like K4
            want = want + p
        if not torch.equal(mix_a, want):
            bad += 1
            print(f"rank {rank} block {blk}: max diff {(mix_a - want).abs().max().item()}")
    # Runtime::process with peers attached: every rank's host buffers receive the mix of ALL ranks, written into mapped host memory
    # by K4 itself (kernels.h HostDeliver) — must be the same bits as the rank-ordered sum of the partial mixes
    for blk in range(6):
        host = a.process(None, 2, BS)
        b.enqueue_block(0, 2, BS, FLAG_MIX)
        b.synchronize()
        parts = [torch.empty_like(mix_b) for _ in range(world)]
        dist.all_gather(parts, mix_b.clone())
        want = torch.zeros_like(mix_b)
        for p in parts:
            want = want + p
        if not np.array_equal(host, want[:2].cpu().numpy()):
            bad += 1
            print(f"rank {rank} process() block {blk}: max diff {np.abs(host - want[:2].cpu().numpy()).max()}")
    st = a.peer_status()
    t = torch.tensor([bad, st], device=f"cuda:{local_rank}")
    dist.all_reduce(t)
    if rank == 0:
        print(f"PEER_MIX_TEST world={world} mismatching_blocks={int(t[0])} peer_timeouts={int(t[1])} peak={float(mix_a.abs().max()):.4f}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(t[0]) == 0 and int(t[1]) == 0 else 1)


if __name__ == "__main__":
    main()
