"""Voice sharding and the one collective of the path: the mix-bus reduce.

Voices / graph instances share no state, so rank ``r`` of ``W`` simply owns a contiguous voice range and renders it with its
own :class:`elementary_b200.Runtime` on its own GPU; shared read-only resources (IRs, wavetables) are added on every rank.
The only exchange is the element-wise sum of the per-rank partial mix buses ``[n_out][block]`` (4 KB per block for stereo
at 512 samples — pure latency over NVLink/NVSwitch), done in place by ``torch.distributed.all_reduce`` (NCCL on GPUs, gloo
in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def shard_voices(total_voices: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous balanced partition: returns (first_voice, count) of `rank`; counts differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(int(total_voices), int(world_size))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def reduce_mix(mix, group=None, dst: int | None = None):
    """Sum the partial mix buses of all ranks in place. ``mix`` is a torch tensor (a zero-copy view of
    ``Runtime.mix_device()`` on GPU).  ``dst=None`` -> all_reduce (every rank gets the mix), else reduce to ``dst``."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mix
    if dst is None:
        dist.all_reduce(mix, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(mix, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return mix
