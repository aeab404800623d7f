"""CPU, world_size 2 over gloo: voice sharding + the mix-bus reduce that bench.py performs over NCCL on GPUs.
Each rank renders its shard with the CPU oracle (the checker stands in for the per-rank engine here: no GPU in this
container), all-reduces its partial mix, and must obtain the mix of the whole voice set."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from elementary_b200 import graphs
from elementary_b200.distributed import reduce_mix, shard_voices
from oracle import oracle as orc

SR, BS, TOTAL, BLOCKS = 48000.0, 512, 11, 3


def test_shard_voices_partitions_exactly():
    for total in (1, 7, 8, 4096, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_voices(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (a, ca), (b, _) in zip(spans, spans[1:]):
                assert a + ca == b
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _partial_mix(first, count):
    acc = np.zeros((1, BLOCKS * BS), dtype=np.float64)
    for v in range(first, first + count):
        r = orc.PortRuntime(SR, BS)
        assert r.apply(graphs.subsynth32()) == 0 and r.apply(graphs.subsynth32_voice_props(v)) == 0
        acc += r.render(BLOCKS, 1)
    return acc


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_voices(TOTAL, rank, world)
    mix = torch.from_numpy(_partial_mix(first, count))
    reduce_mix(mix)                                   # all_reduce: every rank ends with the full mix
    np.save(os.path.join(out_dir, f"mix{rank}.npy"), mix.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_gloo_mix_reduce(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    want = _partial_mix(0, TOTAL)
    for r in range(2):
        got = np.load(tmp_path / f"mix{r}.npy")
        assert np.allclose(got, want, rtol=0, atol=1e-9 * np.abs(want).max())
