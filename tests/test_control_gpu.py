"""GPU parity for the sequencing / control nodes beyond the static cases of tests/cases.py
(which tests/test_golden.py already runs against the reference's outputs): live property changes between blocks —
new sequence data, loop points, re-arming — the host-supplied sample clock, and every tile geometry.

The reference's nodes take such changes at the top of their next process() through atomics and SPSC queues
(Core.h:470-495, SparSeq.h:209-257, Core.h:345-361); all results here must be bit-exact (no transcendental involved).
"""
import numpy as np
import pytest

from elementary_b200 import Runtime, el
from control_common import scenarios
from helpers import oracle_cls

pytestmark = pytest.mark.gpu
SR, BS = 48000.0, 512



def lockstep(batch, script, n_blocks, n_voices=3, n_out=1, exact=True, sample_times=None, **opts):
    """Run the GPU runtime and one oracle per voice in lock step; script = {block: batch applied before that block}."""
    rt = Runtime(SR, BS, n_voices, device=0, **opts)
    oracles = [oracle_cls()(SR, BS) for _ in range(n_voices)]
    assert rt.apply_instructions(batch) == 0, rt.last_error()
    for o in oracles:
        assert o.apply(batch) == 0
    got, ref = [], []
    for b in range(n_blocks):
        if b in script:
            assert rt.apply_instructions(script[b]) == 0, rt.last_error()
            for o in oracles:
                assert o.apply(script[b]) == 0
        if sample_times is not None:
            rt.set_current_time(sample_times[b])
            for o in oracles:
                o.set_current_time(sample_times[b])
        got.append(rt.process_voices(None, n_out, BS)[0])
        ref.append(np.stack([o.process(None, n_out, BS) for o in oracles]))
    g, r = np.concatenate(got, axis=2), np.concatenate(ref, axis=2)
    if exact:
        assert np.array_equal(g, r), f"max diff {np.abs(g - r).max()} at {np.argwhere(g != r)[:4]}"
    else:
        assert np.abs(g - r).max() <= 1e-5 * np.abs(r).max()
    return g


SCEN = scenarios()


@pytest.mark.parametrize("sc", SCEN, ids=[s["name"] for s in SCEN])
def test_live_updates_match_oracle(sc):
    g = lockstep(sc["batch"], sc["script"], sc["n_blocks"], n_out=sc["n_out"], exact=sc["exact"], sample_times=sc.get("sample_times"))
    assert np.abs(g).max() > 0


@pytest.mark.parametrize("tile_width", [1, 2, 8, 16, 32])
def test_seq_live_updates_in_every_tile_geometry(tile_width):
    sc = SCEN[0]
    g = lockstep(sc["batch"], sc["script"], sc["n_blocks"], n_voices=5 if tile_width < 16 else 40, tile_width=tile_width)
    assert np.abs(g).max() >= 30.0


def test_once_arm_per_voice_without_splitting_the_group():
    once = el.once({"arm": False, "key": "o"}, el.train(300.0))
    oid = once.id()
    rt = Runtime(SR, BS, 4, device=0)
    oracles = [oracle_cls()(SR, BS) for _ in range(4)]
    batch = el.render(once)
    assert rt.apply_instructions(batch) == 0
    for o in oracles:
        assert o.apply(batch) == 0
    got, ref = [], []
    for b in range(5):
        if b == 2:                                            # arm voice 1 and 3 only
            for v in (1, 3):
                assert rt.apply_instructions([[3, oid, "arm", True]], voices=(v, v + 1)) == 0
                assert oracles[v].apply([[3, oid, "arm", True]]) == 0
            assert len(rt.describe()["groups"]) == 1
        got.append(rt.process_voices(None, 1, BS)[0])
        ref.append(np.stack([o.process(None, 1, BS) for o in oracles]))
    g, r = np.concatenate(got, axis=2), np.concatenate(ref, axis=2)
    assert np.array_equal(g, r)
    assert np.abs(g[1]).max() > 0 and np.abs(g[0]).max() == 0


def test_process_takes_the_sample_clock_from_user_data():
    rt = Runtime(SR, BS, 2, device=0)
    assert rt.apply_instructions(el.render(el.time())) == 0
    for _ in range(3):
        rt.process(None, 1, BS)                                # pass the root fade; engine-kept clock runs 0, 512, 1024
    assert rt.current_time() == 3 * BS
    y = rt.process(None, 1, BS, sample_time=1_000_000)[0]
    assert np.array_equal(y, 2.0 * (1_000_000 + np.arange(BS)).astype(np.float64).astype(np.float32))   # mix of 2 voices
    assert rt.current_time() == 1_000_000 + BS


