"""GPU parity for the sequencing / control nodes beyond the static cases of tests/cases.py
(which tests/test_golden.py already runs against the reference's outputs): live property changes between blocks —
new sequence data, loop points, re-arming — the host-supplied sample clock, and every tile geometry.

The reference's nodes take such changes at the top of their next process() through atomics and SPSC queues
(Core.h:470-495, SparSeq.h:209-257, Core.h:345-361); all results here must be bit-exact (no transcendental involved).
"""
import numpy as np
import pytest

from elementary_b200 import Runtime, el
from helpers import oracle_cls

pytestmark = pytest.mark.gpu
SR, BS = 48000.0, 512

TRIG = el.train(2000.0)
RST = el.train(170.0)


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


@pytest.mark.parametrize("tile_width", [0, 1, 8, 32])
def test_seq_new_data_hold_and_offset_while_running(tile_width):
    node = el.seq({"seq": [1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0], "hold": True, "key": "s"}, TRIG, RST)
    nid = node.id()
    script = {
        2: [[3, nid, "seq", [10.0, 20.0, 30.0]]],            # shorter: seqIndex wraps by modulo, holdValue resampled (Core.h:476-494)
        4: [[3, nid, "hold", False], [3, nid, "offset", 1]],
        6: [[3, nid, "loop", False], [3, nid, "seq", [-1.0, -2.0, -3.0, -4.0, -5.0]]],
    }
    opts = {"tile_width": tile_width} if tile_width else {}
    g = lockstep(el.render(node), script, 9, n_voices=5 if tile_width != 32 else 40, **opts)
    assert np.abs(g).max() >= 30.0


def test_seq2_and_once_rearm():
    seq2 = el.seq2({"seq": [0.5, 1.5, -2.5], "key": "s2"}, TRIG, RST)
    once = el.once({"arm": False, "key": "o"}, el.train(300.0))
    sid, oid = seq2.id(), once.id()
    script = {
        1: [[3, oid, "arm", True]],
        3: [[3, sid, "seq", [9.0, 8.0, 7.0, 6.0]], [3, sid, "offset", 2]],
        4: [[3, oid, "arm", False], [3, oid, "arm", True]],   # re-arm (Core.h:347-358: the prop can never disarm)
        6: [[3, sid, "hold", True], [3, sid, "loop", False]],
    }
    lockstep(el.render(seq2, once), script, 8, n_out=2)


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


@pytest.mark.parametrize("follow", [False, True])
def test_sparseq_loop_points_and_sequence_changes(follow):
    sp = [{"value": 1.0, "tickTime": 0}, {"value": 4.0, "tickTime": 3}, {"value": -2.0, "tickTime": 4}, {"value": 9.0, "tickTime": 11}]
    node = el.sparseq({"seq": sp, "follow": follow, "interpolate": 1, "tickInterval": 0.0005, "key": "sp"}, TRIG, RST)
    nid = node.id()
    sp_b = [{"value": 2.0, "tickTime": 0}, {"value": 6.0, "tickTime": 2}, {"value": 0.5, "tickTime": 9}, {"value": 3.0, "tickTime": 20}]
    script = {
        1: [[3, nid, "loop", [0, 8]]],
        3: [[3, nid, "loop", [2, 6]]],                         # follow: promoted at the end of the running loop (SparSeq.h:157-177)
        4: [[3, nid, "seq", sp_b]],
        6: [[3, nid, "loop", False], [3, nid, "offset", 3]],
        7: [[3, nid, "loop", None], [3, nid, "interpolate", 0]],
    }
    g = lockstep(el.render(node), script, 10, exact=False)
    assert np.abs(g).max() > 1.0


def test_sparseq2_new_sequence_and_interpolation_switch():
    sp2 = [{"value": 0.5, "time": 0.002}, {"value": 2.0, "time": 0.004}, {"value": -1.0, "time": 0.011}, {"value": 3.0, "time": 0.02}]
    t = el.mul(0.03, el.abs_(el.cycle(37.0)))                  # time runs forwards and backwards
    node = el.sparseq2({"seq": sp2, "key": "q"}, t)
    nid = node.id()
    script = {
        2: [[3, nid, "interpolate", 1]],
        4: [[3, nid, "seq", [{"value": 5.0, "time": 0.001}, {"value": -5.0, "time": 0.025}]]],
    }
    lockstep(el.render(node), script, 7, exact=False)


def test_time_and_metro_follow_the_host_sample_clock():
    # wasm/Main.cpp:232-241 setCurrentTime: the host may move the clock; values far beyond float's 2^24 integer range
    g = el.add(el.mul(1e-6, el.time()), el.metro({"interval": 3.0}))
    times = [0, 512, 10_000_000_000, 10_000_000_512, 77, 123456789]
    lockstep(el.render(g), {}, len(times), sample_times=times)


def test_process_takes_the_sample_clock_from_user_data():
    rt = Runtime(SR, BS, 2, device=0)
    assert rt.apply_instructions(el.render(el.time())) == 0
    for _ in range(3):
        rt.process(None, 1, BS)                                # pass the root fade; engine-kept clock runs 0, 512, 1024
    assert rt.current_time() == 3 * BS
    y = rt.process(None, 1, BS, sample_time=1_000_000)[0]
    assert np.array_equal(y, 2.0 * (1_000_000 + np.arange(BS)).astype(np.float64).astype(np.float32))   # mix of 2 voices
    assert rt.current_time() == 1_000_000 + BS


def test_baked_properties_take_effect_at_the_next_block_without_a_commit():
    """svf.mode / delay.size / maxhold.hold are atomics or queued buffers in the reference (SVF.h:30-46, Delays.h:59-76,
    92-95, Core.h:292-303): a bare SET_PROPERTY must be heard in the very next process() call."""
    x = el.saw(220.0)
    svf = el.svf({"mode": "lowpass", "key": "f"}, 800.0, 2.0, x)
    dly = el.delay({"size": 1000, "key": "d"}, 400.25, 0.4, x)
    script = {
        2: [[3, svf.id(), "mode", "highpass"]],
        4: [[3, dly.id(), "size", 600]],                        # new ring: zeroed, write index reset (Delays.h:92-95)
        5: [[3, svf.id(), "mode", "notch"]],
    }
    lockstep(el.render(svf, dly), script, 8, n_out=2, exact=False)
