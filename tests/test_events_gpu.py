"""GPU parity for the event path: elem_b200_process_queued_events against the oracle's
Runtime::processQueuedEvents on the scenarios of tests/events_common.py — every voice gets its own input noise, so
every voice must report its own readings; polling patterns exercise the queue / ring arithmetic of the reference
(SingleWriterSingleReaderQueue.h, MultiChannelRingBuffer.h)."""
import numpy as np
import pytest

from elementary_b200 import Runtime, el
from events_common import scenarios, noise, canon, same
from helpers import oracle_cls

pytestmark = pytest.mark.gpu
SR, BS = 48000.0, 512
SCEN = scenarios()


@pytest.mark.parametrize("sc", SCEN, ids=[s["name"] for s in SCEN])
@pytest.mark.parametrize("tile_width", [0, 2, 16, 32])
def test_events_match_oracle_per_voice(sc, tile_width):
    n_voices = 3 if tile_width in (0, 2) else 35
    batch = el.render(*sc["graph"])
    opts = {"tile_width": tile_width} if tile_width else {}
    rt = Runtime(SR, BS, n_voices, device=0, **opts)
    assert rt.apply_instructions(batch) == 0, rt.last_error()
    oracles = [oracle_cls()(SR, BS) for _ in range(n_voices)]
    for o in oracles:
        assert o.apply(batch) == 0
    n_out = len(sc["graph"])
    seen = False
    for b in range(sc["blocks"]):
        x = None
        if sc["n_in"]:
            x = np.stack([np.stack([noise(BS, 1000 * b + 7 * v + c) for c in range(sc["n_in"])]) for v in range(n_voices)])
        got_audio = rt.process_voices(x, n_out, BS)[0]
        ref_audio = np.stack([o.process(None if x is None else x[v], n_out, BS) for v, o in enumerate(oracles)])
        assert np.abs(got_audio - ref_audio).max() <= 1e-5 * max(1e-9, np.abs(ref_audio).max())
        if sc["poll"](b):
            events = rt.process_queued_events()
            for v, o in enumerate(oracles):
                want = canon(o.process_queued_events())
                got = canon([e for e in events if e["event"]["voice"] == v])
                assert same(got, want), f"{sc['name']}: block {b} voice {v}: {str(got)[:300]} != {str(want)[:300]}"
                seen = seen or bool(want)
    assert seen != bool(sc.get("silent"))


def test_events_for_a_voice_range_leave_the_other_queues_alone():
    rt = Runtime(SR, BS, 8, device=0)
    assert rt.apply_instructions(el.render(el.meter({"name": "m"}, el.in_(0)))) == 0
    x = np.stack([noise(BS, v)[None] for v in range(8)])
    rt.process_voices(x, 1, BS)
    first = rt.process_queued_events(voices=(2, 5))
    assert sorted(e["event"]["voice"] for e in first) == [2, 3, 4]
    rest = rt.process_queued_events()
    assert sorted(e["event"]["voice"] for e in rest) == [0, 1, 5, 6, 7]
    for e in first + rest:
        v = e["event"]["voice"]
        assert e["type"] == "meter" and e["event"]["source"] == "m"
        assert np.float32(e["event"]["min"]) == x[v].min() and np.float32(e["event"]["max"]) == x[v].max()
    assert rt.process_queued_events() == []


def test_inactive_roots_report_nothing():
    # GraphRenderSequence.h:189-198: a root that is fading out still renders but its nodes' events are not processed
    r = el.Renderer()
    a = r.render(el.meter({"name": "old"}, el.cycle(220.0)))
    b = r.render(el.meter({"name": "new"}, el.cycle(330.0)))
    rt = Runtime(SR, BS, 2, device=0)
    o = oracle_cls()(SR, BS)
    for batch in (a,):
        assert rt.apply_instructions(batch) == 0 and o.apply(batch) == 0
    rt.process_voices(None, 1, BS); o.process(None, 1, BS)
    assert rt.apply_instructions(b) == 0 and o.apply(b) == 0
    for _ in range(2):
        rt.process_voices(None, 1, BS); o.process(None, 1, BS)
        got = canon([e for e in rt.process_queued_events() if e["event"]["voice"] == 0])
        want = canon(o.process_queued_events())
        assert got == want and all(evt["source"] == "new" for _, evt in got)


def test_scope_property_validation_matches_reference():
    rt = Runtime(SR, BS, 1, device=0)
    o = oracle_cls()(SR, BS)
    for ins, code in [([[0, 5, "scope"], [3, 5, "size", 100]], 6), ([[3, 5, "size", 9000]], 6), ([[3, 5, "channels", 5]], 6),
                      ([[3, 5, "name", 3]], 5), ([[3, 5, "size", 1024], [3, 5, "channels", 4], [3, 5, "name", "ok"]], 0)]:
        assert rt.apply_instructions(ins) == code
        assert o.apply(ins) == code

