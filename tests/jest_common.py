"""Replay of the reference's jest tests (js/packages/offline-renderer/__tests__/{sparseq,sparseq2,time,events}.test.js) against
any engine: a minimal OfflineRenderer (offline-renderer/index.ts:87-133 — block loop of 512 with zero-padded inputs,
processQueuedEvents after every block) over an `engine` adapter, and the jest tests transcribed one to one.  Each replay returns
{snapshot key: produced value}; tests/test_jest_goldens.py compares them with tests/golden/jest_snapshots.json."""
import numpy as np

from elementary_b200 import el

BS = 512


class OfflineRenderer:
    """engine: an object with apply(batch) -> rc, process(inputs [n_in, 512] | None, n_out, 512) -> [n_out, 512],
    process_queued_events() -> [{type, event}], set_current_time(t)."""

    def __init__(self, make_engine, sample_rate=44100.0, n_in=1, n_out=1):
        self.sr, self.n_in, self.n_out = sample_rate, n_in, n_out
        self.engine = make_engine(sample_rate, BS)
        self.renderer = el.Renderer()
        self.events = []

    def render(self, *graphs):
        assert self.engine.apply(self.renderer.render(*graphs)) == 0

    def process(self, inps, n):
        """inps: list of n_in 1-D arrays (any length); returns [n_out, n]."""
        out = np.zeros((self.n_out, n), dtype=np.float32)
        for k in range(0, n, BS):
            x = None
            if self.n_in:
                x = np.zeros((self.n_in, BS), dtype=np.float32)
                for c, buf in enumerate(inps):
                    seg = np.asarray(buf, dtype=np.float32)[k:k + BS]
                    x[c, :len(seg)] = seg
            y = self.engine.process(x, self.n_out, BS)
            self.events.extend(self.engine.process_queued_events())
            m = min(BS, n - k)
            out[:, k:k + m] = y[:, :m]
        return out

    def warm(self):
        self.process([np.zeros(BS * 10, dtype=np.float32)] * self.n_in, BS * 10)     # "Get past the fade-in"


SEQ = [{"value": 1, "tickTime": 0}, {"value": 2, "tickTime": 2}, {"value": 3, "tickTime": 4}, {"value": 4, "tickTime": 8}]
IN0 = el.in_(0)
CLOCK = lambda n: np.array([(i + 1) % 2 for i in range(n)], dtype=np.float32)        # (new Float32Array(n)).map((x, i) => (i + 1) % 2)


def replay_all(make_engine):
    """Returns {snapshot key: value}.  Citations: sparseq.test.js / sparseq2.test.js / time.test.js / events.test.js of the reference."""
    R = {}
    S = "sparseq.test.js.snap:"

    core = OfflineRenderer(make_engine)                                            # sparseq basics (:5-43)
    core.render(el.sparseq({"seq": SEQ}, IN0, 0)); core.warm()
    x = np.array([0] + [1, 0] * 12, dtype=np.float32)
    R[S + "sparseq basics 1"] = core.process([x], len(x))[0]

    core = OfflineRenderer(make_engine)                                            # sparseq loop (:45-78)
    core.render(el.sparseq({"seq": SEQ, "loop": [2, 4]}, IN0, 0)); core.warm()
    R[S + "sparseq loop 1"] = core.process([CLOCK(32)], 32)[0]

    core = OfflineRenderer(make_engine)                                            # sparseq loop on then off (:80-124)
    core.render(el.sparseq({"key": "test", "seq": SEQ, "loop": [2, 4]}, IN0, 0)); core.warm()
    core.process([CLOCK(32)], 32)
    core.render(el.sparseq({"key": "test", "seq": SEQ, "loop": False}, IN0, 0))
    R[S + "sparseq loop on then off 1"] = core.process([CLOCK(32)], 32)[0]

    core = OfflineRenderer(make_engine)                                            # sparseq no trigger on reset (:126-175)
    core.render(el.sparseq({"seq": SEQ, "loop": False}, IN0, el.const(0, key="reset"))); core.warm()
    core.process([CLOCK(8)], 8)
    core.render(el.sparseq({"seq": SEQ, "loop": False}, IN0, el.const(1, key="reset")))
    R[S + "sparseq no trigger on reset 1"] = core.process([np.zeros(8, dtype=np.float32)], 8)[0]

    core = OfflineRenderer(make_engine)                                            # sparseq interpolation (:177-212)
    core.render(el.sparseq({"seq": SEQ, "interpolate": 1}, IN0, 0)); core.warm()
    R[S + "sparseq interpolation 1"] = core.process([CLOCK(24)], 24)[0]

    core = OfflineRenderer(make_engine)                                            # sparseq interpolation with loop (:214-247)
    core.render(el.sparseq({"seq": SEQ, "interpolate": 1, "loop": [1, 3]}, IN0, 0)); core.warm()
    R[S + "sparseq interpolation with loop 1"] = core.process([CLOCK(24)], 24)[0]

    core = OfflineRenderer(make_engine, sample_rate=1000.0)                        # sparseq sub-tick interpolation (:249-298)
    core.render(el.sparseq({"seq": SEQ, "interpolate": 1, "tickInterval": 0.002}, IN0, el.const(0, key="reset"))); core.warm()
    R[S + "sparseq sub-tick interpolation 1"] = core.process([CLOCK(24)], 24)[0]
    core.render(el.sparseq({"seq": SEQ, "interpolate": 1, "tickInterval": 0.002}, IN0, el.const(1, key="reset")))
    x = np.array([0 if i > 7 else (i + 1) % 2 for i in range(24)], dtype=np.float32)
    R[S + "sparseq sub-tick interpolation 2"] = core.process([x], 24)[0]

    seq3 = [{"value": 0, "tickTime": 0}, {"value": 0, "tickTime": 4}, {"value": 1, "tickTime": 8}]
    core = OfflineRenderer(make_engine, sample_rate=1000.0)                        # sub-tick interpolation with loop (:300-337)
    core.render(el.sparseq({"seq": seq3, "interpolate": 1, "tickInterval": 0.002, "loop": [4, 8], "offset": 4}, IN0, 0)); core.warm()
    R[S + "sparseq sub-tick interpolation with loop 1"] = core.process([CLOCK(24)], 24)[0]

    core = OfflineRenderer(make_engine, sample_rate=2000.0)                        # ... higher res (:339-381)
    core.render(el.sparseq({"seq": seq3, "interpolate": 1, "tickInterval": 0.01, "loop": [4, 8], "offset": 4}, el.train(100), 0)); core.warm()
    R[S + "sparseq sub-tick interpolation with loop higher res 1"] = core.process([np.zeros(BS, dtype=np.float32)], BS)[0]

    seq4 = [{"value": 1, "tickTime": 0}, {"value": 2, "tickTime": 1}, {"value": 3, "tickTime": 2}, {"value": 4, "tickTime": 3}]
    core = OfflineRenderer(make_engine)                                            # sparseq loop follow (:383-428)
    core.render(el.sparseq({"key": "test", "seq": seq4, "loop": [1, 3]}, IN0, 0)); core.warm()
    R[S + "sparseq loop follow 1"] = core.process([CLOCK(32)], 32)[0]
    core.render(el.sparseq({"key": "test", "seq": seq4, "loop": [0, 2], "follow": True}, IN0, 0))
    R[S + "sparseq loop follow 2"] = core.process([CLOCK(32)], 32)[0]

    S2 = "sparseq2.test.js.snap:"
    t0 = BS * 10
    mk = lambda offs: [{"time": t0 + o, "value": v} for o, v in zip(offs, (1, 2, 3, 4))]
    core = OfflineRenderer(make_engine, n_in=0)                                    # sparseq2 basics (:5-37)
    core.render(el.sparseq2({"seq": mk((4, 8, 12, 16))}, el.time())); core.warm()
    R[S2 + "sparseq2 basics 1"] = core.process([], 32)[0]
    core = OfflineRenderer(make_engine, n_in=0)                                    # sparseq2 interp (:39-75)
    core.render(el.sparseq2({"interpolate": 1, "seq": mk((0, 4, 8, 12))}, el.time())); core.warm()
    R[S2 + "sparseq2 interp 1"] = core.process([], 32)[0]
    loop = lambda start, end, t: el.add(start, el.mod(t, el.sub(end, start)))
    core = OfflineRenderer(make_engine, n_in=0)                                    # sparseq2 looping (:77-113)
    core.render(el.sparseq2({"seq": mk((0, 4, 8, 12))}, loop(5120, 5120 + 16, el.time()))); core.warm()
    R[S2 + "sparseq2 looping 1"] = core.process([], 32)[0]
    core = OfflineRenderer(make_engine)                                            # sparseq2 skip ahead (:115-153)
    core.render(el.sparseq2({"seq": mk((0, 4, 8, 12))}, IN0)); core.warm()
    x = np.array([5120] * 8 + [5128] * 8, dtype=np.float32)
    R[S2 + "sparseq2 skip ahead 1"] = core.process([x], 16)[0]

    core = OfflineRenderer(make_engine)                                            # time node (time.test.js:5-30)
    core.render(el.time()); core.warm()
    R["time.test.js.snap:time node 1"] = core.process([np.zeros(32, dtype=np.float32)], 32)[0]
    core = OfflineRenderer(make_engine, n_in=0)                                    # setting time (:32-61): inline expectations
    core.render(el.time()); core.warm()
    core.engine.set_current_time(50)
    R["inline:setting time 1"] = core.process([], 8)[0]
    core.engine.set_current_time(int(1000 / 1000.0 * core.sr))                     # setCurrentTimeMs(1000), wasm/Main.cpp:237-241
    R["inline:setting time 2"] = core.process([], 8)[0]

    core = OfflineRenderer(make_engine, n_in=0)                                    # event propagation (events.test.js:5-45)
    core.render(el.meter({}, 0))
    core.events.clear()
    core.process([], BS * 4)
    R["events.test.js.snap:event propagation 1"] = [e["event"] for e in core.events if e["type"] == "meter"]
    core.events.clear()
    core.render(el.meter({}, 1))
    core.process([], BS * 4)
    R["events.test.js.snap:event propagation 2"] = [e["event"] for e in core.events if e["type"] == "meter"]
    return R


INLINE = {"inline:setting time 1": [50, 51, 52, 53, 54, 55, 56, 57],
          "inline:setting time 2": [44100, 44101, 44102, 44103, 44104, 44105, 44106, 44107]}


def compare(produced, snapshots, atol=2e-6):
    """Returns a list of mismatch descriptions (empty = all golden vectors reproduced)."""
    bad = []
    for key, got in produced.items():
        want = INLINE.get(key, snapshots.get(key))
        if want is None:
            bad.append(f"{key}: no golden value")
            continue
        if key.startswith("events"):
            g = [{k: v for k, v in e.items() if k != "voice"} for e in got]
            # jest prints `undefined` sources as undefined; js::serialize turns them into null
            w = [{k: (v if v != "undefined" else None) for k, v in e.items()} for e in want]
            if len(g) != len(w) or any(abs(a["min"] - b["min"]) > atol or abs(a["max"] - b["max"]) > atol or a.get("source") != b.get("source") for a, b in zip(g, w)):
                bad.append(f"{key}: {g} != {w}")
            continue
        g, w = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
        if g.shape != w.shape or np.abs(g - w).max() > atol * max(1.0, np.abs(w).max()):
            bad.append(f"{key}: got {np.round(g[:12], 6)}... want {np.round(w[:12], 6)}...")
    return bad
