"""Graph cases shared by the CPU oracle tests (port vs compiled reference) — one entry per node family."""
import math

import numpy as np

from elementary_b200 import el, graphs

IN0, IN1, IN2 = el.in_(0), el.in_(1), el.in_(2)
_gate = el.le(el.phasor(900.0), 0.3)
_tab = np.sin(np.linspace(0, 2 * np.pi, 1000, dtype=np.float64)).astype(np.float32)


def _c(name, graph, n_in=0, n_blocks=3, resources=None, n_out=1, batch=None):
    return dict(name=name, batch=batch if batch is not None else el.render(*(graph if isinstance(graph, tuple) else (graph,))),
                n_in=n_in, n_blocks=n_blocks, resources=resources, n_out=n_out)


CASES = []
for _u in ["sin", "cos", "tan", "tanh", "asinh", "ceil", "floor", "round_", "exp", "abs_"]:
    CASES.append(_c(_u, getattr(el, _u)(el.mul(3.0, IN0)), 1))
for _u in ["ln", "log", "log2", "sqrt"]:
    CASES.append(_c(_u, getattr(el, _u)(el.abs_(IN0)), 1))
for _b in ["le", "leq", "ge", "geq", "eq", "and_", "or_"]:
    CASES.append(_c(_b, getattr(el, _b)(el.round_(el.mul(2.0, IN0)), el.round_(el.mul(2.0, IN1))), 2))
CASES.append(_c("pow", el.pow_(el.mul(2.0, IN0), el.round_(el.mul(3.0, IN1))), 2))
for _r in ["add", "sub", "mul", "div", "mod", "min_", "max_"]:
    CASES.append(_c(_r, getattr(el, _r)(IN0, IN1, IN2, 0.37), 3))
CASES += [
    _c("phasor", el.phasor(440.0), n_blocks=6),
    _c("cycle", el.cycle(440.0), n_blocks=8),
    _c("saw", el.saw(110.0)),
    _c("phasor_audio_rate", el.phasor(el.mul(2000.0, IN0)), 1),
    _c("sphasor", el.syncphasor(220.0, _gate), n_blocks=4),
    _c("counter", el.counter(_gate)),
    _c("accum", el.accum(IN0, _gate), 1),
    _c("latch", el.latch(_gate, IN0), 1),
    _c("maxhold", el.maxhold({}, IN0, _gate), 1),
    _c("maxhold_hold", el.maxhold({"hold": 1.0}, IN0, 0.0), 1),
    _c("rand", el.rand(seed=12345)),
    _c("pole", el.pole(el.mul(0.9, IN1), IN0), 2),
    _c("env", el.env(0.9, 0.999, IN0), 1),
    _c("biquad", el.biquad(0.2, 0.4, 0.2, -0.5, 0.3, IN0), 1),
    _c("z", el.z(IN0), 1),
    _c("smooth", el.smooth(el.tau2pole(0.01), IN0), 1),
    _c("svf_extreme", el.svf({}, el.mul(40000.0, IN0), el.mul(30.0, IN1), IN2), 3),
    _c("delay_4800", el.delay({"size": 4800}, 3001.5, 0.35, IN0), 1, 12),
    _c("delay_mod", el.delay({"size": 100}, el.add(50.0, el.mul(49.0, IN1)), el.mul(1.5, IN2), IN0), 3, 4),
    _c("delay_zero", el.delay({"size": 64}, 0.0, 0.9, IN0), 1),
    _c("delay_default", el.delay({}, 17.25, 0.0, IN0), 1),
    _c("sdelay_10", el.sdelay({"size": 10}, IN0), 1),
    _c("sdelay_3000", el.sdelay({"size": 3000}, IN0), 1, 8),
    _c("sdelay_0", el.sdelay({"size": 0}, IN0), 1),
    _c("table", el.table({"path": "wt"}, el.phasor(441.0)), 0, 4, {"wt": _tab}),
    _c("table_clamped", el.table({"path": "wt"}, el.mul(1.3, IN0)), 1, 3, {"wt": _tab}),
    _c("taps", el.tap_out("t", el.add(el.tap_in("t"), IN0)), 1, 5),
    _c("two_roots", (el.mul(0.5, el.cycle(330.0)), el.tanh(el.mul(3.0, el.cycle(330.0)))), 0, 4, None, 2),
    _c("sr", el.div(el.sr(), 48000.0)),
    _c("leaf_sin", None, 1, 3, None, 1, [[0, 1, "root"], [0, 2, "sin"], [2, 1, 2, 0], [3, 1, "channel", 0], [4, [1]], [5]]),
    _c("missing_inputs", None, 0, 2, None, 1,
       [[0, 1, "root"], [0, 2, "svf"], [0, 3, "const"], [2, 2, 3, 0], [2, 1, 2, 0], [3, 1, "channel", 0], [4, [1]], [5]]),
    _c("subsynth32", None, 0, 12, None, 1, graphs.subsynth32(110.0)),
    _c("additive16", None, 0, 4, None, 1, graphs.additive64(110.0, 16)),
    _c("plumbing", None, 0, 4, None, 2, graphs.plumbing()),
]
for _m in ["lowpass", "bandpass", "highpass", "notch", "allpass"]:
    CASES.append(_c("svf_" + _m, el.svf({"mode": _m}, el.add(1500.0, el.mul(1400.0, el.cycle(3.0))), 1.5, IN0), 1, 4))
for _m in ["lowshelf", "highshelf", "bell"]:
    CASES.append(_c("svfshelf_" + _m, el.svfshelf({"mode": _m}, 900.0, 0.8, el.mul(12.0, IN1), IN0), 2))
for _m in ["lowpass", "highpass", "allpass"]:
    CASES.append(_c("mm1p_" + _m, el.mm1p({"mode": _m}, el.prewarp(el.add(2000.0, el.mul(1500.0, IN1))), IN0), 2))
for _k in ["blepsaw", "blepsquare", "bleptriangle"]:
    CASES.append(_c(_k, getattr(el, _k)(el.add(1500.0, el.mul(1400.0, IN0))), 1, 4))
for _s in range(4):
    CASES.append(_c(f"random_graph_{_s}", None, 0, 6, None, 1, graphs.random_graph(_s, 32)))


def lcg_noise(n, seed, lo=-1.0, hi=1.0):
    s = (seed * 2654435761 + 1) & 0xFFFFFFFF
    out = np.empty(n, dtype=np.float32)
    for i in range(n):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        out[i] = lo + (hi - lo) * ((s >> 8) / float(1 << 24))
    return out


def case_inputs(case, bs=512, seed=0):
    if not case["n_in"]:
        return None
    return np.stack([lcg_noise(case["n_blocks"] * bs, seed * 100 + c) for c in range(case["n_in"])])
