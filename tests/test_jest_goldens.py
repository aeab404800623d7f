"""The reference's own golden vectors for the sequencing / time / event nodes (jest snapshots, extracted by
tests/golden/make_jest_golden.py into tests/golden/jest_snapshots.json): the jest tests that made them are replayed
(tests/jest_common.py) against the CPU restatement, the compiled reference (float runtime; the snapshots come from the double
wasm runtime, hence a 2e-6 tolerance) and — on the GPU — the CUDA path."""
import json
import os

import numpy as np
import pytest

from jest_common import compare, replay_all
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SNAP = json.load(open(os.path.join(HERE, "golden", "jest_snapshots.json")))["snapshots"]


def _checkers():
    out = [orc.PortRuntime]
    if orc.ref_available():
        out.append(orc.RefRuntime)
    return out


@pytest.mark.parametrize("cls", _checkers(), ids=lambda c: c.__name__)
def test_oracles_reproduce_the_reference_jest_snapshots(cls):
    produced = replay_all(lambda sr, bs: cls(sr, bs))
    assert len(produced) >= 21
    bad = compare(produced, SNAP)
    assert not bad, "\n".join(bad)


class _GpuEngine:
    """Adapter: voice 0 of a 2-voice CUDA runtime behind the oracle's small interface."""

    def __init__(self, sr, bs):
        from elementary_b200 import Runtime
        self.rt = Runtime(sr, bs, 2, device=0)

    def apply(self, batch):
        return self.rt.apply_instructions(batch)

    def process(self, x, n_out, n):
        inputs = None if x is None else np.stack([x, x])
        return self.rt.process_voices(inputs, n_out, n)[0][0]

    def process_queued_events(self):
        return [e for e in self.rt.process_queued_events() if e["event"]["voice"] == 0]

    def set_current_time(self, t):
        self.rt.set_current_time(t)


@pytest.mark.gpu
def test_gpu_reproduces_the_reference_jest_snapshots():
    produced = replay_all(lambda sr, bs: _GpuEngine(sr, bs))
    bad = compare(produced, SNAP)
    assert not bad, "\n".join(bad)
