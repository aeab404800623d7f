"""Golden fixtures generated from the unmodified reference (tests/golden/make_golden.py -> reference_cases.npz).

CPU: the restatement (oracle/elem_oracle.cpp) must reproduce every fixture bit-exactly (convolver: float rounding).
GPU: the CUDA path through the C ABI must reproduce every fixture within the north_star tolerance
(|gpu - ref| <= 1e-5 * block peak + 1e-7).  These tests need neither /root/reference nor oracle/_ref.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from elementary_b200 import graphs
from oracle import oracle as orc
from cases import CASES, case_inputs
from helpers import block_peak_tolerance_check

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "reference_cases.npz"))
META = json.load(open(os.path.join(HERE, "golden", "reference_cases.json")))
SR, BS = 48000.0, 512
IDS = [c["name"] for c in CASES]


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_fixture_is_current(case):
    assert META["batch_sha256_16"][case["name"]] == hashlib.sha256(json.dumps(case["batch"]).encode()).hexdigest()[:16], \
        "tests/cases.py changed: re-run tests/golden/make_golden.py where /root/reference exists"


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_port_reproduces_reference_fixture_bit_exactly(case):
    r = orc.PortRuntime(SR, BS)
    for k, v in (case["resources"] or {}).items():
        assert r.add_shared_resource(k, v)
    assert r.apply(case["batch"]) == 0
    got = r.render(case["n_blocks"], case["n_out"], BS, case_inputs(case))
    assert np.array_equal(got, GOLD[case["name"]])


@pytest.mark.parametrize("taps", [16384, 700])
def test_port_convolver_reproduces_reference_fixture(taps):
    x = GOLD[f"convolve_{taps}_input"]
    r = orc.PortRuntime(SR, BS)
    assert r.add_shared_resource("ir", np.asarray(graphs.lcg_ir(16384)[:taps], dtype=np.float32))
    assert r.apply(graphs.convolve_channel("ir")) == 0
    got = r.render(x.shape[1] // BS, 1, BS, x)
    want = GOLD[f"convolve_{taps}"]
    assert np.abs(got - want).max() <= 5e-7 * np.abs(want).max()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_gpu_reproduces_reference_fixture(case):
    from elementary_b200 import Runtime
    rt = Runtime(SR, BS, 2, device=0)
    for k, v in (case["resources"] or {}).items():
        assert rt.add_shared_resource(k, v)
    assert rt.apply_instructions(case["batch"]) == 0, rt.last_error()
    inp = case_inputs(case)
    inputs = None if inp is None else np.stack([inp, inp])
    got, _ = rt.render_voices(case["n_blocks"], case["n_out"], inputs)
    want = GOLD[case["name"]]
    for v in range(2):
        ok, worst, ex = block_peak_tolerance_check(got[v], want, BS)
        assert ok, f"{case['name']} voice {v}: worst err/tol {worst:.3g}, bit-exact fraction {ex:.4f}"


@pytest.mark.gpu
@pytest.mark.parametrize("taps", [16384, 700])
def test_gpu_convolver_reproduces_reference_fixture(taps):
    from elementary_b200 import Runtime
    x = GOLD[f"convolve_{taps}_input"]
    rt = Runtime(SR, BS, 1, device=0)
    assert rt.add_shared_resource("ir", np.asarray(graphs.lcg_ir(16384)[:taps], dtype=np.float32))
    assert rt.apply_instructions(graphs.convolve_channel("ir")) == 0, rt.last_error()
    got, _ = rt.render_voices(x.shape[1] // BS, 1, x[None])
    want = GOLD[f"convolve_{taps}"]
    assert np.abs(got[0] - want).max() <= 1e-5 * np.abs(want).max()
