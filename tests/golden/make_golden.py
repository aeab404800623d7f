"""Generate tests/golden/reference_cases.npz from the UNMODIFIED reference compiled in place (oracle/_ref).

Run in the authoring container (where /root/reference exists):
    python -c "import __graft_entry__ as g; g.build()"      # builds oracle/_ref/libelem_ref.so from /root/reference
    python tests/golden/make_golden.py
Every array is the float32 output [n_out, n_blocks*512] of one `elem::Runtime<float>(48000, 512)` instance fed the
instruction batch of tests/cases.py::CASES[name] and the deterministic LCG inputs of cases.case_inputs — i.e. exactly
what the reference engine writes (glibc libm of this image, g++ -O2 -ffp-contract=off).  The fixtures let the parity
tests run against the reference's outputs on machines where neither /root/reference nor oracle/_ref exists.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as orc          # noqa: E402
from cases import CASES, case_inputs      # noqa: E402
from elementary_b200 import graphs       # noqa: E402

SR, BS = 48000.0, 512


def main():
    assert orc.ref_available(), "oracle/_ref/libelem_ref.so missing: build it where /root/reference exists"
    out, meta = {}, {}
    for case in CASES:
        r = orc.RefRuntime(SR, BS)
        for k, v in (case["resources"] or {}).items():
            assert r.add_shared_resource(k, v)
        assert r.apply(case["batch"]) == 0
        y = r.render(case["n_blocks"], case["n_out"], BS, case_inputs(case))
        out[case["name"]] = y
        meta[case["name"]] = hashlib.sha256(json.dumps(case["batch"]).encode()).hexdigest()[:16]
    # the convolver node on the BASELINE config-4 IR (16384 taps) and a short trimmed IR
    for taps, blocks in ((16384, 40), (700, 6)):
        ir = np.asarray(graphs.lcg_ir(16384)[:taps], dtype=np.float32)
        rng = np.random.RandomState(taps)
        x = ((rng.rand(1, blocks * BS) - 0.5) * 0.5).astype(np.float32)
        r = orc.RefRuntime(SR, BS)
        assert r.add_shared_resource("ir", ir) and r.apply(graphs.convolve_channel("ir")) == 0
        out[f"convolve_{taps}"] = r.render(blocks, 1, BS, x)
        out[f"convolve_{taps}_input"] = x
        meta[f"convolve_{taps}"] = "graphs.convolve_channel('ir'), ir = graphs.lcg_ir(16384)[:taps]"
    np.savez_compressed(os.path.join(HERE, "reference_cases.npz"), **out)
    with open(os.path.join(HERE, "reference_cases.json"), "w") as f:
        json.dump({"sample_rate": SR, "block": BS, "generator": "tests/golden/make_golden.py",
                   "engine": "elem::Runtime<float> compiled from /root/reference (oracle/Makefile)",
                   "batch_sha256_16": meta}, f, indent=1, sort_keys=True)
    print(len(out), "arrays,", os.path.getsize(os.path.join(HERE, "reference_cases.npz")), "bytes")


if __name__ == "__main__":
    main()
