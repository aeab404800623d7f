"""CPU: the restatement (oracle/elem_oracle.cpp) against the compiled reference on the LIVE-UPDATE scenarios of the sequencing /
control nodes (tests/control_common.py) — new sequence data, loop points, re-arming, host clock jumps, baked properties changed
between blocks.  Everything except transcendental-free paths must be bit-exact; the others agree to float rounding of libm-free
double arithmetic, i.e. exactly as well."""
import numpy as np
import pytest

from control_common import scenarios
from oracle import oracle as orc

SR, BS = 48000.0, 512
SCEN = scenarios()


@pytest.mark.skipif(not orc.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("sc", SCEN, ids=[s["name"] for s in SCEN])
def test_port_matches_reference_through_live_updates(sc):
    outs = []
    for cls in (orc.PortRuntime, orc.RefRuntime):
        r = cls(SR, BS)
        assert r.apply(sc["batch"]) == 0
        blocks = []
        for b in range(sc["n_blocks"]):
            if b in sc["script"]:
                assert r.apply(sc["script"][b]) == 0
            if sc.get("sample_times"):
                r.set_current_time(sc["sample_times"][b])
            blocks.append(r.process(None, sc["n_out"], BS))
        outs.append(np.concatenate(blocks, axis=1))
    assert np.array_equal(outs[0], outs[1]), f"{sc['name']}: max diff {np.abs(outs[0] - outs[1]).max()}"
    assert np.abs(outs[1]).max() > 0
