"""GPU parity of K3 (the `convolve` node) against the compiled reference ConvolutionNode (TwoStageFFTConvolver).

The reference convolver's own error against an exact convolution is ~1e-7 of the output peak while its pointwise
relative error is already > 1e-5, so the criterion is peak-normalised: max|gpu - ref| <= 1e-5 * max|ref|
over the rendered signal (the reference's own self-test uses 1e-4*ln(irLen) relative AND 1e-3*irLen absolute,
wasm/FFTConvolver/test/Test.cpp:125-140)."""
import numpy as np
import pytest

from elementary_b200 import Runtime, el, graphs
from helpers import oracle_render

pytestmark = pytest.mark.gpu
SR = 48000.0


def noise(n, seed, amp=0.25):
    s = (seed * 2654435761 + 7) & 0xFFFFFFFF
    out = np.empty(n, dtype=np.float32)
    for i in range(n):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        out[i] = amp * (2.0 * ((s >> 8) / float(1 << 24)) - 1.0)
    return out


def run(batch, ir, n_voices, n_blocks, bs=512, call_sizes=None):
    rt = Runtime(SR, bs, n_voices, device=0)
    assert rt.add_shared_resource("ir", ir)
    assert rt.apply_instructions(batch) == 0, rt.last_error()
    x = np.stack([noise(n_blocks * bs, 31 * v + 1)[None, :] for v in range(n_voices)])     # [voice, 1, n]
    got, _ = rt.render_voices(n_blocks, 1, x)
    ref = oracle_render(batch, n_blocks, 1, SR, bs, x, voice_batches=[None] * n_voices, resources={"ir": ir})
    return got, ref, rt


def check(got, ref, tol=1e-5):
    peak = np.abs(ref).max()
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64)).max()
    assert peak > 0 and err <= tol * peak, f"max err {err:.3g} vs peak {peak:.3g} (ratio {err / peak:.3g})"


def test_convolve_16384_tap_ir_40_blocks():
    ir = np.asarray(graphs.lcg_ir(16384), dtype=np.float32)
    got, ref, rt = run(graphs.convolve_channel("ir"), ir, 6, 40)      # 6 channels: ragged against 4 channels per CTA
    check(got, ref)


@pytest.mark.parametrize("taps", [1, 300, 512, 513, 1000, 4096, 4097, 9000])
def test_convolve_ir_lengths(taps):
    ir = np.asarray(graphs.lcg_ir(16384)[:taps], dtype=np.float32) * 3.0
    got, ref, rt = run(graphs.convolve_channel("ir"), ir, 3, 22)
    check(got, ref)


def test_convolve_trailing_silence_is_trimmed_and_zero_ir_is_silent():
    ir = np.concatenate([np.asarray(graphs.lcg_ir(700), dtype=np.float32), np.zeros(900, dtype=np.float32)])
    got, ref, rt = run(graphs.convolve_channel("ir"), ir, 2, 8)
    check(got, ref)
    got, ref, rt = run(graphs.convolve_channel("ir"), np.zeros(64, dtype=np.float32), 2, 3)
    assert not got.any() and not ref.any()


def test_convolve_inside_a_larger_graph_with_dry_path():
    # values cross the K1 | K3 | K1 stage boundary: the dry signal is needed after the convolver
    x = el.tanh(el.mul(2.0, el.in_(0)))
    g = el.add(el.mul(0.7, el.convolve({"path": "ir"}, x)), el.mul(0.3, x), el.mul(0.1, el.cycle(220.0)))
    ir = np.asarray(graphs.lcg_ir(3000), dtype=np.float32) * 2.0
    got, ref, rt = run(el.render(g), ir, 5, 12)
    check(got, ref)


def test_two_convolvers_in_series():
    g = el.convolve({"path": "ir", "key": "b"}, el.mul(0.5, el.convolve({"path": "ir", "key": "a"}, el.in_(0))))
    ir = np.asarray(graphs.lcg_ir(1500), dtype=np.float32) * 4.0
    got, ref, rt = run(el.render(g), ir, 2, 10)
    check(got, ref, tol=2e-5)


@pytest.mark.parametrize("bs", [64, 1024])
def test_convolve_other_block_sizes(bs):
    ir = np.asarray(graphs.lcg_ir(2500), dtype=np.float32) * 2.0
    got, ref, rt = run(graphs.convolve_channel("ir"), ir, 2, 3 * 2048 // bs, bs=bs)
    check(got, ref)


def test_convolve_varying_call_lengths():
    from helpers import oracle_cls
    ir = np.asarray(graphs.lcg_ir(2000), dtype=np.float32) * 2.0
    batch = graphs.convolve_channel("ir")
    rt = Runtime(SR, 512, 1, device=0)
    assert rt.add_shared_resource("ir", ir) and rt.apply_instructions(batch) == 0
    o = oracle_cls()(SR, 512)
    assert o.add_shared_resource("ir", ir) and o.apply(batch) == 0
    outs_g, outs_r = [], []
    for i, n in enumerate([512, 7, 100, 512, 405, 1, 333, 512, 512]):
        x = noise(n, 100 + i)[None, :]
        outs_g.append(rt.process_voices(x[None], 1, n)[0][0]); outs_r.append(o.process(x, 1, n))
    g, r = np.concatenate(outs_g, axis=1), np.concatenate(outs_r, axis=1)
    check(g, r)


@pytest.mark.parametrize("n_out", [1, 2])
def test_convolve_root_epilogue_equals_the_k1_stage(n_out):
    """`convolve -> root` (BASELINE config 4): K3 applies the root's gain fade and writes the per-voice output and the partial mix
    itself (convolve.h ConvEpilogue); with the option off a K1 stage does it.  Same bits — through the 20 ms root fade-in, with ragged
    call lengths (partition fill / wrap inside one call), per-voice outputs AND the mix bus, a silent second output channel, a fully
    trimmed IR, and a graph update (root cross-fade: the old root fades out on the K1 path) — and the same bits as before the update."""
    bs = 512
    ir = np.asarray(graphs.lcg_ir(3000), dtype=np.float32)
    sizes = [512, 512, 100, 412, 37, 512, 1, 474, 512, 512]
    n_voices = 5
    x = np.stack([noise(sum(sizes), 17 * v + 3)[None, :] for v in range(n_voices)])
    outs = {}
    for fuse in (1, 0):
        rt = Runtime(SR, bs, n_voices, device=0, fuse_conv_root=fuse)
        assert rt.add_shared_resource("ir", ir) and rt.add_shared_resource("zero", np.zeros(64, dtype=np.float32))
        assert rt.apply_instructions(graphs.convolve_channel("ir")) == 0, rt.last_error()
        vs, ms, pos = [], [], 0
        for k, n in enumerate(sizes):
            if k == 7:       # live edit: the same graph behind a gain (new root, old one fades out; not the fused shape any more)
                rg = el.Renderer()
                rg.render(el.convolve({"path": "ir"}, el.in_(0)))
                assert rt.apply_instructions(rg.render(el.mul(0.5, el.convolve({"path": "ir"}, el.in_(0))))) == 0, rt.last_error()
            v, m = rt.process_voices(x[:, :, pos:pos + n], n_out, n)
            vs.append(v); ms.append(m); pos += n
        outs[fuse] = (np.concatenate(vs, axis=2), np.concatenate(ms, axis=1))
        z = Runtime(SR, bs, 3, device=0, fuse_conv_root=fuse)
        assert z.add_shared_resource("ir", np.zeros(64, dtype=np.float32))
        assert z.apply_instructions(graphs.convolve_channel("ir")) == 0
        v, m = z.process_voices(x[:3, :, :bs], n_out, bs)
        assert not v.any() and not m.any()
    assert np.array_equal(outs[1][0], outs[0][0]) and np.array_equal(outs[1][1], outs[0][1])
    assert np.abs(outs[1][0][:, 0]).max() > 0
    if n_out == 2:
        assert not outs[1][0][:, 1].any() and not outs[1][1][1].any()
    ref = oracle_render(graphs.convolve_channel("ir"), 2, 1, SR, bs, x[:, :, :1024], voice_batches=[None] * n_voices, resources={"ir": ir})
    check(outs[1][0][:, 0:1, :1024], ref)
