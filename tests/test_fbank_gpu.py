"""GPU: the fused fbank kernel, through the C ABI, against the reference goldens and the oracle.

Mirrors the reference's tests/unittests/test_batched_fbank.py (same inputs, same tolerances where an
independent fp32 FFT can meet them; the four-part gate of SURVEY.md section 7 otherwise).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import kaldi_fbank as OF
from tests.golden import cases

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
FB = np.load(os.path.join(G, "fbank.npz"))
REPORT = json.load(open(os.path.join(G, "REPORT.json")))["cases"]


@pytest.fixture(scope="module")
def fbank():
    from avex_b200.fbank import KaldiFbank

    return KaldiFbank().cuda()


def _stats(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.abs(a - b)
    return d.max(), np.linalg.norm(a - b) / np.linalg.norm(b), (d > 1e-4 + 1e-4 * np.abs(b)).mean()


NOISE = [k for k, (w, m) in cases.fbank_cases().items() if m == 128 and "sine" not in k]
TONES = [k for k, (w, m) in cases.fbank_cases().items() if m == 128 and "sine" in k]


@pytest.mark.parametrize("name", NOISE)
def test_matches_reference_golden(fbank, name):
    wav, _ = cases.fbank_cases()[name]
    ref = FB[name]
    out = fbank(torch.from_numpy(wav).cuda() * 2**15).cpu().numpy()
    assert out.shape == ref.shape
    mx, l2, frac = _stats(out, ref)
    print(f"{name}: max_abs {mx:.3e} rel_l2 {l2:.3e} frac_out {frac:.3e}")
    # (i) rel-L2, (ii) out-of-tolerance fraction at atol=rtol=1e-4, (iv) the reference's own GPU tolerance
    assert l2 <= 1e-5
    assert frac <= 1e-4
    np.testing.assert_allclose(out, ref, atol=2e-2, rtol=1e-3)  # test_batched_fbank.py:111-131
    # (iii) no worse against float64 than the reference itself (x2 slack for summation-order luck)
    f64 = OF.fbank(wav.astype(np.float64) * 32768.0, dtype=np.float64)
    assert np.abs(out - f64).max() <= 2.0 * max(REPORT["fbank/" + name]["ref_vs_f64"]["max_abs"], 1e-4)
    # oracle agrees too
    orc = OF.fbank(wav * np.float32(32768.0))
    assert _stats(out, orc)[1] <= 1e-5


@pytest.mark.parametrize("name", TONES)
def test_pure_tones_on_significant_bins(fbank, name):
    wav, _ = cases.fbank_cases()[name]
    ref = FB[name]
    out = fbank(torch.from_numpy(wav).cuda() * 2**15).cpu().numpy()
    f64 = OF.fbank(wav.astype(np.float64) * 32768.0, dtype=np.float64)
    sig = f64 >= f64.max(axis=-1, keepdims=True) + np.log(1e-7)
    np.testing.assert_allclose(out[sig], ref[sig], atol=1e-3, rtol=1e-4)
    np.testing.assert_allclose(out[sig], f64[sig], atol=1e-3, rtol=1e-4)
    # leakage-floor bins: fp32 rounding noise in the reference too; ours must not be further from f64 than 3x
    assert np.abs(out - f64).max() <= 3.0 * REPORT["fbank/" + name]["ref_vs_f64"]["max_abs"] + 1e-3


def test_preprocess_with_normalisation(fbank):
    wav, _ = cases.fbank_cases()["randn42_4x1s"]
    out = fbank.run(torch.from_numpy(wav).cuda(), prescale=32768.0, norm_mean=15.41663, norm_std2=2 * 6.55582)
    np.testing.assert_allclose(out.cpu().numpy(), FB["randn42_4x1s__pre"], atol=2e-4, rtol=1e-4)


def test_frame_counts_and_edge_shapes(fbank):
    for n in (400, 401, 559, 560, 4000, 16123):
        x = torch.randn(3, n, device="cuda")
        assert fbank(x).shape == (3, 1 + (n - 400) // 160, 128)
    with pytest.raises(RuntimeError):
        fbank(torch.randn(1, 399, device="cuda"))
    # non-contiguous rows / unaligned base take the scalar load path and agree with the vector path
    big = torch.randn(4, 16001, device="cuda")
    a = fbank(big[:, 1:])
    b = fbank(big[:, 1:].contiguous())
    assert torch.equal(a, b)


def test_bf16_output(fbank):
    x = torch.randn(2, 16000, device="cuda") * 3000
    a = fbank.run(x)
    b = fbank.run(x, out_dtype=torch.bfloat16)
    assert torch.equal(a.to(torch.bfloat16), b)


def test_eat_variant():
    from avex_b200.fbank import KaldiFbank

    g = np.load(os.path.join(G, "eat_fbank.npz"))
    fb = KaldiFbank(window_type="hanning").cuda()
    wav = torch.from_numpy(cases.eat_case()).cuda()
    const = fb.run(wav, norm_mean=-4.268, norm_std2=2 * 4.569, out_frames=1024).cpu().numpy()
    np.testing.assert_allclose(const, g["const"], atol=1e-3, rtol=1e-4)
    per = fb.run(wav, out_frames=1024, per_utterance=True).cpu().numpy()
    np.testing.assert_allclose(per, g["perutt"], atol=2e-3, rtol=1e-3)


def test_full_size_properties(fbank):
    """BASELINE config #2 size (256 x 10 s): batch independence and shift consistency (size-independent)."""
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn(256, 160000, device="cuda", generator=g) * 0.1
    full = fbank.run(x, prescale=32768.0, norm_mean=15.41663, norm_std2=2 * 6.55582)
    assert full.shape == (256, 998, 128) and torch.isfinite(full).all()
    for i in (0, 77, 255):
        assert torch.equal(full[i], fbank.run(x[i : i + 1], prescale=32768.0, norm_mean=15.41663, norm_std2=2 * 6.55582)[0])
    # a clip shifted by k hops yields the same frames shifted by k (frames only see their own 400 samples)
    sh = fbank.run(x[:4, 160 * 5 :], prescale=32768.0, norm_mean=15.41663, norm_std2=2 * 6.55582)
    assert torch.equal(sh, full[:4, 5:])
    # oracle on a slice of the big batch
    orc = OF.beats_preprocess(x[3:4, :32000].cpu().numpy())
    got = full[3, : orc.shape[1]].cpu().numpy()
    assert _stats(got, orc[0])[1] <= 1e-5


@pytest.mark.parametrize("B,T", [(3, 16000), (2, 80000), (1, 160000), (2, 16 * 160 + 400 - 1), (5, 4000)])
def test_patch_operand_equals_split_of_fp32_fbank(B, T):
    """`avexk_fbank_patch_operand` (what the BEATs forward consumes) must be bit-identical to im2col + [hi|lo|hi] bf16 split of the
    fp32 fbank the same kernel writes in its plain mode: row = b*N + tp*8 + fp, col = i*16 + j <-> fbank[b, tp*16+i, fp*16+j]."""
    from avex_b200.fbank import KaldiFbank

    fb = KaldiFbank().cuda()
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T)
    wav = torch.randn(B, T, device="cuda", generator=g) * 0.1
    kw = dict(prescale=32768.0, norm_mean=15.41663, norm_std2=2 * 6.55582)
    full = fb.run(wav, **kw)
    F = full.shape[1]
    tp = F // 16
    got = fb.patch_operand(wav, **kw)
    assert got.shape == (B * tp * 8, 768)
    if tp == 0:
        return
    patches = full[:, : tp * 16].reshape(B, tp, 16, 8, 16).permute(0, 1, 3, 2, 4).reshape(B * tp * 8, 256)
    hi = patches.to(torch.bfloat16)
    lo = (patches - hi.float()).to(torch.bfloat16)
    want = torch.cat([hi, lo, hi], dim=1)
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
