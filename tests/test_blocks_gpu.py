"""GPU: every building block of the BEATs path through the C ABI, against fp32 torch / the numpy oracle."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import beats_encoder as OE
from oracle import relpos as OR
from oracle.weights import make_beats_weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from avex_b200 import _lib

    return _lib.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(rc, lib):
    assert rc == 0, lib.avexk_last_error().decode()


GEMM_SHAPES = [
    (128, 256, 64), (128, 256, 256), (77, 512, 256), (300, 768, 512), (1000, 2304, 768),
    (513, 768, 3072), (4096, 3072, 768), (129, 48 * 16, 768), (20000, 768, 768),
]  # fmt: skip


@pytest.fixture(params=[1, 0], ids=["cta_pair", "single_cta"])
def pair(request, lib):
    """Both GEMM kernels: CTA pairs (tcgen05 cta_group::2, the default) and single-CTA tiles."""
    prev = lib.avexk_gemm_config(request.param)
    yield request.param
    lib.avexk_gemm_config(prev)


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_plain_fp32_out(lib, pair, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda")
    _check(lib.avexk_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), 0, None, None, 0.0,
                               out.data_ptr(), N, 0, _stream()), lib)  # fmt: skip
    ref = A.float() @ W.float().T + bias
    err = (out - ref).abs().max().item()
    assert torch.isfinite(out).all()
    assert err <= 2e-3, f"max abs err {err}"


def test_gemm_epilogues(lib, pair):
    M, N, K = 777, 768, 768
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    ref = A.float() @ W.float().T + bias
    # bias + raw store + scaled residual, fp32 out (out_proj / fc2 epilogue)
    out = torch.empty(M, N, device="cuda")
    raw = torch.empty(M, N, device="cuda")
    _check(lib.avexk_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), 0, raw.data_ptr(), res.data_ptr(),
                               2.2133638, out.data_ptr(), N, 0, _stream()), lib)  # fmt: skip
    assert (raw - ref).abs().max().item() <= 2e-3
    assert (out - (ref + 2.2133638 * res)).abs().max().item() <= 3e-3
    # bias + exact GELU, bf16 out (fc1 epilogue)
    outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    _check(lib.avexk_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), 1, None, None, 0.0,
                               outb.data_ptr(), N, 1, _stream()), lib)  # fmt: skip
    gref = torch.nn.functional.gelu(ref)
    assert (outb.float() - gref).abs().max().item() <= 2e-2
    assert ((outb.float() - gref).abs() <= 8e-3 * gref.abs() + 2e-3).all()
    # no bias, bf16 out
    _check(lib.avexk_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, M, N, K, None, 0, None, None, 0.0,
                               outb.data_ptr(), N, 1, _stream()), lib)  # fmt: skip
    assert ((outb.float() - (ref - bias)).abs() <= 8e-3 * (ref - bias).abs() + 2e-3).all()


@pytest.mark.parametrize("M,K,raw,inplace", [(1, 768, False, False), (100, 768, True, False), (129, 3072, False, True),
                                              (1000, 768, True, True), (40000, 768, False, True), (38017, 3072, True, False)])
def test_gemm_fused_layernorm(lib, pair, M, K, raw, inplace):
    """out_proj / fc2 tail in one launch: LN(A W^T + b + alpha * residual), optional raw (hook) store, in-place residual."""
    N, alpha = 768, 2.2133638
    g = torch.Generator(device="cuda").manual_seed(M + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g) + 0.3
    gamma = torch.randn(N, device="cuda", generator=g)
    beta = torch.randn(N, device="cuda", generator=g)
    lin = A.float() @ W.float().T + bias
    ref = torch.nn.functional.layer_norm(lin + alpha * res, (N,), gamma, beta, 1e-5)
    nbytes = lib.avexk_gemm_ln_scratch_bytes(M)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    res_buf = res.clone()
    o32 = res_buf if inplace else torch.full((M, N), float("nan"), device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    rawt = torch.full((M, N), float("nan"), device="cuda") if raw else None
    for _ in range(1 if inplace else 2):  # second pass over a dirty scratch
        _check(lib.avexk_gemm_bf16_ln(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), rawt.data_ptr() if raw else None,
                                      res_buf.data_ptr(), alpha, gamma.data_ptr(), beta.data_ptr(), 1e-5, o32.data_ptr(), o16.data_ptr(),
                                      scratch.data_ptr(), nbytes, _stream()), lib)  # fmt: skip
    assert torch.isfinite(o32).all()
    assert (o32 - ref).abs().max().item() <= 3e-3
    assert torch.equal(o16, o32.to(torch.bfloat16))
    if raw:
        assert (rawt - lin).abs().max().item() <= 2e-3
    # bf16-only output (the last layer keeps no fp32 copy when only xb is wanted) and scratch-size check
    o16b = torch.empty_like(o16)
    _check(lib.avexk_gemm_bf16_ln(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), None, res.data_ptr(), alpha,
                                  gamma.data_ptr(), beta.data_ptr(), 1e-5, None, o16b.data_ptr(), scratch.data_ptr(), nbytes, _stream()), lib)  # fmt: skip
    assert torch.equal(o16b, o16)
    assert lib.avexk_gemm_bf16_ln(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), None, res.data_ptr(), alpha, gamma.data_ptr(),
                                  beta.data_ptr(), 1e-5, None, o16b.data_ptr(), scratch.data_ptr(), 16, _stream()) != 0  # fmt: skip


@pytest.mark.parametrize("clips,rows,K,want_y", [(4, 496, 768, True), (3, 40, 3072, False), (7, 2992, 3072, True), (33, 248, 768, False),
                                                 (1, 32, 768, True)])
def test_gemm_fused_layernorm_pooled(lib, pair, clips, rows, K, want_y):
    """Mean-pooling over the token rows of every clip as a by-product of the fc2 / out_proj epilogue: pooled_raw = mean of the raw
    Linear output (a mean-aggregated hook), pooled_y = mean of the LayerNorm output; clips of `rows` rows straddle the 32-row boxes
    and the 128 / 256-row tiles.  Fixed-point accumulation: two runs are bit-identical."""
    N, alpha, M = 768, 2.2133638, clips * rows
    g = torch.Generator(device="cuda").manual_seed(M + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g) + 0.3
    gamma = torch.randn(N, device="cuda", generator=g)
    beta = torch.randn(N, device="cuda", generator=g)
    lin = A.float() @ W.float().T + bias
    ref = torch.nn.functional.layer_norm(lin + alpha * res, (N,), gamma, beta, 1e-5)
    nbytes = lib.avexk_gemm_ln_scratch_bytes(M)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    pool_ws = torch.empty(2 * clips * N * 8, dtype=torch.uint8, device="cuda")
    outs = []
    for rep in range(2):
        p_raw = torch.full((clips, N), float("nan"), device="cuda")
        p_y = torch.full((clips, N), float("nan"), device="cuda") if want_y else None
        o32 = torch.full((M, N), float("nan"), device="cuda") if rep == 0 else None  # second run: pooled outputs only
        _check(lib.avexk_gemm_bf16_ln_pooled(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), None, res.data_ptr(), alpha,
                                             gamma.data_ptr(), beta.data_ptr(), 1e-5, o32.data_ptr() if o32 is not None else None, None,
                                             scratch.data_ptr(), nbytes, rows, p_raw.data_ptr(), p_y.data_ptr() if want_y else None,
                                             pool_ws.data_ptr(), _stream()), lib)  # fmt: skip
        outs.append((p_raw, p_y))
        if o32 is not None:
            assert (o32 - ref).abs().max().item() <= 3e-3
    p_raw, p_y = outs[0]
    assert (p_raw - lin.view(clips, rows, N).mean(1)).abs().max().item() <= 5e-4
    if want_y:
        assert (p_y - ref.view(clips, rows, N).mean(1)).abs().max().item() <= 5e-4
        assert torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][0], outs[1][0])
    # fewer than 32 rows per clip cannot be pooled in the epilogue: loud refusal
    assert lib.avexk_gemm_bf16_ln_pooled(A.data_ptr(), K, W.data_ptr(), K, 16, N, K, bias.data_ptr(), None, res.data_ptr(), alpha,
                                         gamma.data_ptr(), beta.data_ptr(), 1e-5, None, None, scratch.data_ptr(), nbytes, 16,
                                         p_raw.data_ptr(), None, pool_ws.data_ptr(), _stream()) != 0  # fmt: skip


def test_gemm_rejects_bad_shapes(lib):
    A = torch.zeros(8, 100, device="cuda", dtype=torch.bfloat16)
    assert lib.avexk_gemm_bf16(A.data_ptr(), 100, A.data_ptr(), 100, 8, 16, 100, None, 0, None, None, 0.0, A.data_ptr(), 16, 1, _stream()) != 0
    assert b"unsupported shape" in lib.avexk_last_error()


@pytest.mark.parametrize("M,Cw", [(1, 768), (9, 512), (1000, 768), (333, 1024), (64, 128)])
def test_layernorm(lib, M, Cw):
    g = torch.Generator(device="cuda").manual_seed(M)
    x = torch.randn(M, Cw, device="cuda", generator=g) * 3 + 1
    w = torch.randn(Cw, device="cuda", generator=g)
    b = torch.randn(Cw, device="cuda", generator=g)
    o32 = torch.empty_like(x)
    o16 = torch.empty(M, Cw, device="cuda", dtype=torch.bfloat16)
    _check(lib.avexk_layernorm(x.data_ptr(), M, Cw, w.data_ptr(), b.data_ptr(), 1e-5, o32.data_ptr(), o16.data_ptr(), _stream()), lib)
    ref = torch.nn.functional.layer_norm(x, (Cw,), w, b, 1e-5)
    assert (o32 - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()) * 4
    assert torch.equal(o16, o32.to(torch.bfloat16))


def _attn_ref(qkv, B, N, H, gw, gb, ga, bias_vec, key_pad):
    d = 64
    q, k, v = [t.reshape(B, N, H, d).permute(0, 2, 1, 3).float() for t in qkv.split(H * d, dim=1)]
    gl = q @ gw.T + gb  # [B,H,N,2]
    gate = torch.sigmoid(gl)
    g1 = gate[..., 0:1] * (gate[..., 1:2] * ga.view(1, H, 1, 1) - 1.0) + 2.0
    idx = (torch.arange(N, device="cuda")[None, :] - torch.arange(N, device="cuda")[:, None]) + N - 1
    s = q @ k.transpose(-1, -2) * 0.125 + g1 * bias_vec[:, idx][None]
    if key_pad is not None:
        s = s.masked_fill(key_pad[:, None, None, :].bool(), float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * N, H * d)


@pytest.mark.parametrize("B,N,H,pad", [(2, 48, 12, False), (1, 248, 12, False), (3, 96, 4, True), (2, 496, 2, False), (1, 700, 1, True),
                                         (1, 128, 1, False), (2, 256, 3, False), (1, 2992, 2, False), (40, 496, 12, False), (2, 300, 2, True)])
def test_attention_gated(lib, B, N, H, pad):
    g = torch.Generator(device="cuda").manual_seed(N)
    qkv = (torch.randn(B * N, 3 * H * 64, device="cuda", generator=g)).to(torch.bfloat16)
    gw = torch.randn(2, 64, device="cuda", generator=g) * 0.2
    gb = torch.randn(2, device="cuda", generator=g) * 0.2
    ga = 1.0 + 0.2 * torch.randn(H, device="cuda", generator=g)
    table = torch.randn(320, H, generator=torch.Generator().manual_seed(1))
    bias_vec = torch.from_numpy(OR.bias_vector(table.numpy(), N)).cuda()
    key_pad = None
    if pad:
        key_pad = torch.zeros(B, N, dtype=torch.uint8, device="cuda")
        key_pad[-1, N // 2 + 3 :] = 1
    out = torch.empty(B * N, H * 64, device="cuda", dtype=torch.bfloat16)
    _check(lib.avexk_attention_gated(qkv.data_ptr(), B, N, H, gw.data_ptr(), gb.data_ptr(), ga.data_ptr(), bias_vec.data_ptr(),
                                     key_pad.data_ptr() if pad else None, out.data_ptr(), _stream()), lib)  # fmt: skip
    ref = _attn_ref(qkv, B, N, H, gw, gb, ga, bias_vec, key_pad)
    err = (out.float() - ref).abs().max().item()
    assert torch.isfinite(out.float()).all()
    assert err <= 2e-2, err  # P and the output are rounded to bf16
    assert torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item() >= 0.9995


def test_attention_leading_padding(lib):
    """Keys 0..139 of clip 0 are padded: the first 128-key tile is fully masked (the reference estimate is taken from the chunk of the first valid key)."""
    B, N, H = 2, 300, 3
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = (torch.randn(B * N, 3 * H * 64, device="cuda", generator=g) * 2.0).to(torch.bfloat16)
    gw = torch.randn(2, 64, device="cuda", generator=g) * 0.2
    gb = torch.randn(2, device="cuda", generator=g) * 0.2
    ga = 1.0 + 0.2 * torch.randn(H, device="cuda", generator=g)
    table = torch.randn(320, H, generator=torch.Generator().manual_seed(2)) * 3.0
    bias_vec = torch.from_numpy(OR.bias_vector(table.numpy(), N)).cuda()
    key_pad = torch.zeros(B, N, dtype=torch.uint8, device="cuda")
    key_pad[0, :140] = 1
    key_pad[1, 200:] = 1
    out = torch.empty(B * N, H * 64, device="cuda", dtype=torch.bfloat16)
    _check(lib.avexk_attention_gated(qkv.data_ptr(), B, N, H, gw.data_ptr(), gb.data_ptr(), ga.data_ptr(), bias_vec.data_ptr(),
                                     key_pad.data_ptr(), out.data_ptr(), _stream()), lib)  # fmt: skip
    ref = _attn_ref(qkv, B, N, H, gw, gb, ga, bias_vec, key_pad)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    assert err <= 4e-2, err  # |v| ~ 2, sharp softmax
    assert torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item() >= 0.9995


def test_attention_large_score_growth(lib):
    """Scores that grow along the key axis force the lazy reference max to move on every tile (rescale path)."""
    B, N, H = 1, 640, 2
    g = torch.Generator(device="cuda").manual_seed(9)
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda", generator=g)
    ramp = torch.linspace(0.2, 6.0, N, device="cuda")[:, None]
    qkv[:, H * 64 : 2 * H * 64] *= ramp  # keys get larger with j
    qkv = qkv.to(torch.bfloat16)
    gw = torch.randn(2, 64, device="cuda", generator=g) * 0.2
    gb = torch.randn(2, device="cuda", generator=g) * 0.2
    ga = 1.0 + 0.2 * torch.randn(H, device="cuda", generator=g)
    table = torch.randn(320, H, generator=torch.Generator().manual_seed(3))
    bias_vec = torch.from_numpy(OR.bias_vector(table.numpy(), N)).cuda()
    out = torch.empty(B * N, H * 64, device="cuda", dtype=torch.bfloat16)
    _check(lib.avexk_attention_gated(qkv.data_ptr(), B, N, H, gw.data_ptr(), gb.data_ptr(), ga.data_ptr(), bias_vec.data_ptr(),
                                     None, out.data_ptr(), _stream()), lib)  # fmt: skip
    ref = _attn_ref(qkv, B, N, H, gw, gb, ga, bias_vec, None)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    assert err <= 3e-2, err
    assert torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item() >= 0.9995


def test_attention_reference_lags_row_max(lib):
    """The softmax reference of a row is estimated from its first 16 valid keys; here every later key scores far higher
    (row max ~ 2^50 above the estimate), so P is formed against a lagging reference and rescaled after the first tile."""
    B, N, H = 2, 496, 2
    g = torch.Generator(device="cuda").manual_seed(21)
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda", generator=g)
    k = qkv[:, H * 64 : 2 * H * 64].view(B, N, H * 64)
    k[:, 16:] *= 12.0
    qkv = qkv.to(torch.bfloat16)
    gw = torch.randn(2, 64, device="cuda", generator=g) * 0.2
    gb = torch.randn(2, device="cuda", generator=g) * 0.2
    ga = 1.0 + 0.2 * torch.randn(H, device="cuda", generator=g)
    table = torch.randn(320, H, generator=torch.Generator().manual_seed(4))
    bias_vec = torch.from_numpy(OR.bias_vector(table.numpy(), N)).cuda()
    for key_pad in (None, torch.zeros(B, N, dtype=torch.uint8, device="cuda")):
        if key_pad is not None:
            key_pad[0, :5] = 1  # the estimate chunk starts with masked keys
            key_pad[1, 400:] = 1
        out = torch.empty(B * N, H * 64, device="cuda", dtype=torch.bfloat16)
        _check(lib.avexk_attention_gated(qkv.data_ptr(), B, N, H, gw.data_ptr(), gb.data_ptr(), ga.data_ptr(), bias_vec.data_ptr(),
                                         key_pad.data_ptr() if key_pad is not None else None, out.data_ptr(), _stream()), lib)  # fmt: skip
        ref = _attn_ref(qkv, B, N, H, gw, gb, ga, bias_vec, key_pad)
        assert torch.isfinite(out.float()).all()
        err = (out.float() - ref).abs().max().item()
        assert err <= 3e-2, err
        assert torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item() >= 0.9995


def test_attention_random_shapes_and_masks(lib):
    """Seeded sweep over ragged shapes and mask patterns (none / suffix / prefix / scattered / a fully padded clip): every
    tile-boundary case of the kernel -- single-tile pairs, a second query tile with a few rows, first valid key in any
    16-key chunk (own-half, shared and safe-chunk estimates), masks that leave whole tiles dead."""
    rng = np.random.RandomState(2024)
    worst = 0.0
    for case in range(48):
        N = int(rng.choice([1, 7, 16, 17, 63, 64, 65, 120, 128, 129, 130, 200, 255, 256, 257, 300, 383, 384, 385, 500, 513, 640]))
        B, H = int(rng.randint(1, 4)), int(rng.randint(1, 4))
        kind = ["none", "suffix", "prefix", "scatter", "deadclip"][case % 5]
        g = torch.Generator(device="cuda").manual_seed(1000 + case)
        qkv = (torch.randn(B * N, 3 * H * 64, device="cuda", generator=g) * (1.0 + (case % 3))).to(torch.bfloat16)
        gw = torch.randn(2, 64, device="cuda", generator=g) * 0.2
        gb = torch.randn(2, device="cuda", generator=g) * 0.2
        ga = 1.0 + 0.2 * torch.randn(H, device="cuda", generator=g)
        table = torch.randn(320, H, generator=torch.Generator().manual_seed(case)) * 2.0
        bias_vec = torch.from_numpy(OR.bias_vector(table.numpy(), N)).cuda()
        key_pad = None
        if kind != "none":
            kp = np.zeros((B, N), np.uint8)
            for b in range(B):
                if kind == "suffix":
                    kp[b, rng.randint(1, N + 1):] = 1
                elif kind == "prefix":
                    kp[b, : rng.randint(0, N)] = 1
                elif kind == "scatter":
                    kp[b] = rng.rand(N) < 0.6
                    kp[b, rng.randint(0, N)] = 0  # at least one valid key
            if kind == "deadclip":
                kp[0, :] = 1
            key_pad = torch.from_numpy(kp).cuda()
        out = torch.full((B * N, H * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        _check(lib.avexk_attention_gated(qkv.data_ptr(), B, N, H, gw.data_ptr(), gb.data_ptr(), ga.data_ptr(), bias_vec.data_ptr(),
                                         key_pad.data_ptr() if key_pad is not None else None, out.data_ptr(), _stream()), lib)  # fmt: skip
        ref = _attn_ref(qkv, B, N, H, gw, gb, ga, bias_vec, key_pad)
        o = out.float()
        assert torch.isfinite(o).all(), (case, N, B, H, kind)
        live = torch.ones(B, dtype=torch.bool, device="cuda") if key_pad is None else ~(key_pad.bool().all(dim=1))
        rows = live[:, None].expand(B, N).reshape(-1)
        if (~rows).any():  # a clip without a valid key: the reference softmax is NaN there, the kernel returns zeros
            assert (o[~rows] == 0).all(), (case, N, kind)
        err = (o[rows] - ref[rows]).abs().max().item() if rows.any() else 0.0
        scale = max(1.0, ref[rows].abs().max().item()) if rows.any() else 1.0
        assert err <= 2e-2 * scale, (case, N, B, H, kind, err)
        worst = max(worst, err / scale)
    print("attention sweep: worst relative max-abs error %.3e" % worst)


@pytest.mark.parametrize("B,N,pad", [(2, 48, False), (1, 248, False), (2, 131, True), (3, 496, False)])
def test_posconv(lib, B, N, pad):
    dims = OE.BeatsDims()
    W = make_beats_weights(OE.BeatsDims(layers=1), seed=7)
    Cw = 768
    g = torch.Generator(device="cuda").manual_seed(N)
    x0 = torch.randn(B, N, Cw, device="cuda", generator=g)
    key_pad = None
    if pad:
        key_pad = torch.zeros(B, N, dtype=torch.uint8, device="cuda")
        key_pad[0, N - 20 :] = 1
    xin = x0.clone()
    x_ref = x0.cpu().numpy().copy()
    if pad:
        x_ref[key_pad.cpu().numpy().astype(bool)] = 0
    ref = x_ref + OE.pos_conv(x_ref, W, dims)
    wg = torch.from_numpy(W["backbone.encoder.pos_conv.0.parametrizations.weight.original0"]).cuda().contiguous()
    wv = torch.from_numpy(W["backbone.encoder.pos_conv.0.parametrizations.weight.original1"]).cuda().contiguous()
    bias = torch.from_numpy(W["backbone.encoder.pos_conv.0.bias"]).cuda()
    need = lib.avexk_posconv_workspace_bytes(B, N, Cw, 16, 128)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    out = torch.empty_like(x0)
    _check(lib.avexk_posconv(xin.data_ptr(), B, N, Cw, 16, 128, wg.data_ptr(), wv.data_ptr(), bias.data_ptr(),
                             key_pad.data_ptr() if pad else None, out.data_ptr(), ws.data_ptr(), need, _stream()), lib)  # fmt: skip
    got = out.cpu().numpy()
    conv_scale = np.abs(ref - x_ref).max()
    err = np.abs(got - ref).max()
    print(f"posconv B={B} N={N}: max err {err:.3e} (conv magnitude {conv_scale:.3f})")
    assert err <= 2e-2 * max(1.0, conv_scale)
    if pad:
        assert (xin[key_pad.bool()] == 0).all()
