"""Oracle: BEATs encoder forward (numpy, fp32 or fp64).  Test infrastructure -- see oracle/__init__.py.

Restates, for eval mode (all dropouts / layerdrop inert) and the post-LN DeepNorm configuration of
avex/api/configs/official_models/esp_aves2_sl_beats_all.yml:
  avex/models/beats/beats.py:325-382        BEATs.extract_features (front end)
  avex/models/beats/beats.py:283-302        forward_padding_mask
  avex/models/beats/backbone.py:151-221     TransformerEncoder.extract_features
  avex/models/beats/backbone.py:350-373     _TransformerSentenceEncoderLayer.forward (post-LN branch)
  avex/models/beats/backbone.py:494-574     _MultiheadAttention.forward (gated rel-pos bias + SDPA)
  avex/models/beats_model.py:232-277        Model.forward (features-only / masked mean-pool + classifier)
Weights are a dict keyed by the reference `state_dict()` names (prefix "backbone.").
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

try:  # exact erf for GELU (modules.py:191-200 / nn.GELU())
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

from . import kaldi_fbank, relpos


@dataclass
class BeatsDims:
    """Architecture hyper-parameters (BEATsConfig, beats.py:166-228)."""

    layers: int = 12
    embed: int = 768
    ffn: int = 3072
    heads: int = 12
    patch: int = 16
    patch_embed: int = 512
    conv_pos: int = 128
    conv_groups: int = 16
    num_buckets: int = 320
    max_distance: int = 800
    n_mels: int = 128
    fbank_mean: float = 15.41663
    fbank_std: float = 6.55582

    @property
    def alpha(self) -> float:
        return math.pow(2 * self.layers, 0.25)  # backbone.py:306


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * w + b


def gelu(x):
    return (0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))).astype(x.dtype)


def n_tokens(num_samples: int, dims: BeatsDims = BeatsDims()) -> int:
    F = kaldi_fbank.frame_count(num_samples)
    return (F // dims.patch) * (dims.n_mels // dims.patch)


def forward_padding_mask(n_feat: int, mask: np.ndarray) -> np.ndarray:
    """beats.py:283-302: trim `T mod n_feat`, view [B, n_feat, -1], all(-1)."""
    extra = mask.shape[1] % n_feat
    if extra > 0:
        mask = mask[:, :-extra]
    return mask.reshape(mask.shape[0], n_feat, -1).all(-1)


def token_padding_mask(sample_mask: np.ndarray, dims: BeatsDims = BeatsDims()) -> np.ndarray:
    """Sample mask [B,T] -> frame mask [B,F] -> token mask [B,N] (beats.py:346-347, :355-356)."""
    T = sample_mask.shape[1]
    F = kaldi_fbank.frame_count(T)
    m = forward_padding_mask(F, np.asarray(sample_mask, dtype=bool))
    return forward_padding_mask(n_tokens(T, dims), m)


def pos_conv_weight(W: dict, dtype) -> np.ndarray:
    """weight_norm(dim=2): w = g * v / ||v||, norm over dims (0,1); backbone.py:67."""
    g = W["backbone.encoder.pos_conv.0.parametrizations.weight.original0"].astype(np.float64)
    v = W["backbone.encoder.pos_conv.0.parametrizations.weight.original1"].astype(np.float64)
    nrm = np.sqrt((v**2).sum(axis=(0, 1), keepdims=True))
    return (g * v / nrm).astype(dtype)


def patchify(fb: np.ndarray, patch: int = 16) -> np.ndarray:
    """Conv2d(1,E,16,16,stride 16) as im2col: [B,F,128] -> [B, N, 256], token = t*8 + f (beats.py:349-352)."""
    B, F, M = fb.shape
    Tp, Fp = F // patch, M // patch
    x = fb[:, : Tp * patch, : Fp * patch].reshape(B, Tp, patch, Fp, patch)
    return x.transpose(0, 1, 3, 2, 4).reshape(B, Tp * Fp, patch * patch)


def pos_conv(x: np.ndarray, W: dict, dims: BeatsDims) -> np.ndarray:
    """Conv1d(C,C,k=128,pad=64,groups=16) over tokens, SamePad drops the last output, GELU.

    backbone.py:52-68,172-173; modules.py:67-94.  x [B,N,C] -> [B,N,C].
    """
    B, N, C = x.shape
    K, G = dims.conv_pos, dims.conv_groups
    cg = C // G
    w = pos_conv_weight(W, x.dtype)  # [C, cg, K]
    bias = W["backbone.encoder.pos_conv.0.bias"].astype(x.dtype)
    xp = np.zeros((B, N + K, C), dtype=x.dtype)
    xp[:, K // 2 : K // 2 + N] = x
    out = np.empty((B, N, C), dtype=x.dtype)
    for g in range(G):
        xg = xp[:, :, g * cg : (g + 1) * cg]  # [B, N+K, cg]
        win = np.lib.stride_tricks.sliding_window_view(xg, K, axis=1)[:, :N]  # [B, N, cg, K]
        wg = w[g * cg : (g + 1) * cg].reshape(cg, cg * K)  # [co, (ci,t)]
        out[:, :, g * cg : (g + 1) * cg] = (win.reshape(B * N, cg * K) @ wg.T).reshape(B, N, cg)
    return gelu(out + bias)


def attention(x: np.ndarray, W: dict, pfx: str, bias_vec: np.ndarray, key_pad, dims: BeatsDims) -> np.ndarray:
    """backbone.py:494-574 on batch-first x [B,N,C]; returns out_proj(attn)."""
    B, N, C = x.shape
    H, d = dims.heads, C // dims.heads
    dt = x.dtype

    def lin(name, t):
        return t @ W[f"{pfx}.{name}.weight"].astype(dt).T + W[f"{pfx}.{name}.bias"].astype(dt)

    q = lin("q_proj", x).reshape(B, N, H, d).transpose(0, 2, 1, 3)  # [B,H,N,d]  backbone.py:531-538
    k = lin("k_proj", x).reshape(B, N, H, d).transpose(0, 2, 1, 3)
    v = lin("v_proj", x).reshape(B, N, H, d).transpose(0, 2, 1, 3)
    # gated relative position bias, from *unscaled* q; backbone.py:544-551
    gw = W[f"{pfx}.grep_linear.weight"].astype(dt)
    gb = W[f"{pfx}.grep_linear.bias"].astype(dt)
    gl = (q @ gw.T + gb).reshape(B, H, N, 2, 4).sum(-1)
    gate = 1.0 / (1.0 + np.exp(-gl))
    gate_a, gate_b = gate[..., 0:1], gate[..., 1:2]
    grep_a = W[f"{pfx}.grep_a"].astype(dt).reshape(1, H, 1, 1)
    gate_a_1 = gate_a * (gate_b * grep_a - 1.0) + 2.0  # [B,H,N,1]
    idx = (np.arange(N)[None, :] - np.arange(N)[:, None]) + (N - 1)  # j - i + N - 1
    pos_bias = bias_vec.astype(dt)[:, idx]  # [H,N,N]
    scores = (q @ k.transpose(0, 1, 3, 2)) * dt.type(d**-0.5) + gate_a_1 * pos_bias[None]
    if key_pad is not None:
        scores = np.where(key_pad[:, None, None, :], dt.type(-np.inf), scores)  # backbone.py:555-558
    scores = scores - scores.max(axis=-1, keepdims=True)
    p = np.exp(scores)
    p = p / p.sum(axis=-1, keepdims=True)
    o = (p @ v).transpose(0, 2, 1, 3).reshape(B, N, C)  # backbone.py:571
    return lin("out_proj", o)


def encoder_from_fbank(W: dict, fb: np.ndarray, key_pad=None, dims: BeatsDims = BeatsDims(), dtype=np.float32) -> dict:
    """Everything after `preprocess`: normalised fbank [B,F,128] -> dict of tensors.

    Returns {"hook0": post_extract_proj out [B,N,C], "fc2": [L x raw fc2 out [B,N,C]],
             "attn": [L x out_proj out], "x": final features [B,N,C], "posconv": encoder input after LN}.
    """
    dt = np.dtype(dtype)
    fb = np.asarray(fb, dtype=dt)
    Wc = W

    def g(name):
        return Wc[name].astype(dt)

    a = patchify(fb, dims.patch)
    x = a @ g("backbone.patch_embedding.weight").reshape(dims.patch_embed, -1).T  # beats.py:350-352
    x = layer_norm(x, g("backbone.layer_norm.weight"), g("backbone.layer_norm.bias"))  # beats.py:353
    x = x @ g("backbone.post_extract_proj.weight").T + g("backbone.post_extract_proj.bias")  # beats.py:359
    B, N, C = x.shape
    if key_pad is not None:
        # backbone.py:169-170 zeroes padded rows IN PLACE on the tensor post_extract_proj returned
        # (dropout_input is identity in eval, beats.py:361), so the tensor a forward hook captured on
        # `backbone.post_extract_proj` shows those zeros too -- observable reference behaviour, kept.
        x = np.where(key_pad[:, :, None], dt.type(0), x)
    out = {"hook0": x.copy(), "fc2": [], "attn": []}
    x = x + pos_conv(x, Wc, dims)  # backbone.py:172-174
    x = layer_norm(x, g("backbone.encoder.layer_norm.weight"), g("backbone.encoder.layer_norm.bias"))  # :176-177
    out["posconv"] = x.copy()
    table = g("backbone.encoder.layers.0.self_attn.relative_attention_bias.weight")
    bias_vec = relpos.bias_vector(table, N, dims.num_buckets, dims.max_distance)
    alpha = dt.type(dims.alpha)
    for li in range(dims.layers):
        p = f"backbone.encoder.layers.{li}"
        att = attention(x, Wc, f"{p}.self_attn", bias_vec, key_pad, dims)
        out["attn"].append(att)
        x = layer_norm(x * alpha + att, g(f"{p}.self_attn_layer_norm.weight"), g(f"{p}.self_attn_layer_norm.bias"))
        h = gelu(x @ g(f"{p}.fc1.weight").T + g(f"{p}.fc1.bias"))
        f2 = h @ g(f"{p}.fc2.weight").T + g(f"{p}.fc2.bias")
        out["fc2"].append(f2)
        x = layer_norm(x * alpha + f2, g(f"{p}.final_layer_norm.weight"), g(f"{p}.final_layer_norm.bias"))
    out["x"] = x
    return out


def beats_forward(W: dict, wav: np.ndarray, padding_mask=None, dims: BeatsDims = BeatsDims(), dtype=np.float32) -> dict:
    """Waveform [B,T] (+ optional sample padding mask [B,T], True = pad) -> encoder_from_fbank dict + "fbank"."""
    wav = np.asarray(wav)
    fb = kaldi_fbank.beats_preprocess(wav, dims.fbank_mean, dims.fbank_std, dtype=dtype)
    key_pad = None
    if padding_mask is not None:
        key_pad = token_padding_mask(np.asarray(padding_mask, dtype=bool), dims)
    out = encoder_from_fbank(W, fb, key_pad, dims, dtype)
    out["fbank"] = fb
    out["key_pad"] = key_pad
    return out


def mean_pool(x: np.ndarray, key_pad=None) -> np.ndarray:
    """beats_model.py:269-275: masked mean over tokens when any token is padded, else plain mean."""
    if key_pad is not None and key_pad.any():
        keep = (~key_pad)[:, :, None].astype(x.dtype)
        cnt = np.maximum(keep.sum(axis=1), 1)
        return (x * keep).sum(axis=1) / cnt
    return x.mean(axis=1)
