"""Oracle, torch-CPU flavour: the same restatement as beats_encoder.py but on torch CPU ops, i.e. the op mix the
reference itself executes on a CPU (conv2d patch embed, F.linear -> addmm, grouped conv1d, a materialised
[B,H,N,N] gate*bias mask handed to scaled_dot_product_attention; backbone.py:544-568).  It exists so that bench.py's
CPU baseline / `--impl reference` arm is timed on the SAME libraries (MKL / oneDNN, all host threads) the reference
would use, instead of a slower numpy + OpenBLAS port.  Test infrastructure -- see oracle/__init__.py; validated
against the reference goldens in tests/test_oracle_vs_golden.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .beats_encoder import BeatsDims, token_padding_mask
from .relpos import bias_vector


def to_torch(W: dict) -> dict:
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in W.items()}


def fbank(wav: torch.Tensor, window: torch.Tensor, mel_fb: torch.Tensor) -> torch.Tensor:
    """beats.py:120-163 on torch CPU ops."""
    fr = wav.unfold(-1, 400, 160)
    fr = fr - fr.mean(dim=-1, keepdim=True)
    sh = F.pad(fr, (1, 0), mode="replicate")[..., :-1]
    fr = (fr - 0.97 * sh) * window
    fr = F.pad(fr, (0, 112))
    power = torch.fft.rfft(fr).abs().pow(2.0)
    return torch.clamp(power @ mel_fb, min=torch.finfo(torch.float32).eps).log()


@torch.no_grad()
def beats_forward(Wt: dict, wav: torch.Tensor, padding_mask=None, dims: BeatsDims = BeatsDims(), tables=None) -> dict:
    from . import kaldi_fbank as OF

    if tables is None:
        tables = (torch.from_numpy(OF.povey_window()).to(wav.device), torch.from_numpy(OF.mel_filterbank()).to(wav.device))
    g = lambda n: Wt[n]  # noqa: E731
    fb = fbank(wav.float() * 2**15, *tables)
    fb = (fb - dims.fbank_mean) / (2 * dims.fbank_std)  # beats.py:323
    key_pad = None
    if padding_mask is not None:
        key_pad = torch.from_numpy(token_padding_mask(np.asarray(padding_mask.cpu() if torch.is_tensor(padding_mask) else padding_mask, dtype=bool), dims)).to(wav.device)
    x = F.conv2d(fb.unsqueeze(1), g("backbone.patch_embedding.weight"), stride=16)  # beats.py:349-350
    x = x.reshape(x.shape[0], x.shape[1], -1).transpose(1, 2)
    x = F.layer_norm(x, (dims.patch_embed,), g("backbone.layer_norm.weight"), g("backbone.layer_norm.bias"))
    x = F.linear(x, g("backbone.post_extract_proj.weight"), g("backbone.post_extract_proj.bias"))
    if key_pad is not None:
        x[key_pad] = 0
    out = {"hook0": x.clone(), "fc2": []}
    B, N, C = x.shape
    gw = g("backbone.encoder.pos_conv.0.parametrizations.weight.original0")
    v = g("backbone.encoder.pos_conv.0.parametrizations.weight.original1")
    w = gw * v / v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
    xc = F.conv1d(x.transpose(1, 2), w, g("backbone.encoder.pos_conv.0.bias"), padding=dims.conv_pos // 2, groups=dims.conv_groups)
    xc = F.gelu(xc[:, :, :-1]).transpose(1, 2)  # SamePad + GELU
    x = x + xc
    x = F.layer_norm(x, (C,), g("backbone.encoder.layer_norm.weight"), g("backbone.encoder.layer_norm.bias"))
    H, d = dims.heads, C // dims.heads
    table = g("backbone.encoder.layers.0.self_attn.relative_attention_bias.weight")
    dev = x.device  # CPU for the baseline; tools/bench_eager_gpu.py runs the same op mix on the GPU
    bv = torch.from_numpy(bias_vector(table.cpu().numpy(), N, dims.num_buckets, dims.max_distance)).to(dev)
    ar = torch.arange(N, device=dev)
    idx = (ar[None, :] - ar[:, None]) + (N - 1)
    pos_bias = bv[:, idx].unsqueeze(0).expand(B, -1, -1, -1)  # [B,H,N,N], backbone.py:526-528
    alpha = dims.alpha
    for li in range(dims.layers):
        p = f"backbone.encoder.layers.{li}"
        sa = p + ".self_attn"
        q = F.linear(x, g(sa + ".q_proj.weight"), g(sa + ".q_proj.bias")).view(B, N, H, d).permute(0, 2, 1, 3)
        k = F.linear(x, g(sa + ".k_proj.weight"), g(sa + ".k_proj.bias")).view(B, N, H, d).permute(0, 2, 1, 3)
        vv = F.linear(x, g(sa + ".v_proj.weight"), g(sa + ".v_proj.bias")).view(B, N, H, d).permute(0, 2, 1, 3)
        ga, gb = torch.sigmoid(F.linear(q, g(sa + ".grep_linear.weight"), g(sa + ".grep_linear.bias")).view(B, H, N, 2, 4).sum(-1)).chunk(2, dim=-1)
        mask = (ga * (gb * g(sa + ".grep_a") - 1.0) + 2.0) * pos_bias  # materialised, like the reference (backbone.py:551)
        if key_pad is not None:
            pm = torch.zeros(B, 1, 1, N, device=dev)
            pm.masked_fill_(key_pad[:, None, None, :], float("-inf"))
            mask = mask + pm
        a = F.scaled_dot_product_attention(q, k, vv, attn_mask=mask, scale=d**-0.5)
        a = a.permute(0, 2, 1, 3).contiguous().view(B, N, C)
        a = F.linear(a, g(sa + ".out_proj.weight"), g(sa + ".out_proj.bias"))
        x = F.layer_norm(x * alpha + a, (C,), g(p + ".self_attn_layer_norm.weight"), g(p + ".self_attn_layer_norm.bias"))
        f2 = F.linear(F.gelu(F.linear(x, g(p + ".fc1.weight"), g(p + ".fc1.bias"))), g(p + ".fc2.weight"), g(p + ".fc2.bias"))
        out["fc2"].append(f2)
        x = F.layer_norm(x * alpha + f2, (C,), g(p + ".final_layer_norm.weight"), g(p + ".final_layer_norm.bias"))
    out["x"] = x
    out["key_pad"] = key_pad
    return out
