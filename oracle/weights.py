"""Deterministic synthetic BEATs weights (numpy).  Test infrastructure -- see oracle/__init__.py.

There is no network for checkpoints, so tests / bench use random-init weights of the named
architecture.  Keys and shapes are exactly the reference `beats_model.Model.state_dict()`
(dumped live), so the same dict loads into the reference module
(`load_state_dict`, used by tests/golden/make_golden.py) and into `avex_b200`'s model.

init="reference": the reference's init *distributions* (backbone.py:59-62,109-122,426-436,577-600;
                  torch defaults for patch_embedding / post_extract_proj): biases 0, LayerNorm (1,0).
init="perturbed": same weight scales, but non-zero biases, non-trivial LayerNorm affine, grep_a != 1 and
                  an O(1) relative-position table, so that every parameter influences the output and
                  indexing bugs are visible.  Used for the parity goldens.
"""
from __future__ import annotations

import math

import numpy as np

from .beats_encoder import BeatsDims


def make_beats_weights(dims: BeatsDims = BeatsDims(), seed: int = 1, init: str = "perturbed") -> dict:
    rs = np.random.RandomState(seed)
    pert = init == "perturbed"

    def normal(shape, std):
        return (rs.standard_normal(shape) * std).astype(np.float32)

    def uniform(shape, bound):
        return rs.uniform(-bound, bound, size=shape).astype(np.float32)

    def bias(n, std=0.02):
        return normal((n,), std) if pert else np.zeros((n,), np.float32)

    def ln(n):
        if pert:
            return (1.0 + normal((n,), 0.1)).astype(np.float32), normal((n,), 0.05)
        return np.ones((n,), np.float32), np.zeros((n,), np.float32)

    C, E, Ff, H = dims.embed, dims.patch_embed, dims.ffn, dims.heads
    d = C // H
    W: dict = {}
    p2 = dims.patch * dims.patch
    W["backbone.post_extract_proj.weight"] = uniform((C, E), 1.0 / math.sqrt(E))
    W["backbone.post_extract_proj.bias"] = uniform((C,), 1.0 / math.sqrt(E))
    W["backbone.patch_embedding.weight"] = uniform((E, 1, dims.patch, dims.patch), 1.0 / math.sqrt(p2))
    W["backbone.layer_norm.weight"], W["backbone.layer_norm.bias"] = ln(E)
    K, G = dims.conv_pos, dims.conv_groups
    v = normal((C, C // G, K), math.sqrt(4.0 / (K * C)))  # backbone.py:59-61
    W["backbone.encoder.pos_conv.0.bias"] = bias(C)
    g = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(0, 1), keepdims=True)).astype(np.float32)
    if pert:
        g = (g * (1.0 + normal((1, 1, K), 0.1))).astype(np.float32)
    W["backbone.encoder.pos_conv.0.parametrizations.weight.original0"] = g
    W["backbone.encoder.pos_conv.0.parametrizations.weight.original1"] = v
    W["backbone.encoder.layer_norm.weight"], W["backbone.encoder.layer_norm.bias"] = ln(C)
    beta = math.pow(8 * dims.layers, -0.25)  # backbone.py:112
    table = normal((dims.num_buckets, H), 1.0 if pert else 0.02)
    for li in range(dims.layers):
        p = f"backbone.encoder.layers.{li}"
        W[f"{p}.self_attn.grep_a"] = (
            (1.0 + normal((1, H, 1, 1), 0.2)) if pert else np.ones((1, H, 1, 1), np.float32)
        ).astype(np.float32)
        W[f"{p}.self_attn.relative_attention_bias.weight"] = table  # shared storage, backbone.py:100-103
        xav = math.sqrt(2.0 / (C + C))
        W[f"{p}.self_attn.k_proj.weight"] = normal((C, C), xav)
        W[f"{p}.self_attn.k_proj.bias"] = bias(C)
        W[f"{p}.self_attn.v_proj.weight"] = normal((C, C), xav * beta)
        W[f"{p}.self_attn.v_proj.bias"] = bias(C)
        W[f"{p}.self_attn.q_proj.weight"] = normal((C, C), xav)
        W[f"{p}.self_attn.q_proj.bias"] = bias(C)
        W[f"{p}.self_attn.out_proj.weight"] = normal((C, C), xav * beta)
        W[f"{p}.self_attn.out_proj.bias"] = bias(C)
        W[f"{p}.self_attn.grep_linear.weight"] = normal((8, d), 0.2 if pert else 0.02)
        W[f"{p}.self_attn.grep_linear.bias"] = bias(8, 0.2)
        W[f"{p}.self_attn_layer_norm.weight"], W[f"{p}.self_attn_layer_norm.bias"] = ln(C)
        xf = math.sqrt(2.0 / (C + Ff)) * beta
        W[f"{p}.fc1.weight"] = normal((Ff, C), xf)
        W[f"{p}.fc1.bias"] = bias(Ff)
        W[f"{p}.fc2.weight"] = normal((C, Ff), xf)
        W[f"{p}.fc2.bias"] = bias(C)
        W[f"{p}.final_layer_norm.weight"], W[f"{p}.final_layer_norm.bias"] = ln(C)
    return W


def make_predictor_weights(seed: int = 9, embed: int = 768, classes: int = 527) -> dict:
    """`backbone.predictor` (beats.py:262-264, the AudioSet-527 head of fine-tuned checkpoints): seeded N(0, 0.05) weight and
    N(0, 0.02) bias so that every logit is observable."""
    rs = np.random.RandomState(seed)
    return {
        "backbone.predictor.weight": (rs.standard_normal((classes, embed)) * 0.05).astype(np.float32),
        "backbone.predictor.bias": (rs.standard_normal((classes,)) * 0.02).astype(np.float32),
    }


def make_effnet_weights(seed: int = 3, num_classes: int = 0, bn_stats: dict | None = None) -> dict:
    """Deterministic synthetic EfficientNet-B0 weights with torchvision's state_dict keys (prefix `model.`, as
    avex/models/efficientnet.py:61-66 holds the network).  Conv weights: fan-out normal (torchvision's init,
    efficientnet.py `_efficientnet` -> kaiming_normal_ fan_out); BatchNorm affine mildly perturbed so it is observable.
    BatchNorm running statistics: `bn_stats` (tests/golden/effnet_bn_stats.npz, the calibration pass of
    tests/golden/make_golden_effnet.py -- a random network with mean 0 / var 1 statistics collapses to 1e-13 at the
    head) or the identity statistics when None."""
    from .effnet import B0_STAGES, HEAD_OUT, STEM_OUT, block_list

    rs = np.random.RandomState(seed)
    W: dict = {}

    def conv(name, co, ci, k):
        std = math.sqrt(2.0 / (co * k * k))
        W[name + ".weight"] = (rs.standard_normal((co, ci, k, k)) * std).astype(np.float32)

    def bnorm(name, c):
        W[name + ".weight"] = (1.0 + 0.1 * rs.standard_normal(c)).astype(np.float32)
        W[name + ".bias"] = (0.1 * rs.standard_normal(c)).astype(np.float32)
        if bn_stats is not None:
            W[name + ".running_mean"] = np.asarray(bn_stats[name + ".running_mean"], np.float32)
            W[name + ".running_var"] = np.asarray(bn_stats[name + ".running_var"], np.float32)
        else:
            W[name + ".running_mean"] = np.zeros(c, np.float32)
            W[name + ".running_var"] = np.ones(c, np.float32)
        W[name + ".num_batches_tracked"] = np.zeros((), np.int64)

    conv("model.features.0.0", STEM_OUT, 3, 3)
    bnorm("model.features.0.1", STEM_OUT)
    for prefix, k, _s, cin, cexp, cout, csq in block_list(B0_STAGES):
        j = 0
        if cexp != cin:
            conv(f"{prefix}.0.0", cexp, cin, 1)
            bnorm(f"{prefix}.0.1", cexp)
            j = 1
        conv(f"{prefix}.{j}.0", cexp, 1, k)
        bnorm(f"{prefix}.{j}.1", cexp)
        conv(f"{prefix}.{j + 1}.fc1", csq, cexp, 1)
        W[f"{prefix}.{j + 1}.fc1.bias"] = (0.1 * rs.standard_normal(csq)).astype(np.float32)
        conv(f"{prefix}.{j + 1}.fc2", cexp, csq, 1)
        W[f"{prefix}.{j + 1}.fc2.bias"] = (0.1 * rs.standard_normal(cexp)).astype(np.float32)
        conv(f"{prefix}.{j + 2}.0", cout, cexp, 1)
        bnorm(f"{prefix}.{j + 2}.1", cout)
    hp = f"model.features.{len(B0_STAGES) + 1}"
    conv(hp + ".0", HEAD_OUT, B0_STAGES[-1][4], 1)
    bnorm(hp + ".1", HEAD_OUT)
    if num_classes:
        bound = 1.0 / math.sqrt(num_classes)
        W["model.classifier.1.weight"] = rs.uniform(-bound, bound, size=(num_classes, HEAD_OUT)).astype(np.float32)
        W["model.classifier.1.bias"] = (0.1 * rs.standard_normal(num_classes)).astype(np.float32)
    return W
