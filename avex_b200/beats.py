"""Host side of the BEATs encoder: module tree identical to the reference, numerics in libavexk.

Mirrors avex/models/beats/beats.py:166-401 (`BEATsConfig`, `BEATs`) and backbone.py:38-221
(`TransformerEncoder`, `_TransformerSentenceEncoderLayer`, `_MultiheadAttention`):

* the SAME sub-module names and parameter shapes, so `state_dict()` keys equal the reference's
  (`post_extract_proj.*`, `patch_embedding.weight`, `layer_norm.*`, `encoder.pos_conv.0.bias`,
  `encoder.pos_conv.0.parametrizations.weight.original{0,1}`, `encoder.layers.{i}.self_attn.{q,k,v,out}_proj.*`,
  `...grep_linear.*`, `...grep_a`, `...relative_attention_bias.weight`, `fc1`, `fc2`, layer norms, `fbank.*`)
  and reference checkpoints load with `load_state_dict(strict=False)`;
* `get_submodule(name).register_forward_hook(...)` keeps working: after the fused forward the model fires the
  forward hooks of `post_extract_proj` and every `encoder.layers.{i}.fc2` with the tensors the kernels
  materialised (raw Linear outputs, `(T,B,C)` for the blocks like the reference, backbone.py:182);
* the parameter holders never compute: the whole forward is ONE C-ABI call (`avexk_beats_forward`).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, fields
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .fbank import KaldiFbank


@dataclass(init=False)
class BEATsConfig:
    """Same fields and defaults as the reference's pydantic BEATsConfig (beats.py:166-228); unknown keys are kept."""

    input_patch_size: int = 16
    embed_dim: int = 512
    conv_bias: bool = False
    encoder_layers: int = 12
    encoder_embed_dim: int = 768
    encoder_ffn_embed_dim: int = 3072
    encoder_attention_heads: int = 12
    activation_fn: str = "gelu"
    layer_wise_gradient_decay_ratio: float = 1.0
    layer_norm_first: bool = False
    deep_norm: bool = True
    dropout: float = 0.1
    attention_dropout: float = 0.1
    activation_dropout: float = 0.0
    encoder_layerdrop: float = 0.05
    dropout_input: float = 0.0
    conv_pos: int = 128
    conv_pos_groups: int = 16
    relative_position_embedding: bool = True
    num_buckets: int = 320
    max_distance: int = 800
    gru_rel_pos: bool = True
    sample_frequency: float = 16000.0
    num_mel_bins: int = 128
    frame_length: float = 25.0
    frame_shift: float = 10.0
    fbank_mean: float = 15.41663
    fbank_std: float = 6.55582
    finetuned_model: bool = False
    predictor_dropout: float = 0.0
    predictor_class: int = 527

    def __init__(self, **kw):
        known = {f.name: f for f in fields(type(self))}
        for name, f in known.items():
            setattr(self, name, kw.pop(name, f.default))
        self.extra = kw  # model_config = extra "allow"

    def model_dump(self) -> dict:
        d = {f.name: getattr(self, f.name) for f in fields(type(self))}
        d.update(self.extra)
        return d

    def check_supported(self) -> None:
        bad = []
        if not self.deep_norm or self.layer_norm_first:
            bad.append("only the post-LN DeepNorm block (deep_norm=True, layer_norm_first=False) is implemented")
        if self.activation_fn != "gelu":
            bad.append("activation_fn must be 'gelu'")
        if not (self.relative_position_embedding and self.gru_rel_pos):
            bad.append("gated relative position bias must be enabled")
        if self.conv_bias:
            bad.append("conv_bias=True is not supported")
        if self.input_patch_size != 16 or self.num_mel_bins != 128:
            bad.append("patch 16 / 128 mel bins only")
        if self.encoder_embed_dim != 64 * self.encoder_attention_heads:
            bad.append("head_dim must be 64")
        if bad:
            raise ValueError("avex_b200 BEATs kernels: " + "; ".join(bad))


def relative_position_bucket(rel: torch.Tensor, num_buckets: int, max_distance: int) -> torch.Tensor:
    """Bidirectional T5 buckets with the reference's exact fp32 arithmetic (backbone.py:438-473)."""
    half = num_buckets // 2
    side = (rel > 0).to(torch.long) * half
    dist = rel.abs()
    exact = half // 2
    log_part = exact + (torch.log(dist.float() / exact) / math.log(max_distance / exact) * (half - exact)).to(torch.long)
    log_part = torch.min(log_part, torch.full_like(log_part, half - 1))
    return side + torch.where(dist < exact, dist, log_part)


def relative_bias_vector(table: torch.Tensor, n_tokens: int, num_buckets: int, max_distance: int) -> torch.Tensor:
    """[H, 2N-1] with vec[h, (j-i)+N-1] = table[bucket(j-i), h] -- the Toeplitz generator of compute_bias (backbone.py:475-492)."""
    rel = torch.arange(-(n_tokens - 1), n_tokens, dtype=torch.long)
    idx = relative_position_bucket(rel, num_buckets, max_distance).to(table.device)
    return table.detach().float()[idx].t().contiguous()


def call_forward_hooks(mod: nn.Module, out):
    """Fire `mod`'s forward hooks as `nn.Module._call_impl` would after a forward that produced `out` (positional inputs are
    not available: the fused kernels never materialise them).  Hooks registered `with_kwargs=True` get the 4-argument form."""
    for hid, hook in list(mod._forward_hooks.items()):
        if hid in getattr(mod, "_forward_hooks_with_kwargs", {}):
            r = hook(mod, (), {}, out)
        else:
            r = hook(mod, (), out)
        if r is not None:
            out = r
    return out


def forward_padding_mask(n_feat: int, mask: torch.Tensor) -> torch.Tensor:
    """beats.py:283-302."""
    extra = mask.size(1) % n_feat
    if extra > 0:
        mask = mask[:, :-extra]
    return mask.reshape(mask.size(0), n_feat, -1).all(-1)


class _SelfAttnParams(nn.Module):
    """Parameter holder for `_MultiheadAttention` (backbone.py:378-424)."""

    def __init__(self, dim: int, heads: int, num_buckets: int) -> None:
        super().__init__()
        self.k_proj = nn.Linear(dim, dim)
        self.v_proj = nn.Linear(dim, dim)
        self.q_proj = nn.Linear(dim, dim)
        self.out_proj = nn.Linear(dim, dim)
        self.grep_linear = nn.Linear(dim // heads, 8)
        self.grep_a = nn.Parameter(torch.ones(1, heads, 1, 1))
        self.relative_attention_bias = nn.Embedding(num_buckets, heads)


class _BlockParams(nn.Module):
    """Parameter holder for `_TransformerSentenceEncoderLayer` (backbone.py:224-308)."""

    def __init__(self, cfg: BEATsConfig) -> None:
        super().__init__()
        d = cfg.encoder_embed_dim
        self.self_attn = _SelfAttnParams(d, cfg.encoder_attention_heads, cfg.num_buckets)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, cfg.encoder_ffn_embed_dim)
        self.fc2 = nn.Linear(cfg.encoder_ffn_embed_dim, d)
        self.final_layer_norm = nn.LayerNorm(d)


class _EncoderParams(nn.Module):
    """Parameter holder for `TransformerEncoder` (backbone.py:38-124), with the reference's init distributions."""

    def __init__(self, cfg: BEATsConfig) -> None:
        super().__init__()
        d = cfg.encoder_embed_dim
        conv = nn.Conv1d(d, d, kernel_size=cfg.conv_pos, padding=cfg.conv_pos // 2, groups=cfg.conv_pos_groups)
        nn.init.normal_(conv.weight, mean=0, std=math.sqrt(4.0 / (cfg.conv_pos * d)))
        nn.init.constant_(conv.bias, 0)
        conv = nn.utils.parametrizations.weight_norm(conv, name="weight", dim=2)
        self.pos_conv = nn.Sequential(conv)  # index 0 carries the parameters; SamePad / GELU are fused in the kernel
        self.layers = nn.ModuleList([_BlockParams(cfg) for _ in range(cfg.encoder_layers)])
        for i in range(1, cfg.encoder_layers):  # one shared table, backbone.py:100-103
            self.layers[i].self_attn.relative_attention_bias = self.layers[0].self_attn.relative_attention_bias
        self.layer_norm = nn.LayerNorm(d)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0.0, 0.02)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, 0.0, 0.02)
        beta = math.pow(8 * cfg.encoder_layers, -0.25)
        for blk in self.layers:
            nn.init.xavier_normal_(blk.self_attn.k_proj.weight, gain=1)
            nn.init.xavier_normal_(blk.self_attn.q_proj.weight, gain=1)
            nn.init.xavier_normal_(blk.self_attn.v_proj.weight, gain=beta)
            nn.init.xavier_normal_(blk.self_attn.out_proj.weight, gain=beta)
            nn.init.xavier_normal_(blk.fc1.weight, gain=beta)
            nn.init.xavier_normal_(blk.fc2.weight, gain=beta)


class BEATs(nn.Module):
    """Drop-in for the reference `BEATs` module: same attributes, `preprocess`, `extract_features`, `forward`."""

    def __init__(self, cfg: BEATsConfig) -> None:
        super().__init__()
        cfg.check_supported()
        self.cfg = cfg
        self.embed = cfg.embed_dim
        self.post_extract_proj = nn.Linear(cfg.embed_dim, cfg.encoder_embed_dim)
        self.fbank = KaldiFbank(
            num_mel_bins=cfg.num_mel_bins, sample_frequency=cfg.sample_frequency,
            frame_length_ms=cfg.frame_length, frame_shift_ms=cfg.frame_shift,
        )  # fmt: skip
        self.fbank_mean = cfg.fbank_mean
        self.fbank_std = cfg.fbank_std
        self.input_patch_size = cfg.input_patch_size
        self.patch_embedding = nn.Conv2d(1, cfg.embed_dim, kernel_size=16, stride=16, bias=False)
        self.encoder = _EncoderParams(cfg)
        self.layer_norm = nn.LayerNorm(cfg.embed_dim)
        self.predictor = nn.Linear(cfg.encoder_embed_dim, cfg.predictor_class) if cfg.finetuned_model else None
        self._engine = None
        self._engine_key = None
        # "bf16" (default): bf16 tensor-core operands, max-abs <= 2e-2 vs the fp32 reference.  "fp32": 3-term split GEMMs with
        # fp32 attention / pos-conv (max-abs <= 1e-3; ~10x slower) -- set `model.backbone.precision = "fp32"` before a forward.
        self.precision = "bf16"
        self._bias_cache: dict = {}
        self._ws: Optional[torch.Tensor] = None

    # ---- engine management ------------------------------------------------------------------------------------
    def _weight_version(self):
        """(storage, autograd version) per parameter.  In-place writes through `.data` (as the reference's own init does with
        `.data.copy_`) do NOT bump `_version`: call `invalidate()` after such an update."""
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def invalidate(self) -> None:
        """Drop the packed bf16 weight copies and cached bias vectors; the next forward re-packs from the parameters."""
        self.release()
        self._bias_cache.clear()

    refresh_weights = invalidate

    def load_state_dict(self, *a, **kw):
        out = super().load_state_dict(*a, **kw)  # copies in place under no_grad: versions may not change
        self.invalidate()
        return out

    # native handles are per process: copies / pickles carry parameters only and rebuild lazily
    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"], st["_engine_key"], st["_ws"], st["_bias_cache"] = None, None, None, {}
        return st

    def __deepcopy__(self, memo):
        import copy

        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def _ensure_engine(self, device: torch.device):
        key = (device, self._weight_version(), self.precision)
        if self._engine is not None and self._engine_key == key:
            return self._engine
        lib = _lib.load()
        self.release()
        cfg = self.cfg
        dims = _lib.BeatsDims(
            cfg.encoder_layers, cfg.encoder_embed_dim, cfg.encoder_ffn_embed_dim, cfg.encoder_attention_heads,
            cfg.embed_dim, cfg.conv_pos, cfg.conv_pos_groups, float(cfg.fbank_mean), float(cfg.fbank_std), 1e-5,
        )  # fmt: skip
        h = C.c_void_p()
        keep = []

        def ptr(t: torch.Tensor) -> int:
            t = t.detach()
            if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(device=device, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        with torch.cuda.device(device):
            _lib.check(lib.avexk_beats_create(C.byref(dims), C.byref(h)), "avexk_beats_create")
            _lib.check(lib.avexk_beats_set_precision(h, 1 if self.precision == "fp32" else 0), "avexk_beats_set_precision")
            enc = self.encoder
            layers = (_lib.BeatsLayerWeights * cfg.encoder_layers)()
            for i, blk in enumerate(enc.layers):
                a = blk.self_attn
                vals = dict(
                    q_w=a.q_proj.weight, q_b=a.q_proj.bias, k_w=a.k_proj.weight, k_b=a.k_proj.bias,
                    v_w=a.v_proj.weight, v_b=a.v_proj.bias, o_w=a.out_proj.weight, o_b=a.out_proj.bias,
                    grep_w=a.grep_linear.weight, grep_b=a.grep_linear.bias, grep_a=a.grep_a,
                    ln1_w=blk.self_attn_layer_norm.weight, ln1_b=blk.self_attn_layer_norm.bias,
                    fc1_w=blk.fc1.weight, fc1_b=blk.fc1.bias, fc2_w=blk.fc2.weight, fc2_b=blk.fc2.bias,
                    ln2_w=blk.final_layer_norm.weight, ln2_b=blk.final_layer_norm.bias,
                )  # fmt: skip
                for k, v in vals.items():
                    setattr(layers[i], k, ptr(v))
            conv = enc.pos_conv[0]
            w = _lib.BeatsWeights(
                patch_w=ptr(self.patch_embedding.weight), ln0_w=ptr(self.layer_norm.weight), ln0_b=ptr(self.layer_norm.bias),
                proj_w=ptr(self.post_extract_proj.weight), proj_b=ptr(self.post_extract_proj.bias),
                posconv_g=ptr(conv.parametrizations.weight.original0), posconv_v=ptr(conv.parametrizations.weight.original1),
                posconv_b=ptr(conv.bias), enc_ln_w=ptr(enc.layer_norm.weight), enc_ln_b=ptr(enc.layer_norm.bias),
                rel_bias_table=ptr(enc.layers[0].self_attn.relative_attention_bias.weight), layers=layers,
            )  # fmt: skip
            st = torch.cuda.current_stream(device).cuda_stream
            _lib.check(lib.avexk_beats_load_weights(h, C.byref(w), st), "avexk_beats_load_weights")
        del keep
        self._engine, self._engine_key = h, key
        self._bias_cache.clear()
        return h

    def release(self) -> None:
        if self._engine is not None:
            _lib.load().avexk_beats_destroy(self._engine)
            self._engine = None
            self._engine_key = None

    def __del__(self):  # pragma: no cover
        try:
            self.release()
        except Exception:
            pass

    def _bias_vec(self, n_tokens: int, device: torch.device) -> torch.Tensor:
        key = (n_tokens, device)
        if key not in self._bias_cache:
            table = self.encoder.layers[0].self_attn.relative_attention_bias.weight
            self._bias_cache[key] = relative_bias_vector(table, n_tokens, self.cfg.num_buckets, self.cfg.max_distance).to(device)
        return self._bias_cache[key]

    # ---- reference-compatible methods -------------------------------------------------------------------------
    def forward_padding_mask(self, features: torch.Tensor, padding_mask: torch.Tensor) -> torch.Tensor:
        return forward_padding_mask(features.size(1), padding_mask)

    def num_tokens(self, num_samples: int) -> int:
        return 8 * (self.fbank.num_frames(num_samples) // 16)

    def preprocess(self, source: torch.Tensor) -> torch.Tensor:
        """beats.py:304-323: fbank(source * 2**15), then (x - mean) / (2 std); one fused kernel, fp32."""
        return self.fbank.run(source, prescale=32768.0, norm_mean=self.fbank_mean, norm_std2=2.0 * self.fbank_std)

    def run(
        self,
        source: torch.Tensor,
        padding_mask: Optional[torch.Tensor] = None,
        *,
        want_features: bool = True,
        want_pooled: bool = False,
        hook_layers: Optional[list[int]] = None,
        hook_pool: bool = False,
    ) -> dict:
        """One fused forward.  Returns {"features", "pooled", "hooks": {idx: tensor}, "padding_mask"}.

        Hooked tensors are [B,N,C] -- or, with `hook_pool=True`, their mean over the N tokens, [B,C], formed inside the fc2
        epilogue without ever materialising [B,N,C] (what `extract_embeddings(aggregation="mean")` reduces them to)."""
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            raise _lib.AvexkError(
                "avex_b200 BEATs is forward-only: call under torch.no_grad() / freeze_backbone=True "
                "(autograd through the fused kernels is not implemented)"
            )
        if not source.is_cuda:
            raise _lib.AvexkError("avex_b200 BEATs runs on CUDA tensors only (no CPU fallback)")
        if source.dim() != 2:
            raise ValueError(f"expected [B, T] waveforms, got {tuple(source.shape)}")
        device = source.device
        x = source if source.dtype == torch.float32 else source.float()
        if x.stride(1) != 1:
            x = x.contiguous()
        B, T = x.shape
        N = self.num_tokens(T)
        if N <= 0:
            raise RuntimeError(f"clip too short for one BEATs patch row: {T} samples")
        cfg = self.cfg
        Cdim = cfg.encoder_embed_dim
        key_pad = None
        tok_mask = None
        if padding_mask is not None:
            fm = forward_padding_mask(self.fbank.num_frames(T), padding_mask.to(device=device, dtype=torch.bool))
            tok_mask = forward_padding_mask(N, fm)  # beats.py:346-347, :355-356
            key_pad = tok_mask.to(torch.uint8).contiguous()
        engine = self._ensure_engine(device)
        lib = _lib.load()
        need = lib.avexk_beats_workspace_bytes(engine, B, T)
        if self._ws is None or self._ws.device != device or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        feats = torch.empty((B, N, Cdim), device=device, dtype=torch.float32) if want_features else None
        pooled = torch.empty((B, Cdim), device=device, dtype=torch.float32) if want_pooled else None
        hooks: dict[int, torch.Tensor] = {}
        hook_ptrs = (C.c_void_p * (cfg.encoder_layers + 1))()
        pool_ptrs = (C.c_void_p * (cfg.encoder_layers + 1))()
        for li in hook_layers or []:
            hooks[li] = torch.empty((B, Cdim) if hook_pool else (B, N, Cdim), device=device, dtype=torch.float32)
            (pool_ptrs if hook_pool else hook_ptrs)[li] = hooks[li].data_ptr()
        bias_vec = self._bias_vec(N, device)
        with torch.cuda.device(device):
            rc = lib.avexk_beats_forward(
                engine, x.data_ptr(), B, T, x.stride(0) if B > 1 else max(T, x.stride(0)), self.fbank.handle(device),
                key_pad.data_ptr() if key_pad is not None else None, bias_vec.data_ptr(),
                feats.data_ptr() if feats is not None else None, hook_ptrs, pool_ptrs,
                pooled.data_ptr() if pooled is not None else None,
                self._ws.data_ptr(), self._ws.numel(), torch.cuda.current_stream(device).cuda_stream,
            )  # fmt: skip
        _lib.check(rc, "avexk_beats_forward")
        return {"features": feats, "pooled": pooled, "hooks": hooks, "padding_mask": tok_mask}

    def _fire_hooks(self, hooks: dict) -> None:
        """Run the forward hooks registered on post_extract_proj / fc2 with the tensors the kernels produced."""
        for li, t in hooks.items():
            mod = self.post_extract_proj if li == 0 else self.encoder.layers[li - 1].fc2
            out = t if (li == 0 or t.dim() == 2) else t.transpose(0, 1)  # blocks run (T,B,C) in the reference, backbone.py:182
            out = call_forward_hooks(mod, out)

    def hooks_are_only(self, handles) -> bool:
        """True when every forward hook on the servable modules is one of `handles` (the ModelBase capture hooks): only then may
        `extract_embeddings(aggregation="mean")` hand the hooks token-pooled [B,C] tensors instead of [B,N,C]."""
        own = {h.id for h in handles}
        mods = [self.post_extract_proj] + [blk.fc2 for blk in self.encoder.layers]
        return all(hid in own for m in mods for hid in m._forward_hooks)

    def _hooked_layers(self) -> list[int]:
        idx = []
        if self.post_extract_proj._forward_hooks:
            idx.append(0)
        for i, blk in enumerate(self.encoder.layers):
            if blk.fc2._forward_hooks:
                idx.append(i + 1)
        return idx

    def extract_features(self, source, padding_mask=None, feature_only: bool = False, disable_layerdrop: bool = False):
        """beats.py:325-382.  Returns (features [B,N,C], token padding mask) or (logits, mask) with a predictor."""
        res = self.run(source, padding_mask, want_features=True, hook_layers=self._hooked_layers())
        self._fire_hooks(res["hooks"])
        x, mask = res["features"], res["padding_mask"]
        if not feature_only and self.predictor is not None:
            logits = torch.nn.functional.linear(x, self.predictor.weight, self.predictor.bias)
            if mask is not None and mask.any():
                logits[mask] = 0
                summed = logits.sum(dim=1)
                logits = summed / (~mask).sum(dim=1).unsqueeze(-1).expand_as(summed)  # beats.py:373-376
            else:
                logits = logits.mean(dim=1)
            return logits, mask
        return x, mask

    def forward(self, source, padding_mask=None, disable_layerdrop: bool = False):
        return self.extract_features(source, padding_mask, feature_only=True, disable_layerdrop=disable_layerdrop)
