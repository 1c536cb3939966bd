"""Host side of the EfficientNet feature extractor: wraps a torchvision `EfficientNet` module (the parameter holder, so
`state_dict()` keys, `named_modules()` and forward hooks are torchvision's own) and runs its forward through the
NHWC bf16 CUDA path behind `avexk_effnet_forward` (avex/models/efficientnet.py:163-215).  No torch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .beats import call_forward_hooks
from .melspec import N_MELS, MelSpectrogram


class EfficientNetEngine:
    def __init__(self, net: nn.Module) -> None:
        self.net = net  # torchvision.models.EfficientNet
        self.mel = MelSpectrogram()
        self._engine = None
        self._engine_key = None
        self._ws: Optional[torch.Tensor] = None
        # MBConv blocks in forward order + the modules whose forward hooks we serve
        self.blocks = [blk for stage in list(net.features)[1:-1] for blk in stage]
        self.hook_modules: list[nn.Module] = [net.features[0][0]]
        self.cfgs = []
        for blk in self.blocks:
            layers = list(blk.block)
            has_expand = len(layers) == 4
            dw = layers[1 if has_expand else 0][0]
            se = layers[2 if has_expand else 1]
            proj = layers[-1][0]
            self.cfgs.append(dict(kernel=dw.kernel_size[0], stride=dw.stride[0], cin=(layers[0][0].in_channels if has_expand else dw.in_channels),
                                  cexp=dw.in_channels, cout=proj.out_channels, csq=se.fc1.out_channels, has_expand=has_expand))
            self.hook_modules.append(proj)
        self.hook_modules.append(net.features[-1][0])
        self.head_out = net.features[-1][0].out_channels

    # ---- engine management ------------------------------------------------------------------------------------
    def _weight_version(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.net.parameters()) + list(self.net.buffers()))

    def _ensure_engine(self, device: torch.device, with_classifier: bool):
        key = (device, with_classifier, self._weight_version())
        if self._engine is not None and self._engine_key == key:
            return self._engine
        lib = _lib.load()
        self.release()
        keep = []

        def ptr(t: torch.Tensor) -> int:
            t = t.detach()
            if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(device=device, dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        def bn(m: nn.BatchNorm2d) -> _lib.BnParams:
            return _lib.BnParams(ptr(m.weight), ptr(m.bias), ptr(m.running_mean), ptr(m.running_var))

        nb = len(self.blocks)
        cfg_arr = (_lib.EffnetBlockCfg * nb)()
        bw = (_lib.EffnetBlockWeights * nb)()
        for i, (blk, c) in enumerate(zip(self.blocks, self.cfgs)):
            cfg_arr[i] = _lib.EffnetBlockCfg(c["kernel"], c["stride"], c["cin"], c["cexp"], c["cout"], c["csq"])
            layers = list(blk.block)
            j = 0
            if c["has_expand"]:
                bw[i].expand_w = ptr(layers[0][0].weight)
                bw[i].expand_bn = bn(layers[0][1])
                j = 1
            bw[i].dw_w = ptr(layers[j][0].weight)
            bw[i].dw_bn = bn(layers[j][1])
            se = layers[j + 1]
            bw[i].se1_w, bw[i].se1_b = ptr(se.fc1.weight), ptr(se.fc1.bias)
            bw[i].se2_w, bw[i].se2_b = ptr(se.fc2.weight), ptr(se.fc2.bias)
            bw[i].proj_w = ptr(layers[j + 2][0].weight)
            bw[i].proj_bn = bn(layers[j + 2][1])
        stem, head = self.net.features[0], self.net.features[-1]
        w = _lib.EffnetWeights(stem_w=ptr(stem[0].weight), stem_bn=bn(stem[1]), blocks=bw, head_w=ptr(head[0].weight), head_bn=bn(head[1]))
        if with_classifier:
            lin = self.net.classifier[-1]
            w.cls_w, w.cls_b, w.num_classes = ptr(lin.weight), ptr(lin.bias), lin.out_features
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.avexk_effnet_create(cfg_arr, nb, stem[0].out_channels, self.head_out, C.byref(h)), "avexk_effnet_create")
            st = torch.cuda.current_stream(device).cuda_stream
            rc = lib.avexk_effnet_load_weights(h, C.byref(w), st)
            if rc != 0:
                msg = lib.avexk_last_error().decode("utf-8", "replace")
                lib.avexk_effnet_destroy(h)
                raise _lib.AvexkError(f"avexk_effnet_load_weights failed (code {rc}): {msg}")
        del keep
        self._engine, self._engine_key = h, key
        return h

    def invalidate(self) -> None:
        """Drop the packed weights (needed after in-place `.data` updates, which do not bump tensor versions)."""
        self.release()

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"], st["_engine_key"], st["_ws"] = None, None, None
        return st

    def release(self) -> None:
        if self._engine is not None:
            _lib.load().avexk_effnet_destroy(self._engine)
            self._engine = None
            self._engine_key = None

    def __del__(self):  # pragma: no cover
        try:
            self.release()
        except Exception:
            pass

    # ---- compute ----------------------------------------------------------------------------------------------
    def hooked(self) -> list[int]:
        return [i for i, m in enumerate(self.hook_modules) if m._forward_hooks]

    def fire_hooks(self, hooks: dict) -> None:
        for i, t in hooks.items():
            mod = self.hook_modules[i]
            call_forward_hooks(mod, t)

    def run(self, image: torch.Tensor, minmax: Optional[torch.Tensor], *, want_features: bool = True, want_logits: bool = False,
            hook_layers: Optional[list[int]] = None) -> dict:
        """image [B, 128, frames] fp32 CUDA (log-mel; un-normalised when `minmax` is given)."""
        if not image.is_cuda:
            raise _lib.AvexkError("avex_b200 EfficientNet runs on CUDA tensors only (no CPU fallback)")
        if self.net.training:
            raise _lib.AvexkError(
                "avex_b200 EfficientNet is an inference path: BatchNorm uses running statistics -- call model.eval() first "
                "(train-mode batch statistics and autograd are not implemented)"
            )
        device = image.device
        image = image.float().contiguous()
        B, H0, W0 = image.shape
        lib = _lib.load()
        eng = self._ensure_engine(device, want_logits)
        hf, wf = C.c_int(), C.c_int()
        _lib.check(lib.avexk_effnet_out_hw(eng, H0, W0, C.byref(hf), C.byref(wf)), "avexk_effnet_out_hw")
        need = lib.avexk_effnet_workspace_bytes(eng, B, H0, W0)
        if self._ws is None or self._ws.device != device or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        feats = torch.empty((B, self.head_out, hf.value, wf.value), device=device, dtype=torch.float32) if want_features else None
        logits = torch.empty((B, self.net.classifier[-1].out_features), device=device, dtype=torch.float32) if want_logits else None
        hooks: dict[int, torch.Tensor] = {}
        nb = len(self.blocks)
        hook_ptrs = (C.c_void_p * (nb + 2))()
        if hook_layers:
            h, w = (H0 + 2 - 3) // 2 + 1, (W0 + 2 - 3) // 2 + 1
            shapes = [(self.hook_modules[0].out_channels, h, w)]
            for c in self.cfgs:
                p = (c["kernel"] - 1) // 2
                h, w = (h + 2 * p - c["kernel"]) // c["stride"] + 1, (w + 2 * p - c["kernel"]) // c["stride"] + 1
                shapes.append((c["cout"], h, w))
            shapes.append((self.head_out, h, w))
            for li in hook_layers:
                hooks[li] = torch.empty((B, *shapes[li]), device=device, dtype=torch.float32)
                hook_ptrs[li] = hooks[li].data_ptr()
        with torch.cuda.device(device):
            rc = lib.avexk_effnet_forward(
                eng, image.data_ptr(), minmax.data_ptr() if minmax is not None else None, B, H0, W0,
                feats.data_ptr() if feats is not None else None, logits.data_ptr() if logits is not None else None,
                hook_ptrs if hook_layers else None, self._ws.data_ptr(), self._ws.numel(),
                torch.cuda.current_stream(device).cuda_stream,
            )  # fmt: skip
        _lib.check(rc, "avexk_effnet_forward")
        return {"features": feats, "logits": logits, "hooks": hooks}
