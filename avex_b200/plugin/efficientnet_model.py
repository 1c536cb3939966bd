"""EfficientNet `Model(ModelBase)`: the drop-in for avex/models/efficientnet.py:22-322.

Same constructor keywords (what `build_model_from_spec` passes after signature filtering, factory.py:140-161), the
torchvision network itself as parameter holder under `self.model` (so `state_dict()` keys, `named_modules()` and the 17
discoverable layer names are identical), same `process_audio` / `forward` / `extract_embeddings` semantics -- with the
mel front end and the whole CNN in the CUDA path (`avex_b200.effnet.EfficientNetEngine`).
"""
from __future__ import annotations

import logging
from typing import List, Optional, Union

import torch

from ..effnet import EfficientNetEngine
from .base_model import ModelBase
from .registry import register_model_class

logger = logging.getLogger(__name__)


def make_model_class(base):
    class Model(base):
        name = "efficientnet"

        def __init__(
            self,
            num_classes: Optional[int] = None,
            pretrained: bool = True,
            device: str = "cuda",
            audio_config=None,
            return_features_only: bool = False,
            efficientnet_variant: str = "b0",
        ) -> None:
            super().__init__(device=device, audio_config=audio_config)
            from torchvision.models import efficientnet_b0, efficientnet_b1

            if num_classes is None:
                return_features_only = True  # efficientnet.py:42-52: no head without a class count
                self.num_classes = None
            else:
                self.num_classes = num_classes
            self.return_features_only = return_features_only
            self.gradient_checkpointing = False
            self.audio_config = audio_config
            if pretrained:
                raise RuntimeError(
                    "pretrained=True downloads ImageNet weights (efficientnet.py:61-63); pass pretrained=False and load "
                    "weights with load_model(..., checkpoint_path=...) or load_state_dict().  NOTE: the reference factory "
                    "never forwards ModelSpec.pretrained (factory.py:30-46), so it must be given explicitly."
                )
            if efficientnet_variant == "b0":
                self.model = efficientnet_b0(weights=None)
            elif efficientnet_variant == "b1":
                self.model = efficientnet_b1(weights=None)
            else:
                raise ValueError(f"Unsupported EfficientNet variant: {efficientnet_variant}")
            if not self.return_features_only:
                in_features = self.model.classifier[-1].in_features
                self.model.classifier[-1] = torch.nn.Linear(in_features, num_classes)
            self.model = self.model.to(self.device)
            self._engine = EfficientNetEngine(self.model)

        def _discover_embedding_layers(self) -> None:
            """efficientnet.py:82-114: stem conv, every `block.3.0` project conv, head conv."""
            if len(self._layer_names) == 0:
                self._layer_names = [
                    n for n, _ in self.named_modules()
                    if n == "model.features.0.0" or (n.endswith(".block.3.0") and "model.features." in n) or n == "model.features.8.0"
                ]  # fmt: skip

        def register_hooks_for_layers(self, target_layers) -> List[str]:
            """base_model.py:101-200, restricted to the conv outputs the NHWC kernels materialise (the 17 layers
            `_discover_embedding_layers` reports); any other module name is refused up front."""
            names = super().register_hooks_for_layers(target_layers)
            self._discover_embedding_layers()
            ok = set(self._layer_names)
            bad = [n for n in names if n not in ok]
            if bad:
                self.deregister_all_hooks()
                self._hook_layers = []
                raise ValueError(
                    f"avex_b200 EfficientNet cannot serve forward hooks on {bad}: the fused CUDA forward materialises only "
                    f"{self._layer_names} (select them by name, index, 'all' or 'last_layer')."
                )
            return names

        def _mel(self, x: torch.Tensor, normalize: bool):
            if x is None:
                raise ValueError("Input tensor cannot be None")
            if x.dtype != torch.float32:
                x = x.to(torch.float32)  # efficientnet.py:131-133
            cfg = self.audio_config
            if cfg is not None and (cfg.representation != "mel_spectrogram" or (cfg.n_fft, cfg.hop_length or cfg.n_fft // 4, cfg.win_length or cfg.n_fft, cfg.n_mels) != (800, 160, 800, 128)
                                    or cfg.window != "hann" or not cfg.normalize or not getattr(cfg, "center", True) or cfg.sample_rate != 16000):
                raise NotImplementedError("avex_b200 mel kernel is specialised to the esp_aves2_effnet audio_config (16 kHz, n_fft 800, hop 160, 128 mels, hann, normalize)")
            x = x.to(next(self.parameters()).device)
            return self._engine.mel.run(x, normalize=normalize, return_minmax=True)

        def process_audio(self, x: torch.Tensor) -> torch.Tensor:
            """efficientnet.py:116-142: normalised mel image repeated to 3 channels, [B, 3, 128, frames]."""
            img, _ = self._mel(x, normalize=True)
            return img.unsqueeze(1).repeat(1, 3, 1, 1)

        def enable_gradient_checkpointing(self) -> None:
            self.gradient_checkpointing = True

        def forward(self, x: torch.Tensor, padding_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
            img, minmax = self._mel(x, normalize=False)  # the stem kernel normalises on load
            eng = self._engine
            res = eng.run(img, minmax, want_features=self.return_features_only, want_logits=not self.return_features_only,
                          hook_layers=eng.hooked())
            eng.fire_hooks(res["hooks"])
            return res["features"] if self.return_features_only else res["logits"]

        def extract_embeddings(
            self,
            x,
            *,
            padding_mask: Optional[torch.Tensor] = None,
            aggregation: str = "none",
            freeze_backbone: bool = True,
        ) -> Union[torch.Tensor, List[torch.Tensor]]:
            """efficientnet.py:217-322: hooked [B, C, H, W] tensors; 'mean' / 'max' reduce the last (time) axis, then
            flatten C*H and concatenate across layers."""
            if not self._hooks:
                raise ValueError("No hooks are registered in the model.")
            if not freeze_backbone:
                raise NotImplementedError("avex_b200 EfficientNet is a frozen-backbone (forward-only) path")
            self._clear_hook_outputs()
            try:
                wav = x["raw_wav"] if isinstance(x, dict) else x
                was_training = self.training
                if was_training:
                    self.eval()
                try:
                    with torch.no_grad():
                        self.forward(wav, padding_mask)
                finally:
                    if was_training:
                        self.train()
                order = self._hook_layers if self._hook_layers else list(self._hook_outputs.keys())
                embeddings = [self._hook_outputs[n] for n in order]
                if not embeddings:
                    raise ValueError("No outputs were captured from registered hooks.")
                if aggregation == "none":
                    return embeddings[0] if len(embeddings) == 1 else embeddings
                for i in range(len(embeddings)):
                    if embeddings[i].dim() == 2:
                        continue
                    if aggregation == "mean":
                        embeddings[i] = embeddings[i].mean(dim=-1)
                    elif aggregation == "max":
                        embeddings[i] = embeddings[i].max(dim=-1)[0]
                    elif aggregation == "cls_token":
                        embeddings[i] = embeddings[i][:, 0, :]
                    else:
                        raise ValueError(f"Unsupported aggregation method: {aggregation}")
                    if embeddings[i].dim() == 3:
                        embeddings[i] = embeddings[i].reshape(embeddings[i].shape[0], -1)
                    else:
                        raise ValueError(f"Unexpected embedding dimension: {embeddings[i].dim()}. Expected 2, 3, or 4.")
                return torch.cat(embeddings, dim=1)
            finally:
                self._clear_hook_outputs()

    return Model


Model = register_model_class(make_model_class(ModelBase))
