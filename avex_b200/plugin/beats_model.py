"""BEATs `Model(ModelBase)`: the drop-in for avex/models/beats_model.py:72-435.

Same constructor keywords (what `build_model_from_spec` passes after signature filtering, factory.py:140-161),
same sub-module names / state_dict keys (`backbone.*`, `classifier.*`), same layer discovery
(`backbone.post_extract_proj` + every `backbone.encoder.layers.{i}.fc2`), same `forward` / `extract_embeddings`
semantics -- with all numerics in the fused CUDA path (`avex_b200.beats.BEATs.run`).
"""
from __future__ import annotations

import logging
from typing import List, Optional, Union

import torch
import torch.nn as nn

from ..beats import BEATs, BEATsConfig
from .base_model import ModelBase, aggregate_embeddings
from .registry import register_model_class

logger = logging.getLogger(__name__)


def make_model_class(base):
    """Build the BEATs wrapper on top of `base` (this package's ModelBase mirror, or avex's own ModelBase)."""

    class Model(base):
        name = "beats"

        def __init__(
            self,
            *,
            num_classes: Optional[int] = None,
            pretrained: bool = False,
            device: str = "cuda",
            audio_config=None,
            return_features_only: bool = False,
            use_naturelm: bool = False,
            fine_tuned: bool = False,
            disable_layerdrop: bool = False,
            init_config: Optional[dict] = None,
        ) -> None:
            super().__init__(device=device, audio_config=audio_config)
            if num_classes is None:
                return_features_only = True  # beats_model.py:112-119: no head without a class count
                self.num_classes = None
            else:
                self.num_classes = num_classes
            self.disable_layerdrop = disable_layerdrop
            self.use_naturelm = use_naturelm
            self.fine_tuned = fine_tuned
            if pretrained:
                raise RuntimeError(
                    "pretrained=True downloads the official BEATs checkpoint (gs:// / hf://); pass pretrained=False and "
                    "load weights with load_model(..., checkpoint_path=...) or load_state_dict()."
                )
            if init_config is not None:
                cfg = BEATsConfig(**init_config)
            else:
                # packaged YAML `beats_cfg` of the reference (deep_norm=True); fine-tuned variants carry the predictor
                cfg = BEATsConfig(finetuned_model=bool(fine_tuned or use_naturelm))
            self.backbone = BEATs(cfg)
            self.backbone.to(device)
            self._return_features_only = return_features_only
            if not return_features_only:
                self.classifier = nn.Linear(cfg.encoder_embed_dim, num_classes)
            else:
                self.register_module("classifier", None)

        @property
        def return_features_only(self) -> bool:
            return self._return_features_only

        def _discover_embedding_layers(self) -> None:
            """beats_model.py:206-227: the projection after the conv front end, then every block's fc2."""
            if not self._layer_names:
                self._layer_names = [
                    n
                    for n, _ in self.named_modules()
                    if n.endswith("post_extract_proj") or (n.endswith(".fc2") and "backbone.encoder.layers." in n)
                ]

        def servable_layers(self) -> List[str]:
            """Layers whose forward hooks the fused forward can feed: the tensors the kernels materialise."""
            self._discover_embedding_layers()
            return list(self._layer_names)

        def register_hooks_for_layers(self, target_layers) -> List[str]:
            """base_model.py:101-200, restricted to the layers the fused kernels materialise.  The reference accepts any
            `get_submodule` name (e.g. `backbone.encoder.layers.3`, `...fc1`); here such a hook would never fire, so it is
            refused up front with the list of servable layers instead of failing later inside `extract_embeddings`."""
            names = super().register_hooks_for_layers(target_layers)
            ok = set(self.servable_layers())
            bad = [n for n in names if n not in ok]
            if bad:
                self.deregister_all_hooks()
                self._hook_layers = []
                raise ValueError(
                    f"avex_b200 BEATs cannot serve forward hooks on {bad}: the fused CUDA forward materialises only "
                    f"{sorted(ok, key=lambda n: (len(n), n))} (the layers `_discover_embedding_layers` reports; "
                    "select them by name, index, 'all' or 'last_layer')."
                )
            return names

        def process_audio(self, x: torch.Tensor) -> torch.Tensor:
            audio = super().process_audio(x)
            if self.use_naturelm:
                audio = torch.clamp(audio, -1.0, 1.0)  # beats_model.py:431-435
            return audio

        def forward(self, x: torch.Tensor, padding_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
            x = self.process_audio(x)
            bk = self.backbone
            hooked = bk._hooked_layers()
            if getattr(self, "_pool_hooks", False):
                # extract_embeddings(aggregation="mean"): the hooked tensors leave the kernels already mean-pooled over the tokens
                # ([B,768] per layer, a by-product of the fc2 epilogue); nothing of size [B,N,768] is written for them
                res = bk.run(x, padding_mask, want_features=False, hook_layers=hooked, hook_pool=True)
                bk._fire_hooks(res["hooks"])
                return None
            if self._return_features_only:
                res = bk.run(x, padding_mask, want_features=True, hook_layers=hooked)
                bk._fire_hooks(res["hooks"])
                return res["features"]
            # classifier mode: masked mean-pool fused into the last LayerNorm's consumer, then the tiny head
            res = bk.run(x, padding_mask, want_features=False, want_pooled=True, hook_layers=hooked)
            bk._fire_hooks(res["hooks"])
            return torch.nn.functional.linear(res["pooled"], self.classifier.weight, self.classifier.bias)

        def extract_embeddings(
            self,
            x,
            *,
            padding_mask: Optional[torch.Tensor] = None,
            aggregation: str = "none",
            freeze_backbone: bool = True,
        ) -> Union[torch.Tensor, List[torch.Tensor]]:
            if x is None:
                raise ValueError("Input tensor cannot be None")
            wav = x["raw_wav"] if isinstance(x, dict) else x
            if wav.numel() == 0 or wav.shape[-1] == 0:
                raise ValueError("Audio tensor cannot be empty")
            if not self._hooks:
                raise ValueError("No hooks are registered in the model.")
            if not freeze_backbone:
                raise NotImplementedError(
                    "avex_b200 BEATs is a frozen-backbone (forward-only) path; freeze_backbone=False needs autograd "
                    "through the backbone, which the fused kernels do not provide."
                )
            was_training = self.training
            if was_training:
                self.eval()
            try:
                self._clear_hook_outputs()
                mask = x.get("padding_mask") if isinstance(x, dict) else padding_mask
                # "mean" over the token axis commutes with nothing downstream (aggregate_embeddings passes 2-D tensors through),
                # so it is done on the device -- unless someone else's forward hooks would see the pooled tensors
                self._pool_hooks = aggregation == "mean" and self.backbone.hooks_are_only(self._hooks.values())
                try:
                    with torch.no_grad():
                        self.forward(wav, mask)
                finally:
                    self._pool_hooks = False
                order = self._hook_layers if self._hook_layers else list(self._hook_outputs.keys())
                embeddings = [self._hook_outputs[n] for n in order]
                if not embeddings:
                    raise ValueError(f"No layers found matching: {self._hook_outputs.keys()}")
                return aggregate_embeddings(embeddings, aggregation, wav.shape[0])
            finally:
                self._clear_hook_outputs()
                if was_training:
                    self.train()

    return Model


Model = register_model_class(make_model_class(ModelBase))
