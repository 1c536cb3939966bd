"""`ModelBase`: hook registry + embedding extraction shared by every model wrapper.

Behavioural mirror of avex/models/base_model.py:19-457 (written from its contract, pinned by the reference's
tests/unittests/test_base_model.py and test_base_model_all_layers.py): layer selectors may be names, 0-based
indices (negatives allowed; bool -> TypeError; out of range -> ValueError("... out of range ...")), "all" or
"last_layer"; duplicates collapse in first-seen order; hook outputs are kept un-detached; aggregation is
none | mean | max | cls_token over dim 1 with concatenation across layers.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional, Union

import torch
import torch.nn as nn

logger = logging.getLogger(__name__)

_HEAD_MARKERS = ("classifier", "head")


def _dedup(names: List[str]) -> List[str]:
    return list(dict.fromkeys(names))


def aggregate_embeddings(embeddings: List[torch.Tensor], aggregation: str, batch_size: int):
    """Shared tail of extract_embeddings (base_model.py:419-453 / beats_model.py:390-423)."""
    embeddings = [e if e.shape[0] == batch_size else e.transpose(0, 1) for e in embeddings]
    if aggregation == "none":
        return embeddings[0] if len(embeddings) == 1 else embeddings
    pooled = []
    for e in embeddings:
        if e.dim() == 2:
            pooled.append(e)
        elif e.dim() == 3:
            if aggregation == "mean":
                pooled.append(e.mean(dim=1))
            elif aggregation == "max":
                pooled.append(e.max(dim=1)[0])
            elif aggregation == "cls_token":
                pooled.append(e[:, 0, :])
            else:
                raise ValueError(f"Unsupported aggregation method: {aggregation}")
        else:
            raise ValueError(f"Unexpected embedding dimension: {e.dim()}. Expected 2 or 3.")
    return pooled[0] if len(pooled) == 1 else torch.cat(pooled, dim=1)


class ModelBase(nn.Module):
    def __init__(self, device: str, audio_config: Optional[Any] = None) -> None:
        super().__init__()
        self.device = device
        self.audio_processor = None
        if audio_config:
            from .audio import AudioProcessor  # only the raw / mel representations the hot path uses

            self.audio_processor = AudioProcessor(audio_config)
        self._hooks: Dict[str, torch.utils.hooks.RemovableHandle] = {}
        self._hook_outputs: Dict[str, torch.Tensor] = {}
        self._layer_names: List[str] = []
        self._hook_layers: List[str] = []

    # ---- layer discovery ---------------------------------------------------------------------------------------
    def _discover_embedding_layers(self) -> None:
        if not self._layer_names:
            self._layer_names = [n for n, m in self.named_modules() if isinstance(m, nn.Linear)]

    def get_model_layers(self) -> list[str]:
        self._discover_embedding_layers()
        return list(self._layer_names)

    def get_model_layer_map(self) -> dict[int, str]:
        return dict(enumerate(self.get_model_layers()))

    def _get_last_non_classification_layer(self) -> Optional[str]:
        for name in reversed(self._layer_names):
            if not any(m in name.lower() for m in _HEAD_MARKERS):
                return name
        return self._layer_names[-1] if self._layer_names else None

    # ---- hooks ------------------------------------------------------------------------------------------------
    def _create_hook_fn(self, layer_name: str):
        def hook_fn(module, inputs, output):
            if isinstance(output, dict):
                output = output["x"]
            elif isinstance(output, tuple):
                output = output[0]
            self._hook_outputs[layer_name] = output  # not detached: gradients may flow (base_model.py:89)

        return hook_fn

    def register_hooks_for_layers(self, target_layers: List[Union[str, int]]) -> List[str]:
        self._discover_embedding_layers()
        names: List[str] = []
        for sel in target_layers:
            if isinstance(sel, bool):
                raise TypeError("target_layers entries must be str or int (bool is not allowed).")
            if isinstance(sel, int):
                n = len(self._layer_names)
                if not -n <= sel < n:
                    raise ValueError(
                        f"Layer index {sel} is out of range for {n} layers "
                        f"(valid indices: 0..{n - 1} and negative indices like -1)."
                    )
                names.append(self._layer_names[sel])
            else:
                names.append(sel)
        if "all" in names:
            names = _dedup([n for n in names if n != "all"] + list(self._layer_names))
        if "last_layer" in names:
            last = self._get_last_non_classification_layer()
            if not last:
                raise ValueError("No layers available for 'last_layer'")
            names = [last if n == "last_layer" else n for n in names]
        names = _dedup(names)
        self.deregister_all_hooks()
        self._hook_layers = names
        for name in names:
            try:
                module = self.get_submodule(name)
            except AttributeError as err:
                raise ValueError(f"Layer '{name}' not found in model") from err
            self._hooks[name] = module.register_forward_hook(self._create_hook_fn(name))
        return names

    def ensure_hooks_registered(self) -> None:
        if not self._hooks and self._hook_layers:
            self.register_hooks_for_layers(self._hook_layers)

    def deregister_all_hooks(self) -> None:
        for handle in self._hooks.values():
            handle.remove()
        self._hooks.clear()
        self._hook_outputs.clear()  # _hook_layers is kept on purpose (base_model.py:225-226)

    def _clear_hook_outputs(self) -> None:
        self._hook_outputs.clear()

    def _cleanup_hooks(self) -> None:
        self.deregister_all_hooks()

    def __del__(self) -> None:
        try:
            self._cleanup_hooks()
        except Exception:
            pass

    # ---- audio -------------------------------------------------------------------------------------------------
    def process_audio(self, x: torch.Tensor) -> torch.Tensor:
        if x is None:
            raise ValueError("Input tensor cannot be None")
        if self.audio_processor is not None:
            x = self.audio_processor(x)
        return x.to(next(self.parameters()).device)

    def enable_gradient_checkpointing(self) -> None:
        raise NotImplementedError(
            f"{self.__class__.__name__} does not support gradient checkpointing. "
            f"Please implement the enable_gradient_checkpointing method."
        )

    # ---- embeddings --------------------------------------------------------------------------------------------
    def extract_embeddings(self, x, *, padding_mask: Optional[torch.Tensor] = None, aggregation: str = "none"):
        self._clear_hook_outputs()
        self.ensure_hooks_registered()
        if not self._hooks:
            raise ValueError("No hooks registered. Call register_hooks_for_layers() first.")
        try:
            wav, mask = (x["raw_wav"], x.get("padding_mask")) if isinstance(x, dict) else (x, padding_mask)
            self.forward(wav, mask)
            if self._hook_layers:
                missing = [n for n in self._hook_layers if n not in self._hook_outputs]
                if missing and not self._hook_outputs:
                    raise ValueError(f"No layers found matching: {missing}")
                if missing:
                    raise ValueError(
                        f"Some requested layers did not produce hook outputs: {missing}. "
                        f"Available outputs: {list(self._hook_outputs.keys())}"
                    )
                order = self._hook_layers
            else:
                order = list(self._hook_outputs.keys())
            embeddings = [self._hook_outputs[n] for n in order]
            if not embeddings:
                raise ValueError(f"No layers found matching: {self._hook_outputs.keys()}")
            return aggregate_embeddings(embeddings, aggregation, wav.shape[0])
        finally:
            self._clear_hook_outputs()
