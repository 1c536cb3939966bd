"""`load_model` (avex/models/utils/load.py:35-150, :152-330, :521-570) for registered names and ModelSpec objects."""
from __future__ import annotations

import inspect
import logging
from pathlib import Path
from typing import Optional, Union

import torch

from .configs import ModelSpec
from .factory import build_model_from_spec
from .registry import get_model_class, get_model_spec, list_models

logger = logging.getLogger(__name__)


def _read_checkpoint(path: str) -> dict:
    if str(path).endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(str(path), device="cpu")
    obj = torch.load(str(path), map_location="cpu", weights_only=True)
    for key in ("model_state_dict", "state_dict", "model"):  # utils/utils.py:509-570 wrappers
        if isinstance(obj, dict) and key in obj and isinstance(obj[key], dict):
            obj = obj[key]
    return obj


def _load_checkpoint(model: torch.nn.Module, path: str, keep_classifier: bool) -> None:
    """load.py:521-570: strip DDP prefixes, drop classifier keys in features-only mode, fix the `backbone.` prefix, strict=False."""
    sd = {}
    for k, v in _read_checkpoint(path).items():
        for pre in ("module.", "model."):
            if k.startswith(pre):
                k = k[len(pre):]
        if not keep_classifier and k.startswith(("classifier.", "head.")):
            continue
        sd[k] = v
    target = set(model.state_dict().keys())
    wants_prefix = any(k.startswith("backbone.") for k in target)
    has_prefix = any(k.startswith("backbone.") for k in sd)
    if wants_prefix and not has_prefix:
        sd = {"backbone." + k: v for k, v in sd.items()}
    elif has_prefix and not wants_prefix:
        sd = {k[len("backbone."):] if k.startswith("backbone.") else k: v for k, v in sd.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    logger.info(f"Loaded checkpoint {path}: {len(missing)} missing, {len(unexpected)} unexpected keys")


def _load_from_modelspec(spec: ModelSpec, device: str, checkpoint_path: Optional[str], return_features_only: bool):
    cls = get_model_class(spec.name)
    if cls is None:
        raise KeyError(f"Model class '{spec.name}' is not registered.")
    kwargs = {}
    if "return_features_only" in inspect.signature(cls.__init__).parameters:  # load.py:215-220
        kwargs["return_features_only"] = return_features_only
    if checkpoint_path and not return_features_only:
        sd = _read_checkpoint(checkpoint_path)
        for key in ("classifier.weight", "module.classifier.weight", "model.classifier.weight"):  # num_classes sniffing
            if key in sd:
                kwargs["num_classes"] = int(sd[key].shape[0])
    model = build_model_from_spec(spec, device, **kwargs)
    if checkpoint_path:
        _load_checkpoint(model, checkpoint_path, keep_classifier=not return_features_only)
    return model.to(device)


def load_model(
    model: Union[str, Path, ModelSpec],
    device: str = "cpu",
    checkpoint_path: Optional[str] = None,
    return_features_only: bool = False,
):
    if isinstance(model, Path):
        model = str(model)
    if isinstance(model, str):
        spec = get_model_spec(model)
        if spec is None:
            raise ValueError(
                f"Unknown model identifier: '{model}'. Available models: {list(list_models().keys())}. "
                "Or provide a path to a YAML config file."
            )
        return _load_from_modelspec(spec, device, checkpoint_path, return_features_only)
    if isinstance(model, ModelSpec):
        return _load_from_modelspec(model, device, checkpoint_path, return_features_only)
    raise TypeError(f"Unsupported model type: {type(model)}. Expected str, Path, or ModelSpec.")
