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


_HEAD_TERMS = ("classifier", "head", "classification", "classification_head")


def _read_checkpoint(path: str) -> dict:
    if str(path).endswith(".safetensors"):
        from safetensors.torch import load_file

        return load_file(str(path), device="cpu")
    obj = torch.load(str(path), map_location="cpu", weights_only=True)
    if isinstance(obj, dict):  # utils/utils.py:536-540 wrappers
        if "model_state_dict" in obj:
            obj = obj["model_state_dict"]
        elif "model" in obj and isinstance(obj["model"], dict):
            obj = obj["model"]
        elif "state_dict" in obj and isinstance(obj["state_dict"], dict):
            obj = obj["state_dict"]
    return obj


def _map_key(key: str, keep_classifier: bool, drop_model_prefix: bool) -> Optional[str]:
    """`_process_state_dict` (utils/utils.py:509-570) for one key: strip `module.` or (optionally) `model.`, drop head layers."""
    if not keep_classifier and key in ("classifier.weight", "classifier.bias", "model.classifier.1.weight", "model.classifier.1.bias"):
        return None
    if key.startswith("module."):
        key = key[7:]
    elif drop_model_prefix and key.startswith("model."):
        key = key[6:]
    if not keep_classifier and any(t in key.lower() for t in _HEAD_TERMS):
        return None
    return key


def _resolve_keys(keys, target_keys, keep_classifier: bool) -> dict:
    """checkpoint key -> model key, with the `backbone.` prefix adapted either way (load.py:548-556)."""
    drop_model = not any(k.startswith("model.") for k in target_keys)
    mapped = {}
    for k in keys:
        m = _map_key(k, keep_classifier, drop_model)
        if m is not None:
            mapped[k] = m
    wants = any(k.startswith("backbone.") for k in target_keys)
    has = any(m.startswith("backbone.") for m in mapped.values())
    if wants and not has:
        mapped = {k: "backbone." + m for k, m in mapped.items()}
    elif has and not wants:
        mapped = {k: (m[len("backbone."):] if m.startswith("backbone.") else m) for k, m in mapped.items()}
    return mapped


def _stream_safetensors(model: torch.nn.Module, path: str, keep_classifier: bool):
    """safetensors -> the model's (device) parameters, one tensor at a time straight from the memory-mapped file: the whole
    checkpoint is never materialised as a CPU state dict (the reference reads everything, then `load_state_dict`s).
    Same key handling and strict=False semantics as `_load_checkpoint`; returns (missing, unexpected)."""
    from safetensors import safe_open

    own = dict(model.state_dict(keep_vars=True))  # parameters and buffers, by name
    seen, unexpected = set(), []
    with safe_open(str(path), framework="pt", device="cpu") as f:
        mapped = _resolve_keys(list(f.keys()), own.keys(), keep_classifier)
        with torch.no_grad():
            for ck, mk in mapped.items():
                dst = own.get(mk)
                if dst is None:
                    unexpected.append(mk)
                    continue
                src = f.get_tensor(ck)
                if tuple(src.shape) != tuple(dst.shape):
                    raise RuntimeError(f"size mismatch for {mk}: checkpoint {tuple(src.shape)} vs model {tuple(dst.shape)}")
                dst.copy_(src, non_blocking=True)  # host (mmap) -> device, converted to the parameter's dtype
                seen.add(mk)
    for m in model.modules():  # in-place copies do not bump tensor versions: packed weight copies must be rebuilt
        if hasattr(m, "invalidate"):
            m.invalidate()
    eng = getattr(model, "_engine", None)
    if eng is not None and hasattr(eng, "invalidate"):
        eng.invalidate()
    return [k for k in own if k not in seen], unexpected


def _load_checkpoint(model: torch.nn.Module, path: str, keep_classifier: bool) -> None:
    """load.py:521-570: unwrap, strip DDP prefixes, drop head keys in features-only mode, fix the `backbone.` prefix, strict=False."""
    if str(path).endswith(".safetensors"):
        missing, unexpected = _stream_safetensors(model, path, keep_classifier)
    else:
        raw = _read_checkpoint(path)
        mapped = _resolve_keys(list(raw.keys()), model.state_dict().keys(), keep_classifier)
        missing, unexpected = model.load_state_dict({m: raw[k] for k, m in mapped.items()}, strict=False)
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
