"""Two module-level dictionaries, like the reference (avex/models/utils/registry.py:27-30): model name ->
ModelSpec and model type -> class.  Single-threaded use is assumed, as upstream."""
from __future__ import annotations

import logging
from typing import Dict, Optional, Type

from .configs import ModelSpec

logger = logging.getLogger(__name__)

_MODEL_REGISTRY: Dict[str, ModelSpec] = {}
_MODEL_CLASSES: Dict[str, Type] = {}


def register_model(name: str, model_spec: ModelSpec) -> None:
    """registry.py:296-308: later registrations overwrite, with a warning."""
    if name in _MODEL_REGISTRY:
        logger.warning(f"Model '{name}' is already registered. Overwriting with new configuration.")
    _MODEL_REGISTRY[name] = model_spec


def get_model_spec(name: str) -> Optional[ModelSpec]:
    return _MODEL_REGISTRY.get(name)


def list_models() -> Dict[str, dict]:
    return {k: {"model_type": v.name, "pretrained": v.pretrained, "device": v.device} for k, v in _MODEL_REGISTRY.items()}


def register_model_class(cls: Type) -> Type:
    """registry.py:600-621: key = `cls.name` if present, else the lower-cased class name; usable as a decorator."""
    key = getattr(cls, "name", None) or cls.__name__.lower()
    if key in _MODEL_CLASSES:
        logger.warning(f"Model class '{key}' is already registered. Overwriting.")
    _MODEL_CLASSES[key] = cls
    return cls


def get_model_class(name: str) -> Optional[Type]:
    return _MODEL_CLASSES.get(name)


def list_model_classes() -> list[str]:
    return list(_MODEL_CLASSES.keys())
