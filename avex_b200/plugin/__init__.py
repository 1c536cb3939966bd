"""Host-side mirror of avex's plugin surface for the hot path (same names, argument meaning, error behaviour).

Reference: avex/__init__.py:9-51, avex/models/base_model.py, avex/models/utils/{registry,factory,load}.py,
avex/configs.py (`ModelSpec`, `AudioConfig`).  When the real `avex` package is importable, use
`avex_b200.integrate.install()` instead and keep using avex's own registry / load_model.
"""
from .base_model import ModelBase
from .configs import AudioConfig, ModelSpec
from .factory import build_model, build_model_from_spec
from .load import load_model
from .registry import (
    get_model_class,
    get_model_spec,
    list_model_classes,
    list_models,
    register_model,
    register_model_class,
)

__all__ = [
    "ModelBase", "AudioConfig", "ModelSpec", "build_model", "build_model_from_spec", "load_model",
    "get_model_class", "get_model_spec", "list_model_classes", "list_models", "register_model", "register_model_class",
]  # fmt: skip
