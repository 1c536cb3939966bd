"""`build_model` / `build_model_from_spec` (avex/models/utils/factory.py:19-166)."""
from __future__ import annotations

import inspect

from .configs import AudioConfig, ModelSpec
from .registry import get_model_class, get_model_spec, list_model_classes, list_models

# ModelSpec fields forwarded to the constructor when set (factory.py:30-46); note `pretrained` is NOT among them.
_FORWARDED = ("efficientnet_variant", "use_naturelm", "fine_tuned", "init_config")


def build_model_from_spec(model_spec: ModelSpec, device: str, **kwargs):
    cls = get_model_class(model_spec.name)
    if cls is None:
        raise KeyError(f"Model class '{model_spec.name}' is not registered. Available classes: {list_model_classes()}")
    audio_config = model_spec.audio_config
    if isinstance(audio_config, dict):
        audio_config = AudioConfig(**audio_config)
    init_kwargs = {"device": device, "audio_config": audio_config or None, **kwargs}
    for name in _FORWARDED:
        value = getattr(model_spec, name, None)
        if value is not None and value != "":
            init_kwargs[name] = value
    accepted = set(inspect.signature(cls.__init__).parameters)  # factory.py:152-154
    return cls(**{k: v for k, v in init_kwargs.items() if k in accepted})


def build_model(model_name: str, device: str, **kwargs):
    spec = get_model_spec(model_name)
    if spec is None:
        raise KeyError(f"Model '{model_name}' is not registered. Available models: {list(list_models().keys())}")
    return build_model_from_spec(spec, device, **kwargs)
