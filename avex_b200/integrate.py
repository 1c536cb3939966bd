"""Registration into the real `avex` package when it is importable."""
from __future__ import annotations


def make_avex_model_class(name: str = "beats"):
    """The BEATs drop-in built on `avex.models.base_model.ModelBase`, named `name` for avex's class registry."""
    from avex.models.base_model import ModelBase as AvexModelBase

    from .plugin.beats_model import make_model_class

    cls = make_model_class(AvexModelBase)
    cls.name = name
    return cls


def install(override: bool = False):
    """Register the class with avex: as "beats_b200", or over "beats" (so `load_model("esp_aves2_sl_beats_all")` uses it)."""
    from avex.models.utils.registry import register_model_class

    return register_model_class(make_avex_model_class("beats" if override else "beats_b200"))
