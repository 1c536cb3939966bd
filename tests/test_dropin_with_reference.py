"""CPU, build container only: with the real reference importable, the drop-in class registers into avex's own
registry and is constructed by avex.load_model with state_dict keys / shapes / layer names identical to the reference's.
Skipped where /root/reference is absent (the GPU box)."""
import os
import sys

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "avex")), reason="reference tree not present")


def test_state_dict_and_layers_identical_to_reference():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import ref_shim

    avex = ref_shim.install()
    from avex.models.base_model import ModelBase
    from avex.models.utils.registry import get_model_class, register_model_class

    ref_cls = get_model_class("beats")
    try:
        spec = avex.get_model_spec("esp_aves2_sl_beats_all").model_copy(deep=True)
        spec.device = "cpu"
        avex.register_model("dropin_ref", spec)
        ref = avex.load_model("dropin_ref", device="cpu", return_features_only=True)

        from avex_b200.integrate import make_avex_model_class

        register_model_class(make_avex_model_class("beats"))  # overwrite, as registry.py:616-618 allows
        mine = avex.load_model("dropin_ref", device="cpu", return_features_only=True)
        assert isinstance(mine, ModelBase) and type(mine).__module__.startswith("avex_b200")
        a, b = ref.state_dict(), mine.state_dict()
        assert set(a) == set(b), set(a) ^ set(b)
        for k in a:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
        assert ref.get_model_layers() == mine.get_model_layers()
        assert ref.register_hooks_for_layers(["last_layer", 0, "all"]) == mine.register_hooks_for_layers(["last_layer", 0, "all"])
        missing, unexpected = mine.load_state_dict(a, strict=False)
        assert not missing and not unexpected
        with pytest.raises(ValueError, match="out of range"):
            mine.register_hooks_for_layers([99])
        with pytest.raises(TypeError):
            mine.register_hooks_for_layers([True])
        with pytest.raises(ValueError, match="not found"):
            mine.register_hooks_for_layers(["nope.layer"])
    finally:
        register_model_class(ref_cls)
        # the reference ships its own top-level `tests` package: take it off the path again
        for pth in (REF, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")):
            while pth in sys.path:
                sys.path.remove(pth)
