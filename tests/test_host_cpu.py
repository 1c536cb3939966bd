"""CPU: host-side logic of avex_b200 that needs no GPU (tables, C-ABI surface, loud failure without CUDA)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import avex_b200
from avex_b200 import _lib
from avex_b200.fbank import KaldiFbank

G = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tables_bit_exact_with_reference():
    t = np.load(os.path.join(G, "fbank_tables.npz"))
    fb = KaldiFbank()
    assert np.array_equal(fb.window.numpy(), t["window"])
    assert np.array_equal(fb.mel_fb.numpy(), t["mel_fb"])
    assert set(fb.state_dict().keys()) == {"window", "mel_fb"}  # reference buffer names (beats.py:76,80)


def test_fbank_geometry_guard():
    with pytest.raises(ValueError):
        KaldiFbank(num_mel_bins=64)
    fb = KaldiFbank()
    for n in (4000, 8000, 16000, 32000, 160000):
        assert fb.num_frames(n) == 1 + (n - 400) // 160


def test_library_exports_every_declared_symbol():
    """include/avexk.h is the contract: every `avexk_*` function it declares must be exported and bound."""
    header = open(os.path.join(ROOT, "include", "avexk.h")).read()
    declared = set(re.findall(r"\b(avexk_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)  # built by __graft_entry__.build() / python -m avex_b200.build
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in avexk.h but not exported by libavexk.so"
    _lib.load()
    assert _lib.load().avexk_version() >= 100
    assert _lib.load().avexk_fbank_num_frames(160000) == 998
    assert _lib.load().avexk_beats_num_tokens(160000) == 496


def test_no_cpu_fallback():
    fb = KaldiFbank()
    with pytest.raises(_lib.AvexkError):
        fb(torch.zeros(1, 16000))


@pytest.mark.parametrize("n_tokens", [248, 496, 2992])
def test_product_relpos_bucket_and_bias_vector_against_reference_golden(n_tokens):
    """`avex_b200.beats.relative_position_bucket` / `relative_bias_vector` are what the CUDA attention consumes; hold THEM (not
    only the oracle's copy) to the reference's `_relative_positions_bucket` output over offsets -3100..3100 (backbone.py:438-473),
    including the log-bucket region (|d| >= 80) and the saturation at max_distance = 800 that only 60 s clips reach."""
    from avex_b200.beats import relative_bias_vector, relative_position_bucket

    g = np.load(os.path.join(G, "relpos_buckets.npz"))
    rel, want = torch.from_numpy(g["rel"].astype(np.int64)), g["bucket"].astype(np.int64)
    got = relative_position_bucket(rel, 320, 800).numpy()
    assert np.array_equal(got, want)
    assert got[rel.numpy() == 800].tolist() == [319] and got[rel.numpy() == -3100].tolist() == [159]
    table = torch.arange(320 * 12, dtype=torch.float32).reshape(320, 12)
    vec = relative_bias_vector(table, n_tokens, 320, 800).numpy()
    assert vec.shape == (12, 2 * n_tokens - 1)
    lo = 3100 - (n_tokens - 1)
    idx = want[lo : lo + 2 * n_tokens - 1]  # offsets -(N-1) .. N-1
    assert np.array_equal(vec, table.numpy()[idx].T)
    # Toeplitz property the kernel relies on: bias[h, i, j] = vec[h, j - i + N - 1] equals the reference's [N, N] gather
    i, j = 3, n_tokens - 2
    assert vec[5, j - i + n_tokens - 1] == table[want[3100 + (j - i)], 5]


def test_loader_refuses_stale_binary(tmp_path, monkeypatch):
    """The .so is git-ignored and ships with the snapshot: `_lib.load()` must notice a binary built from other sources."""
    from avex_b200 import build

    assert build.binary_id() == build.source_id() == _lib.load().avexk_build_id().decode()
    monkeypatch.setattr(build, "source_id", lambda: "0" * 32)
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.AvexkError, match="stale"):
        _lib.load()


def test_hook_selectors_outside_the_fused_path_are_refused():
    """names the reference would accept but the fused forward never materialises must fail at registration."""
    from avex_b200 import plugin
    from avex_b200.plugin import beats_model  # noqa: F401

    plugin.register_model("cpu_hook_test", plugin.ModelSpec(name="beats", device="cpu", init_config=dict(encoder_layers=2)))
    model = plugin.load_model("cpu_hook_test", device="cpu", return_features_only=True)
    assert model.register_hooks_for_layers([0, -1]) == ["backbone.post_extract_proj", "backbone.encoder.layers.1.fc2"]
    for bad in ("backbone.encoder.layers.1", "backbone.encoder.layers.0.fc1", "backbone.encoder.layer_norm"):
        with pytest.raises(ValueError, match="cannot serve forward hooks"):
            model.register_hooks_for_layers([bad])
        assert not model._hooks and not model._hook_layers
    with pytest.raises(ValueError, match="not found"):
        model.register_hooks_for_layers(["backbone.nope"])


def test_model_copies_and_pickles_without_native_handles():
    import copy
    import pickle

    from avex_b200.beats import BEATs, BEATsConfig

    m = BEATs(BEATsConfig(encoder_layers=1))
    m._engine, m._engine_key = ctypes.c_void_p(1234), ("x",)  # what a forward leaves behind
    m.fbank._handle = ctypes.c_void_p(99)
    c = copy.deepcopy(m)
    assert c._engine is None and c.fbank._handle is None
    assert torch.equal(c.post_extract_proj.weight, m.post_extract_proj.weight)
    assert c.encoder.layers[0].self_attn.relative_attention_bias is c.encoder.layers[0].self_attn.relative_attention_bias
    # whole-module pickling is refused by torch itself for the weight-norm parametrised pos-conv (the reference has the same
    # limit); the pieces that hold native handles must not be what blocks it
    f = pickle.loads(pickle.dumps(m.fbank))
    assert f._handle is None and torch.equal(f.window, m.fbank.window)
    assert m.__getstate__()["_engine"] is None
    m._engine, m.fbank._handle = None, None  # nothing real to destroy


def test_checkpoint_key_handling_and_streamed_safetensors(tmp_path):
    """checkpoints load with the reference's key rules (utils/utils.py:509-570, load.py:521-570) -- `module.` / `model.`
    prefixes, head layers dropped in features-only mode, `backbone.` added or removed to fit -- and a .safetensors file is streamed
    tensor by tensor into the parameters (same result as reading it whole)."""
    from safetensors.torch import save_file

    from avex_b200 import plugin
    from avex_b200.plugin import beats_model  # noqa: F401
    from avex_b200.plugin.load import _resolve_keys

    target = ["backbone.post_extract_proj.weight", "backbone.encoder.layers.0.fc1.weight", "classifier.weight"]
    got = _resolve_keys(["module.post_extract_proj.weight", "model.encoder.layers.0.fc1.weight", "classifier.weight", "head.bias"],
                        target, keep_classifier=False)
    assert got == {"module.post_extract_proj.weight": "backbone.post_extract_proj.weight",
                   "model.encoder.layers.0.fc1.weight": "backbone.encoder.layers.0.fc1.weight"}
    # a target whose own keys start with `model.` (EfficientNet wrapper) keeps that prefix
    assert _resolve_keys(["model.features.0.0.weight"], ["model.features.0.0.weight"], True) == {"model.features.0.0.weight": "model.features.0.0.weight"}
    assert _resolve_keys(["backbone.x.weight"], ["x.weight"], True) == {"backbone.x.weight": "x.weight"}

    spec = plugin.ModelSpec(name="beats", device="cpu", init_config=dict(encoder_layers=1, finetuned_model=True))
    plugin.register_model("cpu_ckpt_test", spec)
    torch.manual_seed(3)
    src = plugin.load_model("cpu_ckpt_test", device="cpu", return_features_only=True)
    sd = {("module." + k[len("backbone."):]): v.detach().clone().contiguous() for k, v in src.state_dict().items()
          if "relative_attention_bias" not in k or ".layers.0." in k}  # DDP-style names, no `backbone.` prefix
    sd["module.classifier.weight"] = torch.zeros(3, 768)  # must be dropped in features-only mode
    path = str(tmp_path / "ckpt.safetensors")
    save_file(sd, path)
    torch.manual_seed(4)
    dst = plugin.load_model("cpu_ckpt_test", device="cpu", checkpoint_path=path, return_features_only=True)
    for k, v in src.state_dict().items():
        assert torch.equal(dst.state_dict()[k], v), k
    pt = str(tmp_path / "ckpt.pt")
    torch.save({"model_state_dict": sd}, pt)
    torch.manual_seed(5)
    dst2 = plugin.load_model("cpu_ckpt_test", device="cpu", checkpoint_path=pt, return_features_only=True)
    for k, v in src.state_dict().items():
        assert torch.equal(dst2.state_dict()[k], v), k
