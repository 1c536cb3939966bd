"""GPU: the evaluation-style extraction loop (avex/evaluation/embedding_utils.py:26-144 contract) on top of the plugin model."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import beats_encoder as OE
from oracle.weights import make_beats_weights

pytestmark = pytest.mark.gpu


def _model(layers=2):
    from avex_b200 import plugin
    from avex_b200.plugin import beats_model  # noqa: F401

    plugin.register_model("extract_beats", plugin.ModelSpec(name="beats", device="cuda", init_config=dict(encoder_layers=layers)))
    model = plugin.load_model("extract_beats", device="cuda", return_features_only=True).eval()
    W = make_beats_weights(OE.BeatsDims(layers=layers), seed=1)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=False)
    return model


def _batches(n_batches, bs, T, with_mask):
    g = torch.Generator().manual_seed(11)
    out = []
    for i in range(n_batches):
        b = {"raw_wav": torch.randn(bs, T, generator=g) * 0.1, "label": torch.arange(i * bs, (i + 1) * bs)}
        if with_mask:
            m = torch.zeros(bs, T, dtype=torch.bool)
            m[0, T // 2:] = True
            b["padding_mask"] = m
        out.append(b)
    return out


@pytest.mark.parametrize("aggregation,with_mask", [("mean", False), ("none", False), ("mean", True)])
def test_extraction_loop_matches_direct_calls(tmp_path, aggregation, with_mask):
    from avex_b200.extract import extract_embeddings_for_split, load_embeddings_arrays, save_embeddings_arrays

    model = _model()
    batches = _batches(5, 3, 16000, with_mask)
    emb, labels, dims = extract_embeddings_for_split(model, batches, [0, -1], "cuda", aggregation=aggregation, depth=2)
    assert not model._hooks  # deregistered on exit, like the reference loop
    assert torch.equal(labels, torch.arange(15))
    # direct, blocking calls for comparison
    names = model.register_hooks_for_layers([0, -1])
    ref = []
    with torch.no_grad():
        for b in batches:
            x = b["raw_wav"].cuda()
            inp = {"raw_wav": x, "padding_mask": b["padding_mask"].cuda()} if with_mask else x
            ref.append(model.extract_embeddings(inp, aggregation=aggregation))
    model.deregister_all_hooks()
    if aggregation == "mean":
        assert list(emb.keys()) == [names[0]]
        want = torch.cat([r.cpu() for r in ref])
        assert emb[names[0]].shape == (15, 2 * 768) and dims == [(1536,)]
        assert torch.equal(emb[names[0]], want)
    else:
        assert list(emb.keys()) == names
        for li, n in enumerate(names):
            want = torch.cat([r[li].cpu() for r in ref])
            assert torch.equal(emb[n], want) and emb[n].shape[1:] == (48, 768)
    path = save_embeddings_arrays(emb, labels, str(tmp_path / "split"), num_labels=15, aggregation=aggregation)
    assert os.path.exists(path)
    if path.endswith(".npz"):
        z = np.load(path)
        attrs = json.loads(bytes(z["__attrs__"]).decode())
        assert attrs["extraction_complete"] and attrs["layer_names"] == list(emb.keys()) and attrs["multi_layer"] is True
        assert attrs["embedding_dims"] == [str(tuple(emb[n].shape[1:])) for n in emb] and attrs["num_labels"] == 15
        assert attrs["stored_embedding_rank"] == [emb[n].dim() - 1 for n in emb]
        assert z["labels"].shape == (15,) and z["labels"].dtype == np.int64
        assert all(z[f"embeddings_{n}"].dtype == np.float32 for n in emb)
    back, lab, nl = load_embeddings_arrays(path)
    assert nl == 15 and torch.equal(lab, labels) and all(torch.equal(back[n], emb[n]) for n in emb)


def test_extraction_loop_errors():
    from avex_b200.extract import extract_embeddings_for_split

    model = _model()
    with pytest.raises(ValueError):
        extract_embeddings_for_split(model, [], [0], "cuda")
    with pytest.raises(ValueError):
        extract_embeddings_for_split(model, [], [0], "cpu")
    assert not model._hooks
