"""CPU: the embedding cache layout.  avex_b200.extract.save_embeddings_arrays must write what the reference's
`save_embeddings_arrays` / `_write_embedding_metadata` write (avex/evaluation/embedding_utils.py:147-161, :1433-1580), and the
reference's `load_embeddings_arrays` (:1583-1678) must read it back.  h5py is absent from this image, so both sides run against
the dict-backed double in tests/doubles/h5py.py; the reference half is skipped where the reference is not importable."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture()
def fake_h5py(monkeypatch):
    monkeypatch.syspath_prepend(os.path.join(HERE, "doubles"))
    saved = sys.modules.pop("h5py", None)
    mod = importlib.import_module("h5py")
    assert mod.__file__.startswith(os.path.join(HERE, "doubles"))
    yield mod
    sys.modules.pop("h5py", None)
    if saved is not None:
        sys.modules["h5py"] = saved


def _data(multi=True):
    g = torch.Generator().manual_seed(5)
    labels = torch.tensor([3, 1, 0, 2, 2], dtype=torch.int32)
    if multi:
        emb = {"backbone.post_extract_proj": torch.randn(5, 48, 768, generator=g), "backbone.encoder.layers.11.fc2": torch.randn(5, 48, 768, generator=g)}
    else:
        emb = torch.randn(5, 1536, generator=g)
    return emb, labels


def test_npz_container_roundtrip(tmp_path):
    from avex_b200.extract import load_embeddings_arrays, save_embeddings_arrays

    for multi in (True, False):
        emb, labels = _data(multi)
        path = save_embeddings_arrays(emb, labels, tmp_path / f"split{multi}.h5", num_labels=4, aggregation="none")
        assert path.endswith(".npz")
        got, lab, n = load_embeddings_arrays(tmp_path / f"split{multi}.h5")
        assert n == 4 and lab.dtype == torch.int64 and torch.equal(lab, labels.long())
        if multi:
            assert list(got.keys()) == list(emb.keys()) and all(torch.equal(got[k], emb[k]) for k in emb)
        else:
            assert torch.equal(got, emb)


def test_hdf5_branch_attrs_are_native(tmp_path, fake_h5py):
    from avex_b200.extract import load_embeddings_arrays, save_embeddings_arrays

    emb, labels = _data(True)
    path = save_embeddings_arrays(emb, labels, tmp_path / "a.h5", num_labels=4, aggregation="mean")
    assert path.endswith("a.h5")
    f = fake_h5py.File(path, "r")
    assert sorted(f.keys()) == sorted(["labels"] + [f"embeddings_{k}" for k in emb])
    assert list(f.attrs["layer_names"]) == list(emb.keys())  # a list, not one JSON string
    assert list(f.attrs["embedding_dims"]) == ["(48, 768)", "(48, 768)"]
    assert list(f.attrs["stored_embedding_rank"]) == [2, 2]
    assert f.attrs["multi_layer"] is True and f.attrs["num_labels"] == 4 and f.attrs["extraction_complete"] is True
    assert f.attrs["aggregation"] == f.attrs["embedding_aggregation"] == "mean"
    assert f["labels"].dtype == np.int64 and all(f[f"embeddings_{k}"].dtype == np.float32 for k in emb)
    assert f.creation["labels"] == {"compression": "gzip", "compression_opts": 4}
    got, lab, n = load_embeddings_arrays(path)
    assert n == 4 and all(torch.equal(got[k], emb[k]) for k in emb)


def _reference_embedding_utils():
    """The reference's embedding_utils, imported from baseline/_ref or /root/reference with the absent cloud modules stubbed."""
    import types

    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "avex")):
            break
    else:
        pytest.skip("reference not available here")
    for name in ("gcsfs", "s3fs"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.GCSFileSystem = type("GCSFileSystem", (), {})
            m.S3FileSystem = type("S3FileSystem", (), {})
            sys.modules[name] = m
    if "esp_data" not in sys.modules:  # absent dev dependency: only its local-path helpers are reached here
        import pathlib

        pkg, io, paths = types.ModuleType("esp_data"), types.ModuleType("esp_data.io"), types.ModuleType("esp_data.io.paths")
        pkg.__path__, io.__path__ = [], []
        io.anypath, io.exists, io.filesystem_from_path = pathlib.Path, (lambda p: pathlib.Path(p).exists()), (lambda p: None)
        paths.PureCloudPath = type("PureCloudPath", (), {})
        pkg.io, io.paths = io, paths
        sys.modules.update({"esp_data": pkg, "esp_data.io": io, "esp_data.io.paths": paths})
    import importlib.metadata as md

    orig = md.version
    md.version = lambda n: "0.0.0+ref" if n == "avex" else orig(n)
    if cand not in sys.path:
        sys.path.insert(0, cand)
    try:
        for k in [k for k in sys.modules if k == "avex" or k.startswith("avex.")]:
            del sys.modules[k]  # re-import so that `import h5py` inside binds the double
        return importlib.import_module("avex.evaluation.embedding_utils")
    except Exception as e:  # pragma: no cover
        pytest.skip(f"reference not importable: {e}")


@pytest.mark.parametrize("multi", [True, False])
def test_same_file_as_reference_and_cross_loading(tmp_path, fake_h5py, multi):
    from avex_b200.extract import load_embeddings_arrays, save_embeddings_arrays

    ref = _reference_embedding_utils()
    emb, labels = _data(multi)
    ours = save_embeddings_arrays(emb, labels, tmp_path / "ours.h5", num_labels=4, aggregation="mean")
    ref.save_embeddings_arrays(emb, labels, tmp_path / "ref.h5", num_labels=4, aggregation="mean")
    a, b = fake_h5py.File(ours, "r"), fake_h5py.File(str(tmp_path / "ref.h5"), "r")
    assert sorted(a.keys()) == sorted(b.keys())
    for k in a.keys():
        assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]) and a.creation[k] == b.creation[k]
    assert set(dict.keys(a.attrs)) == set(dict.keys(b.attrs))
    for k in dict.keys(a.attrs):
        va, vb = dict.__getitem__(a.attrs, k), dict.__getitem__(b.attrs, k)
        assert type(va) is type(vb) and va == vb, (k, va, vb)
    # the reference loader reads our file; our loader reads the reference's
    got, lab, n = ref.load_embeddings_arrays(ours)
    got2, lab2, n2 = load_embeddings_arrays(tmp_path / "ref.h5")
    assert n == n2 == 4 and torch.equal(lab, labels.long()) and torch.equal(lab2, labels.long())
    if multi:
        assert all(torch.equal(got[k], emb[k]) and torch.equal(got2[k], emb[k]) for k in emb)
    else:
        assert torch.equal(got, emb) and torch.equal(got2, emb)


# ---- multi-rank feeder: world_size 2 on gloo --------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _embed(wav):  # deterministic per clip, independent of batch neighbours
    return torch.stack([wav.sum(dim=1), wav[:, 0] * 3.0, wav[:, -1]], dim=1)


def _cpu_loop(model, dataloader, target_layers, device, aggregation="mean", depth=3, disable_layerdrop=None):
    """Stand-in for the CUDA loop with its return contract (the gloo test covers sharding + gather + ordering only)."""
    embs, labels = [], []
    for b in dataloader:
        e = _embed(b["raw_wav"])
        embs.append(e if aggregation != "none" else e.unsqueeze(1).repeat(1, 4, 1))
        labels.append(b["label"])
    out = {"layer0": torch.cat(embs)}
    return out, torch.cat(labels), [tuple(out["layer0"].shape[1:])]


def _worker(rank, world, port, n_clips, aggregation, q):
    import torch.distributed as dist

    from avex_b200 import extract, parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = parallel.init_from_env("gloo")
    clips = torch.arange(n_clips * 6, dtype=torch.float32).reshape(n_clips, 6)
    idx = parallel.shard_indices(n_clips, r, w)
    batches = [{"raw_wav": clips[idx[i : i + 2]], "label": torch.tensor(idx[i : i + 2])} for i in range(0, len(idx), 2)]
    emb, labels, dims, gathered = extract.extract_embeddings_distributed(None, batches, [0], "cpu", aggregation=aggregation,
                                                                         num_samples=n_clips, loop=_cpu_loop)
    q.put((r, emb["layer0"], labels, gathered))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("aggregation", ["mean", "none"])
def test_distributed_feeder_world2_gloo(aggregation):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    for n_clips in (8, 7):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, aggregation, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
        clips = torch.arange(n_clips * 6, dtype=torch.float32).reshape(n_clips, 6)
        want = _embed(clips)
        for r, emb, labels, gathered in res:
            if aggregation == "mean":  # every rank holds the full arrays in clip order
                assert gathered and torch.equal(emb, want) and labels.tolist() == list(range(n_clips))
            else:  # frame-level outputs stay rank-local
                from avex_b200.parallel import shard_indices

                idx = shard_indices(n_clips, r, 2)
                assert not gathered and torch.equal(emb[:, 0], want[idx]) and labels.tolist() == idx
