"""Embedding-extraction loop over a dataloader: the caller of the hot path in avex's evaluation pipeline
(avex/evaluation/embedding_utils.py:26-144, `_extract_embeddings_in_memory`).

Same contract -- batches are dicts with `raw_wav`, optional `padding_mask`, `label`; hooks are registered once for
`target_layers`, every batch goes through `model.extract_embeddings(..., aggregation=...)`, the result is
`(embeddings: {layer_name: tensor on CPU}, labels, embedding_dims)` and hooks are deregistered on exit -- but the per-batch
blocking `.cpu()` of the reference (embedding_utils.py:107,114,122) is replaced by an asynchronous device->host ring (and the
per-batch `.to(device)` by a one-batch-ahead host->device prefetch on a side stream, `_DevicePrefetcher`): each
batch's embeddings are copied into pinned host buffers on a side stream and only collected `depth` batches later, so the
D2H transfer and the host-side concatenation overlap the next batches' kernels.  At B200 speed (a 256 x 10 s batch every
30 ms) the blocking copy is otherwise the bottleneck the moment frame-level (`aggregation="none"`) outputs are kept.

`save_embeddings_arrays` / `load_embeddings_arrays` have the reference's signatures and on-disk layout (embedding_utils.py:147-161,
:1433-1678): float32 datasets `embeddings_{layer}` (or `embeddings`), int64 `labels`, attrs `embedding_aggregation`,
`aggregation`, `stored_embedding_rank` (list), `layer_names` (list), `embedding_dims` (list of `str(tuple)`), `multi_layer`,
`num_labels`, `extraction_complete` -- HDF5 when h5py is importable, else an `.npz` with the same names (h5py is not in this
image; tests/test_embedding_io_cpu.py drives the HDF5 branch and the reference's own loader through a dict-backed h5py double).
`extract_embeddings_distributed` is the multi-GPU feeder: the same loop per rank over a sharded dataloader + one all-gather.
"""
from __future__ import annotations

import json
import logging
import os
from typing import Dict, List, Optional, Tuple

import torch

logger = logging.getLogger(__name__)


class _PinnedRing:
    """`depth` slots of pinned host tensors; `push` enqueues non-blocking copies on a side stream, `pop` waits for the oldest."""

    def __init__(self, device: torch.device, depth: int) -> None:
        self.device, self.depth = device, depth
        self.stream = torch.cuda.Stream(device=device)
        self.slots: List[Tuple[List[torch.Tensor], torch.cuda.Event]] = []

    def push(self, tensors: List[torch.Tensor]) -> None:
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))  # the producing kernels
        host = []
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            for t in tensors:
                t.record_stream(self.stream)
                h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                h.copy_(t, non_blocking=True)
                host.append(h)
            done = torch.cuda.Event()
            done.record(self.stream)
        self.slots.append((host, done))

    def pop(self) -> List[torch.Tensor]:
        host, done = self.slots.pop(0)
        done.synchronize()
        return host

    def __len__(self) -> int:
        return len(self.slots)


class _DevicePrefetcher:
    """Iterates a dataloader one batch ahead: the NEXT batch's `raw_wav` / `padding_mask` go host -> device on a side stream,
    enqueued before the caller launches the current batch's kernels, so the transfer (3 ms per 256 x 10 s batch) overlaps
    the forward instead of preceding it.  Overlap needs pinned host tensors (`DataLoader(pin_memory=True)`); pageable ones
    are copied synchronously by torch, as in the reference loop (embedding_utils.py:96-101)."""

    _KEYS = ("raw_wav", "padding_mask")

    def __init__(self, dataloader, device: torch.device) -> None:
        self.it, self.device = iter(dataloader), device
        self.stream = torch.cuda.Stream(device=device)
        self.next = None
        self._preload()

    def _preload(self) -> None:
        try:
            batch = next(self.it)
        except StopIteration:
            self.next = None
            return
        out = dict(batch)
        with torch.cuda.stream(self.stream):
            for k in self._KEYS:
                v = batch.get(k)
                if torch.is_tensor(v):
                    out[k] = v.to(self.device, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        self.next = (out, ready)

    def __iter__(self):
        return self

    def __next__(self):
        if self.next is None:
            raise StopIteration
        out, ready = self.next
        main = torch.cuda.current_stream(self.device)
        main.wait_event(ready)
        for k in self._KEYS:
            if torch.is_tensor(out.get(k)):
                out[k].record_stream(main)
        self._preload()
        return out


def extract_embeddings_for_split(model, dataloader, target_layers, device, aggregation: str = "mean", depth: int = 3,
                                 disable_layerdrop: Optional[bool] = None) -> Tuple[Dict[str, torch.Tensor], torch.Tensor, list]:
    """Mirror of `_extract_embeddings_in_memory` (embedding_utils.py:26-144) with an asynchronous D2H ring of `depth` batches."""
    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError("avex_b200 extraction runs on CUDA devices only (no CPU fallback)")
    original = None
    if disable_layerdrop is not None and hasattr(model, "disable_layerdrop"):
        original = model.disable_layerdrop
        model.disable_layerdrop = disable_layerdrop
    layer_embeds: Dict[str, List[torch.Tensor]] = {}
    labels: List[torch.Tensor] = []
    ring = _PinnedRing(device, depth)
    pending_names: List[List[str]] = []

    def collect() -> None:
        host = ring.pop()
        for name, h in zip(pending_names.pop(0), host):
            layer_embeds.setdefault(name, []).append(h)

    try:
        with torch.no_grad():
            resolved = model.register_hooks_for_layers(target_layers)
            for batch in _DevicePrefetcher(dataloader, device):
                wav = batch["raw_wav"]
                mask = batch.get("padding_mask")
                if mask is not None:
                    emb = model.extract_embeddings({"raw_wav": wav, "padding_mask": mask}, aggregation=aggregation)
                else:
                    emb = model.extract_embeddings(wav, aggregation=aggregation)
                if isinstance(emb, dict):
                    names, tensors = list(emb.keys()), list(emb.values())
                elif isinstance(emb, (list, tuple)):
                    names = [resolved[i] if i < len(resolved) else f"layer_{i}" for i in range(len(emb))]
                    tensors = list(emb)
                else:
                    names, tensors = [resolved[0] if resolved else "embeddings"], [emb]
                ring.push(tensors)
                pending_names.append(names)
                labels.append(batch["label"].cpu())
                while len(ring) > depth:
                    collect()
            while len(ring):
                collect()
        if not labels:
            raise ValueError("No data processed. Check if dataloader is empty or has invalid batches.")
        final = {name: torch.cat(parts) for name, parts in layer_embeds.items()}
        dims = [tuple(t.shape[1:]) for t in final.values()]
        return final, torch.cat(labels), dims
    finally:
        if original is not None and hasattr(model, "disable_layerdrop"):
            model.disable_layerdrop = original
        model.deregister_all_hooks()


def _write_embedding_metadata(h5f, *, aggregation: str, layer_names: list, embedding_dims: list, multi_layer: bool) -> None:
    """The attrs of embedding_utils.py:147-161, written natively (lists stay lists, dims are `str(tuple)`)."""
    h5f.attrs["embedding_aggregation"] = aggregation
    h5f.attrs["aggregation"] = aggregation
    h5f.attrs["stored_embedding_rank"] = [len(tuple(d)) for d in embedding_dims]
    h5f.attrs["layer_names"] = layer_names
    h5f.attrs["embedding_dims"] = [str(tuple(d)) for d in embedding_dims]
    h5f.attrs["multi_layer"] = multi_layer


class _NpzFile:
    """Stand-in container with h5py.File's write surface (`create_dataset`, `.attrs[...]`) for images without h5py: datasets
    become arrays of an .npz, attrs one JSON document under `__attrs__` (lists stay lists)."""

    def __init__(self, path: str) -> None:
        self.path, self.arrays, self.attrs = path, {}, {}

    def create_dataset(self, name, data=None, **_compression):
        self.arrays[name] = data

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *_):
        if exc_type is None:
            import numpy as np

            np.savez(self.path, __attrs__=np.frombuffer(json.dumps(self.attrs).encode(), dtype=np.uint8), **self.arrays)
        return False


def _open_for_write(save_path: str):
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is not None and hasattr(h5py, "File"):
        return h5py.File(save_path, "w"), save_path
    path = save_path if save_path.endswith(".npz") else save_path + ".npz"
    return _NpzFile(path), path


def save_embeddings_arrays(embeddings, labels: torch.Tensor, save_path, num_labels: Optional[int] = None, compression: str = "gzip",
                           compression_level: int = 4, aggregation: str = "unknown") -> str:
    """`save_embeddings_arrays` of the reference (embedding_utils.py:1433-1580), same arguments and the same file:
    float32 datasets `embeddings_{layer}` (dict input, `multi_layer=True`) or `embeddings` (tensor input), int64 `labels`,
    attrs from `_write_embedding_metadata` plus `num_labels` and `extraction_complete`.  HDF5 when h5py is importable,
    else an .npz with the same names.  Returns the path written."""
    import numpy as np

    save_path = str(save_path)
    os.makedirs(os.path.dirname(os.path.abspath(save_path)), exist_ok=True)
    labels_np = labels.detach().cpu().numpy().astype(np.int64)
    ckw: dict = {}
    if compression and str(compression).lower() not in {"none", "null", "false"}:
        ckw["compression"] = compression
        if str(compression).lower() != "lzf":
            ckw["compression_opts"] = int(compression_level)
    fh, path = _open_for_write(save_path)
    with fh as h5f:
        if isinstance(embeddings, dict):
            names = list(embeddings.keys())
            dims = [tuple(e.shape[1:]) for e in embeddings.values()]
            for n, e in embeddings.items():
                h5f.create_dataset(f"embeddings_{n}", data=e.detach().cpu().numpy().astype(np.float32), **ckw)
            _write_embedding_metadata(h5f, aggregation=aggregation, layer_names=names, embedding_dims=dims, multi_layer=True)
        else:
            arr = embeddings.detach().cpu().numpy().astype(np.float32)
            h5f.create_dataset("embeddings", data=arr, **ckw)
            _write_embedding_metadata(h5f, aggregation=aggregation, layer_names=["embeddings"], embedding_dims=[arr.shape[1:]],
                                      multi_layer=False)
        h5f.create_dataset("labels", data=labels_np, **ckw)
        h5f.attrs["num_labels"] = int(num_labels) if num_labels is not None else int(labels_np.max()) + 1 if labels_np.size else 0
        h5f.attrs["extraction_complete"] = True
    return path


def load_embeddings_arrays(path):
    """`load_embeddings_arrays` of the reference (embedding_utils.py:1583-1678): (embeddings, labels, num_labels) on CPU."""
    import numpy as np

    path = str(path)
    if not os.path.exists(path) and os.path.exists(path + ".npz"):
        path = path + ".npz"
    if not os.path.exists(path):
        raise FileNotFoundError(f"Embeddings file not found: {path}")
    if path.endswith(".npz"):
        z = np.load(path)
        attrs = json.loads(bytes(z["__attrs__"]).decode())
        keys = [k for k in z.files if k != "__attrs__"]
        get = lambda k: z[k]  # noqa: E731
    else:
        import h5py

        h5f = h5py.File(path, "r")
        attrs, keys, get = h5f.attrs, list(h5f.keys()), lambda k: np.asarray(h5f[k])  # noqa: E731
    labels = torch.from_numpy(np.asarray(get("labels")))
    has_prefixed = any(k.startswith("embeddings_") for k in keys)
    if has_prefixed or (attrs.get("multi_layer", False) and "embeddings" not in keys):
        embeds = {n: torch.from_numpy(np.asarray(get(f"embeddings_{n}"), dtype=np.float32)) for n in attrs.get("layer_names", [])}
    elif "embeddings" in keys:
        embeds = torch.from_numpy(np.asarray(get("embeddings"), dtype=np.float32))
    else:
        raise KeyError("No embeddings dataset found in file")
    return embeds, labels, attrs.get("num_labels", None)


# ------------------------------------------------------------------------------------------------------------------
# multi-GPU feeder: one process per GPU, each over its own shard of the clip list
# ------------------------------------------------------------------------------------------------------------------
def extract_embeddings_distributed(model, dataloader, target_layers, device, aggregation: str = "mean", depth: int = 3,
                                   num_samples: Optional[int] = None, group=None, disable_layerdrop: Optional[bool] = None,
                                   loop=None):
    """The extraction loop on every rank over a rank-local `dataloader` whose sampler yields clips `rank::world`
    (`parallel.shard_indices` / `DistributedSampler(shuffle=False)`, avex/data/dataset.py:525-526; the tail wraps around so every
    rank sees ceil(n / world) clips), followed by at most ONE collective per layer:

    * pooled outputs (`aggregation` != "none": [n_local, D] per layer) and labels are all-gathered and put back into clip
      order (`parallel.unshard`), trimmed to `num_samples` -- every rank returns the full arrays, rank 0 typically saves them;
    * frame-level outputs (`aggregation == "none"`) stay rank-local (955 MB per rank at config #5) -- each rank
      returns and saves its own shard (`save_path` + f".rank{r}").

    Returns (embeddings, labels, dims, gathered: bool).  Works on gloo (CPU tensors) and nccl (staged through `device`)."""
    import torch.distributed as dist

    from . import parallel

    loop = loop or extract_embeddings_for_split  # injectable so the collective half is testable on gloo without a GPU
    emb, labels, dims = loop(model, dataloader, target_layers, device, aggregation=aggregation, depth=depth,
                             disable_layerdrop=disable_layerdrop)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1 or aggregation == "none":
        return emb, labels, dims, False
    on_gpu = dist.get_backend(group) == "nccl"

    def gather(t: torch.Tensor) -> torch.Tensor:
        src = t.to(device, non_blocking=True) if on_gpu else t
        out = src.new_empty((world * src.shape[0],) + tuple(src.shape[1:]))
        dist.all_gather_into_tensor(out, src.contiguous(), group=group)
        n = num_samples if num_samples is not None else out.shape[0]
        return parallel.unshard(out, n, world).cpu()

    return {k: gather(v) for k, v in emb.items()}, gather(labels), dims, True
