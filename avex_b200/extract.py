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

`save_embeddings_arrays` writes the reference's on-disk layout (embedding_utils.py:147-161, :1433-1580): one float32 dataset
`embeddings_{layer}` per layer plus `labels` and the attrs `aggregation`, `layer_names`, `embedding_dims`, `multi_layer`,
`extraction_complete` -- as HDF5 when h5py is importable, else as an `.npz` with the same names (h5py is not in this image).
"""
from __future__ import annotations

import json
import logging
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


def save_embeddings_arrays(embeddings: Dict[str, torch.Tensor], labels: torch.Tensor, save_path: str, aggregation: str = "mean") -> str:
    """Reference layout (embedding_utils.py:147-161): datasets `embeddings_{layer}` (float32, [num_samples, *dims]) + `labels`,
    attrs aggregation / embedding_aggregation / layer_names / embedding_dims / multi_layer / extraction_complete."""
    names = list(embeddings.keys())
    attrs = {"aggregation": aggregation, "embedding_aggregation": aggregation, "layer_names": names,
             "embedding_dims": [list(embeddings[n].shape[1:]) for n in names], "multi_layer": len(names) > 1,
             "stored_embedding_rank": int(embeddings[names[0]].dim()) if names else 0, "extraction_complete": True}  # fmt: skip
    try:
        import h5py  # noqa: F401
    except ImportError:
        import numpy as np

        path = save_path if save_path.endswith(".npz") else save_path + ".npz"
        arrays = {f"embeddings_{n}": embeddings[n].float().numpy() for n in names}
        arrays["labels"] = labels.numpy()
        arrays["__attrs__"] = np.frombuffer(json.dumps(attrs).encode(), dtype=np.uint8)
        np.savez(path, **arrays)
        return path
    import h5py

    with h5py.File(save_path, "w") as f:
        for n in names:
            f.create_dataset(f"embeddings_{n}", data=embeddings[n].float().numpy())
        f.create_dataset("labels", data=labels.numpy())
        for k, v in attrs.items():
            f.attrs[k] = json.dumps(v) if isinstance(v, (list, dict)) else v
    return save_path
