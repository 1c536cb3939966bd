"""Clip-level data parallelism for embedding extraction: one process per GPU, weights replicated, the clip list
sharded `rank::world` (the reference's DistributedSampler pattern, avex/data/dataset.py:525-526), zero communication
during the forward, and ONE collective per batch: an all-gather of the pooled [B_local, D] embeddings (NCCL over
NVLink on GPUs; gloo on CPU for tests).  The reference's extraction path is single-process
(avex/run_evaluate.py:1053); this is the multi-GPU extension BASELINE.json's north_star asks for.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple[int, int, int]:
    """torchrun-style bring-up (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*), like avex/training/distributed.py:163-183.
    `ModelSpec.device` only admits "cuda" (avex/configs.py:355-372), so per-rank placement is torch.cuda.set_device."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = os.environ.get("PYTORCH_DISTRIBUTED_BACKEND", "nccl" if torch.cuda.is_available() else "gloo")
    if backend == "nccl":
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return rank, local, world


def shard_indices(n_clips: int, rank: int, world: int) -> list[int]:
    """Clips `rank::world`; every rank gets ceil(n/world) entries (the tail wraps around, DistributedSampler-style) so
    the all-gather is rectangular; `unshard` drops the duplicates."""
    per = (n_clips + world - 1) // world
    return [(rank + i * world) % n_clips for i in range(per)] if n_clips > 0 else []


def gather_embeddings(local: torch.Tensor, world: Optional[int] = None) -> torch.Tensor:
    """[B_local, D] on every rank -> [world * B_local, D] on every rank (rank-major), one collective."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    out = local.new_empty((world * local.shape[0],) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def unshard(gathered: torch.Tensor, n_clips: int, world: int) -> torch.Tensor:
    """Invert `shard_indices`: rank-major [world * per, D] -> original clip order [n_clips, D]."""
    per = gathered.shape[0] // world
    g = gathered.view(world, per, *gathered.shape[1:]).transpose(0, 1).reshape(world * per, *gathered.shape[1:])
    return g[:n_clips]


def extract_sharded(embed_fn: Callable[[torch.Tensor], torch.Tensor], clips: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Run `embed_fn` ([b, T] -> [b, D]) on this rank's shard of `clips` and return all embeddings in clip order."""
    idx = shard_indices(clips.shape[0], rank, world)
    local = embed_fn(clips[idx])
    return unshard(gather_embeddings(local, world), clips.shape[0], world)
