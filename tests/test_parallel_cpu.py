"""CPU: the N>1 path (clip sharding + one all-gather of pooled embeddings) with world_size 2 on gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avex_b200 import parallel


def test_shard_indices_cover_all_clips():
    for n in (1, 2, 7, 8, 256, 257):
        for world in (1, 2, 4, 8):
            shards = [parallel.shard_indices(n, r, world) for r in range(world)]
            assert len({len(s) for s in shards}) == 1  # rectangular
            assert set(i for s in shards for i in s) == set(range(n))
            flat = torch.tensor(shards, dtype=torch.float32).reshape(world * len(shards[0]), 1)
            back = parallel.unshard(flat, n, world)
            assert back[:, 0].tolist() == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = parallel.init_from_env("gloo")
    clips = torch.arange(n_clips * 5, dtype=torch.float32).reshape(n_clips, 5)

    def embed(x):  # stand-in for the CUDA forward: deterministic per clip, independent of batch neighbours
        return torch.stack([x.sum(dim=1), x[:, 0] * 3.0], dim=1)

    out = parallel.extract_sharded(embed, clips, r, w)
    if r == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_gather_matches_single_process():
    ctx = mp.get_context("spawn")
    for n_clips in (6, 7):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
        for p in procs:
            p.start()
        got = q.get(timeout=120)
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
        clips = torch.arange(n_clips * 5, dtype=torch.float32).reshape(n_clips, 5)
        want = torch.stack([clips.sum(dim=1), clips[:, 0] * 3.0], dim=1)
        assert torch.equal(got, want)
