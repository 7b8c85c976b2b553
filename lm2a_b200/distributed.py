"""Clip sharding across the GPUs of one box: one process per GPU, clips are independent, so
there is NO collective inside the denoising loop; the finished mels are exchanged with at
most one all-gather at the end. Host logic only — the per-batch sampler is
passed in, so the same driver runs under `gloo` on CPU in the tests.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style init (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*). Returns (rank, world,
    local_rank). Single-process when WORLD_SIZE is unset or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_indices(n_clips, rank, world):
    """Round-robin shard: rank r owns clips r, r+W, r+2W, ... (<= 1 clip imbalance)."""
    return list(range(rank, n_clips, world))


def padded_shard_len(n_clips, world):
    return (n_clips + world - 1) // world


def batches(indices, batch):
    for i in range(0, len(indices), batch):
        yield indices[i:i + batch]


def sample_sharded(n_clips, batch, sample_batch_fn, clip_shape, device, rank=0, world=1,
                   gather=True):
    """Runs `sample_batch_fn(list_of_clip_indices) -> tensor [len, *clip_shape]` over this
    rank's shard in batches of `batch`, then all-gathers. Returns a tensor
    [n_clips, *clip_shape] ordered by clip index (on every rank) when gather=True, else this
    rank's [shard_len, *clip_shape]."""
    mine = shard_indices(n_clips, rank, world)
    per = padded_shard_len(n_clips, world)
    local = torch.zeros((per,) + tuple(clip_shape), dtype=torch.float32, device=device)
    pos = 0
    for idx in batches(mine, batch):
        out = sample_batch_fn(idx)
        local[pos:pos + len(idx)].copy_(out)
        pos += len(idx)
    if not gather:
        return local[:len(mine)]
    if world == 1:
        return local[:n_clips]
    full = torch.empty((world * per,) + tuple(clip_shape), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(full, local)
    # rank r's j-th entry is clip r + j*W  ->  clip c lives at [c % W, c // W]
    full = full.view(world, per, *clip_shape)
    order = torch.arange(n_clips, device=device)
    return full[order % world, order // world]
