"""Data parallelism for the training step: one process per GPU, utterances sharded across ranks, parameters and
optimiser state replicated, ONE all-reduce(SUM) of the flat gradient arena per step over NCCL/NVLink.

The reference has no distributed code; this is the one parallelism strategy the build adds.  What keeps the sharded
step identical to the reference's single-process step on the global batch:
  * the loss normaliser is global: `temporal_softmax_loss` divides by the *global* mask sum
    (`custom/objectives.py:32,37`) — every rank all-reduces its mask count first (a 4-byte device-side all-reduce, no
    host round trip) and the loss kernel normalises by it, so SUM over ranks reproduces the global gradient;
    `mean(categorical_crossentropy)` divides by the global utterance count;
  * the ±5 gate-gradient clip of the LSTMs acts on correctly normalised gradients for the same reason;
  * BatchNorm (adenet_v1) all-reduces its per-feature sum / sum-of-squares (sync-BN), so batch statistics are global.
Inference shards utterances with no collective at all.
"""
import os

import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous utterance slice [lo, hi) of rank `rank`: sizes differ by at most one, earlier ranks get the extras."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(arrays, rank, world):
    """Slice every array of a batch (leading axis = utterances) for this rank."""
    n = len(arrays[0])
    lo, hi = shard_bounds(n, rank, world)
    return [a[lo:hi] for a in arrays]


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def attach(engine, group=None):
    """Make an Engine data-parallel over the default (or given) process group."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        engine.world = (dist.get_rank(group), dist.get_world_size(group), group if group is not None else dist.group.WORLD)
        # replicas must start from identical parameters: broadcast rank 0's arena
        dist.broadcast(engine.arena.flat, src=0, group=group)
        dist.broadcast(engine.arena.aux, src=0, group=group)
        engine.arena.split_dirty = True
    return engine


def global_normaliser(local_count, group=None):
    """Host-side helper (gloo or nccl): the global loss normaliser from per-rank counts."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(local_count)], dtype=torch.float64)
    if dist.is_initialized():
        if dist.get_backend(group) == 'nccl':
            t = t.cuda()
        dist.all_reduce(t, group=group)
    return float(t.item())


def allreduce_host_grads(grads, group=None):
    """Sum a list of host gradient arrays across ranks (used by the CPU/gloo tests of the sharding algebra)."""
    import torch
    import torch.distributed as dist
    flat = torch.from_numpy(np.concatenate([np.asarray(g, np.float64).ravel() for g in grads]))
    if dist.is_initialized():
        dist.all_reduce(flat, group=group)
    out, o = [], 0
    for g in grads:
        n = int(np.prod(np.shape(g))) if np.shape(g) else 1
        out.append(flat[o:o + n].numpy().reshape(np.shape(g)))
        o += n
    return out
