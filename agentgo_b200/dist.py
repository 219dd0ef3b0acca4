"""Multi-GPU plumbing: one process per GPU, `torch.distributed` over NCCL.

Playouts shard by global playout id (g % world == rank); the only exchange on
the path is ONE allreduce (sum, int64, 3*C counters <= 8.7 KB) per evaluation.  torch is used
for the process group and the collective only — plumbing, not the product.
"""
import os

import numpy as np


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_ids(total, rank, world):
    """Global playout ids owned by `rank` (the kernel's rule: g % world == rank)."""
    return np.arange(rank, total, world, dtype=np.int64)


def local_units(total, rank, world):
    return (total - rank + world - 1) // world if total > rank else 0


class _DevInt64:
    """__cuda_array_interface__ view of the engine's counter buffer."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def attach_allreduce(engine, group=None):
    """Install the NCCL allreduce hook on a sharded engine (agb_set_allreduce).

    The hook runs on the engine's own CUDA stream right after the playout kernel, so no host
    synchronisation separates the kernel epilogue from the collective."""
    import torch
    import torch.distributed as dist

    def hook(ptr, count, stream):
        ext = torch.cuda.ExternalStream(int(stream))
        with torch.cuda.stream(ext):
            t = torch.as_tensor(_DevInt64(ptr, count), device=torch.device("cuda", torch.cuda.current_device()))
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return 0

    engine.set_allreduce(hook)
    return hook


def batch_layout(B, first_position_id, world):
    """Where the counters of batch position b live in the gathered table (the rule of
    Engine::playouts_batch): position b belongs to rank (first_position_id + b) % world and is that
    rank's k-th position in batch order.  -> (stride = largest local count, [(rank, k)] * B)"""
    nxt = [0] * world
    where = []
    for b in range(B):
        r = (first_position_id + b) % world
        where.append((r, nxt[r]))
        nxt[r] += 1
    return (max(nxt) if B else 0), where


def attach_allgather(engine, group=None):
    """Install the NCCL allgather hook of the batch path (agb_set_allgather): positions are sharded
    by id, every shard's [3 x stride] int64 table is gathered in rank order with ONE collective on
    the engine's own stream."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)

    def hook(send, recv, count, stream):
        ext = torch.cuda.ExternalStream(int(stream))
        dev = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.stream(ext):
            src = torch.as_tensor(_DevInt64(send, count), device=dev)
            dst = torch.as_tensor(_DevInt64(recv, count * world), device=dev)
            dist.all_gather_into_tensor(dst, src, group=group)
        return 0

    engine.set_allgather(hook)
    return hook


def gather_host_tables(table, group=None):
    """The same gather for int64 numpy tables on the host (gloo; the world_size-2 CPU tests)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(table, dtype=np.int64).copy())
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return [o.numpy() for o in out]


def allreduce_host_counts(arrays, group=None):
    """Sum int64 numpy arrays over the process group (works with gloo on CPU; used by the
    world_size-2 tests of the sharding rule)."""
    import torch
    import torch.distributed as dist
    out = []
    for a in arrays:
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        out.append(t.numpy())
    return out
