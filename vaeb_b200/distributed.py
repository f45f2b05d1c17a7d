"""Single-node multi-GPU plumbing: one process per GPU, `torch.distributed` for rendezvous,
NCCL (attached inside the C library) for the gradient all-reduce, plain sharding for the
importance-sampling estimator.  Nothing here touches the data path."""
from __future__ import annotations

import os

import numpy as np


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object from `src` through the default process group."""
    import torch.distributed as dist
    obj = [payload]
    dist.broadcast_object_list(obj, src=src)
    return obj[0]


def attach_data_parallel(model):
    """Give `model` (one per rank, built on this rank's GPU with batch_size = per-rank share)
    an NCCL communicator: afterwards every update all-reduces gradients and the bound."""
    import torch.distributed as dist
    from .model import comm_unique_id
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 0)
    model.attach_comm(uid, rank, world)
    # ranks of one box map each other's buffers (CUDA IPC over NVLink): the tail kernel of a large-batch tensor-core
    # update then IS the gradient all-reduce (vaeb_comm_p2p_attach); VAEB_DP_P2P=0 keeps ncclAllReduce
    if 1 < world <= 8 and os.environ.get("VAEB_DP_P2P", "1") != "0" and getattr(model, "precision", "fp32") != "fp32":
        mine = model.p2p_export()
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        model.p2p_attach(handles)
        dist.barrier()
    return rank, world


def close_data_parallel(model):
    """Collective counterpart of model.close(): every rank stops using its peers' buffers (barrier) before any rank
    unmaps and frees its own (the exported allocations must outlive every importer's mapping)."""
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
        _lib_detach(model)
        dist.barrier()
    model.close()


def _lib_detach(model):
    from . import _lib
    _lib.check(model._lib.vaeb_comm_detach(model._h))


def shard_rows(n, rank, world):
    """Contiguous block of rows owned by `rank` (N/G points per GPU)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def sharded_log_px(model, x, L, rank, world, gather=True):
    """IS estimator over the shard of x owned by this rank.  Philox counters are keyed by the
    GLOBAL row, so the concatenated result is identical for any world size."""
    lo, hi = shard_rows(len(x), rank, world)
    local = model.log_px(x[lo:hi], L=L, row_offset=lo) if hi > lo else np.empty(0, np.float32)
    if not gather or world == 1:
        return local
    import torch
    import torch.distributed as dist
    outs = [None] * world
    dist.all_gather_object(outs, local)
    return np.concatenate(outs)
