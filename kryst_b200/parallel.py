"""Host-side plumbing of the row-block partition (one process per GPU).

The device library builds the partition maps itself (kb_csr_create_dist); this module holds what the host
program needs around it: bootstrap of the communicator through an existing torch.distributed process group
(any backend — gloo on CPU test boxes, nccl on GPU boxes), shard generation, and a numpy restatement of the
ghost / send-list construction used to cross-check the device maps.
The reference's only partition formula is the uniform chunk of src/preconditioner/asm.rs:46-57.
"""
import os

import numpy as np

from .api import Context, partition_range
from . import stencils


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def broadcast_unique_id(make_id, rank, src=0):
    """Rank `src` creates the 128-byte communicator id, every rank receives it (torch.distributed plumbing)."""
    import torch.distributed as dist
    box = [make_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def init_comm(ctx, rank=None, world=None):
    """Collective: give `ctx` a communicator spanning the process group (Comm::rank/size/all_reduce)."""
    r, w, _ = dist_env()
    rank = r if rank is None else rank
    world = w if world is None else world
    if world > 1:
        uid = broadcast_unique_id(Context.comm_unique_id, rank)
        ctx.comm_init(rank, world, uid)
    return ctx


def shard_stencil(kind, N, world, rank, pe=(0.4, 0.2, 0.1)):
    """-> (n_global, lo, hi, row_ptr, col_idx(global), vals) of this rank's row block."""
    n = stencils.dim(kind, N)
    lo, hi = partition_range(n, world, rank)
    _, rp, ci, v = stencils.stencil(kind, N, lo, hi, pe)
    return n, lo, hi, rp, ci, v


def host_ghost_plan(n_global, world, rank, row_ptr, col_idx):
    """numpy restatement of the shard maps:
         ghosts      sorted unique off-range global columns
         local_col   owned -> c - lo ; ghost -> n_loc + index in `ghosts`   (stored order unchanged)
         recv_from   {owner: (offset into ghosts, count)}  (contiguous because ghosts are sorted)
    """
    lo, hi = partition_range(n_global, world, rank)
    ci = np.asarray(col_idx, dtype=np.int64)
    off = (ci < lo) | (ci >= hi)
    ghosts = np.unique(ci[off])
    local = np.where(off, (hi - lo) + np.searchsorted(ghosts, ci), ci - lo)
    chunk = (n_global + world - 1) // world
    owners = ghosts // chunk
    recv_from = {}
    for q in np.unique(owners):
        idx = np.nonzero(owners == q)[0]
        recv_from[int(q)] = (int(idx[0]), int(idx.size))
    return ghosts.astype(np.uint64), local.astype(np.int64), recv_from


def send_lists(all_ghosts, n_global, world, rank):
    """Given every rank's ghost list, the owned global rows this rank must send to each peer (sorted)."""
    lo, hi = partition_range(n_global, world, rank)
    out = {}
    for q, g in enumerate(all_ghosts):
        if q == rank:
            continue
        g = np.asarray(g, dtype=np.int64)
        mine = g[(g >= lo) & (g < hi)]
        if mine.size:
            out[q] = mine
    return out
