"""One -2logL evaluation sharded over several GPUs (one process per GPU, torchrun / torch.distributed).

The reference has no counterpart: it is a single R process (optimParallel PSOCK workers evaluate
independent thetas, R-pkg/R/SVC_mle.R:164-200).  This exists for matrices that do not fit one GPU
(n = 150 000: 180 GB of fp64 Sigma_y).  torch.distributed is only the plumbing that ships the
128-byte NCCL id to every rank; the data path is the library's own NCCL communicator.
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import lib, check
from .objective import SVCObjective


def process_grid(world: int):
    """pr x pc with pr >= pc, as square as possible: 1->1x1, 2->2x1, 4->2x2, 8->4x2.  More process
    ROWS than columns: the rows of a panel (its TRSM / inner-update work, which overlaps the
    trailing update on a handful of SMs only) are split over pr ranks."""
    pc = 1
    for c in range(1, int(world ** 0.5) + 1):
        if world % c == 0:
            pc = c
    return world // pc, pc


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib().vc_nccl_unique_id(buf), None)
    return buf.raw


def distributed_objective(y, X, locs, W=None, grid=None, device=None, **kw) -> SVCObjective:
    """Collective constructor: every rank of the default torch.distributed group calls it with the
    same (replicated) inputs; returns an SVCObjective whose n2LL is collective."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed must be initialised (torchrun) before distributed_objective")
    rank, world = dist.get_rank(), dist.get_world_size()
    import os
    if grid is None and os.environ.get("VC_GRID"):
        grid = tuple(int(v) for v in os.environ["VC_GRID"].split("x"))
    pr, pc = grid if grid is not None else process_grid(world)
    if pr * pc != world:
        raise ValueError("grid %dx%d does not match world size %d" % (pr, pc, world))
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    if device is None:
        device = torch.cuda.current_device()
    return SVCObjective(y, X, locs, W, device=device, _dist_grid=(rank, world, pr, pc, box[0]), **kw)
