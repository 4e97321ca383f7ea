"""world_size-2 gloo tests (CPU) of the host-side logic of the multi-GPU paths:

* the 2D block-cyclic schedule of vc_dist.cu -- ownership maps and per-rank trailing-update tile
  lists exported by the C library (vc_bc_owner / vc_bc_trailing_tiles) -- driven by a numpy
  emulation of the distributed right-looking Cholesky with real broadcasts between two processes;
* the theta-parallel sharding bench.py uses for N > 1 (rank r evaluates theta_{s N + r}, max-time
  reduction), with the CPU oracle standing in for the GPU objective.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _bc_cholesky_worker(rank, world, port, pr, pc, nt, nb, T, q_out):
    import ctypes as C  # noqa: N812
    from varycoef_b200 import _lib
    L = _lib.lib()
    _init(rank, world, port)
    rng = np.random.default_rng(0)
    n = nt * T
    A0 = rng.standard_normal((n, n))
    S = A0 @ A0.T + n * np.eye(n)                      # same SPD matrix on every rank
    my_pr, my_pc = rank // pc, rank % pc
    owner = lambda I, J: L.vc_bc_owner(I // nb, J // nb, pr, pc)  # noqa: E731
    # local storage: only the tiles this rank owns (lower part)
    loc = {(I, J): S[I * T:(I + 1) * T, J * T:(J + 1) * T].copy()
           for J in range(nt) for I in range(J, nt) if owner(I, J) == rank}
    nblocks = (nt + nb - 1) // nb
    for b in range(nblocks):
        ob, oe = b * nb, min(b * nb + nb, nt)
        K = (oe - ob) * T
        pcp = b % pc
        diag_owner = L.vc_bc_owner(b, b, pr, pc)
        assert diag_owner == (b % pr) * pc + pcp
        # 1. diagonal block: factor on its owner, broadcast
        dd = torch.zeros((K, K), dtype=torch.float64)
        if rank == diag_owner:
            blk = np.zeros((K, K))
            for J in range(ob, oe):
                for I in range(J, oe):
                    blk[(I - ob) * T:(I - ob + 1) * T, (J - ob) * T:(J - ob + 1) * T] = loc[(I, J)]
            Ld = np.linalg.cholesky(np.tril(blk) + np.tril(blk, -1).T)
            for J in range(ob, oe):
                for I in range(J, oe):
                    loc[(I, J)] = Ld[(I - ob) * T:(I - ob + 1) * T, (J - ob) * T:(J - ob + 1) * T].copy()
            dd = torch.from_numpy(Ld.copy())
        dist.broadcast(dd, src=diag_owner)
        Ld = dd.numpy()
        # 2. panel solve on the panel's process column
        if my_pc == pcp:
            for I in range(oe, nt):
                if (I // nb) % pr == my_pr:
                    row = np.hstack([loc[(I, J)] for J in range(ob, oe)])
                    row = np.linalg.solve(Ld, row.T).T           # X Ld^T = A
                    for k, J in enumerate(range(ob, oe)):
                        loc[(I, J)] = row[:, k * T:(k + 1) * T].copy()
        if oe >= nt:
            continue
        # 3. the gathered panel buffer is OWNER-major (vc_bc_panel_layout): every process row's part is one contiguous
        #    range and travels in ONE broadcast from its owner in the panel's process column (round 2; round 1 sent
        #    one broadcast per distribution block)
        ntl = nt - oe
        off = (C.c_int32 * 8)()
        cnt = (C.c_int32 * 8)()
        pos = (C.c_int32 * ntl)()
        assert L.vc_bc_panel_layout(nt, 0, nb, pr, oe, off, cnt, pos) == ntl
        buf = torch.zeros((ntl * T, K), dtype=torch.float64)
        for I in range(oe, nt):                                   # pack: this rank's rows of the panel at their positions
            if my_pc == pcp and (I // nb) % pr == my_pr:
                q = pos[I - oe]
                buf[q * T:(q + 1) * T] = torch.from_numpy(np.hstack([loc[(I, J)] for J in range(ob, oe)]))
        nbcast = 0
        for r in range(pr):
            if cnt[r] > 0:
                chunk = buf[off[r] * T:(off[r] + cnt[r]) * T]     # a view: received in place, like the NCCL broadcast
                dist.broadcast(chunk, src=r * pc + pcp)
                nbcast += 1
        assert nbcast <= pr
        panel = {I: buf.numpy()[pos[I - oe] * T:(pos[I - oe] + 1) * T] for I in range(oe, nt)}
        # 4. trailing update of the tiles the C library assigns to this rank
        cnt = L.vc_bc_trailing_tiles(nt, 0, nb, pr, pc, rank, oe, None, 0)
        out = (C.c_int32 * max(cnt, 1))()
        assert L.vc_bc_trailing_tiles(nt, 0, nb, pr, pc, rank, oe, out, cnt) == cnt
        mine = set()
        for t in list(out)[:cnt]:
            I, J = t & 0xffff, t >> 16
            assert owner(I, J) == rank and oe <= J <= I < nt
            mine.add((I, J))
            loc[(I, J)] -= panel[I] @ panel[J].T
        assert mine == {(I, J) for (I, J) in loc if J >= oe}
    # gather and compare on rank 0
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: v for k, v in loc.items()})
    if rank == 0:
        Lfull = np.zeros((n, n))
        seen = set()
        for d in gathered:
            for (I, J), v in d.items():
                assert (I, J) not in seen
                seen.add((I, J))
                Lfull[I * T:(I + 1) * T, J * T:(J + 1) * T] = v
        assert len(seen) == nt * (nt + 1) // 2
        ref = np.linalg.cholesky(S)
        q_out.put(float(np.max(np.abs(np.tril(Lfull) - ref)) / np.max(np.abs(ref))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("pr,pc,nt,nb", [(1, 2, 7, 2), (2, 1, 9, 4), (1, 2, 10, 4)])
def test_block_cyclic_cholesky_schedule_two_ranks(pr, pc, nt, nb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + nt * 7 + nb) % 300
    procs = [ctx.Process(target=_bc_cholesky_worker, args=(r, 2, port, pr, pc, nt, nb, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-12


def test_trailing_tile_lists_partition_the_trapezoid():
    import ctypes as C
    from varycoef_b200 import _lib
    L = _lib.lib()
    for (pr, pc) in ((1, 1), (1, 2), (2, 2), (2, 4)):
        for nt, nbt, nb, oe in ((13, 1, 4, 4), (37, 1, 4, 12), (9, 0, 2, 2), (20, 1, 4, 20)):
            seen = []
            for rank in range(pr * pc):
                cnt = L.vc_bc_trailing_tiles(nt, nbt, nb, pr, pc, rank, oe, None, 0)
                out = (C.c_int32 * max(cnt, 1))()
                L.vc_bc_trailing_tiles(nt, nbt, nb, pr, pc, rank, oe, out, cnt)
                for t in list(out)[:cnt]:
                    I, J = t & 0xffff, t >> 16
                    orow = (I // nb) % pr if I < nt else 0
                    assert orow * pc + (J // nb) % pc == rank
                    seen.append((I, J))
            want = {(I, J) for J in range(oe, nt) for I in range(J, nt + nbt)}
            assert len(seen) == len(set(seen)) and set(seen) == want
    assert L.vc_bc_trailing_tiles(10, 1, 4, 2, 2, 7, 4, None, 0) == -1


def _theta_parallel_worker(rank, world, port, q_out):
    from oracle import svc_oracle as o
    sys.path.insert(0, ROOT)
    import bench
    _init(rank, world, port)
    n, p = 300, 2
    locs, X, y, Lside = bench.make_inputs(n, p, 1)
    thetas = bench.theta_stencil(n, p, p, Lside, y)
    D = o.own_dist(locs)
    K = 3
    vals = [o.n2ll(thetas[(s * world + rank) % len(thetas)], D, X, X, y, profile=True) for s in range(K)]
    t = torch.tensor([0.01 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    allv = [None] * world
    dist.all_gather_object(allv, vals)
    if rank == 0:
        q_out.put((float(t[0]), allv, [o.n2ll(th, D, X, X, y, profile=True) for th in thetas[:K * world]]))
    dist.barrier()
    dist.destroy_process_group()


def test_theta_parallel_sharding_two_ranks():
    """bench.py --gpus N: rank r takes theta_{sN+r}; together the ranks cover the stencil once, in
    order, and the reported time is the max over ranks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_theta_parallel_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    tmax, allv, ref = q.get(timeout=5)
    assert tmax == pytest.approx(0.02)
    inter = [allv[r][s] for s in range(3) for r in range(2)]
    assert np.allclose(inter, ref, rtol=1e-14)
