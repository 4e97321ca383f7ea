"""Lean in-place CPU evaluation of -2 logL for sizes the op-for-op oracle cannot hold in RAM.

TEST / BASELINE INFRASTRUCTURE ONLY (same rule as oracle/svc_oracle.py): used by the CPU arm of
bench.py and by tests; nothing under varycoef_b200/ may import it.

Same maths as n2LL (R-pkg/R/objective_functions.R:4-62) with the cheapest op sequence a CPU can use:
  * Sigma_y (R-pkg/R/utils.R:175-194) is written once, lower triangle only, column block by column
    block, from the coordinates -- no n x n distance matrix, no outer.W list (the reference keeps
    (q + 2) such matrices: 100 GB at n = 50 000);
  * dpotrf in place (R-pkg/R/objective_functions.R:26);
  * dtrtrs for L^-1 [X y] instead of the three dgesv calls (R-pkg/R/utils.R:94,100,
    objective_functions.R:49);  sum of log diag(L) instead of the LU determinant (:59).
The blocks of the build run on a thread pool (numpy ufuncs release the GIL); LAPACK uses the BLAS
threads.  Peak memory: 8 n^2 bytes + O(block * n).
"""
from __future__ import annotations

import math
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
from scipy.linalg import lapack

from . import svc_oracle as o


def build_lower_inplace(S, theta, locs, W, cov_name="exp", block=512, threads=1):
    """Fill the lower triangle (incl. diagonal) of the Fortran-ordered n x n array S with Sigma_y."""
    n, q = W.shape
    f = o.cov_func(cov_name)
    theta = np.asarray(theta, dtype=np.float64)

    def job(j0):
        j1 = min(j0 + block, n)
        D = o._pair_dist(locs[j0:], locs[j0:j1])          # rows j0.. against columns j0..j1
        acc = np.zeros_like(D)
        for k in range(q):
            C = f(D, theta[2 * k:2 * k + 2])
            C *= W[j0:, k][:, None]
            C *= W[j0:j1, k][None, :]
            acc += C
        idx = np.arange(j1 - j0)
        acc[idx, idx] += theta[2 * q]
        S[j0:, j0:j1] = acc
        return j1 - j0

    starts = list(range(0, n, block))
    if threads <= 1:
        for j0 in starts:
            job(j0)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(job, starts))


def n2ll_lean_inplace(theta, locs, X, W, y, cov_name="exp", threads=1, block=512, time_dgesv=False):
    """Returns {"n2ll", "mu", "logdet", "quad", "t_build", "t_chol", "t_solve"} (profile / GLS mode).

    time_dgesv=True additionally times ONE of the three solve(t(R), .) calls of the reference exactly as R
    runs it (R-pkg/R/utils.R:94: La_solve -> copy of t(R), dgesv = LU with partial pivoting of the triangular
    factor) on the factor just computed, at this n: {"t_dgesv"}.  Needs a second n x n array."""
    locs = np.asarray(locs, dtype=np.float64)
    if locs.ndim == 1:
        locs = locs[:, None]
    X = np.asarray(X, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    n, p = X.shape
    q = W.shape[1]
    S = np.empty((n, n), order="F")
    t0 = time.perf_counter()
    build_lower_inplace(S, theta, locs, W, cov_name, block, threads)
    t1 = time.perf_counter()
    c, info = lapack.dpotrf(S, lower=1, clean=0, overwrite_a=1)
    if info > 0:
        raise np.linalg.LinAlgError("the leading minor of order %d is not positive" % info)
    t2 = time.perf_counter()
    B = np.asfortranarray(np.column_stack([X, y]))
    Z, info = lapack.dtrtrs(c, B, lower=1, trans=0)
    G = Z[:, :p].T @ Z[:, :p]
    mu = np.linalg.solve(G, Z[:, :p].T @ Z[:, p])
    r = Z[:, p] - Z[:, :p] @ mu
    quad = float(r @ r)
    logdet = 2.0 * float(np.sum(np.log(np.diagonal(c))))
    t3 = time.perf_counter()
    val = n * math.log(2.0 * math.pi) + logdet + quad
    out = {"n2ll": val, "mu": mu, "logdet": logdet, "quad": quad,
           "t_build": t1 - t0, "t_chol": t2 - t1, "t_solve": t3 - t2}
    if time_dgesv:
        # t(R) = L as a proper triangular matrix: the strict upper triangle was never written, zero it first
        for j0 in range(0, n, 2048):
            j1 = min(j0 + 2048, n)
            blk = c[:j1, j0:j1]
            blk[np.triu_indices(j1, 1 - j0, j1 - j0)] = 0.0
        t4 = time.perf_counter()
        lu, piv, x, info = lapack.dgesv(c, np.asfortranarray(X))      # overwrite_a = 0: copies t(R), as La_solve does
        out["t_dgesv"] = time.perf_counter() - t4
        out["dgesv_check"] = float(np.max(np.abs(x - Z[:, :p])) / max(np.max(np.abs(Z[:, :p])), 1e-300))
    return out
