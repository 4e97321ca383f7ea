"""Extended-precision (x87 80-bit long double, eps = 1.1e-19) evaluation of -2 logL for small n.

TEST INFRASTRUCTURE ONLY.  At the optimiser's default bounds (nugget 1e-6, range ten times the median
distance, R-pkg/R/utils.R:336-396) Sigma_y is so ill-conditioned that two correct fp64 implementations --
the reference's LAPACK path included -- differ by far more than 1e-10: the error of -2logL is
~ cond(Sigma) * eps.  This module provides the value both are measured against: the same formulas as
n2LL (R-pkg/R/objective_functions.R:4-62) and Sigma_y (R-pkg/R/utils.R:175-194) in long double, about
three more digits than fp64.  No LAPACK: a left-looking column Cholesky with numpy vector operations,
O(n^3 / 3) long-double flops -- seconds at n ~ 1000.
"""
from __future__ import annotations

import numpy as np

LD = np.longdouble


def _shape(name, t):
    e = np.exp(-t)
    if name == "exp":
        return e
    if name == "mat32":
        return (1 + t) * e
    if name == "mat52":
        return (1 + t + t * t / 3) * e
    raise ValueError("long-double oracle covers exp / mat32 / mat52")


def n2ll_longdouble(theta, locs, X, W, y, cov_name="exp"):
    """profile / GLS mode; returns {"n2ll", "logdet", "quad", "mu"} as Python floats (rounded from long double)."""
    locs = np.asarray(locs, dtype=LD)
    if locs.ndim == 1:
        locs = locs[:, None]
    X = np.asarray(X, dtype=LD)
    W = np.asarray(W, dtype=LD)
    y = np.asarray(y, dtype=LD).ravel()
    theta = np.asarray(theta, dtype=LD)
    n, p = X.shape
    q = W.shape[1]
    acc = np.zeros((n, n), dtype=LD)
    for c in range(locs.shape[1]):
        d = locs[:, c][:, None] - locs[:, c][None, :]
        acc += d * d
    D = np.sqrt(acc)
    S = np.zeros((n, n), dtype=LD)
    for k in range(q):
        S += theta[2 * k + 1] * _shape(cov_name, D / theta[2 * k]) * np.outer(W[:, k], W[:, k])
    S[np.diag_indices(n)] += theta[2 * q]
    # left-looking Cholesky, L stored row-major so that L[j:, :j] @ L[j, :j] is a contiguous mat-vec
    L = np.zeros((n, n), dtype=LD)
    for j in range(n):
        col = S[j:, j] - L[j:, :j] @ L[j, :j]
        if not col[0] > 0:
            raise np.linalg.LinAlgError("the leading minor of order %d is not positive" % (j + 1))
        col = col / np.sqrt(col[0])
        L[j:, j] = col
    B = np.column_stack([X, y])
    Z = np.zeros_like(B)
    for i in range(n):
        Z[i] = (B[i] - L[i, :i] @ Z[:i]) / L[i, i]
    G = Z[:, :p].T @ Z[:, :p]
    g = Z[:, :p].T @ Z[:, p]
    # p x p normal equations in long double (Gaussian elimination; G is SPD and tiny)
    A = np.column_stack([G, g]).astype(LD)
    for i in range(p):
        A[i] = A[i] / A[i, i]
        for r in range(p):
            if r != i:
                A[r] = A[r] - A[r, i] * A[i]
    mu = A[:, p]
    r = Z[:, p] - Z[:, :p] @ mu
    quad = r @ r
    logdet = 2 * np.sum(np.log(np.diagonal(L)))
    val = n * np.log(2 * LD(np.pi)) + logdet + quad
    return {"n2ll": float(val), "logdet": float(logdet), "quad": float(quad), "mu": np.asarray(mu, dtype=np.float64),
            "_n2ll_ld": val}
