"""High-precision known answers for the objective ("oracle precision plan").

Evaluates -2 log L = n log 2pi + log|Sigma| + r' Sigma^-1 r with the GLS mean in 40-digit arithmetic
(mpmath) from the SAME fp64 inputs the tests use, so both the CPU oracle and the GPU path can be shown to sit
within 1e-12 (relative) of the true value.  Run once here; the output is committed:

    python tests/golden/make_highprec.py        # writes tests/golden/highprec_golden.json

Formulas: R-pkg/R/objective_functions.R:4-62 (n2LL), R-pkg/R/utils.R:93-101 (GLS_chol), :171-194 (Sigma_y),
covariance shapes as spam::cov.exp / cov.mat (nu = 3/2).
"""
import json
import os
import sys

import numpy as np
from mpmath import mp, mpf

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))            # tests/ (conftest.synth)
mp.dps = 40


def case_inputs(name):
    if name == "vignette_n100_exp":
        d = np.load(os.path.join(HERE, "svcdata.npz"))
        ids = d["id_train"] - 1
        y, X, locs = d["y"][ids], d["X"][ids], d["locs"][ids].reshape(-1, 1)
        theta = np.array([1.68358, 0.25786, 1.31196, 0.51061, 0.26822])
        return y, X, X, locs, theta, "exp"
    if name == "synthetic_n260_mat32":
        from conftest import synth
        locs, X, W, y = synth(260, 3, 777, 2, q=2)
        theta = np.array([0.9, 1.3, 1.7, 0.6, 0.35])
        return y, X, W, locs, theta, "mat32"
    raise KeyError(name)


def shape(cov, t):
    if cov == "exp":
        return mp.exp(-t)
    if cov == "mat32":
        return (1 + t) * mp.exp(-t)
    raise KeyError(cov)


def evaluate(y, X, W, locs, theta, cov):
    n, p = X.shape
    q = W.shape[1]
    L = [[mpf(0)] * n for _ in range(n)]
    for i in range(n):
        for j in range(i + 1):
            h = mp.sqrt(sum((mpf(float(locs[i, c])) - mpf(float(locs[j, c]))) ** 2 for c in range(locs.shape[1])))
            s = mpf(0)
            for k in range(q):
                s += mpf(float(theta[2 * k + 1])) * shape(cov, h / mpf(float(theta[2 * k]))) * mpf(float(W[i, k])) * mpf(float(W[j, k]))
            if i == j:
                s += mpf(float(theta[2 * q]))
            L[i][j] = s
    # in-place Cholesky (lower)
    for j in range(n):
        s = L[j][j] - sum(L[j][k] ** 2 for k in range(j))
        L[j][j] = mp.sqrt(s)
        for i in range(j + 1, n):
            L[i][j] = (L[i][j] - sum(L[i][k] * L[j][k] for k in range(j))) / L[j][j]
    B = [[mpf(float(X[i, c])) for c in range(p)] + [mpf(float(y[i]))] for i in range(n)]
    Z = [[mpf(0)] * (p + 1) for _ in range(n)]
    for i in range(n):
        for c in range(p + 1):
            Z[i][c] = (B[i][c] - sum(L[i][k] * Z[k][c] for k in range(i))) / L[i][i]
    G = mp.matrix(p, p)
    g = mp.matrix(p, 1)
    for a in range(p):
        g[a] = sum(Z[i][a] * Z[i][p] for i in range(n))
        for b in range(p):
            G[a, b] = sum(Z[i][a] * Z[i][b] for i in range(n))
    mu = mp.lu_solve(G, g)
    quad = sum((Z[i][p] - sum(Z[i][a] * mu[a] for a in range(p))) ** 2 for i in range(n))
    logdet = 2 * sum(mp.log(L[i][i]) for i in range(n))
    val = n * mp.log(2 * mp.pi) + logdet + quad
    f = lambda v: mp.nstr(v, 25)  # noqa: E731
    return {"n2ll": f(val), "logdet": f(logdet), "quad": f(quad), "mu": [f(mu[a]) for a in range(p)]}


if __name__ == "__main__":
    out = {}
    for name in ("vignette_n100_exp", "synthetic_n260_mat32"):
        y, X, W, locs, theta, cov = case_inputs(name)
        out[name] = {"theta": theta.tolist(), "cov": cov, "n": int(len(y)), **evaluate(y, X, W, locs, theta, cov)}
        print(name, out[name]["n2ll"])
    json.dump(out, open(os.path.join(HERE, "highprec_golden.json"), "w"), indent=1)
