"""Independent third path for sizes no CPU oracle finishes in seconds: cuSOLVER (through torch) factors
the GPU-built Sigma_y, cuBLAS solves, and -2 logL is assembled on the host.

CHECKER ONLY -- used by tests/ and by the `parity` / `library_bar` records of bench.py (outside every
timed region).  The product path never calls torch, cuSOLVER or cuBLAS.

What is shared with the product path: the covariance build (vc_sigma_device; that kernel is checked
element-wise against the oracle in tests/test_gpu_parity.py).  What is independent: the Cholesky
factorisation, the triangular solves, the GLS normal equations, the log-determinant.
Reference op sequence: R-pkg/R/objective_functions.R:23-61, R-pkg/R/utils.R:93-101.
"""
from __future__ import annotations

import math

import numpy as np


def third_path_n2ll(obj, theta, X, y, device=None):
    """Returns {"n2ll", "logdet", "quad", "mu", "potrf_ms"}; potrf_ms is the CUDA-event time of
    torch.linalg.cholesky_ex (cusolverDn potrf plus torch's copy of the input)."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    n, p = obj.n, obj.p
    S = torch.empty((n, n), dtype=torch.float64, device=dev)      # symmetric: row- vs column-major is moot
    torch.cuda.synchronize(dev)
    obj.Sigma_y_device(theta, S.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L, info = torch.linalg.cholesky_ex(S)
    e1.record()
    torch.cuda.synchronize(dev)
    del S
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError("cuSOLVER potrf: leading minor of order %d is not positive" % int(info.item()))
    logdet = 2.0 * float(torch.log(torch.diagonal(L)).sum().item())
    B = torch.from_numpy(np.ascontiguousarray(np.column_stack([X, y]))).to(dev)
    Z = torch.linalg.solve_triangular(L, B, upper=False)
    del L
    G = (Z.T @ Z).cpu().numpy()
    mu = np.linalg.solve(G[:p, :p], G[:p, p])
    r = Z[:, p] - Z[:, :p] @ torch.from_numpy(mu).to(dev)
    quad = float((r @ r).item())
    return {"n2ll": n * math.log(2.0 * math.pi) + logdet + quad, "logdet": logdet, "quad": quad, "mu": mu,
            "potrf_ms": float(e0.elapsed_time(e1))}


def library_bar(n_potrf, device=None, dgemm_n=8192, reps=3):
    """cublasDgemm (dgemm_n^3) and cusolverDn potrf (n_potrf, lower) timed with CUDA events on this GPU: the
    bars SURVEY.md section 2a / 7.4 set for the hand-written kernels.  Best of `reps`."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    out = {}
    a = torch.randn((dgemm_n, dgemm_n), dtype=torch.float64, device=dev)
    b = torch.randn((dgemm_n, dgemm_n), dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    out["dgemm_n"] = dgemm_n
    out["dgemm_ms"] = best
    out["dgemm_tflops"] = 2.0 * dgemm_n ** 3 / (best * 1e-3) / 1e12
    del a, b, c
    if n_potrf:
        # SPD test matrix: diagonally dominant, built in place (the values do not matter for the timing)
        S = torch.rand((n_potrf, n_potrf), dtype=torch.float64, device=dev)
        S.mul_(1.0 / n_potrf)
        S.diagonal().add_(2.0)
        L = torch.empty_like(S)
        info = torch.empty((), dtype=torch.int32, device=dev)
        best = float("inf")
        for _ in range(max(1, reps - 1)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.linalg.cholesky_ex(S, out=(L, info))
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        out["potrf_n"] = n_potrf
        out["potrf_ms"] = best
        out["potrf_tflops"] = n_potrf ** 3 / 3.0 / (best * 1e-3) / 1e12
        out["potrf_note"] = ("torch.linalg.cholesky_ex -> cusolverDn potrf, lower; includes torch's device copy of the "
                             "input (%.1f GB, a few ms)" % (8.0 * n_potrf * n_potrf / 1e9))
        del S, L
    torch.cuda.empty_cache()
    return out
