"""`stats::optim(method = "L-BFGS-B")` semantics for the fit driver (host logic, stays on the CPU).

The reference calls base R's optim with numerical gradients (R-pkg/R/SVC_mle.R:145-163).  base R is
not part of the reference repository; this module restates the parts of R's optim.c that shape
which points the objective is evaluated at:

* the optimiser works on par/parscale and fn/fnscale;
* with no `gr`, the gradient is a central difference in the scaled space with step `ndeps`
  (default 1e-3); with bounds the probes are clipped to [lower, upper] and the divisor is the
  span actually used (fmingr);
* `hessian = TRUE` -> optimhess: central differences of that numerical gradient (without bounds),
  step ndeps/parscale in the scaled space, symmetrised;
* defaults maxit = 100, factr = 1e7, pgtol = 0, lmm = 5;
* a non-finite objective value is an error ("L-BFGS-B needs finite values of 'fn'").

The L-BFGS-B iteration itself is scipy's (Fortran L-BFGS-B 3.0; R ships the 2.x C translation),
so iteration counts can differ slightly from R while optima agree.

What is B200-specific: the value at a point and all probes of its gradient (1 + 2 npar points), or the
whole Hessian stencil, go to the objective in ONE batch call (`fn_batch`), which the GPU library sweeps
with a single launch sequence (evaluation slots, shared factorisations for repeated thetas) -- the role
optimParallel plays in the reference (R-pkg/R/SVC_mle.R:164-200).
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import minimize


class OptimResult(dict):
    __getattr__ = dict.get


def _as_batch(fn, fn_batch):
    if fn_batch is not None:
        return fn_batch
    return lambda xs: np.array([fn(x) for x in xs])


def fd_points(x_scaled, parscale, ndeps, lower_s=None, upper_s=None):
    """Probe points (original space) and divisors of R's fmingr at scaled point x."""
    npar = x_scaled.size
    pts = np.empty((2 * npar, npar))
    div = np.empty(npar)
    for i in range(npar):
        eps = ndeps[i]
        hi = x_scaled[i] + eps
        eps_hi = eps
        if upper_s is not None and hi > upper_s[i]:
            hi = upper_s[i]
            eps_hi = hi - x_scaled[i]
        lo = x_scaled[i] - eps
        eps_lo = eps
        if lower_s is not None and lo < lower_s[i]:
            lo = lower_s[i]
            eps_lo = x_scaled[i] - lo
        a = x_scaled.copy()
        a[i] = hi
        b = x_scaled.copy()
        b[i] = lo
        pts[2 * i] = a * parscale
        pts[2 * i + 1] = b * parscale
        div[i] = eps_hi + eps_lo
    return pts, div


def optimhess(par, fn=None, fn_batch=None, parscale=None, ndeps=None, fnscale=1.0):
    """R's optimhess for a function without analytic gradient; one batch of 4 npar^2 probes."""
    par = np.asarray(par, dtype=np.float64)
    npar = par.size
    parscale = np.ones(npar) if parscale is None else np.asarray(parscale, dtype=np.float64)
    ndeps = np.full(npar, 1e-3) if ndeps is None else np.asarray(ndeps, dtype=np.float64)
    fb = _as_batch(fn, fn_batch)
    dpar = par / parscale
    pts, divs, eps_list = [], [], []
    for i in range(npar):
        eps = ndeps[i] / parscale[i]
        eps_list.append(eps)
        for sgn in (+1.0, -1.0):
            d = dpar.copy()
            d[i] += sgn * eps
            pp, dv = fd_points(d, parscale, ndeps)
            pts.append(pp)
            divs.append(dv)
    vals = np.asarray(fb(np.vstack(pts)), dtype=np.float64) / fnscale
    if not np.all(np.isfinite(vals)):
        raise FloatingPointError("non-finite finite-difference value in optimhess")
    vals = vals.reshape(2 * npar, 2 * npar)
    H = np.empty((npar, npar))
    for i in range(npar):
        df1 = (vals[2 * i, 0::2] - vals[2 * i, 1::2]) / divs[2 * i]
        df2 = (vals[2 * i + 1, 0::2] - vals[2 * i + 1, 1::2]) / divs[2 * i + 1]
        H[i] = fnscale * (df1 - df2) / (2.0 * eps_list[i] * parscale[i] * parscale)
    return 0.5 * (H + H.T)


def optim_lbfgsb(par, fn=None, fn_batch=None, lower=None, upper=None, hessian=False, control=None):
    """optim(par, fn, method = "L-BFGS-B", lower, upper, hessian, control).

    fn(x) -> float, and/or fn_batch(xs[m, npar]) -> m floats (preferred for the probes).
    control: parscale, fnscale, ndeps, maxit, factr, pgtol, lmm, trace.
    Returns par, value, counts = (fn, gr) at optimiser level, convergence, message, hessian and
    `objective_evals` (every objective evaluation, probes included).
    """
    control = dict(control or {})
    par = np.asarray(par, dtype=np.float64)
    npar = par.size
    parscale = np.asarray(control.get("parscale", np.ones(npar)), dtype=np.float64)
    fnscale = float(control.get("fnscale", 1.0))
    ndeps = np.asarray(control.get("ndeps", np.full(npar, 1e-3)), dtype=np.float64)
    maxit = int(control.get("maxit", 100))
    factr = float(control.get("factr", 1e7))
    pgtol = float(control.get("pgtol", 0.0))
    lmm = int(control.get("lmm", 5))
    lower = np.full(npar, -np.inf) if lower is None else np.asarray(lower, dtype=np.float64)
    upper = np.full(npar, np.inf) if upper is None else np.asarray(upper, dtype=np.float64)
    lower_s, upper_s = lower / parscale, upper / parscale
    fb = _as_batch(fn, fn_batch)
    counts = {"fn": 0, "gr": 0, "evals": 0}

    # L-BFGS-B asks for the value and the gradient at EVERY point it visits (R's lbfgsb.c calls fn then gr; scipy's
    # driver calls fun_and_grad), and the numeric gradient is a stencil of 2 npar more evaluations around that point.
    # On a GPU one batch of 1 + 2 npar evaluations costs little more than a single evaluation (vc_n2ll_batch sweeps
    # them with one launch sequence), so the value call evaluates the stencil along with it and the gradient call
    # that follows at the same point is served from that batch.  Counts keep R's meaning (optimiser-level calls).
    memo = {"x": None, "vals": None, "div": None}

    def f_scaled(xs):
        counts["fn"] += 1
        pts, div = fd_points(xs, parscale, ndeps, lower_s, upper_s)
        vals = np.asarray(fb(np.vstack([(xs * parscale)[None, :], pts])), dtype=np.float64) / fnscale
        counts["evals"] += vals.size
        v = float(vals[0])
        if not np.isfinite(v):
            raise FloatingPointError("L-BFGS-B needs finite values of 'fn'")
        memo["x"], memo["vals"], memo["div"] = np.array(xs, copy=True), vals[1:], div
        return v

    def g_scaled(xs):
        counts["gr"] += 1
        if memo["x"] is not None and np.array_equal(memo["x"], xs):
            vals, div = memo["vals"], memo["div"]
        else:
            pts, div = fd_points(xs, parscale, ndeps, lower_s, upper_s)
            vals = np.asarray(fb(pts), dtype=np.float64) / fnscale
            counts["evals"] += vals.size
        if not np.all(np.isfinite(vals)):
            raise FloatingPointError("L-BFGS-B needs finite values of 'fn'")
        return (vals[0::2] - vals[1::2]) / div

    bounds = [(None if not np.isfinite(lo) else lo, None if not np.isfinite(hi) else hi)
              for lo, hi in zip(lower_s, upper_s)]
    res = minimize(f_scaled, par / parscale, jac=g_scaled, method="L-BFGS-B", bounds=bounds,
                   options={"maxiter": maxit, "ftol": factr * np.finfo(float).eps, "gtol": pgtol,
                            "maxcor": lmm, "maxls": 20, "maxfun": 10 ** 9})
    out = OptimResult()
    out["par"] = res.x * parscale
    out["value"] = float(res.fun) * fnscale
    out["counts"] = {"function": counts["fn"], "gradient": counts["gr"]}
    msg = res.message.decode() if isinstance(res.message, bytes) else str(res.message)
    if res.status == 0:
        out["convergence"] = 0
    elif res.status == 1:
        out["convergence"] = 1      # maxit reached
    else:
        out["convergence"] = 52     # R: error / warning from the L-BFGS-B code
    out["message"] = msg
    out["iterations"] = int(res.nit)
    if hessian:
        out["hessian"] = optimhess(out["par"], fn_batch=fb, parscale=parscale, ndeps=ndeps, fnscale=fnscale)
        counts["evals"] += 4 * npar * npar
    else:
        out["hessian"] = None
    out["objective_evals"] = counts["evals"]
    return out
