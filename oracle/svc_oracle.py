"""CPU oracle: numpy/LAPACK fp64 restatement of varycoef's likelihood hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``varycoef_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / the CPU
arm that is timed next to the GPU path.

Parity status: PINNED to the only numeric golden values the reference ships --
the knitted vignette ``examples/01_Introduction.html`` (theta-hat, mu-hat, logLik,
BIC, standard errors, first fitted rows; see tests/test_oracle_golden.py).  The
reference's own testthat suite pins no likelihood value.

Third-party arithmetic that is NOT in /root/reference and is restated here from
its published definition (R package ``spam``, version unpinned in
R-pkg/DESCRIPTION:38; base R ``stats::dist``, ``chol``, ``solve``,
``determinant``):

* ``spam::cov.exp/cov.mat/cov.sph/cov.wend1/cov.wend2`` -- with t = h/range:
  value = sill + nugget where t < spam.eps (= .Machine$double.eps), else
  sill*f(t), f as in ``_COV_SHAPES`` below.  ``cov.mat`` is the Matern
  ``2^(1-nu)/Gamma(nu) t^nu K_nu(t)`` (no sqrt(2 nu) scaling), which for
  nu = 3/2, 5/2 is (1+t)e^-t and (1+t+t^2/3)e^-t.
* ``stats::dist`` (C ``R_euclidean`` etc.): direct differences, summed over the
  coordinate columns in order, then sqrt.
* ``spam::nearest.dist(x, delta)``: lower triangle *including the zero diagonal*,
  keeping pairs with distance <= delta.

All reference citations are relative to /root/reference/R-pkg/.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import lapack

SPAM_EPS = 2.220446049250313e-16  # getOption("spam.eps") == .Machine$double.eps

COV_NAMES = ("exp", "mat32", "mat52", "sph", "wend1", "wend2")
DIST_METHODS = ("euclidean", "maximum", "manhattan", "minkowski")


# --------------------------------------------------------------------------
# distances: own_dist, R/utils.R:508-543
# --------------------------------------------------------------------------
def _pair_dist(a, b, method="euclidean", p=2.0):
    """Distance between every row of a (m x d) and b (n x d), direct differences
    accumulated column by column like R's distance.c."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    if b.ndim == 1:
        b = b[:, None]
    m, d = a.shape
    acc = np.zeros((m, b.shape[0]))
    for c in range(d):
        dev = np.abs(a[:, c][:, None] - b[:, c][None, :])
        if method == "euclidean":
            acc += dev * dev
        elif method == "maximum":
            np.maximum(acc, dev, out=acc)
        elif method == "manhattan":
            acc += dev
        elif method == "minkowski":
            acc += dev ** p
        else:
            raise ValueError("unknown dist method %r" % (method,))
    if method == "euclidean":
        return np.sqrt(acc)
    if method == "minkowski":
        return acc ** (1.0 / p)
    return acc


def own_dist(x, y=None, taper=None, method="euclidean", p=2.0):
    """R/utils.R:508-543.  Dense case returns the full n x n matrix (as.matrix(dist(.)));
    tapered case returns (D, mask) with mask = structural non-zeros of the spam
    object: lower triangle incl. diagonal with d <= taper (self distances), or the
    full rectangle with d <= taper (cross distances)."""
    if y is None:
        D = _pair_dist(x, x, method, p)
        if taper is None:
            return D
        mask = np.tril(D <= taper)
        return D, mask
    D = _pair_dist(x, y, method, p)
    if taper is None:
        return D
    return D, (D <= taper)


# --------------------------------------------------------------------------
# covariance functions: MLE.cov.func / cov.mat32 / cov.mat52, R/utils.R:2-29
# --------------------------------------------------------------------------
def _shape_exp(t):
    return np.exp(-t)


def _shape_mat32(t):
    return (1.0 + t) * np.exp(-t)


def _shape_mat52(t):
    return (1.0 + t + t * t / 3.0) * np.exp(-t)


def _shape_sph(t):
    return np.where(t < 1.0, 1.0 - 1.5 * t + 0.5 * t ** 3, 0.0)


def _shape_wend1(t):
    return np.where(t < 1.0, (1.0 - t) ** 4 * (4.0 * t + 1.0), 0.0)


def _shape_wend2(t):
    return np.where(t < 1.0, (1.0 - t) ** 6 * (35.0 * t * t + 18.0 * t + 3.0) / 3.0, 0.0)


_COV_SHAPES = {
    "exp": _shape_exp, "mat32": _shape_mat32, "mat52": _shape_mat52,
    "sph": _shape_sph, "wend1": _shape_wend1, "wend2": _shape_wend2,
}


def cov_func(name):
    """MLE.cov.func(name), R/utils.R:18-29: returns f(h, theta) with
    theta = (range, sill[, nugget]).  cov.mat32/52 insist on length(theta)==2
    (R/utils.R:3,10; pinned by tests/testthat/test_utils.R:34-41)."""
    if name not in _COV_SHAPES:
        raise ValueError("'arg' should be one of %s" % (COV_NAMES,))
    shape = _COV_SHAPES[name]

    def f(h, theta):
        theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
        if name in ("mat32", "mat52") and theta.size != 2:
            raise ValueError("length(theta) == 2L is not TRUE")
        rng = theta[0]
        sill = theta[1] if theta.size > 1 else 1.0
        nug = theta[2] if theta.size > 2 else 0.0
        t = np.asarray(h, dtype=np.float64) / rng
        return np.where(t < SPAM_EPS, sill + nug, sill * shape(t))

    return f


def get_taper(cov_name, d, tapering):
    """R/utils.R:545-552: wend1 taper for exp & mat32, wend2 for mat52, else NULL."""
    if cov_name in ("exp", "mat32"):
        return cov_func("wend1")(d, (tapering, 1.0, 0.0))
    if cov_name == "mat52":
        return cov_func("wend2")(d, (tapering, 1.0, 0.0))
    return None


# --------------------------------------------------------------------------
# Sigma_y, R/utils.R:171-233 (+ outer.W construction R/SVC_mle.R:65-72)
# --------------------------------------------------------------------------
def sigma_y(theta, D, W, cov_name="exp", taper_range=None, mask=None):
    """Full symmetric n x n covariance of y.

    dense branch (R/utils.R:175-194): sum_k cov(theta_k) * (W_k W_k^T) + diag(tau^2).
    tapered branch (R/utils.R:195-232): same on the structural non-zeros of the
    lower-triangular spam distance object, with the Wendland taper pre-folded
    into outer.W (R/SVC_mle.R:69-72), then lower.tri(S) + t(S) + diag(tau^2).
    `mask` is the spam structure (from own_dist); entries outside are exact 0.
    """
    theta = np.asarray(theta, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    n, q = W.shape
    f = cov_func(cov_name)
    if taper_range is None:
        S = np.zeros((n, n))
        for k in range(q):
            C = f(D, theta[2 * k:2 * k + 2])
            S = S + C * np.outer(W[:, k], W[:, k])
        return S + np.diag(np.full(n, theta[2 * q]))
    if mask is None:
        mask = np.tril(D <= taper_range)
    tap = get_taper(cov_name, D, taper_range)
    if tap is None:
        raise ValueError("tapering undefined for cov.name=%r (get_taper returns NULL)" % cov_name)
    S = np.zeros((n, n))
    for k in range(q):
        th = theta[2 * k:2 * k + 2]
        if k > 0:
            th = np.append(th, 0.0)  # R/utils.R:211 passes a 3-vector for k >= 2
        C = f(D, th)
        S = S + C * (np.outer(W[:, k], W[:, k]) * tap)
    S = np.where(mask, S, 0.0)
    lower = np.tril(S, -1)
    upper = np.tril(S).T
    return lower + upper + np.diag(np.full(n, theta[2 * q]))


# --------------------------------------------------------------------------
# GLS_chol.matrix, R/utils.R:93-101  (faithful: dgesv on the triangular factor)
# --------------------------------------------------------------------------
def _dgesv(A, B):
    lu, piv, x, info = lapack.dgesv(A, B)
    if info != 0:
        raise np.linalg.LinAlgError("dgesv info=%d" % info)
    return x


def gls_chol(R, X, y):
    """R is the upper factor (R^T R = Sigma) as base::chol returns it."""
    Rt = np.asfortranarray(R.T)
    RiX = _dgesv(Rt, np.asfortranarray(X))
    Riy = _dgesv(Rt, np.asfortranarray(np.asarray(y).reshape(-1, 1)))
    G = RiX.T @ RiX
    return np.linalg.solve(G, RiX.T @ Riy).ravel()


def chol_upper(S):
    """base::chol -> LAPACK dpotrf('U'); raises like R when a leading minor is not PD."""
    c, info = lapack.dpotrf(np.asfortranarray(S), lower=0, clean=1)
    if info > 0:
        raise np.linalg.LinAlgError("the leading minor of order %d is not positive" % info)
    if info < 0:
        raise ValueError("dpotrf argument %d" % -info)
    return c


def pc_penalty(x, q, pc_prior):
    """R/objective_functions.R:110-123 with the density of R/SVC_mle.R:85-99;
    pc_prior = (rho0, alpha_rho, sigma0, alpha_sigma) or None."""
    if pc_prior is None:
        return 0.0
    lam_r = -math.log(pc_prior[1]) * 2.0 * pc_prior[0]
    lam_s = -math.log(pc_prior[3]) / pc_prior[2]
    s = 0.0
    for j in range(q):
        rho, sig2 = x[2 * j], x[2 * j + 1]
        s += 4.0 * math.log(rho) + lam_r / rho + 2.0 * lam_s * math.sqrt(sig2)
    return s


def n2ll(x, D, X, W, y, cov_name="exp", taper_range=None, mask=None,
         mean_est=None, pc_prior=None, profile=True, faithful=True, parts=False):
    """n2LL, R/objective_functions.R:4-62, dense path.

    faithful=True follows the op sequence R executes (dpotrf, dgesv-based
    solve(t(R), .), sequential log-diag sum as determinant() does on a triangular
    matrix); faithful=False is the lean variant (dtrtrs) with identical maths.
    """
    x = np.asarray(x, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    n, p = X.shape
    q = W.shape[1]
    S = sigma_y(x[:2 * q + 1], D, W, cov_name, taper_range, mask)
    R = chol_upper(S)
    del S
    if profile:
        if mean_est is None:
            if faithful:
                mu = gls_chol(R, X, y)
            else:
                Z, info = lapack.dtrtrs(R, np.asfortranarray(np.column_stack([X, y])), lower=0, trans=1)
                G = Z[:, :p].T @ Z[:, :p]
                mu = np.linalg.solve(G, Z[:, :p].T @ Z[:, p])
        else:
            mu = np.asarray(mean_est, dtype=np.float64)
    else:
        mu = x[2 * q + 1:2 * q + 1 + p]
    res = y - X @ mu
    if faithful:
        z = _dgesv(np.asfortranarray(R.T), np.asfortranarray(res.reshape(-1, 1))).ravel()
    else:
        z, info = lapack.dtrtrs(R, res, lower=0, trans=1)
    quad = float(z @ z)
    # determinant(cholS)$modulus: log|diag| summed sequentially
    logdet_R = float(np.cumsum(np.log(np.abs(np.diag(R))))[-1])
    val = n * math.log(2.0 * math.pi) + 2.0 * logdet_R + quad + pc_penalty(x, q, pc_prior)
    if parts:
        return {"n2ll": val, "mu": mu, "logdet": 2.0 * logdet_R, "quad": quad}
    return val


# --------------------------------------------------------------------------
# med_dist R/SVC_mle.R:39-45 for n too large to hold the distance matrix
# --------------------------------------------------------------------------
def _pair_block(a, b, method, p):
    acc = np.zeros((a.shape[0], b.shape[0]))
    for c in range(a.shape[1]):
        dev = np.abs(a[:, c][:, None] - b[:, c][None, :])
        if method == "euclidean":
            acc += dev * dev
        elif method == "maximum":
            np.maximum(acc, dev, out=acc)
        elif method == "manhattan":
            acc += dev
        else:
            acc += dev ** p
    if method == "euclidean":
        return np.sqrt(acc)
    if method == "minkowski":
        return acc ** (1.0 / p)
    return acc


def median_dist(locs, method="euclidean", p=2.0, block=2048):
    """med_dist, R/SVC_mle.R:39-45 (dense branch): median(as.numeric(d)) over the FULL n x n distance
    matrix, zero diagonal and both triangles included, without holding the matrix: exact selection by a
    histogram pass plus a refinement pass over the strictly-lower pairs.  Checker of csrc/vc_median.cu."""
    locs = np.asarray(locs, dtype=np.float64)
    if locs.ndim == 1:
        locs = locs[:, None]
    n = locs.shape[0]
    N = n * n
    # order statistics needed (0-based) in the full multiset = n zeros + each lower pair twice
    ranks = [N // 2] if N % 2 == 1 else [N // 2 - 1, N // 2]

    def lower_blocks():
        for i0 in range(0, n, block):
            a = locs[i0:i0 + block]
            for j0 in range(0, i0 + 1, block):
                d = _pair_block(a, locs[j0:j0 + block], method, p)
                if j0 == i0:
                    yield d[np.tril_indices(d.shape[0], -1, d.shape[1])]
                else:
                    yield d.ravel()

    if n * (n - 1) // 2 <= 4_000_000:
        low = np.concatenate([b for b in lower_blocks()]) if n > 1 else np.zeros(0)
        low.sort()
        vals = []
        for r in ranks:
            vals.append(0.0 if r < n else float(low[(r - n) // 2]))
        return float(np.mean(vals))
    dmax = 0.0
    for b in lower_blocks():
        if b.size:
            dmax = max(dmax, float(b.max()))
    nb = 1 << 16
    hist = np.zeros(nb, dtype=np.int64)
    scale = nb / (dmax * (1 + 1e-12) + 1e-300)
    for b in lower_blocks():
        hist += np.bincount(np.minimum((b * scale).astype(np.int64), nb - 1), minlength=nb)
    cum = np.cumsum(hist)
    vals = []
    for r in ranks:
        if r < n:
            vals.append(0.0)
            continue
        k = (r - n) // 2                      # rank among the lower pairs
        bi = int(np.searchsorted(cum, k, side="right"))
        before = int(cum[bi - 1]) if bi > 0 else 0
        sel = []
        for b in lower_blocks():
            idx = np.minimum((b * scale).astype(np.int64), nb - 1)
            sel.append(b[idx == bi])
        sel = np.sort(np.concatenate(sel))
        vals.append(float(sel[k - before]))
    return float(np.mean(vals))



# --------------------------------------------------------------------------
# init_bounds_optim R/utils.R:336-396, med_dist / y_var / OLS R/SVC_mle.R:39-54
# --------------------------------------------------------------------------
def check_cov_lower(cv, q):
    """R/utils.R:291-297 (note: only ranges+nugget > 0 and variances >= 0)."""
    cv = np.asarray(cv, dtype=np.float64)
    l_rp = bool(np.all(cv[0:2 * q + 1:2] > 0))
    l_vp = bool(np.all(cv[1:2 * q:2] >= 0))
    return l_rp and l_vp


def ols(X, y):
    return np.linalg.lstsq(np.asarray(X, float), np.asarray(y, float), rcond=None)[0]


def init_bounds_optim(p, q, med_dist, y_var, ols_mu, profile_lik=False):
    lower = [med_dist / 1000.0, 0.0] * q + [1e-6]
    upper = [10.0 * med_dist, 10.0 * y_var] * q + [10.0 * y_var]
    init = [med_dist / 4.0, y_var / (q + 1)] * q + [y_var / (q + 1)]
    if not profile_lik:
        lower += [-np.inf] * p
        upper += [np.inf] * p
        init += list(ols_mu)
    return {"lower": np.array(lower), "init": np.array(init), "upper": np.array(upper)}


def default_liu(D, X, W, y, profile_lik=False, tapered=False):
    """med_dist (R/SVC_mle.R:41-45: dense = median over the FULL matrix incl. zero
    diagonal; tapered = median of a logical vector == 1), var(y), OLS."""
    med = 1.0 if tapered else float(np.median(D))
    y = np.asarray(y, float)
    return init_bounds_optim(X.shape[1], W.shape[1], med, float(np.var(y, ddof=1)),
                             ols(X, y), profile_lik)


# --------------------------------------------------------------------------
# sample_SVCdata, R/example.R:63-134 (numpy RNG instead of R's stream)
# --------------------------------------------------------------------------
def sample_svcdata(rng, locs, var, scale, mean, nugget_sd, cov_name="exp", X=None):
    locs = np.asarray(locs, float)
    if locs.ndim == 1:
        locs = locs[:, None]
    n = locs.shape[0]
    p = len(var)
    D = _pair_dist(locs, locs)
    f = cov_func(cov_name)
    beta = np.empty((n, p))
    for k in range(p):
        if var[k] == 0:
            beta[:, k] = mean[k]
        else:
            L = np.linalg.cholesky(f(D, (scale[k], var[k])))
            beta[:, k] = mean[k] + L @ rng.standard_normal(n)
    eps = rng.standard_normal(n) * nugget_sd
    if X is None:
        X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
    y = (beta * X).sum(axis=1) + eps
    return {"y": y, "X": X, "beta": beta, "eps": eps, "locs": locs}


# --------------------------------------------------------------------------
# next-tier rows: prep_par_output, eff_dof, predict
# --------------------------------------------------------------------------
def prep_par_output(par, S_final, X, y, H, q, profile_lik):
    """R/utils.R:423-473."""
    X = np.asarray(X, float)
    p = X.shape[1]
    if profile_lik:
        mu = gls_chol(chol_upper(S_final), X, y)
    else:
        mu = np.asarray(par[2 * q + 1:2 * q + 1 + p])
    if H is None:
        se_all = np.full(len(par), np.nan)
    else:
        try:
            se_all = np.sqrt(np.diag(np.linalg.inv(np.asarray(H) / 2.0)))
        except np.linalg.LinAlgError:
            se_all = np.full(len(par), np.nan)
    if profile_lik:
        se_re = se_all
        se_fe = np.sqrt(np.diag(np.linalg.inv(X.T @ np.linalg.solve(S_final, X))))
    else:
        se_re = se_all[:2 * q + 1]
        se_fe = se_all[2 * q + 1:2 * q + 1 + p]
    return {"RE": {"est": np.asarray(par[:2 * q + 1]), "SE": se_re}, "FE": {"est": mu, "SE": se_fe}}


def eff_dof(cov_par, X, S):
    """R/eff_dof.R:7-21 with S = Sigma_y(cov_par)."""
    X = np.asarray(X, float)
    n = X.shape[0]
    nug = cov_par[-1]
    iS = np.linalg.inv(S)
    XtiS = X.T @ iS
    return float(nug * np.trace(np.linalg.solve(XtiS @ X, XtiS @ iS @ X)) + n - nug * np.trace(iS))


def predict(cov_par, mu, locs, X, W, y, newlocs=None, newX=None, newW=None,
            cov_name="exp", taper_range=None, compute_y_var=False, method="euclidean", p=2.0):
    """predict.SVC_mle, R/predict-SVC_mle.R:80-267 (EBLUP of the SVC deviations,
    the response and, optionally, its predictive variance)."""
    locs = np.asarray(locs, float)
    if locs.ndim == 1:
        locs = locs[:, None]
    W = np.asarray(W, float)
    X = np.asarray(X, float)
    q = W.shape[1]
    n = W.shape[0]
    if taper_range is None:
        D = own_dist(locs, method=method, p=p)
        mask = None
    else:
        D, mask = own_dist(locs, taper=taper_range, method=method, p=p)
    S = sigma_y(cov_par, D, W, cov_name, taper_range, mask)
    if newlocs is None:
        newlocs = locs
    newlocs = np.asarray(newlocs, float)
    if newlocs.ndim == 1:
        newlocs = newlocs[:, None]
    n_new = newlocs.shape[0]
    Dc = _pair_dist(newlocs, locs, method, p)
    f = cov_func(cov_name)
    if taper_range is None:
        def cf_cross(th):
            return f(Dc, th)
    else:
        tc = np.where(Dc <= taper_range, get_taper(cov_name, Dc, taper_range), 0.0)

        def cf_cross(th):
            return f(Dc, th) * tc
    alpha = np.linalg.solve(S, np.asarray(y, float) - X @ mu)
    eff = np.empty((n_new, q))
    for k in range(q):
        # Sigma_b_y, R/utils.R:238-249
        eff[:, k] = (cf_cross(cov_par[2 * k:2 * k + 2]) * W[:, k][None, :]) @ alpha
    out = {"eff": eff}
    if newX is not None and newW is not None:
        newX = np.asarray(newX, float)
        newW = np.asarray(newW, float)
        out["y_pred"] = (newW * eff).sum(axis=1) + newX @ mu
        if compute_y_var:
            B = np.zeros((n_new, n))
            for k in range(q):  # Sigma_y_y, R/utils.R:252-269
                B += cf_cross(np.append(cov_par[2 * k:2 * k + 2], 0.0) if cov_name not in ("mat32", "mat52")
                              else cov_par[2 * k:2 * k + 2]) * np.outer(newW[:, k], W[:, k])
            if taper_range is None:
                Dn = own_dist(newlocs, method=method, p=p)
                An = sigma_y(cov_par, Dn, newW, cov_name)
            else:
                Dn, mn = own_dist(newlocs, taper=taper_range, method=method, p=p)
                An = sigma_y(cov_par, Dn, newW, cov_name, taper_range, mn)
            out["y_var"] = np.diag(An) - np.einsum("ij,ji->i", B, np.linalg.solve(S, B.T))
    return out
