"""CPU restatement of the reference's SVC_selection inner loops -- TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module; the product (varycoef_b200/) never does.

Follows R-pkg/R/SVC_selection.R: CD_mu :17-58 (explicit solve(t(R)) whitening + LASSO), CD_theta :60-123,
PMLE_CD :126-262, BW_pen / VB_pen :1-14, with the dense numpy pieces of oracle/svc_oracle.py.

Parity unpinned: the reference's test-suite holds no value for this path (tests/testthat/
test_SVC_selection.R checks one error only) and glmnet is not available here.  The LASSO is solved by an
independent implementation (scikit-learn's coordinate descent, tol 1e-14) of glmnet's published objective
    1/(2n) ||y - X b||^2 + lambda sum_j vq_j xs_j |b_j|
(standardize = TRUE, intercept = FALSE: xs_j = sqrt(mean(x_j^2) - mean(x_j)^2), vq = pf * p / sum(pf)).
"""
import math

import numpy as np

from . import svc_oracle as so


def lasso_glmnet_sklearn(x, y, lam, penalty_factor=None):
    from sklearn.linear_model import Lasso
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    n, p = x.shape
    pf = np.ones(p) if penalty_factor is None else np.asarray(penalty_factor, dtype=np.float64).copy()
    excl = ~np.isfinite(pf)
    pf[excl] = 1.0
    vq = pf * p / pf.sum()
    xs = np.sqrt((x ** 2).mean(0) - x.mean(0) ** 2)
    cj = vq * xs                                    # b'_j = cj b_j, x'_j = x_j / cj  =>  plain LASSO in b'
    keep = np.where(~excl)[0]
    beta = np.zeros(p)
    if lam <= 0 or keep.size == 0:
        if keep.size:
            beta[keep] = np.linalg.lstsq(x[:, keep], y, rcond=None)[0]
        return beta
    m = Lasso(alpha=lam, fit_intercept=False, tol=1e-14, max_iter=200000)
    m.fit(x[:, keep] / cj[keep], y)
    beta[keep] = m.coef_ / cj[keep]
    return beta


def _n2ll(model, x, mean_est):
    return so.n2ll(x, model["D"], model["X"], model["W"], model["y"], model["cov_name"],
                   mean_est=mean_est, pc_prior=model.get("pc_prior"), profile=True, faithful=False)


def cd_mu(theta_k, mle_mu, model, lambda_mu, adaptive=False):
    X, y, W = model["X"], model["y"], model["W"]
    q = W.shape[1]
    S = so.sigma_y(np.asarray(theta_k)[:2 * q + 1], model["D"], W, model["cov_name"])
    R = so.chol_upper(S)
    Rti = np.linalg.inv(R.T)                         # solve(t(R)), :32
    yt, Xt = Rti @ y, Rti @ X
    with np.errstate(divide="ignore"):
        pf = 1.0 / np.abs(np.asarray(mle_mu, dtype=np.float64)) if adaptive else np.ones(X.shape[1])
    return lasso_glmnet_sklearn(Xt, yt, lambda_mu, pf)


def cd_theta(mu_k, theta_k, mle_par, model, lambda_theta, optim, optim_args=None, adaptive=False):
    q = model["W"].shape[1]
    n = model["y"].shape[0]
    vi = 2 * np.arange(q) + 1
    if adaptive:
        a = np.abs(np.asarray(mle_par)[vi])
        with np.errstate(divide="ignore"):
            lambda2 = np.where(a < 1e-9, 1e99, lambda_theta / a)
    else:
        lambda2 = np.full(q, float(lambda_theta))

    def pl(x):
        return _n2ll(model, x, mu_k) + 2.0 * n * float(np.sum(lambda2 * np.abs(x[vi])))

    a = dict(optim_args or {})
    return optim(np.asarray(theta_k, dtype=np.float64), fn=pl, lower=a.get("lower"), upper=a.get("upper"),
                 control=a.get("control"))["par"]


def pmle_cd(lam, mle_par, model, optim, optim_args=None, adaptive=False, IC_type="BIC", CD_conv=None):
    conv = dict(CD_conv or {"N": 20, "delta": 1e-6, "logLik": True})
    N = int(conv["N"])
    X, y, W = model["X"], model["y"], model["W"]
    n, p = X.shape
    q = W.shape[1]
    c_par = np.full((N + 1, 2 * q + 1), np.nan)
    mu_par = np.full((N + 1, p), np.nan)
    ll = np.full(N + 1, np.nan)
    c_par[0] = mle_par
    iS = np.linalg.inv(so.sigma_y(c_par[0], model["D"], W, model["cov_name"]))       # :146-150
    B = X.T @ iS
    mu_par[0] = np.linalg.solve(B @ X, B @ y)
    ll[0] = -0.5 * _n2ll(model, c_par[0], mu_par[0])
    k = 0
    for k in range(1, N + 1):
        mu_par[k] = cd_mu(c_par[k - 1], mu_par[0], model, lam[0], adaptive)
        c_par[k] = cd_theta(mu_par[k], c_par[k - 1], c_par[0], model, lam[1], optim, optim_args, adaptive)
        ll[k] = -0.5 * _n2ll(model, c_par[k], mu_par[k])
        if conv.get("logLik", True):
            if abs(ll[k - 1] - ll[k]) / abs(ll[k - 1]) < conv["delta"]:
                break
        else:
            prev = np.concatenate([c_par[k - 1], mu_par[k - 1]])
            cur = np.concatenate([c_par[k], mu_par[k]])
            if np.sum(np.abs(prev - cur)) / np.sum(np.abs(prev)) < conv["delta"]:
                break
    if IC_type == "BIC":
        MC = math.log(n) * float(np.sum(np.concatenate([mu_par[k], c_par[k][2 * np.arange(q) + 1]]) > 1e-10))
    else:
        S = so.sigma_y(c_par[k], model["D"], W, model["cov_name"])
        e = so.eff_dof(c_par[k], X, S)
        MC = 2.0 * (e + 2 * q + 1) if IC_type == "cAIC_BW" else (2.0 * n) / (n - p - 2) * (e + 1 - (e - p) / (n - p))
    return {"mu.par": mu_par, "c.par": c_par, "loglik.CD": ll,
            "final.IC": _n2ll(model, c_par[k], mu_par[k]) + MC, "n.iter": k}
