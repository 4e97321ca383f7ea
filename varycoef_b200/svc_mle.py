"""Host-side mirror of the fit driver around the GPU objective.

Reference (R-pkg/R/): SVC_mle_control SVC_mle.R:427-475; SVC_mle.default SVC_mle.R:649-693;
MLE_computation SVC_mle.R:12-242; init_bounds_optim utils.R:336-396; prep_par_output
utils.R:423-473; eff_dof eff_dof.R:7-21; create_SVC_mle SVC_mle.R:254-299; predict.SVC_mle
predict-SVC_mle.R:80-267; logLik/BIC logLik-SVC_mle.R:22-34, BIC-SVC_mle.R:23-43.

In the reference this layer is R and stays R (INTEGRATION.md shows the `.Call` shim); R is not
available in the build image, so the same control flow is restated here in Python on top of the
same C ABI so that fits can be run and checked end to end.  Nothing n x n is ever held on the
host: distances, outer products, Sigma_final and its factor live on the GPU.
"""
from __future__ import annotations

import math
import warnings

import numpy as np

from .objective import SVCObjective, check_cov_lower
from .optim import optim_lbfgsb


def SVC_mle_control(cov_name="exp", tapering=None, parallel=None, init=None, lower=None, upper=None,
                    save_fitted=True, profileLik=False, mean_est="GLS", pc_prior=None,
                    extract_fun=False, hessian=True, dist=None, parscale=True, **extra):
    """SVC_mle_control.default (SVC_mle.R:427-475); same defaults, same checks."""
    names = ("exp", "sph", "mat32", "mat52", "wend1", "wend2")
    if cov_name not in names:
        raise ValueError("'arg' should be one of %s" % (names,))
    if mean_est not in ("GLS", "OLS"):
        raise ValueError("'arg' should be one of ('GLS', 'OLS')")
    if not (tapering is None or tapering >= 0):
        raise ValueError("is.null(tapering) | (tapering >= 0) is not TRUE")
    for nm, v in (("save_fitted", save_fitted), ("profileLik", profileLik),
                  ("extract_fun", extract_fun), ("hessian", hessian), ("parscale", parscale)):
        if not isinstance(v, (bool, np.bool_)):
            raise ValueError("is.logical(%s) is not TRUE" % nm)
    ctl = dict(cov_name=cov_name, tapering=tapering, parallel=parallel, init=init, lower=lower,
               upper=upper, save_fitted=save_fitted, profileLik=profileLik, mean_est=mean_est,
               pc_prior=pc_prior, extract_fun=extract_fun, hessian=hessian,
               dist=dict(dist or {"method": "euclidean"}), parscale=parscale)
    ctl.update(extra)
    return ctl


def init_bounds_optim(control, p, q, id_obj, med_dist, y_var, OLS_mu):
    """utils.R:336-396, including its error messages."""
    prof = control["profileLik"]
    tail_lo = [] if prof else [-np.inf] * p
    tail_hi = [] if prof else [np.inf] * p
    if control.get("lower") is None:
        lower = np.array([med_dist / 1000.0, 0.0] * q + [1e-6] + tail_lo)
    else:
        lower = np.asarray(control["lower"], dtype=np.float64)
        if lower.size != len(id_obj):
            raise ValueError("Lower boundary vector has wrong length. Check SVC_mle_control.")
        if not check_cov_lower(lower[:2 * q + 1], q):
            raise ValueError("Lower boundary vector is not greater (or equal) than 0. Call ?check_cov_lower.")
    if control.get("upper") is None:
        upper = np.array([10.0 * med_dist, 10.0 * y_var] * q + [10.0 * y_var] + tail_hi)
    else:
        upper = np.asarray(control["upper"], dtype=np.float64)
        if upper.size != len(id_obj):
            raise ValueError("Upper boundary vector has wrong length. Check SVC_mle_control.")
        if not check_cov_lower(upper[:2 * q + 1], q):
            raise ValueError("Upper boundary vector is not greater (or equal) than 0. Call ?check_cov_lower.")
        if np.any(lower > upper):
            raise ValueError("Upper boundary vector smaller than lower boundary.")
    if control.get("init") is None:
        init = np.array([med_dist / 4.0, y_var / (q + 1)] * q + [y_var / (q + 1)]
                        + ([] if prof else list(OLS_mu)))
    else:
        init = np.asarray(control["init"], dtype=np.float64)
        if init.size != len(id_obj):
            raise ValueError("Initial vector has wrong length. Check SVC_mle_control.")
        if not check_cov_lower(init[:2 * q + 1], q):
            raise ValueError("Initial value vector is not greater (or equal) than 0. Call ?check_cov_lower.")
        if not (np.all(lower <= init) and np.all(init <= upper)):
            raise ValueError("Initial values do not lie between lower and upper boundarys.")
    return {"lower": lower, "init": init, "upper": upper}


def median_dist(locs, method="euclidean", p=2.0, device=0):
    """median(as.numeric(d)) over the FULL n x n distance matrix, zero diagonal and both triangles included
    (SVC_mle.R:39-45), on the GPU: exact radix select over recomputed tile-local distances
    (csrc/vc_median.cu), nothing n x n is ever held.  Bit-identical to the host median of stats::dist."""
    import ctypes as C
    from . import _lib
    locs = _lib.as_f64(locs, 2)
    out = C.c_double()
    _lib.check(_lib.lib().vc_median_dist(int(device), locs.shape[0], locs.shape[1], _lib.ptr(locs),
                                         _lib.DIST_IDS[method], float(p), C.byref(out), None), None)
    return out.value


def _ols(X, y):
    return np.linalg.lstsq(X, y, rcond=None)[0]


class SVCFit:
    """The `SVC_mle` object (create_SVC_mle, SVC_mle.R:254-299) with the S3 accessors that
    sit on the likelihood path: coef, cov_par, logLik, BIC, nobs, fitted, residuals, predict."""

    def __init__(self, obj, optim_output, liu, par_SE, edof, control, optim_control):
        self.objective = obj
        self.optim_output = optim_output
        self.liu = liu
        self.par_SE = par_SE
        self.control = control
        self.optim_control = optim_control
        self.coefficients = np.asarray(par_SE["FE"]["est"])
        self.cov_par = np.asarray(par_SE["RE"]["est"])
        q = obj.q
        df = int(np.sum(np.abs(np.concatenate([self.coefficients, self.cov_par[1:2 * q:2]])) > 1e-10))
        self.df = {"df": df, "edof": edof}
        self.fitted = None
        self.residuals = None

    def coef(self):
        return self.coefficients

    def nobs(self):
        return self.objective.n

    def logLik(self):
        """-1/2 of the optimiser's final value (logLik-SVC_mle.R:27)."""
        return -0.5 * self.optim_output["value"]

    def BIC(self):
        """BIC-SVC_mle.R:23-27: -2 logLik + log(n) * df."""
        return -2.0 * self.logLik() + math.log(self.nobs()) * self.df["df"]

    def predict(self, newlocs=None, newX=None, newW=None, compute_y_var=False):
        return self.objective.predict(self.cov_par, self.coefficients, newlocs, newX, newW, compute_y_var)


def prep_par_output(output_par, obj, profileLik, H, q):
    """utils.R:423-473 with Sigma_final replaced by the GPU-resident factor."""
    p = obj.p
    output_par = np.asarray(output_par, dtype=np.float64)
    post = None
    if profileLik:
        post = obj.post(output_par[:2 * q + 1], edof=False)
        mu = post["mu_gls"]
    else:
        mu = output_par[2 * q + 1:2 * q + 1 + p]
    if H is None:
        warnings.warn("MLE without Hessian. Cannot return standard errors of covariance parameters.")
        se_all = np.full(output_par.size, np.nan)
    else:
        try:
            with np.errstate(invalid="ignore"):
                se_all = np.sqrt(np.diag(np.linalg.inv(np.asarray(H) / 2.0)))
        except np.linalg.LinAlgError:
            warnings.warn("Could not invert Hessian.")
            se_all = np.full(output_par.size, np.nan)
    if profileLik:
        se_re = se_all
        se_fe = np.sqrt(np.diag(post["Sigma_FE"]))
    else:
        se_re = se_all[:2 * q + 1]
        se_fe = se_all[2 * q + 1:2 * q + 1 + p]
    return {"RE": {"est": output_par[:2 * q + 1], "SE": se_re}, "FE": {"est": mu, "SE": se_fe}}


def SVC_mle(y, X, locs, W=None, control=None, optim_control=None, device=0, compute_edof=True):
    """SVC_mle.default (SVC_mle.R:649-693) -> MLE_computation (SVC_mle.R:12-242).

    Returns an SVCFit, or -- with control["extract_fun"] -- the SVCObjective bundle
    (the reference returns list(obj_fun, args), SVC_mle.R:111-132)."""
    control = SVC_mle_control() if control is None else dict(control)
    optim_control = dict(optim_control or {})
    # control$parallel (SVC_mle.R:164-200: the optimParallel cluster) names the GPUs here: a list of device
    # ordinals gives a multi-device objective whose batches (optim's stencils) are spread over them
    if isinstance(control.get("parallel"), (list, tuple)) and len(control["parallel"]) > 0:
        device = [int(v) for v in control["parallel"]]
    dev0 = device[0] if isinstance(device, (list, tuple)) else int(device)
    y = np.asarray(y, dtype=np.float64).ravel()
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    W = X if W is None else np.asarray(W, dtype=np.float64)
    locs = np.asarray(locs, dtype=np.float64)
    if locs.ndim == 1:
        locs = locs[:, None]
    p, q = X.shape[1], W.shape[1]
    prof = control["profileLik"]
    id_obj = list(range(2 * q + 1 if prof else 2 * q + 1 + p))
    if control.get("lower") is None or control.get("upper") is None or control.get("init") is None:
        if control["tapering"] is None:
            med = median_dist(locs, control["dist"].get("method", "euclidean"),
                              float(control["dist"].get("p", 2.0)), device=dev0)
        else:
            med = 1.0   # median(lower.tri(d@entries)) of a logical matrix (SVC_mle.R:44): always 1
        y_var = float(np.var(y, ddof=1))
        ols_mu = _ols(X, y)
    else:
        med = y_var = ols_mu = None
    liu = init_bounds_optim(control, p, q, id_obj, med, y_var, ols_mu)
    mean_est = None if control["mean_est"] == "GLS" else _ols(X, y)
    obj = SVCObjective(y, X, locs, W, cov_name=control["cov_name"], tapering=control["tapering"],
                       dist=control["dist"], profileLik=prof, mean_est=mean_est,
                       pc_prior=control["pc_prior"], device=device)
    if control["extract_fun"]:
        return obj
    if control["parscale"]:
        init = liu["init"]
        optim_control["parscale"] = np.abs(np.where(init == 0, 0.001, init))
    out = optim_lbfgsb(liu["init"], fn=obj.n2LL, fn_batch=obj.n2LL_batch, lower=liu["lower"],
                       upper=liu["upper"], hessian=control["hessian"], control=optim_control)
    par_SE = prep_par_output(out["par"], obj, prof, out["hessian"], q)
    edof = float("nan")
    if compute_edof:
        edof = obj.post(par_SE["RE"]["est"])["edof"]
    fit = SVCFit(obj, out, liu, par_SE, edof, control, optim_control)
    if control["save_fitted"]:
        pred = fit.predict(newlocs=None, newX=X, newW=W)
        fit.fitted = pred
        fit.residuals = y - pred["y_pred"]
    return fit
