"""Host-side mirror of the reference's objective-function contract on top of the C ABI.

Reference (all paths relative to R-pkg/R/):
  n2LL          objective_functions.R:4-62      -> SVCObjective.n2LL / n2LL()
  pc_penalty    objective_functions.R:110-123   -> pc_penalty()   (host scalar, as in R)
  Sigma_y       utils.R:171-233                 -> SVCObjective.Sigma_y
  GLS_chol      utils.R:77-101 (exported)       -> GLS_chol()
  the bundle list(obj_fun, args) of SVC_mle_control(extract_fun=TRUE), SVC_mle.R:111-132
                                                -> SVCObjective (callable: obj(x))
Distances / outer products / taper matrices are never materialised on the host: the handle
keeps the O(n) inputs on the GPU and every evaluation runs build -> Cholesky -> finalize there.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import lib, as_f64, ptr, check


def pc_penalty(x, q, pc_prior):
    """objective_functions.R:110-123 with the -2 log PC-prior density of SVC_mle.R:85-99.
    pc_prior = (rho_0, alpha_rho, sigma_0, alpha_sigma) or None."""
    if pc_prior is None:
        return 0.0
    lam_r = -math.log(pc_prior[1]) * 2.0 * pc_prior[0]
    lam_s = -math.log(pc_prior[3]) / pc_prior[2]
    s = 0.0
    for j in range(q):
        rho, sig2 = float(x[2 * j]), float(x[2 * j + 1])
        s += 4.0 * math.log(rho) + lam_r / rho + 2.0 * lam_s * math.sqrt(sig2)
    return s


def taper_for(cov_name):
    """get_taper, utils.R:545-552: which Wendland taper goes with which covariance."""
    return {"exp": "wend1", "mat32": "wend1", "mat52": "wend2"}.get(cov_name)


class SVCObjective:
    """GPU-resident objective of one SVC model: the drop-in for `fn = n2LL` and its `args`.

    Parameters mirror MLE_computation's inputs (SVC_mle.R:12-14): y (n), X (n x p), locs (n x d or n),
    W (n x q, defaults to X as in SVC_mle.default, SVC_mle.R:660), and the control entries that
    shape the objective: cov_name, tapering, dist (method, p), profileLik, mean_est, pc_prior.
    """

    def __init__(self, y, X, locs, W=None, cov_name="exp", tapering=None, dist=None,
                 profileLik=False, mean_est=None, pc_prior=None, device=0, nb_outer=0,
                 lookahead=True, legacy_syrk=False, legacy_potrf=False, _dist_grid=None):
        self._h = C.c_void_p()
        self.y = as_f64(y).ravel()
        self.X = as_f64(X, 2)
        self.W = self.X if W is None else as_f64(W, 2)
        self.locs = as_f64(locs, 2)
        n = self.y.shape[0]
        if self.X.shape[0] != n or self.W.shape[0] != n or self.locs.shape[0] != n:
            raise ValueError("y, X, W and locs must have the same number of rows")
        self.n, self.p, self.q, self.d = n, self.X.shape[1], self.W.shape[1], self.locs.shape[1]
        if cov_name not in _lib.COV_IDS:
            raise ValueError("'arg' should be one of %s" % (tuple(_lib.COV_IDS),))
        dist = dict(dist or {"method": "euclidean"})
        method = dist.get("method", "euclidean")
        if method not in _lib.DIST_IDS:
            raise ValueError("unsupported dist method %r" % (method,))
        self.cov_name, self.tapering, self.dist = cov_name, tapering, dist
        self.profileLik, self.mean_est, self.pc_prior = bool(profileLik), mean_est, pc_prior
        cfg = _lib.VcConfig()
        check(lib().vc_config_default(C.byref(cfg)))
        cfg.cov_id = _lib.COV_IDS[cov_name]
        cfg.dist_method = _lib.DIST_IDS[method]
        cfg.minkowski_p = float(dist.get("p", 2.0))
        if tapering is not None:
            tap = taper_for(cov_name)
            if tap is None:
                # get_taper returns NULL for sph/wend1/wend2 (utils.R:545-552): tapering undefined
                raise ValueError("covariance tapering is only defined for exp, mat32, mat52")
            if cov_name in ("mat32", "mat52") and self.q > 1:
                # utils.R:211 hands a length-3 theta to cov.mat32/52, which stop() (utils.R:3,10)
                raise ValueError("length(theta) == 2L is not TRUE (tapering with %s needs q = 1)" % cov_name)
            cfg.taper_id = _lib.TAPER_IDS[tap]
            cfg.taper_range = float(tapering)
        devices = [int(v) for v in device] if isinstance(device, (list, tuple)) else None
        cfg.device = devices[0] if devices else int(device)
        cfg.nb_outer = int(nb_outer)
        cfg.flags = (0 if lookahead else _lib.VC_FLAG_NO_LOOKAHEAD) | (_lib.VC_FLAG_LEGACY_SYRK if legacy_syrk else 0) | (_lib.VC_FLAG_LEGACY_POTRF if legacy_potrf else 0)
        self._cfg = cfg
        if devices is not None and _dist_grid is None:
            # one replica per GPU; n2LL_batch spreads the stencil over them (optimParallel, SVC_mle.R:164-200)
            dv = (C.c_int32 * len(devices))(*devices)
            code = lib().vc_create_multi(C.byref(self._h), C.byref(cfg), n, self.d, self.p, self.q,
                                         ptr(self.locs), ptr(self.X), ptr(self.W), ptr(self.y), dv, len(devices))
        elif _dist_grid is None:
            code = lib().vc_create(C.byref(self._h), C.byref(cfg), n, self.d, self.p, self.q,
                                   ptr(self.locs), ptr(self.X), ptr(self.W), ptr(self.y))
        else:
            rank, world, pr, pc, uid = _dist_grid
            code = lib().vc_create_dist(C.byref(self._h), C.byref(cfg), n, self.d, self.p, self.q,
                                        ptr(self.locs), ptr(self.X), ptr(self.W), ptr(self.y),
                                        rank, world, pr, pc, uid)
        if code != _lib.VC_OK:
            self._h = C.c_void_p()
            check(code, None)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().vc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- the objective ----------------------------------------------------------------------
    def _split(self, x, profile, mean_est):
        x = np.asarray(x, dtype=np.float64).ravel()
        nth = 2 * self.q + 1
        if profile:
            if x.size < nth:
                raise ValueError("x must have at least 2q+1 entries")
            if mean_est is None:
                return x[:nth], _lib.VC_MU_GLS, None
            mu = np.asarray(mean_est, dtype=np.float64).ravel()
        else:
            if x.size < nth + self.p:
                raise ValueError("non-profile objective needs x of length 2q+1+p")
            mu = x[nth:nth + self.p]
        if mu.size != self.p:
            raise ValueError("mean vector must have length p")
        return x[:nth], _lib.VC_MU_FIXED, np.ascontiguousarray(mu)

    def n2LL(self, x, mean_est="__default__", profile=None, pc_prior="__default__", parts=False):
        """-2 log-likelihood at x (objective_functions.R:4-62), PC penalty added on the host."""
        profile = self.profileLik if profile is None else bool(profile)
        mean_est = self.mean_est if isinstance(mean_est, str) else mean_est
        pc_prior = self.pc_prior if isinstance(pc_prior, str) else pc_prior
        theta, mode, mu = self._split(x, profile, mean_est)
        theta = np.ascontiguousarray(theta)
        out = C.c_double()
        logdet = C.c_double()
        quad = C.c_double()
        mu_out = np.empty(self.p)
        code = lib().vc_n2ll(self._h, ptr(theta), mode, ptr(mu), C.byref(out), ptr(mu_out),
                             C.byref(logdet), C.byref(quad))
        check(code, self._h)
        val = out.value + pc_penalty(theta, self.q, pc_prior)
        if parts:
            return {"n2ll": val, "mu": mu_out, "logdet": logdet.value, "quad": quad.value}
        return val

    __call__ = n2LL

    def n2LL_batch(self, xs, mean_est="__default__", profile=None, pc_prior="__default__",
                   return_status=False):
        """Evaluate the objective at every row of xs with one synchronisation
        (optim's finite-difference stencil / optimhess / optimParallel, SVC_mle.R:145-199)."""
        profile = self.profileLik if profile is None else bool(profile)
        mean_est = self.mean_est if isinstance(mean_est, str) else mean_est
        pc_prior = self.pc_prior if isinstance(pc_prior, str) else pc_prior
        xs = np.atleast_2d(np.asarray(xs, dtype=np.float64))
        m = xs.shape[0]
        nth = 2 * self.q + 1
        thetas = np.empty((m, nth))
        mus = None
        mode = _lib.VC_MU_GLS
        for e in range(m):
            th, mode, mu = self._split(xs[e], profile, mean_est)
            thetas[e] = th
            if mu is not None:
                if mus is None:
                    mus = np.empty((m, self.p))
                mus[e] = mu
        out = np.empty(m)
        status = np.zeros(m, dtype=np.int32)
        code = lib().vc_n2ll_batch(self._h, m, ptr(thetas), mode, ptr(mus), ptr(out), None,
                                   status.ctypes.data_as(C.POINTER(C.c_int32)))
        if code not in (_lib.VC_OK, _lib.VC_ERR_NOT_PD, _lib.VC_ERR_INVALID):
            check(code, self._h)
        for e in range(m):
            if status[e] == _lib.VC_OK:
                out[e] += pc_penalty(thetas[e], self.q, pc_prior)
        if return_status:
            return out, status
        if code != _lib.VC_OK:
            check(code, self._h)
        return out

    # -- pieces of the path, for parity checks and the next-tier rows -------------------------
    def Sigma_y(self, x):
        """Sigma_y(x[1:(2q+1)]) as a dense symmetric n x n matrix (utils.R:171-233)."""
        theta = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel()[:2 * self.q + 1])
        out = np.empty((self.n, self.n), order="F")
        check(lib().vc_sigma(self._h, ptr(theta), ptr(out)), self._h)
        return out

    def Sigma_y_device(self, x, out_ptr):
        """Sigma_y written to device memory (n x n column-major doubles at the raw pointer out_ptr)."""
        theta = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel()[:2 * self.q + 1])
        check(lib().vc_sigma_device(self._h, ptr(theta), C.c_void_p(int(out_ptr))), self._h)

    def chol(self):
        """Upper factor R (R'R = Sigma_y) of the last evaluation -- base::chol's convention."""
        out = np.empty((self.n, self.n), order="F")
        check(lib().vc_get_chol(self._h, ptr(out)), self._h)
        return out

    def whitened(self):
        """solve(t(R), cbind(X, y)) of the last evaluation (n x (p+1))."""
        out = np.empty((self.n, self.p + 1), order="F")
        check(lib().vc_get_whitened(self._h, ptr(out)), self._h)
        return out

    def set_data(self, y=None, X=None, locs=None, W=None):
        if y is not None:
            self.y = as_f64(y).ravel()
        if X is not None:
            self.X = as_f64(X, 2)
        if W is not None:
            self.W = as_f64(W, 2)
        if locs is not None:
            self.locs = as_f64(locs, 2)
        check(lib().vc_set_data(self._h, ptr(self.locs) if locs is not None else None,
                                ptr(self.X) if X is not None else None,
                                ptr(self.W) if W is not None else None,
                                ptr(self.y) if y is not None else None), self._h)

    def timings(self):
        t = _lib.VcTiming()
        check(lib().vc_timings(self._h, C.byref(t)), self._h)
        return {"build_ms": t.build_ms, "chol_ms": t.chol_ms, "finalize_ms": t.finalize_ms,
                "total_ms": t.total_ms}

    @property
    def launch_count(self):
        return int(lib().vc_launch_count(self._h))

    def dist_traffic(self):
        """(bytes received over NVLink per evaluation by this rank, row/column-communicator bound); sharded handles."""
        a, b = C.c_double(), C.c_double()
        check(lib().vc_dist_traffic(self._h, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value

    @property
    def device_count(self):
        return int(lib().vc_device_count(self._h))

    @property
    def padded_n(self):
        return int(lib().vc_padded_n(self._h))

    def post(self, x, edof=True):
        """prep_par_output / eff_dof pieces (utils.R:423-473, eff_dof.R:7-21) from the GPU factor."""
        theta = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel()[:2 * self.q + 1])
        mu = np.empty(self.p)
        sfe = np.empty((self.p, self.p), order="F")
        ed = C.c_double(float("nan"))
        check(lib().vc_post(self._h, ptr(theta), ptr(mu), ptr(sfe), C.byref(ed) if edof else None), self._h)
        return {"mu_gls": mu, "Sigma_FE": sfe, "edof": ed.value}

    def predict(self, x, mu, newlocs=None, newX=None, newW=None, compute_y_var=False):
        """predict.SVC_mle (predict-SVC_mle.R:80-267) on the GPU."""
        theta = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel()[:2 * self.q + 1])
        mu = np.ascontiguousarray(np.asarray(mu, dtype=np.float64).ravel())
        newlocs = self.locs if newlocs is None else as_f64(newlocs, 2)
        n_new = newlocs.shape[0]
        eff = np.empty((n_new, self.q), order="F")
        have_xw = newX is not None and newW is not None
        nX = as_f64(newX, 2) if have_xw else None
        nW = as_f64(newW, 2) if have_xw else None
        y_pred = np.empty(n_new) if have_xw else None
        y_var = np.empty(n_new) if (have_xw and compute_y_var) else None
        check(lib().vc_predict(self._h, ptr(theta), ptr(mu), n_new, ptr(newlocs), ptr(nX), ptr(nW),
                               ptr(eff), ptr(y_pred), ptr(y_var)), self._h)
        out = {"eff": eff}
        if have_xw:
            out["y_pred"] = y_pred
            if compute_y_var:
                out["y_var"] = y_var
        return out


def n2LL(x, obj: SVCObjective, mean_est=None, pc_prior=None, profile=True):
    """Free-function form with the reference's argument meaning (objective_functions.R:4-11):
    the heavy closure arguments (cov_func, outer.W, y, X, W, taper, Rstruct) live in `obj`."""
    return obj.n2LL(x, mean_est=mean_est, profile=profile, pc_prior=pc_prior)


def GLS_chol(R, X, y, device=0):
    """GLS_chol.matrix (utils.R:93-101): mu_GLS from an upper Cholesky factor R of Sigma."""
    R = as_f64(R, 2)
    X = as_f64(X, 2)
    y = as_f64(y).ravel()
    n, p = X.shape
    if R.shape != (n, n) or y.shape[0] != n:
        raise ValueError("non-conformable arguments")
    mu = np.empty(p)
    check(lib().vc_gls_chol(int(device), n, p, ptr(R), ptr(X), ptr(y), ptr(mu)), None)
    return mu


def check_cov_lower(cv, q):
    """utils.R:291-297."""
    cv = np.asarray(cv, dtype=np.float64)
    return bool(np.all(cv[0:2 * q + 1:2] > 0) and np.all(cv[1:2 * q:2] >= 0))
