"""Bring-up diagnostics on a real B200: every stage of the path against the numpy oracle."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import svc_oracle as o  # noqa: E402
from varycoef_b200 import SVCObjective, _lib  # noqa: E402
import ctypes as C  # noqa: E402


def synth(n, p, seed, dim=2):
    rng = np.random.default_rng(seed)
    L = 10.0 * np.sqrt(n / 1000.0)
    locs = rng.uniform(0, L, size=(n, dim))
    X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
    y = X @ np.arange(1, p + 1) + rng.standard_normal(n) * 1.5
    return locs, X, y


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def case(n, p, cov, lookahead=True, nb=0, dim=2, taper=None):
    locs, X, y = synth(n, p, 7 + n, dim)
    q = p
    theta = np.array([1.3, 0.7, 2.1, 0.4, 0.9, 1.1, 1.7, 0.3, 2.5, 0.2][:2 * q] + [0.35])
    t0 = time.time()
    obj = SVCObjective(y, X, locs, cov_name=cov, profileLik=True, lookahead=lookahead, nb_outer=nb,
                       tapering=taper)
    out = {"n": n, "p": p, "cov": cov, "lookahead": lookahead, "nb": nb, "taper": taper}
    r = obj.n2LL(theta, parts=True)
    out["gpu_first_s"] = time.time() - t0
    if taper is None:
        D = o.own_dist(locs)
        mask = None
    else:
        D, mask = o.own_dist(locs, taper=taper)
    ref = o.n2ll(theta, D, X, X, y, cov, taper_range=taper, mask=mask, profile=True, faithful=False, parts=True)
    out["n2ll_gpu"], out["n2ll_ref"] = r["n2ll"], ref["n2ll"]
    out["n2ll_rel"] = abs(r["n2ll"] - ref["n2ll"]) / abs(ref["n2ll"])
    out["logdet_rel"] = abs(r["logdet"] - ref["logdet"]) / abs(ref["logdet"])
    out["quad_rel"] = abs(r["quad"] - ref["quad"]) / abs(ref["quad"])
    out["mu_rel"] = rel(r["mu"], ref["mu"])
    if n <= 3000:
        S = obj.Sigma_y(theta)
        Sref = o.sigma_y(theta, D, X, cov, taper, mask)
        out["sigma_rel"] = rel(S, Sref)
        out["sigma_sym"] = float(np.max(np.abs(S - S.T)))
        obj.n2LL(theta)
        R = obj.chol()
        Rref = o.chol_upper(Sref)
        out["chol_rel"] = rel(R, Rref)
        Z = obj.whitened()
        from scipy.linalg import solve_triangular
        Zref = solve_triangular(Rref, np.column_stack([X, y]), trans=1, lower=False)
        out["z_rel"] = rel(Z, Zref)
    # fixed-mu mode
    mu = ref["mu"] * 1.01
    r2 = obj.n2LL(np.concatenate([theta, mu]), profile=False)
    ref2 = o.n2ll(np.concatenate([theta, mu]), D, X, X, y, cov, taper_range=taper, mask=mask, profile=False, faithful=False)
    out["fixed_rel"] = abs(r2 - ref2) / abs(ref2)
    t0 = time.time()
    for _ in range(3):
        obj.n2LL(theta)
    out["eval_s"] = (time.time() - t0) / 3
    out["timing"] = obj.timings()
    out["launches"] = obj.launch_count
    obj.close()
    return out


def main():
    res = []
    L = _lib.lib()
    tf = C.c_double()
    print("dmma probe rc", L.vc_probe_dmma_peak(0, 20000, C.byref(tf)), tf.value, "TFLOP/s", flush=True)
    gb = C.c_double()
    print("hbm write rc", L.vc_probe_hbm_write(0, 8 << 30, 5, C.byref(gb)), gb.value, "GB/s", flush=True)
    cases = [(100, 2, "exp", True, 0, 1), (300, 2, "exp", False, 0, 2), (700, 3, "mat32", True, 0, 2),
             (1500, 3, "exp", False, 256, 2), (1500, 3, "exp", True, 256, 2), (2500, 3, "mat52", True, 0, 3),
             (3000, 5, "sph", True, 0, 2), (6000, 3, "exp", True, 0, 2), (6000, 3, "exp", False, 0, 2)]
    for (n, p, cov, la, nb, dim) in cases:
        try:
            r = case(n, p, cov, la, nb, dim)
        except Exception as e:  # keep going: we want the whole picture from one GPU call
            r = {"n": n, "cov": cov, "error": repr(e)}
        print(json.dumps(r), flush=True)
        res.append(r)
    try:
        r = case(2000, 3, "exp", True, 0, 2, taper=0.6)
    except Exception as e:
        r = {"taper": True, "error": repr(e)}
    print(json.dumps(r), flush=True)
    res.append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"dmma_tflops": tf.value, "hbm_write_gbs": gb.value, "cases": res},
              open("gpurun_out/diag.json", "w"), indent=1)
    # timing at a larger size (no oracle)
    for n in (10000, 20000):
        locs, X, y = synth(n, 3, 1)
        obj = SVCObjective(y, X, locs, profileLik=True)
        theta = np.array([1.3, 0.7, 2.1, 0.4, 0.9, 1.1, 0.35])
        v = obj.n2LL(theta)
        t0 = time.time()
        v = obj.n2LL(theta)
        dt = time.time() - t0
        tm = obj.timings()
        print("n=%d n2ll=%.10g wall=%.3fs timing=%s chol_tflops=%.2f" % (
            n, v, dt, tm, (obj.padded_n ** 3 / 3) / (tm["chol_ms"] * 1e-3) / 1e12), flush=True)
        obj.close()


if __name__ == "__main__":
    main()
