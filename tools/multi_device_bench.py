"""theta fan-out behind the ABI (vc_create_multi): optim's stencil and a full fit on 1 device vs all devices.

    python tools/multi_device_bench.py [--n 10000] [--cov mat32] [--out file.json]

BASELINE config 2 (n = 10 000, p = q = 3, Matern-3/2): the 15-point central-difference stencil of the profile
objective (npar = 7) and the 196-point Hessian stencil through SVCObjective.n2LL_batch, then SVC_mle(...) with
control["parallel"] = device list -- the analogue of the reference's optimParallel cluster (R-pkg/R/SVC_mle.R:164-200).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from varycoef_b200 import SVCObjective  # noqa: E402
from varycoef_b200.optim import fd_points  # noqa: E402
from varycoef_b200.svc_mle import SVC_mle, SVC_mle_control  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10000)
    ap.add_argument("--cov", default="mat32")
    ap.add_argument("--fit-n", type=int, default=4000)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    ng = torch.cuda.device_count()
    n, p = a.n, 3
    rng = np.random.default_rng(20261019 + 2)
    L = 10.0 * np.sqrt(n / 1000.0)
    locs = rng.uniform(0, L, (n, 2))
    X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
    y = (1.0 + np.sin(locs[:, 0] / 2.0)) + X[:, 1] * (2.0 + np.cos(locs[:, 1] / 3.0)) + 0.5 * X[:, 2] + 0.5 * rng.standard_normal(n)
    theta = np.array([1.5, 0.6, 2.0, 0.5, 1.0, 0.3, 0.3])
    stencil = np.vstack([theta] + [theta + s * 1e-3 * abs(theta[i]) * np.eye(7)[i] for i in range(7) for s in (1, -1)])
    hess = np.vstack([theta + 1e-3 * (si * abs(theta[i]) * np.eye(7)[i] + sj * abs(theta[j]) * np.eye(7)[j])
                      for i in range(7) for j in range(7) for si in (1, -1) for sj in (1, -1)])
    out = {"n": n, "cov": a.cov, "gpus": ng, "rows": []}
    ref = None
    for devs in ([0], list(range(ng))) if ng > 1 else ([0], [0, 0]):
        with SVCObjective(y, X, locs, cov_name=a.cov, profileLik=True, device=devs if len(devs) > 1 else devs[0]) as obj:
            obj.n2LL_batch(stencil)
            t0 = time.perf_counter(); v = obj.n2LL_batch(stencil); t_st = time.perf_counter() - t0
            obj.n2LL_batch(hess)
            t0 = time.perf_counter(); vh = obj.n2LL_batch(hess); t_h = time.perf_counter() - t0
        if ref is None:
            ref = (v, vh)
        row = {"devices": devs, "stencil15_ms": t_st * 1e3, "hessian196_ms": t_h * 1e3,
               "bit_identical_to_one_device": bool(np.array_equal(v, ref[0]) and np.array_equal(vh, ref[1]))}
        out["rows"].append(row)
        print(json.dumps(row), flush=True)
    # a full fit (smaller n so that it finishes in seconds)
    nf = a.fit_n
    fits = []
    for devs in ([0], list(range(ng))) if ng > 1 else ([0], [0, 0]):
        ctl = SVC_mle_control(cov_name=a.cov, profileLik=True, hessian=False, parallel=devs if len(devs) > 1 else None)
        t0 = time.perf_counter()
        f = SVC_mle(y[:nf], X[:nf], locs[:nf], control=ctl, compute_edof=False)
        fits.append({"devices": devs, "fit_s": time.perf_counter() - t0, "cov_par": [float(v) for v in f.cov_par],
                     "counts": int(f.optim_output["counts"]["function"])})
        print(json.dumps(fits[-1]), flush=True)
    out["fit_n"] = nf
    out["fits"] = fits
    a0, a1 = np.array(fits[0]["cov_par"]), np.array(fits[1]["cov_par"])
    out["fit_max_rel_diff"] = float(np.max(np.abs(a0 - a1) / np.maximum(np.abs(a0), 1e-300)))
    print("fit max rel diff", out["fit_max_rel_diff"])
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
