"""Timing of the next-tier rows on the resident factor at scale:
vc_post (prep_par_output + eff_dof) and vc_predict (EBLUP, y_pred, y_var), n = 20 000 and 50 000.

    python tools/post_predict_bench.py [--sizes 20000,50000] [--out file.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varycoef_b200 import SVCObjective  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="20000,50000")
    ap.add_argument("--n-new", type=int, default=4096)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rows = []
    for n in [int(v) for v in a.sizes.split(",")]:
        p = 3
        rng = np.random.default_rng(11 + n)
        L = 10.0 * np.sqrt(n / 1000.0)
        locs = rng.uniform(0, L, (n, 2))
        X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
        y = X @ np.arange(1, p + 1) + rng.standard_normal(n) * 1.5
        theta = np.array([1.3, 0.7, 2.1, 0.4, 0.9, 1.1, 0.35])
        newlocs = rng.uniform(0, L, (a.n_new, 2))
        newX = np.column_stack([np.ones(a.n_new), rng.standard_normal((a.n_new, p - 1))])
        with SVCObjective(y, X, locs, profileLik=True) as obj:
            t0 = time.perf_counter(); obj.n2LL(theta); t_eval = time.perf_counter() - t0
            t0 = time.perf_counter(); post = obj.post(theta, edof=False); t_post = time.perf_counter() - t0
            t0 = time.perf_counter(); post = obj.post(theta, edof=True); t_edof = time.perf_counter() - t0
            t0 = time.perf_counter(); pr = obj.predict(theta, post["mu_gls"], newlocs, newX, newX); t_pred = time.perf_counter() - t0
            t0 = time.perf_counter(); pr = obj.predict(theta, post["mu_gls"], newlocs, newX, newX, compute_y_var=True); t_var = time.perf_counter() - t0
        rows.append({"n": n, "n_new": a.n_new, "eval_s": t_eval, "post_cached_s": t_post, "post_with_edof_s": t_edof,
                     "predict_s": t_pred, "predict_with_y_var_s": t_var, "edof": post["edof"]})
        print(json.dumps(rows[-1]), flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
