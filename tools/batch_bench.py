"""Throughput of vc_n2ll_batch at small / mid n (the sizes real varycoef fits live at): batch of m evaluations
vs one-at-a-time, per-evaluation time and fraction of the fp64 DMMA peak.  CUDA-event-free wall clock around
synchronous ABI calls (every call ends with a stream synchronise), best of `reps`.

    python tools/batch_bench.py [--sizes 1000,2000,5000,10000] [--m 16] [--reps 5] [--out file.json]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varycoef_b200 import SVCObjective, _lib  # noqa: E402


def synth(n, p, seed):
    rng = np.random.default_rng(seed)
    L = 10.0 * np.sqrt(n / 1000.0)
    locs = rng.uniform(0, L, size=(n, 2))
    X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
    y = X @ np.arange(1, p + 1) + rng.standard_normal(n) * 1.5
    return locs, X, y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1000,2000,5000,10000")
    ap.add_argument("--m", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cov", default="exp")
    ap.add_argument("--out", default="")
    ap.add_argument("--nb", type=int, default=0, help="outer Cholesky block (0 = auto)")
    a = ap.parse_args()
    peak = C.c_double(0.0)
    _lib.lib().vc_probe_dmma_peak(0, 20000, C.byref(peak))
    rows = []
    for n in [int(v) for v in a.sizes.split(",")]:
        p = 3
        locs, X, y = synth(n, p, 5 + n)
        base = np.array([1.3, 0.7, 2.1, 0.4, 0.9, 1.1, 0.35])
        xs = np.vstack([base * (1 + 0.003 * k) for k in range(a.m)])
        with SVCObjective(y, X, locs, cov_name=a.cov, profileLik=True, nb_outer=a.nb) as obj:
            single = np.array([obj.n2LL(x) for x in xs])
            batch = obj.n2LL_batch(xs)
            t_single, t_batch = 1e30, 1e30
            for _ in range(a.reps):
                t0 = time.perf_counter()
                for x in xs[:4]:
                    obj.n2LL(x)
                t_single = min(t_single, (time.perf_counter() - t0) / 4)
                t0 = time.perf_counter()
                obj.n2LL_batch(xs)
                t_batch = min(t_batch, (time.perf_counter() - t0) / a.m)
            fl = n ** 3 / 3.0
            rows.append({"n": n, "m": a.m, "nb": a.nb, "single_ms": t_single * 1e3, "batch_ms_per_eval": t_batch * 1e3,
                         "single_tflops": fl / t_single / 1e12, "batch_tflops": fl / t_batch / 1e12,
                         "batch_frac_of_peak": fl / t_batch / 1e12 / peak.value, "speedup": t_single / t_batch,
                         "bit_identical": bool(np.array_equal(single, batch))})
            print(json.dumps(rows[-1]), flush=True)
    out = {"dmma_peak_tflops": peak.value, "rows": rows}
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
