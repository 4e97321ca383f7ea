"""One warm-up + `reps` timed -2logL evaluations at size n (for ncu captures and quick timing)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varycoef_b200 import SVCObjective  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
la = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
legacy = (sys.argv[5] == "1") if len(sys.argv) > 5 else False
legacy_potrf = (sys.argv[6] == "1") if len(sys.argv) > 6 else False
rng = np.random.default_rng(1)
L = 10.0 * np.sqrt(n / 1000.0)
locs = rng.uniform(0, L, (n, 2))
X = np.column_stack([np.ones(n), rng.standard_normal((n, 2))])
y = X @ np.array([1.0, 2.0, 0.5]) + rng.standard_normal(n) * 1.8
theta = np.array([1.3, 0.7, 2.1, 0.4, 0.9, 1.1, 0.35])
obj = SVCObjective(y, X, locs, profileLik=True, nb_outer=nb, lookahead=la, legacy_syrk=legacy, legacy_potrf=legacy_potrf)
v = obj.n2LL(theta)
for rep in range(reps):
    th = theta * (1.0 + 1e-3 * (rep + 1))      # a new theta every time: no factor-cache hit
    t0 = time.time()
    v = obj.n2LL(th)
    dt = time.time() - t0
    tm = obj.timings()
    print("legacy=%d lp=%d " % (legacy, legacy_potrf), end="")
    print("n=%d nb=%d la=%d n2ll=%.12g wall=%.4fs build=%.3fms chol=%.3fms fin=%.3fms chol_tflops=%.2f (n^3/3) build_gbs=%.0f" % (
        n, nb, la, v, dt, tm["build_ms"], tm["chol_ms"], tm["finalize_ms"],
        n ** 3 / 3 / (tm["chol_ms"] * 1e-3) / 1e12, 4.0 * n * (n + 1) / (tm["build_ms"] * 1e-3) / 1e9), flush=True)
obj.close()
