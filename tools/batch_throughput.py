"""Batch (multi-lane) vs one-at-a-time throughput of the objective at small / mid n."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varycoef_b200 import SVCObjective
for n in [int(a) for a in sys.argv[1:]] or [1000, 2000, 5000, 10000, 16000]:
    rng = np.random.default_rng(1)
    L = 10.0 * np.sqrt(n / 1000.0)
    locs = rng.uniform(0, L, (n, 2))
    X = np.column_stack([np.ones(n), rng.standard_normal((n, 2))])
    y = X @ np.array([1.0, 2.0, 0.5]) + rng.standard_normal(n) * 1.8
    theta = np.array([1.3, 0.7, 2.1, 0.4, 0.9, 1.1, 0.35])
    xs = np.vstack([theta * (1 + 1e-3 * k) for k in range(16)])
    with SVCObjective(y, X, locs, profileLik=True) as obj:
        obj.n2LL_batch(xs)
        t0 = time.time(); single = np.array([obj.n2LL(x * (1 + 1e-7)) for x in xs]); t1 = time.time() - t0
        t0 = time.time(); batch = obj.n2LL_batch(xs); t2 = time.time() - t0
        chk = np.array([obj.n2LL(x) for x in xs])
    print("n=%6d  one-at-a-time %.2f ms/eval   batch of 16 %.2f ms/eval   speed-up %.2fx   batch==single: %s" % (
        n, 1e3 * t1 / 16, 1e3 * t2 / 16, t1 / t2, np.array_equal(batch, chk)), flush=True)
