"""VC_TRACE timeline of one batched call: python tools/trace_batch.py n m out.txt  (run with VC_TRACE=out.txt)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varycoef_b200 import SVCObjective  # noqa: E402

n, m = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(5 + n)
L = 10.0 * np.sqrt(n / 1000.0)
locs = rng.uniform(0, L, size=(n, 2))
X = np.column_stack([np.ones(n), rng.standard_normal((n, 2))])
y = X @ np.arange(1, 4) + rng.standard_normal(n) * 1.5
base = np.array([1.3, 0.7, 2.1, 0.4, 0.9, 1.1, 0.35])
xs = np.vstack([base * (1 + 0.003 * k) for k in range(m)])
with SVCObjective(y, X, locs, profileLik=True) as obj:
    obj.n2LL_batch(xs)
    t0 = time.perf_counter()
    obj.n2LL_batch(xs)
    print("batch wall ms", (time.perf_counter() - t0) * 1e3)
