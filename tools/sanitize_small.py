"""Small end-to-end walk over every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize_small.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import synth, theta_for  # noqa: E402
from varycoef_b200 import SVCObjective, GLS_chol  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 700
locs, X, W, y = synth(n, 3, 99, 2)
theta = theta_for(3)
vals = {}
for kw in ({}, {"cov_name": "mat32", "lookahead": False}, {"tapering": 1.2}, {"legacy_syrk": True, "legacy_potrf": True},
           {"dist": {"method": "manhattan"}, "cov_name": "sph"}):
    with SVCObjective(y, X, locs, profileLik=True, **kw) as obj:
        v = obj.n2LL(theta)
        b = obj.n2LL_batch(np.stack([theta * (1 + 0.01 * i) for i in range(6)]))
        assert abs(b[0] - v) <= 1e-12 * abs(v)
        post = obj.post(theta)
        pr = obj.predict(theta, post["mu_gls"], locs[:50] + 0.01, X[:50], W[:50], compute_y_var=True)
        R = obj.chol() if not kw else None
        Z = obj.whitened()
        vals[str(kw)] = v
    if R is not None:
        mu = GLS_chol(R, X, y)
        assert np.allclose(mu, post["mu_gls"], rtol=1e-9)
print("sanitize_small OK", vals)
