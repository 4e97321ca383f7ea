import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varycoef_b200 import SVCObjective

def run(n, la, legacy, nb=0):
    rng = np.random.default_rng(1)
    L = 10.0 * np.sqrt(n / 1000.0)
    locs = rng.uniform(0, L, (n, 2))
    X = np.column_stack([np.ones(n), rng.standard_normal((n, 2))])
    y = X @ np.array([1.0, 2.0, 0.5]) + rng.standard_normal(n) * 1.8
    theta = np.array([1.3, 0.7, 2.1, 0.4, 0.9, 1.1, 0.35])
    obj = SVCObjective(y, X, locs, profileLik=True, lookahead=la, legacy_syrk=legacy, nb_outer=nb)
    v = obj.n2LL(theta)
    R = obj.chol()
    Z = obj.whitened()
    obj.close()
    return v, R, Z

for n in [int(a) for a in sys.argv[1:]] or [4100]:
    v0, R0, Z0 = run(n, False, True)
    for la in (False, True):
        v1, R1, Z1 = run(n, la, False)
        bad = np.abs(R1 - R0) > 1e-9 * np.abs(R0).max()
        nt = (n + 127) // 128
        tiles = set()
        ii, jj = np.nonzero(bad)
        for a, b in zip(ii // 128, jj // 128):
            tiles.add((int(b), int(a)))    # R is upper: R[j, i] = L[i, j] -> tile (I=col tile, J=row tile)
        zb = np.nonzero(np.abs(Z1 - Z0).max(axis=1) > 1e-9)[0]
        print("n=%d nt=%d la=%d  v_legacy=%.10g v_ws=%.10g  bad L tiles=%d  bad Z row-tiles=%s" % (
            n, nt, la, v0, v1, len(tiles), sorted(set((zb // 128).tolist()))[:20]), flush=True)
        st = sorted(tiles)
        print("   first bad tiles (I,J):", st[:24], flush=True)
        if st:
            I, J = st[0]
            blk = (R1 - R0)[J * 128:(J + 1) * 128, I * 128:(I + 1) * 128].T   # L tile (rows I, cols J)
            r, c = np.nonzero(np.abs(blk) > 1e-9 * np.abs(R0).max())
            print("   in first bad tile: rows", sorted(set(r.tolist()))[:40], "cols", sorted(set(c.tolist()))[:40], flush=True)
