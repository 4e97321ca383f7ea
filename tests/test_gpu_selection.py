"""GPU parity of the SVC_selection mirror: CD_mu / CD_theta / PMLE_CD on the GPU
objective against the CPU restatement (oracle/svc_selection_oracle.py).  The reference's own test-suite pins no
value on this path (parity unpinned); tolerances are the north-star ones."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_PAR = 1e-6


def _model(n=400, seed=21):
    from oracle import svc_oracle as so
    rng = np.random.default_rng(seed)
    locs = rng.uniform(0, 10, (n, 2))
    X = np.column_stack([np.ones(n), rng.standard_normal(n), rng.standard_normal(n)])
    d = so.sample_svcdata(rng, locs, [2.0, 1.0], [1.5, 1.0], [1.0, 2.0, 0.0], 0.5, "exp", X=X[:, :2])
    y = d["y"]                                   # third regressor has a zero coefficient
    W = X[:, :2]
    return so, locs, X, W, y


def test_cd_mu_whitening_and_lasso_match_oracle():
    from varycoef_b200 import SVCObjective
    from varycoef_b200 import svc_selection as sel
    from oracle import svc_selection_oracle as selo
    so, locs, X, W, y = _model()
    theta = np.array([1.5, 2.0, 1.0, 1.0, 0.25])
    model = dict(D=so.own_dist(locs), X=X, W=W, y=y, cov_name="exp")
    with SVCObjective(y, X, locs, W, profileLik=True) as obj:
        for lam, adaptive in [(1e-8, False), (0.05, False), (0.05, True), (1.0, False)]:
            mle_mu = obj.post(theta, edof=False)["mu_gls"]
            got = sel.CD_mu(theta, mle_mu, obj, lam, adaptive)
            ref = selo.cd_mu(theta, mle_mu, model, lam, adaptive)
            print("CD_mu lambda=%g adaptive=%s: %s vs %s" % (lam, adaptive, got, ref))
            assert np.allclose(got, ref, rtol=1e-8, atol=1e-10)
            assert np.array_equal(got == 0.0, ref == 0.0)


@pytest.mark.parametrize("lam,adaptive,ic", [((0.02, 0.002), False, "BIC"), ((0.02, 0.01), True, "cAIC_BW"),
                                              ((0.08, 0.0005), False, "cAIC_VB")])
def test_pmle_cd_matches_oracle(lam, adaptive, ic):
    from varycoef_b200 import SVCObjective
    from varycoef_b200 import svc_selection as sel
    from varycoef_b200.optim import optim_lbfgsb
    from oracle import svc_selection_oracle as selo
    so, locs, X, W, y = _model()
    q = 2
    mle = np.array([1.4, 1.9, 0.9, 1.1, 0.26])
    oa = {"lower": np.array([1e-3, 0.0] * q + [1e-3]), "upper": np.array([30.0, 20.0] * q + [20.0]),
          "control": {"parscale": np.abs(mle)}}
    conv = {"N": 3, "delta": 1e-9, "logLik": True}
    model = dict(D=so.own_dist(locs), X=X, W=W, y=y, cov_name="exp")
    with SVCObjective(y, X, locs, W, profileLik=True) as obj:
        got = sel.PMLE_CD(lam, mle, obj, oa, adaptive, True, ic, conv)
    ref = selo.pmle_cd(lam, mle, model, optim_lbfgsb, oa, adaptive, ic, conv)
    k = got["n.iter"]
    assert k == ref["n.iter"]
    cp, cr = got["c.par"][k], ref["c.par"][k]
    mp, mr = got["mu.par"][k], ref["mu.par"][k]
    scale = np.maximum(np.abs(cr), 1e-3)
    rel = float(np.max(np.abs(cp - cr) / scale))
    relm = float(np.max(np.abs(mp - mr) / np.maximum(np.abs(mr), 1e-3)))
    reli = abs(got["final.IC"] - ref["final.IC"]) / abs(ref["final.IC"])
    print("PMLE_CD lam=%s adaptive=%s %s: k=%d theta %s mu %s IC %.6f; rel diff theta %.2e mu %.2e IC %.2e" % (
        lam, adaptive, ic, k, np.round(cp, 5), np.round(mp, 5), got["final.IC"], rel, relm, reli))
    assert np.allclose(got["mu.par"][0], ref["mu.par"][0], rtol=1e-10)
    assert abs(got["loglik.CD"][0] - ref["loglik.CD"][0]) <= 1e-10 * abs(ref["loglik.CD"][0])
    assert rel <= TOL_PAR and relm <= TOL_PAR, (rel, relm)
    assert reli <= 5e-9
    assert np.array_equal(mp == 0.0, mr == 0.0)


def test_selection_grid_runs_and_selects_minimum():
    from varycoef_b200 import SVCObjective
    from varycoef_b200 import svc_selection as sel
    so, locs, X, W, y = _model(n=300, seed=3)
    q = 2
    mle = np.array([1.4, 1.9, 0.9, 1.1, 0.26])
    ctl = sel.SVC_selection_control(r_lambda=(1e-3, 0.5), n_lambda=2,
                                    CD_conv={"N": 2, "delta": 1e-6, "logLik": True},
                                    optim_args={"lower": np.array([1e-3, 0.0] * q + [1e-3]),
                                                "upper": np.array([30.0, 20.0] * q + [20.0]),
                                                "control": {"parscale": np.abs(mle), "maxit": 30}})
    with SVCObjective(y, X, locs, W, profileLik=True) as obj:
        out = sel.SVC_selection(obj, mle, ctl)
        assert out["PMLE_opt"]["IC_grid"].shape == (4,)
        j = int(np.argmin(out["PMLE_opt"]["IC_grid"]))
        assert np.allclose(out["lambda"], out["PMLE_opt"]["lambda_grid"][j])
        assert abs(out["PMLE_pars"]["final.IC"] - out["PMLE_opt"]["IC_grid"][j]) <= 1e-8 * abs(out["PMLE_pars"]["final.IC"])
        # heavy shrinkage removes the zero-mean regressor
        heavy = sel.PMLE_CD((0.5, 1e-3), mle, obj, ctl["optim.args"], False, True, "BIC", ctl["CD.conv"])
        assert heavy["mu.par"][heavy["n.iter"]][2] == 0.0
    with SVCObjective(y, X, locs, W, profileLik=False) as obj2:
        with pytest.raises(ValueError):
            sel.PMLE_CD((0.1, 0.1), mle, obj2)
        with pytest.raises(ValueError, match="profileLik"):
            sel.SVC_selection(obj2, mle, ctl)
