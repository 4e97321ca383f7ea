"""Host logic of the SVC_selection mirror; no GPU needed."""
import numpy as np
import pytest

from varycoef_b200 import svc_selection as sel
from oracle import svc_selection_oracle as selo


def _kkt(x, y, beta, w):
    n = x.shape[0]
    g = x.T @ (y - x @ beta) / n
    for j in range(x.shape[1]):
        if beta[j] != 0.0:
            assert abs(g[j] - w[j] * np.sign(beta[j])) < 1e-12
        else:
            assert abs(g[j]) <= w[j] + 1e-12


@pytest.mark.parametrize("lam", [1e-10, 0.05, 0.4, 10.0])
def test_lasso_matches_independent_solver_and_kkt(lam):
    rng = np.random.default_rng(5)
    n, p = 400, 5
    x = rng.normal(size=(n, p)) * np.array([1.0, 3.0, 0.2, 1.0, 1.0]) + np.array([0.0, 1.0, 0.0, -2.0, 0.5])
    y = x @ np.array([1.0, 0.0, -2.0, 0.3, 0.0]) + rng.normal(size=n)
    b = sel.lasso_glmnet(x, y, lam)
    b2 = selo.lasso_glmnet_sklearn(x, y, lam)
    assert np.allclose(b, b2, rtol=1e-9, atol=1e-11)
    xs = np.sqrt((x ** 2).mean(0) - x.mean(0) ** 2)
    _kkt(x, y, b, lam * xs)
    if lam == 10.0:
        assert np.all(b == 0.0)


def test_lasso_adaptive_weights_and_excluded_variable():
    rng = np.random.default_rng(6)
    n, p = 300, 4
    x = rng.normal(size=(n, p))
    y = x @ np.array([2.0, -1.0, 0.5, 0.0]) + 0.3 * rng.normal(size=n)
    mle = np.array([2.0, -1.0, 0.5, 0.0])
    with np.errstate(divide="ignore"):
        pf = 1.0 / np.abs(mle)                       # last one infinite -> excluded (glmnet)
    b = sel.lasso_glmnet(x, y, 0.05, pf)
    b2 = selo.lasso_glmnet_sklearn(x, y, 0.05, pf)
    assert b[3] == 0.0 and np.allclose(b, b2, rtol=1e-9, atol=1e-11)
    pf1 = pf.copy(); pf1[3] = 1.0
    vq = pf1 * p / pf1.sum()
    xs = np.sqrt((x ** 2).mean(0) - x.mean(0) ** 2)
    w = 0.05 * vq * xs
    w[3] = np.inf
    _kkt(x, y, b, w)


def test_constant_column_is_excluded():
    rng = np.random.default_rng(7)
    x = np.column_stack([np.ones(50), rng.normal(size=50)])
    y = 3.0 + x[:, 1] + 0.1 * rng.normal(size=50)
    b = sel.lasso_glmnet(x, y, 0.01)
    assert b[0] == 0.0 and b[1] != 0.0


def test_control_defaults_and_errors():
    c = sel.SVC_selection_control()
    assert c["IC.type"] == "BIC" and c["method"] == "grid" and c["r.lambda"] == (1e-10, 10.0)
    assert c["n.lambda"] == 10 and c["CD.conv"] == {"N": 20, "delta": 1e-6, "logLik": True}
    assert c["adaptive"] is False and c["hessian"] is False and c["optim.args"] == {}
    with pytest.raises(ValueError):
        sel.SVC_selection_control(IC_type="AIC")
    with pytest.raises(ValueError):
        sel.SVC_selection_control(method="random")
    # reference test: expect_error(SVC_selection(1, 1))  (tests/testthat/test_SVC_selection.R:3)
    with pytest.raises(Exception):
        sel.SVC_selection(1, 1)


def test_grid_order_is_expand_grid():
    seen = []
    out = sel.IC_opt_grid(lambda lam: (seen.append(tuple(lam)), lam[0] + 10 * lam[1])[1], (1e-2, 1.0), 3)
    ll = np.exp(np.linspace(np.log(1e-2), 0.0, 3))
    assert np.allclose([s[0] for s in seen[:3]], ll) and np.allclose([s[1] for s in seen[:3]], ll[0])
    assert out["IC_grid"].shape == (9,) and int(np.argmin(out["IC_grid"])) == 0
