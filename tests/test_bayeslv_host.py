"""BayesLV: the host-side model of the log-variances (api.LogVarModel, functions.jl:446-485) against the statement-by-statement
restatement, with the single-site loop of the restatement standing in for the device sweep (CPU test; the device leg is test_gpu_bayeslv.py)."""
import numpy as np
import pytest

import nextgp.jl_b200 as ngp
from oracle import restate_numpy as RN


def _model_and_twin(p, k, est, seed):
    rng = np.random.default_rng(seed)
    cov = np.column_stack([np.ones(p)] + [rng.normal(size=p) for _ in range(k - 1)])
    prior = ngp.BayesLV(0.02, None, cov, 0.3, estimateVarZeta=est)
    m = ngp.LogVarModel(prior, p, np.random.default_rng(seed + 1))
    twin = dict(logVar=m.logVar.copy(), SNPVARRESID=m.SNPVARRESID.copy(), covariates=cov.copy(), iCpC=m.iCpC.copy(), c=m.c.copy(),
                varZeta=[m.varZeta], estVarZeta=est)
    return m, twin, rng


@pytest.mark.parametrize("est", [False, True, 0.01])
def test_log_variance_model_matches_the_restatement(est):
    n, p, k = 60, 150, 3
    m, twin, rng = _model_and_twin(p, k, est, 11)
    X = rng.normal(size=(n, p)); X -= X.mean(axis=0)
    mpm = (X * X).sum(axis=0)
    y = X[:, :10] @ rng.normal(size=10) + rng.normal(size=n)
    beta_o, e_o, vb_o = np.zeros(p), y - y.mean(), np.full(p, 0.02)
    vb = vb_o.copy()
    zero = np.zeros(p)
    for it in range(6):
        z, u, zc = rng.normal(size=p), rng.random((p, 4)), rng.normal(size=k)
        trapped = RN.bayes_lv(X, mpm, zero, zero, beta_o, e_o, 1.1, vb_o, twin, z, u, zc)
        m.update(beta_o, vb, u=u, z=zc)                      # same effects: the device sweep is compared in test_gpu_bayeslv.py
        assert m.trapped == trapped
        assert np.allclose(vb, vb_o, rtol=1e-12, atol=0) and np.allclose(m.logVar, twin["logVar"], rtol=1e-12, atol=1e-14)
        assert np.allclose(m.c, twin["c"], rtol=1e-10, atol=1e-13) and np.allclose(m.SNPVARRESID, twin["SNPVARRESID"], rtol=1e-9, atol=1e-12)
        assert abs(m.varZeta / twin["varZeta"][0] - 1) < 1e-10


def test_design_matrix_from_formula_and_columns():
    cols = {"maf": np.array([0.1, 0.2, 0.3]), "cons": np.array([1.0, 0.0, 2.0])}
    pr = ngp.BayesLV(0.01, "1 + maf + cons", cols, 0.5)
    m = ngp.LogVarModel(pr, 3, np.random.default_rng(0))
    assert m.covariates.shape == (3, 3) and np.array_equal(m.covariates[:, 0], np.ones(3)) and np.array_equal(m.covariates[:, 2], cols["cons"])
    assert np.allclose(m.logVar, np.log(0.01))
    CpC = m.covariates.T @ m.covariates
    assert np.allclose(np.linalg.inv(CpC + np.eye(3) * np.min(np.abs(np.diag(CpC))) / 10000), m.iCpC)
    with pytest.raises(ValueError):
        ngp.LogVarModel(pr, 4, np.random.default_rng(0))
