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


def test_grn_lambda2_is_a_bayespr_sweep_with_an_improper_prior():
    """The recipe behind api.sampleLambda2 / ngp_set_marker_summary, on the CPU: sampleΛ2! (GRN.jl:150-164) for one gene equals the BayesPR
    loop (functions.jl:118-137) with varBeta = Inf, lhs0 = 0 and rhs0 = pMeans[g] / σ2τ[g] on every SNP (α pMeans / σ2ϵ), same normals."""
    rng = np.random.default_rng(3)
    n, p, nGenes = 80, 40, 3
    X = rng.integers(0, 3, size=(n, p)).astype(np.float64); X -= X.mean(axis=0)
    mpm = (X * X).sum(axis=0)
    Y = rng.normal(size=(nGenes, n))
    L2_a, yC_a = np.zeros((nGenes, p)), Y.copy()
    L2_b, yC_b = np.zeros((nGenes, p)), Y.copy()
    var_tau, pMeans, varE = np.array([0.3, 0.05, 1.7]), np.array([0.2, -0.1, 0.05]), 0.9
    for sweep in range(3):
        z = rng.normal(size=(nGenes, p))
        RN.grn_sample_lambda2(L2_a, X.T.copy(), yC_a, var_tau, varE, pMeans, z)
        for g in range(nGenes):
            vb = np.array([np.inf])
            with np.errstate(divide="ignore"):
                RN.bayes_pr(X, mpm, np.zeros(p), np.full(p, pMeans[g] / var_tau[g]), [0, p], 0.01, 4.0, L2_b[g], yC_b[g], varE, vb, z[g], np.array([50.0]))
        assert np.allclose(L2_a, L2_b, rtol=1e-12, atol=1e-14) and np.allclose(yC_a, yC_b, rtol=1e-11, atol=1e-13)
