"""SURVEY §8 row f4, GRN single-site loop: sampleΛ2! of the gene-regulatory-network sampler (GRN.jl:150-164) through the sweep-level C ABI
(ngp_set_marker_summary + ngp_sweep with an improper BayesPR prior), against the statement-by-statement numpy restatement."""
import numpy as np
import pytest

import nextgp.jl_b200 as ngp
from common import make_problem, rel
from oracle import oracle as O
from oracle import restate_numpy as RN

pytestmark = pytest.mark.gpu


def _normals(seed, chain, it, p):
    L = O.lib()
    return np.array([L.ngo_stream_normal(seed, chain, it, 0, O.P_Z, j, 0) for j in range(p)])


@pytest.mark.parametrize("kernel", ["blocked", "literal"])
def test_sample_lambda2_matches_the_restatement(gpu, kernel):
    n, nSNPs, nGenes = 700, 96, 5
    prob = make_problem(n, nSNPs, 77)
    Xc = O.center_codes(prob["codes"])[0].T.copy()                       # SNPs x individuals, row-centred (GRN.jl:23)
    rng = np.random.default_rng(5)
    Y = rng.normal(size=(nGenes, n)) + 0.3 * prob["y"][None, :]
    Lambda2 = np.zeros((nGenes, nSNPs)); L2_o = Lambda2.copy()
    yCorr = Y - Y.mean(axis=1, keepdims=True); yC_o = yCorr.copy()
    pMeans = rng.normal(size=nGenes) * 0.1
    g = ngp.Sampler(0, kernel=kernel)
    g.upload_genotypes(0, prob["codes"])
    g.set_prior(0, ngp.BAYESPR, 4.0, 0.01, 0.01)
    g.set_rng(12, 3)
    nu_S, df_b = 0.02, 4.0
    var_tau = np.full(nGenes, 0.05)
    it = 0
    for sweep in range(3):
        varE = 0.8 + 0.1 * sweep
        z = np.empty((nGenes, nSNPs))
        for gi in range(nGenes):
            it += 1
            z[gi] = _normals(12, 3, it, nSNPs)
        RN.grn_sample_lambda2(L2_o, Xc, yC_o, var_tau, varE, pMeans, z)
        ngp.sampleLambda2(g, 0, Lambda2, yCorr, var_tau, varE, pMeans)
        assert rel(Lambda2, L2_o) < 1e-8 and rel(yCorr, yC_o) < 1e-8
        # GRN.jl:131-133 stays on the host: varBeta[g] = (nuS + Lambda2[g,:]'Lambda2[g,:]) / chi2(df + nSNPs)
        var_tau = (nu_S + (Lambda2 ** 2).sum(axis=1)) / rng.chisquare(df_b + nSNPs, size=nGenes)
    g.close()


def test_marker_summary_keeps_the_chain_state(gpu):
    prob = make_problem(300, 40, 3)
    g = ngp.Sampler(0)
    g.upload_genotypes(0, prob["codes"])
    g.set_prior(0, ngp.BAYESPR, 4.0, 0.01, 0.01)
    g.set_phenotype(prob["y"]); g.set_residual_prior(4.0, 0.5); g.set_intercept(True); g.set_rng(1, 0)
    g.run(2)
    before = g.state()
    g.set_marker_summary(0, np.full(40, 0.5), np.full(40, 0.1))
    after = g.state()
    assert np.array_equal(before["sets"][0]["beta"], after["sets"][0]["beta"]) and before["iter"] == after["iter"]
    g.run(1)                                                               # the offsets are used
    with_prior = g.state()["sets"][0]["beta"].copy()
    g.set_state(e=before["e"], mu=before["mu"], varE=before["varE"], iter=before["iter"], sets={0: before["sets"][0]})
    g.set_marker_summary(0, None, None)
    g.run(1)
    assert not np.array_equal(with_prior, g.state()["sets"][0]["beta"])
    with pytest.raises(AssertionError):
        g.set_marker_summary(0, np.zeros(3), None)
    g.close()
