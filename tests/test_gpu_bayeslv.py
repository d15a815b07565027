"""SURVEY §8 row f4, BayesLV: sampleBayesLV! (functions.jl:421-486) = the BayesPR sweep with one region per locus on the device
(ngp_sweep) + the host-side model of the log-variances, against the statement-by-statement numpy restatement; and the run-level wiring
(getMME / runSampler with VCV = {M: BayesLV(...)})."""
import os

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


@pytest.mark.parametrize("kernel,est", [("blocked", True), ("literal", False), ("blocked", 0.05)])
def test_sample_bayeslv_matches_the_restatement(gpu, kernel, est):
    n, p, k = 600, 128, 2
    prob = make_problem(n, p, 91)
    X, _, mpm = O.center_codes(prob["codes"])
    rng = np.random.default_rng(2)
    cov = np.column_stack([np.ones(p), rng.normal(size=p)])
    prior = ngp.BayesLV(0.02, None, cov, 0.3, estimateVarZeta=est)
    model = ngp.LogVarModel(prior, p, np.random.default_rng(9))
    twin = dict(logVar=model.logVar.copy(), SNPVARRESID=model.SNPVARRESID.copy(), covariates=cov.copy(), iCpC=model.iCpC.copy(),
                c=model.c.copy(), varZeta=[model.varZeta], estVarZeta=est)
    g = ngp.Sampler(0, kernel=kernel)
    g.upload_genotypes(0, prob["codes"])
    g.set_prior(0, ngp.BAYESPR, 4.0, 0.01, 0.02, region_off=np.arange(p + 1, dtype=np.int64))
    g.set_rng(21, 1)
    beta, delta, vb = np.zeros(p), np.ones(p, dtype=np.int64), np.full(p, 0.02)
    ycorr = prob["y"] - prob["y"].mean()
    beta_o, e_o, vb_o, zero = beta.copy(), ycorr.copy(), vb.copy(), np.zeros(p)
    for it in range(1, 6):
        varE = 0.9 + 0.05 * it
        u, zc = rng.random((p, 4)), rng.normal(size=k)
        RN.bayes_lv(X, mpm, zero, zero, beta_o, e_o, varE, vb_o, twin, _normals(21, 1, it, p), u, zc)
        ngp.sampleBayesLV(g, 0, model, beta, delta, ycorr, varE, vb, u=u, z=zc)
        assert rel(beta, beta_o) < 1e-8 and rel(ycorr, e_o) < 1e-8 and rel(vb, vb_o) < 1e-7
        assert rel(model.c, twin["c"]) < 1e-7 and abs(model.varZeta / twin["varZeta"][0] - 1) < 1e-7
    assert np.ptp(vb) > 0                                     # the variances moved apart
    g.close()


def test_bayeslv_at_run_level_writes_the_reference_files(gpu, tmp_path):
    n, p = 400, 60
    prob = make_problem(n, p, 5)
    out = str(tmp_path / "outMCMC")
    cov = {"x": np.linspace(-1, 1, p)}
    s = ngp.runLMEM("y ~ 1 + SNP(M,geno)", {"y": prob["y"]}, 12, 2, 2, outFolder=out, matrices={"M": prob["codes"]},
                    VCV={"M": ngp.BayesLV(0.01, "1 + x", cov, 0.5, estimateVarZeta=True), "e": ngp.Random("I", prob["var_y"] / 2)}, seed=4)
    rows = lambda f: open(os.path.join(out, f)).read().strip().split("\n")
    assert len(rows("cMOut")) == 1 + 5 and rows("cMOut")[0].split("\t") == ["c1", "c2"]
    assert len(rows("varZetaMOut")) == 1 + 5 and rows("varZetaMOut")[0] == "varZeta"
    var = np.array([[float(x) for x in r.split("\t")] for r in rows("varMOut")[1:]])
    assert var.shape == (5, p) and np.all(var > 0) and np.ptp(var[-1]) > 0
    st = s.state(want_e=False)
    s.close()
