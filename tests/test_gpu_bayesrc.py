"""GPU parity of BayesRCpi (functions.jl:291-360) and BayesRCplus (functions.jl:362-419; wiring mme.jl:385-418) through the
C ABI against the CPU oracle (ngo_rc_sweep, itself checked against a literal numpy restatement in tests/test_oracle.py): native Philox
chains and replayed variates, fixed and Dirichlet-updated class proportions per annotation, annotProb / annotCat.  Swept by the per-marker
kernel (one lane of the chain warp per annotation and class)."""
import numpy as np
import pytest

import nextgp.jl_b200 as ngp
from nextgp.jl_b200 import _lib as L
from common import make_problem, rel
from oracle import oracle as O

pytestmark = pytest.mark.gpu

VCLASS = np.array([0.0, 0.001, 0.01, 0.1])
PI0 = np.array([0.7, 0.15, 0.1, 0.05])


def _annot(p, nA, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2, size=(p, nA)).astype(np.int32)
    a[a.sum(1) == 0, 0] = 1
    a[1, :] = np.arange(nA, 0, -1)                    # counts, not only 0 / 1
    return a


def _pair(prob, plus, est_pi, nA=3, v=0.6, **kw):
    X, _, mpm = O.center_codes(prob["codes"])
    p = X.shape[1]
    annot = _annot(p, nA, 3)
    R = O.BayesRCOracle(X, mpm, PI0, VCLASS, v=v, annot=annot, est_pi=est_pi, plus=plus)
    ch = O.OracleChain(prob["y"], [], v_e=prob["var_y"] / 2, intercept=True)
    g = ngp.Sampler(0, **kw)
    g.upload_genotypes(0, prob["codes"])
    g.set_rc_prior(0, plus, *O.marker_hyper(v), v, VCLASS, PI0, annot, est_pi=est_pi)
    g.set_phenotype(prob["y"]); g.set_residual_prior(*O.residual_hyper(prob["var_y"] / 2)); g.set_intercept(True)
    return ch, R, g


def _check(g, ch, R, plus, tol=1e-8):
    st = g.state()
    rs = g.rc_state(0)
    assert np.array_equal(st["sets"][0]["delta"], R.delta)
    assert rel(st["sets"][0]["beta"], R.beta) < tol and rel(st["sets"][0]["varBeta"], R.varBeta) < tol and rel(st["e"], ch.e) < tol
    assert abs(st["varE"] / ch.varE - 1) < 1e-9 and rel(rs["piHat"], R.piHat) < 1e-9
    if not plus:
        assert np.array_equal(rs["annot_cat"], R.annot_cat) and rel(rs["annot_prob"], R.annot_prob) < 1e-9


@pytest.mark.parametrize("plus", [False, True], ids=["RCpi", "RCplus"])
@pytest.mark.parametrize("est_pi", [False, True])
@pytest.mark.parametrize("n,p,nA,kw", [(600, 90, 3, {}), (1501, 70, 5, dict(min_rows=64)), (300, 40, 8, dict(max_ctas=4))])
def test_bayesrc_native_chain_matches_oracle(gpu, plus, est_pi, n, p, nA, kw):
    prob = make_problem(n, p, 37)
    ch, R, g = _pair(prob, plus, est_pi, nA=nA, **kw)
    g.set_rng(17, 1)
    for it in range(5):
        ch.iteration(seed=17, chain=1)
        R.sweep(ch.e, ch.varE, it=ch.iter, seed=17, chain=1)
        g.run(1)
        _check(g, ch, R, plus)
    assert g.timing()["kernel_variant"] == 3
    g.close()


@pytest.mark.parametrize("plus", [False, True], ids=["RCpi", "RCplus"])
def test_bayesrc_replay(gpu, plus):
    prob = make_problem(700, 80, 41)
    ch, R, g = _pair(prob, plus, True)
    logs, rlogs = [], []
    for _ in range(4):
        logs.append(ch.iteration(seed=3, chain=0))
        rlogs.append(R.sweep(ch.e, ch.varE, it=ch.iter, seed=3, chain=0))
    g.set_rng(999, 5)                       # a different stream: everything must come from the logs
    g.set_replay(logs)
    g.set_rc_replay(0, rlogs)
    g.run(3)
    g.run(1)
    _check(g, ch, R, plus)
    g.close()


def test_bayesrc_argument_checks(gpu):
    prob = make_problem(200, 30, 2)
    g = ngp.Sampler(0)
    g.upload_genotypes(0, prob["codes"])
    annot = _annot(30, 3, 1)
    with pytest.raises(ngp.NgpError):
        g.set_rc_prior(0, False, 4.0, 0.1, 0.2, np.arange(5.0), np.full(5, 0.2), _annot(30, 8, 1))       # 8 annotations x 5 classes > 32 lanes
    bad = annot.copy(); bad[7, :] = 0
    with pytest.raises(ngp.NgpError) as ei:
        g.set_rc_prior(0, False, 4.0, 0.1, 0.2, VCLASS, PI0, bad)                                       # a locus without annotation
    assert ei.value.code == L.EDATA
    g.set_rc_prior(0, True, 4.0, 0.1, 0.2, VCLASS, PI0, annot)
    g.set_phenotype(prob["y"]); g.set_residual_prior(4.0, 1.0); g.set_intercept(True)
    with pytest.raises(ngp.NgpError):
        g.sweep(0, prob["y"].copy(), 1.0, np.zeros(30), np.ones(30, dtype=np.int64), np.full(3, 0.2))     # run level only
    g.run(2)
    assert g.state()["sets"][0]["varBeta"].shape == (3,)
    g.close()


def test_runLMEM_bayesrc_writes_reference_output_files(gpu, tmp_path):
    """runLMEM with priorVCV[M] = BayesRCπ(...): the files and header rows of mme.jl:571-576 / samplers.jl:85-88"""
    import os
    prob = make_problem(300, 40, 9)
    annot = _annot(40, 2, 5)
    out = str(tmp_path / "outMCMC")
    VCV = {"M": ngp.BayesRCpi(PI0, VCLASS, 0.5, annot, estimatePi=True), "e": ngp.Random("I", prob["var_y"] / 2)}
    s = ngp.runLMEM("y ~ 1 + SNP(M,x)", {"y": prob["y"]}, 20, 10, 5, outFolder=out, VCV=VCV, seed=4, matrices={"M": prob["codes"]})
    assert sorted(os.listdir(out)) == ["annotMOut", "bOut", "betaMOut", "deltaMOut", "piMOut", "varEOut", "varMOut"]
    assert open(os.path.join(out, "piMOut")).readline().strip().split("\t") == [f"pi{v}" for v in range(1, 9)]
    assert open(os.path.join(out, "varMOut")).readline().strip().split("\t") == ["reg_1", "reg_2"]
    an = np.loadtxt(os.path.join(out, "annotMOut"), delimiter="\t", skiprows=1)
    pis = np.loadtxt(os.path.join(out, "piMOut"), delimiter="\t", skiprows=1)
    assert an.shape == (2, 40) and set(np.unique(an)) <= {1, 2} and pis.shape == (2, 8) and np.allclose(pis.reshape(2, 2, 4).sum(2), 1.0)
    s.close()


@pytest.mark.parametrize("name", ["bayesrc_pi", "bayesrc_plus"])
def test_bayesrc_replay_against_committed_golden(gpu, name):
    """The device consumes the committed variate logs (tests/golden/bayesrc_*.npz) and must land on the committed states — no oracle
    code runs in this test."""
    import importlib.util
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(gold, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    c = mg.RC[name]; g = np.load(os.path.join(gold, name + ".npz"))
    prob = make_problem(c["n"], c["p"], c["seed"])
    s = ngp.Sampler(0)
    s.upload_genotypes(0, prob["codes"])
    v = c["v"]
    s.set_rc_prior(0, c["plus"], 4.0, v * 0.5, v, np.array(c["v_class"]), np.array(c["pi"]), mg.rc_annot(c["p"], c["n_annot"], c["seed"]), est_pi=True)
    s.set_phenotype(prob["y"]); s.set_residual_prior(4.0, prob["var_y"] / 2 * 0.5); s.set_intercept(True)
    ni = c["iters"]
    s.set_replay([{"chi2_e": g["chi2_e"][i], "z_mu": g["z_mu"][i], "sets": []} for i in range(ni)])
    s.set_rc_replay(0, [{k: g[k][i] for k in ("u_annot", "dirp", "u", "z", "chi2_b", "dir_pi")} for i in range(ni)])
    s.run(ni)
    st, rs = s.state(), s.rc_state(0)
    assert np.array_equal(st["sets"][0]["delta"], g["delta"][-1]) and rel(st["sets"][0]["beta"], g["beta"][-1]) < 1e-8
    assert rel(st["sets"][0]["varBeta"], g["varBeta"][-1]) < 1e-8 and rel(rs["piHat"], g["pi"][-1]) < 1e-12 and rel(st["e"], g["e_final"]) < 1e-8
    if not c["plus"]:
        assert np.array_equal(rs["annot_cat"], g["annot_cat"][-1]) and rel(rs["annot_prob"], g["annot_prob_final"]) < 1e-12
    s.close()
