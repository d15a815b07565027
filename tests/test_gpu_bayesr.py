"""GPU parity of BayesR (functions.jl:238-289) through the C ABI against the CPU oracle (ngo_r_sweep): native Philox
chains, replayed variates (one uniform per cumulative comparison), fixed and Dirichlet-updated class proportions, the sweep-level
plugin call.  BayesR is sampled by the per-marker kernel whatever NGP_CFG_KERNEL says."""
import numpy as np
import pytest

import nextgp.jl_b200 as ngp
from nextgp.jl_b200 import _lib as L
from common import make_problem, rel
from oracle import oracle as O

pytestmark = pytest.mark.gpu

VCLASS = np.array([0.0, 0.0001, 0.001, 0.01])
PI0 = np.array([0.8, 0.1, 0.07, 0.03])


def _pair(prob, v, est_pi, v_class=VCLASS, pi=PI0, kernel="blocked", **kw):
    X, mean, mpm = O.center_codes(prob["codes"])
    R = O.BayesROracle(X, mpm, pi, v_class, v=v, est_pi=est_pi)
    ch = O.OracleChain(prob["y"], [], v_e=prob["var_y"] / 2, intercept=True)
    g = ngp.Sampler(0, kernel=kernel, **kw)
    g.upload_genotypes(0, prob["codes"])
    df, scale = O.marker_hyper(v)
    g.set_prior(0, L.BAYESR, df, scale, v, est_pi=est_pi, v_class=v_class, pi_class=pi)
    g.set_phenotype(prob["y"])
    g.set_residual_prior(*O.residual_hyper(prob["var_y"] / 2))
    g.set_intercept(True)
    return ch, R, g


@pytest.mark.parametrize("est_pi", [False, True])
@pytest.mark.parametrize("n,p,kw", [(600, 120, {}), (1501, 77, dict(min_rows=64)), (400, 90, dict(kernel="literal", max_ctas=4))])
def test_bayesr_native_chain_matches_oracle(gpu, n, p, kw, est_pi):
    prob = make_problem(n, p, 31)
    ch, R, g = _pair(prob, 0.5, est_pi, **kw)
    g.set_rng(17, 1)
    for _ in range(6):
        ch.iteration(seed=17, chain=1)
        R.sweep(ch.e, ch.varE, it=ch.iter, seed=17, chain=1)
    g.run(4)
    g.run(2)
    st = g.state()
    assert np.array_equal(st["sets"][0]["delta"], R.delta)
    assert rel(st["sets"][0]["beta"], R.beta) < 1e-8 and rel(st["sets"][0]["varBeta"], R.varBeta) < 1e-8
    assert rel(st["sets"][0]["piHat"], R.piHat) < 1e-9 and rel(st["e"], ch.e) < 1e-8
    assert abs(st["varE"] / ch.varE - 1) < 1e-9
    g.close()


def test_bayesr_replay_and_sweep_level_call(gpu):
    prob = make_problem(800, 100, 44)
    ch, R, g = _pair(prob, 0.5, True)
    logs = []
    for _ in range(4):
        lg = ch.iteration(seed=3, chain=0)
        lg["sets"] = [R.sweep(ch.e, ch.varE, it=ch.iter, seed=3, chain=0)]
        logs.append(lg)
    g.set_rng(999, 5)                       # a different stream: everything must come from the log
    g.set_replay(logs)
    g.run(4)
    st = g.state()
    assert np.array_equal(st["sets"][0]["delta"], R.delta) and rel(st["sets"][0]["beta"], R.beta) < 1e-8
    assert rel(st["sets"][0]["piHat"], R.piHat) < 1e-12 and rel(st["e"], ch.e) < 1e-8
    g.close()
    # M[mSet].funct(mSet, M, beta, delta, ycorr, varE, varBeta) with host buffers
    ch2, R2, g = _pair(prob, 0.5, True)
    g.set_rng(8, 0)
    e = prob["y"] - prob["y"].mean(); e_o = e.copy()
    beta, delta, vb, ph = np.zeros(100), np.ones(100, dtype=np.int64), np.array([0.5]), PI0.copy()
    for it in (1, 2):
        R2.sweep(e_o, 2.3, it=it, seed=8, chain=0)
        g.sweep(0, e, 2.3, beta, delta, vb, ph)
        assert np.array_equal(delta, R2.delta) and rel(beta, R2.beta) < 1e-8 and rel(vb, R2.varBeta) < 1e-9
        assert rel(ph, R2.piHat) < 1e-9 and rel(e, e_o) < 1e-8
    g.close()


def test_bayesr_argument_checks(gpu):
    prob = make_problem(200, 30, 2)
    g = ngp.Sampler(0)
    g.upload_genotypes(0, prob["codes"])
    with pytest.raises(ngp.NgpError):
        g.set_prior(0, L.BAYESR, 4.0, 0.1, 0.2, v_class=np.arange(9.0), pi_class=np.full(9, 1 / 9))      # too many classes
    with pytest.raises(ngp.NgpError):
        g.set_prior(0, L.BAYESR, 4.0, 0.1, 0.2, v_class=VCLASS, pi_class=np.array([1.0, 0.0, 0.0, 0.0]))  # log(0) proportions
    g.close()


def test_runLMEM_bayesr_writes_class_proportions(gpu, tmp_path):
    import os
    prob = make_problem(300, 40, 9)
    out = str(tmp_path / "outMCMC")
    VCV = {"M": ngp.BayesR(PI0, VCLASS, 0.5, estimatePi=True), "e": ngp.Random("I", prob["var_y"] / 2)}
    s = ngp.runLMEM("y ~ 1 + SNP(M,x)", {"y": prob["y"]}, 20, 10, 5, outFolder=out, VCV=VCV, seed=4, matrices={"M": prob["codes"]})
    assert sorted(os.listdir(out)) == ["bOut", "betaMOut", "deltaMOut", "piMOut", "varEOut", "varMOut"]
    assert open(os.path.join(out, "piMOut")).readline().strip().split("\t") == ["pi1", "pi2", "pi3", "pi4"]
    pis = np.loadtxt(os.path.join(out, "piMOut"), delimiter="\t", skiprows=1)
    dl = np.loadtxt(os.path.join(out, "deltaMOut"), delimiter="\t", skiprows=1)
    assert pis.shape == (2, 4) and np.allclose(pis.sum(1), 1.0) and set(np.unique(dl)) <= {1, 2, 3, 4}
    s.close()


def test_bayesr_and_tuple_replay_against_committed_golden(gpu):
    """The device consumes the committed variate logs (tests/golden/bayesr.npz, tuple2.npz) and must land on the committed states —
    no oracle code runs in this test."""
    import importlib.util
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(gold, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    c = mg.EXTRA["bayesr"]; g = np.load(os.path.join(gold, "bayesr.npz"))
    prob = make_problem(c["n"], c["p"], c["seed"])
    s = ngp.Sampler(0)
    s.upload_genotypes(0, prob["codes"])
    v = c["v"]
    s.set_prior(0, L.BAYESR, 4.0, v * 0.5, v, est_pi=True, v_class=np.array(c["v_class"]), pi_class=np.array(c["pi"]))
    s.set_phenotype(prob["y"]); s.set_residual_prior(4.0, prob["var_y"] / 2 * 0.5); s.set_intercept(True)
    ni = c["iters"]
    s.set_replay([{"chi2_e": g["chi2_e"][i], "z_mu": g["z_mu"][i],
                   "sets": [{"u": g["u"][i], "z": g["z"][i], "chi2_b": g["chi2_b"][i], "dir_pi": g["dir_pi"][i]}]} for i in range(ni)])
    s.run(ni)
    st = s.state()
    assert np.array_equal(st["sets"][0]["delta"], g["delta"][-1]) and rel(st["sets"][0]["beta"], g["beta"][-1]) < 1e-8
    assert rel(st["sets"][0]["piHat"], g["pi"][-1]) < 1e-12 and rel(st["e"], g["e_final"]) < 1e-8
    s.close()
    c = mg.EXTRA["tuple2"]; g = np.load(os.path.join(gold, "tuple2.npz"))
    probs, y = mg.breeds(c["n"], c["p"], 2, c["seed"])
    V = np.array(c["V"])
    s = ngp.Sampler(0)
    for b, pr in enumerate(probs):
        s.upload_genotypes(b, pr["codes"])
    s.set_joint_prior([0, 1], 5.0, V * 2.0, V, region_off=np.array(c["region_off"], dtype=np.int64))
    s.set_phenotype(y); s.set_residual_prior(4.0, float(np.var(y)) / 2 * 0.5); s.set_intercept(True)
    ni = c["iters"]
    s.set_replay([{"chi2_e": g["chi2_e"][i], "z_mu": g["z_mu"][i], "sets": []} for i in range(ni)])
    s.set_joint_replay([{"z": g["z"][i], "iw_chi2": g["iw_chi2"][i], "iw_z": g["iw_z"][i]} for i in range(ni)])
    s.run(ni)
    js = s.joint_state()
    assert rel(js["beta"], g["beta"][-1]) < 1e-8 and rel(js["varBeta"], g["varBeta"][-1]) < 1e-8 and rel(s.state()["e"], g["e_final"]) < 1e-8
    s.close()
