"""GPU parity of BayesR (functions.jl:238-289) through the C ABI against the CPU oracle (ngo_r_sweep): native Philox
chains, replayed variates (one uniform per cumulative comparison), fixed and Dirichlet-updated class proportions, the sweep-level
plugin call.  BayesR with <= 4 classes and no summary-statistic prior is swept by the blocked kernel's BayesR instantiation (class algebra
per lane in the chain warp: ngp_timing.kernel_variant == 7); more classes, summary statistics or NGP_KERNEL_LITERAL: the per-marker kernel."""
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
@pytest.mark.parametrize("n,p,kw", [(600, 120, {}), (1501, 77, dict(min_rows=64)), (400, 90, dict(kernel="literal", max_ctas=4)),
                                    (3001, 60, dict(max_ctas=3))])       # 1,504 rows per CTA with blocks of 64: the library falls back to the per-marker sweep
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
    assert g.timing()["kernel_variant"] == (3 if kw.get("kernel") == "literal" or kw.get("max_ctas") == 3 else 7)
    g.close()


@pytest.mark.parametrize("storage,block", [("i8", 32), ("2bit", 0), ("i8", 16)])
def test_bayesr_blocked_sweep_geometries_and_many_classes(gpu, storage, block):
    """the blocked BayesR sweep over block sizes and 2-bit tiles; 5 classes fall back to the per-marker kernel with the same results"""
    prob = make_problem(1300, 260, 35)
    ch, R, g = _pair(prob, 0.5, True, storage=storage, block=block, min_rows=8)
    g.set_rng(4, 2)
    for _ in range(5):
        ch.iteration(seed=4, chain=2)
        R.sweep(ch.e, ch.varE, it=ch.iter, seed=4, chain=2)
    g.run(5)
    st = g.state()
    assert g.timing()["kernel_variant"] == 7
    assert np.array_equal(st["sets"][0]["delta"], R.delta) and rel(st["sets"][0]["beta"], R.beta) < 1e-8 and rel(st["e"], ch.e) < 1e-8
    assert rel(st["sets"][0]["piHat"], R.piHat) < 1e-9 and rel(st["sets"][0]["varBeta"], R.varBeta) < 1e-8
    g.close()
    if storage == "i8" and block == 32:
        v5, p5 = np.array([0.0, 1e-5, 1e-4, 1e-3, 1e-2]), np.array([0.7, 0.1, 0.1, 0.07, 0.03])
        ch, R, g = _pair(prob, 0.5, True, v_class=v5, pi=p5)
        g.set_rng(4, 2)
        for _ in range(3):
            ch.iteration(seed=4, chain=2)
            R.sweep(ch.e, ch.varE, it=ch.iter, seed=4, chain=2)
        g.run(3)
        st = g.state()
        assert g.timing()["kernel_variant"] == 3
        assert np.array_equal(st["sets"][0]["delta"], R.delta) and rel(st["sets"][0]["beta"], R.beta) < 1e-8
        g.close()


def test_bayesr_at_headline_rows(gpu):
    """n = 50,000 x 1,024 markers, device-generated genotypes, blocked BayesR sweep against the oracle (native stream, 3 iterations).
    400 causal loci: the reference's class likelihoods exp(rhs^2 / 2 lhs) (functions.jl:255-256, no log-sum-exp) overflow for a locus whose
    chi-square exceeds ~1400, which at n = 50,000 is any locus explaining > 3 % of the variance — oracle and GPU both report that case as
    the error the reference would throw (test_bayesr_overflow_is_reported_like_the_reference)."""
    n, p, seed = 50000, 1024, 20261024
    pr = ngp.synth.problem(n, p, seed, q=400)
    codes = O.synth_codes(seed, n, 0, p, pr["thr0"], pr["thr1"])
    X, _, mpm = O.center_codes(codes)
    del codes
    v = pr["var_y"] / 2
    R = O.BayesROracle(X, mpm, PI0, VCLASS, v=v, est_pi=True)
    ch = O.OracleChain(pr["y"], [], v_e=pr["var_y"] / 2, intercept=True)
    g = ngp.Sampler(0)
    g.synth_genotypes(0, n, p, seed, pr["thr0"], pr["thr1"])
    g.set_prior(0, L.BAYESR, *O.marker_hyper(v), v, est_pi=True, v_class=VCLASS, pi_class=PI0)
    g.set_phenotype(pr["y"]); g.set_residual_prior(*O.residual_hyper(pr["var_y"] / 2)); g.set_intercept(True)
    g.set_rng(seed, 1)
    O.set_threads(max(1, len(__import__("os").sched_getaffinity(0))))
    try:
        for it in range(3):
            ch.iteration(seed=seed, chain=1)
            R.sweep(ch.e, ch.varE, it=ch.iter, seed=seed, chain=1)
            g.run(1)
            st = g.state()
            assert np.array_equal(st["sets"][0]["delta"], R.delta), f"class mismatch at iteration {it + 1}"
            assert rel(st["sets"][0]["beta"], R.beta) < 1e-7 and rel(st["e"], ch.e) < 1e-7 and abs(st["varE"] / ch.varE - 1) < 1e-8
    finally:
        O.set_threads(1)
    assert g.timing()["kernel_variant"] == 7 and g.timing()["block"] == 64
    g.close()


def test_bayesr_overflow_is_reported_like_the_reference(gpu):
    """a locus with a huge chi-square: every class likelihood is +Inf, the proportions are NaN and `findfirst` finds no class — the
    reference throws (functions.jl:261-262); the oracle returns its error code and the library NGP_ENUMERIC, on both kernels"""
    n, p, seed = 20000, 256, 20261025
    pr = ngp.synth.problem(n, p, seed, q=3)
    codes = O.synth_codes(seed, n, 0, p, pr["thr0"], pr["thr1"])
    X, _, mpm = O.center_codes(codes)
    y = pr["y"] - pr["y"].mean()
    R = O.BayesROracle(X, mpm, PI0, VCLASS, v=pr["var_y"] / 2, est_pi=True)
    ch = O.OracleChain(y, [], v_e=pr["var_y"] / 2, intercept=True)
    ch.iteration(seed=1, chain=0)
    with pytest.raises(AssertionError):
        R.sweep(ch.e, ch.varE, it=1, seed=1, chain=0)
    for kernel in ("blocked", "literal"):
        g = ngp.Sampler(0, kernel=kernel)
        g.upload_genotypes(0, codes)
        g.set_prior(0, L.BAYESR, *O.marker_hyper(pr["var_y"] / 2), pr["var_y"] / 2, est_pi=True, v_class=VCLASS, pi_class=PI0)
        g.set_phenotype(y); g.set_residual_prior(*O.residual_hyper(pr["var_y"] / 2)); g.set_intercept(True); g.set_rng(1, 0)
        with pytest.raises(ngp.NgpError) as ei:
            g.run(1)
        assert ei.value.code == L.ENUMERIC
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
