"""GPU parity of weighted residuals, E.str == "D" (mme.jl:70-73, 133-136, 299-303; functions.jl:526-528), through the C ABI
against the CPU oracle: the residual variance uses sum(w .* e.^2), the intercept xpx = 1'W1 and Xp = w', every marker Mp_j = (x_j .* w)'
and mpm_j = sum(x_j .* w .* x_j) — while the add-back and the BayesB/C inclusion dot stay unweighted (functions.jl:168, :208), so those
samplers form two dots per marker.  Sampled by the per-marker kernel whatever NGP_CFG_KERNEL says.
Tolerance as in test_gpu_parity.py (BASELINE.json north_star): replayed draws reproduce effects and variances to 1e-9 relative."""
import ctypes as C

import numpy as np
import pytest

import nextgp.jl_b200 as ngp
from common import make_problem, gpu_sampler, rel
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-9

CASES = [
    ("BayesPR-RR", 0, dict(v=0.01)),
    ("BayesPR-regions", 0, dict(v=0.01, region_off=[0, 17, 18, 90, 200, 333])),
    ("BayesB", 1, dict(v=0.05, pi=0.1, est_pi=True)),
    ("BayesC-pi", 2, dict(v=0.05, pi=0.05, est_pi=True)),
    ("BayesC-fixedpi", 2, dict(v=0.05, pi=0.5, est_pi=False)),
]


def _weights(n, seed):
    return np.random.default_rng(seed).uniform(0.25, 4.0, n)


def _oracle(prob, w, method, intercept=True, lhs0=None, rhs0=None, **kw):
    X, mean, _ = O.center_codes(prob["codes"])
    Mp, mpm = O.weighted_marker_arrays(X, w)
    ro = kw.pop("region_off", None)
    S = O.MarkerSet(X=X, mpm=mpm, method=method, region_off=None if ro is None else np.array(ro), lhs0=lhs0, rhs0=rhs0, **kw)
    S.Mp = Mp
    return O.OracleChain(prob["y"], [S], v_e=prob["var_y"] / 2, intercept=intercept, weights=w), S


def _check(st, o, method, intercept=True):
    worst = max(rel(st["sets"][0]["beta"], o["sets"][0]["beta"]), abs(st["varE"] / o["varE"] - 1),
                abs(st["mu"] - o["mu"]) / max(abs(o["mu"]), 1e-300) if intercept else 0.0,
                rel(st["sets"][0]["varBeta"], o["sets"][0]["varBeta"]), rel(st["e"], o["e"]))
    assert np.array_equal(st["sets"][0]["delta"], o["sets"][0]["delta"])
    if method:
        worst = max(worst, rel(st["sets"][0]["piHat"], o["sets"][0]["piHat"]))
    return worst


@pytest.mark.parametrize("kernel", ["blocked", "literal"])
@pytest.mark.parametrize("name,method,kw", CASES, ids=[c[0] for c in CASES])
def test_weighted_replay_parity(gpu, kernel, name, method, kw):
    prob = make_problem(640, 333, 23)
    w = _weights(640, 5)
    ch, S = _oracle(prob, w, method, **dict(kw))
    logs, snaps = [], []
    for _ in range(10):
        logs.append(ch.iteration(seed=42, chain=1))
        snaps.append(ch.snapshot())
    kw2 = dict(kw)
    ro = kw2.pop("region_off", None)
    g = gpu_sampler(prob, method, region_off=None if ro is None else np.array(ro), kernel=kernel, **kw2)
    g.set_residual_weights(w)
    assert rel(g.column_stats(0)[1], S.mpm) < 1e-12     # weighted as soon as the weights are set, whatever the call order
    g.set_replay(logs)
    for it in range(10):
        g.run(1)
        worst = _check(g.state(), snaps[it], method)
        assert worst < TOL, f"iteration {it + 1}: rel diff {worst}"
    mean, mpm = g.column_stats(0)
    assert rel(mpm, S.mpm) < 1e-12                      # the weighted mpm of mme.jl:301
    g.close()


def test_weighted_ragged_shape_no_intercept_summary_priors(gpu):
    n, p = 257, 71
    prob = make_problem(n, p, 29)
    w = _weights(n, 9)
    rng = np.random.default_rng(3)
    lhs0, rhs0 = rng.uniform(0.0, 2.0, p), rng.normal(0, 0.5, p)
    for method, kw in [(0, dict(v=0.02)), (1, dict(v=0.05, pi=0.3, est_pi=True)), (2, dict(v=0.05, pi=0.3, est_pi=True))]:
        ch, S = _oracle(prob, w, method, intercept=False, lhs0=lhs0, rhs0=rhs0, **kw)
        logs, snaps = [], []
        for _ in range(6):
            logs.append(ch.iteration(seed=1, chain=0))
            snaps.append(ch.snapshot())
        g = gpu_sampler(prob, method, intercept=False, lhs0=lhs0, rhs0=rhs0, min_rows=8, **kw)
        g.set_residual_weights(w)
        g.set_replay(logs)
        for it in range(6):
            g.run(1)
            assert _check(g.state(), snaps[it], method, intercept=False) < TOL
        g.close()


def test_unit_weights_equal_the_unweighted_chain(gpu):
    """w = 1 is E.str == "I": the weighted path must reproduce the unweighted chain (same Philox draws)."""
    prob = make_problem(500, 200, 31)
    kw = dict(v=0.05, pi=0.1, est_pi=True)
    a = gpu_sampler(prob, 2, kernel="literal", **kw)
    b = gpu_sampler(prob, 2, kernel="literal", **kw)
    b.set_residual_weights(np.ones(500))
    for s in (a, b):
        s.set_rng(77, 0)
        s.run(15)
    sa, sb = a.state(), b.state()
    assert np.array_equal(sa["sets"][0]["delta"], sb["sets"][0]["delta"])
    assert rel(sb["sets"][0]["beta"], sa["sets"][0]["beta"]) < 1e-9 and abs(sb["varE"] / sa["varE"] - 1) < 1e-9
    # back to "I"
    b.set_residual_weights(None)
    a.run(3); b.run(3)
    assert rel(b.state()["sets"][0]["beta"], a.state()["sets"][0]["beta"]) < 1e-9
    a.close(); b.close()


def test_weighted_native_chain_matches_oracle_native_chain(gpu):
    prob = make_problem(400, 150, 37)
    w = _weights(400, 11)
    kw = dict(v=0.05, pi=0.2, est_pi=True)
    ch, S = _oracle(prob, w, 2, **kw)
    g = gpu_sampler(prob, 2, **kw)
    g.set_residual_weights(w)
    g.set_rng(123, 2)
    for it in range(8):
        ch.iteration(seed=123, chain=2)
        g.run(1)
        assert _check(g.state(), ch.snapshot(), 2) < 1e-8
    g.close()


def test_weighted_sweep_level_plugin_call(gpu):
    """M[mSet].funct(...) with E.str == "D": Julia keeps varE and the intercept (weighted, here by the oracle's scalar routines)."""
    prob = make_problem(450, 180, 41)
    w = _weights(450, 13)
    kw = dict(v=0.05, pi=0.1, est_pi=True)
    ch, S = _oracle(prob, w, 2, **kw)
    g = gpu_sampler(prob, 2, **kw)
    g.set_residual_weights(w)
    beta, delta = np.zeros(180), np.ones(180, dtype=np.int64)
    varBeta, piHat = np.full(1, kw["v"]), np.array([1 - kw["pi"], kw["pi"]])
    ycorr, mu = prob["y"].copy(), 0.0
    Lo = O.lib()
    for it in range(1, 6):
        log = ch.iteration(seed=3, chain=0)
        c2 = C.c_double(log["chi2_e"]); zm = C.c_double(log["z_mu"])
        varE = Lo.ngo_sample_varE_w(len(ycorr), ycorr.ctypes.data, w.ctypes.data, 4.0, ch.scale_e, 1, 0, 0, it, C.byref(c2))
        mu = Lo.ngo_sample_intercept_w(len(ycorr), ycorr.ctypes.data, w.ctypes.data, mu, varE, 0.0, 0.0, 1, 0, 0, it, C.byref(zm))
        g.set_replay([log])
        g.sweep(0, ycorr, varE, beta, delta, varBeta, piHat)
        assert rel(beta, S.beta) < TOL and rel(ycorr, ch.e) < TOL and rel(varBeta, S.varBeta) < TOL
        assert np.array_equal(delta, S.delta) and rel(piHat, S.piHat) < TOL
    g.close()


def test_weight_argument_errors(gpu):
    prob = make_problem(128, 40, 43)
    g = gpu_sampler(prob, 0, v=0.01)
    with pytest.raises(ngp.NgpError):
        g.set_residual_weights(np.ones(127))
    bad = np.ones(128); bad[5] = 0.0
    with pytest.raises(ngp.NgpError):
        g.set_residual_weights(bad)
    g.close()


def test_runLMEM_with_D_vector(gpu, tmp_path):
    """priorVCV[:e] = Random(D, v) with a vector D (d_ii = 1/w_ii, mme.jl:70-73) through the runLMEM mirror; the final state equals an
    oracle chain with iVarStr = inv.(D) on the same native stream."""
    import os
    n, p = 200, 50
    prob = make_problem(n, p, 71)
    dvec = np.random.default_rng(2).uniform(0.5, 3.0, n)
    geno = tmp_path / "geno.txt"
    np.savetxt(geno, prob["codes"], fmt="%d", delimiter=" ")
    out = str(tmp_path / "outMCMC")
    VCV = {"M": ngp.BayesC(0.2, 0.01, estimatePi=True), "e": ngp.Random(dvec, prob["var_y"] / 2)}
    s = ngp.runLMEM(f'y ~ 1 + SNP(M,"{geno}")', {"y": prob["y"]}, 40, 20, 10, outFolder=out, VCV=VCV, seed=5)
    beta = np.loadtxt(os.path.join(out, "betaMOut"), delimiter="\t", skiprows=1)
    ch, S = _oracle(prob, 1.0 / dvec, 2, v=0.01, pi=0.2, est_pi=True)
    for _ in range(40):
        ch.iteration(seed=5, chain=0)
    assert rel(beta[-1], S.beta) < 1e-7
    s.close()
    with pytest.raises(ValueError):
        ngp.runLMEM(f'y ~ 1 + SNP(M,"{geno}")', {"y": prob["y"]}, 4, 2, 1, outFolder=str(tmp_path / "o2"),
                    VCV={"M": ngp.BayesPR(9999, 0.01), "e": ngp.Random(np.ones(n - 1), 1.0)})


@pytest.mark.parametrize("est_pi", [False, True])
def test_weighted_bayesr_native_chain_matches_oracle(gpu, est_pi):
    """BayesR (functions.jl:238-289) takes Mp and the weighted mpm only (rhs at :250; no unweighted dot)."""
    from nextgp.jl_b200 import _lib as L
    n, p = 600, 120
    prob = make_problem(n, p, 31)
    w = _weights(n, 17)
    v_class, pi = np.array([0.0, 0.0001, 0.001, 0.01]), np.array([0.8, 0.1, 0.07, 0.03])
    X, mean, _ = O.center_codes(prob["codes"])
    Mp, mpm = O.weighted_marker_arrays(X, w)
    R = O.BayesROracle(X, mpm, pi, v_class, v=0.5, est_pi=est_pi)
    R.Mp = Mp
    ch = O.OracleChain(prob["y"], [], v_e=prob["var_y"] / 2, intercept=True, weights=w)
    g = ngp.Sampler(0)
    g.upload_genotypes(0, prob["codes"])
    df, scale = O.marker_hyper(0.5)
    g.set_prior(0, L.BAYESR, df, scale, 0.5, est_pi=est_pi, v_class=v_class, pi_class=pi)
    g.set_phenotype(prob["y"])
    g.set_residual_prior(*O.residual_hyper(prob["var_y"] / 2))
    g.set_residual_weights(w)
    g.set_intercept(True)
    g.set_rng(17, 1)
    for _ in range(6):
        ch.iteration(seed=17, chain=1)
        R.sweep(ch.e, ch.varE, it=ch.iter, seed=17, chain=1)
    g.run(6)
    st = g.state()
    assert np.array_equal(st["sets"][0]["delta"], R.delta)
    assert rel(st["sets"][0]["beta"], R.beta) < 1e-8 and rel(st["sets"][0]["varBeta"], R.varBeta) < 1e-8
    assert rel(st["sets"][0]["piHat"], R.piHat) < 1e-9 and rel(st["e"], ch.e) < 1e-8
    assert abs(st["varE"] / ch.varE - 1) < 1e-9
    g.close()


@pytest.mark.parametrize("method,kw", [(2, dict(v=0.05, pi=0.1, est_pi=True)), (0, dict(v=0.01)), (1, dict(v=0.05, pi=0.3, est_pi=True))])
def test_weighted_fixed_effects_native_chain_matches_oracle(gpu, method, kw):
    """Covariates and factor levels besides the intercept under E.str == "D": xpx = X'(w .* X), Xp = (X .* w)' (mme.jl:133-136) in
    sampleX! / sampleb! (functions.jl:22-54); a single column with prior information and two multi-column sets."""
    n, p = 777, 130
    prob = make_problem(n, p, 41)
    w = _weights(n, 19)
    rng = np.random.default_rng(1)
    age = rng.normal(size=n) * 2 + 5
    herd = np.eye(3)[rng.integers(0, 3, n)][:, 1:]
    cov2 = np.column_stack([rng.uniform(size=n), rng.normal(size=n), (rng.uniform(size=n) > 0.5).astype(float)])
    specs = [(age, 0.7, -0.3), (herd, 0.0, 0.0), (cov2, 0.0, 0.0)]
    y = prob["y"] + 0.8 * age
    fixed_o, col = [], 1
    for data, l0, r0 in specs:
        F = O.FixedSet(data, col0=col, lhs0=l0, rhs0=r0, weights=w)
        col += F.c
        fixed_o.append(F)
    X, mean, _ = O.center_codes(prob["codes"])
    Mp, mpm = O.weighted_marker_arrays(X, w)
    S = O.MarkerSet(X=X, mpm=mpm, method=method, **kw)
    S.Mp = Mp
    ch = O.OracleChain(y, [S], v_e=prob["var_y"] / 2, intercept=True, fixed=fixed_o, weights=w)
    g = ngp.Sampler(0)
    g.upload_genotypes(0, prob["codes"])
    df, scale = O.marker_hyper(kw["v"])
    g.set_prior(0, method, df, scale, kw["v"], pi_in=kw.get("pi", 0.0), est_pi=kw.get("est_pi", False))
    g.set_fixed_effects(specs)
    g.set_phenotype(y)
    g.set_residual_prior(*O.residual_hyper(prob["var_y"] / 2))
    g.set_residual_weights(w)
    g.set_intercept(True)
    g.set_rng(12, 2)
    for _ in range(6):
        ch.iteration(seed=12, chain=2)
    g.run(4); g.run(2)
    st = g.state()
    assert rel(g.fixed_effects(), np.concatenate([F.b for F in fixed_o])) < 1e-8 and abs(st["mu"] - ch.mu) < 1e-8 * max(1.0, abs(ch.mu))
    if method:
        assert np.array_equal(st["sets"][0]["delta"], S.delta)
    assert rel(st["sets"][0]["beta"], S.beta) < 1e-8 and rel(st["e"], ch.e) < 1e-8 and abs(st["varE"] / ch.varE - 1) < 1e-9
    g.close()


@pytest.mark.parametrize("name", ["bayesc_weighted", "bayesb_weighted"])
def test_weighted_replay_against_committed_golden(gpu, name):
    """Same comparison against tests/golden (does not execute the oracle)."""
    import importlib.util
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(gold, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    c = mg.WEIGHTED[name]
    gd = np.load(os.path.join(gold, name + ".npz"))
    prob = make_problem(c["n"], c["p"], c["seed"])
    g = gpu_sampler(prob, c["method"], c["v"], pi=c["pi"], est_pi=c["est_pi"])
    g.set_residual_weights(mg.weights_of(c))
    logs = [{"chi2_e": gd["chi2_e"][i], "z_mu": gd["z_mu"][i],
             "sets": [{"u": gd["u"][i], "z": gd["z"][i], "chi2_b": gd["chi2_b"][i], "beta_pi": gd["beta_pi"][i]}]}
            for i in range(c["iters"])]
    g.set_replay(logs)
    for i in range(c["iters"]):
        g.run(1)
        st = g.state()
        assert rel(st["sets"][0]["beta"], gd["beta"][i]) < TOL and abs(st["varE"] / gd["varE"][i] - 1) < TOL
        assert rel(st["sets"][0]["varBeta"], gd["varBeta"][i]) < TOL and np.array_equal(st["sets"][0]["delta"], gd["delta"][i])
    assert rel(g.state()["e"], gd["e_final"]) < TOL
    g.close()
