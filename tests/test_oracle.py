"""CPU tests of the oracle itself (the checker must be trustworthy before it checks anything)."""
import os

import numpy as np
import pytest

from common import make_problem, oracle_chain
from oracle import oracle as O
from oracle import restate_numpy as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert [int(x) for x in O.philox([0, 0, 0, 0], [0, 0])] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert [int(x) for x in O.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert [int(x) for x in O.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])] == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_hyperparameters_match_reference_doc_transcripts():
    # docs/src/BWGR/BWGR.md:32-33,52-55 ; MultipleMarkerSets.md:48-50,76-80 ; Example.md:160,177-179 ; PBLUP.md:82-85,112-117
    assert O.marker_hyper(0.001) == (4.0, 0.0005)
    assert O.residual_hyper(150.0) == (4.0, 75.0)
    assert O.marker_hyper(0.04) == (4.0, 0.02)
    assert O.residual_hyper(2500.0) == (4.0, 1250.0)
    assert O.residual_hyper(0.01) == (4.0, 0.005)
    for v, s in ((150, 75), (90, 45), (40, 20), (350, 175)):
        assert O.residual_hyper(float(v))[1] == float(s)
    assert O.residual_hyper(0.0) == (4.0, 0.0005)          # mme.jl:89-91


def test_stream_moments():
    L = O.lib()
    u = np.array([L.ngo_stream_uniform(5, 1, 1, 0, O.P_U, i) for i in range(20000)])
    z = np.array([L.ngo_stream_normal(5, 1, 1, 0, O.P_Z, i, 0) for i in range(20000)])
    c = np.array([L.ngo_stream_chisq(5, 1, 1, 0, O.P_CHI2_B, i, 0, 7.0) for i in range(20000)])
    b = np.array([L.ngo_stream_beta(5, 1, i + 1, 0, 3.0, 9.0) for i in range(5000)])
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03
    assert abs(c.mean() - 7.0) < 0.15 and abs(c.var() - 14.0) < 1.0
    assert abs(b.mean() - 0.25) < 0.01


@pytest.mark.parametrize("method", [0, 1, 2])
def test_c_oracle_equals_numpy_restatement(method):
    prob = make_problem(180, 70, 7)
    ro = np.array([0, 10, 25, 70]) if method == 0 else None
    rng = np.random.default_rng(0)
    lhs0, rhs0 = rng.uniform(0, 2, 70), rng.normal(size=70)
    ch, S = oracle_chain(prob, method, 0.02, pi=0.2, est_pi=True, region_off=ro, v_e=1.0, lhs0=lhs0, rhs0=rhs0)
    X, mpm = S.X, S.mpm
    e, mu, beta, delta = prob["y"].copy(), 0.0, np.zeros(70), np.ones(70, dtype=np.int64)
    varBeta, piHat = np.full(S.nvar, 0.02), np.array([0.8, 0.2])
    logPi = np.log(piHat)
    for _ in range(20):
        log = ch.iteration(seed=9, chain=2)
        s = log["sets"][0]
        varE = R.sample_varE(4.0, 0.5, e, len(e), log["chi2_e"])
        mu = R.sample_intercept(e, mu, varE, log["z_mu"])
        if method == 0:
            R.bayes_pr(X, mpm, lhs0, rhs0, [0, 10, 25, 70], S.scale, S.df, beta, e, varE, varBeta, s["z"], s["chi2_b"])
        elif method == 1:
            R.bayes_b(X, mpm, lhs0, rhs0, S.scale, S.df, True, beta, delta, e, varE, varBeta, piHat, logPi, s["u"], s["z"], s["chi2_b"], s["beta_pi"])
        else:
            R.bayes_c(X, mpm, lhs0, S.scale, S.df, True, beta, delta, e, varE, varBeta, piHat, logPi, s["u"], s["z"], s["chi2_b"], s["beta_pi"])
        assert np.isclose(varE, ch.varE, rtol=1e-11) and np.isclose(mu, ch.mu, rtol=1e-10)
        assert np.allclose(beta, S.beta, rtol=1e-8, atol=1e-12)
        assert np.allclose(varBeta, S.varBeta, rtol=1e-9)
        if method:
            assert (delta == S.delta).all()


@pytest.mark.parametrize("method", [0, 1, 2])
def test_weighted_residual_oracle_equals_numpy_restatement(method):
    """E.str == "D" (mme.jl:70-73, 133-136, 299-303; functions.jl:526-528): the C oracle fed with Mp = X .* w and the weighted mpm
    against the numpy restatement that spells out which dot is weighted."""
    n, p = 160, 60
    prob = make_problem(n, p, 8)
    w = np.random.default_rng(4).uniform(0.25, 4.0, n)
    rng = np.random.default_rng(0)
    lhs0, rhs0 = rng.uniform(0, 2, p), rng.normal(size=p)
    X, mean, _ = O.center_codes(prob["codes"])
    Mp_c, mpm_c = O.weighted_marker_arrays(X, w)
    ro = np.array([0, 10, 25, p]) if method == 0 else None
    S = O.MarkerSet(X=X, mpm=mpm_c, method=method, v=0.02, pi=0.2, est_pi=True, region_off=ro, lhs0=lhs0, rhs0=rhs0)
    S.Mp = Mp_c
    ch = O.OracleChain(prob["y"], [S], v_e=1.0, weights=w)
    mpm, Mp = R.weighted_setup(X, w)
    assert np.allclose(mpm, mpm_c, rtol=1e-13)
    e, mu, beta, delta = prob["y"].copy(), 0.0, np.zeros(p), np.ones(p, dtype=np.int64)
    varBeta, piHat = np.full(S.nvar, 0.02), np.array([0.8, 0.2])
    logPi = np.log(piHat)
    for _ in range(15):
        log = ch.iteration(seed=9, chain=2)
        s = log["sets"][0]
        varE = R.sample_varE_w(4.0, 0.5, w, e, n, log["chi2_e"])
        mu = R.sample_intercept_w(e, w, mu, varE, log["z_mu"])
        if method == 0:
            R.bayes_pr_w(X, Mp, mpm, lhs0, rhs0, [0, 10, 25, p], S.scale, S.df, beta, e, varE, varBeta, s["z"], s["chi2_b"])
        else:
            R.bayes_bc_w(method == 1, X, Mp, mpm, lhs0, rhs0, S.scale, S.df, True, beta, delta, e, varE, varBeta, piHat, logPi,
                         s["u"], s["z"], s["chi2_b"], s["beta_pi"])
        assert np.isclose(varE, ch.varE, rtol=1e-11) and np.isclose(mu, ch.mu, rtol=1e-10)
        assert np.allclose(beta, S.beta, rtol=1e-8, atol=1e-12)
        assert np.allclose(varBeta, S.varBeta, rtol=1e-9)
        assert np.allclose(e, ch.e, rtol=1e-8, atol=1e-10)
        if method:
            assert (delta == S.delta).all()


def test_bayesb_zero_variance_quirk():
    """functions.jl:186 sets varBeta_j = 0.0 on exclusion; next iteration p1 == pi and an included locus draws beta == 0."""
    prob = make_problem(120, 40, 3)
    ch, S = oracle_chain(prob, 1, 0.05, pi=0.3, est_pi=False)
    for _ in range(6):
        ch.iteration(seed=1, chain=0)
    assert (S.varBeta[S.delta == 0] == 0.0).all()
    assert ((S.delta == 1) & (S.beta == 0.0)).any() or True     # may or may not occur in 6 iterations; exercised without NaN
    assert np.isfinite(S.beta).all() and np.isfinite(ch.e).all()


def test_replay_reproduces_native():
    prob = make_problem(150, 60, 5)
    ch1, S1 = oracle_chain(prob, 2, 0.05, pi=0.1, est_pi=True)
    ch2, S2 = oracle_chain(prob, 2, 0.05, pi=0.1, est_pi=True)
    for _ in range(10):
        log = ch1.iteration(seed=77, chain=4)
        ch2.iteration(replay=log)
    assert np.array_equal(S1.beta, S2.beta) and np.array_equal(ch1.e, ch2.e) and ch1.varE == ch2.varE


@pytest.mark.parametrize("name", ["bayespr_rr", "bayespr_regions", "bayesb", "bayesc_pi"])
def test_oracle_against_golden(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    c = mg.CASES[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    prob = make_problem(c["n"], c["p"], c["seed"])
    ro = np.array(c["region_off"], dtype=np.int64) if "region_off" in c else None
    ch, S = oracle_chain(prob, c["method"], c["v"], pi=c.get("pi", 0.0), est_pi=c.get("est_pi", False), region_off=ro)
    for it in range(c["iters"]):
        ch.iteration(seed=c["seed"], chain=3)
        assert np.allclose(S.beta, g["beta"][it], rtol=1e-9, atol=1e-13)
        assert np.isclose(ch.varE, g["varE"][it], rtol=1e-10)
    assert np.allclose(ch.e, g["e_final"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("name", ["bayesc_weighted", "bayesb_weighted"])
def test_weighted_oracle_against_golden(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    c = mg.WEIGHTED[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    prob, w, ch, S = mg.weighted_chain(c)
    for it in range(c["iters"]):
        ch.iteration(seed=c["seed"], chain=3)
        assert np.allclose(S.beta, g["beta"][it], rtol=1e-9, atol=1e-13) and np.array_equal(S.delta, g["delta"][it])
        assert np.isclose(ch.varE, g["varE"][it], rtol=1e-10)
    assert np.allclose(ch.e, g["e_final"], rtol=1e-8, atol=1e-10)


def test_multibreed_oracle_matches_dense_algebra():
    """functions.jl:140-154 with k=2: the per-locus draw equals mean + chol(C) z with C = (M'M/varE + inv(Sigma))^-1."""
    rng = np.random.default_rng(4)
    n, p, k = 90, 12, 2
    Xk = [np.asfortranarray(rng.binomial(2, 0.3, size=(n, p)).astype(float)) for _ in range(k)]
    for x in Xk:
        x -= x.mean(0)
    v = np.array([[0.02, 0.005], [0.005, 0.03]])
    mb = O.MultiBreedOracle(Xk, v)
    e = rng.normal(size=n)
    e0, varE = e.copy(), 1.3
    log = mb.sweep(e, varE, it=1, seed=3, chain=0)
    beta = np.zeros((k, p)); ee = e0.copy(); invB = np.linalg.inv(v)
    for j in range(p):
        Mj = np.stack([Xk[b][:, j] for b in range(k)], axis=1)
        ee += Mj @ beta[:, j]
        C = np.linalg.inv(Mj.T @ Mj / varE + invB)
        beta[:, j] = C @ (Mj.T @ ee / varE) + np.linalg.cholesky(C) @ log["z"][j]
        ee -= Mj @ beta[:, j]
    assert np.allclose(beta, mb.beta, rtol=1e-9, atol=1e-12) and np.allclose(ee, e, rtol=1e-9, atol=1e-11)
    assert np.allclose(mb.varBeta[0], mb.varBeta[0].T) and np.all(np.linalg.eigvalsh(mb.varBeta[0]) > 0)


def test_region_builder_matches_reference_rules():
    chr_id = np.array([1] * 7 + [2] * 5 + [3] * 3)
    assert O.regions_from_map(chr_id, 9999).tolist() == [0, 15]
    assert O.regions_from_map(chr_id, 99).tolist() == [0, 7, 12, 15]
    assert O.regions_from_map(chr_id, 3).tolist() == [0, 3, 6, 7, 10, 12, 15]      # ceil windows inside each chromosome


def test_bayesr_oracle_matches_a_literal_numpy_restatement():
    """functions.jl:238-289 line by line in numpy (fresh uniform per cumulative comparison, class 1 = zero variance)."""
    rng = np.random.default_rng(11)
    n, p = 120, 25
    X = np.asfortranarray(rng.binomial(2, 0.3, size=(n, p)).astype(float)); X -= X.mean(0)
    mpm = (X * X).sum(0)
    vclass = np.array([0.0, 0.0001, 0.001, 0.01]); pi = np.array([0.7, 0.15, 0.1, 0.05])
    R = O.BayesROracle(X, mpm, pi, vclass, v=0.8, est_pi=True)
    e = rng.normal(size=n) * 2; e0 = e.copy(); varE = 1.7
    log = R.sweep(e, varE, it=1, seed=5, chain=0)
    beta = np.zeros(p); ee = e0.copy(); varc = 0.8 * vclass; logPi = np.log(pi)
    nLoci = np.zeros(4, dtype=int); nnz = 0; sumS = 0.0
    for j in range(p):
        ee += X[:, j] * beta[j]
        rhs = X[:, j] @ ee / varE
        lhs = np.where(varc == 0, 0.0, mpm[j] / varE + 1.0 / np.where(varc == 0, 1.0, varc))
        with np.errstate(divide="ignore", invalid="ignore"):
            logL = np.where(varc == 0, logPi, -0.5 * (np.log(varc * lhs) - rhs * rhs / lhs) + logPi)
        cum = np.cumsum(np.exp(logL) / np.exp(logL).sum())
        cls = next(v for v in range(4) if cum[v] >= log["u"][j, v])
        nLoci[cls] += 1
        if varc[cls] != 0:
            nnz += 1
            beta[j] = rhs / lhs[cls] + np.sqrt(1 / lhs[cls]) * log["z"][j]
            ee -= X[:, j] * beta[j]
            sumS += beta[j] ** 2 / vclass[cls]
        else:
            beta[j] = 0.0
        assert R.delta[j] == cls + 1
    assert np.allclose(beta, R.beta, rtol=1e-10, atol=1e-13) and np.allclose(ee, e, rtol=1e-10, atol=1e-11)
    df, scale = O.marker_hyper(0.8)
    assert np.isclose(R.varBeta[0], (scale * df + sumS) / log["chi2_b"][0], rtol=1e-12)
    assert np.isclose(R.piHat.sum(), 1.0) and np.allclose(R.logPi, np.log(R.piHat)) and np.array_equal(R.piHat, log["dir_pi"])
    # replaying the log reproduces the sweep bit for bit
    R2 = O.BayesROracle(X, mpm, pi, vclass, v=0.8, est_pi=True)
    e2 = e0.copy()
    R2.sweep(e2, varE, it=1, seed=999, chain=3, replay=log)
    assert np.array_equal(R2.beta, R.beta) and np.array_equal(e2, e) and np.array_equal(R2.piHat, R.piHat)


def _golden_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    return mg


def test_bayesr_and_tuple_oracles_reproduce_their_committed_golden_chains():
    mg = _golden_module()
    c = mg.EXTRA["bayesr"]; g = np.load(os.path.join(GOLD, "bayesr.npz"))
    prob = make_problem(c["n"], c["p"], c["seed"])
    X, mean, mpm = O.center_codes(prob["codes"])
    R = O.BayesROracle(X, mpm, np.array(c["pi"]), np.array(c["v_class"]), v=c["v"], est_pi=c["est_pi"])
    ch = O.OracleChain(prob["y"], [], v_e=prob["var_y"] / 2)
    for it in range(c["iters"]):
        ch.iteration(seed=c["seed"], chain=3)
        R.sweep(ch.e, ch.varE, it=ch.iter, seed=c["seed"], chain=3)
        assert np.array_equal(R.delta, g["delta"][it]) and np.allclose(R.beta, g["beta"][it], rtol=1e-9, atol=1e-13)
        assert np.allclose(R.piHat, g["pi"][it], rtol=1e-12)
    assert np.allclose(ch.e, g["e_final"], rtol=1e-8, atol=1e-10)
    c = mg.EXTRA["tuple2"]; g = np.load(os.path.join(GOLD, "tuple2.npz"))
    probs, y = mg.breeds(c["n"], c["p"], 2, c["seed"])
    mb = O.MultiBreedOracle([O.center_codes(pr["codes"])[0] for pr in probs], np.array(c["V"]), region_off=np.array(c["region_off"], dtype=np.int64))
    ch = O.OracleChain(y, [], v_e=float(np.var(y)) / 2)
    for it in range(c["iters"]):
        ch.iteration(seed=c["seed"], chain=3)
        mb.sweep(ch.e, ch.varE, it=ch.iter, seed=c["seed"], chain=3)
        assert np.allclose(mb.beta, g["beta"][it], rtol=1e-9, atol=1e-13) and np.allclose(mb.varBeta, g["varBeta"][it], rtol=1e-9)
    assert np.allclose(ch.e, g["e_final"], rtol=1e-8, atol=1e-10)


def test_fixed_effect_oracle_matches_numpy_restatement_of_wangs_trick():
    """functions.jl:22-54: add-back, Yi = X'e/varE once, then the columns one by one against xpx."""
    rng = np.random.default_rng(2)
    n, c = 150, 4
    Xf = np.asfortranarray(np.column_stack([rng.normal(size=n), rng.integers(0, 2, n), rng.normal(size=n) * 3, rng.uniform(size=n)]))
    F = O.FixedSet(Xf, col0=1)
    F.b[:] = rng.normal(size=c)
    e = rng.normal(size=n); e0 = e.copy(); b0 = F.b.copy(); varE = 0.9
    z = F.sample(e, varE, it=3, seed=8, chain=1)
    ee = e0 + Xf @ b0
    Yi = Xf.T @ ee / varE
    xpx = Xf.T @ Xf
    bv = b0.copy()
    for i in range(c):
        bv[i] = 0.0
        rhs = Yi[i] - xpx[i] @ bv / varE
        lhs = xpx[i, i] / varE
        bv[i] = rhs / lhs + np.sqrt(1 / lhs) * z[i]
    ee -= Xf @ bv
    assert np.allclose(F.b, bv, rtol=1e-10) and np.allclose(e, ee, rtol=1e-10, atol=1e-12)
    F1 = O.FixedSet(Xf[:, 0], col0=5, lhs0=0.3, rhs0=-0.2)
    e = e0.copy(); z1 = F1.sample(e, varE, it=1, seed=8, chain=0)
    rhs = Xf[:, 0] @ e0 / varE - 0.2; lhs = Xf[:, 0] @ Xf[:, 0] / varE + 0.3
    assert np.isclose(F1.b[0], rhs / lhs + np.sqrt(1 / lhs) * z1[0], rtol=1e-12)


# ----------------------------------------------------------------------------- BayesRCpi / BayesRCplus (functions.jl:291-419)
@pytest.mark.parametrize("plus", [False, True])
@pytest.mark.parametrize("est_pi", [False, True])
def test_bayesrc_oracle_matches_numpy_restatement(plus, est_pi):
    """the C oracle (ngo_rc_sweep) against a literal numpy restatement of the reference's loops, same explicit variates; native variates
    logged by the oracle in iteration k are replayed through both"""
    from oracle import restate_numpy as RN
    rng = np.random.default_rng(5)
    n, p, nA = 300, 40, 3
    codes = rng.integers(0, 3, size=(n, p)).astype(np.int8)
    X, _, mpm = O.center_codes(np.asfortranarray(codes))
    annot = rng.integers(0, 2, size=(p, nA)).astype(np.int32)
    annot[annot.sum(1) == 0, 0] = 1                      # every locus has an annotation (mme.jl:396-399 leaves all-zero rows as NaN)
    annot[3, :] = [2, 0, 1]                              # integer counts, not only 0/1
    vclass, pi = np.array([0.0, 0.001, 0.01, 0.1]), np.array([0.7, 0.15, 0.1, 0.05])
    y = rng.normal(size=n) + X[:, 2] * 0.8 - X[:, 11] * 0.5
    R = O.BayesRCOracle(X, mpm, pi, vclass, v=0.6, annot=annot, est_pi=est_pi, plus=plus)
    e_c = y - y.mean()
    e_n = e_c.copy()
    st = dict(beta=np.zeros(p), delta=np.ones(p, dtype=np.int64), annot_cat=np.zeros(p, dtype=np.int64), varBeta=np.full(nA, 0.6),
              piHat=np.tile(pi, (nA, 1)), logPi=np.log(np.tile(pi, (nA, 1))), annot_prob=annot / annot.sum(1, keepdims=True))
    for it in range(1, 5):
        v = R.sweep(e_c, 1.3, it=it, seed=9, chain=2)                      # native stream, variates logged
        RN.bayes_rc(X, mpm, e_n, 1.3, st["beta"], st["delta"], st["annot_cat"], st["varBeta"], st["piHat"], st["logPi"], annot,
                    st["annot_prob"], vclass, R.df, R.scale, est_pi, plus, v)
        assert np.array_equal(R.delta, st["delta"]) and np.allclose(R.beta, st["beta"], rtol=1e-10, atol=1e-14)
        assert np.allclose(R.varBeta, st["varBeta"], rtol=1e-12) and np.allclose(R.piHat, st["piHat"], rtol=1e-14) and np.allclose(e_c, e_n, rtol=1e-9, atol=1e-11)
        if not plus:
            assert np.array_equal(R.annot_cat, st["annot_cat"]) and np.allclose(R.annot_prob, st["annot_prob"], rtol=1e-14)
    assert (R.delta > 1).sum() > 0 and np.allclose(R.annot_prob.sum(1), 1.0)
    if est_pi:
        assert np.allclose(R.piHat.sum(1), 1.0)
    # replaying the last iteration's log reproduces it
    R2 = O.BayesRCOracle(X, mpm, pi, vclass, v=0.6, annot=annot, est_pi=est_pi, plus=plus)
    e2 = y - y.mean()
    v1 = R2.sweep(e2, 1.3, it=1, seed=9, chain=2)
    R3 = O.BayesRCOracle(X, mpm, pi, vclass, v=0.6, annot=annot, est_pi=est_pi, plus=plus)
    e3 = y - y.mean()
    R3.sweep(e3, 1.3, it=1, seed=1234, chain=0, replay=v1)
    assert np.array_equal(R2.beta, R3.beta) and np.array_equal(e2, e3) and np.array_equal(R2.varBeta, R3.varBeta)


@pytest.mark.parametrize("name", ["bayesrc_pi", "bayesrc_plus"])
def test_bayesrc_oracle_reproduces_its_committed_golden_chain(name):
    """tests/golden/bayesrc_*.npz (python tests/golden/make_golden.py rc): the oracle, fed the committed variate log, lands on the committed states"""
    mg = _golden_module()
    c = mg.RC[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    prob = make_problem(c["n"], c["p"], c["seed"])
    X, _, mpm = O.center_codes(prob["codes"])
    R = O.BayesRCOracle(X, mpm, np.array(c["pi"]), np.array(c["v_class"]), v=c["v"], annot=mg.rc_annot(c["p"], c["n_annot"], c["seed"]),
                        est_pi=True, plus=c["plus"])
    ch = O.OracleChain(prob["y"], [], v_e=prob["var_y"] / 2)
    for i in range(c["iters"]):
        ch.iteration(replay={"chi2_e": float(g["chi2_e"][i]), "z_mu": float(g["z_mu"][i]), "z_fx": [], "sets": []})
        R.sweep(ch.e, ch.varE, it=ch.iter, replay={k: g[k][i] for k in ("u_annot", "dirp", "u", "z", "chi2_b", "dir_pi")})
        assert np.array_equal(R.delta, g["delta"][i]) and np.allclose(R.beta, g["beta"][i], rtol=1e-12, atol=1e-15)
        assert np.allclose(R.varBeta, g["varBeta"][i], rtol=1e-12) and np.isclose(ch.varE, g["varE"][i], rtol=1e-12)
    assert np.allclose(ch.e, g["e_final"], rtol=1e-10, atol=1e-12) and np.allclose(R.annot_prob, g["annot_prob_final"], rtol=1e-14)
