"""GPU parity of fixed effects besides the intercept: sampleX! / sampleb! of functions.jl:22-54 inside ngp_run,
against the CPU oracle (ngo_sample_fixed): native streams, replay, single columns with prior information, several sets."""
import numpy as np
import pytest

import nextgp.jl_b200 as ngp
from common import make_problem, rel
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _covariates(n, seed):
    rng = np.random.default_rng(seed)
    age = rng.normal(size=n) * 2 + 5
    herd = np.eye(3)[rng.integers(0, 3, n)][:, 1:]                 # a 3-level factor with the first level dropped: 2 columns
    cov2 = np.column_stack([rng.uniform(size=n), rng.normal(size=n), (rng.uniform(size=n) > 0.5).astype(float)])
    return age, herd, cov2


def _pair(prob, fixed_specs, method, kw, kernel="blocked", **geom):
    n = len(prob["y"])
    y = prob["y"] + 0.8 * fixed_specs[0][0].reshape(n, -1)[:, 0]
    fixed_o, col = [], 1
    for data, l0, r0 in fixed_specs:
        F = O.FixedSet(data, col0=col, lhs0=l0, rhs0=r0)
        col += F.c
        fixed_o.append(F)
    X, mean, mpm = O.center_codes(prob["codes"])
    S = O.MarkerSet(X=X, mpm=mpm, method=method, **kw)
    ch = O.OracleChain(y, [S], v_e=prob["var_y"] / 2, intercept=True, fixed=fixed_o)
    g = ngp.Sampler(0, kernel=kernel, **geom)
    g.upload_genotypes(0, prob["codes"])
    df, scale = O.marker_hyper(kw["v"])
    g.set_prior(0, method, df, scale, kw["v"], pi_in=kw.get("pi", 0.0), est_pi=kw.get("est_pi", False))
    g.set_fixed_effects([(d, l0, r0) for d, l0, r0 in fixed_specs])
    g.set_phenotype(y)
    g.set_residual_prior(*O.residual_hyper(prob["var_y"] / 2))
    g.set_intercept(True)
    return ch, S, fixed_o, g


@pytest.mark.parametrize("kernel", ["blocked", "literal"])
@pytest.mark.parametrize("method,kw", [(2, dict(v=0.05, pi=0.1, est_pi=True)), (0, dict(v=0.01))])
def test_fixed_effects_native_chain_matches_oracle(gpu, kernel, method, kw):
    prob = make_problem(777, 130, 41)
    age, herd, cov2 = _covariates(777, 1)
    ch, S, fx, g = _pair(prob, [(age, 0.0, 0.0), (herd, 0.0, 0.0), (cov2, 0.0, 0.0)], method, kw, kernel=kernel)
    g.set_rng(12, 2)
    for _ in range(6):
        ch.iteration(seed=12, chain=2)
    g.run(4); g.run(2)
    st = g.state()
    b_o = np.concatenate([F.b for F in fx])
    assert rel(g.fixed_effects(), b_o) < 1e-8 and abs(st["mu"] - ch.mu) < 1e-8 * max(1.0, abs(ch.mu))
    assert rel(st["sets"][0]["beta"], S.beta) < 1e-8 and rel(st["e"], ch.e) < 1e-8 and abs(st["varE"] / ch.varE - 1) < 1e-9
    g.close()


def test_fixed_effects_replay_and_single_column_prior(gpu):
    prob = make_problem(500, 90, 43)
    age, herd, cov2 = _covariates(500, 2)
    specs = [(age, 0.7, -0.3), (herd, 0.0, 0.0)]
    kw = dict(v=0.05, pi=0.2, est_pi=True)
    ch, S, fx, g = _pair(prob, specs, 2, kw, min_rows=32)
    logs = [ch.iteration(seed=5, chain=0) for _ in range(4)]
    g.set_rng(99, 9)
    g.set_replay(logs)
    g.set_fixed_replay(logs)
    g.run(4)
    st = g.state()
    assert rel(g.fixed_effects(), np.concatenate([F.b for F in fx])) < 1e-8
    assert np.array_equal(st["sets"][0]["delta"], S.delta) and rel(st["sets"][0]["beta"], S.beta) < 1e-8 and rel(st["e"], ch.e) < 1e-8
    g.close()


def test_fixed_effects_argument_checks(gpu):
    prob = make_problem(120, 20, 4)
    g = ngp.Sampler(0)
    with pytest.raises(ngp.NgpError):
        g.set_fixed_effects([np.ones(120)])                       # before the first upload
    g.upload_genotypes(0, prob["codes"])
    with pytest.raises(ngp.NgpError):
        g.set_fixed_effects([np.ones((120, 33))])                 # more than 32 columns
    g.set_fixed_effects([np.arange(120.0)])
    g.set_fixed_effects([])                                       # removing them again is allowed
    g.close()


def test_runLMEM_with_a_covariate_and_a_factor(gpu, tmp_path):
    import os
    prob = make_problem(240, 40, 6)
    rng = np.random.default_rng(0)
    age = rng.normal(size=240) + 4
    herd = np.array(["a", "b", "c"])[rng.integers(0, 3, 240)]
    y = prob["y"] + 0.5 * age + (herd == "c") * 1.5
    out = str(tmp_path / "outMCMC")
    VCV = {"M": ngp.BayesC(0.1, 0.05, estimatePi=True), "e": ngp.Random("I", float(np.var(y)) / 2)}
    s = ngp.runLMEM("y ~ 1 + age + herd + SNP(M,x)", {"y": y, "age": age, "herd": herd}, 400, 100, 10, outFolder=out, VCV=VCV, seed=2,
                    matrices={"M": prob["codes"]})
    assert open(os.path.join(out, "bOut")).readline().strip().split("\t") == ["(Intercept)", "age", "herd: b", "herd: c"]
    b = np.loadtxt(os.path.join(out, "bOut"), delimiter="\t", skiprows=1)
    assert b.shape == (30, 4)
    assert abs(b[:, 1].mean() - 0.5) < 0.25 and abs(b[:, 3].mean() - 1.5) < 0.8      # the planted effects are recovered
    s.close()
