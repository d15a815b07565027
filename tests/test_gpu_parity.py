"""GPU parity tests proper: everything goes through the C ABI (ctypes -> libngp.so -> sm_100a kernels) and is
compared with the CPU oracle on the same seeded inputs, and with the committed golden fixtures.
Tolerance (BASELINE.json north_star): replayed draws must reproduce per-iteration effects and variances to a
relative 1e-5 in fp64; the checks below use 1e-8 (observed ~1e-12).  Packing must be bit-exact."""
import os

import numpy as np
import pytest

import nextgp.jl_b200 as ngp
from common import METHODS, gpu_sampler, make_problem, oracle_chain, rel
from nextgp.jl_b200 import _lib as L
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-8


# ----------------------------------------------------------------------------- codec / set-up
@pytest.mark.parametrize("n,p", [(1, 1), (7, 3), (64, 64), (1000, 129), (4097, 70)])
def test_pack_unpack_bit_exact(gpu, n, p):
    rng = np.random.default_rng(n + p)
    codes = np.asfortranarray(rng.integers(0, 3, size=(n, p)).astype(np.int8))
    s = ngp.Sampler(0)
    s.upload_genotypes(0, codes)
    assert np.array_equal(s.download_genotypes(0), codes)
    # same data through the Float64 and 2-bit host formats
    s.upload_genotypes(1, codes.astype(np.float64))
    assert np.array_equal(s.download_genotypes(1), codes)
    ldp = (n + 3) // 4
    packed = np.zeros((ldp, p), dtype=np.uint8, order="F")
    assert L.lib().ngp_pack2(codes.ctypes.data, n, p, n, packed.ctypes.data, ldp) == 0
    s.upload_genotypes(2, packed, fmt=L.GENO_PACKED2, n=n)
    assert np.array_equal(s.download_genotypes(2), codes)
    mean, mpm = s.column_stats(0)
    X = codes.astype(np.float64) - codes.mean(0)
    assert np.allclose(mean, codes.mean(0), rtol=1e-14) and np.allclose(mpm, (X * X).sum(0), rtol=1e-11, atol=1e-9)
    s.close()


def test_upload_rejects_missing_and_dosages(gpu):
    s = ngp.Sampler(0)
    bad = np.asfortranarray(np.array([[0, 1], [3, 2], [1, 1]], dtype=np.int8))
    with pytest.raises(ngp.NgpError) as ei:
        s.upload_genotypes(0, bad)
    assert ei.value.code == L.EDATA
    with pytest.raises(ngp.NgpError):
        s.upload_genotypes(0, np.array([[0.0, 0.5], [1.0, 2.0]]))
    with pytest.raises(ngp.NgpError):
        s.run(1)                       # nothing uploaded: loud error, no fallback
    s.close()


def test_device_synth_matches_spec_bit_exact(gpu):
    n, p, seed = 1001, 77, 99
    pr = ngp.synth.problem(n, p, seed)
    s = ngp.Sampler(0)
    s.synth_genotypes(0, n, p, seed, pr["thr0"], pr["thr1"])
    dev = s.download_genotypes(0)
    assert np.array_equal(dev, O.synth_codes(seed, n, 0, p, pr["thr0"], pr["thr1"]))
    s.close()


def test_device_stream_matches_oracle_stream(gpu):
    s = ngp.Sampler(0)
    s.set_rng(0x1234567890ABCDEF, 5)
    Lo = O.lib()
    u = s.debug_variates(2, 7, O.P_U, 0.0, 500)
    z = s.debug_variates(2, 7, O.P_Z, 0.0, 500)
    c = s.debug_variates(2, 7, O.P_CHI2_B, 5.0, 500)
    cl = s.debug_variates(2, 7, O.P_CHI2_B, 50004.0, 200)
    uo = np.array([Lo.ngo_stream_uniform(0x1234567890ABCDEF, 5, 7, 2, O.P_U, i) for i in range(500)])
    zo = np.array([Lo.ngo_stream_normal(0x1234567890ABCDEF, 5, 7, 2, O.P_Z, i, 0) for i in range(500)])
    co = np.array([Lo.ngo_stream_chisq(0x1234567890ABCDEF, 5, 7, 2, O.P_CHI2_B, i, 0, 5.0) for i in range(500)])
    clo = np.array([Lo.ngo_stream_chisq(0x1234567890ABCDEF, 5, 7, 2, O.P_CHI2_B, i, 0, 50004.0) for i in range(200)])
    assert np.array_equal(u, uo)                       # integer -> fp64, exact
    assert np.allclose(z, zo, rtol=1e-12, atol=1e-14)  # libm vs CUDA log/cos/sqrt
    assert np.allclose(c, co, rtol=1e-11) and np.allclose(cl, clo, rtol=1e-11)
    s.close()


# ----------------------------------------------------------------------------- replay parity
CASES = [
    ("BayesPR-RR", 0, dict(v=0.01)),
    ("BayesPR-regions", 0, dict(v=0.01, region_off=[0, 17, 18, 90, 200, 333])),
    ("BayesPR-perlocus", 0, dict(v=0.01, region_off="each")),
    ("BayesB", 1, dict(v=0.05, pi=0.1, est_pi=True)),
    ("BayesB-fixedpi", 1, dict(v=0.05, pi=0.3, est_pi=False)),
    ("BayesC-pi", 2, dict(v=0.05, pi=0.05, est_pi=True)),
    ("BayesC-fixedpi", 2, dict(v=0.05, pi=0.5, est_pi=False)),
]


def _run_replay(prob, method, kw, kernel, iters, block=0, min_rows=0, intercept=True, lhs0=None, rhs0=None, max_ctas=0, **geom):
    kw = dict(kw)
    p = prob["codes"].shape[1]
    ro = kw.pop("region_off", None)
    if isinstance(ro, str):
        ro = np.arange(p + 1)
    elif ro is not None:
        ro = np.array(ro)
    ch, S = oracle_chain(prob, method, region_off=ro, intercept=intercept, lhs0=lhs0, rhs0=rhs0, **kw)
    logs, snaps = [], []
    for _ in range(iters):
        logs.append(ch.iteration(seed=42, chain=1))
        snaps.append(ch.snapshot())
    g = gpu_sampler(prob, method, region_off=ro, kernel=kernel, block=block, min_rows=min_rows, intercept=intercept,
                    lhs0=lhs0, rhs0=rhs0, max_ctas=max_ctas, **geom, **kw)
    g.set_replay(logs)
    worst = 0.0
    for it in range(iters):
        g.run(1)
        st = g.state()
        o = snaps[it]
        worst = max(worst, rel(st["sets"][0]["beta"], o["sets"][0]["beta"]), abs(st["varE"] / o["varE"] - 1),
                    abs(st["mu"] - o["mu"]) / max(abs(o["mu"]), 1e-300) if intercept else 0.0,
                    rel(st["sets"][0]["varBeta"], o["sets"][0]["varBeta"]), rel(st["e"], o["e"]))
        assert np.array_equal(st["sets"][0]["delta"], o["sets"][0]["delta"]), f"indicator mismatch at iteration {it + 1}"
        if method != 0:
            worst = max(worst, rel(st["sets"][0]["piHat"], o["sets"][0]["piHat"]))
        assert worst < TOL, f"iteration {it + 1}: rel diff {worst}"
    g.close()
    return worst


@pytest.mark.parametrize("kernel", ["blocked", "literal"])
@pytest.mark.parametrize("name,method,kw", CASES, ids=[c[0] for c in CASES])
def test_replay_parity(gpu, kernel, name, method, kw):
    prob = make_problem(640, 333, 17)
    _run_replay(prob, method, kw, kernel, iters=12)


@pytest.mark.parametrize("n,p,block,min_rows", [(1, 5, 0, 0), (9, 1, 0, 0), (100, 64, 32, 8), (257, 65, 64, 8), (1500, 130, 32, 8),
                                                (3000, 70, 64, 16)])
def test_replay_parity_ragged_shapes(gpu, n, p, block, min_rows):
    """edge shapes: single row / single marker / p not a multiple of the block / many small panels (more CTAs)."""
    prob = make_problem(n, p, 5 + n)
    _run_replay(prob, 2, dict(v=0.05, pi=0.3, est_pi=True), "blocked", iters=6, block=block, min_rows=min_rows)
    _run_replay(prob, 0, dict(v=0.02), "blocked", iters=4, block=block, min_rows=min_rows)


@pytest.mark.parametrize("geom", [dict(lookahead=1, near=1, block=64), dict(lookahead=3, tile_stages=5, versions=2), dict(lookahead=24, block=16),
                                  dict(lookahead=9, versions=3, near=6), dict(lookahead=6, near=5, block=64),
                                  dict(profile=True), dict(refetch=1, lookahead=13), dict(refetch=1, lookahead=20, tile_stages=4, block=16),
                                  dict(refetch=1, lookahead=5, tile_stages=1, block=64)],
                         ids=["D1-B64", "D3-2versions", "D24-B16", "D9-near6", "D6-near5-B64", "instrumented", "refetch-D13", "refetch-D20-4stages-B16",
                              "refetch-1stage-B64"])
def test_replay_parity_over_pipeline_geometries(gpu, geom):
    """The look-ahead depth, ring sizes (incl. a 2- or 4-stage record ring), residual-version count and the instrumented variant change the schedule, never the result."""
    prob = make_problem(1300, 700, 77)
    geom = dict(geom)
    block = geom.pop("block", 0)
    _run_replay(prob, 2, dict(v=0.05, pi=0.2, est_pi=True), "blocked", iters=5, block=block, min_rows=8, **geom)
    _run_replay(prob, 0, dict(v=0.02), "blocked", iters=3, block=block, min_rows=8, **geom)


@pytest.mark.parametrize("first,then", [(0, 2), (2, 0)])
def test_dense_sets_get_the_short_lookahead_and_back(gpu, first, then):
    """All marker sets BayesPR (every effect changes in every sweep): the rings are re-sized to look-ahead 3 / near 3 and the banded Gram is
    rebuilt; a spike-and-slab prior brings the default geometry back.  Either way the chain is the oracle's."""
    prob = make_problem(1300, 700, 5)
    kws = {0: dict(v=0.02), 2: dict(v=0.05, pi=0.2, est_pi=True)}
    g = gpu_sampler(prob, first, **kws[first])
    t1 = g.timing()
    df, scale = O.marker_hyper(kws[then]["v"])
    g.set_prior(0, then, df, scale, kws[then]["v"], pi_in=kws[then].get("pi", 0.0), est_pi=kws[then].get("est_pi", False))
    t2 = g.timing()
    dense, default = (t1, t2) if first == 0 else (t2, t1)
    assert dense["lookahead"] == 3 and dense["near_depth"] == 3 and default["lookahead"] > 3
    assert dense["block"] == default["block"] and dense["rows_per_cta"] == default["rows_per_cta"] and dense["ctas"] == default["ctas"]
    ch, S = oracle_chain(prob, then, **kws[then])
    g.set_rng(21, 2)
    for _ in range(4):
        ch.iteration(seed=21, chain=2)
    g.run(4)
    st = g.state()
    assert rel(st["sets"][0]["beta"], S.beta) < TOL and rel(st["e"], ch.e) < TOL and abs(st["varE"] / ch.varE - 1) < TOL
    assert np.array_equal(st["sets"][0]["delta"], S.delta)
    g.close()


def test_replay_parity_many_rows_per_cta(gpu):
    """Few CTAs => > 512 rows per worker CTA => 4 row groups per updater thread (the C5 / C3 geometry), for every block size."""
    prob = make_problem(3001, 150, 78)
    for method, kw in ((2, dict(v=0.05, pi=0.3, est_pi=True)), (1, dict(v=0.05, pi=0.3, est_pi=True)), (0, dict(v=0.02))):
        _run_replay(prob, method, kw, "blocked", iters=4, max_ctas=3)
    g = gpu_sampler(prob, 2, 0.05, pi=0.3, est_pi=True, max_ctas=3)
    t = g.timing()
    assert t["block"] == 64 and t["rows_per_cta"] > 1024 and t["ctas"] == 3 and t["tile_stages"] == 1 and t["refetch"] == 1     # one tile, shared by the 8 dot warps
    g.close()
    _run_replay(prob, 2, dict(v=0.05, pi=0.3, est_pi=True), "literal", iters=3, max_ctas=3)          # the per-marker sweep takes any panel with any block size
    _run_replay(prob, 2, dict(v=0.05, pi=0.3, est_pi=True), "blocked", iters=4, max_ctas=3, block=16)
    g = gpu_sampler(prob, 2, 0.05, pi=0.3, est_pi=True, max_ctas=3, block=16)
    assert g.timing()["block"] == 16
    g.close()
    for block in (32, 64):
        _run_replay(prob, 2, dict(v=0.05, pi=0.3, est_pi=True), "blocked", iters=4, max_ctas=3, block=block)
        _run_replay(prob, 0, dict(v=0.02), "blocked", iters=3, max_ctas=3, block=block)
    big = make_problem(8300, 40, 79)
    with pytest.raises(ngp.NgpError) as ei:        # more rows than the updater warps hold in registers
        gpu_sampler(big, 2, 0.05, pi=0.3, est_pi=True, max_ctas=3)
    assert ei.value.code == L.EUNSUPPORTED


def test_profile_and_trace_of_instrumented_kernel(gpu):
    prob = make_problem(900, 640, 79)
    g = gpu_sampler(prob, 2, 0.05, pi=0.1, est_pi=True, profile=True)
    g.set_rng(3, 0)
    g.run(2)
    pr, tr = g.profile(), g.trace()
    t = g.timing()
    assert pr.shape == (t["ctas"], 32) and (pr[-1, 4] > 0) and (pr[:-1, 2] > 0).all()      # chain warp and every updater warp ticked
    assert (tr[: 640 // 64, 1] > 0).all() and t["lookahead"] >= 1 and t["record_stages"] in (2, 4, 8)
    g.close()


def test_replay_parity_no_intercept_and_summary_stat_priors(gpu):
    prob = make_problem(400, 150, 8)
    rng = np.random.default_rng(1)
    lhs0, rhs0 = rng.uniform(0, 3, 150), rng.normal(size=150)
    for method, kw in ((0, dict(v=0.02)), (1, dict(v=0.05, pi=0.2, est_pi=True)), (2, dict(v=0.05, pi=0.2, est_pi=True))):
        _run_replay(prob, method, kw, "blocked", iters=8, intercept=False, lhs0=lhs0, rhs0=rhs0)


@pytest.mark.parametrize("name", ["bayespr_rr", "bayespr_regions", "bayesb", "bayesc_pi"])
def test_replay_against_committed_golden(gpu, name):
    """Same comparison against tests/golden (does not execute the oracle)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    c = mg.CASES[name]
    gd = np.load(os.path.join(GOLD, name + ".npz"))
    prob = make_problem(c["n"], c["p"], c["seed"])
    ro = np.array(c["region_off"], dtype=np.int64) if "region_off" in c else None
    g = gpu_sampler(prob, c["method"], c["v"], pi=c.get("pi", 0.0), est_pi=c.get("est_pi", False), region_off=ro)
    logs = [{"chi2_e": gd["chi2_e"][i], "z_mu": gd["z_mu"][i],
             "sets": [{"u": gd["u"][i], "z": gd["z"][i], "chi2_b": gd["chi2_b"][i], "beta_pi": gd["beta_pi"][i]}]}
            for i in range(c["iters"])]
    g.set_replay(logs)
    for i in range(c["iters"]):
        g.run(1)
        st = g.state()
        assert rel(st["sets"][0]["beta"], gd["beta"][i]) < TOL and abs(st["varE"] / gd["varE"][i] - 1) < TOL
        assert rel(st["sets"][0]["varBeta"], gd["varBeta"][i]) < TOL
        if c["method"]:
            assert np.array_equal(st["sets"][0]["delta"], gd["delta"][i])
    assert rel(g.state()["e"], gd["e_final"]) < TOL
    g.close()


# ----------------------------------------------------------------------------- native stream
@pytest.mark.parametrize("method,kw", [(0, dict(v=0.01)), (1, dict(v=0.05, pi=0.1, est_pi=True)), (2, dict(v=0.05, pi=0.05, est_pi=True))])
def test_native_philox_chain_matches_oracle_native_chain(gpu, method, kw):
    """Same seed, device-generated variates: the two Philox implementations + transforms agree, so the chains agree."""
    prob = make_problem(500, 300, 23)
    ch, S = oracle_chain(prob, method, **kw)
    g = gpu_sampler(prob, method, **kw)
    g.set_rng(20261018, 3)
    for it in range(10):
        ch.iteration(seed=20261018, chain=3)
    g.run(10)                                   # ten iterations inside ONE persistent launch
    st = g.state()
    assert rel(st["sets"][0]["beta"], S.beta) < 1e-7
    assert abs(st["varE"] / ch.varE - 1) < 1e-7 and rel(st["e"], ch.e) < 1e-7
    assert np.array_equal(st["sets"][0]["delta"], S.delta) and st["iter"] == 10
    g.close()


def test_run_is_reproducible_and_batching_invariant(gpu):
    prob = make_problem(700, 200, 31)
    outs = []
    for batches in ((6,), (1, 2, 3)):
        g = gpu_sampler(prob, 2, 0.05, pi=0.05, est_pi=True)
        g.set_rng(7, 0)
        for b in batches:
            g.run(b)
        outs.append(g.state())
        g.close()
    assert np.array_equal(outs[0]["sets"][0]["beta"], outs[1]["sets"][0]["beta"])      # bit-identical
    assert np.array_equal(outs[0]["e"], outs[1]["e"]) and outs[0]["varE"] == outs[1]["varE"]


def test_blocked_and_literal_kernels_agree(gpu):
    prob = make_problem(900, 260, 33)
    sts = []
    for kernel in ("blocked", "literal"):
        g = gpu_sampler(prob, 2, 0.05, pi=0.1, est_pi=True, kernel=kernel)
        g.set_rng(5, 1)
        g.run(8)
        sts.append(g.state())
        g.close()
    assert rel(sts[0]["sets"][0]["beta"], sts[1]["sets"][0]["beta"]) < 1e-9
    assert np.array_equal(sts[0]["sets"][0]["delta"], sts[1]["sets"][0]["delta"])


# ----------------------------------------------------------------------------- sweep-level drop-in (M[mSet].funct)
@pytest.mark.parametrize("method,kw", [(0, dict(v=0.01)), (1, dict(v=0.05, pi=0.2, est_pi=True)), (2, dict(v=0.05, pi=0.1, est_pi=True))])
def test_sweep_level_plugin_call(gpu, method, kw):
    """ngp_sweep(mSet, ycorr, varE, beta, delta, varBeta): host buffers in/out, Julia keeps varE and the intercept."""
    prob = make_problem(450, 180, 41)
    ch, S = oracle_chain(prob, method, **kw)
    g = gpu_sampler(prob, method, **kw)
    beta, delta = np.zeros(180), np.ones(180, dtype=np.int64)
    varBeta = np.full(S.nvar, kw["v"])
    piHat = np.array([1 - kw.get("pi", 0.5), kw.get("pi", 0.5)])
    ycorr, mu = prob["y"].copy(), 0.0
    Lo = O.lib()
    import ctypes as C
    for it in range(1, 7):
        log = ch.iteration(seed=3, chain=0)
        # host side does what Julia would do: varE and intercept (here with the oracle's scalar routines)
        c2 = C.c_double(log["chi2_e"]); zm = C.c_double(log["z_mu"])
        varE = Lo.ngo_sample_varE(len(ycorr), ycorr.ctypes.data, 4.0, ch.scale_e, 1, 0, 0, it, C.byref(c2))
        mu = Lo.ngo_sample_intercept(len(ycorr), ycorr.ctypes.data, mu, varE, 0.0, 0.0, 1, 0, 0, it, C.byref(zm))
        g.set_replay([log])
        g.sweep(0, ycorr, varE, beta, delta, varBeta, piHat)
        assert rel(beta, S.beta) < TOL and rel(ycorr, ch.e) < TOL and rel(varBeta, S.varBeta) < TOL
        if method:
            assert np.array_equal(delta, S.delta) and rel(piHat, S.piHat) < TOL
    g.close()


def test_two_marker_sets_in_one_model(gpu):
    """for mSet in keys(M): two sets with different priors share the residual (samplers.jl:50-53)."""
    pa, pb = make_problem(380, 100, 51), make_problem(380, 70, 52)
    Xa, _, da = O.center_codes(pa["codes"]); Xb, _, db = O.center_codes(pb["codes"])
    y = pa["y"] + pb["y"] - 10.0
    Sa = O.MarkerSet(X=Xa, mpm=da, method=0, v=0.01, set_id=0)
    Sb = O.MarkerSet(X=Xb, mpm=db, method=2, v=0.05, pi=0.2, est_pi=True, set_id=1)
    ch = O.OracleChain(y, [Sa, Sb], v_e=y.var() / 2)
    g = ngp.Sampler(0)
    g.upload_genotypes(0, pa["codes"]); g.upload_genotypes(1, pb["codes"])
    g.set_prior(0, 0, *O.marker_hyper(0.01), 0.01)
    g.set_prior(1, 2, *O.marker_hyper(0.05), 0.05, pi_in=0.2, est_pi=True)
    g.set_phenotype(y); g.set_residual_prior(*O.residual_hyper(y.var() / 2)); g.set_intercept(True)
    g.set_rng(99, 2)
    for _ in range(8):
        ch.iteration(seed=99, chain=2)
    g.run(8)
    st = g.state()
    assert rel(st["sets"][0]["beta"], Sa.beta) < 1e-7 and rel(st["sets"][1]["beta"], Sb.beta) < 1e-7
    assert np.array_equal(st["sets"][1]["delta"], Sb.delta) and rel(st["e"], ch.e) < 1e-7
    g.close()


# ----------------------------------------------------------------------------- posterior agreement (native RNG, different seeds)
def test_posterior_means_agree_within_monte_carlo_error(gpu):
    """config 1 in miniature: BayesRR + intercept; GPU chain (seed A) vs oracle chain (seed B)."""
    prob = make_problem(400, 120, 61, q=10)
    v_e, v, _ = ngp.synth.priors(prob, "BayesRR")
    iters, burn = 1500, 300
    ch, S = oracle_chain(prob, 0, v, v_e=v_e)
    sb = np.zeros(120); sv = []; k = 0
    for it in range(iters):
        ch.iteration(seed=1, chain=0)
        if it >= burn:
            sb += S.beta; sv.append((ch.varE, S.varBeta[0])); k += 1
    ob, ov = sb / k, np.array(sv)
    g = gpu_sampler(prob, 0, v, v_e=v_e)
    g.set_rng(2, 0)
    g.run(burn); g.reset_posterior()
    gv = []
    for _ in range((iters - burn) // 10):
        g.run(10); st = g.state(want_e=False); gv.append((st["varE"], st["sets"][0]["varBeta"][0]))
    post = g.posterior(0)
    gv = np.array(gv)
    assert post["n"] == iters - burn
    X = prob["codes"] - prob["codes"].mean(0)
    gebv_o, gebv_g = X @ ob, X @ post["mean_beta"]
    assert np.corrcoef(gebv_o, gebv_g)[0, 1] > 0.995
    assert abs(gebv_o - gebv_g).max() < 0.25 * gebv_o.std() + 0.05
    # variance components: means within 5 MC standard errors (autocorrelation-inflated by using every 10th draw)
    for c in range(2):
        se = np.sqrt(ov[::10, c].var() / len(ov[::10]) + gv[:, c].var() / len(gv))
        assert abs(ov[:, c].mean() - gv[:, c].mean()) < 5 * se + 1e-12
    g.close()


# ----------------------------------------------------------------------------- whole driver through the reference-shaped API
def test_runLMEM_writes_reference_output_files(gpu, tmp_path):
    prob = make_problem(200, 50, 71)
    geno = tmp_path / "geno.txt"
    np.savetxt(geno, prob["codes"], fmt="%d", delimiter=" ")
    mp = tmp_path / "map.txt"
    mp.write_text("snpID,snpOrder,chrID\n" + "\n".join(f"s{i},{i + 1},{1 + i // 20}" for i in range(50)) + "\n")
    out = str(tmp_path / "outMCMC")
    VCV = {"M": ngp.BayesPR(10, 0.01), "e": ngp.Random("I", prob["var_y"] / 2)}
    s = ngp.runLMEM(f'y ~ 1 + SNP(M,"{geno}","{mp}")', {"y": prob["y"]}, 60, 20, 10, outFolder=out, VCV=VCV, seed=5)
    files = sorted(os.listdir(out))
    assert files == ["bOut", "betaMOut", "deltaMOut", "groupInfo_M.txt", "varEOut", "varMOut"]
    beta = np.loadtxt(os.path.join(out, "betaMOut"), delimiter="\t", skiprows=1)
    assert beta.shape == (4, 50)                     # kept iterations 30,40,50,60 (samplers.jl:26)
    assert open(os.path.join(out, "varMOut")).readline().strip().split("\t") == [f"reg_{r}" for r in range(1, 6)]   # chromosomes of 20,20,10 SNPs in windows of 10
    assert ngp.summaryMCMC("varE", outFolder=out).shape == (1, 1)
    # final state equals an oracle chain with the same native stream
    ro = ngp.prep2RegionData(None, "M", str(mp), 10)
    ch, S = oracle_chain(prob, 0, 0.01, region_off=ro)
    for _ in range(60):
        ch.iteration(seed=5, chain=0)
    assert rel(beta[-1], S.beta) < 1e-7
    s.close()


# ----------------------------------------------------------------------------- full-size properties (BASELINE config 2 rows, fewer markers)
def test_large_shape_properties(gpu):
    """n = 50,000 (the headline row count) x 4,096 markers generated on device: (i) packing round-trips on sampled
    columns, (ii) 1'e is invariant under a sweep (centred columns), (iii) blocked == literal, (iv) e == y - mu - X beta."""
    n, p, seed = 50000, 4096, 20261020
    pr = ngp.synth.problem(n, p, seed)
    v_e, v, pi = ngp.synth.priors(pr, "BayesC")
    sts = []
    for kernel in ("blocked", "literal"):
        g = ngp.Sampler(0, kernel=kernel)
        g.synth_genotypes(0, n, p, seed, pr["thr0"], pr["thr1"])
        if kernel == "blocked":
            cols = np.array([0, 1, 63, 64, 2047, 4095])
            for j in cols:
                assert np.array_equal(g.download_genotypes(0, int(j), int(j) + 1)[:, 0], ngp.synth.codes(seed, n, [j], pr["thr0"], pr["thr1"])[:, 0])
        g.set_prior(0, 2, *O.marker_hyper(v), v, pi_in=pi, est_pi=True)
        g.set_phenotype(pr["y"]); g.set_residual_prior(*O.residual_hyper(v_e)); g.set_intercept(True)
        g.set_rng(11, 0)
        g.run(3)
        sts.append(g.state())
        if kernel == "blocked":
            st = sts[0]
            nz = np.nonzero(st["sets"][0]["beta"])[0]
            Xnz = ngp.synth.codes(seed, n, nz, pr["thr0"], pr["thr1"]).astype(np.float64)
            Xnz -= Xnz.mean(0)
            e_expect = pr["y"] - st["mu"] - Xnz @ st["sets"][0]["beta"][nz]
            assert rel(st["e"], e_expect) < 1e-9
        g.close()
    assert rel(sts[0]["sets"][0]["beta"], sts[1]["sets"][0]["beta"]) < 1e-8
    assert np.array_equal(sts[0]["sets"][0]["delta"], sts[1]["sets"][0]["delta"])


# ----------------------------------------------------------------------------- edge cases the reference's arithmetic defines
@pytest.mark.parametrize("method,kw", [(2, dict(v=0.05, pi=0.3, est_pi=True)), (1, dict(v=0.05, pi=0.3)), (0, dict(v=0.02))])
def test_monomorphic_columns_follow_the_reference_arithmetic(gpu, method, kw):
    """A column without variation centres to zero: mpm = 0, so BayesB/C get log(0) and 0/0 in the inclusion probability (NaN ->
    `rand() < NaN` is false -> never included, functions.jl:169-174, 209-216) and BayesPR draws from the prior (lhs = 1/varBeta)."""
    prob = make_problem(300, 70, 13)
    codes = prob["codes"].copy()
    codes[:, 5] = 2; codes[:, 40] = 0; codes[:, 69] = 1
    prob = dict(prob, codes=np.asfortranarray(codes))
    with np.errstate(all="ignore"):
        ch, S = oracle_chain(prob, method, **kw)
        g = gpu_sampler(prob, method, **kw)
        g.set_rng(21, 0)
        for _ in range(5):
            ch.iteration(seed=21, chain=0)
    g.run(5)
    st = g.state()
    assert np.array_equal(st["sets"][0]["delta"], S.delta)
    assert rel(st["sets"][0]["beta"], S.beta) < 1e-8 and rel(st["e"], ch.e) < 1e-8
    if method:
        assert not st["sets"][0]["delta"][[5, 40, 69]].any() and not st["sets"][0]["beta"][[5, 40, 69]].any()
    g.close()


@pytest.mark.parametrize("n,p", [(5, 3), (31, 1), (33, 65)])
def test_tiny_shapes(gpu, n, p):
    """Fewer rows than one MMA chunk, a single marker, one marker more than a padded block."""
    prob = make_problem(n, p, 3 + n)
    for method, kw in ((2, dict(v=0.05, pi=0.4, est_pi=True)), (0, dict(v=0.02))):
        for kernel in ("blocked", "literal"):
            _run_replay(prob, method, kw, kernel, iters=3)


def test_call_order_and_argument_errors(gpu):
    s = ngp.Sampler(0)
    with pytest.raises(ngp.NgpError):
        s.set_phenotype(np.zeros(10))                       # no marker set yet
    prob = make_problem(100, 20, 1)
    s.upload_genotypes(0, prob["codes"])
    with pytest.raises(ngp.NgpError):
        s.run(1)                                            # no prior / phenotype
    with pytest.raises(ngp.NgpError):
        s.set_phenotype(np.zeros(99))                       # wrong n
    s.set_prior(0, 2, 4.0, 0.01, 0.02, pi_in=0.1)
    s.set_phenotype(prob["y"]); s.set_residual_prior(4.0, 1.0)
    with pytest.raises(ngp.NgpError):
        s.run(0)                                            # n_iter must be positive
    with pytest.raises(ngp.NgpError):
        s.upload_genotypes(1, prob["codes"][:50])           # all sets of a handle share n
    with pytest.raises(ngp.NgpError):
        s.configure(L.CFG_DEBUG, 1)                         # only the terminating experiments are accepted
    s.run(2)
    assert s.state()["iter"] == 2
    s.close()


# ----------------------------------------------------------------------------- oracle parity AT the BASELINE row counts and kernel geometries
def _native_vs_oracle_at_scale(n, p, model, iters, seed, expect=None, tol=1e-7, **geom):
    """Device-generated genotypes (ngp_synth_genotypes) against the CPU oracle on the SAME codes (regenerated on the host by the
    oracle's own Philox), native variate stream on both sides: per-iteration beta, delta, varE, mu, varBeta, pi and the residual."""
    pr = ngp.synth.problem(n, p, seed)
    v_e, v, pi = ngp.synth.priors(pr, model)
    method = METHODS["BayesPR" if model == "BayesRR" else model]
    codes = O.synth_codes(seed, n, 0, p, pr["thr0"], pr["thr1"])
    X, _, mpm = O.center_codes(codes)
    del codes
    S = O.MarkerSet(X=X, mpm=mpm, method=method, v=v, pi=pi, est_pi=(method == 2))
    ch = O.OracleChain(pr["y"], [S], v_e=v_e)
    O.set_threads(max(1, len(os.sched_getaffinity(0))))
    g = ngp.Sampler(0, **geom)
    g.synth_genotypes(0, n, p, seed, pr["thr0"], pr["thr1"])
    g.set_prior(0, method, *O.marker_hyper(v), v, pi_in=pi, est_pi=(method == 2))
    g.set_phenotype(pr["y"]); g.set_residual_prior(*O.residual_hyper(v_e)); g.set_intercept(True)
    g.set_rng(seed, 3)
    g.run(1)                                        # (an untimed, unchecked first launch: ngp_timing then names the kernel variant)
    t = g.timing()
    if expect:
        for k, val in expect.items():
            assert t[k] == val if not callable(val) else val(t[k]), f"geometry {k} = {t[k]}"
    g.close()
    g = ngp.Sampler(0, **geom)
    g.synth_genotypes(0, n, p, seed, pr["thr0"], pr["thr1"])
    g.set_prior(0, method, *O.marker_hyper(v), v, pi_in=pi, est_pi=(method == 2))
    g.set_phenotype(pr["y"]); g.set_residual_prior(*O.residual_hyper(v_e)); g.set_intercept(True)
    g.set_rng(seed, 3)
    worst = 0.0
    try:
        for it in range(iters):
            ch.iteration(seed=seed, chain=3)
            g.run(1)
            st = g.state()
            assert np.array_equal(st["sets"][0]["delta"], S.delta), f"indicator mismatch at iteration {it + 1}"
            worst = max(worst, rel(st["sets"][0]["beta"], S.beta), abs(st["varE"] / ch.varE - 1), abs(st["mu"] / ch.mu - 1),
                        rel(st["sets"][0]["varBeta"], S.varBeta), rel(st["e"], ch.e))
            if method:
                worst = max(worst, rel(st["sets"][0]["piHat"], S.piHat))
            assert worst < tol, f"iteration {it + 1}: rel diff {worst}"
    finally:
        O.set_threads(1)
        g.close()
    return worst


def test_oracle_parity_at_headline_rows_bayescpi(gpu):
    """BASELINE config 2 rows: n = 50,000 x 2,000 markers BayesCpi (functions.jl:197-236), default geometry of the headline run
    (blocks of 64, 352 rows per CTA, refetch ring, look-ahead 12), native stream, 3 iterations against the oracle to 1e-7; also with
    the round-1 geometry (blocks of 32, resident tiles, look-ahead 14) and with 2-bit storage."""
    _native_vs_oracle_at_scale(50000, 2000, "BayesC", 3, 20261020,
                               expect=dict(block=64, rows_per_cta=352, lookahead=12, ctas=lambda c: c >= 140))
    _native_vs_oracle_at_scale(50000, 2000, "BayesC", 2, 20261020, block=32, expect=dict(block=32, lookahead=14))
    _native_vs_oracle_at_scale(50000, 2000, "BayesC", 2, 20261020, storage="2bit", expect=dict(block=64, lookahead=12))


def test_oracle_parity_at_c5_geometry_refetch_ring(gpu):
    """BASELINE config 5 rows: n = 200,000 x 512 markers BayesCpi, 1,376+ rows per CTA (4 row groups per updater thread): blocks of 64 on a refetch
    ring of ONE tile whose dots the 8 dot warps share (default since the end of round 2), and the earlier default, blocks of 16;
    changed columns are re-read from L2."""
    _native_vs_oracle_at_scale(200000, 512, "BayesC", 3, 20261021,
                               expect=dict(block=64, tile_stages=1, refetch=1, kernel_variant=11, rows_per_cta=lambda r: r >= 1376, ctas=lambda c: c >= 140))
    _native_vs_oracle_at_scale(200000, 512, "BayesC", 2, 20261021, block=16,
                               expect=dict(block=16, rows_per_cta=lambda r: r >= 1376, ctas=lambda c: c >= 140))
    # the same rows as 2-bit tiles: blocks of 64, 4 row groups per updater thread (the BIGR instantiation), refetch ring
    _native_vs_oracle_at_scale(200000, 512, "BayesC", 2, 20261021, storage="2bit",
                               expect=dict(block=64, rows_per_cta=lambda r: r >= 1376))


def test_oracle_parity_at_c3_geometry_bayesb(gpu):
    """BASELINE config 3 rows: n = 100,000 x 512 markers BayesB (functions.jl:157-195), 704 rows per CTA: blocks of 64 on the refetch ring
    (the BIGR instantiation; default since round 2) and the round-1 geometry (blocks of 16, resident tiles)."""
    _native_vs_oracle_at_scale(100000, 512, "BayesB", 3, 20261022, expect=dict(block=64, rows_per_cta=704, refetch=1, kernel_variant=6))
    _native_vs_oracle_at_scale(100000, 512, "BayesB", 2, 20261022, block=16, expect=dict(block=16, rows_per_cta=704))
    _native_vs_oracle_at_scale(100000, 512, "BayesB", 2, 20261022, storage="2bit", expect=dict(block=64, rows_per_cta=704))
