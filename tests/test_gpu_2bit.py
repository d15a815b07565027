"""2-bit device storage (NGP_STORE_2BIT; north_star "stored packed (2-bit or int8)", BASELINE config 3): four codes per byte in HBM,
expanded to the INT8 tensor-core operands on chip.  It replaces the dense Float64 matrix of prepMatVec.jl:116-131 / mme.jl:299-311.
Bit-exact packing round trips, replay parity against the oracle, and equality with the int8 storage (the integer dots are exact, so
the two storages must give the SAME chain)."""
import numpy as np
import pytest

import nextgp.jl_b200 as ngp
from common import gpu_sampler, make_problem, oracle_chain, rel
from nextgp.jl_b200 import _lib as L
from oracle import oracle as O
from test_gpu_parity import CASES, _run_replay

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,p,block", [(1, 3, 0), (5, 70, 16), (33, 65, 32), (1001, 130, 64), (3001, 40, 0)])
@pytest.mark.parametrize("fmt", ["i8", "f64", "packed2"])
def test_2bit_storage_round_trip_is_bit_exact(gpu, n, p, block, fmt):
    rng = np.random.default_rng(n * 131 + p)
    codes = np.asfortranarray(rng.integers(0, 3, size=(n, p), dtype=np.int8))
    s = ngp.Sampler(0, block=block, storage="2bit", min_rows=8)
    if fmt == "packed2":
        ld = (n + 3) // 4
        packed = np.zeros((ld, p), dtype=np.uint8, order="F")
        assert L.lib().ngp_pack2(codes.ctypes.data, n, p, n, packed.ctypes.data, ld) == 0
        s.upload_genotypes(0, packed, fmt=L.GENO_PACKED2, n=n)
    elif fmt == "f64":
        s.upload_genotypes(0, codes.astype(np.float64))
    else:
        s.upload_genotypes(0, codes)
    assert np.array_equal(s.download_genotypes(0), codes)
    mean, mpm = s.column_stats(0)
    _, mean_o, mpm_o = O.center_codes(codes)
    assert np.allclose(mean, mean_o, rtol=1e-14, atol=0) and np.allclose(mpm, mpm_o, rtol=1e-12, atol=1e-12)
    s.close()


def test_2bit_rejects_codes_outside_0_1_2(gpu):
    codes = np.zeros((40, 8), dtype=np.int8, order="F")
    codes[7, 3] = 3
    s = ngp.Sampler(0, storage="2bit")
    with pytest.raises(ngp.NgpError) as ei:
        s.upload_genotypes(0, codes)
    assert ei.value.code == L.EDATA
    s.close()


@pytest.mark.parametrize("name,method,kw", CASES, ids=[c[0] for c in CASES])
def test_2bit_replay_parity(gpu, name, method, kw):
    """the reference's loops (functions.jl:118-137, 157-195, 197-236) replayed through the 2-bit tiles"""
    prob = make_problem(640, 333, 17)
    _run_replay(prob, method, kw, "blocked", iters=10, storage="2bit")


@pytest.mark.parametrize("geom", [dict(block=16, max_ctas=3), dict(block=64, min_rows=8), dict(refetch=1, lookahead=13, min_rows=8),
                                  dict(block=16, refetch=1, lookahead=20, tile_stages=4, min_rows=8), dict(lookahead=24, block=16, min_rows=8)],
                         ids=["B16-many-rows", "B64", "refetch", "refetch-B16", "D24-B16"])
def test_2bit_replay_parity_over_geometries(gpu, geom):
    prob = make_problem(3001 if "max_ctas" in geom else 1300, 200, 91)
    _run_replay(prob, 2, dict(v=0.05, pi=0.2, est_pi=True), "blocked", iters=4, storage="2bit", **geom)
    _run_replay(prob, 0, dict(v=0.02), "blocked", iters=3, storage="2bit", **geom)


@pytest.mark.parametrize("method,kw", [(2, dict(v=0.05, pi=0.1, est_pi=True)), (1, dict(v=0.05, pi=0.2)), (0, dict(v=0.01))])
def test_2bit_and_int8_storage_give_the_same_chain(gpu, method, kw):
    """integer-exact dots and Gram matrices: the storage format cannot change a single bit of the chain"""
    prob = make_problem(2000, 500, 23)
    sts = []
    for storage in ("i8", "2bit"):
        g = gpu_sampler(prob, method, storage=storage, **kw)
        g.set_rng(77, 1)
        g.run(6)
        sts.append(g.state())
        g.close()
    assert np.array_equal(sts[0]["sets"][0]["beta"], sts[1]["sets"][0]["beta"])
    assert np.array_equal(sts[0]["sets"][0]["delta"], sts[1]["sets"][0]["delta"])
    assert sts[0]["varE"] == sts[1]["varE"] and np.array_equal(sts[0]["e"], sts[1]["e"])


def test_2bit_at_headline_rows_matches_int8_and_oracle_identity(gpu):
    """n = 50,000 x 2,048 generated on device straight into 2-bit tiles: packing round trip on sampled columns, the int8 chain
    bit for bit, and e == y - mu - X beta."""
    n, p, seed = 50000, 2048, 20261023
    pr = ngp.synth.problem(n, p, seed)
    v_e, v, pi = ngp.synth.priors(pr, "BayesC")
    sts = []
    for storage in ("2bit", "i8"):
        g = ngp.Sampler(0, storage=storage)
        g.synth_genotypes(0, n, p, seed, pr["thr0"], pr["thr1"])
        if storage == "2bit":
            for j in (0, 31, 32, 1000, 2047):
                assert np.array_equal(g.download_genotypes(0, j, j + 1)[:, 0], ngp.synth.codes(seed, n, [j], pr["thr0"], pr["thr1"])[:, 0])
        g.set_prior(0, 2, *O.marker_hyper(v), v, pi_in=pi, est_pi=True)
        g.set_phenotype(pr["y"]); g.set_residual_prior(*O.residual_hyper(v_e)); g.set_intercept(True)
        g.set_rng(5, 0)
        g.run(3)
        sts.append(g.state())
        g.close()
    assert np.array_equal(sts[0]["sets"][0]["beta"], sts[1]["sets"][0]["beta"]) and np.array_equal(sts[0]["e"], sts[1]["e"])
    st = sts[0]
    nz = np.nonzero(st["sets"][0]["beta"])[0]
    Xnz = ngp.synth.codes(seed, n, nz, pr["thr0"], pr["thr1"]).astype(np.float64)
    Xnz -= Xnz.mean(0)
    assert rel(st["e"], pr["y"] - st["mu"] - Xnz @ st["sets"][0]["beta"][nz]) < 1e-9


def test_2bit_unsupported_paths_fail_loudly(gpu):
    prob = make_problem(300, 64, 4)
    g = gpu_sampler(prob, 2, 0.05, pi=0.2, est_pi=True, storage="2bit", kernel="literal")
    with pytest.raises(ngp.NgpError) as ei:
        g.run(1)                                   # per-marker kernel reads int8 tiles
    assert ei.value.code == L.EUNSUPPORTED
    g.close()
    g = gpu_sampler(prob, 2, 0.05, pi=0.2, est_pi=True, storage="2bit")
    g.set_residual_weights(np.full(300, 1.5))
    with pytest.raises(ngp.NgpError) as ei:
        g.run(1)                                   # weighted residuals run on the per-marker kernel
    assert ei.value.code == L.EUNSUPPORTED
    g.close()
    s = ngp.Sampler(0, storage="2bit")
    s.upload_genotypes(0, prob["codes"])
    s2 = ngp.Sampler(0)
    s2.upload_genotypes(0, prob["codes"])
    s2.storage = L.STORE_2BIT
    with pytest.raises(ngp.NgpError):
        s2.upload_genotypes(1, prob["codes"])      # one storage format per handle
    s.close(); s2.close()
