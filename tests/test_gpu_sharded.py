"""Row-sharded single chain (BASELINE config 5): the individuals are split over several handles whose
persistent kernels reduce every marker's partial dots through each other's synchronisation areas (peer memory; NVLink
between GPUs).  On a one-GPU box all ranks run as ONE cooperative grid over all ranks' data (ngp_run_group,
ngp::gibbs_group_kernel) — the same protocol, the same code path (`ld/red ...sys` on every rank's synchronisation area),
which is how SURVEY §8e / §4(iv) ask the multi-rank logic to be tested when GPUs are scarce: kernels that wait for one
another are never separate launches on one GPU."""
import numpy as np
import pytest

import nextgp.jl_b200 as ngp
from common import make_problem, oracle_chain, rel
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _sharded(prob, devices, method, v, pi=0.0, est_pi=False, region_off=None, max_ctas=6, **kw):
    ch = ngp.ShardedChain(devices, max_ctas=max_ctas, **kw)
    ch.upload_genotypes(0, prob["codes"])
    df, scale = O.marker_hyper(v)
    for s in ch.shards:
        s.set_prior(0, method, df, scale, v, pi_in=pi, est_pi=est_pi, region_off=region_off)
        s.set_residual_prior(*O.residual_hyper(prob["var_y"] / 2))
        s.set_intercept(True)
    ch.set_phenotype(prob["y"])
    return ch


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("method,kw", [(2, dict(v=0.05, pi=0.1, est_pi=True)), (0, dict(v=0.01, region_off=[0, 50, 51, 130])),
                                       (1, dict(v=0.05, pi=0.2))])
def test_sharded_chain_on_one_device_matches_oracle(gpu, world, method, kw):
    prob = make_problem(1001, 130, 8)
    kw = dict(kw)
    if "region_off" in kw:
        kw["region_off"] = np.array(kw["region_off"], dtype=np.int64)
    ch_o, S = oracle_chain(prob, method, **kw)
    ch = _sharded(prob, [0] * world, method, **kw)
    for s in ch.shards:
        s.set_rng(31, 4)
    for _ in range(4):
        ch_o.iteration(seed=31, chain=4)
    ch.run(3)
    ch.run(1)          # a second launch: the cross-shard barrier counter and accumulators carry over
    st = ch.state()
    assert rel(st["sets"][0]["beta"], S.beta) < 1e-8 and np.array_equal(st["sets"][0]["delta"], S.delta)
    assert rel(st["sets"][0]["varBeta"], S.varBeta) < 1e-8
    assert abs(st["varE"] / ch_o.varE - 1) < 1e-9 and rel(st["e"], ch_o.e) < 1e-8
    # every shard holds the identical chain state (same sums, same draws)
    for other in st["shards"][1:]:
        assert np.array_equal(other["sets"][0]["beta"], st["shards"][0]["sets"][0]["beta"])
        assert other["varE"] == st["shards"][0]["varE"] and other["mu"] == st["shards"][0]["mu"]
    mean, mpm = ch.shards[-1].column_stats(0)
    _, mean_o, mpm_o = O.center_codes(prob["codes"])
    assert np.allclose(mean, mean_o, rtol=1e-14) and np.allclose(mpm, mpm_o, rtol=1e-12)
    ch.close()


@pytest.mark.parametrize("world,hier", [(2, False), (3, False), (3, True), (4, None)])
@pytest.mark.parametrize("block,method,kw", [(32, 2, dict(v=0.05, pi=0.1, est_pi=True)), (64, 1, dict(v=0.05, pi=0.2)),
                                             (16, 0, dict(v=0.01, region_off=[0, 50, 51, 130])), (64, 2, dict(v=0.05, pi=0.3, est_pi=True))])
def test_sharded_blocked_kernel_on_one_device_matches_oracle(gpu, world, hier, block, method, kw):
    """Row sharding of the BLOCKED kernel (SURVEY 8e "B-many scalars blocked"): every worker CTA pushes the B partial sums of a block into
    the accumulator ring of every rank (or, with the rank-local pre-reduction, the rank's prep warps push one total per marker), every rank's
    chain CTA derives the identical lists; the banded Gram is summed over the ranks at set-up.  All ranks as ONE cooperative grid (ngp_run_group)."""
    prob = make_problem(1500, 200, 9)
    kw = dict(kw)
    if "region_off" in kw:
        kw["region_off"] = np.array(kw["region_off"] + [200] if kw["region_off"][-1] != 200 else kw["region_off"], dtype=np.int64)
    ch_o, S = oracle_chain(prob, method, **kw)
    ch = _sharded(prob, [0] * world, method, kernel="blocked", block=block, lookahead=5, max_ctas=5, min_rows=8, **kw)
    for s in ch.shards:
        s.set_rng(31, 4)
        if hier is False:           # rank-local pre-reduction (one push per rank and marker) is the default; 512 switches it off
            s.configure(ngp._lib.CFG_OPT, 512)
    for _ in range(5):
        ch_o.iteration(seed=31, chain=4)
    ch.run(3)
    ch.run(2)
    st = ch.state()
    assert st["shards"][0]["sets"][0]["beta"].shape == S.beta.shape
    assert rel(st["sets"][0]["beta"], S.beta) < 1e-8 and np.array_equal(st["sets"][0]["delta"], S.delta)
    assert rel(st["sets"][0]["varBeta"], S.varBeta) < 1e-8 and abs(st["varE"] / ch_o.varE - 1) < 1e-9 and rel(st["e"], ch_o.e) < 1e-8
    for other in st["shards"][1:]:
        assert np.array_equal(other["sets"][0]["beta"], st["shards"][0]["sets"][0]["beta"]) and other["varE"] == st["shards"][0]["varE"]
    assert ch.shards[0].timing()["kernel_variant"] == 8
    ch.close()


def test_sharded_blocked_needs_the_global_gram(gpu):
    prob = make_problem(400, 64, 3)
    ch = ngp.ShardedChain([0, 0], max_ctas=4, kernel="literal")
    ch.upload_genotypes(0, prob["codes"])
    df, scale = O.marker_hyper(0.05)
    for s in ch.shards:
        s.configure(ngp._lib.CFG_KERNEL, ngp._lib.KERNEL_BLOCKED)      # blocked kernel, but the Grams were never summed over the ranks
        s.set_prior(0, 2, df, scale, 0.05, pi_in=0.1, est_pi=True)
        s.set_residual_prior(*O.residual_hyper(1.0)); s.set_intercept(True)
    ch.set_phenotype(prob["y"])
    with pytest.raises(ngp.NgpError) as ei:
        ch.run(1)
    assert ei.value.code == ngp._lib.EINVAL
    ch.close()


def test_sharded_blocked_chain_over_two_gpus(gpu):
    if ngp._lib.lib().ngp_device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    prob = make_problem(6000, 400, 12)
    kw = dict(v=0.05, pi=0.1, est_pi=True)
    ch_o, S = oracle_chain(prob, 2, **kw)
    ch = _sharded(prob, [0, 1], 2, max_ctas=40, kernel="blocked", block=64, lookahead=8, **kw)
    for s in ch.shards:
        s.set_rng(5, 0)
    for _ in range(5):
        ch_o.iteration(seed=5, chain=0)
    ch.run(5)
    st = ch.state()
    assert rel(st["sets"][0]["beta"], S.beta) < 1e-8 and rel(st["e"], ch_o.e) < 1e-8
    assert np.array_equal(st["shards"][0]["sets"][0]["beta"], st["shards"][1]["sets"][0]["beta"])
    ch.close()


def test_sharded_chain_over_two_gpus(gpu):
    if ngp._lib.lib().ngp_device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    prob = make_problem(3000, 200, 12)
    kw = dict(v=0.05, pi=0.1, est_pi=True)
    ch_o, S = oracle_chain(prob, 2, **kw)
    ch = _sharded(prob, [0, 1], 2, max_ctas=40, **kw)
    for s in ch.shards:
        s.set_rng(5, 0)
    for _ in range(5):
        ch_o.iteration(seed=5, chain=0)
    ch.run(5)
    st = ch.state()
    assert rel(st["sets"][0]["beta"], S.beta) < 1e-8 and rel(st["e"], ch_o.e) < 1e-8
    ch.close()


def test_same_device_shards_refuse_separate_launches(gpu):
    """ngp_run on a handle whose peer lives on the same device would be a separate launch waiting for another: refused."""
    prob = make_problem(400, 40, 3)
    ch = _sharded(prob, [0, 0], 2, v=0.05, pi=0.1, est_pi=True)
    with pytest.raises(ngp.NgpError) as ei:
        ch.shards[0].run(1)
    assert ei.value.code == ngp._lib.EUNSUPPORTED
    ch.run(1)                                  # the group launch works
    ch.close()


def test_sharded_argument_checks(gpu):
    s = ngp.Sampler(0, kernel="blocked")
    s.shard_init(0, 2)
    prob = make_problem(300, 40, 2)
    s.upload_genotypes(0, prob["codes"][:152])
    with pytest.raises(ngp.NgpError):
        s.shard_init(1, 2)                     # after the first upload
    df, scale = O.marker_hyper(0.01)
    s.set_prior(0, 0, df, scale, 0.01)
    s.set_phenotype(prob["y"][:152])
    with pytest.raises(ngp.NgpError):
        s.run(1)                               # not attached
    s.close()
