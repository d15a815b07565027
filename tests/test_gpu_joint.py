"""GPU parity of the multi-breed ("Tuple") BayesPR path — sampleBayesPR!(mSet::Tuple, ...) of functions.jl:140-154 with
the inverse-Wishart covariance draw of functions.jl:513-516 — through the C ABI (ngp_set_joint_prior, ngp_run,
ngp_joint_sweep) against the CPU oracle (ngo_mb_sweep), with native Philox streams and with replayed variates."""
import numpy as np
import pytest

import nextgp.jl_b200 as ngp
from common import make_problem, rel
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _breeds(n, p, k, seed):
    probs = [make_problem(n, p, seed + 17 * b) for b in range(k)]
    y = probs[0]["y"].copy()
    for b in range(1, k):
        y += probs[b]["y"] - probs[b]["y"].mean()
    return probs, y


def _oracle(probs, y, v, region_off, v_e):
    Xk = [O.center_codes(pr["codes"])[0] for pr in probs]
    mb = O.MultiBreedOracle(Xk, v, region_off=region_off, set_id=0)
    ch = O.OracleChain(y, [], v_e=v_e, intercept=True)
    return ch, mb


def _gpu(probs, y, v, region_off, v_e, **kw):
    k = len(probs)
    s = ngp.Sampler(0, **kw)
    for b, pr in enumerate(probs):
        s.upload_genotypes(b, pr["codes"])
    df = 3.0 + k
    s.set_joint_prior(list(range(k)), df, np.asarray(v) * (df - k - 1.0), v, region_off=region_off)
    s.set_phenotype(y)
    s.set_residual_prior(*O.residual_hyper(v_e))
    s.set_intercept(True)
    return s


V2 = np.array([[0.02, 0.005], [0.005, 0.03]])
V3 = np.array([[0.02, 0.004, 0.002], [0.004, 0.03, -0.003], [0.002, -0.003, 0.025]])


@pytest.mark.parametrize("n,p,k,v,regions,kw", [
    (700, 96, 2, V2, None, {}),                                          # blocked tuple sweep (interleaved copy, joint draw in the chain warp)
    (700, 96, 2, V2, None, dict(kernel="literal")),                      # per-locus kernel
    (1203, 150, 3, V3, [0, 40, 41, 150], dict(min_rows=64)),             # k = 3 does not divide the block: per-locus kernel
    (333, 70, 2, V2, [0, 10, 70], dict(block=16, max_ctas=5)),
    (900, 200, 2, V2, [0, 64, 65, 130, 200], dict(block=64, lookahead=3)),
    (500, 48, 4, np.eye(4) * 0.02 + 0.004, [0, 20, 48], dict(min_rows=32)),
    (3001, 64, 2, V2, [0, 30, 64], dict(max_ctas=3)),                    # 1,504 rows per CTA, blocks of 64: no tuple instantiation of the blocked sweep -> per-locus kernel
])
def test_joint_native_chain_matches_oracle(gpu, n, p, k, v, regions, kw):
    probs, y = _breeds(n, p, k, 11)
    ro = None if regions is None else np.array(regions, dtype=np.int64)
    v_e = float(np.var(y)) / 2
    ch, mb = _oracle(probs, y, v, ro, v_e)
    g = _gpu(probs, y, v, ro, v_e, **kw)
    g.set_rng(77, 2)
    for _ in range(4):
        ch.iteration(seed=77, chain=2)
        mb.sweep(ch.e, ch.varE, it=ch.iter, seed=77, chain=2)
    g.run(3)
    g.run(1)
    st, js = g.state(), g.joint_state()
    assert rel(js["beta"], mb.beta) < 1e-8
    assert rel(js["varBeta"], mb.varBeta) < 1e-8
    assert abs(st["varE"] / ch.varE - 1) < 1e-9 and abs(st["mu"] - ch.mu) < 1e-9 * max(1.0, abs(ch.mu))
    assert rel(st["e"], ch.e) < 1e-8
    g.close()


def test_joint_replay_and_sweep_level_call(gpu):
    n, p, k = 640, 80, 2
    probs, y = _breeds(n, p, k, 5)
    ro = np.array([0, 30, 80], dtype=np.int64)
    v_e = float(np.var(y)) / 2
    ch, mb = _oracle(probs, y, V2, ro, v_e)
    logs, jlogs = [], []
    for _ in range(3):
        logs.append(ch.iteration(seed=901, chain=0))
        jlogs.append(mb.sweep(ch.e, ch.varE, it=ch.iter, seed=901, chain=0))
    # replay: the device consumes the logged variates instead of its own stream (different seed on purpose)
    for kern in ("blocked", "literal"):
        g = _gpu(probs, y, V2, ro, v_e, kernel=kern)
        g.set_rng(1, 0)
        g.set_replay(logs)
        g.set_joint_replay(jlogs)
        g.run(3)
        st, js = g.state(), g.joint_state()
        assert rel(js["beta"], mb.beta) < 1e-8 and rel(js["varBeta"], mb.varBeta) < 1e-8 and rel(st["e"], ch.e) < 1e-8
        g.close()
    # sweep level: host buffers in, mutated in place, like the reference's M[mSet].funct call
    ch2, mb2 = _oracle(probs, y, V2, ro, v_e)
    g = _gpu(probs, y, V2, ro, v_e)
    g.set_rng(55, 1)
    e = y - y.mean()
    beta = np.zeros((k, p)); vb = np.ascontiguousarray(np.broadcast_to(V2, (2, k, k)).copy())
    e_o = e.copy()
    for it in (1, 2):
        mb2.sweep(e_o, 1.7, it=it, seed=55, chain=1)     # every call advances the handle's iteration counter (it numbers the stream)
        g.joint_sweep(e, 1.7, beta, vb)
        assert rel(beta, mb2.beta) < 1e-8 and rel(vb, mb2.varBeta) < 1e-8 and rel(e, e_o) < 1e-8
    g.close()


def test_joint_prior_argument_checks(gpu):
    probs, y = _breeds(200, 40, 2, 3)
    s = ngp.Sampler(0)
    s.upload_genotypes(0, probs[0]["codes"])
    s.upload_genotypes(1, probs[1]["codes"][:, :30])
    with pytest.raises(ngp.NgpError):
        s.set_joint_prior([0, 1], 5.0, V2 * 2, V2)          # different numbers of loci
    s.upload_genotypes(1, probs[1]["codes"])
    with pytest.raises(ngp.NgpError):
        s.set_joint_prior([0, 0], 5.0, V2 * 2, V2)          # a set listed twice
    s.set_joint_prior([0, 1], 5.0, V2 * 2, V2)
    with pytest.raises(ngp.NgpError):
        s.set_prior(0, 0, 4.0, 0.01, 0.02)                  # members have no prior of their own
    s.close()


def test_runLMEM_with_a_tuple_of_marker_sets(gpu, tmp_path):
    """`(:M1,:M2) => BayesPR(9999, V)` through the reference-shaped driver: per-member beta files + one covariance file."""
    import os
    probs, y = _breeds(150, 30, 2, 21)
    out = str(tmp_path / "outMCMC")
    VCV = {("M1", "M2"): ngp.BayesPR(9999, V2), "e": ngp.Random("I", float(np.var(y)) / 2)}
    s = ngp.runLMEM("y ~ 1 + SNP(M1,a) + SNP(M2,b)", {"y": y}, 30, 10, 10, outFolder=out, VCV=VCV, seed=9,
                    matrices={"M1": probs[0]["codes"], "M2": probs[1]["codes"]})
    assert sorted(os.listdir(out)) == ["bOut", "betaM1Out", "betaM2Out", "varEOut", "varM1_M2Out"]
    b2 = np.loadtxt(os.path.join(out, "betaM2Out"), delimiter="\t", skiprows=1)
    vv = np.loadtxt(os.path.join(out, "varM1_M2Out"), delimiter="\t", skiprows=1)
    assert b2.shape == (2, 30) and vv.shape == (2, 4)             # kept iterations 20, 30
    ch, mb = _oracle(probs, y, V2, None, float(np.var(y)) / 2)
    for _ in range(30):
        ch.iteration(seed=9, chain=0)
        mb.sweep(ch.e, ch.varE, it=ch.iter, seed=9, chain=0)
    assert rel(b2[-1], mb.beta[1]) < 1e-7 and rel(vv[-1], mb.varBeta.ravel()) < 1e-7
    s.close()
