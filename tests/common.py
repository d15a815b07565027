"""Shared builders for the parity tests: small synthetic problems + oracle chains (test infrastructure)."""
import numpy as np

from oracle import oracle as O

import nextgp.jl_b200 as ngp

METHODS = {"BayesPR": 0, "BayesB": 1, "BayesC": 2}


def make_problem(n, p, seed, q=None):
    prob = ngp.synth.problem(n, p, seed, q=q)
    codes = ngp.synth.codes(seed, n, np.arange(p), prob["thr0"], prob["thr1"])
    # guarantee polymorphic columns
    for j in np.where(codes.min(0) == codes.max(0))[0]:
        codes[0, j] = (codes[0, j] + 1) % 3
    prob["codes"] = np.asfortranarray(codes)
    return prob


def oracle_chain(prob, method, v, pi=0.0, est_pi=False, region_off=None, v_e=None, lhs0=None, rhs0=None,
                 intercept=True, set_id=0):
    X, mean, mpm = O.center_codes(prob["codes"])
    S = O.MarkerSet(X=X, mpm=mpm, method=method, v=v, pi=pi, est_pi=est_pi, region_off=region_off,
                    lhs0=lhs0, rhs0=rhs0, set_id=set_id)
    v_e = prob["var_y"] / 2 if v_e is None else v_e
    return O.OracleChain(prob["y"], [S], v_e=v_e, intercept=intercept), S


def gpu_sampler(prob, method, v, pi=0.0, est_pi=False, region_off=None, v_e=None, lhs0=None, rhs0=None,
                intercept=True, kernel="blocked", block=0, min_rows=0, max_ctas=0, upload="i8", storage="i8", **geom):
    s = ngp.Sampler(0, kernel=kernel, block=block, min_rows=min_rows, max_ctas=max_ctas, storage=storage, **geom)
    if upload == "f64":
        s.upload_genotypes(0, prob["codes"].astype(np.float64))
    else:
        s.upload_genotypes(0, prob["codes"])
    df, scale = O.marker_hyper(v)
    s.set_prior(0, method, df, scale, v, pi_in=pi, est_pi=est_pi, region_off=region_off, lhs0=lhs0, rhs0=rhs0)
    s.set_phenotype(prob["y"])
    v_e = prob["var_y"] / 2 if v_e is None else v_e
    s.set_residual_prior(*O.residual_hyper(v_e))
    s.set_intercept(intercept)
    return s


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / scale)
