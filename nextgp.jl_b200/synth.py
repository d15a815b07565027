"""Synthetic genotype / phenotype generator shared by tests and bench.py.

  f_j ~ U(0.05, 0.5);  g_ij ~ Binomial(2, f_j) realised as
      g_ij = (w >= thr0_j) + (w >= thr1_j),  w = word (i & 3) of Philox4x32-10(key = seed,
      ctr = (i >> 2, j, 0, 0x47454e4f)),  thr0 = floor((1-f)^2 2^32), thr1 = floor((1-f^2) 2^32)
  so any column can be regenerated anywhere (device: ngp_synth_genotypes; numpy: codes()).
  q = max(10, p // 100) causal loci, effects N(0,1) rescaled to h2 = 0.5, y = 10 + X beta + eps.
"""
from __future__ import annotations

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10 over uint32 numpy arrays (counter words), scalar key."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def frequencies(p: int, seed: int) -> np.ndarray:
    return np.random.default_rng([seed, 1]).uniform(0.05, 0.5, size=p)


def thresholds(f: np.ndarray):
    thr0 = np.floor((1.0 - f) ** 2 * 4294967296.0).astype(np.uint64).clip(0, 0xFFFFFFFF).astype(np.uint32)
    thr1 = np.floor((1.0 - f ** 2) * 4294967296.0).astype(np.uint64).clip(0, 0xFFFFFFFF).astype(np.uint32)
    return thr0, thr1


def codes(seed: int, n: int, cols: np.ndarray, thr0: np.ndarray, thr1: np.ndarray) -> np.ndarray:
    """int8 codes (n, len(cols)) Fortran order for the listed marker indices."""
    cols = np.asarray(cols, dtype=np.int64)
    n4 = (n + 3) // 4
    out = np.empty((n, len(cols)), dtype=np.int8, order="F")
    i4 = np.arange(n4, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for c, j in enumerate(cols):
        w = philox4x32_10(i4, np.full(n4, j, dtype=np.uint64), np.zeros(n4, dtype=np.uint64),
                          np.full(n4, 0x47454E4F, dtype=np.uint64), k0, k1)
        ws = np.stack(w, axis=1).reshape(-1)[:n]
        out[:, c] = (ws >= thr0[j]).astype(np.int8) + (ws >= thr1[j]).astype(np.int8)
    return out


def problem(n: int, p: int, seed: int, h2: float = 0.5, q: int | None = None):
    """Returns dict(f, thr0, thr1, causal, beta_true, y, var_y, sum2pq).  Only the q causal columns are materialised."""
    f = frequencies(p, seed)
    thr0, thr1 = thresholds(f)
    rng = np.random.default_rng([seed, 2])
    q = max(10, p // 100) if q is None else q
    q = min(q, p)
    causal = np.sort(rng.choice(p, size=q, replace=False))
    Xc = codes(seed, n, causal, thr0, thr1).astype(np.float64)
    Xc -= Xc.mean(axis=0)
    b = rng.normal(size=q)
    g = Xc @ b
    sg = g.std()
    if sg > 0:
        b *= 1.0 / sg
        g /= sg
    eps = rng.normal(size=n) * np.sqrt((1.0 - h2) / h2)
    y = 10.0 + g + eps
    beta_true = np.zeros(p)
    beta_true[causal] = b
    return {"f": f, "thr0": thr0, "thr1": thr1, "causal": causal, "beta_true": beta_true, "y": y,
            "var_y": float(y.var()), "var_g": 1.0, "sum2pq": float(np.sum(2.0 * f * (1.0 - f)))}


def priors(prob: dict, model: str):
    """Prior guesses of SURVEY §8(d): :e => Random("I", var(y)/2); BayesRR BayesPR(9999, vg/sum2pq);
    BayesC(0.05, vg/(0.05 sum2pq), estimatePi=true); BayesB(0.05, same, estimatePi=false)."""
    v_e = prob["var_y"] / 2.0
    vg = prob["var_y"] / 2.0
    if model in ("BayesRR", "BayesPR"):
        return v_e, vg / prob["sum2pq"], 0.0
    return v_e, vg / (0.05 * prob["sum2pq"]), 0.05
