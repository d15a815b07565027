"""numpy model of the ARITHMETIC of the blocked CUDA sweep (nextgp.jl_b200/csrc/ngp_sweep.cuh):
the (1 + g/4) operand encoding, per-panel partial sums in fixed point, the invariant 1'e, the hoisted
per-marker constants (A, B, t, C, QSZ) and the Gram corrections inside a block.  Used on CPU to show
that this re-association of the reference's arithmetic reproduces the oracle's chain; the kernel itself
is checked against the oracle on the GPU (tests/test_gpu_parity.py)."""
import math

import numpy as np

PR, BB, BC = 0, 1, 2


class BlockedModel:
    def __init__(self, codes, y, method, v, pi=0.0, est_pi=False, region_off=None, v_e=1.0, B=64, T=4,
                 lhs0=None, rhs0=None, intercept=True):
        self.g = np.asarray(codes, dtype=np.int64)
        self.n, self.p = self.g.shape
        self.method, self.est_pi, self.B, self.T = method, est_pi, B, T
        self.colsum = self.g.sum(0)
        self.mean = self.colsum / self.n
        self.d = (self.n * (self.g ** 2).sum(0) - self.colsum ** 2) / self.n
        self.e = np.asarray(y, dtype=np.float64).copy()
        self.mu = 0.0
        self.intercept = intercept
        self.df_e, self.scale_e = 4.0, (0.0005 if v_e == 0 else v_e * 0.5)
        self.df, self.scale = 4.0, v * 0.5
        self.region_off = np.array([0, self.p]) if region_off is None else np.asarray(region_off)
        nvar = {PR: len(self.region_off) - 1, BB: self.p, BC: 1}[method]
        self.varBeta = np.full(nvar, float(v))
        self.region_of = np.repeat(np.arange(len(self.region_off) - 1), np.diff(self.region_off))
        self.beta = np.zeros(self.p)
        self.delta = np.ones(self.p, dtype=np.int64)
        self.piHat = np.array([1 - pi, pi]) if method != PR else np.array([0.0, 1.0])
        with np.errstate(divide="ignore"):
            self.logPi = np.log(self.piHat)
        self.lhs0 = np.zeros(self.p) if lhs0 is None else lhs0
        self.rhs0 = np.zeros(self.p) if rhs0 is None else rhs0
        R8 = -(-self.n // (8 * T))
        if R8 % 2 == 0:
            R8 += 1
        self.R = 8 * R8
        self.varE = float("nan")

    def iteration(self, log):
        n, p, B = self.n, self.p, self.B
        s = log["sets"][0]
        ee, se = float(self.e @ self.e), float(self.e.sum())
        self.varE = varE = (self.df_e * self.scale_e + ee) / log["chi2_e"]
        dmu = 0.0
        if self.intercept:
            iVarE = 1.0 / varE
            rhs = (se + n * self.mu) * iVarE
            lhs = n * iVarE
            mu_new = rhs / lhs + math.sqrt(1.0 / lhs) * log["z_mu"]
            dmu = self.mu - mu_new
            self.mu = mu_new
        M = max(2.0 * math.sqrt(n) * (math.sqrt(ee) + math.sqrt(n) * abs(dmu)), 1e-300)
        sh = 62 - 8 - math.frexp(M)[1]
        Stot = se + n * dmu
        self.e += dmu
        # phase 1 constants
        iVarE = 1.0 / varE
        if self.method == PR:
            vb = self.varBeta[self.region_of]
        elif self.method == BB:
            vb = self.varBeta
        else:
            vb = np.full(p, self.varBeta[0])
        with np.errstate(divide="ignore", invalid="ignore"):
            lhs = self.d * iVarE + self.lhs0 + 1.0 / vb
            ilhs = 1.0 / lhs
            cC = iVarE * ilhs
            cQ = np.sqrt(ilhs) * s["z"] + (0.0 if self.method == BC else self.rhs0 * ilhs)
            if self.method != PR:
                v0 = self.d * varE
                v1 = self.d * self.d * vb + v0
                cA = 0.5 * (np.log(v1) - np.log(v0)) + (self.logPi[0] - self.logPi[1])
                cB = 0.5 * (1.0 / v1 - 1.0 / v0)
                cT = np.log(1.0 / s["u"] - 1.0)
            else:
                cA, cB, cT = np.zeros(p), np.zeros(p), np.full(p, np.inf)
        enc = 1.0 + self.g / 4.0
        bb, nl = 0.0, 0
        for k0 in range(0, p, B):
            k1 = min(p, k0 + B)
            # per-panel partial sums, fixed point
            tot = np.zeros(k1 - k0, dtype=object)
            for t in range(self.T):
                rows = slice(t * self.R, min(n, (t + 1) * self.R))
                A_t = enc[rows, k0:k1].T @ self.e[rows]
                tot += np.array([int(np.rint(math.ldexp(a, sh))) for a in A_t], dtype=object)
            A = np.array([math.ldexp(float(x), -sh) for x in tot])
            r = 4.0 * (A - Stot) - self.mean[k0:k1] * Stot
            G = self.g[:, k0:k1].T @ self.g[:, k0:k1]
            cs = self.colsum[k0:k1].astype(np.float64)
            db = np.zeros(k1 - k0)
            for q in range(k1 - k0):
                j = k0 + q
                rr = self.d[j] * self.beta[j] + r[q]
                with np.errstate(invalid="ignore"):
                    inc = bool(cB[j] * rr * rr + cA[j] < cT[j])
                bn = rr * cC[j] + cQ[j] if inc else 0.0
                db[q] = bn - self.beta[j]
                if db[q] != 0.0:
                    gc = G[q, :].astype(np.float64) - cs[q] * cs / n
                    r[q + 1:] -= gc[q + 1:] * db[q]
                self.beta[j] = bn
                if self.method != PR:
                    self.delta[j] = int(inc)
                    nl += int(inc)
                if self.method == BB:
                    self.varBeta[j] = (self.scale * self.df + bn * bn) / s["chi2_b"][j] if inc else 0.0
                bb += bn * bn
            nz = np.nonzero(db)[0]
            if len(nz):
                K = np.sum(4.0 * db[nz] * (1.0 + 0.25 * self.mean[k0 + nz]))
                self.e -= enc[:, k0 + nz] @ (4.0 * db[nz]) - K
        if self.method == PR:
            for rg in range(len(self.region_off) - 1):
                sl = slice(self.region_off[rg], self.region_off[rg + 1])
                self.varBeta[rg] = (self.scale * self.df + self.beta[sl] @ self.beta[sl]) / s["chi2_b"][rg]
        elif self.method == BC:
            self.varBeta[0] = (self.scale * self.df + bb) / s["chi2_b"][0]
        if self.method != PR and self.est_pi:
            self.piHat = np.array([1.0 - s["beta_pi"], s["beta_pi"]])
            self.logPi = np.log(self.piHat)
