"""ctypes front-end of the CPU oracle (oracle/ngp_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of ngp_oracle.c.  Imported from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs;
never from the product package.  PARITY UNPINNED (the reference has no tests).

The Python layer restates the reference's set-up arithmetic:
  * residual prior  df=4, scale=v(df-2)/df, v==0 -> 0.0005        (/root/reference/src/mme.jl:87-94)
  * marker prior    df=3+size(v,1), scale=v(df-2)/df              (mme.jl:492-506)
  * initial state   varBeta=v per slot, beta=0, delta=1, e=y      (mme.jl:57,443-444,513-520)
  * logPi/piHat     [1-pi, pi]                                    (mme.jl:351-372)
  * regionArray     [1:p] / [j:j] / map windows                   (mme.jl:335-348, misc.jl:163-215)
and drives one iteration in the order of samplers.jl:32-53.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libngp_oracle.so")

BAYESPR, BAYESB, BAYESC = 0, 1, 2
P_CHI2_E, P_Z_MU, P_U, P_Z, P_CHI2_B, P_PI_A, P_PI_B, P_IW = range(8)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ngp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


class _Set(C.Structure):
    _fields_ = [("n", C.c_int64), ("p", C.c_int64), ("X", C.c_void_p), ("Mp", C.c_void_p),
                ("mpm", C.c_void_p), ("lhs0", C.c_void_p), ("rhs0", C.c_void_p),
                ("method", C.c_int32), ("est_pi", C.c_int32), ("n_regions", C.c_int64),
                ("region_off", C.c_void_p), ("df", C.c_double), ("scale", C.c_double),
                ("set_id", C.c_int32), ("pad_", C.c_int32)]


class _SetState(C.Structure):
    _fields_ = [("beta", C.c_void_p), ("delta", C.c_void_p), ("varBeta", C.c_void_p),
                ("piHat", C.c_double * 2), ("logPi", C.c_double * 2)]


class _Variates(C.Structure):
    _fields_ = [("replay", C.c_int32), ("pad_", C.c_int32), ("seed", C.c_uint64),
                ("chain", C.c_uint32), ("iter", C.c_uint32), ("u", C.c_void_p), ("z", C.c_void_p),
                ("chi2_b", C.c_void_p), ("beta_pi", C.c_void_p)]


class _MbSet(C.Structure):
    _fields_ = [("n", C.c_int64), ("p", C.c_int64), ("k", C.c_int32), ("set_id", C.c_int32),
                ("Xk", C.c_void_p), ("n_regions", C.c_int64), ("region_off", C.c_void_p),
                ("df", C.c_double), ("scale", C.c_void_p)]


class _FxSet(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_cols", C.c_int32), ("col0", C.c_int32), ("data", C.c_void_p), ("xpx", C.c_void_p),
                ("lhs0", C.c_double), ("rhs0", C.c_double), ("Xp", C.c_void_p)]


class _RSet(C.Structure):
    _fields_ = [("n", C.c_int64), ("p", C.c_int64), ("X", C.c_void_p), ("mpm", C.c_void_p), ("lhs0", C.c_void_p),
                ("rhs0", C.c_void_p), ("n_class", C.c_int32), ("est_pi", C.c_int32), ("v_class", C.c_void_p),
                ("df", C.c_double), ("scale", C.c_double), ("set_id", C.c_int32), ("pad_", C.c_int32), ("Mp", C.c_void_p)]


class _RState(C.Structure):
    _fields_ = [("beta", C.c_void_p), ("delta", C.c_void_p), ("varBeta", C.c_void_p), ("piHat", C.c_void_p), ("logPi", C.c_void_p)]


class _RVariates(C.Structure):
    _fields_ = [("replay", C.c_int32), ("pad_", C.c_int32), ("seed", C.c_uint64), ("chain", C.c_uint32), ("iter", C.c_uint32),
                ("u", C.c_void_p), ("z", C.c_void_p), ("chi2_b", C.c_void_p), ("dir_pi", C.c_void_p)]


class _RCSet(C.Structure):
    _fields_ = [("n", C.c_int64), ("p", C.c_int64), ("X", C.c_void_p), ("mpm", C.c_void_p), ("lhs0", C.c_void_p), ("rhs0", C.c_void_p),
                ("n_class", C.c_int32), ("n_annot", C.c_int32), ("est_pi", C.c_int32), ("plus", C.c_int32), ("v_class", C.c_void_p),
                ("annot", C.c_void_p), ("df", C.c_double), ("scale", C.c_double), ("set_id", C.c_int32), ("pad_", C.c_int32)]


class _RCState(C.Structure):
    _fields_ = [("beta", C.c_void_p), ("delta", C.c_void_p), ("annot_cat", C.c_void_p), ("varBeta", C.c_void_p), ("piHat", C.c_void_p),
                ("logPi", C.c_void_p), ("annot_prob", C.c_void_p)]


class _RCVariates(C.Structure):
    _fields_ = [("replay", C.c_int32), ("pad_", C.c_int32), ("seed", C.c_uint64), ("chain", C.c_uint32), ("iter", C.c_uint32),
                ("u_annot", C.c_void_p), ("dirp", C.c_void_p), ("u", C.c_void_p), ("z", C.c_void_p), ("chi2_b", C.c_void_p), ("dir_pi", C.c_void_p)]


class _MbVariates(C.Structure):
    _fields_ = [("replay", C.c_int32), ("pad_", C.c_int32), ("seed", C.c_uint64),
                ("chain", C.c_uint32), ("iter", C.c_uint32), ("z", C.c_void_p),
                ("iw_chi2", C.c_void_p), ("iw_z", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.ngo_stream_uniform.restype = C.c_double
        L.ngo_stream_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.ngo_stream_normal.restype = C.c_double
        L.ngo_stream_normal.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.ngo_stream_chisq.restype = C.c_double
        L.ngo_stream_chisq.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double]
        L.ngo_stream_beta.restype = C.c_double
        L.ngo_stream_beta.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_double]
        L.ngo_sample_varE.restype = C.c_double
        L.ngo_sample_varE.argtypes = [C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_uint64,
                                      C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
        L.ngo_sample_intercept.restype = C.c_double
        L.ngo_sample_intercept.argtypes = [C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                           C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
        L.ngo_sample_varE_w.restype = C.c_double
        L.ngo_sample_varE_w.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_uint64,
                                        C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
        L.ngo_sample_intercept_w.restype = C.c_double
        L.ngo_sample_intercept_w.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                             C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
        L.ngo_sweep.restype = C.c_int
        L.ngo_sweep.argtypes = [C.POINTER(_Set), C.POINTER(_SetState), C.c_void_p, C.c_double, C.POINTER(_Variates)]
        L.ngo_fill_marker_variates.restype = None
        L.ngo_fill_marker_variates.argtypes = [C.POINTER(_Set), C.POINTER(_Variates)]
        L.ngo_center_codes.restype = None
        L.ngo_center_codes.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ngo_center_f64.restype = None
        L.ngo_center_f64.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ngo_synth_codes.restype = None
        L.ngo_synth_codes.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ngo_philox4x32_10.restype = None
        L.ngo_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ngo_rc_sweep.restype = C.c_int
        L.ngo_rc_sweep.argtypes = [C.POINTER(_RCSet), C.POINTER(_RCState), C.c_void_p, C.c_double, C.POINTER(_RCVariates)]
        L.ngo_set_threads.argtypes = [C.c_int]
        L.ngo_max_threads.restype = C.c_int
        L.ngo_sample_fixed.restype = None
        L.ngo_sample_fixed.argtypes = [C.POINTER(_FxSet), C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        L.ngo_r_sweep.restype = C.c_int
        L.ngo_r_sweep.argtypes = [C.POINTER(_RSet), C.POINTER(_RState), C.c_void_p, C.c_double, C.POINTER(_RVariates)]
        L.ngo_r_fill_variates.restype = None
        L.ngo_r_fill_variates.argtypes = [C.POINTER(_RSet), C.POINTER(_RVariates)]
        L.ngo_mb_sweep.restype = C.c_int
        L.ngo_mb_sweep.argtypes = [C.POINTER(_MbSet), C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.POINTER(_MbVariates)]
        L.ngo_mb_fill_variates.restype = None
        L.ngo_mb_fill_variates.argtypes = [C.POINTER(_MbSet), C.POINTER(_MbVariates)]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().ngo_philox4x32_10(_ptr(c), _ptr(k), _ptr(out))
    return out


def set_threads(t: int) -> None:
    lib().ngo_set_threads(int(t))


def max_threads() -> int:
    return int(lib().ngo_max_threads())


def center_codes(codes: np.ndarray):
    """codes: (n,p) int8 Fortran-ordered -> (X centred f64 F-order, mean, mpm).  prepMatVec.jl:129, mme.jl:305-307"""
    codes = np.asfortranarray(codes, dtype=np.int8)
    n, p = codes.shape
    X = np.empty((n, p), dtype=np.float64, order="F")
    mean = np.empty(p)
    mpm = np.empty(p)
    lib().ngo_center_codes(n, p, _ptr(codes), _ptr(X), _ptr(mean), _ptr(mpm))
    return X, mean, mpm


def synth_codes(seed: int, n: int, j0: int, j1: int, thr0: np.ndarray, thr1: np.ndarray) -> np.ndarray:
    out = np.empty((n, j1 - j0), dtype=np.int8, order="F")
    thr0 = np.ascontiguousarray(thr0, dtype=np.uint32)
    thr1 = np.ascontiguousarray(thr1, dtype=np.uint32)
    lib().ngo_synth_codes(C.c_uint64(seed), n, j0, j1, _ptr(thr0), _ptr(thr1), _ptr(out))
    return out


# --------------------------------------------------------------------------- set-up restatement
def residual_hyper(v_e: float):
    """mme.jl:87-94"""
    df = 4.0
    scale = 0.0005 if v_e == 0.0 else v_e * (df - 2.0) / df
    return df, scale


def marker_hyper(v: float):
    """mme.jl:492-506 for scalar v"""
    df = 3.0 + 1.0
    return df, v * (df - 2.0) / df


def regions_from_map(chr_id: np.ndarray, region_size: int) -> np.ndarray:
    """misc.jl:163-215 -> 0-based half-open offsets.  chr_id in file order (integer 1..C)."""
    chr_id = np.asarray(chr_id)
    p = len(chr_id)
    if region_size == 9999:
        return np.array([0, p], dtype=np.int64)
    offs = [0]
    if region_size == 99:
        # groupID = chrID ; searchsorted(groupID, g) for g in 1:nChr
        for c in range(1, len(np.unique(chr_id)) + 1):
            lo = np.searchsorted(chr_id, c, "left")
            hi = np.searchsorted(chr_id, c, "right")
            assert lo == offs[-1]
            offs.append(int(hi))
        return np.array(offs, dtype=np.int64)
    # fixed-size windows inside each chromosome, chromosomes visited in order of first appearance
    _, first = np.unique(chr_id, return_index=True)
    for c in chr_id[np.sort(first)]:
        tot = int(np.sum(chr_id == c))
        nreg = -(-tot // region_size)
        for g in range(nreg):
            offs.append(offs[-1] + min(region_size, tot - g * region_size))
    return np.array(offs, dtype=np.int64)


@dataclass
class MarkerSet:
    X: np.ndarray                 # centred f64 (n,p) F-order
    mpm: np.ndarray
    method: int
    v: float                      # prior variance guess
    pi: float = 0.0               # inclusion probability (B/C)
    est_pi: bool = False
    region_off: np.ndarray | None = None   # PR only
    lhs0: np.ndarray | None = None
    rhs0: np.ndarray | None = None
    set_id: int = 0
    use_Mp_copy: bool = False
    # state
    beta: np.ndarray = field(default=None)
    delta: np.ndarray = field(default=None)
    varBeta: np.ndarray = field(default=None)
    piHat: np.ndarray = field(default=None)
    logPi: np.ndarray = field(default=None)

    def __post_init__(self):
        n, p = self.X.shape
        self.df, self.scale = marker_hyper(self.v)
        if self.method == BAYESPR:
            if self.region_off is None:
                self.region_off = np.array([0, p], dtype=np.int64)
            self.region_off = np.ascontiguousarray(self.region_off, dtype=np.int64)
            nvar = len(self.region_off) - 1
        elif self.method == BAYESB:
            nvar = p
        else:
            nvar = 1
        self.nvar = nvar
        self.beta = np.zeros(p)
        self.delta = np.ones(p, dtype=np.int64)
        self.varBeta = np.full(nvar, float(self.v))
        if self.method != BAYESPR:
            self.piHat = np.array([1.0 - self.pi, self.pi])
            self.logPi = np.log(self.piHat)
        else:
            self.piHat = np.array([0.0, 1.0])
            self.logPi = np.array([-np.inf, 0.0])
        self.Mp = np.array(self.X, order="F", copy=True) if self.use_Mp_copy else None

    def c_set(self) -> _Set:
        n, p = self.X.shape
        s = _Set()
        s.n, s.p = n, p
        s.X = _ptr(self.X)
        s.Mp = _ptr(self.Mp)
        s.mpm = _ptr(self.mpm)
        s.lhs0 = _ptr(self.lhs0)
        s.rhs0 = _ptr(self.rhs0)
        s.method = self.method
        s.est_pi = int(self.est_pi)
        s.n_regions = (len(self.region_off) - 1) if self.method == BAYESPR else 0
        s.region_off = _ptr(self.region_off) if self.method == BAYESPR else None
        s.df, s.scale = self.df, self.scale
        s.set_id = self.set_id
        return s

    def c_state(self) -> _SetState:
        t = _SetState()
        t.beta, t.delta, t.varBeta = _ptr(self.beta), _ptr(self.delta), _ptr(self.varBeta)
        t.piHat[0], t.piHat[1] = self.piHat
        t.logPi[0], t.logPi[1] = self.logPi
        return t


def weighted_marker_arrays(X: np.ndarray, w: np.ndarray):
    """mme.jl:299-303 for E.str == "D": mpm_j = sum(x_j .* w .* x_j), Mp_j = (x_j .* w)'."""
    w = np.asarray(w, dtype=np.float64)
    Mp = np.asfortranarray(X * w[:, None])
    mpm = np.ascontiguousarray(np.einsum("ij,ij->j", Mp, X))
    return Mp, mpm


class OracleChain:
    """One chain: intercept (optional) + marker sets, iteration order of samplers.jl:32-53."""

    def __init__(self, y: np.ndarray, sets: list[MarkerSet], v_e: float, intercept: bool = True,
                 mu_lhs0: float = 0.0, mu_rhs0: float = 0.0, fixed: list | None = None, weights: np.ndarray | None = None):
        # weights = E.iVarStr of a "D" residual structure (mme.jl:70-73); the marker sets must then carry Mp = X .* w and the
        # weighted mpm (mme.jl:299-303): see weighted_marker_arrays
        self.w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self.fixed = fixed or []          # FixedSet objects, sampled after the intercept in this order (samplers.jl:37-39)
        self.y = np.asarray(y, dtype=np.float64)
        self.n = len(self.y)
        self.sets = sets
        self.df_e, self.scale_e = residual_hyper(v_e)
        self.intercept = intercept
        self.mu_lhs0, self.mu_rhs0 = mu_lhs0, mu_rhs0
        self.e = self.y.copy()          # ycorr = deepcopy(Y), mme.jl:57
        self.mu = 0.0
        self.varE = float("nan")
        self.iter = 0

    def iteration(self, seed: int = 0, chain: int = 0, replay: dict | None = None) -> dict:
        """Runs one iteration; returns the variate log of the iteration (the replay format)."""
        L = lib()
        self.iter += 1
        it = self.iter
        rep = replay is not None
        log: dict = {"iter": it}
        chi2_e = C.c_double(replay["chi2_e"] if rep else 0.0)
        if self.w is not None:      # samplers.jl:32-33
            self.varE = L.ngo_sample_varE_w(self.n, _ptr(self.e), _ptr(self.w), self.df_e, self.scale_e, int(rep), seed, chain, it, C.byref(chi2_e))
        else:
            self.varE = L.ngo_sample_varE(self.n, _ptr(self.e), self.df_e, self.scale_e, int(rep), seed, chain, it, C.byref(chi2_e))
        log["chi2_e"] = chi2_e.value
        if self.intercept:
            z_mu = C.c_double(replay["z_mu"] if rep else 0.0)
            if self.w is not None:
                self.mu = L.ngo_sample_intercept_w(self.n, _ptr(self.e), _ptr(self.w), self.mu, self.varE, self.mu_lhs0, self.mu_rhs0,
                                                   int(rep), seed, chain, it, C.byref(z_mu))
            else:
                self.mu = L.ngo_sample_intercept(self.n, _ptr(self.e), self.mu, self.varE, self.mu_lhs0, self.mu_rhs0,
                                                 int(rep), seed, chain, it, C.byref(z_mu))
            log["z_mu"] = z_mu.value
        log["z_fx"] = []
        for fi, F in enumerate(self.fixed):
            log["z_fx"].append(F.sample(self.e, self.varE, it, seed=seed, chain=chain, replay_z=(replay["z_fx"][fi] if rep else None)))
        log["sets"] = []
        for si, S in enumerate(self.sets):
            p = S.X.shape[1]
            if rep:
                r = replay["sets"][si]
                u = np.ascontiguousarray(r["u"], dtype=np.float64).copy()
                z = np.ascontiguousarray(r["z"], dtype=np.float64).copy()
                chi2_b = np.ascontiguousarray(r["chi2_b"], dtype=np.float64).copy()
                beta_pi = C.c_double(r["beta_pi"])
            else:
                u, z, chi2_b = np.zeros(p), np.zeros(p), np.zeros(S.nvar)
                beta_pi = C.c_double(0.0)
            V = _Variates()
            V.replay, V.seed, V.chain, V.iter = int(rep), seed, chain, it
            V.u, V.z, V.chi2_b = _ptr(u), _ptr(z), _ptr(chi2_b)
            V.beta_pi = C.cast(C.pointer(beta_pi), C.c_void_p)
            cs, ct = S.c_set(), S.c_state()
            if not rep:
                L.ngo_fill_marker_variates(C.byref(cs), C.byref(V))
            rc = L.ngo_sweep(C.byref(cs), C.byref(ct), _ptr(self.e), self.varE, C.byref(V))
            assert rc == 0
            S.piHat = np.array([ct.piHat[0], ct.piHat[1]])
            S.logPi = np.array([ct.logPi[0], ct.logPi[1]])
            log["sets"].append({"u": u, "z": z, "chi2_b": chi2_b, "beta_pi": beta_pi.value})
        return log

    def snapshot(self) -> dict:
        return {"varE": self.varE, "mu": self.mu, "e": self.e.copy(),
                "sets": [{"beta": S.beta.copy(), "delta": S.delta.copy(), "varBeta": S.varBeta.copy(),
                          "piHat": S.piHat.copy()} for S in self.sets]}


# --------------------------------------------------------------------------- fixed effects with several columns
class FixedSet:
    """One X[xSet] of getMME! (covariates / factor levels): functions.jl:22-54.  col0 = index of its first column among all fixed
    columns of the model (0 is the intercept), which addresses the variate stream."""

    def __init__(self, data: np.ndarray, col0: int, lhs0: float = 0.0, rhs0: float = 0.0, weights: np.ndarray | None = None):
        self.data = np.asfortranarray(data, dtype=np.float64)
        if self.data.ndim == 1:
            self.data = np.asfortranarray(self.data[:, None])
        self.n, self.c = self.data.shape
        self.Xp = None
        if weights is not None:           # E.str == "D": xpx = X'(w .* X), Xp = (X .* w)' (mme.jl:133-136)
            self.Xp = np.asfortranarray(self.data * np.asarray(weights, dtype=np.float64)[:, None])
            self.xpx = np.ascontiguousarray(self.data.T @ self.Xp)
        else:
            self.xpx = np.ascontiguousarray(self.data.T @ self.data)
        self.col0, self.lhs0, self.rhs0 = col0, lhs0, rhs0
        self.b = np.zeros(self.c)

    def sample(self, e: np.ndarray, varE: float, it: int, seed: int = 0, chain: int = 0, replay_z=None) -> np.ndarray:
        F = _FxSet()
        F.n, F.n_cols, F.col0, F.data, F.xpx, F.lhs0, F.rhs0 = self.n, self.c, self.col0, _ptr(self.data), _ptr(self.xpx), self.lhs0, self.rhs0
        F.Xp = _ptr(self.Xp)
        z = np.zeros(self.c) if replay_z is None else np.ascontiguousarray(replay_z, dtype=np.float64).copy()
        lib().ngo_sample_fixed(C.byref(F), _ptr(self.b), _ptr(e), varE, int(replay_z is not None), seed, chain, it, _ptr(z))
        return z


# --------------------------------------------------------------------------- BayesR
class BayesROracle:
    """functions.jl:238-289 with the wiring of mme.jl:374-383: variance classes v_class, class proportions pi (Dirichlet update
    when est_pi), one common variance.  delta holds the 1-based class of every locus."""

    def __init__(self, X: np.ndarray, mpm: np.ndarray, pi: np.ndarray, v_class: np.ndarray, v: float, est_pi: bool = False,
                 lhs0=None, rhs0=None, set_id: int = 0):
        self.X = np.asfortranarray(X, dtype=np.float64)
        self.mpm = np.ascontiguousarray(mpm, dtype=np.float64)
        self.n, self.p = self.X.shape
        self.v_class = np.ascontiguousarray(v_class, dtype=np.float64)
        self.nc = len(self.v_class)
        self.piHat = np.ascontiguousarray(pi, dtype=np.float64).copy()
        self.logPi = np.log(self.piHat)                                   # mme.jl:375
        self.df, self.scale = marker_hyper(v)
        self.est_pi = est_pi
        self.lhs0 = None if lhs0 is None else np.ascontiguousarray(lhs0, dtype=np.float64)
        self.rhs0 = None if rhs0 is None else np.ascontiguousarray(rhs0, dtype=np.float64)
        self.beta = np.zeros(self.p)
        self.delta = np.ones(self.p, dtype=np.int64)
        self.varBeta = np.array([float(v)])
        self.set_id = set_id
        self.Mp = None                    # E.str == "D": X .* w (mme.jl:303), with mpm the weighted one

    def sweep(self, e: np.ndarray, varE: float, it: int, seed: int = 0, chain: int = 0, replay: dict | None = None) -> dict:
        L = lib()
        S = _RSet()
        S.n, S.p, S.X, S.mpm = self.n, self.p, _ptr(self.X), _ptr(self.mpm)
        S.Mp = _ptr(self.Mp) if self.Mp is not None else None
        S.lhs0 = _ptr(self.lhs0) if self.lhs0 is not None else None
        S.rhs0 = _ptr(self.rhs0) if self.rhs0 is not None else None
        S.n_class, S.est_pi, S.v_class, S.df, S.scale, S.set_id = self.nc, int(self.est_pi), _ptr(self.v_class), self.df, self.scale, self.set_id
        T = _RState()
        T.beta, T.delta, T.varBeta, T.piHat, T.logPi = _ptr(self.beta), _ptr(self.delta), _ptr(self.varBeta), _ptr(self.piHat), _ptr(self.logPi)
        if replay is None:
            u, z, c2, dp = np.zeros((self.p, self.nc)), np.zeros(self.p), np.zeros(1), np.zeros(self.nc)
        else:
            u = np.ascontiguousarray(replay["u"], dtype=np.float64).copy(); z = np.ascontiguousarray(replay["z"], dtype=np.float64).copy()
            c2 = np.ascontiguousarray(replay["chi2_b"], dtype=np.float64).copy(); dp = np.ascontiguousarray(replay["dir_pi"], dtype=np.float64).copy()
        V = _RVariates()
        V.replay, V.seed, V.chain, V.iter = int(replay is not None), seed, chain, it
        V.u, V.z, V.chi2_b, V.dir_pi = _ptr(u), _ptr(z), _ptr(c2), _ptr(dp)
        if replay is None:
            L.ngo_r_fill_variates(C.byref(S), C.byref(V))
        rc = L.ngo_r_sweep(C.byref(S), C.byref(T), _ptr(e), varE, C.byref(V))
        assert rc == 0, rc
        return {"u": u, "z": z, "chi2_b": c2, "dir_pi": dp}


# --------------------------------------------------------------------------- multi-breed (Tuple) BayesPR
class MultiBreedOracle:
    """functions.jl:140-154,513-516 with the layout of mme.jl:448-467 (unwired in v1.2.0)."""

    def __init__(self, Xk: list[np.ndarray], v: np.ndarray, region_off: np.ndarray | None = None, set_id: int = 0):
        self.Xk = [np.asfortranarray(x, dtype=np.float64) for x in Xk]
        self.k = len(Xk)
        self.n, self.p = self.Xk[0].shape
        self.df = 3.0 + self.k                                   # mme.jl:493
        self.scale = np.ascontiguousarray(np.asarray(v, dtype=np.float64) * (self.df - self.k - 1.0))  # mme.jl:501
        self.region_off = np.ascontiguousarray(region_off if region_off is not None else [0, self.p], dtype=np.int64)
        self.R = len(self.region_off) - 1
        self.beta = np.zeros((self.k, self.p))
        self.varBeta = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.R, self.k, self.k)).copy())
        self.set_id = set_id
        self._ptrs = (C.c_void_p * self.k)(*[x.ctypes.data for x in self.Xk])

    def sweep(self, e: np.ndarray, varE: float, it: int, seed: int = 0, chain: int = 0, replay: dict | None = None) -> dict:
        L = lib()
        S = _MbSet()
        S.n, S.p, S.k, S.set_id = self.n, self.p, self.k, self.set_id
        S.Xk = C.cast(self._ptrs, C.c_void_p)
        S.n_regions, S.region_off = self.R, _ptr(self.region_off)
        S.df, S.scale = self.df, _ptr(self.scale)
        k = self.k
        if replay is None:
            z = np.zeros((self.p, k)); iw_chi2 = np.zeros((self.R, k)); iw_z = np.zeros((self.R, k, k))
        else:
            z = np.ascontiguousarray(replay["z"]).copy(); iw_chi2 = np.ascontiguousarray(replay["iw_chi2"]).copy()
            iw_z = np.ascontiguousarray(replay["iw_z"]).copy()
        V = _MbVariates()
        V.replay, V.seed, V.chain, V.iter = int(replay is not None), seed, chain, it
        V.z, V.iw_chi2, V.iw_z = _ptr(z), _ptr(iw_chi2), _ptr(iw_z)
        if replay is None:
            L.ngo_mb_fill_variates(C.byref(S), C.byref(V))
        rc = L.ngo_mb_sweep(C.byref(S), _ptr(self.beta), _ptr(self.varBeta), _ptr(e), varE, C.byref(V))
        assert rc == 0, rc
        return {"z": z, "iw_chi2": iw_chi2, "iw_z": iw_z}


class BayesRCOracle:
    """BayesRCpi (functions.jl:291-360; plus=False) and BayesRCplus (functions.jl:362-419; plus=True) with the wiring of mme.jl:385-418:
    annot (p x nAnnot integer matrix), variance classes v_class, class proportions pi shared by all annotations at the start, ONE variance
    per annotation (varBeta = fill(v, nAnnot), mme.jl:516).  delta = 1-based class, annot_cat = 1-based annotation (RCpi)."""

    def __init__(self, X, mpm, pi, v_class, v: float, annot, est_pi: bool = False, plus: bool = False, lhs0=None, rhs0=None, set_id: int = 0):
        self.X = np.asfortranarray(X, dtype=np.float64)
        self.mpm = np.ascontiguousarray(mpm, dtype=np.float64)
        self.n, self.p = self.X.shape
        self.v_class = np.ascontiguousarray(v_class, dtype=np.float64)
        self.nc = len(self.v_class)
        self.annot = np.ascontiguousarray(annot, dtype=np.int32)
        assert self.annot.shape[0] == self.p
        self.nA = self.annot.shape[1]
        self.piHat = np.tile(np.asarray(pi, dtype=np.float64), (self.nA, 1)).copy()         # mme.jl:392
        self.logPi = np.log(self.piHat)                                                      # mme.jl:390
        with np.errstate(invalid="ignore", divide="ignore"):
            self.annot_prob = np.ascontiguousarray(self.annot / self.annot.sum(1, keepdims=True), dtype=np.float64)   # mme.jl:395
        self.df, self.scale = marker_hyper(v)
        self.est_pi, self.plus = est_pi, plus
        self.lhs0 = None if lhs0 is None else np.ascontiguousarray(lhs0, dtype=np.float64)
        self.rhs0 = None if rhs0 is None else np.ascontiguousarray(rhs0, dtype=np.float64)
        self.beta = np.zeros(self.p)
        self.delta = np.ones(self.p, dtype=np.int64)
        self.annot_cat = np.zeros(self.p, dtype=np.int64)                                    # mme.jl:403
        self.varBeta = np.full(self.nA, float(v))
        self.set_id = set_id

    def sweep(self, e: np.ndarray, varE: float, it: int, seed: int = 0, chain: int = 0, replay: dict | None = None) -> dict:
        L = lib()
        S = _RCSet()
        S.n, S.p, S.X, S.mpm = self.n, self.p, _ptr(self.X), _ptr(self.mpm)
        S.lhs0 = _ptr(self.lhs0) if self.lhs0 is not None else None
        S.rhs0 = _ptr(self.rhs0) if self.rhs0 is not None else None
        S.n_class, S.n_annot, S.est_pi, S.plus = self.nc, self.nA, int(self.est_pi), int(self.plus)
        S.v_class, S.annot, S.df, S.scale, S.set_id = _ptr(self.v_class), _ptr(self.annot), self.df, self.scale, self.set_id
        T = _RCState()
        T.beta, T.delta, T.annot_cat, T.varBeta = _ptr(self.beta), _ptr(self.delta), _ptr(self.annot_cat), _ptr(self.varBeta)
        T.piHat, T.logPi, T.annot_prob = _ptr(self.piHat), _ptr(self.logPi), _ptr(self.annot_prob)
        p, nA, nc = self.p, self.nA, self.nc
        ushape, zshape = ((p, nA, nc), (p, nA)) if self.plus else ((p, nc), (p,))
        if replay is None:
            v = {"u_annot": np.zeros(p), "dirp": np.zeros((p, nA)), "u": np.zeros(ushape), "z": np.zeros(zshape), "chi2_b": np.zeros(nA),
                 "dir_pi": np.zeros((nA, nc))}
        else:
            v = {k: np.ascontiguousarray(replay[k], dtype=np.float64).copy() for k in ("u_annot", "dirp", "u", "z", "chi2_b", "dir_pi")}
        V = _RCVariates()
        V.replay, V.seed, V.chain, V.iter = int(replay is not None), seed, chain, it
        V.u_annot, V.dirp, V.u, V.z, V.chi2_b, V.dir_pi = (_ptr(v[k]) for k in ("u_annot", "dirp", "u", "z", "chi2_b", "dir_pi"))
        rc = L.ngo_rc_sweep(C.byref(S), C.byref(T), _ptr(e), varE, C.byref(V))
        assert rc == 0, rc
        return v
