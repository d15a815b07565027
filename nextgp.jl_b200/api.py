"""Host-side mirror of NextGP.jl's interface for the marker-effect path, on top of the C ABI.

The reference is Julia (absent from this image), so the host side is written in Python with the
reference's own names and argument meaning; the Julia `ccall` shim a maintainer would add is in
INTEGRATION.md / julia/NextGPB200.jl.  Citations are file:line under NextGP.jl v1.2.0 `src/`.

  BayesPR / BayesB / BayesC / Random / SummaryStatistics   runTime.jl:30-76,135-152
  prep_snp            prepMatVec.jl:113-134 (read text, drop columns with missing, centring is done on device)
  prep2RegionData     misc.jl:163-215
  getMME              mme.jl:87-94 (E), :286-446 (marker wiring), :492-520 (df/scale/varBeta), :543-595 (headers)
  runSampler          samplers.jl:23-106
  runLMEM             MCMC.jl:31-41
  outMCMC             outFiles.jl:17-21
  summaryMCMC         misc.jl:241-244
  Sampler             the handle: one chain on one GPU (ngp_create ... ngp_destroy)
"""
from __future__ import annotations

import ctypes as C
import os
import re
import shutil
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from . import _lib as L


# ----------------------------------------------------------------------------- prior types (runTime.jl)
@dataclass
class BayesPRType:
    r: int
    v: Any                       # Float64, or a k x k matrix for a tuple of correlated marker sets
    name: str = "BayesPR"


@dataclass
class BayesBType:
    pi: float
    v: float
    name: str = "BayesB"
    estimatePi: bool = False


@dataclass
class BayesCType:
    pi: float
    v: float
    name: str = "BayesC"
    estimatePi: bool = False


@dataclass
class BayesRType:
    pi: Any                      # vector of class proportions (runTime.jl:78-84)
    class_: Any                  # vector of class scales, increasing, e.g. [0.0, 0.0001, 0.001, 0.01]
    v: float
    name: str = "BayesR"
    estimatePi: bool = False


@dataclass
class BayesRCType:
    pi: Any                      # starting class proportions of every annotation (runTime.jl:95-102)
    class_: Any
    v: float
    annot: Any                   # (p, n_annot) integer matrix
    name: str = "BayesRCπ"
    estimatePi: bool = False


@dataclass
class BayesLogVarType:
    v: float
    f: Any                       # None: `covariates` already is the (p, k) design matrix; else "1 + x1 + x2" over the columns of `covariates`
    covariates: Any
    varZeta: float
    name: str = "BayesLV"
    estimateVarZeta: Any = False  # False: fixed varZeta; True: var(residuals of the log-variances); a float: that share of var(logVar)


@dataclass
class RandomEffectType:
    str: Any
    v: float
    type: int = 1


@dataclass
class SummaryStatistics:
    m: Any
    v: Any


def BayesPR(r: int, v, name: str = "BayesPR") -> BayesPRType:
    """runTime.jl:36-45 — r: 1 per-SNP variance, 99 per chromosome, 9999 one common variance, else window size."""
    return BayesPRType(int(r), float(v) if np.ndim(v) == 0 else np.asarray(v, dtype=np.float64), name)


def BayesB(pi: float, v: float, name: str = "BayesB", estimatePi: bool = False) -> BayesBType:
    """runTime.jl:55-61 — pi is the proportion of SNPs INCLUDED."""
    return BayesBType(float(pi), float(v), name, bool(estimatePi))


def BayesC(pi: float, v: float, name: str = "BayesC", estimatePi: bool = False) -> BayesCType:
    """runTime.jl:70-76"""
    return BayesCType(float(pi), float(v), name, bool(estimatePi))


def BayesR(pi, class_, v: float, name: str = "BayesR", estimatePi: bool = False) -> BayesRType:
    """runTime.jl:87-93."""
    return BayesRType(np.asarray(pi, dtype=np.float64), np.asarray(class_, dtype=np.float64), float(v), name, bool(estimatePi))


def BayesRCpi(pi, class_, v: float, annot, name: str = "BayesRCπ", estimatePi: bool = False) -> BayesRCType:
    """runTime.jl:112 (BayesRCπ)."""
    return BayesRCType(np.asarray(pi, dtype=np.float64), np.asarray(class_, dtype=np.float64), float(v), np.asarray(annot, dtype=np.int32), name, bool(estimatePi))


def BayesRCplus(pi, class_, v: float, annot, name: str = "BayesRCplus", estimatePi: bool = False) -> BayesRCType:
    """runTime.jl:113."""
    return BayesRCType(np.asarray(pi, dtype=np.float64), np.asarray(class_, dtype=np.float64), float(v), np.asarray(annot, dtype=np.int32), name, bool(estimatePi))


def BayesLV(v: float, f, covariates, varZeta: float, name: str = "BayesLV", estimateVarZeta=False) -> BayesLogVarType:
    """runTime.jl:116-133: per-SNP variances with a log-linear model on SNP covariates."""
    return BayesLogVarType(float(v), f, covariates, float(varZeta), name, estimateVarZeta)


def _design_matrix(f, covariates) -> np.ndarray:
    """modelmatrix(f, covariates) (mme.jl:426) for the terms this mirror knows: "1" and numeric columns, joined by "+"."""
    if f is None:
        X = np.asarray(covariates, dtype=np.float64)
        return X[:, None] if X.ndim == 1 else X
    cols = []
    for t in [t.strip() for t in str(f).lstrip("~").split("+")]:
        if t == "1":
            cols.append(None)
        elif t in ("0", "-1", ""):
            continue
        else:
            cols.append(np.asarray(covariates[t], dtype=np.float64))
    n = next(len(c) for c in cols if c is not None) if any(c is not None for c in cols) else len(next(iter(covariates.values())))
    return np.column_stack([np.ones(n) if c is None else c for c in cols])


class LogVarModel:
    """Host-side state of a BayesLV marker set (mme.jl:418-440) and the model of the log-variances that follows the single-site loop in
    sampleBayesLV! (functions.jl:446-485).  O(p k) work per iteration next to the O(n p) sweep on the device."""

    def __init__(self, prior: BayesLogVarType, p: int, rng: np.random.Generator):
        X = _design_matrix(prior.f, prior.covariates)
        if X.shape[0] != p:
            raise ValueError("BayesLV: one row of covariates per SNP")
        self.covariates = X
        self.logVar = np.full(p, np.log(prior.v))                               # mme.jl:425
        self.c = rng.random(X.shape[1])                                         # mme.jl:429
        self.SNPVARRESID = rng.random(p)                                        # mme.jl:430
        CpC = X.T @ X
        CpC = CpC + np.eye(X.shape[1]) * np.min(np.abs(np.diag(CpC) / 10000.0))  # mme.jl:432-434
        self.iCpC = np.linalg.inv(CpC)                                          # mme.jl:437
        self.varZeta = float(prior.varZeta)
        self.estVarZeta = prior.estimateVarZeta
        self.trapped = 0

    def update(self, beta: np.ndarray, varBeta: np.ndarray, rng: np.random.Generator | None = None, u: np.ndarray | None = None,
               z: np.ndarray | None = None) -> None:
        """functions.jl:446-485 after the effects were sampled: slice-type update of every locus' variance (varBeta, in place), the
        covariate coefficients c ~ MvNormal(iCpC C'logVar, iCpC varZeta), the residuals and varZeta.  u (p, 4) uniforms in the order the
        reference calls rand() per locus, z (k) standard normals of the MvNormal draw; drawn from rng when not given."""
        p, k = self.covariates.shape
        u = rng.random((p, 4)) if u is None else u
        z = rng.standard_normal(k) if z is None else z
        vv = self.varZeta
        bi2 = beta * beta
        zeta = self.SNPVARRESID
        var_mui = self.logVar - zeta
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            c1 = varBeta ** -1.5 * u[:, 0]
            c2 = np.exp(-0.5 * bi2 / varBeta) * u[:, 1]
            c3 = np.exp(-0.5 * zeta * zeta / vv) * u[:, 2]
            temp = np.sqrt(-2.0 * vv * np.log(c3))
            lbound, rbound = np.exp(var_mui - temp), np.exp(var_mui + temp)
            r1 = np.exp((-2.0 / 3.0) * np.log(c1))
            rbound = np.where(r1 < rbound, r1, rbound)
            l1 = -0.5 * bi2 / np.log(c2)
            lbound = np.where(l1 > lbound, l1, lbound)
            ok = ~(lbound >= rbound)
            vari = lbound + u[:, 3] * (rbound - lbound)
        self.trapped = int(p - ok.sum())
        varBeta[ok] = vari[ok]
        self.logVar[ok] = np.log(vari[ok])
        meanC = self.iCpC @ (self.covariates.T @ self.logVar)
        cov = self.iCpC * vv
        cov = np.triu(cov) + np.triu(cov, 1).T                                  # Symmetric(): the upper triangle
        self.c[:] = meanC + np.linalg.cholesky(cov) @ z
        self.SNPVARRESID[:] = self.logVar - self.covariates @ self.c
        if isinstance(self.estVarZeta, float):
            self.varZeta = self.estVarZeta * float(np.var(self.logVar, ddof=1))
        elif self.estVarZeta is True:
            self.varZeta = float(np.var(self.SNPVARRESID, ddof=1))


def sampleBayesLV(sampler: "Sampler", set_id: int, model: LogVarModel, beta: np.ndarray, delta: np.ndarray, ycorr: np.ndarray, varE: float,
                  varBeta: np.ndarray, rng: np.random.Generator | None = None, u=None, z=None) -> None:
    """sampleBayesLV!(mSet,M,beta,delta,ycorr,varE,varBeta), functions.jl:421-486, host buffers mutated in place.  The single-site loop
    (functions.jl:431-443) is the BayesPR sweep with one region per locus and runs on the device (ngp_sweep); the variance draw the device
    appends to that sweep is discarded and the model of the log-variances (functions.jl:446-485) follows on the host."""
    scratch = varBeta.copy()
    sampler.sweep(set_id, ycorr, varE, beta, delta, scratch)
    model.update(beta, varBeta, rng, u, z)


def Random(str_: Any, v: float, type: int = 1) -> RandomEffectType:
    """runTime.jl:141-146"""
    return RandomEffectType(str_, float(v), type)


# ----------------------------------------------------------------------------- output (outFiles.jl)
def outMCMC(folder: str, thisVar: str, output) -> None:
    """Append row(s) to <folder>/<thisVar>Out, tab-delimited like writedlm (outFiles.jl:17-21)."""
    arr = np.asarray(output)
    if arr.ndim == 0:
        arr = arr.reshape(1, 1)
    elif arr.ndim == 1:
        arr = arr.reshape(1, -1)
    with open(os.path.join(folder, f"{thisVar}Out"), "a") as f:
        for row in arr:
            f.write("\t".join(_fmt(x) for x in row) + "\n")


def _fmt(x) -> str:
    if isinstance(x, (str, np.str_)):
        return str(x)
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    return repr(float(x))


def summaryMCMC(param: str, outFolder: str = os.path.join(os.getcwd(), "outMCMC")) -> np.ndarray:
    """Posterior mean of every column of <outFolder>/<param>Out (misc.jl:241-244)."""
    data = np.loadtxt(os.path.join(outFolder, f"{param}Out"), delimiter="\t", skiprows=1, ndmin=2)
    return data.mean(axis=0, keepdims=True)


def folderHandler(outFolder: str) -> None:
    """misc.jl:221-232: an existing output folder is removed."""
    if os.path.isdir(outFolder):
        shutil.rmtree(outFolder)
    os.mkdir(outFolder)


# ----------------------------------------------------------------------------- ingest (prepMatVec.jl:113-134)
def prep_snp(path_or_matrix) -> np.ndarray:
    """Space-delimited text, no header -> int8 codes (n,p) Fortran order.  Columns containing a missing
    or non-{0,1,2} value are dropped (the reference drops columns with `missing`, prepMatVec.jl:118).
    Centring (prepMatVec.jl:129) happens on device: the library stores raw codes + column means."""
    if isinstance(path_or_matrix, np.ndarray):
        raw = np.asarray(path_or_matrix, dtype=np.float64)
    else:
        packed, n, _ = read_text_packed(str(path_or_matrix))      # native reader: text -> 2-bit codes, no Float64 matrix
        return unpack2(packed, n)
    keep = ~np.isnan(raw).any(axis=0)
    raw = raw[:, keep]
    ok = np.isin(raw, (0.0, 1.0, 2.0)).all(axis=0)
    if not ok.all():
        raise ValueError("genotype codes must be 0/1/2 for packed storage; columns %s are not" % np.where(~ok)[0][:10])
    return np.asfortranarray(raw.astype(np.int8))


def read_text_packed(path: str):
    """Native reader of the reference's genotype text format (prepMatVec.jl:116-118) straight to 2-bit codes:
    returns (packed uint8 [ceil(n/4), p_kept] Fortran order, n, keep mask over the file's columns)."""
    lib = L.lib()
    n, p = C.c_int64(), C.c_int64()
    rc = lib.ngp_read_text_genotypes(path.encode(), C.byref(n), C.byref(p), None, 0, None, None)
    if rc != 0:
        raise L.NgpError(rc, f"cannot read {path}")
    ld = (n.value + 3) // 4
    packed = np.zeros((ld, p.value), dtype=np.uint8, order="F")
    keep = np.zeros(p.value, dtype=np.uint8)
    pk = C.c_int64()
    rc = lib.ngp_read_text_genotypes(path.encode(), C.byref(n), C.byref(p), _p(packed), ld, _p(keep), C.byref(pk))
    if rc == L.EDATA:
        raise ValueError(f"{path}: genotype codes must be 0/1/2 (packed storage) in rows of equal length")
    if rc != 0:
        raise L.NgpError(rc, f"cannot read {path}")
    return np.asfortranarray(packed[:, :pk.value]), n.value, keep.astype(bool)


def read_bed_packed(path: str, n: int, p: int, count_a1: bool = True):
    """PLINK .bed (SNP-major) -> 2-bit codes; variants with missing calls are dropped like columns with missing values."""
    lib = L.lib()
    ld = (n + 3) // 4
    packed = np.zeros((ld, p), dtype=np.uint8, order="F")
    keep = np.zeros(p, dtype=np.uint8)
    pk = C.c_int64()
    rc = lib.ngp_read_bed_genotypes(path.encode(), n, p, int(count_a1), _p(packed), ld, _p(keep), C.byref(pk))
    if rc != 0:
        raise L.NgpError(rc, f"cannot read {path} as a SNP-major PLINK .bed of {n} samples x {p} variants")
    return np.asfortranarray(packed[:, :pk.value]), keep.astype(bool)


def unpack2(packed: np.ndarray, n: int) -> np.ndarray:
    ld, p = packed.shape
    out = np.empty((n, p), dtype=np.int8, order="F")
    rc = L.lib().ngp_unpack2(_p(np.asfortranarray(packed)), n, p, ld, _p(out), n)
    if rc != 0:
        raise L.NgpError(rc, "ngp_unpack2")
    return out


def _tok(t: str) -> float:
    t = t.strip()
    if t == "" or t.upper() in ("NA", "NAN", "MISSING"):
        return float("nan")
    return float(t)


def prep2RegionData(outPutFolder: str | None, markerSet: str, mapFile, fixedRegSize: int) -> np.ndarray:
    """misc.jl:163-215.  mapFile: path of a delimited file with header snpID,snpOrder,chrID (or a dict of
    arrays).  Returns 0-based half-open region offsets (the reference returns Vector{UnitRange})."""
    if isinstance(mapFile, dict):
        snp_id, snp_order, chr_id = (np.asarray(mapFile[k]) for k in ("snpID", "snpOrder", "chrID"))
    else:
        with open(mapFile) as f:
            header = re.split(r"[,\t ]", f.readline().strip())
            cols = {h: [] for h in header}
            for line in f:
                for h, tok in zip(header, re.split(r"[,\t ]", line.strip())):
                    cols[h].append(tok)
        snp_id = np.asarray(cols["snpID"])
        snp_order = np.asarray(cols["snpOrder"])
        chr_id = np.asarray(cols["chrID"], dtype=np.int64)
    p = len(chr_id)
    if fixedRegSize == 99:
        group = chr_id.astype(np.int64).copy()
        n_reg = len(np.unique(chr_id))
    elif fixedRegSize == 9999:
        group = np.ones(p, dtype=np.int64)
        n_reg = 1
    else:
        group = np.empty(p, dtype=np.int64)
        acc = 0
        pos = 0
        _, first = np.unique(chr_id, return_index=True)
        for c in chr_id[np.sort(first)]:
            tot = int(np.sum(chr_id == c))
            nreg = -(-tot // fixedRegSize)
            g = np.repeat(np.arange(acc + 1, acc + nreg + 1), fixedRegSize)[:tot]
            group[pos:pos + tot] = g
            acc += nreg
            pos += tot
        n_reg = acc
    if outPutFolder is not None:
        with open(os.path.join(outPutFolder, f"groupInfo_{markerSet}.txt"), "w") as f:
            f.write("snpID\tsnpOrder\tchrID\tgroupID\n")
            for a, b, c, g in zip(snp_id, snp_order, chr_id, group):
                f.write(f"{a}\t{b}\t{c}\t{g}\n")
    offs = [int(np.searchsorted(group, g, "left")) for g in range(1, n_reg + 1)] + [p]
    return np.asarray(offs, dtype=np.int64)


# ----------------------------------------------------------------------------- the handle
def _p(a):
    return None if a is None else a.ctypes.data


class Sampler:
    """One chain on one GPU.  Thin, explicit wrapper over the C ABI (include/ngp.h)."""

    def __init__(self, device: int = 0, kernel: str = "blocked", block: int = 0, min_rows: int = 0, max_ctas: int = 0,
                 lookahead: int = 0, tile_stages: int = 0, near: int = 0, versions: int = 0, profile: bool = False,
                 refetch: int = -1, storage: str = "i8"):
        self._lib = L.lib()
        self.storage = {"i8": L.STORE_I8, "2bit": L.STORE_2BIT}[storage]
        hp = C.c_void_p()
        rc = self._lib.ngp_create(device, C.byref(hp))
        if rc != 0:
            raise L.NgpError(rc, self._lib.ngp_last_error(None).decode())
        self._h = hp
        self.device = device
        self.sets: dict[int, dict] = {}
        self.n = 0
        self._keep = []
        self.configure(L.CFG_KERNEL, L.KERNEL_BLOCKED if kernel == "blocked" else L.KERNEL_LITERAL)
        self._configured = set()
        if block:
            self.configure(L.CFG_BLOCK, block)
            self._configured.add("block")
        if min_rows:
            self.configure(L.CFG_MIN_ROWS, min_rows)
        if max_ctas:
            self.configure(L.CFG_MAX_CTAS, max_ctas)
        if lookahead:
            self.configure(L.CFG_LOOKAHEAD, lookahead)
        if tile_stages:
            self.configure(L.CFG_TILE_STAGES, tile_stages)
        if near:
            self.configure(L.CFG_NEAR, near)
        if versions:
            self.configure(L.CFG_VERSIONS, versions)
        if profile:
            self.configure(L.CFG_PROFILE, 1)
        if refetch != -1:
            self.configure(L.CFG_REFETCH, refetch)

    # -- plumbing
    def _ck(self, rc: int) -> None:
        if rc != 0:
            raise L.NgpError(rc, self._lib.ngp_last_error(self._h).decode())

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.ngp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def configure(self, key: int, value: int) -> None:
        self._ck(self._lib.ngp_configure(self._h, key, value))

    def set_kernel(self, kernel: str) -> None:
        self.configure(L.CFG_KERNEL, L.KERNEL_BLOCKED if kernel == "blocked" else L.KERNEL_LITERAL)

    def set_stream(self, cuda_stream: int | None) -> None:
        self._ck(self._lib.ngp_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    # -- data
    def upload_genotypes(self, set_id: int, data: np.ndarray, fmt: int | None = None, n: int | None = None) -> None:
        if fmt is None:
            fmt = L.GENO_F64 if data.dtype == np.float64 else L.GENO_I8
        if fmt == L.GENO_PACKED2:
            assert n is not None and data.dtype == np.uint8
            data = np.asfortranarray(data)
            ld, p = data.shape
        else:
            data = np.asfortranarray(data, dtype=np.float64 if fmt == L.GENO_F64 else np.int8)
            n, p = data.shape
            ld = n
        self._ck(self._lib.ngp_upload_genotypes(self._h, set_id, n, p, _p(data), fmt, ld, self.storage))
        self.n = n
        self.sets[set_id] = {"p": p}

    def synth_genotypes(self, set_id: int, n: int, p: int, seed: int, thr0: np.ndarray, thr1: np.ndarray) -> None:
        thr0 = np.ascontiguousarray(thr0, dtype=np.uint32)
        thr1 = np.ascontiguousarray(thr1, dtype=np.uint32)
        self._ck(self._lib.ngp_synth_genotypes(self._h, set_id, n, p, C.c_uint64(seed), _p(thr0), _p(thr1), self.storage))
        self.n = n
        self.sets[set_id] = {"p": p}

    def synth_genotypes_rows(self, set_id: int, row0: int, n: int, p: int, seed: int, thr0: np.ndarray, thr1: np.ndarray) -> None:
        thr0 = np.ascontiguousarray(thr0, dtype=np.uint32)
        thr1 = np.ascontiguousarray(thr1, dtype=np.uint32)
        self._ck(self._lib.ngp_synth_genotypes_rows(self._h, set_id, row0, n, p, C.c_uint64(seed), _p(thr0), _p(thr1), self.storage))
        self.n = n
        self.sets[set_id] = {"p": p}

    # -- row-sharded chain (include/ngp.h: ngp_shard_*)
    def shard_init(self, rank: int, world: int) -> None:
        self._ck(self._lib.ngp_shard_init(self._h, rank, world))

    def shard_export(self) -> bytes:
        info = L.ShardInfo()
        self._ck(self._lib.ngp_shard_export(self._h, C.byref(info)))
        return bytes(info)

    def shard_attach(self, infos: list[bytes]) -> None:
        arr = (L.ShardInfo * len(infos))(*[L.ShardInfo.from_buffer_copy(b) for b in infos])
        self._ck(self._lib.ngp_shard_attach(self._h, arr))

    def column_sums(self, set_id: int):
        p = self.sets[set_id]["p"]
        a, b = np.empty(p, dtype=np.int64), np.empty(p, dtype=np.int64)
        self._ck(self._lib.ngp_get_column_sums(self._h, set_id, _p(a), _p(b)))
        return a, b

    def gram(self, set_id: int) -> np.ndarray:
        """banded raw Gram of this handle's rows (int32): ngp_get_gram"""
        cnt = C.c_int64()
        self._ck(self._lib.ngp_gram_size(self._h, set_id, C.byref(cnt)))
        out = np.empty(cnt.value, dtype=np.int32)
        self._ck(self._lib.ngp_get_gram(self._h, set_id, _p(out)))
        return out

    def set_gram(self, set_id: int, gram: np.ndarray) -> None:
        gram = np.ascontiguousarray(gram, dtype=np.int32)
        cnt = C.c_int64()
        self._ck(self._lib.ngp_gram_size(self._h, set_id, C.byref(cnt)))
        assert gram.size == cnt.value
        self._ck(self._lib.ngp_set_gram(self._h, set_id, _p(gram)))

    def set_column_sums(self, set_id: int, n_total: int, colsum: np.ndarray, colsumsq: np.ndarray) -> None:
        a = np.ascontiguousarray(colsum, dtype=np.int64)
        b = np.ascontiguousarray(colsumsq, dtype=np.int64)
        self._ck(self._lib.ngp_set_column_sums(self._h, set_id, n_total, _p(a), _p(b)))

    def download_genotypes(self, set_id: int, j0: int = 0, j1: int | None = None) -> np.ndarray:
        p = self.sets[set_id]["p"]
        j1 = p if j1 is None else j1
        out = np.empty((self.n, j1 - j0), dtype=np.int8, order="F")
        self._ck(self._lib.ngp_download_genotypes(self._h, set_id, j0, j1, _p(out)))
        return out

    def column_stats(self, set_id: int):
        p = self.sets[set_id]["p"]
        mean, mpm = np.empty(p), np.empty(p)
        self._ck(self._lib.ngp_get_column_stats(self._h, set_id, _p(mean), _p(mpm)))
        return mean, mpm

    # -- model
    def set_phenotype(self, y: np.ndarray) -> None:
        y = np.ascontiguousarray(y, dtype=np.float64)
        self._ck(self._lib.ngp_set_phenotype(self._h, _p(y), len(y)))

    def set_residual_prior(self, df_e: float, scale_e: float) -> None:
        self._ck(self._lib.ngp_set_residual_prior(self._h, df_e, scale_e))

    def set_residual_weights(self, w) -> None:
        """E.iVarStr of a "D" residual structure (mme.jl:70-73): one positive weight per individual; None returns to "I"."""
        if w is None:
            self._ck(self._lib.ngp_set_residual_weights(self._h, None, 0))
            return
        w = np.ascontiguousarray(w, dtype=np.float64)
        self._ck(self._lib.ngp_set_residual_weights(self._h, _p(w), len(w)))

    def set_intercept(self, enabled: bool = True, lhs0: float = 0.0, rhs0: float = 0.0) -> None:
        self._ck(self._lib.ngp_set_intercept(self._h, int(enabled), lhs0, rhs0))

    def set_fixed_effects(self, sets: list) -> None:
        """Fixed-effect terms besides the intercept, sampled after it in this order (functions.jl:22-54).
        sets: arrays (n,) / (n, c), or (array, lhs0, rhs0) for a single column with prior information."""
        arr = (L.FixedSet * max(len(sets), 1))()
        keep = []
        self.fixed_cols = 0
        for i, s in enumerate(sets):
            data, l0, r0 = (s if isinstance(s, tuple) else (s, 0.0, 0.0))
            data = np.asfortranarray(data, dtype=np.float64)
            if data.ndim == 1:
                data = np.asfortranarray(data[:, None])
            keep.append(data)
            arr[i].n_cols, arr[i].data, arr[i].lhs0, arr[i].rhs0 = data.shape[1], _p(data), l0, r0
            self.fixed_cols += data.shape[1]
        self._ck(self._lib.ngp_set_fixed_effects(self._h, len(sets), arr))

    def fixed_effects(self) -> np.ndarray:
        b = np.empty(self.fixed_cols)
        self._ck(self._lib.ngp_get_fixed_effects(self._h, _p(b)))
        return b

    def set_fixed_replay(self, logs: list[dict]) -> None:
        """logs: the oracle's per-iteration logs; their "z_fx" entries (one array per fixed set) are concatenated per iteration."""
        z = np.ascontiguousarray(np.stack([np.concatenate([np.atleast_1d(v) for v in g["z_fx"]]) for g in logs]), dtype=np.float64)
        self._ck(self._lib.ngp_set_fixed_replay(self._h, len(logs), _p(z)))

    def set_prior(self, set_id: int, method: int, df: float, scale: float, var_init: float, pi_in: float = 0.0,
                  est_pi: bool = False, region_off: np.ndarray | None = None, lhs0: np.ndarray | None = None,
                  rhs0: np.ndarray | None = None, v_class: np.ndarray | None = None, pi_class: np.ndarray | None = None) -> None:
        pr = L.Prior()
        if method == L.BAYESR:
            v_class = np.ascontiguousarray(v_class, dtype=np.float64)
            pi_class = np.ascontiguousarray(pi_class, dtype=np.float64)
            pr.n_class, pr.v_class, pr.pi_class = len(v_class), _p(v_class), _p(pi_class)
        pr.method, pr.est_pi, pr.df, pr.scale, pr.var_init, pr.pi_in = method, int(est_pi), df, scale, var_init, pi_in
        p = self.sets[set_id]["p"]
        if region_off is not None:
            region_off = np.ascontiguousarray(region_off, dtype=np.int64)
            pr.n_regions, pr.region_off = len(region_off) - 1, _p(region_off)
        if lhs0 is not None:
            lhs0 = np.ascontiguousarray(lhs0, dtype=np.float64)
            pr.lhs0 = _p(lhs0)
        if rhs0 is not None:
            rhs0 = np.ascontiguousarray(rhs0, dtype=np.float64)
            pr.rhs0 = _p(rhs0)
        self._ck(self._lib.ngp_set_prior(self._h, set_id, C.byref(pr)))
        nvar = (len(region_off) - 1 if region_off is not None else 1) if method == L.BAYESPR else (p if method == L.BAYESB else 1)
        self.sets[set_id].update(method=method, nvar=nvar, est_pi=bool(est_pi), n_class=(len(v_class) if method == L.BAYESR else 0))

    def set_joint_prior(self, set_ids: list[int], df: float, scale: np.ndarray, var_init: np.ndarray,
                        region_off: np.ndarray | None = None) -> None:
        """(:M1,:M2,...) => BayesPR(r, V): the member sets' effects are drawn jointly per locus (mme.jl:448-489)."""
        k = len(set_ids)
        pr = L.JointPrior()
        pr.k, pr.df = k, df
        for b, sid in enumerate(set_ids):
            pr.set_id[b] = sid
        scale = np.ascontiguousarray(scale, dtype=np.float64).reshape(k, k)
        var_init = np.ascontiguousarray(var_init, dtype=np.float64).reshape(k, k)
        pr.scale, pr.var_init = _p(scale), _p(var_init)
        if region_off is not None:
            region_off = np.ascontiguousarray(region_off, dtype=np.int64)
            pr.n_regions, pr.region_off = len(region_off) - 1, _p(region_off)
        self._ck(self._lib.ngp_set_joint_prior(self._h, C.byref(pr)))
        self.joint = {"k": k, "sets": list(set_ids), "p": self.sets[set_ids[0]]["p"],
                      "n_regions": len(region_off) - 1 if region_off is not None else 1}

    def set_joint_replay(self, logs: list[dict]) -> None:
        """logs: per iteration the oracle's tuple variate log {z [p,k], iw_chi2 [R,k], iw_z [R,k,k]}."""
        z = np.ascontiguousarray(np.stack([g["z"] for g in logs]), dtype=np.float64)
        c2 = np.ascontiguousarray(np.stack([g["iw_chi2"] for g in logs]), dtype=np.float64)
        zl = np.ascontiguousarray(np.stack([g["iw_z"] for g in logs]), dtype=np.float64)
        self._ck(self._lib.ngp_set_joint_replay(self._h, len(logs), _p(z), _p(c2), _p(zl)))

    def joint_sweep(self, ycorr: np.ndarray, varE: float, beta: np.ndarray, varBeta: np.ndarray) -> None:
        """sampleBayesPR!(mSet::Tuple, M, beta, delta, ycorr, varE, varBeta): host arrays, mutated in place."""
        for a in (ycorr, beta, varBeta):
            assert a.dtype == np.float64 and a.flags.c_contiguous
        self._ck(self._lib.ngp_joint_sweep(self._h, _p(ycorr), varE, _p(beta), _p(varBeta)))

    def joint_state(self) -> dict:
        j = self.joint
        beta = np.empty((j["k"], j["p"]))
        vb = np.empty((j["n_regions"], j["k"], j["k"]))
        self._ck(self._lib.ngp_get_joint_state(self._h, _p(beta), _p(vb)))
        return {"beta": beta, "varBeta": vb}

    def set_rc_prior(self, set_id: int, plus: bool, df: float, scale: float, var_init: float, v_class, pi_class, annot, est_pi: bool = False) -> None:
        """BayesRCpi (plus=False) / BayesRCplus (plus=True): annot = (p, n_annot) integer matrix (mme.jl:385-418)."""
        v_class = np.ascontiguousarray(v_class, dtype=np.float64)
        pi_class = np.ascontiguousarray(pi_class, dtype=np.float64)
        annot = np.ascontiguousarray(annot, dtype=np.int32)
        assert annot.shape[0] == self.sets[set_id]["p"] and len(pi_class) == len(v_class)
        pr = L.RCPrior()
        pr.plus, pr.est_pi, pr.df, pr.scale, pr.var_init = int(plus), int(est_pi), df, scale, var_init
        pr.n_class, pr.n_annot, pr.v_class, pr.pi_class, pr.annot = len(v_class), annot.shape[1], _p(v_class), _p(pi_class), _p(annot)
        self._ck(self._lib.ngp_set_rc_prior(self._h, set_id, C.byref(pr)))
        self.sets[set_id].update(method=L.BAYESRCPLUS if plus else L.BAYESRCPI, nvar=annot.shape[1], est_pi=bool(est_pi), n_class=0,
                                 rc=(annot.shape[1], len(v_class)))

    def rc_state(self, set_id: int) -> dict:
        nA, nc = self.sets[set_id]["rc"]
        p = self.sets[set_id]["p"]
        cat, prob, pi = np.zeros(p, dtype=np.int64), np.zeros((p, nA)), np.zeros((nA, nc))
        self._ck(self._lib.ngp_get_rc_state(self._h, set_id, _p(cat), _p(prob), _p(pi)))
        return {"annot_cat": cat, "annot_prob": prob, "piHat": pi}

    def set_rc_replay(self, set_id: int, logs: list[dict]) -> None:
        """logs: the oracle's BayesRCOracle.sweep() variate dicts, one per iteration (after set_replay with chi2_e / z_mu)"""
        st = lambda k: np.ascontiguousarray(np.stack([g[k] for g in logs]), dtype=np.float64)
        arrs = {k: st(k) for k in ("u_annot", "dirp", "u", "z", "chi2_b", "dir_pi")}
        self._keep.append(arrs)
        self._ck(self._lib.ngp_set_rc_replay(self._h, set_id, len(logs), *(_p(arrs[k]) for k in ("u_annot", "dirp", "u", "z", "chi2_b", "dir_pi"))))

    def set_rng(self, seed: int, chain_id: int = 0) -> None:
        self._ck(self._lib.ngp_set_rng(self._h, C.c_uint64(seed), chain_id))
        self.seed, self.chain_id = int(seed), int(chain_id)

    def set_replay(self, logs: list[dict] | None) -> None:
        """logs: list (one per iteration) of the oracle's variate-log dicts
        {chi2_e, z_mu, sets:[{u,z,chi2_b,beta_pi}]}; None switches back to Philox."""
        if logs is None:
            self._ck(self._lib.ngp_set_replay(self._h, None))
            return
        rp = L.Replay()
        ni = len(logs)
        rp.n_iter, rp.n_sets = ni, len(logs[0]["sets"])
        keep = []
        chi2_e = np.array([g["chi2_e"] for g in logs], dtype=np.float64)
        z_mu = np.array([g.get("z_mu", 0.0) for g in logs], dtype=np.float64)
        keep += [chi2_e, z_mu]
        rp.chi2_e, rp.z_mu = _p(chi2_e), _p(z_mu)
        for s in range(rp.n_sets):
            if "dir_pi" in logs[0]["sets"][s]:                      # BayesR log: u [p, n_class], the Dirichlet draw instead of the Beta draw
                for g in logs:
                    g["sets"][s].setdefault("beta_pi", g["sets"][s]["dir_pi"])
            u = np.ascontiguousarray(np.stack([g["sets"][s]["u"] for g in logs]), dtype=np.float64)
            z = np.ascontiguousarray(np.stack([g["sets"][s]["z"] for g in logs]), dtype=np.float64)
            cb = np.ascontiguousarray(np.stack([g["sets"][s]["chi2_b"] for g in logs]), dtype=np.float64)
            bp = np.ascontiguousarray(np.array([g["sets"][s]["beta_pi"] for g in logs], dtype=np.float64))
            keep += [u, z, cb, bp]
            rp.u[s], rp.z[s], rp.chi2_b[s], rp.beta_pi[s] = _p(u), _p(z), _p(cb), _p(bp)
        self._ck(self._lib.ngp_set_replay(self._h, C.byref(rp)))

    # -- sampling
    def run(self, n_iter: int = 1) -> None:
        self._ck(self._lib.ngp_run(self._h, n_iter))

    def set_var_beta(self, set_id: int, varBeta: np.ndarray) -> None:
        """Overwrite the effect variances of one set (nvar values), nothing else."""
        varBeta = np.ascontiguousarray(varBeta, dtype=np.float64)
        assert varBeta.shape == (self.sets[set_id]["nvar"],)
        self._ck(self._lib.ngp_set_var_beta(self._h, set_id, _p(varBeta)))

    def set_marker_summary(self, set_id: int, lhs0: np.ndarray | None, rhs0: np.ndarray | None) -> None:
        """Replace only the per-marker prior information (M[pSet][:lhs] / [:rhs], mme.jl:314-322) of a set; the chain state stays."""
        if lhs0 is not None:
            lhs0 = np.ascontiguousarray(lhs0, dtype=np.float64)
            assert lhs0.shape == (self.sets[set_id]["p"],)
        if rhs0 is not None:
            rhs0 = np.ascontiguousarray(rhs0, dtype=np.float64)
            assert rhs0.shape == (self.sets[set_id]["p"],)
        self._ck(self._lib.ngp_set_marker_summary(self._h, set_id, _p(lhs0), _p(rhs0)))

    def sweep(self, set_id: int, ycorr: np.ndarray, varE: float, beta: np.ndarray, delta: np.ndarray,
              varBeta: np.ndarray, piHat: np.ndarray | None = None) -> None:
        """M[mSet].funct(mSet,M,beta,delta,ycorr,varE,varBeta): host arrays, mutated in place."""
        for a, dt in ((ycorr, np.float64), (beta, np.float64), (delta, np.int64), (varBeta, np.float64)):
            assert a.dtype == dt and a.flags.c_contiguous
        self._ck(self._lib.ngp_sweep(self._h, set_id, _p(ycorr), varE, _p(beta), _p(delta), _p(varBeta), _p(piHat)))

    def state(self, want_e: bool = True) -> dict:
        st = L.State()
        e = np.empty(self.n) if want_e else None
        st.e = _p(e)
        bufs = {}
        for s, info in self.sets.items():
            if "method" not in info:
                continue
            b, d, v = np.empty(info["p"]), np.empty(info["p"], dtype=np.int64), np.empty(info["nvar"])
            bufs[s] = (b, d, v)
            st.beta[s], st.delta[s], st.varBeta[s] = _p(b), _p(d), _p(v)
        self._ck(self._lib.ngp_get_state(self._h, C.byref(st)))
        out = {"e": e, "mu": st.mu, "varE": st.varE, "iter": st.iter,
               "sets": {s: {"beta": b, "delta": d, "varBeta": v, "piHat": np.array([st.pi[s][0], st.pi[s][1]])}
                        for s, (b, d, v) in bufs.items()}}
        for s in bufs:
            nc = self.sets[s].get("n_class", 0)
            if nc:
                ph = np.empty(nc)
                self._ck(self._lib.ngp_get_class_pi(self._h, s, _p(ph)))
                out["sets"][s]["piHat"] = ph
        return out

    def set_state(self, e=None, mu=0.0, varE=0.0, iter=0, sets: dict | None = None) -> None:
        st = L.State()
        keep = []
        if e is not None:
            e = np.ascontiguousarray(e, dtype=np.float64); keep.append(e); st.e = _p(e)
        st.mu, st.varE, st.iter = mu, varE, iter
        for s, d in (sets or {}).items():
            for key, dt, fld in (("beta", np.float64, st.beta), ("delta", np.int64, st.delta), ("varBeta", np.float64, st.varBeta)):
                if key in d:
                    a = np.ascontiguousarray(d[key], dtype=dt); keep.append(a); fld[s] = _p(a)
            ph = d.get("piHat", (0.5, 0.5))
            st.pi[s][0], st.pi[s][1] = ph[0], ph[1]
        self._ck(self._lib.ngp_set_state(self._h, C.byref(st)))

    def reset_posterior(self) -> None:
        self._ck(self._lib.ngp_reset_posterior(self._h))

    def posterior(self, set_id: int) -> dict:
        p = self.sets[set_id]["p"]
        n = C.c_int64()
        sb, sb2, sd = np.empty(p), np.empty(p), np.empty(p)
        self._ck(self._lib.ngp_get_posterior(self._h, set_id, C.byref(n), _p(sb), _p(sb2), _p(sd)))
        k = max(n.value, 1)
        return {"n": n.value, "mean_beta": sb / k, "mean_beta2": sb2 / k, "mean_delta": sd / k}

    def timing(self) -> dict:
        t = L.Timing()
        self._ck(self._lib.ngp_get_timing(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in L.Timing._fields_}

    def profile(self) -> np.ndarray:
        """(ctas, 32) int64 cycle counters of the last launch, see ngp_get_profile."""
        out = np.zeros((160, 32), dtype=np.int64)
        nc = self._lib.ngp_get_profile(self._h, _p(out), 160)
        if nc < 0:
            self._ck(nc)
        return out[:nc]

    def trace(self) -> np.ndarray:
        """(steps, 2) int64: start clock and cycles waited for r_base of the chain warp's first 2048 steps (instrumented kernel)."""
        out = np.zeros((2048, 2), dtype=np.int64)
        nc = self._lib.ngp_get_trace(self._h, _p(out), 4096)
        if nc < 0:
            self._ck(nc)
        return out

    def debug_variates(self, set_id: int, it: int, purpose: int, df: float, n: int) -> np.ndarray:
        out = np.empty(n)
        self._ck(self._lib.ngp_debug_variates(self._h, set_id, it, purpose, df, n, _p(out)))
        return out


def sampleLambda2(sampler: Sampler, set_id: int, Lambda2: np.ndarray, yCorr: np.ndarray, var_tau: np.ndarray, varE: float,
                  pMeans: np.ndarray) -> None:
    """sampleΛ2!(Λ2,Xc,yCorr,σ2τ,σ2ϵ,pMeans) of the gene-regulatory-network sampler (GRN.jl:150-164), in place: for every gene g one
    single-site sweep over the SNPs of the uploaded (device-centred, GRN.jl:23) marker set with
        beta_q ~ N((x_q'yCorr_g + alpha_g pMeans_g) / x_q'x_q, varE / x_q'x_q),   alpha_g = varE / var_tau[g].
    That is the BayesPR sweep of ngp_sweep with an improper effect prior (varBeta = +Inf, so 1/varBeta = 0: the reference's LHS holds no
    prior precision) and the right-hand-side offset rhs0 = alpha_g pMeans_g / varE = pMeans_g / var_tau[g] on every marker.
    Lambda2 (genes, SNPs) and yCorr (genes, individuals) are C-contiguous float64 and are mutated row by row; the set needs a one-region
    BayesPR prior (set_prior(set_id, BAYESPR, ...)).  Draws: the handle's counter stream, one iteration index per gene sweep."""
    info = sampler.sets[set_id]
    assert info.get("method") == L.BAYESPR and info["nvar"] == 1, "sampleLambda2 needs a one-region BayesPR set"
    assert Lambda2.dtype == np.float64 and Lambda2.flags.c_contiguous and yCorr.dtype == np.float64 and yCorr.flags.c_contiguous
    nGenes, nSNPs = Lambda2.shape
    assert yCorr.shape == (nGenes, sampler.n) and nSNPs == info["p"]
    delta = np.ones(nSNPs, dtype=np.int64)
    for g in range(nGenes):
        sampler.set_marker_summary(set_id, None, np.full(nSNPs, pMeans[g] / var_tau[g]))
        vb = np.array([np.inf])
        sampler.sweep(set_id, yCorr[g], varE, Lambda2[g], delta, vb)


class ShardedChain:
    """One chain whose individuals are row-sharded over several handles of THIS process (one per device, or several on one
    device for testing).  Each handle runs the same persistent per-marker kernel; the kernels of all shards must be
    co-resident, so every ngp_run is issued from one host thread per shard.  With one process per GPU use the Sampler
    methods directly and exchange shard_export() blobs / column sums with torch.distributed (bench.py --sharded)."""

    def __init__(self, devices: list[int], max_ctas: int = 0, min_rows: int = 0, kernel: str = "literal", block: int = 0, lookahead: int = 0, **geom):
        """kernel = "literal": one reduction per marker; "blocked": the look-ahead kernel, B partial sums per block pushed to every rank
        (all ranks must share the block size and the look-ahead: pass them when the shards' row counts differ much)."""
        self.world = len(devices)
        self.kernel = kernel
        self.shards = [Sampler(d, kernel=kernel, max_ctas=max_ctas, min_rows=min_rows, block=block, lookahead=lookahead, **geom) for d in devices]
        for r, s in enumerate(self.shards):
            s.shard_init(r, self.world)
        self.rows: list[tuple[int, int]] = []

    def split_rows(self, n: int) -> list[tuple[int, int]]:
        per = -(-n // self.world)
        per = -(-per // 4) * 4                                   # shard boundaries on multiples of 4 rows
        self.rows = [(min(n, r * per), min(n, (r + 1) * per)) for r in range(self.world)]
        return self.rows

    def upload_genotypes(self, set_id: int, codes: np.ndarray) -> None:
        n = codes.shape[0]
        if not self.rows:
            self.split_rows(n)
        for s, (a, b) in zip(self.shards, self.rows):
            s.upload_genotypes(set_id, np.asfortranarray(codes[a:b]))
        self._finish(set_id, n)

    def _finish(self, set_id: int, n: int) -> None:
        if not self.shards[0].__dict__.get("_attached"):
            infos = [s.shard_export() for s in self.shards]
            for s in self.shards:
                s.shard_attach(infos)
                s._attached = True
        sums = [s.column_sums(set_id) for s in self.shards]
        cs, css = sum(x[0] for x in sums), sum(x[1] for x in sums)
        for s in self.shards:
            s.set_column_sums(set_id, n, cs, css)
        if self.kernel == "blocked":                      # the banded Gram is a sum over individuals
            grams = [s.gram(set_id) for s in self.shards]
            if len({g.shape for g in grams}) != 1:
                raise ValueError("ShardedChain: the shards chose different block sizes / look-aheads; pass block= and lookahead=")
            total = np.sum(grams, axis=0, dtype=np.int64).astype(np.int32)
            for s in self.shards:
                s.set_gram(set_id, total)

    def each(self, fn) -> list:
        return [fn(s, r) for r, s in enumerate(self.shards)]

    def set_phenotype(self, y: np.ndarray) -> None:
        for s, (a, b) in zip(self.shards, self.rows):
            s.set_phenotype(y[a:b])

    def run(self, n_iter: int = 1) -> None:
        """Shards on ONE device run as one cooperative grid (ngp_run_group): kernels that wait for one another must not be separate
        launches on one GPU.  Shards on distinct devices: one launch per device, each from its own host thread; every shard is
        validated (a dry ngp_get_timing / state check) before any is launched, and a rank that never arrives ends the others' waits
        after 4 s (NGP_ETIMEOUT)."""
        devs = {s.device for s in self.shards}
        if len(devs) == 1:
            arr = (C.c_void_p * self.world)(*[s._h for s in self.shards])
            rc = L.lib().ngp_run_group(arr, self.world, n_iter)
            if rc != L.OK:
                raise L.NgpError(rc, (L.lib().ngp_last_error(self.shards[0]._h) or b"").decode())
            return
        if len(devs) != self.world:
            raise ValueError("ShardedChain: shards must be on one device (one grid) or on pairwise distinct devices")
        import threading
        errs: list = []

        def work(s):
            try:
                s.run(n_iter)
            except Exception as ex:  # noqa: BLE001
                errs.append(ex)
        th = [threading.Thread(target=work, args=(s,)) for s in self.shards]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]

    def state(self) -> dict:
        sts = [s.state() for s in self.shards]
        out = sts[0]
        out["e"] = np.concatenate([st["e"] for st in sts])
        out["shards"] = sts
        return out

    def close(self) -> None:
        for s in self.shards:
            s.close()


# ----------------------------------------------------------------------------- getMME / runSampler / runLMEM
@dataclass
class MarkerTerm:
    name: str
    codes: Any                   # int8 (n,p), or None when `packed` is given
    map: Any = None              # map file path / dict or None
    levels: list = field(default_factory=list)
    packed: Any = None           # uint8 (ceil(n/4), p) 2-bit codes from the native readers (NGP_GENO_PACKED2)
    n: int = 0                   # individuals (needed with `packed`)

    @property
    def p(self) -> int:
        return self.packed.shape[1] if self.packed is not None else self.codes.shape[1]

    def upload(self, sampler: "Sampler", sid: int) -> None:
        if self.packed is not None:
            sampler.upload_genotypes(sid, self.packed, fmt=L.GENO_PACKED2, n=self.n)
        else:
            sampler.upload_genotypes(sid, self.codes)


def getMME(sampler: Sampler, Y: np.ndarray, M: list[MarkerTerm], priorVCV: dict, summaryStat: dict | None, outPut: str | None,
           intercept: bool = True, fixed: list | None = None):
    """mme.getMME! for intercept + marker sets: derives df/scale/regions (mme.jl:87-94, 324-373, 492-520), uploads
    everything through the C ABI and writes the header rows of the output files (mme.jl:543-595)."""
    summaryStat = summaryStat or {}
    if "e" not in priorVCV:
        priorVCV = dict(priorVCV, e=Random("I", 100.0))                       # mme.jl:79-84
    e_prior = priorVCV["e"]
    weights = None
    if isinstance(e_prior.str, (list, tuple, np.ndarray)) and len(e_prior.str) > 0:      # "D": mme.jl:70-73, iVarStr = inv.(D)
        dvec = np.asarray(e_prior.str, dtype=np.float64)
        if dvec.shape != (len(Y),):
            raise ValueError("the \"D\" vector of priorVCV[:e] needs one entry per record")
        weights = 1.0 / dvec
        if any(isinstance(k, tuple) for k in priorVCV):
            raise NotImplementedError("weighted residuals: not with a tuple of marker sets (stays in Julia)")
    elif not (isinstance(e_prior.str, str) and e_prior.str in ("I", "")) and not (e_prior.str is None or len(e_prior.str) == 0):
        raise ValueError("provide a valid prior var-cov structure (\"I\", \"D\" or leave it empty \"[]\") for \"e\" ")   # mme.jl:76
    df_e = 4.0
    scale_e = 0.0005 if e_prior.v == 0.0 else e_prior.v * (df_e - 2.0) / df_e  # mme.jl:87-94
    tuples = [k for k in priorVCV if isinstance(k, tuple)]
    if tuples:                                                                  # (:M1,:M2) => BayesPR(r, V): mme.jl:448-489
        return _getMME_tuple(sampler, Y, M, priorVCV, tuples, outPut, intercept, df_e, scale_e)
    info = []
    # every marker prior BayesPR (dense updates) on panels of more than 512 rows per CTA: blocks of 16 with resident tiles beat the default
    # blocks of 64 on the refetch ring (ngp_api.cu: apply_ring_geometry); the block size is fixed at the first upload, so say it now
    n_rows = len(Y)
    if (M and not sampler.sets and n_rows > 512 * 147 and "block" not in getattr(sampler, "_configured", ())
            and all((priorVCV.get(t.name) is None) or getattr(priorVCV.get(t.name), "name", "") == "BayesPR" for t in M)):
        sampler.configure(L.CFG_BLOCK, 16)
    for sid, term in enumerate(M):
        term.upload(sampler, sid)
        p = term.p
        pr = priorVCV.get(term.name)
        lhs0 = rhs0 = None
        if term.name in summaryStat:                                           # mme.jl:314-322
            ss = summaryStat[term.name]
            v = np.asarray(ss.v, dtype=np.float64)
            v = np.diag(v) if v.ndim == 2 else np.broadcast_to(v, (p,))
            with np.errstate(divide="ignore", invalid="ignore"):
                lhs0 = 1.0 / v
                rhs0 = lhs0 * np.broadcast_to(np.asarray(ss.m, dtype=np.float64), (p,))
            lhs0 = np.where(np.isinf(lhs0), 0.0, lhs0)
            rhs0 = np.where(np.isnan(rhs0), 0.0, rhs0)
        if pr is None:                                                         # mme.jl:324-329, 503-504, 518
            method, v, region_off, pi, est = L.BAYESPR, 0.05, None, 0.0, False
            name = "BayesPR"
        else:
            name, v = pr.name, pr.v
            pi, est = getattr(pr, "pi", 0.0), getattr(pr, "estimatePi", False)
            if name == "BayesPR":
                method = L.BAYESPR
                if term.map is None or (isinstance(term.map, str) and term.map == ""):   # mme.jl:334-343
                    if pr.r == 1:
                        region_off = np.arange(p + 1, dtype=np.int64)
                    elif pr.r == 9999:
                        region_off = None
                    else:
                        raise ValueError("Please enter a valid region size (1 or 9999)")
                else:
                    region_off = prep2RegionData(outPut, term.name, term.map, pr.r)      # mme.jl:345-347
            elif name == "BayesB":
                method, region_off = L.BAYESB, None
            elif name == "BayesC":
                method, region_off = L.BAYESC, None
            elif name == "BayesR":                                              # mme.jl:374-383
                method, region_off, pi = L.BAYESR, None, 0.0
            elif name in ("BayesRCπ", "BayesRCplus"):                             # mme.jl:385-418
                method, region_off, pi = (L.BAYESRCPI if name == "BayesRCπ" else L.BAYESRCPLUS), None, 0.0
            elif name == "BayesLV":                                             # mme.jl:418-440: theseRegions = [r:r for r in 1:p]
                method, region_off, pi = L.BAYESPR, np.arange(p + 1, dtype=np.int64), 0.0
            else:
                raise NotImplementedError(f"{name} stays in Julia: SURVEY §8(f2)")
        df = 3.0 + 1.0                                                          # mme.jl:493 (scalar v)
        scale = v * (df - 2.0) / df                                             # mme.jl:501
        extra = dict(v_class=pr.class_, pi_class=pr.pi) if name == "BayesR" else {}
        if name in ("BayesRCπ", "BayesRCplus"):
            if lhs0 is not None:
                raise NotImplementedError("summary-statistic priors with BayesRCπ / BayesRCplus stay in Julia")
            sampler.set_rc_prior(sid, name == "BayesRCplus", df, scale, v, pr.class_, pr.pi, pr.annot, est_pi=est)
        else:
            sampler.set_prior(sid, method, df, scale, v, pi_in=pi, est_pi=est, region_off=region_off, lhs0=lhs0, rhs0=rhs0, **extra)
        nvar = sampler.sets[sid]["nvar"]
        info.append({"name": term.name, "method": name, "p": p, "nvar": nvar, "df": df, "scale": scale,
                     "n_pi": (len(pr.class_) if name == "BayesR" else len(pr.class_) * nvar if name in ("BayesRCπ", "BayesRCplus") else 2)})
        if name == "BayesLV":
            info[-1]["lv"] = LogVarModel(pr, p, np.random.default_rng([getattr(sampler, "seed", 0), getattr(sampler, "chain_id", 0), sid]))
    fixed = fixed or []                                                          # [(name, data (n,c), level names)]
    if fixed:
        sampler.set_fixed_effects([d for _, d, _ in fixed])                      # X[xSet] besides the intercept (functions.jl:22-54)
    sampler.set_phenotype(Y)
    sampler.set_residual_prior(df_e, scale_e)
    sampler.set_residual_weights(weights)
    sampler.set_intercept(intercept)
    fx_names = [lv for _, _, lvs in fixed for lv in lvs]
    if outPut is not None:                                                       # header rows, mme.jl:543-595
        outMCMC(outPut, "b", [(["(Intercept)"] if intercept else []) + fx_names])
        for term, inf in zip(M, info):
            levels = term.levels or [f"M{i}" for i in range(1, inf["p"] + 1)]
            outMCMC(outPut, f"beta{term.name}", [levels])
            outMCMC(outPut, f"delta{term.name}", [levels])
            if inf["method"] in ("BayesB", "BayesC", "BayesR", "BayesRCπ", "BayesRCplus"):                  # mme.jl:571-576
                outMCMC(outPut, f"pi{term.name}", [[f"pi{v}" for v in range(1, inf["n_pi"] + 1)]])
            if inf["method"] in ("BayesRCπ", "BayesRCplus"):
                outMCMC(outPut, f"annot{term.name}", [levels])
            if inf["method"] == "BayesLV":                                      # mme.jl:577-579
                outMCMC(outPut, f"c{term.name}", [[f"c{v}" for v in range(1, len(inf["lv"].c) + 1)]])
                outMCMC(outPut, f"varZeta{term.name}", [["varZeta"]])
        for term, inf in zip(M, info):
            outMCMC(outPut, f"var{term.name}", [[f"reg_{r}" for r in range(1, inf["nvar"] + 1)]])
        outMCMC(outPut, "varE", [["e"]])
    return {"df_e": df_e, "scale_e": scale_e, "sets": info, "n_fixed": len(fx_names)}


def _getMME_tuple(sampler, Y, M, priorVCV, tuples, outPut, intercept, df_e, scale_e):
    """Correlated (multi-breed) marker sets: one tuple whose members are all marker terms of the model."""
    if len(tuples) != 1 or sorted(tuples[0]) != sorted(t.name for t in M):
        raise NotImplementedError("one tuple covering every marker set of the model is supported on device")
    names = list(tuples[0])
    pr = priorVCV[tuples[0]]
    if pr.name != "BayesPR":
        raise NotImplementedError("correlated marker sets are sampled with BayesPR (functions.jl:140)")
    by_name = {t.name: t for t in M}
    maps = [by_name[nm].map for nm in names]
    if any(m != maps[0] for m in maps):
        raise ValueError("correlated marker sets must have the same map file!")      # mme.jl:453
    k = len(names)
    V = np.asarray(pr.v, dtype=np.float64).reshape(k, k)
    for sid, nm in enumerate(names):
        by_name[nm].upload(sampler, sid)
    p = by_name[names[0]].p
    if maps[0] is None or maps[0] == "":                                          # mme.jl:470-481
        if pr.r == 1:
            region_off = np.arange(p + 1, dtype=np.int64)
        elif pr.r == 9999:
            region_off = None
        else:
            raise ValueError("Please enter a valid region size (1 or 9999)")
    else:
        region_off = prep2RegionData(outPut, "_".join(names), maps[0], pr.r)        # mme.jl:483-484
    df = 3.0 + k                                                                  # mme.jl:493
    sampler.set_joint_prior(list(range(k)), df, V * (df - k - 1.0), V, region_off=region_off)   # mme.jl:501,516
    sampler.set_phenotype(Y)
    sampler.set_residual_prior(df_e, scale_e)
    sampler.set_intercept(intercept)
    R = sampler.joint["n_regions"]
    tname = "_".join(names)
    if outPut is not None:
        outMCMC(outPut, "b", [["(Intercept)"]] if intercept else [[]])
        for nm in names:
            levels = by_name[nm].levels or [f"M{i}" for i in range(1, p + 1)]
            outMCMC(outPut, f"beta{nm}", [levels])
        outMCMC(outPut, f"var{tname}", [[f"reg_{r}_{a}_{b}" for r in range(1, R + 1) for a in names for b in names]])
        outMCMC(outPut, "varE", [["e"]])
    return {"df_e": df_e, "scale_e": scale_e, "sets": [], "tuple": {"names": names, "name": tname, "k": k, "p": p, "df": df}}


def runSampler(sampler: Sampler, M: list[MarkerTerm], info: dict, chainLength: int, burnIn: int, outputFreq: int, outPut: str | None,
               intercept: bool = True, on_sample=None) -> None:
    """samplers.runSampler! (samplers.jl:23-106): iterations run on device in batches that end on a kept iteration;
    kept samples are written in the order of samplers.jl:57-103."""
    these2Keep = list(range(burnIn + outputFreq, chainLength + 1, outputFreq))   # samplers.jl:26
    lv = [(sid, inf["lv"]) for sid, inf in enumerate(info["sets"]) if "lv" in inf]

    def run(k: int) -> None:
        if not lv:
            return sampler.run(k)
        # BayesLV: one iteration per launch; after each one the host replaces the variances of the LV sets (functions.jl:446-485)
        rng = info.setdefault("lv_rng", np.random.default_rng([getattr(sampler, "seed", 0), getattr(sampler, "chain_id", 0), 1 << 20]))
        for _ in range(k):
            sampler.run(1)
            if "lv_var" not in info:
                info["lv_var"] = {sid: np.full(m.logVar.shape, np.exp(m.logVar[0])) for sid, m in lv}
            st = sampler.state(want_e=False)
            for sid, m in lv:
                m.update(st["sets"][sid]["beta"], info["lv_var"][sid], rng)
                sampler.set_var_beta(sid, info["lv_var"][sid])

    done = 0
    if burnIn > 0:
        run(min(burnIn, chainLength))
        done = min(burnIn, chainLength)
    sampler.reset_posterior()                                                    # device-side posterior sums exclude burn-in
    for it in these2Keep:
        run(it - done)
        done = it
        st = sampler.state(want_e=False)
        if outPut is not None:
            bfx = list(sampler.fixed_effects()) if info.get("n_fixed") else []
            st["b_fixed"] = np.array(bfx)
            outMCMC(outPut, "b", [([st["mu"]] if intercept else []) + bfx])
            outMCMC(outPut, "varE", st["varE"])
            for sid, (term, inf) in enumerate(zip(M, info["sets"])):
                outMCMC(outPut, f"beta{term.name}", st["sets"][sid]["beta"])
                outMCMC(outPut, f"delta{term.name}", st["sets"][sid]["delta"])
                if inf["method"] in ("BayesB", "BayesC", "BayesR"):              # samplers.jl:82
                    outMCMC(outPut, f"pi{term.name}", st["sets"][sid]["piHat"])
                if inf["method"] in ("BayesRCπ", "BayesRCplus"):                   # samplers.jl:85-88: vcat(piHat...), annotCat
                    rs = sampler.rc_state(sid)
                    st["sets"][sid]["rc"] = rs
                    outMCMC(outPut, f"pi{term.name}", rs["piHat"].ravel())
                    outMCMC(outPut, f"annot{term.name}", rs["annot_cat"])
                if inf["method"] == "BayesLV":                                   # samplers.jl:89-92
                    outMCMC(outPut, f"c{term.name}", inf["lv"].c)
                    outMCMC(outPut, f"varZeta{term.name}", inf["lv"].varZeta)
            for sid, term in enumerate(M):
                if not info.get("tuple"):
                    outMCMC(outPut, f"var{term.name}", st["sets"][sid]["varBeta"])
            if info.get("tuple"):
                js = sampler.joint_state()
                st["joint"] = js
                for b, nm in enumerate(info["tuple"]["names"]):
                    outMCMC(outPut, f"beta{nm}", js["beta"][b])
                outMCMC(outPut, f"var{info['tuple']['name']}", js["varBeta"].ravel())
        if on_sample is not None:
            on_sample(it, st)
    if done < chainLength:
        run(chainLength - done)


_SNP_RE = re.compile(r"SNP\(\s*([A-Za-z_]\w*)\s*,\s*([^,\)]+?)\s*(?:,\s*([^\)]+?)\s*)?\)")


def runLMEM(formula: str, userData, nChain: int, nBurn: int, nThin: int, outFolder: str = "outMCMC", VCV: dict | None = None,
            summaryStat: dict | None = None, device: int = 0, seed: int = 0, chain_id: int = 0, matrices: dict | None = None,
            sampler: Sampler | None = None) -> Sampler:
    """MCMC.runLMEM (MCMC.jl:31-41) for formulas of the form  "y ~ 1 + SNP(M,geno.txt[,map.txt]) [+ SNP(...)]".
    userData: mapping with the response column.  matrices: optional {name: ndarray of codes} instead of files.
    Any other term (fixed covariates, PED, (1|g), GBLUP) stays in Julia — use ngp_sweep from the Julia shim."""
    VCV = VCV or {}
    lhs, rhs = [s.strip() for s in formula.split("~")]
    terms = [t.strip() for t in re.split(r"\+(?![^(]*\))", rhs)]
    intercept = False
    M: list[MarkerTerm] = []
    fixed: list = []
    for t in terms:
        if t == "1":
            intercept = True
        elif t == "0" or t == "-1":
            intercept = False
        else:
            m = _SNP_RE.fullmatch(t)
            if not m and re.fullmatch(r"[A-Za-z_]\w*", t) and t in userData:
                # a covariate (numeric column) or a factor (anything else: one column per level, first level dropped like
                # StatsModels' DummyCoding); sampled on device right after the intercept (functions.jl:39-54)
                col = np.asarray(userData[t])
                if np.issubdtype(col.dtype, np.number):
                    fixed.append((t, col.astype(np.float64)[:, None], [t]))
                else:
                    lv = sorted(set(col.tolist()))
                    fixed.append((t, np.column_stack([(col == v).astype(np.float64) for v in lv[1:]]), [f"{t}: {v}" for v in lv[1:]]))
                continue
            if not m:
                raise NotImplementedError(f"term '{t}' is outside the B200 hot path (stays in Julia)")
            name, path, mp = m.group(1), m.group(2).strip("\"'"), (m.group(3) or "").strip("\"'")
            if matrices and name in matrices:
                M.append(MarkerTerm(name, prep_snp(matrices[name]), mp or None))
            elif path.lower().endswith(".bed"):
                # PLINK triple: sample / variant counts from the .fam / .bim files next to the .bed
                stem = path[:-4]
                n_s = sum(1 for ln in open(stem + ".fam") if ln.strip())
                p_v = sum(1 for ln in open(stem + ".bim") if ln.strip())
                packed, _ = read_bed_packed(path, n_s, p_v)
                M.append(MarkerTerm(name, None, mp or None, packed=packed, n=n_s))
            else:
                packed, n_s, _ = read_text_packed(path)      # text -> 2-bit codes -> device, never a Float64 matrix (prepMatVec.jl:116-120)
                M.append(MarkerTerm(name, None, mp or None, packed=packed, n=n_s))
    Y = np.asarray(userData[lhs], dtype=np.float64)
    folderHandler(outFolder)
    sampler = sampler or Sampler(device)
    sampler.set_rng(seed, chain_id)
    info = getMME(sampler, Y, M, VCV, summaryStat, outFolder, intercept=intercept, fixed=fixed)
    runSampler(sampler, M, info, nChain, nBurn, nThin, outFolder, intercept=intercept)
    return sampler
