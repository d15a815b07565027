"""ctypes binding of libngp.so (include/ngp.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
SO_PATH = os.environ.get("LIBNGP") or os.path.join(_HERE, "libngp.so")      # LIBNGP: another build of the same ABI (A/B timing)
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(_ROOT, "include", "ngp.h")

NGP_MAX_SETS = 8
BAYESPR, BAYESB, BAYESC, BAYESR, BAYESRCPI, BAYESRCPLUS = 0, 1, 2, 3, 5, 6
GENO_I8, GENO_F64, GENO_PACKED2 = 0, 1, 2
STORE_I8, STORE_2BIT = 0, 1
KERNEL_BLOCKED, KERNEL_LITERAL = 0, 1
CFG_KERNEL, CFG_BLOCK, CFG_MIN_ROWS, CFG_MAX_CTAS, CFG_LOOKAHEAD, CFG_TILE_STAGES, CFG_NEAR, CFG_PROFILE, CFG_DEBUG, CFG_VERSIONS, CFG_REFETCH, CFG_OPT = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11
OK, EINVAL, ECUDA, EDATA, ERANGE, ENOMEM, EUNSUPPORTED, ENUMERIC, ETIMEOUT = 0, -1, -2, -3, -4, -5, -6, -7, -8

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
BUILD_DIR = os.path.join(_HERE, "build")          # object files (git-ignored); the linked library is nextgp.jl_b200/libngp.so
STAMP = SO_PATH + ".srchash"                      # hash of the sources + flags the library was linked from


class NgpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libngp error {code}: {msg}")
        self.code = code


class Prior(C.Structure):
    _fields_ = [("method", C.c_int32), ("est_pi", C.c_int32), ("df", C.c_double), ("scale", C.c_double),
                ("var_init", C.c_double), ("pi_in", C.c_double), ("n_regions", C.c_int64),
                ("region_off", C.c_void_p), ("lhs0", C.c_void_p), ("rhs0", C.c_void_p),
                ("n_class", C.c_int32), ("pad_", C.c_int32), ("v_class", C.c_void_p), ("pi_class", C.c_void_p)]


class RCPrior(C.Structure):
    _fields_ = [("plus", C.c_int32), ("est_pi", C.c_int32), ("df", C.c_double), ("scale", C.c_double), ("var_init", C.c_double),
                ("n_class", C.c_int32), ("n_annot", C.c_int32), ("v_class", C.c_void_p), ("pi_class", C.c_void_p), ("annot", C.c_void_p)]


class JointPrior(C.Structure):
    _fields_ = [("k", C.c_int32), ("set_id", C.c_int32 * NGP_MAX_SETS), ("pad_", C.c_int32), ("df", C.c_double),
                ("scale", C.c_void_p), ("var_init", C.c_void_p), ("n_regions", C.c_int64), ("region_off", C.c_void_p)]


class FixedSet(C.Structure):
    _fields_ = [("n_cols", C.c_int32), ("pad_", C.c_int32), ("data", C.c_void_p), ("lhs0", C.c_double), ("rhs0", C.c_double)]


class ShardInfo(C.Structure):
    _fields_ = [("ipc", C.c_ubyte * 64), ("n_local", C.c_int64), ("worker_ctas", C.c_int32), ("device", C.c_int32),
                ("pid", C.c_int64), ("local_ptr", C.c_uint64)]


class Replay(C.Structure):
    _fields_ = [("n_iter", C.c_int32), ("n_sets", C.c_int32), ("chi2_e", C.c_void_p), ("z_mu", C.c_void_p),
                ("u", C.c_void_p * NGP_MAX_SETS), ("z", C.c_void_p * NGP_MAX_SETS),
                ("chi2_b", C.c_void_p * NGP_MAX_SETS), ("beta_pi", C.c_void_p * NGP_MAX_SETS)]


class State(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_sets", C.c_int32), ("pad_", C.c_int32), ("e", C.c_void_p),
                ("mu", C.c_double), ("varE", C.c_double), ("iter", C.c_int64),
                ("beta", C.c_void_p * NGP_MAX_SETS), ("delta", C.c_void_p * NGP_MAX_SETS),
                ("varBeta", C.c_void_p * NGP_MAX_SETS), ("pi", (C.c_double * 2) * NGP_MAX_SETS)]


class Timing(C.Structure):
    _fields_ = [("last_run_ms", C.c_double), ("launches", C.c_int64), ("ctas", C.c_int32), ("threads", C.c_int32),
                ("block", C.c_int32), ("rows_per_cta", C.c_int32), ("smem_bytes", C.c_int64),
                ("lookahead", C.c_int32), ("near_depth", C.c_int32), ("tile_stages", C.c_int32), ("record_stages", C.c_int32),
                ("kernel_variant", C.c_int32), ("refetch", C.c_int32), ("storage_2bit", C.c_int32), ("pad_", C.c_int32),
                ("sum_run_ms", C.c_double)]


def translation_units() -> list[tuple[str, str, list[str]]]:
    """(object name, source file, extra flags): one instantiation of the sweep kernels per unit (csrc/ngp_kernels.h)."""
    tus = [("ngp_api", "ngp_api.cu", []), ("ngp_ingest", "ngp_ingest.cpp", [])]
    for B in (16, 32, 64):
        for v in range(12):
            tus.append((f"ngp_k_gibbs_{B}_{v}", "ngp_k_gibbs.cu", [f"-DNGP_KB={B}", f"-DNGP_KV={v}"]))
    for k in range(2, 9):
        tus.append((f"ngp_k_joint_{k}", "ngp_k_joint.cu", [f"-DNGP_JK={k}"]))
    return tus


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [HEADER]


def _hash(extra: list[str]) -> str:
    import hashlib
    h = hashlib.sha256()
    for f in sources():
        h.update(os.path.basename(f).encode()); h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS + extra).encode())
    return h.hexdigest()


def source_hash() -> str:
    return _hash([])


def _stale() -> bool:
    """The library is rebuilt when the hash of its sources (not their mtime) differs from the one it was linked from."""
    if not os.path.exists(SO_PATH) or not os.path.exists(STAMP):
        return True
    return open(STAMP).read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False, jobs: int = 0) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> nextgp.jl_b200/libngp.so (in-tree, travels to the GPU box).
    Units are compiled in parallel; an object is reused only if the hash of ALL sources + its flags is unchanged."""
    if os.environ.get("LIBNGP"):
        return SO_PATH
    if not (force or _stale()):
        return SO_PATH
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(BUILD_DIR, exist_ok=True)
    tus = translation_units()

    def compile_one(tu):
        name, src, extra = tu
        obj, stamp = os.path.join(BUILD_DIR, name + ".o"), os.path.join(BUILD_DIR, name + ".hash")
        hh = _hash(extra)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == hh:
            return obj, ""
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + extra + ["-c", "-o", obj, os.path.join(CSRC, src)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {name}:\n" + r.stdout + r.stderr)
        open(stamp, "w").write(hh)
        return obj, r.stderr
    with ThreadPoolExecutor(max_workers=jobs or max(1, min(16, os.cpu_count() or 1))) as ex:
        res = list(ex.map(compile_one, tus))
    if verbose:
        for _, err in res:
            sys.stderr.write(err)
    r = subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO_PATH] + [o for o, _ in res], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    open(STAMP, "w").write(source_hash())
    return SO_PATH


_lib = None

_SIGS = {
    "ngp_abi_version": (C.c_int, []),
    "ngp_device_count": (C.c_int, []),
    "ngp_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "ngp_destroy": (C.c_int, [C.c_void_p]),
    "ngp_last_error": (C.c_char_p, [C.c_void_p]),
    "ngp_configure": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "ngp_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ngp_upload_genotypes": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_int]),
    "ngp_synth_genotypes": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]),
    "ngp_synth_genotypes_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]),
    "ngp_shard_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ngp_shard_export": (C.c_int, [C.c_void_p, C.POINTER(ShardInfo)]),
    "ngp_run_group": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int32]),
    "ngp_shard_attach": (C.c_int, [C.c_void_p, C.POINTER(ShardInfo)]),
    "ngp_get_column_sums": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ngp_set_column_sums": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "ngp_gram_size": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]),
    "ngp_get_gram": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ngp_set_gram": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ngp_download_genotypes": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p]),
    "ngp_get_column_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ngp_pack2": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]),
    "ngp_unpack2": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]),
    "ngp_read_text_genotypes": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64)]),
    "ngp_read_bed_genotypes": (C.c_int, [C.c_char_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64)]),
    "ngp_set_phenotype": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "ngp_set_residual_prior": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "ngp_set_residual_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "ngp_set_intercept": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double]),
    "ngp_set_fixed_effects": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(FixedSet)]),
    "ngp_get_fixed_effects": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ngp_set_fixed_replay": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "ngp_set_prior": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Prior)]),
    "ngp_set_marker_summary": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ngp_set_var_beta": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ngp_set_joint_prior": (C.c_int, [C.c_void_p, C.POINTER(JointPrior)]),
    "ngp_set_joint_replay": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ngp_joint_sweep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "ngp_get_joint_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ngp_set_rng": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32]),
    "ngp_set_replay": (C.c_int, [C.c_void_p, C.POINTER(Replay)]),
    "ngp_run": (C.c_int, [C.c_void_p, C.c_int32]),
    "ngp_sweep": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ngp_get_class_pi": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ngp_set_rc_prior": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(RCPrior)]),
    "ngp_get_rc_state": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ngp_set_rc_replay": (C.c_int, [C.c_void_p, C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ngp_get_state": (C.c_int, [C.c_void_p, C.POINTER(State)]),
    "ngp_set_state": (C.c_int, [C.c_void_p, C.POINTER(State)]),
    "ngp_reset_posterior": (C.c_int, [C.c_void_p]),
    "ngp_get_posterior": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_void_p]),
    "ngp_get_timing": (C.c_int, [C.c_void_p, C.POINTER(Timing)]),
    "ngp_get_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "ngp_get_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "ngp_debug_variates": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_double, C.c_int64, C.c_void_p]),
}


def exported_symbols() -> list[str]:
    return sorted(_SIGS)


def lib() -> C.CDLL:
    """Loads libngp.so.  Raises if it has not been built — never falls back to anything else."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a).  nextgp.jl_b200 has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in _SIGS.items():
            if os.environ.get("LIBNGP") and not hasattr(L, name):
                continue                     # an older build of the ABI loaded for A/B timing: its missing entry points stay unbound
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
