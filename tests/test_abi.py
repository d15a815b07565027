"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/ngp.h declares,
fails loudly without a device, and the host codec round-trips bit-exactly."""
import os
import re

import numpy as np
import pytest

import nextgp.jl_b200 as ngp
from nextgp.jl_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ngp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ngp_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.lib()
    decl = declared_symbols()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/ngp.h but not exported by libngp.so"
    assert sorted(L.exported_symbols()) == decl, "ctypes signature table and header disagree"
    assert lib.ngp_abi_version() == 1


def test_no_cpu_fallback():
    if L.lib().ngp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(ngp.NgpError) as ei:
        ngp.Sampler(0)
    assert ei.value.code == L.ECUDA and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    """The oracle is the checker: nothing under the product package may import, link or call it."""
    pkg = os.path.join(ROOT, "nextgp.jl_b200")
    pat_py = re.compile(r"^\s*(from|import)\s+oracle|libngp_oracle|oracle\.", re.M)
    pat_c = re.compile(r"#include[^\n]*oracle|ngo_[a-z0-9_]+\s*\(|libngp_oracle")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            src = open(os.path.join(dirpath, f), errors="ignore").read() if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) else ""
            pat = pat_py if f.endswith(".py") else pat_c
            assert not pat.search(src), f"{f} uses the oracle"


@pytest.mark.parametrize("n,p", [(1, 1), (4, 3), (5, 2), (1023, 7), (64, 64)])
def test_pack2_roundtrip_bit_exact(n, p):
    lib = L.lib()
    rng = np.random.default_rng(n * 31 + p)
    codes = np.asfortranarray(rng.integers(0, 3, size=(n, p)).astype(np.int8))
    ldp = (n + 3) // 4 + 2
    packed = np.full((ldp, p), 0xAA, dtype=np.uint8, order="F")
    assert lib.ngp_pack2(codes.ctypes.data, n, p, n, packed.ctypes.data, ldp) == 0
    ref = np.zeros(((n + 3) // 4, p), dtype=np.uint8)
    for i in range(n):
        ref[i // 4] |= (codes[i].astype(np.uint8) << (2 * (i % 4)))
    assert np.array_equal(packed[: (n + 3) // 4], ref) and (packed[(n + 3) // 4:] == 0).all()
    out = np.full((n, p), -7, dtype=np.int8, order="F")
    assert lib.ngp_unpack2(packed.ctypes.data, n, p, ldp, out.ctypes.data, n) == 0
    assert np.array_equal(out, codes)


def test_pack2_rejects_missing():
    lib = L.lib()
    codes = np.asfortranarray(np.array([[0], [3], [1]], dtype=np.int8))
    packed = np.zeros((1, 1), dtype=np.uint8)
    assert lib.ngp_pack2(codes.ctypes.data, 3, 1, 3, packed.ctypes.data, 1) == L.EDATA
    bad = np.array([[0b11]], dtype=np.uint8)
    out = np.zeros((1, 1), dtype=np.int8)
    assert lib.ngp_unpack2(bad.ctypes.data, 1, 1, 1, out.ctypes.data, 1) == L.EDATA
