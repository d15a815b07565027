#!/usr/bin/env python
"""Turns a directory written by julia/record_variates.jl (raw little-endian Float64 / Int64 files + header.txt) into the fixture
tests/golden/julia/<case>.npz that tests/test_julia_golden.py replays through the oracle (CPU) and the GPU sampler.
    python tests/golden/import_julia_log.py <recorder_out_dir> tests/golden/julia/<case>
The genotypes (0/1/2 text file named in header.txt) and phenotypes travel inside the fixture, so keep the case small."""
import sys

import numpy as np


def load(d):
    hdr = {}
    for ln in open(f"{d}/header.txt"):
        k, _, v = ln.strip().partition(" ")
        hdr[k] = v
    assert hdr["format"] == "ngp-replay-log 1"
    n, p, it, nvar = (int(hdr[k]) for k in ("n", "p", "iters", "nvar"))
    f = lambda name, shape: np.fromfile(f"{d}/{name}.f64", dtype="<f8").reshape(shape)
    out = {"chi2_e": f("chi2_e", (it,)), "z_mu": f("z_mu", (it,)), "u": f("u", (it, p)), "z": f("z", (it, p)),
           "chi2_b": f("chi2_b", (it, nvar)), "beta_pi": f("beta_pi", (it,)), "varE": f("varE", (it,)), "mu": f("mu", (it,)),
           "beta": f("beta", (it, p)), "varBeta": f("varBeta", (it, nvar)), "pi": f("pi", (it, 2)),
           "delta": np.fromfile(f"{d}/delta.i64", dtype="<i8").reshape(it, p),
           "ycorr_final": np.fromfile(f"{d}/ycorr_final.f64", dtype="<f8")}
    codes = np.loadtxt(hdr["genotypes"], dtype=np.int8, ndmin=2)
    import csv
    rows = list(csv.DictReader(open(hdr["phenotypes"])))
    out["codes"] = np.asfortranarray(codes)
    out["y"] = np.array([float(r["y"]) for r in rows])
    for k in ("method",):
        out[k] = np.array(hdr[k])
    for k in ("est_pi", "n_regions"):
        out[k] = np.array(int(hdr[k]))
    for k in ("df", "scale", "v", "pi", "df_e", "scale_e"):
        out[k] = np.array(float(hdr[k]))
    return out


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    np.savez_compressed(dst, **load(src))
    print("wrote", dst + ".npz")
