#!/usr/bin/env python
"""Sweep-time of ngp::gibbs_kernel over geometry knobs (block, look-ahead, tile stages, near depth) at a bench workload.
Usage (GPU box): python profiles/tune.py --config c2 --combos 32:12:16:2,32:8:12:2 [--iters 30]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nextgp.jl_b200 as ngp  # noqa: E402
from bench import CONFIGS, SEED0  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--combos", default="0:0:0:0")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--model", default="")
    ap.add_argument("--max_ctas", type=int, default=0)
    ap.add_argument("--refetch", type=int, default=-1)
    ap.add_argument("--min_rows", type=int, default=0)
    ap.add_argument("--storage", default="i8")
    ap.add_argument("--opt", default="0", help="comma list of NGP_CFG_OPT masks, each run for every combo")
    a = ap.parse_args()
    n, p, model = CONFIGS[a.config]
    model = a.model or model
    seed = SEED0 + 2
    prob = ngp.synth.problem(n, p, seed)
    v_e, v, pi = ngp.synth.priors(prob, model)
    method = 0 if model in ("BayesRR", "BayesPR") else (1 if model == "BayesB" else 2)
    for combo, opt in [(c, int(o)) for c in a.combos.split(",") for o in a.opt.split(",")]:
        b, d, nt, dn, dbg, nv = (list(int(x) for x in combo.split(":")) + [0, 0])[:6]
        t0 = time.time()
        try:
            s = ngp.Sampler(0, block=b, lookahead=d, tile_stages=nt, near=dn, max_ctas=a.max_ctas, refetch=a.refetch, min_rows=a.min_rows, storage=a.storage)
            if nv:
                s.configure(ngp._lib.CFG_VERSIONS, nv)
            if dbg:
                s.configure(ngp._lib.CFG_DEBUG, dbg)
            if opt:
                s.configure(ngp._lib.CFG_OPT, opt)
            s.synth_genotypes(0, n, p, seed, prob["thr0"], prob["thr1"])
            t_up = time.time() - t0
            s.set_prior(0, method, 4.0, v * 0.5, v, pi_in=pi, est_pi=(method == 2))
            s.set_phenotype(prob["y"]); s.set_residual_prior(4.0, v_e * 0.5); s.set_intercept(True); s.set_rng(seed, 0)
            ms = []
            for it in range(a.iters):
                s.run(1)
                ms.append(s.timing()["last_run_ms"])
            g = s.timing()
            st = s.state(want_e=False)
            print(json.dumps({"combo": combo, "opt": opt, "ms_last10_mean": float(np.mean(ms[-10:])), "ms_min": float(np.min(ms)), "ms_first": ms[0],
                              "upload_s": round(t_up, 2), "included": int(st["sets"][0]["delta"].sum()),
                              "geom": {k: g[k] for k in ("ctas", "block", "rows_per_cta", "smem_bytes", "lookahead", "near_depth", "tile_stages", "record_stages")}}), flush=True)
            s.close()
        except Exception as ex:  # noqa: BLE001
            print(json.dumps({"combo": combo, "error": str(ex)}), flush=True)


if __name__ == "__main__":
    main()
