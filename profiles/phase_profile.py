#!/usr/bin/env python
"""Per-phase cycle breakdown of ngp::gibbs_kernel (thread 0 of every CTA, clock64) at a bench workload.
Usage (on the GPU box): python profiles/phase_profile.py --config c2 [--kernel blocked] [--warm 20] > gpurun_out/phases.json"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nextgp.jl_b200 as ngp  # noqa: E402
from bench import CONFIGS, SEED0  # noqa: E402

NAMES = ["tile_wait", "w_dot_red", "w_axpy", "acc_poll", "chain", "chain_waits_workers", "changed_effects", "spec_evals",
         "phase0", "phase1", "w_waits_chain", "phase3", "w_tile_wait", "w_dot_loop", "w_dot_bar", "x15"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--kernel", default="blocked")
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--warm", type=int, default=20)
    ap.add_argument("--model", default="")
    a = ap.parse_args()
    n, p, model = CONFIGS[a.config]
    model = a.model or model
    seed = SEED0 + 2
    prob = ngp.synth.problem(n, p, seed)
    v_e, v, pi = ngp.synth.priors(prob, model)
    method = 0 if model in ("BayesRR", "BayesPR") else (1 if model == "BayesB" else 2)
    s = ngp.Sampler(0, kernel=a.kernel, block=a.block)
    s.synth_genotypes(0, n, p, seed, prob["thr0"], prob["thr1"])
    s.set_prior(0, method, 4.0, v * 0.5, v, pi_in=pi, est_pi=(method == 2))
    s.set_phenotype(prob["y"]); s.set_residual_prior(4.0, v_e * 0.5); s.set_intercept(True); s.set_rng(seed, 0)
    out = {"config": a.config, "n": n, "p": p, "model": model, "kernel": a.kernel, "iters": []}
    for it in range(1, a.warm + 4):
        s.run(1)
        t = s.timing()
        rec = {"iter": it, "ms": t["last_run_ms"]}
        if it in (1, 2, 3, 5, 10) or it > a.warm:
            pr = s.profile().astype(np.float64)
            st = s.state(want_e=False)
            rec["included"] = int(st["sets"][0]["delta"].sum()) if method else p
            rec["varE"] = st["varE"]
            rec["phases_mean_cycles"] = {k: float(pr[:, i].mean()) for i, k in enumerate(NAMES)}
            rec["phases_max_cycles"] = {k: float(pr[:, i].max()) for i, k in enumerate(NAMES)}
            nblk = (p + t["block"] - 1) // t["block"]
            rec["per_block_mean_cycles"] = {k: float(pr[:, i].mean()) / nblk for i, k in enumerate(NAMES)}
        out["iters"].append(rec)
    out["geometry"] = s.timing()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
