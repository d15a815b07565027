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

NAMES = ["x0", "d_imma", "u_axpy_quant", "c_rbase_wait", "c_load", "u_dotdone_wait", "changed_effects", "spec_evals",
         "phase0", "phase1", "u_list_wait", "phase3", "p_corrections", "p_acc_poll", "d_version_wait", "d_reduce_red",
         "c_near1", "c_spec_loop", "c_publish", "c_outputs", "d_tile_wait", "p_looplat_sum", "p_looplat_n", "p_looplat_max",
         "l_poll_ns", "l_poll_n", "l_upd_ns", "l_upd_n", "l_dot_ns", "l_dot_n", "l_acc_ns", "l_acc_n"]
# u_* updater warp 0, d_* first dot warp (handles every 8th block) of a worker CTA; c_* chain warp, p_* first prep warp (every 8th block)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--kernel", default="blocked")
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--warm", type=int, default=20)
    ap.add_argument("--model", default="")
    ap.add_argument("--lookahead", type=int, default=0)
    ap.add_argument("--tiles", type=int, default=0)
    ap.add_argument("--near", type=int, default=0)
    a = ap.parse_args()
    n, p, model = CONFIGS[a.config]
    model = a.model or model
    seed = SEED0 + 2
    prob = ngp.synth.problem(n, p, seed)
    v_e, v, pi = ngp.synth.priors(prob, model)
    method = 0 if model in ("BayesRR", "BayesPR") else (1 if model == "BayesB" else 2)
    s = ngp.Sampler(0, kernel=a.kernel, block=a.block, lookahead=a.lookahead, tile_stages=a.tiles, near=a.near)
    s.configure(ngp._lib.CFG_PROFILE, 1)
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
            nblk = (p + t["block"] - 1) // t["block"]
            w, c = pr[:-1], pr[-1]           # worker CTAs, chain CTA
            rec["worker_per_block_mean_cycles"] = {k: float(w[:, i].mean()) / nblk for i, k in enumerate(NAMES) if k[:2] in ("u_", "d_") or k.startswith("phase")}
            rec["worker_per_block_max_cycles"] = {k: float(w[:, i].max()) / nblk for i, k in enumerate(NAMES) if k[:2] in ("u_", "d_")}
            rec["loop_latency_cycles"] = {"mean": float(c[21] / max(c[22], 1)), "max": float(c[23]), "samples": int(c[22])}
            rec["ns_since_list_published"] = {"list_received_by_poll_warp": float(w[:, 24].sum() / max(w[:, 25].sum(), 1)),
                                              "residual_version_ready": float(w[:, 26].sum() / max(w[:, 27].sum(), 1)),
                                              "dots_of_block_plus_D_plus_1_red": float(w[:, 28].sum() / max(w[:, 29].sum(), 1)),
                                              "their_sums_complete_in_chain_cta": float(c[30] / max(c[31], 1))}
            rec["chain_cta_per_block_cycles"] = {k: float(c[i]) / nblk for i, k in enumerate(NAMES) if k[0] in "cp" or k in ("changed_effects", "spec_evals")}
        out["iters"].append(rec)
    tr = s.trace()
    nst = min(2048, (p + 63) // 64)
    out["chain_trace"] = {"step_start_delta": np.diff(tr[:nst, 0]).tolist(), "wait": tr[:nst, 1].tolist()}
    out["geometry"] = s.timing()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
