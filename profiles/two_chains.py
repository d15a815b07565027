#!/usr/bin/env python
"""Two (or more) independent chains sharing ONE GPU: each handle gets a slice of the SMs (NGP_CFG_MAX_CTAS) and its own copy of the
genotypes; the persistent kernels run concurrently from one host thread each.  Reports the aggregate marker-updates/s against one chain
owning the whole GPU.  Usage (GPU box): python profiles/two_chains.py --config c2 --chains 2 --iters 30"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nextgp.jl_b200 as ngp  # noqa: E402
from bench import CONFIGS, SEED0  # noqa: E402


def make(n, p, model, seed, chain, max_ctas):
    prob = ngp.synth.problem(n, p, seed)
    v_e, v, pi = ngp.synth.priors(prob, model)
    method = 0 if model in ("BayesRR", "BayesPR") else (1 if model == "BayesB" else 2)
    s = ngp.Sampler(0, max_ctas=max_ctas)
    s.synth_genotypes(0, n, p, seed, prob["thr0"], prob["thr1"])
    s.set_prior(0, method, 4.0, v * 0.5, v, pi_in=pi, est_pi=(method == 2))
    s.set_phenotype(prob["y"]); s.set_residual_prior(4.0, v_e * 0.5); s.set_intercept(True); s.set_rng(seed, chain)
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--chains", type=int, default=2)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--sms", type=int, default=148)
    a = ap.parse_args()
    n, p, model = CONFIGS[a.config]
    seed = SEED0 + 2
    per = a.sms // a.chains
    hs = [make(n, p, model, seed, c, per) for c in range(a.chains)]
    for s in hs:
        s.run(5)                                  # warm-up, one after the other
    def work(s):
        s.run(a.iters)                            # ONE launch of a.iters iterations per chain; the launches overlap
    th = [threading.Thread(target=work, args=(s,)) for s in hs]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    ms = [s.timing()["last_run_ms"] for s in hs]
    print(json.dumps({"config": a.config, "chains": a.chains, "ctas_per_chain": hs[0].timing()["ctas"], "geometry": hs[0].timing(),
                      "wall_s": dt, "kernel_ms_per_chain": ms, "aggregate_marker_updates_per_s": a.chains * a.iters * p / dt,
                      "per_sweep_ms_wall": 1e3 * dt / a.iters}))
    for s in hs:
        s.close()


if __name__ == "__main__":
    main()
