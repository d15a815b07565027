#!/usr/bin/env python
"""bench.py — marker-updates/s of the Gibbs sweep on B200 (BASELINE.json metric), with roofline and CPU baseline.

  python bench.py --gpus N --steps K --warmup W            # our arm (one chain per GPU, weak scaling)
  python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of NextGP.jl's sampler on the host cores

A "step" is ONE Gibbs iteration (varE -> intercept -> full sweep over all p markers -> variance/pi updates) of the
configured workload; default workload = BASELINE.json configs[1]: 50,000 x 50,000 single-trait BayesCpi, int8 genotypes.
`value` times K steps with everything resident in HBM (CUDA events on the launching stream); `e2e` times the same
sweep through the reference-facing plugin call ngp_sweep(ycorr, varE, beta, delta, varBeta) with HOST buffers.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (n, p, model)    — BASELINE.json configs
    "c1": (1000, 5000, "BayesRR"),
    "c2": (50000, 50000, "BayesC"),
    "c3": (100000, 600000, "BayesB"),
    "c4": (30000, 50000, "MultiBreed2"),
    "c5": (200000, 50000, "BayesC"),
    "tiny": (2000, 4096, "BayesC"),
    "c4rr": (30000, 100000, "BayesRR"),     # diagnostic: dense updates at the size of the interleaved C4 tuple
    "c2r": (50000, 50000, "BayesR"),        # BayesR (4 variance classes, Dirichlet-updated proportions) at the headline shape
}
SEED0 = 20261018
METRIC = "marker-updates/sec"



def host_threads() -> int:
    """All the host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently make the
    CPU arm single-threaded at N > 1 (omp_get_max_threads() == 1): count the cores from the affinity mask instead."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def workload_name(cfg, n, p, model, storage="i8", regions=0):
    st = "2-bit packed" if storage == "2bit" else "int8"
    if regions and model in ("BayesPR", "BayesRR"):
        model = f"BayesPR({regions}) region-wise variances"
    if model.startswith("MultiBreed"):
        return f"{cfg}: {n} individuals x {p} SNPs, {model[10:]}-breed tuple BayesPR (joint effects per locus) + intercept, {st} genotypes in HBM"
    return f"{cfg}: {n} individuals x {p} SNPs single-trait {model}{'pi' if model == 'BayesC' else ''} + intercept, {st} genotypes in HBM"


def clocks_sampler(stop, out):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    try:
        pr = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                               "-i", os.environ.get("LOCAL_RANK", "0")], stdout=subprocess.PIPE, text=True)
    except Exception:
        return
    def reader():
        for line in pr.stdout:
            out.append(line.strip())
    th = threading.Thread(target=reader, daemon=True)
    th.start()
    stop.wait()
    pr.terminate()


def summarize_clocks(lines):
    sm, mx, reasons = [], 0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        parts = [x.strip() for x in ln.split(",")]
        if len(parts) < 6:
            continue
        try:
            sm.append(float(parts[0])); mx = max(mx, float(parts[1]))
        except ValueError:
            continue
        for nm, v in zip(names, parts[2:6]):
            if v.lower().startswith("active"):
                reasons.add(nm)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_baseline(n, p, model, seed, budget_s=15.0, max_cols=2000):
    """The oracle port (CPU restatement with the reference's dense-fp64 memory behaviour) on a bounded sample:
    all n rows, the first `max_cols` markers, whole Gibbs iterations; per-marker cost is independent of p."""
    from oracle import oracle as O
    import nextgp.jl_b200 as ngp
    pc = min(p, max_cols)
    prob = ngp.synth.problem(n, p, seed)
    codes = O.synth_codes(seed, n, 0, pc, prob["thr0"], prob["thr1"])
    X, mean, mpm = O.center_codes(codes)
    del codes
    v_e, v, pi = ngp.synth.priors(prob, model)
    method = 0 if model in ("BayesRR", "BayesPR") else (1 if model == "BayesB" else 2)
    results = {}
    for threads in sorted({1, host_threads()}):
        O.set_threads(threads)
        S = O.MarkerSet(X=X, mpm=mpm, method=method, v=v, pi=pi, est_pi=(method == 2))
        ch = O.OracleChain(prob["y"], [S], v_e=v_e)
        ch.iteration(seed=seed, chain=0)          # warm-up
        t0 = time.perf_counter(); it = 0
        while True:
            ch.iteration(seed=seed, chain=0); it += 1
            dt = time.perf_counter() - t0
            if dt > budget_s / 2 or it >= 50:
                break
        results[threads] = it * pc / dt
    best_t = max(results, key=results.get)
    O.set_threads(1)
    return {"value": results[best_t], "unit": METRIC, "cores": best_t, "kind": "port",
            "sample": f"first {pc} of {p} markers x all {n} individuals, dense fp64 column-major, whole Gibbs iterations "
                      f"(varE, intercept, add-back axpy + dot + axpy per marker); per-thread-count marker-updates/s: "
                      + ", ".join(f"{t}t={v:.3g}" for t, v in sorted(results.items()))
                      + "; CPU restatement of NextGP.jl v1.2.0 (Julia absent on the box; JULIA_NUM_THREADS n/a)",
            "host_cpus": os.cpu_count()}


def run_reference(args, rank):
    n, p, model = CONFIGS[args.config] if not args.n else (args.n, args.p, args.model)
    if rank != 0:
        return
    from oracle import oracle as O
    import nextgp.jl_b200 as ngp
    seed = SEED0 + 2
    pc = min(p, args.ref_cols)
    prob = ngp.synth.problem(n, p, seed)
    codes = O.synth_codes(seed, n, 0, pc, prob["thr0"], prob["thr1"])
    X, mean, mpm = O.center_codes(codes)
    v_e, v, pi = ngp.synth.priors(prob, model)
    method = 0 if model in ("BayesRR", "BayesPR") else (1 if model == "BayesB" else 2)
    S = O.MarkerSet(X=X, mpm=mpm, method=method, v=v, pi=pi, est_pi=(method == 2))
    ch = O.OracleChain(prob["y"], [S], v_e=v_e)
    # the thread count that serves the CPU arm best: all host cores for long columns, one for short ones (the OpenMP fork/join of a
    # level-1 call costs more than a 1000-row dot) — calibrated on two untimed iterations each
    cal = {}
    for threads in sorted({1, host_threads()}):
        O.set_threads(threads)
        ch.iteration(seed=seed, chain=0)
        tc = time.perf_counter()
        ch.iteration(seed=seed, chain=0)
        cal[threads] = time.perf_counter() - tc
    threads = min(cal, key=cal.get)
    O.set_threads(threads)
    for _ in range(args.warmup):
        ch.iteration(seed=seed, chain=0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ch.iteration(seed=seed, chain=0)
    dt = time.perf_counter() - t0
    val = args.steps * pc / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "marker-updates/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps * (p / pc), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, n, p, model), "sample": f"first {pc} of {p} markers, all rows; ms_per_step extrapolated linearly in p",
                       "chains": 1, "note": "ONE CPU chain on rank 0's host cores whatever --gpus is: at N > 1 compare it with value / N of the GPU arm (N chains)"},
            "cpu_baseline": {"value": val, "unit": "marker-updates/s", "cores": threads, "kind": "port",
                             "sample": f"first {pc} of {p} markers x {n} rows, dense fp64, OpenMP over rows in dot/axpy, {threads} of {host_threads()} host threads (the faster of 1 / all, calibrated); CPU restatement of NextGP.jl (no Julia on the box)"},
            "e2e": {"value": val, "unit": "marker-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def allreduce_gram(s, set_id, dist):
    """row-sharded chain on the blocked kernel: the banded Gram is a sum over individuals (NCCL all-reduce of the int32 band at set-up)"""
    import torch
    g = torch.from_numpy(s.gram(set_id)).cuda()
    sizes = [None] * dist.get_world_size()
    dist.all_gather_object(sizes, int(g.numel()))
    assert len(set(sizes)) == 1, f"ranks chose different block sizes / look-aheads: {sizes}"
    dist.all_reduce(g)
    s.set_gram(set_id, g.cpu().numpy())


def sharded_leg(rank, world, local, dist, n, model, seed, p_s=2048, iters=3, kernel="literal"):
    """Row-sharded single chain over the N GPUs of this run on a bounded slice of the workload (all n rows, the first p_s markers),
    CHECKED: every rank must hold the bitwise identical chain, and it must agree with the unsharded chain that rank 0 runs on the
    same data with the same variate stream.  Returns {per_marker_us, max_rel_err, ranks_identical} on rank 0."""
    import torch
    import nextgp.jl_b200 as ngp
    prob = ngp.synth.problem(n, p_s, seed)
    v_e, v, pi = ngp.synth.priors(prob, model)
    method = 0 if model in ("BayesRR", "BayesPR") else (1 if model == "BayesB" else 2)
    per = -(-(-(-n // world)) // 4) * 4
    a, b = min(n, rank * per), min(n, (rank + 1) * per)
    s = ngp.Sampler(local, kernel=kernel)
    s.shard_init(rank, world)
    s.synth_genotypes_rows(0, a, b - a, p_s, seed, prob["thr0"], prob["thr1"])
    infos = [None] * world
    dist.all_gather_object(infos, s.shard_export())
    s.shard_attach(infos)
    cs, css = s.column_sums(0)
    tcs = torch.from_numpy(np.stack([cs, css])).cuda()
    dist.all_reduce(tcs)
    tcs = tcs.cpu().numpy()
    s.set_column_sums(0, n, tcs[0], tcs[1])
    if kernel == "blocked":
        allreduce_gram(s, 0, dist)
    s.set_prior(0, method, 4.0, v * 0.5, v, pi_in=pi, est_pi=(method == 2))
    s.set_phenotype(prob["y"][a:b]); s.set_residual_prior(4.0, v_e * 0.5); s.set_intercept(True); s.set_rng(seed, 0)
    dist.barrier()
    s.run(1)                                            # warm-up launch (also exercises the cross-launch barrier counter)
    torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); s.run(iters - 1); ev1.record()
    torch.cuda.synchronize()
    tms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    st = s.state(want_e=False)
    mine = torch.from_numpy(np.concatenate([st["sets"][0]["beta"], [st["varE"], st["mu"]]])).cuda()
    allb = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allb, mine)
    identical = all(bool(torch.equal(allb[0].view(torch.int64), x.view(torch.int64))) for x in allb)
    s.close()
    out = None
    if rank == 0:
        u = ngp.Sampler(local)                          # the same chain, unsharded, blocked kernel
        u.synth_genotypes(0, n, p_s, seed, prob["thr0"], prob["thr1"])
        u.set_prior(0, method, 4.0, v * 0.5, v, pi_in=pi, est_pi=(method == 2))
        u.set_phenotype(prob["y"]); u.set_residual_prior(4.0, v_e * 0.5); u.set_intercept(True); u.set_rng(seed, 0)
        u.run(iters)
        su = u.state(want_e=False)
        u.close()
        ref = np.concatenate([su["sets"][0]["beta"], [su["varE"], su["mu"]]])
        got = allb[0].cpu().numpy()
        err = float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300))
        out = {"per_marker_us": 1e3 * float(tms.item()) / (iters - 1) / p_s, "max_rel_err": err, "ranks_identical": bool(identical),
               "checked_against": f"unsharded blocked chain of rank 0, {iters} iterations, beta / varE / mu",
               "slice": (f"all {n} rows x first {p_s} markers, per-marker kernel, one fixed-point RED per marker into every rank's accumulator over NVLink"
                         if kernel == "literal" else
                         f"all {n} rows x first {p_s} markers, blocked look-ahead kernel, the B partial sums of a block pushed into every rank's accumulator ring over NVLink")}
    dist.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--kernel", default="blocked", choices=["blocked", "literal"])
    ap.add_argument("--block", type=int, default=0)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--p", type=int, default=0)
    ap.add_argument("--model", default="BayesC")
    ap.add_argument("--storage", default="i8", choices=["i8", "2bit"], help="device storage of the genotype codes (2bit: NGP_STORE_2BIT, a quarter of the HBM bytes per sweep)")
    ap.add_argument("--regions", type=int, default=0, help="BayesPR with one variance per window of this many SNPs (BayesPR(r, v): mme.jl:345-348, misc.jl:163-215) on 30 equal chromosomes")
    ap.add_argument("--cfg-opt", type=int, default=-1, help="NGP_CFG_OPT mask (-1 = the library's defaults)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--weighted", action="store_true",
                    help="diagnostic: residual weights w ~ U(0.5, 2) (E.str == \"D\", mme.jl:70-73); the set is swept by the per-marker kernel")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sharded-leg", action="store_true", help="N > 1: skip the checked row-sharded leg")
    ap.add_argument("--long-seconds", type=float, default=1.0, help="extra untimed-by-contract run of at least this many seconds after the K timed steps (0 = off)")
    ap.add_argument("--ref-cols", type=int, default=2000)
    ap.add_argument("--shard-kernel", default="blocked", choices=["blocked", "literal"], help="--sharded: kernel of the row-sharded chain")
    ap.add_argument("--sharded", action="store_true",
                    help="ONE chain whose individuals are row-sharded over the N GPUs (per-marker reduction over NVLink peer memory); strong scaling")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import nextgp.jl_b200 as ngp
    from nextgp.jl_b200 import _lib as L

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, p, model = CONFIGS[args.config] if not args.n else (args.n, args.p, args.model)
    seed = SEED0 + 2
    # (BayesR: a polygenic trait, p / 10 causal loci — the reference's class likelihoods overflow for a locus of chi-square > ~1400, functions.jl:255)
    prob = ngp.synth.problem(n, p, seed, q=(p // 10 if model == "BayesR" else None))
    v_e, v, pi = ngp.synth.priors(prob, "BayesRR" if model.startswith("MultiBreed") else model)
    method = 0 if model in ("BayesRR", "BayesPR") else (1 if model == "BayesB" else 3 if model == "BayesR" else 2)
    r_vclass, r_pi = np.array([0.0, 1e-4, 1e-3, 1e-2]), np.array([0.95, 0.02, 0.02, 0.01])
    if method == 3:
        v = prob["var_y"] / 2.0                   # BayesR: the class variances are fractions of the genetic variance
        args.no_cpu = True                      # (the timed CPU port covers BayesPR / BayesB / BayesC)

    sharded = args.sharded and world > 1
    kbreeds = int(model[10:]) if model.startswith("MultiBreed") else 0
    if sharded:
        args.kernel = args.shard_kernel
        args.no_e2e = True
    if not args.block and method == 0 and not kbreeds and not sharded and args.storage == "i8" and n > 512 * 147:
        args.block = 16       # dense updates on panels of more than 512 rows: the choice api.getMME makes for all-BayesPR models (ngp_api.cu: apply_ring_geometry)
    s = ngp.Sampler(local, kernel=args.kernel, block=args.block, storage=args.storage)
    if args.cfg_opt >= 0:
        s.configure(L.CFG_OPT, args.cfg_opt)
    stream = torch.cuda.current_stream()
    s.set_stream(stream.cuda_stream)
    df, scale = 4.0, v * 0.5
    y = prob["y"]
    if sharded:
        # one chain, rows [a, b) on this rank; NCCL only for set-up (handles, column sums) and the timing barrier
        per = -(-(-(-n // world)) // 4) * 4
        a, b = min(n, rank * per), min(n, (rank + 1) * per)
        s.shard_init(rank, world)
        s.synth_genotypes_rows(0, a, b - a, p, seed, prob["thr0"], prob["thr1"])
        infos = [None] * world
        dist.all_gather_object(infos, s.shard_export())
        s.shard_attach(infos)
        cs, css = s.column_sums(0)
        tcs = torch.from_numpy(np.stack([cs, css])).cuda()
        dist.all_reduce(tcs)
        tcs = tcs.cpu().numpy()
        s.set_column_sums(0, n, tcs[0], tcs[1])
        if args.kernel == "blocked":
            allreduce_gram(s, 0, dist)
        y = prob["y"][a:b]
    elif kbreeds:
        for bset in range(kbreeds):
            pb = ngp.synth.problem(n, p, seed + 101 * bset)
            s.synth_genotypes(bset, n, p, seed + 101 * bset, pb["thr0"], pb["thr1"])
        V = np.eye(kbreeds) * v + (np.ones((kbreeds, kbreeds)) - np.eye(kbreeds)) * 0.3 * v
        dfk = 3.0 + kbreeds
        s.set_joint_prior(list(range(kbreeds)), dfk, V * (dfk - kbreeds - 1.0), V)
        args.no_e2e = True
    else:
        s.synth_genotypes(0, n, p, seed, prob["thr0"], prob["thr1"])       # every rank: same X, generated on device
    region_off = None
    if args.regions and method == 0:
        chrom = -(-p // 30)                         # 30 equal chromosomes; windows never span two chromosomes
        edges = sorted({min(p, c0 + w) for c0 in range(0, p, chrom) for w in range(0, chrom + args.regions, args.regions)} | {0, p})
        region_off = np.array(edges, dtype=np.int64)
    if not kbreeds:
        if method == 3:
            s.set_prior(0, method, df, v * 0.5, v, est_pi=True, v_class=r_vclass, pi_class=r_pi)
        else:
            s.set_prior(0, method, df, scale, v, pi_in=pi, est_pi=(method == 2), region_off=region_off)
    s.set_phenotype(y)
    s.set_residual_prior(4.0, v_e * 0.5)
    if args.weighted:
        s.set_residual_weights(np.random.default_rng(seed).uniform(0.5, 2.0, n))
        args.kernel, args.no_e2e, args.no_cpu = "literal", True, True
    s.set_intercept(True)
    s.set_rng(seed, 0 if sharded else rank)                            # independent chains: chain id = rank

    for _ in range(args.warmup):
        s.run(1)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    stop, lines = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, lines), daemon=True)
    if rank == 0:
        th.start(); time.sleep(0.3)
    l0 = s.timing()["launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    torch.cuda.synchronize()
    if dist:
        dist.barrier()          # (rank 0 has just slept while its clock sampler started: a sharded chain's other ranks would wait for it inside their kernels)
    torch.cuda.synchronize()
    ev0.record(stream)
    if sharded:
        # one launch runs all K iterations (ngp_run(h, K)): the shards' persistent kernels stay co-resident, no host in the loop
        s.run(args.steps)
        kern_ms.append(s.timing()["last_run_ms"] / args.steps)
    else:
        k0 = s.timing()["sum_run_ms"]
        for _ in range(args.steps):
            s.run(1)                                   # (one launch per step; the library sums the launches' event times)
    ev1.record(stream)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    if not sharded:
        kern_ms.append((s.timing()["sum_run_ms"] - k0) / args.steps)
    launches = s.timing()["launches"] - l0
    if rank == 0:
        time.sleep(0.2); stop.set()
    tms = torch.tensor([ms], device="cuda")
    if dist:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_all = float(tms.item())
    nchains = 1 if sharded else world
    value = nchains * p * args.steps / (ms_all * 1e-3)

    # ---- a longer run (the contract's K steps last ~20 ms at C2): at least --long-seconds of back-to-back launches
    long_run = None
    if args.long_seconds > 0 and not sharded:
        per = max(ms / max(args.steps, 1), 1e-3)
        k2 = int(min(20000, max(args.steps, np.ceil(args.long_seconds * 1e3 / per))))
        lms = []
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k2):
            s.run(1)
            lms.append(s.timing()["last_run_ms"])
        e1.record(stream)
        torch.cuda.synchronize()
        long_run = {"steps": k2, "ms_per_step": e0.elapsed_time(e1) / k2, "kernel_ms_mean": float(np.mean(lms)), "kernel_ms_p10_p90": [float(np.percentile(lms, 10)), float(np.percentile(lms, 90))]}
    # ---- how many effects change per sweep (the cost of a sweep grows with it): three more iterations, counted on the host
    changed = None
    if not sharded and not kbreeds:
        b_prev = s.state(want_e=False)["sets"][0]["beta"].copy()
        cnt = []
        for _ in range(3):
            s.run(1)
            b_new = s.state(want_e=False)["sets"][0]["beta"]
            cnt.append(int(np.count_nonzero(b_new != b_prev)))
            b_prev = b_new.copy()
        changed = float(np.mean(cnt))
    # ---- N > 1: the row-sharded single chain over the same GPUs, checked
    shard_rec = None
    if dist and not sharded and not kbreeds and not args.no_sharded_leg and not args.weighted and args.storage == "i8":
        try:
            shard_rec = sharded_leg(rank, world, local, dist, n, model, seed)                               # per-marker kernel
            blk = sharded_leg(rank, world, local, dist, n, model, seed, p_s=8192, iters=6, kernel="blocked")        # blocked kernel
            if shard_rec is not None:
                shard_rec = dict(shard_rec, blocked=blk)
        except Exception as ex:  # noqa: BLE001
            shard_rec = {"error": str(ex)}

    # ---- e2e: the plugin call with host buffers (pinned), H2D + sweep + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        st = s.state()
        nvar = len(st["sets"][0]["varBeta"])
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        ycorr, beta, delta = pin(st["e"]), pin(st["sets"][0]["beta"]), pin(st["sets"][0]["delta"])
        varBeta, piHat = pin(st["sets"][0]["varBeta"]), pin(st["sets"][0]["piHat"])
        varE = st["varE"]
        s.sweep(0, ycorr, varE, beta, delta, varBeta, piHat)          # warm
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        ke = max(3, args.steps // 2)
        t0 = time.perf_counter()
        for _ in range(ke):
            s.sweep(0, ycorr, varE, beta, delta, varBeta, piHat)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tdt = torch.tensor([dt], device="cuda")
        if dist:
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
        bytes_io = 8 * n + 8 * p + 4 * p + 8 * nvar + 32
        e2e = {"value": world * p * ke / float(tdt.item()), "unit": "marker-updates/s", "h2d_bytes_per_step": bytes_io,
               "d2h_bytes_per_step": bytes_io, "steps": ke,
               "call": "ngp_sweep(set, ycorr, varE, beta, delta, varBeta, piHat) == M[mSet].funct(...) with pinned host buffers"}

    if rank == 0:
        peak, peak_src = measured_peak()
        tm = s.timing()
        gbytes = 0.25 if args.storage == "2bit" else 1.0
        alg_bytes = p * (n * gbytes + 40.0) * max(1, kbreeds)     # n*g + 40 B of per-marker scalars, g = 1 B int8 / 0.25 B 2-bit (per breed column)
        if sharded:
            alg_bytes /= world                                 # bytes streamed by ONE GPU (its row slice)
        k_ms = float(np.mean(kern_ms))
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        traffic, traffic_source = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                key = f"{args.config}:{args.kernel}" + (":2bit" if args.storage == "2bit" else "")
                traffic = tj.get(key)
                if traffic is not None:
                    traffic_source = tj.get("_source", {}).get(key, "ncu --set full capture of this kernel at this config (dram__bytes_read.sum + dram__bytes_write.sum per launch), "
                                                                     "recorded in profiles/traffic.json — NOT measured in this run")
            except Exception:
                traffic = None
        line = {"metric": METRIC, "value": value, "unit": "marker-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(args.config, n, p, model, args.storage, args.regions) + (" [diagnostic: weighted residuals E.str == \"D\", per-marker kernel]" if args.weighted else ""), "kernel": ("blocked tuple sweep (interleaved copy, joint draw in the chain warp)" if args.kernel == "blocked" and kbreeds in (2, 4) else "joint (per locus)") if kbreeds else args.kernel, "chains": nchains,
                           "parallelism": (f"ONE chain row-sharded over {world} GPUs ({args.kernel} kernel): fixed-point partial sums pushed into every rank's "
                                           f"accumulators over NVLink peer memory (CUDA IPC), identical draws on every rank")
                                          if sharded else f"{world} independent chain(s), one per GPU, no data-path collective",
                           "l2": f"genotype matrix {n * p / 1e9:.2f} GB per sweep vs 126 MB L2 (inputs larger than L2, no flush needed)"
                                 if n * p > 4e8 else "inputs fit in L2 (cache-resident workload; HBM roofline not meaningful)",
                           "gibbs_iters_per_s": nchains * args.steps / (ms_all * 1e-3),
                           "geometry": {k: tm[k] for k in ("ctas", "threads", "block", "rows_per_cta", "smem_bytes", "lookahead", "near_depth", "tile_stages", "record_stages")}},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source,
                             "kernel": ("ngp::gibbs_kernel<TUP> (one launch = one Gibbs iteration)" if args.kernel == "blocked" and kbreeds in (2, 4)
                                        else "ngp::joint_kernel (one launch = one Gibbs iteration)") if kbreeds else
                                       "ngp::gibbs_kernel (one launch = one Gibbs iteration)", "kernel_ms": k_ms,
                             "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": summarize_clocks(lines)}
        line["config"]["changed_effects_per_sweep"] = changed
        line["long_run"] = long_run
        if shard_rec is not None:
            line["sharded"] = shard_rec
        if sharded:
            line["per_marker_us"] = 1e3 * ms_all / args.steps / p
        if not args.no_cpu and not kbreeds:
            line["cpu_baseline"] = cpu_baseline(n, p, model, seed)
        print(json.dumps(line), flush=True)
    s.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
