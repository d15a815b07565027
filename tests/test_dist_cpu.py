"""CPU (gloo, world_size 2): the multi-GPU path shards CHAINS, one per rank, with no data-path collective
(DESIGN.md §5).  What runs across ranks is only (i) the timing barrier / max-over-ranks of bench.py and
(ii) the final reduction of posterior sums.  Both are exercised here on gloo; the per-chain streams are
checked to be disjoint and reproducible per chain id."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank runs its own chain (chain id = rank) on the same data: here the CPU oracle stands in for the device chain
    from common import make_problem, oracle_chain
    prob = make_problem(120, 40, 3)
    ch, S = oracle_chain(prob, 2, 0.05, pi=0.1, est_pi=True)
    sb, k = np.zeros(40), 0
    for it in range(60):
        ch.iteration(seed=77, chain=rank)
        if it >= 20:
            sb += S.beta; k += 1
    # bench.py timing rule: max over ranks
    tms = torch.tensor([10.0 + rank])
    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    # posterior sums pooled over chains at the end of the run
    pooled = torch.from_numpy(sb.copy()); cnt = torch.tensor([float(k)])
    dist.all_reduce(pooled); dist.all_reduce(cnt)
    gathered = [torch.zeros(40, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(sb / k))
    if rank == 0:
        q.put((float(tms.item()), (pooled / cnt).numpy(), [g.numpy() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_chain_partition_over_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    tmax, pooled, per_chain = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 11.0                                   # max over ranks
    assert not np.allclose(per_chain[0], per_chain[1])    # chains are independent streams
    assert np.allclose(pooled, (per_chain[0] + per_chain[1]) / 2)
    # chain id -> reproducible stream: rank 1's chain equals a fresh oracle chain with chain=1
    from common import make_problem, oracle_chain
    prob = make_problem(120, 40, 3)
    ch, S = oracle_chain(prob, 2, 0.05, pi=0.1, est_pi=True)
    sb, k = np.zeros(40), 0
    for it in range(60):
        ch.iteration(seed=77, chain=1)
        if it >= 20:
            sb += S.beta; k += 1
    assert np.allclose(sb / k, per_chain[1])


def test_streams_of_different_chains_do_not_collide():
    L = O.lib()
    a = [L.ngo_stream_uniform(5, 0, 1, 0, O.P_U, i) for i in range(256)]
    b = [L.ngo_stream_uniform(5, 1, 1, 0, O.P_U, i) for i in range(256)]
    assert len(set(a) & set(b)) == 0


# ----------------------------------------------------------------------------- row-sharded single chain, host-side logic on gloo
def _shard_worker(rank, world, port, q):
    """What bench.py --sharded does around the device calls, with numpy standing in for the kernel: rows split on multiples of 4,
    integer column sums all-reduced -> the mean / mpm of the WHOLE column on every rank, and per marker ONE all-reduced scalar
    (the partial dot of the rank's rows); every rank then draws the identical effect and updates only its own rows of e."""
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from common import make_problem
    n, p = 203, 30
    prob = make_problem(n, p, 9)
    per = -(-(-(-n // world)) // 4) * 4
    a, b = min(n, rank * per), min(n, (rank + 1) * per)
    g = prob["codes"][a:b].astype(np.int64)
    sums = torch.from_numpy(np.stack([g.sum(0), (g * g).sum(0)]))
    dist.all_reduce(sums)
    cs, css = sums[0].numpy().astype(np.float64), sums[1].numpy().astype(np.float64)
    mean, mpm = cs / n, (n * css - cs * cs) / n                   # colstats_kernel of csrc/ngp_api.cu with n_total
    x = g - mean
    e = (prob["y"] - prob["y"].mean())[a:b].copy()
    rng = np.random.default_rng(5)                                # the same variates on every rank (same Philox counter on device)
    beta = np.zeros(p); varE, vb = 1.3, 0.02
    for j in range(p):
        part = torch.tensor([x[:, j] @ e])
        dist.all_reduce(part)                                     # the per-marker scalar reduction
        rr = part.item() + mpm[j] * beta[j]
        lhs = mpm[j] / varE + 1.0 / vb
        bn = (rr / varE) / lhs + rng.standard_normal() / np.sqrt(lhs)
        e -= x[:, j] * (bn - beta[j])
        beta[j] = bn
    parts = [None] * world
    dist.all_gather_object(parts, (a, b, e))
    if rank == 0:
        q.put((mean, mpm, beta, parts))
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_chain_logic_over_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    mean, mpm, beta, parts = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    from common import make_problem
    prob = make_problem(203, 30, 9)
    X, mean_o, mpm_o = O.center_codes(prob["codes"])
    assert np.allclose(mean, mean_o, rtol=1e-14) and np.allclose(mpm, mpm_o, rtol=1e-12)
    assert parts[0][0] == 0 and parts[0][1] % 4 == 0 and parts[0][1] == parts[1][0] and parts[1][1] == 203      # slices tile the rows
    # single-process restatement of the same sweep on all rows
    e = prob["y"] - prob["y"].mean(); rng = np.random.default_rng(5); b1 = np.zeros(30)
    for j in range(30):
        rr = X[:, j] @ e + mpm_o[j] * b1[j]
        lhs = mpm_o[j] / 1.3 + 1.0 / 0.02
        bn = (rr / 1.3) / lhs + rng.standard_normal() / np.sqrt(lhs)
        e -= X[:, j] * (bn - b1[j]); b1[j] = bn
    assert np.allclose(beta, b1, rtol=1e-9) and np.allclose(np.concatenate([pt[2] for pt in parts]), e, rtol=1e-8, atol=1e-10)


# ----------------------------------------------------------------------------- row-sharded chain: the host-side exchange steps (round 2)
def _exchange_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from common import make_problem
    n, p, B, D = 203, 96, 32, 2
    prob = make_problem(n, p, 5)
    codes = prob["codes"].astype(np.int64)
    per = -(-(-(-n // world)) // 4) * 4                      # bench.py / ShardedChain: shard boundaries on multiples of 4 rows
    a, b = min(n, rank * per), min(n, (rank + 1) * per)
    loc = codes[a:b]
    # (i) whole-column statistics from the all-reduced integer column sums (ngp_get_column_sums -> all-reduce -> ngp_set_column_sums)
    t = torch.from_numpy(np.stack([loc.sum(0), (loc * loc).sum(0)]))
    dist.all_reduce(t)
    cs, css = t[0].numpy().astype(np.float64), t[1].numpy().astype(np.float64)
    mean, mpm = cs / n, (n * css - cs * cs) / n
    # (ii) the banded raw Gram of the blocked kernel is a sum over individuals: gx[k][d][a][b] = sum_i g_{(k-d)B+a,i} g_{kB+b,i}
    nblk = p // B
    gx = np.zeros((nblk, D + 1, B, B), dtype=np.int64)
    for k in range(nblk):
        for d in range(D + 1):
            if k - d >= 0:
                gx[k, d] = loc[:, (k - d) * B:(k - d + 1) * B].T @ loc[:, k * B:(k + 1) * B]
    g = torch.from_numpy(gx.astype(np.int32))
    sizes = [None] * world
    dist.all_gather_object(sizes, int(g.numel()))            # bench.allreduce_gram: all ranks must have chosen the same block size / look-ahead
    dist.all_reduce(g)
    # (iii) every rank must hold the bitwise identical chain: all-gather and compare as integers (bench.sharded_leg)
    mine = torch.from_numpy(np.concatenate([mean, mpm]))
    allb = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allb, mine)
    identical = all(bool(torch.equal(allb[0].view(torch.int64), x.view(torch.int64))) for x in allb)
    q.put((rank, a, b, mean, mpm, g.numpy(), sizes, identical))
    dist.destroy_process_group()


def test_row_sharding_exchange_steps_gloo_world2():
    """the set-up collectives of the row-sharded chain on CPU (gloo, world size 2): row split, column-sum all-reduce, Gram all-reduce"""
    from common import make_problem
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    prob = make_problem(203, 96, 5)
    X, mean_o, mpm_o = O.center_codes(prob["codes"])
    codes = prob["codes"].astype(np.int64)
    assert res[0][1] == 0 and res[0][2] == res[1][1] == 104 and res[1][2] == 203 and res[0][2] % 4 == 0
    for r in res:
        assert np.allclose(r[3], mean_o, rtol=1e-14) and np.allclose(r[4], mpm_o, rtol=1e-12)
        assert len(set(r[6])) == 1 and r[7]
    full = codes[:, 0:32].T @ codes[:, 32:64]                 # gx[1][1]: block 0 against block 1, all rows
    assert np.array_equal(res[0][5][1, 1], full) and np.array_equal(res[0][5], res[1][5])
