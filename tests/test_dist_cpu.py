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
