"""N > 1 path on CPU: world_size-2 gloo job, items sharded round-robin, one all-gather of (score, theta) records;
every rank must end up with the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from autogpc_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _job(seed=21):
    from autogpc_b200 import engine as eng_mod
    from autogpc_b200 import flexible_function as ff
    from autogpc_b200 import gpckernel
    from autogpc_b200.gpcdata import GPCData
    from tests.fake_engine import OracleEngine
    from tests.util import synth
    eng_mod.set_default_engine(OracleEngine())
    gpckernel.set_rng(seed)
    X, y = synth(64, 2, 9)
    d = GPCData(X, y.reshape(-1, 1), seed=1)
    ks = [gpckernel.GPCKernel(k, d, depth=1) for k in
          (ff.SqExpKernel(0), ff.SqExpKernel(1), ff.ProductKernel([ff.SqExpKernel(0), ff.SqExpKernel(1)]))]
    return ks, gpckernel


def _summary(ks):
    return [(k.kernel.structure(), round(k.error(), 12), [round(e, 12) for e in k.foldErrors],
             np.round(k.model.kern.param_array, 9).tolist(), round(k.getNLML(), 7)) for k in ks]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ks, gpckernel = _job()
        gpckernel.train_kernels(ks, mode="full", n_folds=3, max_evals=25, shard=(rank, world))
        q.put((rank, _summary(ks)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    ks, gpckernel = _job()
    gpckernel.train_kernels(ks, mode="full", n_folds=3, max_evals=25)
    ref = _summary(ks)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == got[1]                     # every rank holds every result, identically ordered
    assert got[0] == ref                        # and it is the single-process result


def test_shard_indices_and_records():
    assert sharding.shard_indices(7, 1, 3) == [1, 4]
    assert sorted(sum((sharding.shard_indices(11, r, 4) for r in range(4)), [])) == list(range(11))
    r = sharding.pack_record(5, 1.5, 2.5, 0.25, 0, [1.0, 2.0, 3.0])
    assert r.shape == (sharding.RECORD_WIDTH,) and r[5] == 3 and r[6:9].tolist() == [1.0, 2.0, 3.0]
    assert sharding.world() == (0, 1)


def test_cost_aware_dealing_balances_and_is_deterministic():
    costs = [sharding.item_cost(6400, nl, npr) for nl, npr in [(1, 0), (1, 1), (2, 0), (2, 1), (3, 1), (1, 0), (2, 0), (3, 0)] * 5]
    parts = sharding.deal_items(costs, 8)
    assert sorted(sum(parts, [])) == list(range(len(costs)))
    load = [sum(costs[i] for i in p) for p in parts]
    assert max(load) <= 1.10 * (sum(costs) / 8)          # LPT: within 10 % of the mean on this mix
    rr = [sum(costs[i] for i in range(len(costs)) if i % 8 == r) for r in range(8)]
    assert max(load) <= max(rr)                          # never worse than round-robin
    assert parts == sharding.deal_items(costs, 8)
    assert sharding.shard_indices(len(costs), 3, 8, costs) == parts[3]


def _queue_worker(rank, world, port, q):
    import time
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = []
        for n in (7, 13):                        # two queues in a row: keys must not collide
            wq = sharding.WorkQueue(n, tag="test")
            mine = []
            while True:
                c = wq.claim()
                if c is None:
                    break
                mine.append(c)
                time.sleep(0.002 * (rank + 1))   # rank 1 is slower: it must end up with fewer chunks
            out.append(mine)
            dist.barrier()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_work_queue_deals_every_chunk_exactly_once():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_queue_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k, n in enumerate((7, 13)):
        assert sorted(got[0][k] + got[1][k]) == list(range(n))
    assert len(got[0][1]) >= len(got[1][1])
    wq = sharding.WorkQueue(3)                   # single process: a plain counter
    assert [wq.claim() for _ in range(4)] == [0, 1, 2, None]
