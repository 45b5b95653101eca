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
