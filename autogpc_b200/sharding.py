"""Multi-GPU sharding of the candidate x fold items.

Items are independent, so there is no data-path collective: every rank holds the (tiny) dataset, optimises its share of
the items -- dealt by estimated cost (``deal_items``, longest-processing-time first) or pulled chunk by chunk from a
shared counter (``WorkQueue``) -- and ONE all-gather per search depth exchanges fixed-width records
``(item id, nlml, bic, cv error, status, n_theta, theta[MAX_THETA])``; afterwards every rank holds every result and
performs the identical sort / beam cut, so no broadcast is needed.  The collective runs through ``torch.distributed``
(NCCL on the GPU box, gloo in the CPU tests) -- plumbing only.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .engine import MAX_THETA

RECORD_WIDTH = 6 + MAX_THETA


def world() -> Tuple[int, int]:
    """(rank, world_size) of the torch.distributed job.  Under torchrun (WORLD_SIZE > 1 in the environment) with no
    process group yet, the group is created here -- NCCL on this rank's GPU (LOCAL_RANK) when CUDA is present, gloo
    otherwise -- so that a sharded search never starts optimising before the exchange step is known to work.  Without
    a launcher: (0, 1)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1:
        return 0, 1
    import torch
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if torch.cuda.is_available():
        dev = torch.device("cuda", local_device())
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    return dist.get_rank(), dist.get_world_size()


def local_device() -> int:
    """This rank's GPU ordinal: LOCAL_RANK under torchrun, else 0."""
    return int(os.environ.get("LOCAL_RANK", "0"))


def item_cost(n_rows: int, n_leaves: int = 1, n_periodic: int = 0) -> float:
    """A-priori cost estimate of scoring / optimising one (candidate, fold) item (SURVEY 8e "P n^3 estimate"): the n^3
    factorisations dominate; EP needs more sweeps on composite and periodic structures (measured on the bench block:
    3 sweeps for one SE leaf, 4-5 for sums / products, up to 9 with a periodic leaf)."""
    return float(n_rows) ** 3 * (3.0 + 0.5 * (n_leaves - 1) + 1.5 * n_periodic)


def deal_items(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first dealing: items sorted by decreasing cost (ties by id), each to the least loaded
    rank (ties by rank).  Deterministic, so every rank computes the same table.  Returns the item ids per rank, sorted."""
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in sorted(range(len(costs)), key=lambda j: (-costs[j], j)):
        r = min(range(world_size), key=lambda q: (load[q], len(out[q]), q))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(x) for x in out]


def shard_indices(n_items: int, rank: int, world_size: int, costs: Optional[Sequence[float]] = None) -> List[int]:
    """Item ids of ``rank``: cost-aware dealing when ``costs`` is given, plain round-robin by id otherwise."""
    if costs is not None:
        assert len(costs) == n_items
        return deal_items(costs, world_size)[rank]
    return [i for i in range(n_items) if i % world_size == rank]


class WorkQueue:
    """Dynamic dealing: chunks of work are claimed with an atomic fetch-and-add on a counter in the process group's
    key-value store (one ~100 us round trip per chunk to rank 0's TCPStore; no collective, so a rank that finishes early
    simply claims more).  Every rank must create its queues in the same order -- the counter key is derived from a
    per-process sequence number.  Single-process: a local counter."""
    _seq = 0

    def __init__(self, n_chunks: int, tag: str = "q"):
        import torch.distributed as dist
        WorkQueue._seq += 1
        self.n = int(n_chunks)
        self.key = "agpc/%s/%d" % (tag, WorkQueue._seq)
        self.store = None
        self._local = 0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from torch.distributed import distributed_c10d as c10d
            self.store = c10d._get_default_store()

    def claim(self) -> Optional[int]:
        """Next unclaimed chunk id, or None when the queue is drained."""
        if self.store is None:
            c = self._local
            self._local += 1
        else:
            c = int(self.store.add(self.key, 1)) - 1
        return c if c < self.n else None


def pack_record(item_id: int, nlml: float, bic: float, err: float, status: int, theta: Sequence[float]) -> np.ndarray:
    r = np.zeros(RECORD_WIDTH)
    r[0], r[1], r[2], r[3], r[4], r[5] = item_id, nlml, bic, err, status, len(theta)
    r[6:6 + len(theta)] = theta
    return r


def allgather_records(local: np.ndarray, n_total: int, per_rank_max: Optional[int] = None) -> np.ndarray:
    """All ranks contribute ``local`` ([n_local, RECORD_WIDTH], ids in column 0) and receive the [n_total, W] table
    ordered by item id.  Ranks may hold different counts: each pads to ``per_rank_max`` (default ceil(n_total / world),
    which covers round-robin dealing; cost-aware or dynamic dealing passes its own bound, at most n_total)."""
    import torch
    import torch.distributed as dist
    rank, ws = dist.get_rank(), dist.get_world_size()
    per = per_rank_max if per_rank_max is not None else (n_total + ws - 1) // ws
    assert len(local) <= per, "rank holds %d records, exchange buffer sized for %d" % (len(local), per)
    buf = np.full((per, RECORD_WIDTH), np.nan)
    buf[:, 0] = -1
    buf[:len(local)] = local
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", local_device()) if use_cuda else torch.device("cpu")
    t = torch.from_numpy(buf).to(dev)
    out = torch.empty((ws * per, RECORD_WIDTH), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, t)
    allr = out.cpu().numpy()
    allr = allr[allr[:, 0] >= 0]
    allr = allr[np.argsort(allr[:, 0], kind="stable")]
    assert allr.shape[0] == n_total, "all-gather lost records: %d of %d" % (allr.shape[0], n_total)
    return allr


def exchange_models(items, mine: List[int], errs: Dict[int, float], shard: Tuple[int, int]) -> None:
    """After local optimisation: publish (nlml, bic, error, status, theta*) of my items, adopt everyone else's.
    Remote models receive theta* and their scores; EP sites are not shipped (n doubles per item) -- a remote model
    re-derives them with one cold inference the first time it has to predict."""
    local = np.zeros((len(mine), RECORD_WIDTH))
    for j, i in enumerate(mine):
        m = items[i][2]
        local[j] = pack_record(i, m._nlml, m.bic, errs[i], m.status, m.kern.param_array)
    table = allgather_records(local, len(items), per_rank_max=len(items))
    mine_set = set(mine)
    for row in table:
        i = int(row[0])
        if i in mine_set:
            continue
        m = items[i][2]
        p = int(row[5])
        m.kern.set_param_array(row[6:6 + p])
        m._nlml, m.bic, m.status = float(row[1]), float(row[2]), int(row[4])
        m._tau = m._nu = None
        errs[i] = float(row[3])
