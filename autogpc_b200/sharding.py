"""Multi-GPU sharding of the candidate x fold items.

Items are independent, so there is no data-path collective: every rank holds the (tiny) dataset, optimises the items
``i % world == rank``, and ONE all-gather per search depth exchanges fixed-width records
``(item id, nlml, bic, cv error, status, n_theta, theta[MAX_THETA])``; afterwards every rank holds every result and
performs the identical sort / beam cut, so no broadcast is needed.  The collective runs through ``torch.distributed``
(NCCL on the GPU box, gloo in the CPU tests) -- plumbing only.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import numpy as np

from .engine import MAX_THETA

RECORD_WIDTH = 6 + MAX_THETA


def world() -> Tuple[int, int]:
    """(rank, world_size) from torch.distributed if initialised, else from the torchrun environment, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_indices(n_items: int, rank: int, world_size: int) -> List[int]:
    """Round-robin by global item id (deterministic; cost-balanced because folds of one candidate cost the same)."""
    return [i for i in range(n_items) if i % world_size == rank]


def pack_record(item_id: int, nlml: float, bic: float, err: float, status: int, theta: Sequence[float]) -> np.ndarray:
    r = np.zeros(RECORD_WIDTH)
    r[0], r[1], r[2], r[3], r[4], r[5] = item_id, nlml, bic, err, status, len(theta)
    r[6:6 + len(theta)] = theta
    return r


def allgather_records(local: np.ndarray, n_total: int) -> np.ndarray:
    """All ranks contribute ``local`` ([n_local, RECORD_WIDTH], ids in column 0) and receive the [n_total, W] table
    ordered by item id.  Ranks may hold different counts: each pads to ceil(n_total / world)."""
    import torch
    import torch.distributed as dist
    rank, ws = dist.get_rank(), dist.get_world_size()
    per = (n_total + ws - 1) // ws
    buf = np.full((per, RECORD_WIDTH), np.nan)
    buf[:, 0] = -1
    buf[:len(local)] = local
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
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
    table = allgather_records(local, len(items))
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
