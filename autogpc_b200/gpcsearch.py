"""AutoGPC kernel search (host side).  Python 3 mirror of /root/reference/gpcsearch.py on the CUDA scoring engine.

Same flow as ``GPCSearch.search`` (gpcsearch.py:25-96): start from ``NoneKernel``; per depth expand every beam member,
keep candidates that add a new dimension, train them all, stable-sort, stop when the best candidate does not improve,
keep ``beam_width``; then split the winner into summands and refresh the best 1-D kernels.

Differences: ``base_kernels`` is actually passed to ``expand`` (the reference stores it and then
calls ``k.expand()`` with the default 'SE', gpcsearch.py:38); ``mode`` defaults to 'full' (the reference's 'auto' would
route every benchmark shape to SVGP, gpckernel.py:181-183); the ranking criterion is selectable -- ``'error'``
(reference: mean cross-validated error, gpcsearch.py:49) or ``'bic'`` (north_star) -- with ties broken on the canonical
structure string so the order is deterministic; all candidates x folds of a depth are trained as one lock-step
batch, sharded over the ranks of a torchrun job.
"""
from __future__ import annotations

from typing import List

import numpy as np

from . import flexible_function as ff
from . import sharding
from .gpcdata import GPCData
from .gpckernel import GPCKernel, train_kernels


class GPCSearch:
    def __init__(self, data=None, base_kernels="SE", max_depth=3, beam_width=1, criterion="error", mode="full",
                 n_folds=5, max_evals=1000, verbose=False):
        assert isinstance(data, GPCData), "data must be of type GPCData"
        assert criterion in ("error", "bic"), "criterion must be 'error' or 'bic'"
        self.data = data
        self.baseKernels = base_kernels
        self.maxDepth = max_depth
        self.beamWidth = beam_width
        self.criterion = criterion
        self.mode = mode
        self.nFolds = n_folds
        self.maxEvals = max_evals
        self.verbose = verbose
        self.history: List[List[GPCKernel]] = []     # all scored candidates per depth, in ranked order

    def _say(self, *a):
        if self.verbose and sharding.world()[0] == 0:
            print(*a)

    def score(self, k: GPCKernel) -> float:
        return k.error() if self.criterion == "error" else k.bic()

    def _rank(self, kernels: List[GPCKernel]) -> List[GPCKernel]:
        # the reference's sort is stable on the un-rounded error; structure string makes ties reproducible
        return sorted(kernels, key=lambda x: (self.score(x), x.kernel.canonical().structure()))

    def _train(self, kernels: List[GPCKernel]) -> None:
        rank, ws = sharding.world()
        train_kernels(kernels, mode=self.mode, n_folds=self.nFolds, max_evals=self.maxEvals,
                      shard=(rank, ws) if ws > 1 else None)

    def search(self):
        depth = 0
        kernels = [GPCKernel(ff.NoneKernel(), self.data, depth=depth)]
        best = [kernels[0]]
        best_score = [float("inf")]
        kernels1d: List[GPCKernel] = []
        self._say("\n=====\nSearch begins:")
        while depth < self.maxDepth:
            newkernels: List[GPCKernel] = []
            for k in kernels:
                # enforce a new dimension at every level (gpcsearch.py:38)
                newkernels.extend(x for x in k.expand(base_kernels=self.baseKernels) if len(x.getActiveDims()) > depth)
            if depth == 0:
                kernels1d = newkernels
            kernels = newkernels
            if not kernels:
                break
            self._train(kernels)
            self._say("\n=====\nFully expanded kernel set at depth %d:" % (depth + 1))
            self._say("\n".join(str(k) for k in kernels))
            kernels = self._rank(kernels)
            self.history.append(list(kernels))
            s0 = self.score(kernels[0])
            prev = best_score[-1] if self.criterion == "bic" else best[-1].error()
            if s0 >= prev:
                break
            best.append(kernels[0])
            best_score.append(s0)
            kernels = kernels[:self.beamWidth]
            depth += 1

        # refresh 1-D kernels from the winner's additive components (gpcsearch.py:62-79)
        summands = best[-1].toSummands() if not isinstance(best[-1].kernel, ff.NoneKernel) else []
        for k in [s for s in summands if len(s.getActiveDims()) == 1]:
            for i in range(len(kernels1d)):
                if k.equals(kernels1d[i]) and k.betterThan(kernels1d[i]):
                    kernels1d[i] = k
                    break
        best1d = bestKernels1D(kernels1d)
        if len(best[-1].getActiveDims()) == 1:
            for k in best1d:
                if k.equals(best[-1]) and k.betterThan(best[-1]):
                    best[-1] = k
                    if summands:
                        summands[0] = k
                    break
        self._say("\n=====\nSearch completed. Best kernels at each depth:")
        for k in best:
            self._say(k)
        self.summands = summands
        return best, best1d

    def baseline(self):
        """Constant-kernel baseline (gpcsearch.py:99-109)."""
        k = GPCKernel(ff.ConstKernel(), self.data)
        k.train(mode=self.mode, n_folds=self.nFolds, max_evals=self.maxEvals)
        return k


def bestKernels1D(kernels):
    """Best 1-D kernel per dimension by cross-validated error, ranked (gpcsearch.py:118-143)."""
    if len(kernels) == 0:
        return []
    assert all(len(k.getActiveDims()) == 1 for k in kernels), "All kernels must be one-dimensional."
    data = kernels[0].data
    best1d = [GPCKernel(ff.NoneKernel(), data) for _ in range(data.getDim())]
    for k in kernels:
        dim = k.getActiveDims()[0]
        if k.error() < best1d[dim].error():
            best1d[dim] = k
    best1d.sort(key=lambda k: k.error())
    return best1d
