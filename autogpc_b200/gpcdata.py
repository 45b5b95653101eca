"""Dataset container for AutoGPC (host side).  Behaviour of /root/reference/gpcdata.py in Python 3.

Difference: the k-fold shuffle takes an explicit seed (the reference's ``KFold(..., shuffle=True)`` draws from the
unseeded global RNG, gpcdata.py:170), and folds are also exposed as index lists (``kFoldIndices``) because the engine
addresses a fold as a row list of the device-resident dataset instead of a copied array.
"""
from __future__ import annotations

import numpy as np


class GPCData:
    """N data points, D input dimensions, {0,1} labels (gpcdata.py:7-38)."""

    def __init__(self, X, Y, XLabel=None, YLabel=None, seed: int = 0):
        assert isinstance(X, np.ndarray) and X.ndim == 2, "X must be a two-dimensional NumPy array"
        assert isinstance(Y, np.ndarray) and Y.ndim == 2 and Y.shape[1] == 1, "Y must be an N x 1 NumPy array"
        assert Y.shape[0] == X.shape[0], "X and Y must contain the same number of entries"
        self.X, self.Y = X, Y
        self.seed = seed
        if XLabel is not None and isinstance(XLabel, (list, tuple)) and len(XLabel) == X.shape[1]:
            self.XLabel = tuple(XLabel)
        else:
            self.XLabel = tuple("$x_{{{0}}}$".format(d + 1) for d in range(X.shape[1]))
        self.YLabel = YLabel if (YLabel is not None and isinstance(YLabel, (list, tuple)) and len(YLabel) == 2) \
            else ["negative", "positive"]
        self.xranges = None
        self.minseps = None
        self._fold_idx = None
        self._n_folds = None

    def __repr__(self):
        return ("GPCData: %d dimensions, %d data points.\n" % (self.getDim(), self.getNum())
                + "XLabel:\n" + ", ".join(self.XLabel) + "\nYLabel:\n" + ", ".join(self.YLabel)
                + "\nX Ranges:\n" + str(self.inputRange().flatten().tolist())
                + "\nX Minimum Separations:\n" + str(self.minSeparation().flatten().tolist()))

    def getNum(self) -> int:
        return self.X.shape[0]

    def getDim(self) -> int:
        return self.X.shape[1]

    def getClass(self, y):
        return self.X[self.Y[:, 0] == y]

    def getDataShape(self) -> dict:
        """gpcdata.py:72-86."""
        return {"x_mu": self.X.mean(axis=0).tolist(), "x_sd": self.X.std(axis=0).tolist(),
                "x_min": self.X.min(axis=0).tolist(), "x_max": self.X.max(axis=0).tolist(),
                "y_sd": float(self.Y.std()), "y_mean": float(self.Y.mean())}

    def inputRange(self, dims=None):
        if dims is None:
            dims = list(range(self.getDim()))
        if self.xranges is None:
            self.xranges = self.X.max(axis=0, keepdims=True) - self.X.min(axis=0, keepdims=True)
        return self.xranges[0, dims]

    def minSeparation(self, dims=None):
        """Minimum separation between distinct values per dimension (gpcdata.py:105-122)."""
        if dims is None:
            dims = list(range(self.getDim()))
        if self.minseps is None:
            self.minseps = np.array([[np.diff(np.unique(self.X[:, i])).min() for i in range(self.getDim())]])
        return self.minseps[0, dims]

    def getLengthscaleBounds(self, dims=None):
        """[min separation, 2 x range] (gpcdata.py:125-136)."""
        if dims is None:
            dims = list(range(self.getDim()))
        return np.vstack((self.minSeparation(dims=dims), self.inputRange(dims=dims) * 2))

    def getPeriodBounds(self, dims=None):
        return self.getLengthscaleBounds(dims=dims)

    def kFoldIndices(self, k: int = 5):
        """[(train_idx, test_idx)] * k; shuffled once with ``seed`` and cached per k (gpcdata.py:149-180).
        Fold sizes follow sklearn's KFold: the first N % k folds get one extra point."""
        n = self.getNum()
        assert 0 < k <= n, "Invalid number of folds"
        if self._fold_idx is not None and self._n_folds == k:
            return self._fold_idx
        if k == 1:
            idx = np.arange(n, dtype=np.int32)
            out = [(idx, idx)]
        else:
            perm = np.random.default_rng(self.seed).permutation(n)
            sizes = np.full(k, n // k)
            sizes[: n % k] += 1
            out, start = [], 0
            for s in sizes:
                test = np.sort(perm[start:start + s]).astype(np.int32)
                mask = np.ones(n, dtype=bool)
                mask[test] = False
                out.append((np.nonzero(mask)[0].astype(np.int32), test))
                start += s
        self._fold_idx, self._n_folds = out, k
        return out

    def kFoldSplits(self, k: int = 5):
        """(X, Y, XT, YT), each a list of k arrays -- the reference's return shape (gpcdata.py:149-180)."""
        folds = self.kFoldIndices(k)
        return ([self.X[tr] for tr, _ in folds], [self.Y[tr] for tr, _ in folds],
                [self.X[te] for _, te in folds], [self.Y[te] for _, te in folds])
