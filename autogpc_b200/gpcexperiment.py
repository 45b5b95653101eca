"""``GPCExperiment``: the experiment drivers of /root/reference/gpcexperiment.py:41-106 on the CUDA scoring engine.

Thin shim.  The reference's four methods download UCI data through ``pods`` (network) and render a
LaTeX report through ``pylatex`` / ``pdflatex`` -- both absent here and out of this path's scope -- so every method
takes the data as arrays (or a loader callable) and returns the search result instead of writing a PDF:

    exp = GPCExperiment()
    out = exp.pima(X, Y, XLabel=[...])            # same defaults as gpcexperiment.py:50 (depth 4, beam 2)
    out["best"], out["best1d"], out["baseline"], out["seconds"]

``run`` is the generic driver the four named methods share (search + constant-kernel baseline, gpcexperiment.py:56-58).
"""
from __future__ import annotations

import time

import numpy as np

from .gpcdata import GPCData
from .gpcsearch import GPCSearch


def timing(f):
    """Wall-clock decorator of gpcexperiment.py:13-20 (prints, returns the wrapped function's value)."""
    def wrap(*args, **kw):
        t0 = time.time()
        ret = f(*args, **kw)
        print("%s function took %0.3f seconds." % (f.__name__, time.time() - t0))
        return ret
    wrap.__name__ = f.__name__
    return wrap


def repeat(f):
    """Retry decorator of gpcexperiment.py:24-38 (<= 5 attempts); unused by default, exactly as in the reference."""
    def wrap(*args, **kw):
        fails = 0
        while fails < 5:
            try:
                return f(*args, **kw)
            except Exception:
                fails += 1
        print("%s function failed after %d times." % (f.__name__, fails))
    wrap.__name__ = f.__name__
    return wrap


class GPCExperiment(object):
    """Experiments for AutoGPC (gpcexperiment.py:41-106)."""

    def __init__(self, base_kernels="SE", criterion="error", n_folds=5, max_evals=1000, verbose=True):
        self.opts = dict(base_kernels=base_kernels, criterion=criterion, n_folds=n_folds, max_evals=max_evals,
                         verbose=verbose)

    def run(self, name, X, Y, XLabel=None, YLabel=None, dims=None, depth=4, width=2):
        X = np.asarray(X, float)
        if dims is not None:
            X = X[:, list(dims)]
            if XLabel is not None:
                XLabel = [XLabel[i] for i in dims]
        d = GPCData(X, np.asarray(Y, float).reshape(-1, 1), XLabel=XLabel, YLabel=YLabel)
        if self.opts["verbose"]:
            print("\n\n%s dataset" % name)
            print("=====\nData size: D = %d, N = %d." % (d.getDim(), d.getNum()))
        t0 = time.time()
        search = GPCSearch(data=d, max_depth=depth, beam_width=width, **self.opts)
        best, best1d = search.search()
        constker = search.baseline()
        return dict(name=name, data=d, best=best, best1d=best1d, baseline=constker, history=search.history,
                    summands=getattr(search, "summands", []), seconds=time.time() - t0)

    @timing
    def pima(self, X, Y, XLabel=None, dims=(1, 5, 6, 7), depth=4, width=2):
        return self.run("Pima Indian Diabetes", X, Y, XLabel, ["not diabetic", "diabetic"], dims, depth, width)

    @timing
    def wisconsin(self, X, Y, XLabel=None, dims=(0, 1, 4, 5, 7), depth=4, width=2):
        return self.run("Wisconsin Breast Cancer", X, Y, XLabel, ["no breast cancer", "with breast cancer"], dims, depth, width)

    @timing
    def bupa(self, X, Y, XLabel=None, dims=tuple(range(5)), depth=4, width=2):
        return self.run("BUPA Liver Disorders", X, Y, XLabel, ["$\\leq 5$ drink units", "$> 5$ drink units"], dims, depth, width)

    @timing
    def cleveland(self, X, Y, XLabel=None, dims=(0, 3, 4, 7, 9, 11), depth=4, width=2):
        # the reference passes depth=5 in the signature but hard-codes 4 in the call (gpcexperiment.py:96,102)
        return self.run("Cleveland Heart Disease", X, Y, XLabel, ["no heart disease", "with heart disease"], dims, 4, width)
