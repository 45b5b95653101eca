"""Lock-step batched driver for ``m.optimize()`` (gpckernel.py:246).

GPy optimises one model at a time with SciPy's L-BFGS-B (paramz ``opt_lbfgsb``: m = 10, factr = 1e7, pgtol = 1e-5,
<= 1000 evaluations; <= 10 linear-algebra failures tolerated).  To keep those semantics bit-for-bit while giving the
GPU a batch, every model gets its own SciPy L-BFGS-B instance in a host thread; each objective request parks its
thread, and a coordinator turns the set of parked requests into ONE ``score_batch`` call.  Threads only wait on the
GPU, so the GIL is irrelevant.
"""
from __future__ import annotations

import threading
from typing import Callable, List, Sequence, Tuple

import numpy as np
from scipy.optimize import fmin_l_bfgs_b

FAIL_VALUE = 1e300
ALLOWED_FAILURES = 10


class _Lockstep:
    def __init__(self, n: int, eval_batch: Callable[[List[int], List[np.ndarray]], List[Tuple[float, np.ndarray]]]):
        self.n = n
        self.eval_batch = eval_batch
        self.cv = threading.Condition()
        self.pending = {}
        self.results = {}
        self.alive = n
        self.error = None

    def request(self, i: int, x: np.ndarray):
        with self.cv:
            self.pending[i] = np.array(x, dtype=float)
            self.cv.notify_all()
            while i not in self.results and self.error is None:
                self.cv.wait()
            if self.error is not None:
                raise RuntimeError("batched evaluation failed") from self.error
            return self.results.pop(i)

    def finish(self, i: int):
        with self.cv:
            self.alive -= 1
            self.cv.notify_all()

    def coordinate(self):
        while True:
            with self.cv:
                while self.alive > 0 and len(self.pending) < self.alive:
                    self.cv.wait()
                if self.alive == 0:
                    return
                idx = sorted(self.pending)
                xs = [self.pending.pop(i) for i in idx]
            try:
                out = self.eval_batch(idx, xs)
            except BaseException as e:  # propagate to every parked worker
                with self.cv:
                    self.error = e
                    self.cv.notify_all()
                raise
            with self.cv:
                for i, r in zip(idx, out):
                    self.results[i] = r
                self.cv.notify_all()


def lbfgsb_lockstep(eval_batch, x0s: Sequence[np.ndarray], max_evals: int = 1000, m: int = 10, factr: float = 1e7,
                    pgtol: float = 1e-5):
    """Minimise len(x0s) independent objectives; ``eval_batch(indices, xs) -> [(f, g), ...]`` is called with every
    currently outstanding request.  Returns [(x, f, info)] in input order."""
    n = len(x0s)
    ls = _Lockstep(n, eval_batch)
    out = [None] * n

    def work(i):
        try:
            out[i] = fmin_l_bfgs_b(lambda x: ls.request(i, x), np.array(x0s[i], float), maxfun=max_evals, m=m,
                                   factr=factr, pgtol=pgtol)
        except BaseException as e:
            out[i] = e
        finally:
            ls.finish(i)

    threads = [threading.Thread(target=work, args=(i,), daemon=True) for i in range(n)]
    for t in threads:
        t.start()
    ls.coordinate()
    for t in threads:
        t.join()
    return out


def optimize_models(models, max_evals: int = 1000, warm_start: bool = True):
    """Optimise GPClassification shims in lock-step.  Each evaluation of every live model is one item of one
    ``score_batch``; EP is warm-started from the model's previous sites.  Models must share an engine; they may
    live on different datasets only if the engine call is split, so they are grouped by dataset id."""
    if not models:
        return
    engine = models[0].engine
    fails = [0] * len(models)
    evals = [0] * len(models)

    def eval_batch(idx, xs):
        by_ds = {}
        for i, x in zip(idx, xs):
            by_ds.setdefault(models[i]._dataset, []).append((i, x))
        res = {}
        for ds, group in by_ds.items():
            progs, thetas, rows, taus, nus, params = [], [], [], [], [], []
            all_warm = warm_start and all(models[i]._tau is not None for i, _ in group)
            for i, x in group:
                mdl = models[i]
                ps = mdl.kern.flat_params()
                theta = np.array([p.from_x(float(v)) for p, v in zip(ps, x)])
                mdl.kern.set_param_array(theta)
                progs.append(mdl._program)
                thetas.append(theta)
                rows.append(mdl._rows)
                taus.append(mdl._tau)
                nus.append(mdl._nu)
                params.append(ps)
            r = engine.score_batch(ds, progs, thetas, rows, tau=taus if all_warm else None,
                                   nu=nus if all_warm else None, **models[group[0][0]].ep)
            for b, (i, x) in enumerate(group):
                mdl = models[i]
                evals[i] += 1
                st = int(r["status"][b])
                if st in (2, 4) or not np.isfinite(r["nlml"][b]):
                    fails[i] += 1
                    if fails[i] > ALLOWED_FAILURES:
                        raise np.linalg.LinAlgError("too many failed evaluations")
                    mdl.status = st
                    res[i] = (FAIL_VALUE, np.zeros_like(x))
                    continue
                mdl._absorb(r, b)
                gf = np.array([p.gradfactor() for p in params[b]])
                res[i] = (mdl._nlml, mdl._grad * gf)
        return [res[i] for i in idx]

    x0s = [np.array([p.to_x() for p in mdl.kern.flat_params()]) for mdl in models]
    out = lbfgsb_lockstep(eval_batch, x0s, max_evals=max_evals)
    # leave every model at its optimum (the last evaluation is not necessarily the best one)
    finals = []
    for i, (mdl, o) in enumerate(zip(models, out)):
        if isinstance(o, BaseException):
            mdl.optimization_error = o
            continue
        x, f, info = o
        ps = mdl.kern.flat_params()
        mdl.kern.set_param_array([p.from_x(float(v)) for p, v in zip(ps, x)])
        mdl.optimization_info = dict(info, evals=evals[i])
        finals.append(mdl)
    by_ds = {}
    for mdl in finals:
        by_ds.setdefault(mdl._dataset, []).append(mdl)
    for ds, group in by_ds.items():
        warm = warm_start and all(m._tau is not None for m in group)
        r = engine.score_batch(ds, [m._program for m in group], [m.kern.param_array for m in group],
                               [m._rows for m in group], tau=[m._tau for m in group] if warm else None,
                               nu=[m._nu for m in group] if warm else None, **group[0].ep)
        for b, mdl in enumerate(group):
            mdl._absorb(r, b)
