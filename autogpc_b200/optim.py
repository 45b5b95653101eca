"""Lock-step batched driver for ``m.optimize()`` (gpckernel.py:246).

GPy optimises one model at a time with SciPy's L-BFGS-B (paramz ``opt_lbfgsb``: m = 10, factr = 1e7, pgtol = 1e-5,
<= 1000 evaluations; <= 10 linear-algebra failures tolerated).  To keep those semantics bit-for-bit while giving the
GPU a batch, every model owns the state of one SciPy L-BFGS-B instance and is advanced through SciPy's own
reverse-communication kernel (``scipy.optimize._lbfgsb.setulb``, the routine ``fmin_l_bfgs_b`` loops over) until it
asks for an objective value; all outstanding requests of a round then become ONE ``score_batch`` call.  No threads:
the first version parked one ``fmin_l_bfgs_b`` per model on a condition variable and spent 55 ms per round for 90
models in wake-up storms -- more than the GPU needed for the evaluations at n = 1600.
"""
from __future__ import annotations

import os
from typing import Callable, List, Sequence, Tuple

import numpy as np
from scipy.optimize import _lbfgsb

FAIL_VALUE = 1e300
ALLOWED_FAILURES = 10

# ``setulb`` task codes (scipy/optimize/_lbfgsb_py.py): 3 = evaluate f and g, 1 = new iteration, 4 = converged, 5 = stop
_TASK_FG, _TASK_NEW_X, _TASK_CONVERGED, _TASK_STOP = 3, 1, 4, 5
_INT = np.int64 if getattr(_lbfgsb, "HAS_ILP64", False) else np.int32
try:  # newer SciPy exposes the flag next to the driver
    from scipy.optimize._lbfgsb_py import HAS_ILP64 as _ILP64
    _INT = np.int64 if _ILP64 else np.int32
except Exception:  # pragma: no cover
    pass


class _Instance:
    """State of one ``_minimize_lbfgsb`` call (same arrays, same loop, see scipy/optimize/_lbfgsb_py.py)."""

    def __init__(self, x0, m, factr, pgtol, maxfun, maxiter, maxls):
        n = x0.shape[0]
        self.m, self.factr, self.pgtol, self.maxfun, self.maxiter, self.maxls = m, factr, pgtol, maxfun, maxiter, maxls
        self.x = np.array(x0, dtype=np.float64)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros(n, dtype=np.float64)
        self.low = np.zeros(n)
        self.up = np.zeros(n)
        self.nbd = np.zeros(n, dtype=_INT)          # unbounded: the transforms live in the objective
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self.iwa = np.zeros(3 * n, dtype=_INT)
        self.task = np.zeros(2, dtype=_INT)
        self.ln_task = np.zeros(2, dtype=_INT)
        self.lsave = np.zeros(4, dtype=_INT)
        self.isave = np.zeros(44, dtype=_INT)
        self.dsave = np.zeros(29, dtype=np.float64)
        self.nit = 0
        self.nfev = 0
        self.last_x = None          # ScalarFunction's cache: f, g are re-used when x has not moved
        self.done = False
        self.error = None

    def set_value(self, x, f, g):
        self.last_x = np.array(x, copy=True)
        self.f = f
        self.g = np.asarray(g, dtype=np.float64)
        self.nfev += 1

    def advance(self):
        """Runs the SciPy loop until a *new* objective value is needed (returns the point) or the run ends (None)."""
        while True:
            self.g = self.g.astype(np.float64)
            _lbfgsb.setulb(self.m, self.x, self.low, self.up, self.nbd, self.f, self.g, self.factr, self.pgtol, self.wa,
                           self.iwa, self.task, self.lsave, self.isave, self.dsave, self.maxls, self.ln_task)
            t = self.task[0]
            if t == _TASK_FG:
                if self.last_x is not None and np.array_equal(self.x, self.last_x):
                    continue                      # cached (the start point was evaluated before the first call)
                return self.x
            if t == _TASK_NEW_X:
                self.nit += 1
                if self.nit >= self.maxiter:
                    self.task[0], self.task[1] = _TASK_STOP, 504
                elif self.nfev > self.maxfun:
                    self.task[0], self.task[1] = _TASK_STOP, 502
                continue
            self.done = True
            return None

    def result(self):
        if self.task[0] == _TASK_CONVERGED:
            warnflag = 0
        elif self.nfev > self.maxfun or self.nit >= self.maxiter:
            warnflag = 1
        else:
            warnflag = 2
        info = {"grad": self.g, "funcalls": self.nfev, "nit": self.nit, "warnflag": warnflag,
                "task": (int(self.task[0]), int(self.task[1]))}
        return self.x, float(self.f), info


def lbfgsb_lockstep(eval_batch: Callable[[List[int], List[np.ndarray]], List[Tuple[float, np.ndarray]]],
                    x0s: Sequence[np.ndarray], max_evals: int = 1000, m: int = 10, factr: float = 1e7,
                    pgtol: float = 1e-5, maxiter: int = 15000, maxls: int = 20):
    """Minimise len(x0s) independent objectives; ``eval_batch(indices, xs) -> [(f, g) | Exception, ...]`` is called
    with every currently outstanding request.  Returns [(x, f, info) | Exception] in input order -- the same
    trajectories, values and ``funcalls`` as ``fmin_l_bfgs_b(..., maxfun=max_evals)`` run once per objective."""
    n = len(x0s)
    inst = [_Instance(np.asarray(x0, dtype=np.float64).ravel(), m, factr, pgtol, max_evals, maxiter, maxls) for x0 in x0s]
    out = [None] * n
    # ScalarFunction evaluates the start point at construction
    pending = {i: inst[i].x for i in range(n)}
    while pending:
        idx = sorted(pending)
        res = eval_batch(idx, [np.array(pending[i], copy=True) for i in idx])
        nxt = {}
        # setulb factorises (2m x 2m) matrices through LAPACK: a multi-threaded BLAS spends ~0.3 ms per call spinning
        # its pool up for them (measured: 28 ms per round of 90 models), so the host BLAS is pinned to one thread here
        with _single_threaded_blas():
            for i, r in zip(idx, res):
                if isinstance(r, BaseException):
                    out[i] = r
                    continue
                inst[i].set_value(pending[i], r[0], r[1])
                x = inst[i].advance()
                if x is None:
                    out[i] = inst[i].result()
                else:
                    nxt[i] = x
        pending = nxt
    return out


def _single_threaded_blas():
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=1, user_api="blas")
    except Exception:  # pragma: no cover - threadpoolctl is optional
        import contextlib
        return contextlib.nullcontext()


_KIND = {"free": 0, "positive": 1, "bounded": 2}


def device_optimizer_default(models) -> bool:
    """The on-device optimiser is the default wherever it applies (EP inference, an engine that has it); AGPC_DEVICE_OPT=0
    keeps every ``optimize_models`` call on the SciPy driver (bit-for-bit SciPy trajectories), =1 insists on the device."""
    env = os.environ.get("AGPC_DEVICE_OPT", "")
    if env == "0":
        return False
    able = hasattr(models[0].engine, "optimize_batch") and all(m.ep.get("inference", "ep") == "ep" for m in models)
    if env not in ("", "0") and not able and hasattr(models[0].engine, "optimize_batch"):
        raise ValueError("AGPC_DEVICE_OPT=1: the on-device optimiser runs EP inference only")
    return able


def optimize_models_device(models, max_evals: int = 1000):
    """``optimize_models`` with the L-BFGS state on the device (``agpc_optimize_batch``, csrc/optim.cu): the iterates, the
    limited-memory pairs, the line searches and the EP sites of all models stay in HBM; per round the host only reads
    the "still running" flags.  Same algorithm and stopping rules as the SciPy driver below, not the same rounding."""
    engine = models[0].engine
    by_ds = {}
    for mdl in models:
        if mdl.ep.get("inference", "ep") != "ep":
            raise ValueError("the on-device optimiser runs EP inference only")
        by_ds.setdefault(mdl._dataset, []).append(mdl)
    for ds, group in by_ds.items():
        params = [m.kern.flat_params() for m in group]
        ep = {k: v for k, v in group[0].ep.items() if k != "inference"}
        r = engine.optimize_batch(
            ds, [m._program for m in group], [m.kern.param_array for m in group],
            [[_KIND[p.kind] for p in ps] for ps in params],
            [[0.0 if p.lo is None else p.lo for p in ps] for ps in params],
            [[1.0 if p.hi is None else p.hi for p in ps] for ps in params],
            [m._rows for m in group], max_evals=max_evals, **ep)
        for b, mdl in enumerate(group):
            mdl.kern.set_param_array(r["theta"][b])
            mdl._absorb(r, b)
            task = int(r["task"][b])
            warnflag = 0 if task in (1, 2) else 1 if task in (3, 4) else 2
            mdl.optimization_info = dict(funcalls=int(r["evals"][b]), evals=int(r["evals"][b]), nit=int(r["iters"][b]),
                                         warnflag=warnflag, task=task, device=True)
            if task == 6:
                mdl.optimization_error = np.linalg.LinAlgError("too many failed evaluations")


def optimize_models(models, max_evals: int = 1000, warm_start: bool = True, device: bool = None):
    """Optimise GPClassification shims in lock-step.  Each evaluation of every live model is one item of one
    ``score_batch``; EP is warm-started from the model's previous sites.  Models must share an engine; they may
    live on different datasets only if the engine call is split, so they are grouped by dataset id.
    By default the optimiser state lives on the GPU (``optimize_models_device``, csrc/optim.cu); ``device=False`` or
    AGPC_DEVICE_OPT=0 selects this host driver, which steps SciPy's own L-BFGS-B."""
    if not models:
        return
    if (device_optimizer_default(models) if device is None else device) and warm_start:   # (the device loop always warm-starts EP)
        return optimize_models_device(models, max_evals=max_evals)
    engine = models[0].engine
    fails = [0] * len(models)
    evals = [0] * len(models)
    last_grad = [None] * len(models)   # transformed gradient of the last successful evaluation (what paramz reports on a failure)
    # L-BFGS-B returns an iterate it has already evaluated: the last few successful evaluations of every model are kept
    # (keyed by the bytes of x) so that leaving the model "at its optimum" does not cost one more engine evaluation
    recent = [dict() for _ in models]
    KEEP = 4

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
                    if fails[i] > ALLOWED_FAILURES:   # paramz gives up on THIS model; the others carry on
                        res[i] = np.linalg.LinAlgError("too many failed evaluations")
                        continue
                    mdl.status = st
                    # paramz: objective = a huge value, gradient = the model's CURRENT gradient, i.e. the one left by
                    # the last successful parameters_changed() -- not zero, which L-BFGS-B would read as convergence
                    res[i] = (FAIL_VALUE, np.zeros_like(x) if last_grad[i] is None else last_grad[i].copy())
                    continue
                fails[i] = 0   # paramz counts CONSECUTIVE failures: _fail_count is reset by every successful evaluation
                mdl._absorb(r, b)
                gf = np.array([p.gradfactor() for p in params[b]])
                last_grad[i] = np.clip(mdl._grad * gf, -1e10, 1e10)           # paramz clips the transformed gradient
                res[i] = (mdl._nlml, last_grad[i])
                rec = recent[i]
                rec[np.asarray(x, dtype=np.float64).tobytes()] = (mdl._nlml, mdl._grad, mdl.bic, mdl.status,
                                                                  mdl.ep_sweeps, mdl._tau, mdl._nu)
                while len(rec) > KEEP:
                    rec.pop(next(iter(rec)))
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
        hit = recent[i].get(np.asarray(x, dtype=np.float64).tobytes())
        if hit is not None:
            mdl._nlml, mdl._grad, mdl.bic, mdl.status, mdl.ep_sweeps, mdl._tau, mdl._nu = hit
        else:
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
