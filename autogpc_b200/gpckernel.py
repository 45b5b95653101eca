"""``GPCKernel``: one node of the AutoGPC search tree = one candidate kernel to be scored (host side).

Python 3 mirror of /root/reference/gpckernel.py with GPy replaced by ``autogpc_b200.gpy_compat`` (CUDA engine):
same class, method and helper names, argument meaning and error behaviour (``train`` skips a failing fold and
raises ``RuntimeError`` only if every fold failed, gpckernel.py:196-215; the "median" model is the one at index
``len(results) // 2`` in *fold order* because the reference discards its sort, gpckernel.py:209-211).

Differences, all called for by SURVEY 7 / BASELINE north_star:
  * ``train(mode='auto')`` resolves to 'full' -- the SVGP branch (gpckernel.py:255-293) is out of scope;
  * random restarts draw from an explicit ``np.random.Generator`` (``set_rng``), not the global unseeded RNG;
  * the 5 folds of one candidate -- and, through ``train_kernels``, of *all* candidates of a search depth -- are
    optimised in lock-step, every L-BFGS-B evaluation of every fold being one item of one ``score_batch`` call;
  * ``LinearKernel`` translates to ``gpy_compat.kern.Linear`` instead of raising (gpckernel.py:573), BIC is exposed
    (``GPCKernel.bic``), and the plotting import is lazy (gpckernel.py:10).
"""
from __future__ import annotations

import sys
import warnings
from typing import List, Optional

import numpy as np

from . import flexible_function as ff
from . import gpy_compat as GPy
from . import grammar
from . import optim
from .engine import EngineError as _EngineError
from .gpcdata import GPCData

_rng = np.random.default_rng(0)


def set_rng(rng) -> None:
    """Seed the restart distribution (resetGpssParams) and kernel initialisation."""
    global _rng
    _rng = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)
    ff.set_rng(_rng)


class GPCKernel:
    """GP classification kernel for AutoGPC: wrapper around a GPSS kernel + its trained model (gpckernel.py:14-57)."""

    def __init__(self, gpssKernel, data, depth=0, parent=None):
        self.kernel = gpssKernel
        self.data = data
        self.depth = depth
        self.parent = parent
        self.model = None
        self.isSparse = None
        self.errorRate = None
        resetGpssParams(self.kernel, data=self.data)

    def __repr__(self):
        return "GPCKernel: depth = %d, NLML = %f, CV error = %.4f\n  %s" % \
               (self.depth, self.getNLML(), self.error(), self.kernel.pretty_print())

    def equals(self, other, strict=False, canonical=True):
        return isKernelEqual(self.kernel, other.kernel, compare_params=strict, use_canonical=canonical)

    def betterThan(self, other, strict=False):
        """Lower error, then lower NLML; rounded to 4 / 2 decimals unless ``strict`` (gpckernel.py:81-102)."""
        e0, e1 = self.error(), other.error()
        l0, l1 = self.getNLML(), other.getNLML()
        if not strict:
            e0, e1 = round(e0, 4), round(e1, 4)
            l0, l1 = round(l0, 2), round(l1, 2)
        if e0 != e1:
            return e0 < e1
        return l0 < l1

    def _untrained_model(self, k):
        """GPClassification on the full data at the kernel's current parameters: one EP inference, no optimize()."""
        did = device_dataset(self.data)
        return GPy.models.GPClassification(self.data.X, self.data.Y, kernel=gpss2gpy(k, data=self.data),
                                           _dataset=did, _rows=np.arange(self.data.getNum(), dtype=np.int32))

    def add(self, other):
        """Sum kernel with an un-optimised full-data model (gpckernel.py:105-128)."""
        k = self.kernel + other.kernel
        ret = GPCKernel(k, self.data)
        ret.isSparse = False
        ret.model = self._untrained_model(k)
        return ret

    def add_many(self, others):
        """``[self.add(o) for o in others]`` with every un-optimised full-data inference in ONE engine batch and every
        training-error prediction in one ``predict_batch`` (the loop of gpcreport.py:672-674 asks for exactly this)."""
        rets = []
        did = device_dataset(self.data)
        rows = np.arange(self.data.getNum(), dtype=np.int32)
        for o in others:
            k = self.kernel + o.kernel
            ret = GPCKernel(k, self.data)
            ret.isSparse = False
            ret.model = GPy.models.GPClassification(self.data.X, self.data.Y, kernel=gpss2gpy(k, data=self.data),
                                                    _dataset=did, _rows=rows, _infer=False)
            rets.append(ret)
        infer_models_with_errors(rets)
        return rets

    def expand(self, base_kernels="SE"):
        """All single-production rewrites, canonical + simplified + de-duplicated (gpckernel.py:131-147)."""
        ndim = self.data.getDim()
        g = grammar.MultiDGrammar(ndim, base_kernels=base_kernels, rules=None)
        kernels = grammar.expand(self.kernel, g)
        kernels = ff.remove_duplicates([k.canonical() for k in kernels])
        for k in kernels:
            k.initialise_params(data_shape=self.data.getDataShape())
        kernels = ff.remove_duplicates([k.simplified() for k in kernels])
        kernels = [k for k in kernels if not isinstance(k, ff.NoneKernel)]
        return [GPCKernel(k, self.data, depth=self.depth + 1, parent=self) for k in kernels]

    def reset(self):
        resetGpssParams(self.kernel, data=self.data)
        self.isSparse = None
        self.errorRate = None

    def train(self, mode="auto", n_folds=5, max_evals=1000):
        """k-fold cross-validated training with one random restart per fold (gpckernel.py:162-216)."""
        train_kernels([self], mode=mode, n_folds=n_folds, max_evals=max_evals)

    def trainFull(self, X=None, Y=None, XT=None, YT=None, randomise=False, rows=None, test_rows=None, max_evals=1000):
        """Train one full GP classifier and compute its validation error; does not mutate ``self``
        (gpckernel.py:219-252).  Folds may be given as arrays (reference signature) or as dataset row lists."""
        k = self.kernel.copy()
        if randomise:
            resetGpssParams(k, data=self.data)
        if rows is not None:
            m = GPy.models.GPClassification(self.data.X, self.data.Y, kernel=gpss2gpy(k, data=self.data),
                                            _dataset=device_dataset(self.data), _rows=rows)
        else:
            m = GPy.models.GPClassification(X, Y, kernel=gpss2gpy(k, data=self.data))
        m.optimize(max_iters=max_evals)
        if test_rows is not None:
            cverror = _error_on_rows(m, self.data, test_rows)
        else:
            if XT is None or YT is None:
                XT, YT = X, Y
            cverror = computeError(m, XT, YT)
        return {"model": m, "error": cverror}

    def trainSVGP(self, X, Y, XT=None, YT=None, randomise=False, inducing=10):
        """Sparse variational GP branch (gpckernel.py:255-293): outside the hot path this engine covers (north_star
        names full EP / Laplace inference only); present so that the call fails with an explanation."""
        raise NotImplementedError("SVGP (gpckernel.py:255-293) is out of scope of the B200 scoring engine; "
                                  "use train(mode='full') / trainFull")

    def toSummands(self):
        """Additive components, each with an un-optimised full-data model (gpckernel.py:296-324).  As in the
        reference, constructing the component ``GPCKernel`` re-draws its hyper-parameters (``__init__`` calls
        ``resetGpssParams``), so the component models sit at restart values, not at the parent's optimum."""
        k = self.kernel.additive_form()
        parts = k.operands if isinstance(k, ff.SumKernel) else [k]
        summands = [GPCKernel(o, self.data) for o in parts]
        for s in summands:
            s.isSparse = False
            s.model = self._untrained_model(s.kernel)
        return summands

    def draw(self, filename, active_dims_only=False, draw_posterior=True):
        raise NotImplementedError("plotting (gpcplot.py) is out of scope of the B200 scoring engine")

    def misclassifiedPoints(self, X=None, Y=None):
        if X is None or Y is None:
            X, Y = self.data.X, self.data.Y
        return misclassifiedPoints(self.model, X, Y)

    def getDepth(self):
        return self.depth

    def getNLML(self):
        return float("inf") if self.model is None else -self.model.log_likelihood()

    def bic(self):
        """BIC = 2 NLML + p_eff ln n of the selected model (flexible_function.py:685-687; north_star's criterion)."""
        return float("inf") if self.model is None else self.model.bic

    def getActiveDims(self):
        if self.model is not None:
            return list(self.model.kern.active_dims)
        return list(gpss2gpy(self.kernel, data=self.data).active_dims)

    def getGPyKernel(self):
        return gpss2gpy(self.kernel, data=self.data)

    def error(self):
        """Cached (cross-validated) error; constant kernels score the majority-class rate (gpckernel.py:393-407)."""
        if isinstance(self.kernel, ff.ConstKernel):
            d = self.data
            return min(d.getClass(0).shape[0], d.getClass(1).shape[0]) / float(d.getNum())
        if self.errorRate is None:
            if self.model is None:
                return 1.0 if isinstance(self.kernel, ff.NoneKernel) else float("inf")
            self.errorRate = computeError(self.model, self.data.X, self.data.Y)
        return self.errorRate

    def monotonicity(self, margin=0.15):
        """+1 / -1 / 0 for a 1-D kernel with increasing / decreasing / non-monotonic posterior mean
        (gpckernel.py:410-439)."""
        assert len(self.getActiveDims()) == 1, "Kernel must be one-dimensional"
        margin = min(max(margin, 0.0), 0.5)
        dim = self.getActiveDims()[0]
        x = self.data.X[:, dim]
        xlo, xhi = x.min() + margin * (x.max() - x.min()), x.max() - margin * (x.max() - x.min())
        X = self.data.X[(x >= xlo) & (x <= xhi)]
        X = X[X[:, dim].argsort()]
        dmu_dx, _ = self.model.predictive_gradients(X)
        d = dmu_dx[:, dim, 0]
        return 1 if np.all(d > 0) else (-1 if np.all(d < 0) else 0)

    def period(self):
        assert len(self.getActiveDims()) == 1, "Kernel must be one-dimensional"
        return self.kernel.period if isinstance(self.kernel, ff.PeriodicKernel) else 0.0

    def sensitivity(self):
        sdict = sensitivityDict(self.kernel)
        return np.array([sdict.get(d, 0.0) for d in self.getActiveDims()])

    def shortInterp(self):
        names = {ff.SqExpKernel: "smooth", ff.PeriodicKernel: "periodic", ff.ConstKernel: "constant",
                 ff.LinearKernel: "linear", ff.SumKernel: "additive", ff.ProductKernel: "interaction",
                 ff.NoneKernel: "null"}
        for cls, s in names.items():
            if isinstance(self.kernel, cls):
                return s
        raise NotImplementedError("Unrecognised kernel type.")

    def latex(self):
        return gpss2latex(self.kernel)


# ---- batched training of many candidates ---------------------------------------------------------------------
def device_dataset(data: GPCData, engine=None) -> int:
    """Uploads ``data`` once per engine and caches the dataset id on the GPCData object."""
    from . import engine as _eng
    engine = engine if engine is not None else _eng.default_engine()
    cache = data.__dict__.setdefault("_device_ids", {})
    key = id(engine)
    if key not in cache:
        cache[key] = engine.dataset(data.X, data.Y)
    return cache[key]


def _error_on_rows(model, data: GPCData, test_rows) -> float:
    p = model.predict_rows(test_rows)
    y = data.Y[np.asarray(test_rows), 0]
    return float(np.mean((p - 0.5) * (y - 0.5) < 0))


def infer_models_with_errors(kernels: List[GPCKernel]) -> None:
    """One EP inference (no optimisation, no gradient) for the models of ``kernels`` in a single ``score_batch`` and
    their training-set error rates (``computeError(model, data.X, data.Y)``, gpckernel.py:393-407) in a single
    ``predict_batch``; fills ``model`` state and ``errorRate``.  Models must share one device dataset."""
    ms_ = [k.model for k in kernels]
    if not ms_:
        return
    eng, did = ms_[0].engine, ms_[0]._dataset
    r = eng.score_batch(did, [m._program for m in ms_], [m.kern.param_array for m in ms_], [m._rows for m in ms_],
                        want_grad=False, **ms_[0].ep)
    for b, m in enumerate(ms_):
        m._absorb(r, b)
    ok = [i for i, m in enumerate(ms_) if m.status not in (2, 4) and np.isfinite(m._nlml)]
    if ok:
        p, _ = eng.predict_batch(did, [ms_[i]._program for i in ok], [ms_[i].kern.param_array for i in ok],
                                 [ms_[i]._rows for i in ok], [ms_[i]._tau for i in ok], [ms_[i]._nu for i in ok],
                                 [ms_[i]._rows for i in ok])
        for i, pi in zip(ok, p):
            y = kernels[i].data.Y[ms_[i]._rows, 0]
            kernels[i].errorRate = float(np.mean((pi - 0.5) * (y - 0.5) < 0))
    for i, k in enumerate(kernels):
        if i not in ok:
            k.errorRate = float("inf")


def cumulateAdditiveKernels(summands: List[GPCKernel]):
    """Incrementally cumulate the additive components of a kernel, best (lowest training error) first
    (gpcreport.py:657-680; the report layer itself is out of scope, this is its one caller of the scoring engine).
    Returns ``(kers, cums)`` like the reference.  Each round scores ``cum + k`` for every remaining ``k`` in one batch
    instead of one GPy model per candidate; ordering and tie-breaking follow the reference's stable sorts."""
    terms = sorted(summands, key=lambda k: k.error(), reverse=True)
    ker = terms.pop()
    cum = ker
    kers, cums = [ker], [cum]
    while len(terms) > 0:
        added = cum.add_many(terms)
        order = sorted(range(len(terms)), key=lambda i: added[i].error(), reverse=True)
        terms = [terms[i] for i in order]
        ker = terms.pop()
        cum = cum.add(ker)   # as in the reference: the winner is added again (its GPCKernel re-draws the parameters)
        kers.append(ker)
        cums.append(cum)
    return kers, cums


def train_kernels(kernels: List[GPCKernel], mode="auto", n_folds=5, max_evals=1000, shard=None):
    """``for k in kernels: k.train()`` (gpcsearch.py:44-45) as ONE lock-step optimisation over every
    (candidate, fold) item.  With ``shard=(rank, world)`` this rank optimises only its share of the items -- dealt by
    estimated cost, longest first (``sharding.deal_items``) -- and results are exchanged with one all-gather
    (autogpc_b200/sharding.py)."""
    mode = mode.lower()
    assert mode in ("full", "svgp", "auto"), "mode must be 'full', 'svgp' or 'auto'"
    assert n_folds is None or (isinstance(n_folds, int) and n_folds > 0), "n_folds must be None or positive integer"
    if mode == "svgp":
        raise NotImplementedError("SVGP (gpckernel.py:255-293) is out of scope; use mode='full'")
    if not kernels:
        return
    randomise = n_folds is not None
    nf = 1 if n_folds is None else n_folds
    items = []  # (kernel index, fold index, model)
    over_limit = set()
    for ci, gk in enumerate(kernels):
        gk.isSparse = False
        folds = gk.data.kFoldIndices(k=nf)
        did = device_dataset(gk.data)
        try:
            GPy.compile_program(gpss2gpy(gk.kernel, data=gk.data))
        except _EngineError as e:
            # a structure beyond the engine's program limits (16 leaves / 16 product terms / 48 hyper-parameters after
            # distributing products over sums) cannot be scored: it ranks last (error = BIC = inf) instead of aborting
            # the search.  GPy has no such limit; the reference's grammar reaches it only beyond depth ~5.
            warnings.warn("candidate %s skipped: %s" % (gk.kernel, e))
            over_limit.add(ci)
            gk.model, gk.errorRate = None, None
            continue
        for fi, (tr, te) in enumerate(folds):
            k = gk.kernel.copy()
            if randomise:
                resetGpssParams(k, data=gk.data)
            m = GPy.models.GPClassification(gk.data.X, gk.data.Y, kernel=gpss2gpy(k, data=gk.data), _dataset=did,
                                            _rows=tr, _infer=False)
            m._test_rows = te
            items.append((ci, fi, m))

    if shard is None:
        mine = list(range(len(items)))
    else:
        from . import sharding
        rank, world = shard
        costs = []
        for _, _, m in items:
            prog = m._program
            types = [prog.leaf_type[l] for l in range(prog.n_leaves)]
            costs.append(sharding.item_cost(len(m._rows), prog.n_leaves, sum(1 for t in types if t == 1)))
        mine = sharding.shard_indices(len(items), rank, world, costs)
    my_models = [items[i][2] for i in mine]
    if my_models:
        optim.optimize_models(my_models, max_evals=max_evals)
    errs = {}
    ok = []
    for i in mine:
        m = items[i][2]
        if getattr(m, "optimization_error", None) is not None or not np.isfinite(m._nlml):
            errs[i] = float("nan")   # failed fold: skipped below, like the reference's bare except
        else:
            ok.append(i)
    # held-out error of every surviving fold model (computeError, gpckernel.py:837-846): one predict_batch per dataset
    by_ds = {}
    for i in ok:
        by_ds.setdefault((id(items[i][2].engine), items[i][2]._dataset), []).append(i)
    for group in by_ds.values():
        ms_ = [items[i][2] for i in group]
        for m in ms_:
            m._ensure_sites()
        p, _ = ms_[0].engine.predict_batch(ms_[0]._dataset, [m._program for m in ms_], [m.kern.param_array for m in ms_],
                                           [m._rows for m in ms_], [m._tau for m in ms_], [m._nu for m in ms_],
                                           [np.asarray(m._test_rows, dtype=np.int32) for m in ms_])
        for i, m, pi in zip(group, ms_, p):
            yte = kernels[items[i][0]].data.Y[np.asarray(m._test_rows), 0]
            errs[i] = float(np.mean((pi - 0.5) * (yte - 0.5) < 0))
    if shard is not None:
        from . import sharding
        sharding.exchange_models(items, mine, errs, shard)

    for ci, gk in enumerate(kernels):
        if ci in over_limit:
            continue
        results = [{"model": items[i][2], "error": errs[i]} for i in range(len(items))
                   if items[i][0] == ci and np.isfinite(errs.get(i, float("nan")))]
        if not results:
            raise RuntimeError("Error during training: none of the %d optimisation attempts were successful." % nf)
        med = len(results) // 2
        gk.model = results[med]["model"]
        gk.kernel = gpy2gpss(gk.model.kern)
        gk.errorRate = float(np.mean([r["error"] for r in results]))
        gk.foldErrors = [r["error"] for r in results]
        gk.foldModels = [r["model"] for r in results]


# ---- helper functions (same names as the reference's module-level helpers) ------------------------------------
def gpss2gpy(kernel, data=None):
    """GPSS kernel -> GPy-shaped kernel with constraints (gpckernel.py:514-574): variance = sf**2, 1-D leaves on
    ``active_dims=[dimension]``, lengthscale / period bounded to [min separation, 2 x range]."""
    assert isinstance(kernel, ff.Kernel), "kernel must be of type flexible_function.Kernel"
    if isinstance(kernel, ff.SqExpKernel):
        k = GPy.kern.RBF(1, variance=kernel.sf ** 2, lengthscale=kernel.lengthscale, active_dims=np.array([kernel.dimension]))
        if data:
            lo, hi = _bounds(data, kernel.dimension)
            k["lengthscale"].constrain_bounded(lo, hi, warning=False)
        return k
    if isinstance(kernel, ff.PeriodicKernel):
        k = GPy.kern.StdPeriodic(1, variance=kernel.sf ** 2, period=kernel.period, lengthscale=kernel.lengthscale,
                                 active_dims=np.array([kernel.dimension]))
        if data:
            lo, hi = _bounds(data, kernel.dimension)
            k["lengthscale"].constrain_bounded(lo, hi, warning=False)
            lo, hi = _bounds(data, kernel.dimension, period=True)
            k["period"].constrain_bounded(lo, hi, warning=False)
        return k
    if isinstance(kernel, ff.LinearKernel):
        return GPy.kern.Linear(1, variance=kernel.sf ** 2, location=kernel.location, active_dims=np.array([kernel.dimension]))
    if isinstance(kernel, ff.ConstKernel):
        assert isinstance(data, GPCData), "Must specify data field for ConstKernel"
        ndim = data.getDim()
        return GPy.kern.Bias(ndim, variance=kernel.sf ** 2, active_dims=np.array(range(ndim)))
    if isinstance(kernel, ff.SumKernel):
        return GPy.kern.Add([gpss2gpy(o, data=data) for o in kernel.operands])
    if isinstance(kernel, ff.ProductKernel):
        return GPy.kern.Prod([gpss2gpy(o, data=data) for o in kernel.operands])
    raise NotImplementedError("Cannot translate kernel of type " + type(kernel).__name__)


def gpy2gpss(kernel):
    """Back-translation of optimised hyper-parameters (gpckernel.py:577-617)."""
    assert isinstance(kernel, GPy.kern.Kern), "kernel must be of type GPy.kern.Kern"
    if isinstance(kernel, GPy.kern.RBF):
        return ff.SqExpKernel(dimension=int(kernel.active_dims[0]), lengthscale=kernel.lengthscale[0],
                              sf=np.sqrt(kernel.variance[0]))
    if isinstance(kernel, GPy.kern.StdPeriodic):
        return ff.PeriodicKernel(dimension=int(kernel.active_dims[0]), lengthscale=kernel.lengthscale[0],
                                 period=kernel.period[0], sf=np.sqrt(kernel.variance[0]))
    if isinstance(kernel, GPy.kern.Linear):
        return ff.LinearKernel(dimension=int(kernel.active_dims[0]), location=kernel.location[0],
                               sf=np.sqrt(kernel.variance[0]))
    if isinstance(kernel, GPy.kern.Bias):
        return ff.ConstKernel(sf=np.sqrt(kernel.variance[0]))
    if isinstance(kernel, GPy.kern.Add):
        return ff.SumKernel([gpy2gpss(p) for p in kernel.parts])
    if isinstance(kernel, GPy.kern.Prod):
        return ff.ProductKernel([gpy2gpss(p) for p in kernel.parts])
    raise NotImplementedError("Cannot translate kernel of type " + type(kernel).__name__)


def _bounds(data, dim, period=False):
    b = data.getPeriodBounds(dims=dim) if period else data.getLengthscaleBounds(dims=dim)
    b = np.asarray(b, float).reshape(-1)
    return float(b[0]), float(b[1])


def _log_uniform(lo, hi):
    return _rng.random() * (hi - lo) + lo


def _draw_sf(y_sd, p_data, sd):
    loc = np.log(y_sd) if _rng.random() < p_data else 0.0
    return float(np.exp(_rng.normal(loc=loc, scale=sd)))


def resetGpssParams(k, data=None, sd=1):
    """Random-restart distribution, linear-scale outputs (gpckernel.py:620-689).  The sequence of draws per leaf
    follows the reference so a seeded generator reproduces a restart."""
    assert isinstance(k, ff.Kernel), "kernel must be of type flexible_function.Kernel"
    shape = data.getDataShape()
    if isinstance(k, ff.NoneKernel):
        return
    if isinstance(k, ff.ConstKernel):
        k.sf = _draw_sf(shape["y_sd"], 2.0 / 3, sd)
    elif isinstance(k, ff.SqExpKernel):
        lo, hi = np.log(_bounds(data, k.dimension))
        if _rng.random() < 2.0 / 3:
            ls = _rng.normal(loc=np.log(shape["x_sd"][k.dimension]), scale=sd)
        else:
            ls = _log_uniform(lo, hi)
        if ls < lo or ls > hi:
            ls = _log_uniform(lo, hi)
        k.lengthscale = float(np.exp(ls))
        k.sf = _draw_sf(shape["y_sd"], 0.5, sd)
    elif isinstance(k, ff.PeriodicKernel):
        lo, hi = np.log(_bounds(data, k.dimension))
        ls = _rng.normal(loc=0, scale=sd)
        if ls < lo or ls > hi:
            ls = _log_uniform(lo, hi)
        k.lengthscale = float(np.exp(ls))
        lo, hi = np.log(_bounds(data, k.dimension, period=True))
        if _rng.random() < 1.0 / 3:
            p = _rng.normal(loc=0, scale=sd)
        elif _rng.random() < 1.0 / 2:          # nested draw, as in the reference (:668-673)
            p = _rng.normal(loc=-3.95, scale=sd)
        else:
            p = _rng.normal(loc=-5.9, scale=sd)
        if p < lo or p > hi:
            p = _log_uniform(lo, hi)
        k.period = float(np.exp(p))
        k.sf = _draw_sf(shape["y_sd"], 0.5, sd)
    elif isinstance(k, ff.LinearKernel):
        # no reference behaviour (gpckernel.py:688 raises): follow ff.LinearKernel.initialise_params' location
        # prior (uniform over 3 x data range, flexible_function.py:1326) and the common sf mixture
        d = k.dimension
        k.location = float(_rng.uniform(2 * shape["x_min"][d] - shape["x_max"][d], 2 * shape["x_max"][d] - shape["x_min"][d]))
        k.sf = _draw_sf(shape["y_sd"], 0.5, sd)
    elif isinstance(k, (ff.SumKernel, ff.ProductKernel)):
        for o in k.operands:
            resetGpssParams(o, data=data, sd=sd)
    else:
        raise NotImplementedError("Unrecognised kernel type " + type(k).__name__)


def gpss2latex(k):
    assert isinstance(k, ff.Kernel), "kernel must be of type flexible_function.Kernel"
    if isinstance(k, ff.NoneKernel):
        return r"{\textsc{Null}}"
    if isinstance(k, ff.ConstKernel):
        return r"{\textsc{C}}"
    for cls, name in ((ff.SqExpKernel, "SE"), (ff.PeriodicKernel, "Per"), (ff.LinearKernel, "Lin")):
        if isinstance(k, cls):
            return r"{{\textsc{{{1}}}_{0}}}".format(k.dimension + 1, name)
    if isinstance(k, ff.SumKernel):
        return " + ".join(gpss2latex(o) for o in k.operands)
    if isinstance(k, ff.ProductKernel):
        return r" \times ".join("( " + gpss2latex(o) + " )" if isinstance(o, ff.SumKernel) else gpss2latex(o)
                                for o in k.operands)
    raise NotImplementedError("Unrecognised kernel type " + type(k).__name__)


def removeKernelParams(kernel):
    assert isinstance(kernel, ff.Kernel), "kernel must be of type flexible_function.Kernel"
    if isinstance(kernel, (ff.SqExpKernel, ff.PeriodicKernel, ff.LinearKernel)):
        return type(kernel)(dimension=kernel.dimension)
    if isinstance(kernel, ff.ConstKernel):
        return ff.ConstKernel()
    if isinstance(kernel, (ff.SumKernel, ff.ProductKernel)):
        return type(kernel)([removeKernelParams(o) for o in kernel.operands])
    if isinstance(kernel, ff.NoneKernel):
        return kernel
    raise NotImplementedError("Unrecognised kernel type " + type(kernel).__name__)


def isKernelEqual(k1, k2, compare_params=False, use_canonical=True):
    """Structural (optionally parametric) equality of two GPSS kernels (gpckernel.py:760-817)."""
    assert isinstance(k1, ff.Kernel), "k1 must be of type flexible_function.Kernel"
    assert isinstance(k2, ff.Kernel), "k2 must be of type flexible_function.Kernel"
    if use_canonical:
        k1, k2 = k1.canonical(), k2.canonical()
    if isinstance(k1, ff.NoneKernel):
        return isinstance(k2, ff.NoneKernel)
    if isinstance(k1, ff.ConstKernel):
        return isinstance(k2, ff.ConstKernel) and (not compare_params or np.array_equal(k1.param_vector, k2.param_vector))
    if isinstance(k1, (ff.SqExpKernel, ff.PeriodicKernel, ff.LinearKernel)):
        ok = type(k1) is type(k2) and k1.dimension == k2.dimension
        return ok and (not compare_params or np.array_equal(k1.param_vector, k2.param_vector))
    if isinstance(k1, (ff.SumKernel, ff.ProductKernel)):
        return type(k1) is type(k2) and len(k1.operands) == len(k2.operands) and \
            all(isKernelEqual(a, b, compare_params=compare_params, use_canonical=False)
                for a, b in zip(k1.operands, k2.operands))
    raise NotImplementedError("Cannot compare kernels of type %s and %s" % (type(k1).__name__, type(k2).__name__))


def misclassifiedPoints(model, XT, YT):
    """Samples with (Phi - 0.5)(y - 0.5) < 0 (gpckernel.py:820-834)."""
    if model is None:
        return {"X": XT, "Y": YT}
    Phi, _ = model.predict(XT)
    bad = (((Phi - 0.5) * (YT - 0.5)) < 0).flatten()
    return {"X": XT[bad], "Y": YT[bad]}


def computeError(model, XT, YT):
    """Error rate in [0, 1] (gpckernel.py:837-846)."""
    return misclassifiedPoints(model, XT, YT)["X"].shape[0] / float(XT.shape[0])


def sensitivityDict(k):
    """Per-dimension input sensitivity (sf / lengthscale)^2, summed over sums and multiplied over products
    (gpckernel.py:849-879)."""
    if isinstance(k, (ff.SqExpKernel, ff.PeriodicKernel)):
        return {k.dimension: (k.sf / k.lengthscale) ** 2}
    if isinstance(k, (ff.ConstKernel, ff.NoneKernel, ff.LinearKernel)):
        return {}
    if isinstance(k, (ff.SumKernel, ff.ProductKernel)):
        ret = {}
        for o in k.operands:
            for d, s in sensitivityDict(o).items():
                if d in ret:
                    ret[d] = ret[d] + s if isinstance(k, ff.SumKernel) else ret[d] * s
                else:
                    ret[d] = s
        return ret
    raise NotImplementedError("Unrecognised kernel type " + type(k).__name__)
