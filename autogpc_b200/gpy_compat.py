"""The slice of the GPy object API that AutoGPC touches, re-created on top of the CUDA engine.

The reference has no plugin/FFI boundary; its scoring path *is* these GPy calls (complete list):

  GPy.kern.RBF(1, variance=, lengthscale=, active_dims=)        gpckernel.py:539   -> kern.RBF
  k['lengthscale'].constrain_bounded(lo, hi, warning=False)      gpckernel.py:543   -> Param.constrain_bounded
  GPy.kern.StdPeriodic(1, variance=, period=, lengthscale=, ..)  gpckernel.py:551   -> kern.StdPeriodic
  GPy.kern.Bias(ndim, variance=, active_dims=)                   gpckernel.py:564   -> kern.Bias
  GPy.kern.Add([...]) / GPy.kern.Prod([...]), .parts             gpckernel.py:568,571,610-614
  GPy.models.GPClassification(X, Y, kernel=ker)                  gpckernel.py:126,245,322 (constructor = 1 inference)
  m.optimize() / m.log_likelihood() / m.predict(XT) / m.kern     gpckernel.py:246,371,832,212
  m.predictive_gradients(X), m._raw_predict(X)                   gpckernel.py:431; gpcplot.py:110-114

``kern.Linear`` is an extension (the reference raises NotImplementedError for ff.LinearKernel, gpckernel.py:573).
Usage in host code: ``from autogpc_b200 import gpy_compat as GPy``.  All numerics run in ``libautogpc_b200.so``;
paramz's Logexp / Logistic transforms and SciPy's L-BFGS-B stay on the host exactly as in GPy.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List, Optional, Sequence

import numpy as np

from . import engine as _eng
from . import optim

_LIM = 36.0


# ---- paramz transformations (paramz/transformations.py Logexp, Logistic) ---------------------------------------
def _logexp_f(x):
    return x if x > _LIM else math.log1p(math.exp(max(x, -700.0)))


def _logexp_finv(f):
    return f if f > _LIM else math.log(math.expm1(f))


class Param:
    """One scalar hyper-parameter with its constraint; indexable like a 1-element paramz Param."""

    def __init__(self, name: str, value: float, kind: str = "positive"):
        self.name = name
        self.values = np.array([float(value)])
        self.kind, self.lo, self.hi = kind, None, None

    def __getitem__(self, i):
        return self.values[i]

    def __float__(self):
        return float(self.values[0])

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def __repr__(self):
        return "Param(%s=%r, %s)" % (self.name, float(self.values[0]), self.kind)

    def constrain_bounded(self, lower, upper, warning=True):
        self.kind, self.lo, self.hi = "bounded", float(lower), float(upper)
        v = float(self.values[0])
        if not (self.lo <= v <= self.hi):  # paramz resets an out-of-bounds start value
            self.values[0] = 0.5 * (self.lo + self.hi)
        return self

    def constrain_positive(self, warning=True):
        self.kind = "positive"
        return self

    def unconstrain(self):
        self.kind = "free"
        return self

    # constrained <-> optimiser space
    def to_x(self) -> float:
        v = float(self.values[0])
        if self.kind == "positive":
            return _logexp_finv(v)
        if self.kind == "bounded":
            w = self.hi - self.lo
            v = min(max(v, self.lo + 1e-10 * w), self.hi - 1e-10 * w)
            return math.log((v - self.lo) / (self.hi - v))
        return v

    def from_x(self, x: float) -> float:
        if self.kind == "positive":
            return _logexp_f(x)
        if self.kind == "bounded":
            return self.lo + (self.hi - self.lo) / (1.0 + math.exp(-x)) if x > -700 else self.lo
        return x

    def gradfactor(self) -> float:
        v = float(self.values[0])
        if self.kind == "positive":
            return 1.0 if v > _LIM else -math.expm1(-v)
        if self.kind == "bounded":
            return (v - self.lo) * (self.hi - v) / (self.hi - self.lo)
        return 1.0


class Kern:
    """Base of the kernel shim; ``parameters`` are in GPy's param_array order."""
    parts: Sequence["Kern"] = ()

    def __init__(self, input_dim, active_dims, name):
        self.input_dim = input_dim
        self.active_dims = np.asarray(active_dims, dtype=int).reshape(-1)
        self.name = name
        self.parameters: List[Param] = []

    def __getitem__(self, name: str) -> Param:
        for p in self.parameters:
            if p.name == name:
                return p
        raise KeyError(name)

    def leaves(self) -> List["Kern"]:
        return [self]

    def flat_params(self) -> List[Param]:
        return [p for leaf in self.leaves() for p in leaf.parameters]

    @property
    def param_array(self) -> np.ndarray:
        return np.array([float(p) for p in self.flat_params()])

    def set_param_array(self, theta) -> None:
        for p, v in zip(self.flat_params(), theta):
            p.values[0] = float(v)

    @property
    def effective_params(self) -> int:
        return len(self.parameters)

    def __add__(self, other):
        return Add([self, other])

    def __mul__(self, other):
        return Prod([self, other])

    def copy(self):
        import copy
        return copy.deepcopy(self)


class RBF(Kern):
    leaf_type = _eng.LEAF_SE

    def __init__(self, input_dim=1, variance=1.0, lengthscale=1.0, active_dims=None, name="rbf"):
        super().__init__(input_dim, [0] if active_dims is None else active_dims, name)
        self.variance = Param("variance", variance)
        self.lengthscale = Param("lengthscale", lengthscale)
        self.parameters = [self.variance, self.lengthscale]


class StdPeriodic(Kern):
    leaf_type = _eng.LEAF_PER

    def __init__(self, input_dim=1, variance=1.0, period=1.0, lengthscale=1.0, active_dims=None, name="std_periodic"):
        super().__init__(input_dim, [0] if active_dims is None else active_dims, name)
        self.variance = Param("variance", variance)
        self.period = Param("period", period)
        self.lengthscale = Param("lengthscale", lengthscale)
        self.parameters = [self.variance, self.period, self.lengthscale]


class Bias(Kern):
    leaf_type = _eng.LEAF_CONST

    def __init__(self, input_dim=1, variance=1.0, active_dims=None, name="bias"):
        super().__init__(input_dim, list(range(input_dim)) if active_dims is None else active_dims, name)
        self.variance = Param("variance", variance)
        self.parameters = [self.variance]


class Linear(Kern):
    """k = variance (x - location)(x' - location): ff.LinearKernel (flexible_function.py:1284-1343).  Extension."""
    leaf_type = _eng.LEAF_LIN

    def __init__(self, input_dim=1, variance=1.0, location=0.0, active_dims=None, name="linear"):
        super().__init__(input_dim, [0] if active_dims is None else active_dims, name)
        self.variance = Param("variance", variance)
        self.location = Param("location", location, kind="free")
        self.parameters = [self.variance, self.location]


class _Combination(Kern):
    def __init__(self, parts, name):
        parts = list(parts)
        dims = sorted(set(int(d) for p in parts for d in p.active_dims))
        super().__init__(len(dims), dims, name)
        self.parts = parts

    def leaves(self):
        return [l for p in self.parts for l in p.leaves()]


class Add(_Combination):
    def __init__(self, parts, name="sum"):
        super().__init__(parts, name)

    @property
    def effective_params(self):
        return sum(p.effective_params for p in self.parts)


class Prod(_Combination):
    def __init__(self, parts, name="mul"):
        super().__init__(parts, name)

    @property
    def effective_params(self):  # flexible_function.py:1647: one scale per product
        return sum(p.effective_params for p in self.parts) - (len(self.parts) - 1)


def compile_program(kern: Kern) -> _eng.Program:
    """GPy-shaped tree -> flat engine program.  theta order = depth-first leaf order = GPy's param_array;
    products are distributed over sums (a leaf may then sit in several terms but keeps one theta slot)."""
    leaf_type, leaf_dim, leaf_off = [], [], []
    off = 0

    def walk(k):
        nonlocal off
        if isinstance(k, Add):
            return [t for p in k.parts for t in walk(p)]
        if isinstance(k, Prod):
            acc = [[]]
            for p in k.parts:
                tp = walk(p)            # once per factor: walking registers the factor's leaves
                acc = [a + t for a in acc for t in tp]
            return acc
        idx = len(leaf_type)
        leaf_type.append(k.leaf_type)
        leaf_dim.append(int(k.active_dims[0]) if k.leaf_type != _eng.LEAF_CONST else 0)
        leaf_off.append(off)
        off += len(k.parameters)
        return [[idx]]

    terms = walk(kern)
    return _eng.Program.build(leaf_type, leaf_dim, leaf_off, terms, off, kern.effective_params)


kern = SimpleNamespace(Kern=Kern, RBF=RBF, StdPeriodic=StdPeriodic, Bias=Bias, Linear=Linear, Add=Add, Prod=Prod)


# ---- model ---------------------------------------------------------------------------------------------------
class _ProbitLink:
    @staticmethod
    def transf(f):
        from math import erfc, sqrt
        return 0.5 * np.vectorize(erfc)(-np.asarray(f, float) / sqrt(2.0))


class GPClassification:
    """``GPy.models.GPClassification(X, Y, kernel)``: Bernoulli/probit likelihood, EP inference -- on the engine.

    Construction runs one inference at the kernel's current hyper-parameters, exactly like GPy's
    ``Model.initialize_parameter`` -> ``parameters_changed``.  Internally the model is (dataset id, row list, kernel,
    EP sites); the row-list form lets the search score many folds of one device-resident dataset in one batch.
    """

    def __init__(self, X=None, Y=None, kernel: Optional[Kern] = None, engine: Optional[_eng.Engine] = None,
                 _dataset: Optional[int] = None, _rows: Optional[np.ndarray] = None, _infer: bool = True,
                 ep_epsilon: float = 1e-6, ep_damping: float = 0.0, ep_max_sweeps: int = 100,
                 inference_method=None):
        self.engine = engine if engine is not None else _eng.default_engine()
        self.ep = dict(epsilon=ep_epsilon, damping=ep_damping, max_sweeps=ep_max_sweeps)
        # GPy: GPClassification(..., inference_method=GPy.inference.latent_function_inference.Laplace()); default EP
        if inference_method is not None and getattr(inference_method, "name", str(inference_method)).lower() == "laplace":
            self.ep = dict(epsilon=1e-10, damping=0.0, max_sweeps=50, inference="laplace")
        self._owns_dataset = _dataset is None
        if _dataset is None:
            X = np.asarray(X, float)
            Y = np.asarray(Y, float).reshape(-1, 1)
            self.X, self.Y = X, Y
            self._dataset = self.engine.dataset(X, Y)
            self._rows = np.arange(X.shape[0], dtype=np.int32)
        else:
            self._dataset, self._rows = _dataset, np.asarray(_rows, dtype=np.int32)
            self.X = None if X is None else np.asarray(X, float)[self._rows]
            self.Y = None if Y is None else np.asarray(Y, float).reshape(-1, 1)[self._rows]
        self.kern = kernel if kernel is not None else RBF(self.input_dim if self.X is not None else 1)
        self.likelihood = SimpleNamespace(gp_link=_ProbitLink())
        self._program = compile_program(self.kern)
        self._tau = self._nu = None
        self._nlml = float("inf")
        self._grad = None
        self.status = _eng.ST_OK
        self.ep_sweeps = 0
        self.bic = float("inf")
        if _infer:
            self.parameters_changed()

    # -- GPy-facing surface --
    @property
    def input_dim(self):
        return self.X.shape[1]

    @property
    def num_data(self):
        return len(self._rows)

    def parameters_changed(self, warm: bool = False):
        r = self.engine.score_batch(self._dataset, [self._program], [self.kern.param_array], [self._rows],
                                    tau=[self._tau] if (warm and self._tau is not None) else None,
                                    nu=[self._nu] if (warm and self._nu is not None) else None, **self.ep)
        self._absorb(r, 0)

    def _absorb(self, r, b):
        self._nlml = float(r["nlml"][b])
        self._grad = np.array(r["grad"][b, :self._program.n_theta])
        self.bic = float(r["bic"][b])
        self.status = int(r["status"][b])
        self.ep_sweeps = int(r["sweeps"][b])
        self._tau, self._nu = r["tau"][b], r["nu"][b]

    def log_likelihood(self) -> float:
        return -self._nlml

    def objective_function(self) -> float:
        return self._nlml

    def optimize(self, optimizer=None, max_iters: int = 1000, messages: bool = False, **kw):
        """``m.optimize()`` (gpckernel.py:246): L-BFGS-B (m=10, factr=1e7, pgtol=1e-5) in the unconstrained space."""
        optim.optimize_models([self], max_evals=max_iters)
        return self

    def _ensure_sites(self):
        if self._tau is None:
            self.parameters_changed()

    def predict(self, Xnew, full_cov=False, **kw):
        """(Phi[m x 1], None): probit predictive mean -- the only output gpckernel.py:832 reads."""
        p, _, _ = self._predict_points(np.asarray(Xnew, float))
        return p.reshape(-1, 1), None

    def _raw_predict(self, Xnew, full_cov=False, **kw):
        _, mu, var = self._predict_points(np.asarray(Xnew, float))
        return mu.reshape(-1, 1), var.reshape(-1, 1)

    def predict_rows(self, test_rows):
        """Predictive probabilities for rows of the model's own device dataset (no upload)."""
        self._ensure_sites()
        p, st = self.engine.predict_batch(self._dataset, [self._program], [self.kern.param_array], [self._rows],
                                          [self._tau], [self._nu], [np.asarray(test_rows, dtype=np.int32)])
        return p[0]

    def _predict_points(self, Xnew):
        self._ensure_sites()
        return self.engine.predict_points(self._dataset, self._program, self.kern.param_array, self._rows, self._tau,
                                          self._nu, Xnew)

    def predictive_gradients(self, Xnew, **kw):
        """(d mu*/d x* [m x D x 1], None) -- gpckernel.py:431 (monotonicity)."""
        self._ensure_sites()
        g = self.engine.predict_points(self._dataset, self._program, self.kern.param_array, self._rows, self._tau,
                                       self._nu, np.asarray(Xnew, float), want_gradients=True)[3]
        return g[:, :, None], None

    def __del__(self):
        try:
            if self._owns_dataset and self.engine is not None:
                self.engine.dataset_free(self._dataset)
        except Exception:
            pass


models = SimpleNamespace(GPClassification=GPClassification)


class _EP:
    name = "ep"


class _Laplace:
    name = "laplace"


inference = SimpleNamespace(latent_function_inference=SimpleNamespace(EP=_EP, Laplace=_Laplace))
