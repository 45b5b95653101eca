"""Kernel algebra for the AutoGPC search (host side, structure only -- no numerics).

Python 3 re-implementation of the part of the reference's GPSS fork that is reachable from ``gpss2gpy``
(/root/reference/flexible_function.py): ``NoneKernel``, ``ConstKernel``, ``SqExpKernel``, ``PeriodicKernel``,
``LinearKernel``, ``SumKernel``, ``ProductKernel``; canonical / simplified / additive forms (``:264-301``,
``:303-348``, ``:404-570``), ``effective_params`` (``:69-74``, ``:1554``, ``:1647``), ``GPModel.bic`` (``:685-687``),
``base_kernels`` (``:2061-2085``) and ``remove_duplicates`` (``:2055-2057``).

Deliberate differences: the reference orders operands with Python 2 ``__cmp__`` on class objects
and ``repr`` and de-duplicates through ``list(set(...))`` -- both hash/address dependent.  Here the order is an
explicit total order (class rank, dimension, structure, then parameters) and de-duplication preserves first
occurrence, so candidate lists are reproducible.  Mean functions, likelihood wrappers, RQ / spectral / change-point
kernels and GPML strings are out of scope (never reachable from gpckernel.py:514-574).
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, Iterable, List, Optional

import numpy as np

_CLASS_RANK = {"NoneKernel": 0, "ConstKernel": 1, "LinearKernel": 2, "PeriodicKernel": 3, "SqExpKernel": 4,
               "ProductKernel": 5, "SumKernel": 6}

_rng = np.random.default_rng(0)


def set_rng(rng) -> None:
    """Generator used by ``initialise_params`` (the reference draws from the unseeded global np.random)."""
    global _rng
    _rng = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)


def _fmt(v, pattern="%1.1f"):
    try:
        return pattern % v
    except TypeError:
        return str(v)


class Kernel:
    """Base class: leaves override the properties, operators override ``operands`` handling."""
    is_operator = False
    is_thunk = False          # dimensionless (Const)
    is_stationary = True
    arity = 0

    # ---- identity / ordering -------------------------------------------------------------------------------
    @property
    def id(self) -> str:
        raise NotImplementedError

    @property
    def param_vector(self) -> np.ndarray:
        raise NotImplementedError

    @property
    def num_params(self) -> int:
        return len(self.param_vector)

    @property
    def effective_params(self) -> int:
        return len(self.param_vector)

    @property
    def depth(self) -> int:
        return 0

    def structure(self) -> str:
        """Parameter-free structural string, e.g. ``SE0*Per2+C``."""
        raise NotImplementedError

    def _key(self):
        return (_CLASS_RANK[type(self).__name__], getattr(self, "dimension", -1) if not self.is_operator else -1,
                self.structure(), repr(self))

    def __eq__(self, other):
        return isinstance(other, Kernel) and repr(self) == repr(other)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __lt__(self, other):
        return self._key() < other._key()

    def __hash__(self):
        return hash(repr(self))

    # ---- sugar ---------------------------------------------------------------------------------------------
    def __add__(self, other):
        assert isinstance(other, Kernel)
        return SumKernel([self.copy(), other.copy()]).canonical()

    def __mul__(self, other):
        assert isinstance(other, Kernel)
        return ProductKernel([self.copy(), other.copy()]).canonical()

    # ---- transforms ----------------------------------------------------------------------------------------
    def copy(self) -> "Kernel":
        raise NotImplementedError

    def canonical(self) -> "Kernel":
        return self.copy()

    def additive_form(self) -> "Kernel":
        return self.canonical()

    def break_into_summands(self) -> List["Kernel"]:
        k = self.additive_form()
        return list(k.operands) if isinstance(k, SumKernel) else [k]

    def simplified(self) -> "Kernel":
        k, prev = self.canonical(), None
        while prev is None or repr(prev) != repr(k):
            prev = k
            k = _merge_sum_constants(k)
            k = _merge_product_duplicates(k)
            k = _drop_product_constants(k)
            k = k.canonical()
        return k

    def multiply_by_const(self, sf) -> None:
        if hasattr(self, "sf") and self.sf is not None and sf is not None:
            self.sf += sf

    def initialise_params(self, sd=1, data_shape=None) -> None:
        raise NotImplementedError

    def load_param_vector(self, params) -> None:
        raise NotImplementedError

    def pretty_print(self) -> str:
        raise NotImplementedError

    def out_of_bounds(self, constraints) -> bool:
        return False


class NoneKernel(Kernel):
    id = "None"

    def copy(self):
        return NoneKernel()

    def structure(self):
        return "None"

    def __repr__(self):
        return "NoneKernel()"

    @property
    def param_vector(self):
        return np.array([])

    def initialise_params(self, sd=1, data_shape=None):
        pass

    def multiply_by_const(self, sf):
        pass

    def pretty_print(self):
        return "NoneKernel"


class ConstKernel(Kernel):
    id = "Const"
    is_thunk = True

    def __init__(self, sf=None):
        self.sf = sf

    def copy(self):
        return ConstKernel(sf=self.sf)

    def structure(self):
        return "C"

    @property
    def param_vector(self):
        return np.array([self.sf])

    def initialise_params(self, sd=1, data_shape=None):
        if self.sf is None:
            r = _rng.random()
            loc = math.log(abs(data_shape["y_mean"]) + 1e-12) if r < 1 / 3 else (data_shape["y_sd"] if r < 2 / 3 else 0.0)
            self.sf = float(_rng.normal(loc, sd))

    def load_param_vector(self, params):
        (self.sf,) = params

    def __repr__(self):
        return "ConstKernel(sf=%s)" % (self.sf,)

    def pretty_print(self):
        return "C(sf=%s)" % _fmt(self.sf)


class _DimKernel(Kernel):
    """Leaf acting on one input dimension."""

    def structure(self):
        return "%s%s" % (self.id, self.dimension)


class SqExpKernel(_DimKernel):
    id = "SE"

    def __init__(self, dimension=None, lengthscale=None, sf=None):
        self.dimension, self.lengthscale, self.sf = dimension, lengthscale, sf

    def copy(self):
        return SqExpKernel(self.dimension, self.lengthscale, self.sf)

    @property
    def param_vector(self):
        return np.array([self.lengthscale, self.sf])

    def initialise_params(self, sd=1, data_shape=None):
        d = self.dimension
        if self.lengthscale is None:
            if _rng.random() < 0.5:
                self.lengthscale = float(_rng.normal(data_shape["x_sd"][d], sd))
            else:
                self.lengthscale = float(_rng.normal(math.log(2 * (data_shape["x_max"][d] - data_shape["x_min"][d])), sd))
        if self.sf is None:
            self.sf = float(_rng.normal(data_shape["y_sd"] if _rng.random() < 0.5 else 0.0, sd))

    def load_param_vector(self, params):
        self.lengthscale, self.sf = params

    def __repr__(self):
        return "SqExpKernel(dimension=%s, lengthscale=%s, sf=%s)" % (self.dimension, self.lengthscale, self.sf)

    def pretty_print(self):
        return "SE(d=%s, l=%s, sf=%s)" % (self.dimension, _fmt(self.lengthscale), _fmt(self.sf))


class PeriodicKernel(_DimKernel):
    id = "Per"

    def __init__(self, dimension=None, lengthscale=None, period=None, sf=None):
        self.dimension, self.lengthscale, self.period, self.sf = dimension, lengthscale, period, sf

    def copy(self):
        return PeriodicKernel(self.dimension, self.lengthscale, self.period, self.sf)

    @property
    def param_vector(self):
        return np.array([self.lengthscale, self.period, self.sf])

    def initialise_params(self, sd=1, data_shape=None):
        if self.lengthscale is None:
            self.lengthscale = float(_rng.normal(0.0, sd))
        if self.period is None:
            r = _rng.random()
            self.period = float(_rng.normal(0.0 if r < 1 / 3 else (-3.95 if r < 2 / 3 else -5.9), sd))
        if self.sf is None:
            self.sf = float(_rng.normal(data_shape["y_sd"] if _rng.random() < 0.5 else 0.0, sd))

    def load_param_vector(self, params):
        self.lengthscale, self.period, self.sf = params

    def out_of_bounds(self, constraints):
        return (self.period < constraints["min_period"][self.dimension]) or \
               (self.period > constraints["max_period"][self.dimension])

    def __repr__(self):
        return "PeriodicKernel(dimension=%s, lengthscale=%s, period=%s, sf=%s)" % \
               (self.dimension, self.lengthscale, self.period, self.sf)

    def pretty_print(self):
        return "Per(d=%s, l=%s, p=%s, sf=%s)" % (self.dimension, _fmt(self.lengthscale), _fmt(self.period), _fmt(self.sf))


class LinearKernel(_DimKernel):
    id = "Lin"
    is_stationary = False

    def __init__(self, dimension=None, location=None, sf=None):
        self.dimension, self.location, self.sf = dimension, location, sf

    def copy(self):
        return LinearKernel(self.dimension, self.location, self.sf)

    @property
    def param_vector(self):
        return np.array([self.sf, self.location])

    def initialise_params(self, sd=1, data_shape=None):
        d = self.dimension
        if self.sf is None:
            r = _rng.random()
            if r < 1 / 3:
                loc = data_shape["y_sd"] - data_shape["x_sd"][d]
            elif r < 2 / 3:
                loc = math.log(abs(1.0 / (data_shape["x_max"][d] - data_shape["x_min"][d])))
            else:
                loc = 0.0
            self.sf = float(_rng.normal(loc, sd))
        if self.location is None:
            lo = 2 * data_shape["x_min"][d] - data_shape["x_max"][d]
            hi = 2 * data_shape["x_max"][d] - data_shape["x_min"][d]
            self.location = float(_rng.uniform(lo, hi))

    def load_param_vector(self, params):
        self.sf, self.location = params

    def __repr__(self):
        return "LinearKernel(dimension=%s, location=%s, sf=%s)" % (self.dimension, self.location, self.sf)

    def pretty_print(self):
        return "Lin(d=%s, loc=%s, sf=%s)" % (self.dimension, _fmt(self.location), _fmt(self.sf))


class _NaryKernel(Kernel):
    is_operator = True
    arity = "n"
    is_abelian = True
    _symbol = "?"

    def __init__(self, operands: Optional[Iterable[Kernel]] = None):
        self.operands = list(operands) if operands is not None else []

    @property
    def is_stationary(self):
        return all(o.is_stationary for o in self.operands)

    @property
    def param_vector(self):
        return np.concatenate([o.param_vector for o in self.operands]) if self.operands else np.array([])

    @property
    def depth(self):
        return max(o.depth for o in self.operands) + 1

    def copy(self):
        return type(self)([o.copy() for o in self.operands])

    def initialise_params(self, sd=1, data_shape=None):
        for o in self.operands:
            o.initialise_params(sd=sd, data_shape=data_shape)

    def load_param_vector(self, params):
        start = 0
        for o in self.operands:
            o.load_param_vector(params[start:start + o.num_params])
            start += o.num_params

    def out_of_bounds(self, constraints):
        return any(o.out_of_bounds(constraints) for o in self.operands)

    def canonical(self):
        """Flatten nested operators of the same type, drop NoneKernel, unwrap singletons, sort operands
        (reference Kernel.canonical, flexible_function.py:278-301)."""
        ops: List[Kernel] = []
        for o in self.operands:
            c = o.canonical()
            if isinstance(c, type(self)):
                ops.extend(c.operands)
            elif not isinstance(c, NoneKernel):
                ops.append(c)
        if not ops:
            return NoneKernel()
        if len(ops) == 1:
            return ops[0]
        return type(self)(sorted(ops))

    def __repr__(self):
        return "%s(operands=[%s])" % (type(self).__name__, ", ".join(repr(o) for o in self.operands))

    def pretty_print(self):
        return "( " + (" %s " % self._symbol).join(o.pretty_print() for o in self.operands) + " ) "


class SumKernel(_NaryKernel):
    id = "Sum"
    _symbol = "+"

    def structure(self):
        return "+".join(o.structure() for o in self.operands)

    @property
    def effective_params(self):
        return sum(o.effective_params for o in self.operands)

    def multiply_by_const(self, sf):
        for o in self.operands:
            o.multiply_by_const(sf)

    def additive_form(self):
        k = self.canonical()
        if not isinstance(k, SumKernel):
            return k.additive_form()
        return SumKernel([o.additive_form() for o in k.operands]).canonical()


class ProductKernel(_NaryKernel):
    id = "Product"
    _symbol = "x"

    def structure(self):
        return "*".join("(%s)" % o.structure() if isinstance(o, SumKernel) else o.structure() for o in self.operands)

    @property
    def effective_params(self):
        return sum(o.effective_params for o in self.operands) - (len(self.operands) - 1)

    def multiply_by_const(self, sf):
        self.operands[0].multiply_by_const(sf)

    def additive_form(self):
        """Distribute the product over sums: result is a sum of products in canonical form
        (reference Kernel.additive_form, flexible_function.py:303-348)."""
        k = self.canonical()
        if not isinstance(k, ProductKernel):
            return k.additive_form()
        choices = []
        for o in k.operands:
            a = o.additive_form()
            choices.append(list(a.operands) if isinstance(a, SumKernel) else [a])
        terms = [ProductKernel([f.copy() for f in combo]).canonical() for combo in itertools.product(*choices)]
        return SumKernel(terms).canonical() if len(terms) > 1 else terms[0]


# ---- simplification passes (reference flexible_function.py:404-570, restricted to the in-scope classes) ---------
def _map_operands(k: Kernel, fn) -> Kernel:
    if not k.is_operator:
        return k
    k = k.copy()
    k.operands = [fn(o) for o in k.operands]
    return k


def _merge_sum_constants(k: Kernel) -> Kernel:
    """Several constants in one sum are one constant (collapse_additive_idempotency)."""
    k = _map_operands(k.canonical(), _merge_sum_constants)
    if isinstance(k, SumKernel):
        consts = [o for o in k.operands if isinstance(o, ConstKernel)]
        if len(consts) > 1:
            sfs = [c.sf for c in consts]
            # log-sum-exp: a trained (linear-scale) sf above ~355 must not overflow -- the reference's np.exp gives inf
            # with a warning and the merged value is re-drawn by GPCKernel.__init__ anyway
            sf = None if any(s is None for s in sfs) else 0.5 * float(np.logaddexp.reduce([2.0 * s for s in sfs]))
            k.operands = [o for o in k.operands if not isinstance(o, ConstKernel)] + [ConstKernel(sf=sf)]
    return k.canonical()


def _merge_product_duplicates(k: Kernel) -> Kernel:
    """SE x SE on one dimension is one SE; several constants in one product are one constant
    (collapse_multiplicative_idempotency)."""
    k = _map_operands(k.canonical(), _merge_product_duplicates)
    if isinstance(k, ProductKernel):
        by_dim: Dict[int, SqExpKernel] = {}
        rest, consts = [], []
        for o in k.operands:
            if isinstance(o, SqExpKernel):
                if o.dimension in by_dim:
                    a = by_dim[o.dimension]
                    if a.lengthscale is not None and o.lengthscale is not None:
                        a.lengthscale = -0.5 * math.log(math.exp(-2 * a.lengthscale) + math.exp(-2 * o.lengthscale))
                    if a.sf is not None and o.sf is not None:
                        a.sf += o.sf
                else:
                    by_dim[o.dimension] = o.copy()
            elif isinstance(o, ConstKernel):
                consts.append(o)
            else:
                rest.append(o)
        ops = rest + [by_dim[d] for d in sorted(by_dim)]
        if consts:
            sfs = [c.sf for c in consts]
            ops.append(ConstKernel(sf=None if any(s is None for s in sfs) else sum(sfs)))
        k.operands = ops
    return k.canonical()


def _drop_product_constants(k: Kernel) -> Kernel:
    """A constant factor is absorbed by the first remaining factor's scale (collapse_multiplicative_identity)."""
    k = _map_operands(k.canonical(), _drop_product_constants)
    if isinstance(k, ProductKernel):
        consts = [o for o in k.operands if isinstance(o, ConstKernel)]
        others = [o for o in k.operands if not isinstance(o, ConstKernel)]
        if consts and others:
            sf = None if any(c.sf is None for c in consts) else sum(c.sf for c in consts)
            others[0].multiply_by_const(sf)
            k.operands = others
    return k.canonical()


# ---- model-level helpers -----------------------------------------------------------------------------------
class GPModel:
    """Kernel + score holder; only the information criteria of the reference are kept (flexible_function.py:685-696)."""

    def __init__(self, kernel=None, nll=None, ndata=None, **_ignored):
        self.kernel, self.nll, self.ndata = kernel, nll, ndata

    @property
    def bic(self):
        return 2 * self.nll + self.kernel.effective_params * np.log(self.ndata)

    @property
    def aic(self):
        return 2 * self.nll + self.kernel.effective_params * 2

    @property
    def pl2(self):
        return self.nll / self.ndata + self.kernel.effective_params / (2 * self.ndata)

    def copy(self):
        return GPModel(self.kernel.copy() if self.kernel is not None else None, self.nll, self.ndata)

    def __repr__(self):
        return "GPModel(kernel=%r, nll=%s, ndata=%s)" % (self.kernel, self.nll, self.ndata)


def repr_to_model(string: str):
    """De-facto (de)serialiser of the reference (flexible_function.py:2052-2053)."""
    return eval(string)  # noqa: S307 - mirrors the reference; input is this module's own repr()


def remove_duplicates(things):
    """Order-preserving (first occurrence wins); the reference's list(set()) order is hash dependent."""
    seen, out = set(), []
    for t in things:
        r = repr(t)
        if r not in seen:
            seen.add(r)
            out.append(t)
    return out


def base_kernels_without_dimension(base_kernel_names: str):
    names = [n.strip() for n in base_kernel_names.split(",")]
    for k in (SqExpKernel(), ConstKernel(), LinearKernel(), PeriodicKernel()):
        if k.id in names:
            yield k


def base_kernels(dimensions: int = 1, base_kernel_names: str = "SE"):
    for kernel in base_kernels_without_dimension(base_kernel_names):
        if kernel.is_thunk:
            yield kernel
        else:
            for d in range(dimensions):
                k = kernel.copy()
                k.dimension = d
                yield k
