"""Search grammar: candidate generation for one greedy expansion step (host side, symbolic).

Python 3 re-implementation of the behaviour of /root/reference/grammar.py: production rules ``A -> A + B`` and
``A -> A x B`` (``:20-21``; B a base kernel, not a constant for the product rule), applied to the whole tree, to every
operand, and to every non-trivial operand subset of an n-ary operator (``:129-161``).  The change-point / window
operators of the GPSS original are out of scope (never produced by MULTI_D_RULES).
"""
from __future__ import annotations

import itertools
from typing import List

from . import flexible_function as ff

# (lhs, operator, type of the new operand)
MULTI_D_RULES = [("A", ("+", "A", "B"), {"A": "kernel", "B": "base"}),
                 ("A", ("*", "A", "B"), {"A": "kernel", "B": "base-not-const"})]


class MultiDGrammar:
    def __init__(self, ndim: int, base_kernels: str = "SE", rules=None):
        self.rules = rules if rules else MULTI_D_RULES
        self.ndim = ndim
        self.base_kernels = base_kernels

    def type_matches(self, candidate, tp: str) -> bool:
        if tp == "kernel":
            return isinstance(candidate, ff.Kernel)
        if tp == "base":
            return isinstance(candidate, ff.Kernel) and not candidate.is_operator
        if tp == "base-not-const":
            return isinstance(candidate, ff.Kernel) and not candidate.is_operator and not isinstance(candidate, ff.ConstKernel)
        if tp == "dimension":
            return isinstance(candidate, int)
        raise RuntimeError("Unknown type: %s" % tp)

    def list_options(self, tp: str) -> list:
        if tp == "base":
            return list(ff.base_kernels(self.ndim, self.base_kernels))
        if tp == "base-not-const":
            return [k for k in self.list_options("base") if not isinstance(k, ff.ConstKernel)]
        if tp == "dimension":
            return list(range(self.ndim))
        raise RuntimeError("Cannot expand the '%s' type" % tp)


def _apply(op: str, kernel: ff.Kernel, new: ff.Kernel) -> ff.Kernel:
    cls = ff.SumKernel if op == "+" else ff.ProductKernel
    return cls([kernel.copy(), new.copy()])


def expand_single_tree(kernel: ff.Kernel, grammar: MultiDGrammar) -> List[ff.Kernel]:
    """All single-production rewrites of the root."""
    assert isinstance(kernel, ff.Kernel)
    out = []
    for lhs, rhs, types in grammar.rules:
        if not grammar.type_matches(kernel, types[lhs]):
            continue
        op = rhs[0]
        free = [v for v in rhs[1:] if v != lhs]
        for choice in itertools.product(*[grammar.list_options(types[v]) for v in free]):
            out.append(_apply(op, kernel, choice[0]))
    return out


def expand(kernel: ff.Kernel, grammar: MultiDGrammar) -> List[ff.Kernel]:
    """Rewrites of the root, of each operand (recursively), and of each non-trivial operand subset."""
    result = expand_single_tree(kernel, grammar)
    if not kernel.is_operator:
        return result
    ops = kernel.operands
    for mask in itertools.product((0, 1), repeat=len(ops)):
        if sum(mask) in (0, len(ops)):
            continue
        keep = [o.copy() for o, m in zip(ops, mask) if not m]
        chosen = [o.copy() for o, m in zip(ops, mask) if m]
        if len(chosen) > 1:
            expansions = expand_single_tree(type(kernel)(chosen), grammar)
        else:
            expansions = expand(chosen[0], grammar)
        for e in expansions:
            result.append(type(kernel)([e] + [o.copy() for o in keep]))
    return result


def expand_kernels(D: int, seed_kernels, base_kernels: str = "SE", rules=None) -> List[ff.Kernel]:
    """All expansions of a set of kernels in D dimensions, canonical and de-duplicated (grammar.py:165-174)."""
    g = MultiDGrammar(D, base_kernels=base_kernels, rules=rules)
    kernels: List[ff.Kernel] = []
    for k in seed_kernels:
        kernels += expand(k, g)
    kernels = ff.remove_duplicates([k.canonical() for k in kernels])
    return [k for k in kernels if not isinstance(k, ff.NoneKernel)]
