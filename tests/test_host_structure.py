"""Host-side structure tests (CPU): the reference's only assertions (gpckerneltest.py:84-107) plus the shapes its
print-driven scripts exercise (gpss2gpy / gpy2gpss / expand), ported to asserts."""
import numpy as np
import pytest

from autogpc_b200 import flexible_function as ff
from autogpc_b200 import gpckernel, grammar
from autogpc_b200 import gpy_compat as GPy
from autogpc_b200.gpcdata import GPCData
from tests.util import synth


@pytest.fixture()
def data():
    X, y = synth(60, 3, 1)
    return GPCData(X, y.reshape(-1, 1))


def test_is_kernel_equal_truth_table():
    # gpckerneltest.py:86-107, assertion for assertion
    k1, k2 = ff.NoneKernel(), ff.NoneKernel()
    assert gpckernel.isKernelEqual(k1, k2)
    k2 = ff.SqExpKernel(dimension=0, lengthscale=1.5, sf=0.5)
    assert not gpckernel.isKernelEqual(k1, k2)
    k1 = ff.SqExpKernel(dimension=0, lengthscale=1, sf=0.5)
    assert gpckernel.isKernelEqual(k1, k2)
    assert not gpckernel.isKernelEqual(k1, k2, compare_params=True)
    assert gpckernel.isKernelEqual(k1, k1.copy(), compare_params=True)
    k2 = ff.SqExpKernel(dimension=1, lengthscale=1, sf=0.5)
    assert not gpckernel.isKernelEqual(k1, k2)
    k1 = ff.SqExpKernel(dimension=0, lengthscale=1, sf=0.5)
    k2 = ff.SqExpKernel(dimension=1, lengthscale=1, sf=0.5)
    k3, k4 = ff.SumKernel([k1, k2]), ff.SumKernel([k2, k1])
    assert gpckernel.isKernelEqual(k3, k4)
    k4 = ff.SumKernel([k1, k1])
    assert not gpckernel.isKernelEqual(k3, k4)


def test_gpss2gpy_shapes_and_bounds(data):
    # gpckerneltest.py:32-56 / gpckernel.py:535-565
    se = gpckernel.gpss2gpy(ff.SqExpKernel(dimension=1, lengthscale=1.5, sf=0.5), data=data)
    assert isinstance(se, GPy.kern.RBF) and list(se.active_dims) == [1]
    assert se.variance[0] == pytest.approx(0.25) and se.lengthscale[0] == pytest.approx(1.5)
    lo, hi = np.ravel(data.getLengthscaleBounds(dims=1))
    assert se["lengthscale"].kind == "bounded" and (se["lengthscale"].lo, se["lengthscale"].hi) == (lo, hi)
    per = gpckernel.gpss2gpy(ff.PeriodicKernel(dimension=0, lengthscale=5, period=4, sf=1.5), data=data)
    assert [p.name for p in per.parameters] == ["variance", "period", "lengthscale"]
    assert per["period"].kind == "bounded" and per["lengthscale"].kind == "bounded"
    c = gpckernel.gpss2gpy(ff.ConstKernel(sf=2.0), data=data)
    assert isinstance(c, GPy.kern.Bias) and list(c.active_dims) == [0, 1, 2] and c.variance[0] == 4.0
    s = gpckernel.gpss2gpy(ff.SumKernel([ff.SqExpKernel(0, 1.0, 1.0), ff.PeriodicKernel(2, 1.0, 1.0, 1.0)]), data=data)
    assert isinstance(s, GPy.kern.Add) and list(s.active_dims) == [0, 2]
    p = gpckernel.gpss2gpy(ff.ProductKernel([ff.SqExpKernel(0, 1.0, 1.0), s_ := ff.SqExpKernel(1, 1.0, 1.0)]), data=data)
    assert isinstance(p, GPy.kern.Prod) and len(p.parts) == 2
    with pytest.raises(NotImplementedError):
        gpckernel.gpss2gpy(ff.NoneKernel(), data=data)


def test_gpy2gpss_roundtrip(data):
    # gpckerneltest.py:58-69
    kk = GPy.kern.Add([GPy.kern.RBF(1, variance=2, lengthscale=3, active_dims=np.array([3])),
                       GPy.kern.RBF(1, variance=4, lengthscale=6, active_dims=np.array([1]))])
    g = gpckernel.gpy2gpss(kk)
    assert isinstance(g, ff.SumKernel)
    assert (g.operands[0].dimension, g.operands[0].lengthscale, g.operands[0].sf) == (3, 3.0, pytest.approx(np.sqrt(2)))
    c = gpckernel.gpy2gpss(GPy.kern.Bias(3, variance=3, active_dims=np.array([0, 1, 2])))
    assert isinstance(c, ff.ConstKernel) and c.sf == pytest.approx(np.sqrt(3))
    k = ff.ProductKernel([ff.SqExpKernel(0, 1.25, 0.5), ff.SumKernel([ff.PeriodicKernel(1, 0.5, 2.0, 1.5), ff.ConstKernel(2.0)])])
    back = gpckernel.gpy2gpss(gpckernel.gpss2gpy(k, data=data))
    assert gpckernel.isKernelEqual(k, back, compare_params=False)
    assert np.allclose(back.param_vector, k.param_vector)


def test_expand_se0_in_2d():
    # gpckerneltest.py:71-77: expected set from grammar.py:20-21 after canonical + simplified
    X, y = synth(40, 2, 0)
    d = GPCData(X, y.reshape(-1, 1))
    ks = gpckernel.GPCKernel(ff.SqExpKernel(dimension=0, lengthscale=1.5, sf=0.5), d, 1).expand()
    assert sorted(k.kernel.structure() for k in ks) == ["SE0", "SE0*SE1", "SE0+SE0", "SE0+SE1"]
    assert all(k.depth == 2 and k.parent is not None for k in ks)


def test_grammar_counts_d8_three_bases():
    # depth-0 = 24 (NoneKernel + B and NoneKernel x B both canonicalise to B), 48 raw
    g = grammar.MultiDGrammar(8, "SE,Lin,Per")
    raw = grammar.expand(ff.NoneKernel(), g)
    assert len(raw) == 48
    uniq = [k for k in ff.remove_duplicates([k.canonical() for k in raw]) if not isinstance(k, ff.NoneKernel)]
    assert len(uniq) == 24
    raw1 = grammar.expand(ff.SqExpKernel(0, 1.0, 1.0), g)
    assert len(raw1) == 48


def test_canonical_additive_and_peff():
    a, b, c = ff.SqExpKernel(0, 1.0, 1.0), ff.SqExpKernel(1, 1.0, 1.0), ff.PeriodicKernel(2, 1.0, 1.0, 1.0)
    k = ff.ProductKernel([a, ff.SumKernel([b, c])])
    assert k.effective_params == 2 + (2 + 3) - 1                           # flexible_function.py:1647
    add = k.additive_form()
    assert isinstance(add, ff.SumKernel) and sorted(o.structure() for o in add.operands) == ["Per2*SE0", "SE0*SE1"]
    nested = ff.SumKernel([ff.SumKernel([b, a]), ff.NoneKernel(), c])
    assert nested.canonical().structure() == "Per2+SE0+SE1"
    assert ff.SumKernel([ff.NoneKernel()]).canonical().structure() == "None"
    # SE x SE on one dimension collapses; constants fold into the first factor
    prod = ff.ProductKernel([ff.SqExpKernel(0, 0.1, 0.2), ff.SqExpKernel(0, 0.3, 0.4), ff.ConstKernel(0.5)])
    assert prod.simplified().structure() == "SE0"
    assert ff.repr_to_model.__doc__ and repr(eval(repr(k), vars(ff))) == repr(k)


def test_compile_program_theta_order_and_terms(data):
    k = ff.ProductKernel([ff.SqExpKernel(0, 1.5, 0.5), ff.SumKernel([ff.PeriodicKernel(1, 0.5, 2.0, 1.5), ff.ConstKernel(2.0)])])
    gk = gpckernel.gpss2gpy(k, data=data)
    prog = GPy.compile_program(gk)
    # depth-first leaf order, GPy parameter order per leaf (gpckernel.py:535-565)
    assert [prog.leaf_type[i] for i in range(prog.n_leaves)] == [0, 1, 3]
    assert [prog.leaf_off[i] for i in range(prog.n_leaves)] == [0, 2, 5]
    assert prog.n_theta == 6 and prog.p_eff == k.effective_params == 5
    assert prog.terms() == [[0, 1], [0, 2]]
    assert np.allclose(gk.param_array, [0.25, 1.5, 2.25, 2.0, 0.5, 4.0])


def test_reset_params_in_bounds_and_seeded(data):
    k = ff.SumKernel([ff.SqExpKernel(0), ff.PeriodicKernel(1), ff.ConstKernel(), ff.LinearKernel(2)])
    gpckernel.set_rng(123)
    gpckernel.resetGpssParams(k, data=data)
    first = k.param_vector.copy()
    lo, hi = np.ravel(data.getLengthscaleBounds(dims=0))
    assert lo <= k.operands[0].lengthscale <= hi
    lo, hi = np.ravel(data.getPeriodBounds(dims=1))
    assert lo <= k.operands[1].period <= hi and k.operands[1].lengthscale > 0
    assert all(o.sf > 0 for o in k.operands)
    gpckernel.set_rng(123)
    gpckernel.resetGpssParams(k, data=data)
    assert np.array_equal(first, k.param_vector)


def test_gpcdata_bounds_and_folds():
    X, y = synth(103, 2, 5)
    d = GPCData(X, y.reshape(-1, 1), seed=7)
    b = d.getLengthscaleBounds()
    assert b.shape == (2, 2) and np.all(b[0] > 0) and np.allclose(b[1], 2 * (X.max(0) - X.min(0)))
    folds = d.kFoldIndices(5)
    assert [len(te) for _, te in folds] == [21, 21, 21, 20, 20]        # sklearn KFold sizes
    assert sorted(np.concatenate([te for _, te in folds]).tolist()) == list(range(103))
    assert all(len(set(tr) & set(te)) == 0 and len(tr) + len(te) == 103 for tr, te in folds)
    assert d.kFoldIndices(5) is folds                                   # cached (gpcdata.py:163-164)
    Xs, Ys, XT, YT = d.kFoldSplits(5)
    assert Xs[0].shape == (82, 2) and YT[4].shape == (20, 1)
    one = d.kFoldIndices(1)
    assert len(one) == 1 and np.array_equal(one[0][0], one[0][1])


def test_param_transforms_roundtrip_and_gradfactor():
    p = GPy.Param("v", 0.7)
    assert p.from_x(p.to_x()) == pytest.approx(0.7)
    h = 1e-6
    fd = (p.from_x(p.to_x() + h) - p.from_x(p.to_x() - h)) / (2 * h)
    assert p.gradfactor() == pytest.approx(fd, rel=1e-7)
    q = GPy.Param("l", 5.0).constrain_bounded(0.1, 3.0)
    assert q[0] == pytest.approx(1.55)                                  # out-of-bounds start -> midpoint
    q.values[0] = 2.0
    assert q.from_x(q.to_x()) == pytest.approx(2.0)
    fd = (q.from_x(q.to_x() + h) - q.from_x(q.to_x() - h)) / (2 * h)
    assert q.gradfactor() == pytest.approx(fd, rel=1e-6)


def test_compile_program_product_of_sums_registers_each_leaf_once():
    """(SE0 + SE1) * (SE2 + SE3): distribution gives 4 terms over the SAME 4 leaves (8 theta slots = GPy's param_array);
    K from the flat program equals the product of the two sums."""
    import numpy as np
    from autogpc_b200 import gpy_compat as GPy
    from oracle import gpc_oracle as orc
    mk = lambda d, v, l: GPy.kern.RBF(1, variance=v, lengthscale=l, active_dims=np.array([d]))
    k = (mk(0, 1.1, 0.9) + mk(1, 0.7, 1.3)) * (mk(2, 1.5, 0.8) + mk(3, 0.6, 2.0))
    prog = GPy.compile_program(k)
    assert prog.n_leaves == 4 and prog.n_theta == 8 == len(k.param_array) and prog.n_terms == 4
    terms = sorted(tuple(prog.term_leaf[t][:prog.term_len[t]]) for t in range(prog.n_terms))
    assert terms == [(0, 2), (0, 3), (1, 2), (1, 3)]
    # three factors, the middle one a sum: still one registration per leaf
    k3 = mk(0, 1.0, 1.0) * (mk(1, 1.0, 1.0) + mk(2, 1.0, 1.0)) * (mk(3, 1.0, 1.0) + GPy.kern.Bias(1, variance=0.5))
    p3 = GPy.compile_program(k3)
    assert p3.n_leaves == 5 and p3.n_theta == len(k3.param_array) == 9 and p3.n_terms == 4
    X = np.random.default_rng(0).uniform(-2, 2, (30, 4))
    oprog = orc.KernelProgram([prog.leaf_type[i] for i in range(4)], [prog.leaf_dim[i] for i in range(4)],
                              [prog.leaf_off[i] for i in range(4)], [list(t) for t in terms], 8)
    se = lambda d, v, l: v * np.exp(-0.5 * (X[:, d][:, None] - X[:, d][None, :]) ** 2 / l ** 2)
    dense = (se(0, 1.1, 0.9) + se(1, 0.7, 1.3)) * (se(2, 1.5, 0.8) + se(3, 0.6, 2.0))
    assert np.allclose(orc.kern_K(oprog, k.param_array, X), dense, rtol=1e-12)


def _dense_K(k, X):
    """Direct recursive evaluation of the GPy-shaped tree (GPy kern.K semantics), independent of compile_program."""
    import numpy as np
    from autogpc_b200 import gpy_compat as GPy
    if isinstance(k, GPy.kern.Add):
        return sum(_dense_K(p, X) for p in k.parts)
    if isinstance(k, GPy.kern.Prod):
        out = np.ones((X.shape[0], X.shape[0]))
        for p in k.parts:
            out = out * _dense_K(p, X)
        return out
    v = float(k.variance)
    if isinstance(k, GPy.kern.Bias):
        return np.full((X.shape[0], X.shape[0]), v)
    x = X[:, int(k.active_dims[0])]
    r = x[:, None] - x[None, :]
    if isinstance(k, GPy.kern.RBF):
        return v * np.exp(-0.5 * r ** 2 / float(k.lengthscale) ** 2)
    if isinstance(k, GPy.kern.StdPeriodic):
        return v * np.exp(-0.5 * np.sin(np.pi * r / float(k.period)) ** 2 / float(k.lengthscale) ** 2)
    c = float(k.location)
    return v * (x[:, None] - c) * (x[None, :] - c)


def test_compile_program_matches_tree_on_grammar_candidates():
    """Every structure the grammar produces to depth 3 (SE / Lin / Per bases, a sample of each level): the flat program
    has GPy's theta count and leaf count, and K evaluated from the program equals K evaluated from the tree."""
    import numpy as np
    from autogpc_b200 import flexible_function as ff, gpckernel as gk, gpy_compat as GPy
    from autogpc_b200.gpcdata import GPCData
    from oracle import gpc_oracle as orc
    rng = np.random.default_rng(5)
    X = rng.uniform(-2, 2, (40, 3))
    y = (X[:, 0] > 0).astype(float).reshape(-1, 1)
    d = GPCData(X, y)
    gk.set_rng(1)
    level = [gk.GPCKernel(ff.NoneKernel(), d, depth=0)]
    checked, seen = 0, set()
    for depth in range(3):
        nxt = []
        for parent in level:
            for c in parent.expand(base_kernels="SE,Lin,Per"):
                name = c.kernel.structure() if hasattr(c.kernel, "structure") else str(c.kernel)
                if name in seen:
                    continue
                seen.add(name)
                nxt.append(c)
        pick = [nxt[i] for i in rng.permutation(len(nxt))[:60]]
        for c in pick:
            kcopy = c.kernel.copy()
            gk.resetGpssParams(kcopy, data=d)
            kern = gk.gpss2gpy(kcopy, data=d)
            prog = GPy.compile_program(kern)
            assert prog.n_theta == len(kern.param_array), name
            assert prog.n_leaves == len(kern.leaves()), name
            oprog = orc.KernelProgram([prog.leaf_type[i] for i in range(prog.n_leaves)],
                                      [prog.leaf_dim[i] for i in range(prog.n_leaves)],
                                      [prog.leaf_off[i] for i in range(prog.n_leaves)], prog.terms(), prog.n_theta)
            Kp = orc.kern_K(oprog, kern.param_array, X)
            Kt = _dense_K(kern, X)
            assert np.allclose(Kp, Kt, rtol=1e-11, atol=1e-13), name
            checked += 1
        level = pick[:8]
    assert checked >= 100
