"""Host logic on CPU with the oracle-backed engine stand-in: lock-step L-BFGS-B, GPy shim, GPCKernel.train,
GPCSearch (reference flow, gpcsearchtest.py:11-21 data), BIC wiring."""
import numpy as np
import pytest
from scipy.optimize import fmin_l_bfgs_b

from autogpc_b200 import engine as eng_mod
from autogpc_b200 import flexible_function as ff
from autogpc_b200 import gpckernel, optim
from autogpc_b200 import gpy_compat as GPy
from autogpc_b200.gpcdata import GPCData
from autogpc_b200.gpcsearch import GPCSearch, bestKernels1D
from oracle import gpc_oracle as orc
from tests.fake_engine import OracleEngine
from tests.util import synth


@pytest.fixture()
def fake():
    e = OracleEngine()
    eng_mod.set_default_engine(e)
    yield e
    eng_mod.set_default_engine(None)


def step_data(n=60, seed=0):
    """gpcsearchtest.py:11-13: X ~ U(1, 10), the last quarter labelled 1 -- here sorted so the step is learnable."""
    rng = np.random.default_rng(seed)
    X = np.sort(rng.uniform(1, 10, (n, 1)), axis=0)
    Y = np.zeros((n, 1))
    Y[-n // 4:] = 1
    return GPCData(X, Y, seed=seed)


def test_lockstep_lbfgsb_equals_individual_runs():
    rs = [np.random.default_rng(i) for i in range(5)]
    As = [np.diag(r.uniform(0.5, 4, 3)) for r in rs]
    bs = [r.normal(size=3) for r in rs]
    f = lambda i, x: (0.5 * x @ As[i] @ x - bs[i] @ x + 0.1 * np.sum(x ** 4), As[i] @ x - bs[i] + 0.4 * x ** 3)
    calls = []

    def eval_batch(idx, xs):
        calls.append(len(idx))
        return [f(i, x) for i, x in zip(idx, xs)]

    out = optim.lbfgsb_lockstep(eval_batch, [np.ones(3) * (i + 1) for i in range(5)])
    for i, (x, fv, info) in enumerate(out):
        xr, fr, ir = fmin_l_bfgs_b(lambda x: f(i, x), np.ones(3) * (i + 1), m=10, factr=1e7, pgtol=1e-5)
        assert np.array_equal(x, xr) and fv == fr and info["funcalls"] == ir["funcalls"]
    assert calls[0] == 5 and max(calls) == 5          # requests really are batched


def test_gpclassification_shim_matches_oracle_optimize(fake):
    X, y = synth(90, 2, 3)
    d = GPCData(X, y.reshape(-1, 1))
    k = ff.ProductKernel([ff.SqExpKernel(0, 1.0, 1.0), ff.SqExpKernel(1, 1.5, 0.8)])
    m = GPy.models.GPClassification(X, y.reshape(-1, 1), kernel=gpckernel.gpss2gpy(k, data=d))
    prog = orc.KernelProgram.from_tree(("*", ("SE", 0), ("SE", 1)))
    th0 = m.kern.param_array.copy()
    assert -m.log_likelihood() == pytest.approx(orc.score(prog, th0, X, y, damping=0.0)["nlml"], rel=1e-12)   # ctor = 1 inference
    m.optimize()
    lo, hi = d.getLengthscaleBounds()
    tr = orc.Transforms(np.array([0, 1, 0, 1]), np.array([0, lo[0], 0, lo[1]]), np.array([0, hi[0], 0, hi[1]]))
    ref = orc.optimize(prog, th0, X, y, tr, damping=0.0)
    assert -m.log_likelihood() == pytest.approx(ref["nlml"], rel=1e-6)   # warm- vs cold-started final EP (eps 1e-6)
    assert np.allclose(m.kern.param_array, ref["theta"], rtol=1e-5)
    Phi, _ = m.predict(X[:7])
    assert Phi.shape == (7, 1) and np.all((Phi > 0) & (Phi < 1))
    assert m.bic == pytest.approx(2 * -m.log_likelihood() + 3 * np.log(90))                      # p_eff = 3


def test_train_semantics_median_fold_and_mean_error(fake):
    gpckernel.set_rng(5)
    d = step_data(50, 1)
    k = gpckernel.GPCKernel(ff.SqExpKernel(dimension=0), d, depth=1)
    calls0 = fake.calls
    k.train(mode="full", n_folds=5)
    assert k.model is not None and isinstance(k.kernel, ff.SqExpKernel)
    assert len(k.foldErrors) == 5 and k.error() == pytest.approx(np.mean(k.foldErrors))
    assert k.model.num_data == len(d.kFoldIndices(5)[2][0])           # results[len // 2] in fold order
    assert k.getNLML() == pytest.approx(-k.model.log_likelihood())
    assert k.bic() == pytest.approx(2 * k.getNLML() + 2 * np.log(k.model.num_data))
    # all five folds advance in lock-step: far fewer engine calls than items
    assert fake.items / (fake.calls - calls0) > 2.0
    with pytest.raises(NotImplementedError):
        k.train(mode="svgp")
    assert gpckernel.GPCKernel(ff.ConstKernel(), d).error() == pytest.approx(min(d.Y.sum(), 50 - d.Y.sum()) / 50.0)
    mis = k.misclassifiedPoints()
    assert mis["X"].shape[0] == mis["Y"].shape[0]                      # gpckerneltest.py:121


def test_search_runs_and_is_deterministic(fake):
    def run(criterion):
        gpckernel.set_rng(11)
        d = step_data(48, 2)
        s = GPCSearch(d, base_kernels="SE,Per", max_depth=2, beam_width=1, criterion=criterion, n_folds=3, max_evals=40)
        best, best1d = s.search()
        return s, best, best1d

    s1, best, best1d = run("error")
    s2, best2, _ = run("error")
    assert [k.kernel.structure() for k in best] == [k.kernel.structure() for k in best2]
    assert [[k.kernel.structure() for k in lvl] for lvl in s1.history] == [[k.kernel.structure() for k in lvl] for lvl in s2.history]
    assert isinstance(best[0].kernel, ff.NoneKernel) and len(best) >= 2
    assert best[1].error() < 0.25 and len(best1d) == 1
    assert sorted(k.kernel.structure() for k in s1.history[0]) == ["Per0", "SE0"]
    sb, bestb, _ = run("bic")
    assert np.isfinite(bestb[-1].bic()) and [k.bic() for k in sb.history[0]] == sorted(k.bic() for k in sb.history[0])
    base = s1.baseline()
    assert isinstance(base.kernel, ff.ConstKernel) and base.error() == pytest.approx(0.25)


def test_add_and_tosummands_score_unoptimised_models(fake):
    gpckernel.set_rng(3)
    X, y = synth(70, 2, 4)
    d = GPCData(X, y.reshape(-1, 1))
    a = gpckernel.GPCKernel(ff.SqExpKernel(0), d)
    b = gpckernel.GPCKernel(ff.SqExpKernel(1), d)
    c = a.add(b)
    assert c.kernel.structure() == "SE0+SE1" and c.model.num_data == 70 and np.isfinite(c.getNLML())
    assert 0 <= c.error() <= 1
    parts = gpckernel.GPCKernel(ff.ProductKernel([ff.SqExpKernel(0), ff.SumKernel([ff.SqExpKernel(1), ff.ConstKernel()])]), d).toSummands()
    assert sorted(p.kernel.structure() for p in parts) == ["C*SE0", "SE0*SE1"]   # additive form is canonical, not simplified
    assert all(p.model is not None for p in parts)
    one = gpckernel.GPCKernel(ff.SqExpKernel(0), d)
    one.train(mode="full", n_folds=2, max_evals=20)
    assert one.monotonicity() in (-1, 0, 1)
    assert bestKernels1D([a.add(b)][:0]) == []


def test_experiment_shim_runs_search_and_baseline(fake):
    """gpcexperiment.py:50-61 flow (data -> search -> baseline) on arrays instead of pods downloads."""
    from autogpc_b200.gpcexperiment import GPCExperiment
    gpckernel.set_rng(2)
    X, y = synth(60, 8, 9)
    out = GPCExperiment(base_kernels="SE", n_folds=2, max_evals=15, verbose=False).pima(X, y, depth=1, width=1)
    assert out["data"].getDim() == 4 and out["data"].getNum() == 60        # dims=(1,5,6,7) as in the reference
    assert len(out["best"]) == 2 and out["best"][1].kernel.structure().startswith("SE")
    assert isinstance(out["baseline"].kernel, ff.ConstKernel) and out["seconds"] > 0


def test_candidate_beyond_program_limits_ranks_last_instead_of_aborting(fake):
    """(SE0+SE1)*(SE0+SE1)*(SE0+SE1)*(SE0+SE1)*(SE0+SE1) distributes to 32 product terms (engine limit 16): the candidate
    is skipped with a warning and scores inf; the other candidate of the same call trains normally."""
    gpckernel.set_rng(2)
    X, y = synth(60, 2, 1)
    d = GPCData(X, y.reshape(-1, 1))
    s2 = lambda: ff.SumKernel([ff.SqExpKernel(0, 1.0, 1.0), ff.SqExpKernel(1, 1.0, 1.0)])
    big = gpckernel.GPCKernel(ff.ProductKernel([s2() for _ in range(5)]), d, depth=5)
    ok = gpckernel.GPCKernel(ff.SqExpKernel(0, 1.0, 1.0), d, depth=1)
    with pytest.warns(UserWarning, match="skipped"):
        gpckernel.train_kernels([big, ok], mode="full", n_folds=2, max_evals=15)
    assert big.model is None and big.error() == float("inf") and big.bic() == float("inf")
    assert ok.model is not None and np.isfinite(ok.error()) and np.isfinite(ok.bic())


def test_cumulate_additive_kernels_matches_the_serial_reference_loop(fake):
    """gpcreport.py:657-680 transcribed with one ``add`` (one model) per candidate, against the batched version."""
    gpckernel.set_rng(4)
    X, y = synth(90, 3, 6)
    d = GPCData(X, y.reshape(-1, 1))
    parent = gpckernel.GPCKernel(ff.SumKernel([ff.SqExpKernel(0, 1.0, 1.2), ff.SqExpKernel(1, 1.5, 0.7),
                                               ff.ProductKernel([ff.SqExpKernel(2, 0.8, 1.0), ff.SqExpKernel(0, 1.0, 2.0)])]),
                                 d, depth=3)
    summands = parent.toSummands()
    assert len(summands) == 3

    def serial(summands):
        terms = sorted(summands, key=lambda k: k.error(), reverse=True)
        ker = terms.pop()
        cum = ker
        kers, cums = [ker], [cum]
        while len(terms) > 0:
            terms = sorted(terms, key=lambda k: cum.add(k).error(), reverse=True)
            ker = terms.pop()
            cum = cum.add(ker)
            kers.append(ker)
            cums.append(cum)
        return kers, cums

    gpckernel.set_rng(11)      # GPCKernel.__init__ re-draws hyper-parameters: same stream for both versions
    k1, c1 = serial(list(summands))
    gpckernel.set_rng(11)
    k2, c2 = gpckernel.cumulateAdditiveKernels(list(summands))
    assert [k.kernel.structure() for k in k1] == [k.kernel.structure() for k in k2]
    assert [c.kernel.structure() for c in c1] == [c.kernel.structure() for c in c2]
    assert np.allclose([c.error() for c in c1], [c.error() for c in c2])
    assert np.allclose([c.getNLML() for c in c1], [c.getNLML() for c in c2], rtol=1e-9)


def test_optimum_is_taken_from_the_evaluation_cache(fake):
    """L-BFGS-B returns an iterate it has evaluated: optimize_models restores that evaluation instead of scoring theta*
    once more (engine items == objective evaluations), and the restored state is the model's state at theta*."""
    X, y = synth(80, 2, 9)
    d = GPCData(X, y.reshape(-1, 1))
    did = gpckernel.device_dataset(d)
    rows = np.arange(80, dtype=np.int32)
    ms = []
    for v in (1.0, 2.0, 0.5):
        k = ff.ProductKernel([ff.SqExpKernel(0, v, 1.0), ff.SqExpKernel(1, 1.5, 0.8)])
        ms.append(GPy.models.GPClassification(X, y.reshape(-1, 1), kernel=gpckernel.gpss2gpy(k, data=d), _dataset=did,
                                              _rows=rows, _infer=False))
    items0 = fake.items
    optim.optimize_models(ms, max_evals=30)
    assert fake.items - items0 == sum(m.optimization_info["evals"] for m in ms)
    prog = orc.KernelProgram.from_tree(("*", ("SE", 0), ("SE", 1)))
    for m in ms:
        o = orc.score(prog, m.kern.param_array, X, y, damping=0.0)
        # warm-started (cached) vs cold evaluation at GPy's epsilon = 1e-6: both within ~1e-6 of the tight fixed point
        assert -m.log_likelihood() == pytest.approx(o["nlml"], rel=5e-6)
        assert m.bic == pytest.approx(o["bic"], rel=5e-6)
        assert m._tau.shape == (80,) and np.all(np.isfinite(m._grad))


def test_failed_evaluations_are_tolerated_like_paramz(fake):
    """A failed evaluation (not-PD / non-finite) is reported to L-BFGS-B as a huge objective with the gradient of the last
    successful evaluation (paramz: the model's current gradient), it is counted per model, the count is reset by the next
    success (paramz gives up after more than 10 CONSECUTIVE failures), and the model is left at its best evaluated point."""
    class Flaky(OracleEngine):
        n_fail = 0

        def score_batch(self, *a, **kw):
            r = super().score_batch(*a, **kw)
            if self.calls > 1 and self.calls % 3 == 0:
                r["status"][:] = orc.ST_NOT_PD
                r["nlml"][:] = np.inf
                self.n_fail += 1
            return r
    e = Flaky()
    eng_mod.set_default_engine(e)
    try:
        X, y = synth(60, 2, 5)
        d = GPCData(X, y.reshape(-1, 1))
        k = ff.ProductKernel([ff.SqExpKernel(0, 1.0, 1.0), ff.SqExpKernel(1, 1.5, 0.8)])
        m = GPy.models.GPClassification(X, y.reshape(-1, 1), kernel=gpckernel.gpss2gpy(k, data=d))
        f0 = m._nlml
        optim.optimize_models([m], max_evals=40)
        assert e.n_fail >= 1 and getattr(m, "optimization_error", None) is None
        assert np.isfinite(m._nlml) and m._nlml <= f0 and m.status == 0
    finally:
        eng_mod.set_default_engine(None)


def test_device_optimizer_selection_rules(monkeypatch):
    """Which driver ``optimize_models`` picks: the device loop wherever the engine has it and every model runs EP, unless
    AGPC_DEVICE_OPT=0; an engine without ``optimize_batch`` (the oracle-backed test double) or a Laplace model stays on the
    SciPy driver; AGPC_DEVICE_OPT=1 with a Laplace model is an error rather than a silent fallback."""
    from types import SimpleNamespace
    from autogpc_b200 import optim

    with_dev = SimpleNamespace(optimize_batch=lambda *a, **k: None)
    without = SimpleNamespace()
    ep = lambda e: SimpleNamespace(engine=e, ep=dict(epsilon=1e-6, damping=0.0, max_sweeps=100))
    lap = lambda e: SimpleNamespace(engine=e, ep=dict(epsilon=1e-10, damping=0.0, max_sweeps=50, inference="laplace"))
    monkeypatch.delenv("AGPC_DEVICE_OPT", raising=False)
    assert optim.device_optimizer_default([ep(with_dev), ep(with_dev)]) is True
    assert optim.device_optimizer_default([ep(without)]) is False
    assert optim.device_optimizer_default([ep(with_dev), lap(with_dev)]) is False
    monkeypatch.setenv("AGPC_DEVICE_OPT", "0")
    assert optim.device_optimizer_default([ep(with_dev)]) is False
    monkeypatch.setenv("AGPC_DEVICE_OPT", "1")
    assert optim.device_optimizer_default([ep(with_dev)]) is True
    assert optim.device_optimizer_default([ep(without)]) is False
    import pytest
    with pytest.raises(ValueError):
        optim.device_optimizer_default([lap(with_dev)])
