"""Pins the oracle (CPU, no GPU): internal-consistency checks that stand in for the golden vectors the reference
does not have ("parity unpinned"), plus the committed fixtures under tests/golden/."""
import os

import numpy as np
import pytest

from oracle import gpc_oracle as orc
from tests.util import TREES, random_theta, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def case():
    X, y = synth(160, 4, 0)
    prog = orc.KernelProgram.from_tree(TREES["mix"])
    theta = random_theta(prog, np.random.default_rng(1))
    return X, y, prog, theta


def test_kernel_gradients_fd(case):
    X, _, prog, theta = case
    dK = orc.kern_dK(prog, theta, X[:40])
    for j in range(prog.n_theta):
        h = 1e-6 * max(1.0, abs(theta[j]))
        tp, tm = theta.copy(), theta.copy()
        tp[j] += h
        tm[j] -= h
        fd = (orc.kern_K(prog, tp, X[:40]) - orc.kern_K(prog, tm, X[:40])) / (2 * h)
        assert np.allclose(dK[j], fd, rtol=1e-6, atol=1e-8), j


def test_distributed_form_equals_tree():
    """(SE0 + C) x LIN2 evaluated as a tree equals the distributed sum-of-products program."""
    X, _ = synth(30, 3, 2)
    prog = orc.KernelProgram.from_tree(TREES["lin2x(se0+c)"])
    th = np.array([0.7, 0.3, 1.4, 0.9, 0.5])       # LIN[v, c], SE[v, l], CONST[v]
    lin = th[0] * np.outer(X[:, 2] - th[1], X[:, 2] - th[1])
    se = th[2] * np.exp(-0.5 * (X[:, 0][:, None] - X[:, 0][None, :]) ** 2 / th[3] ** 2)
    assert np.allclose(orc.kern_K(prog, th, X), lin * (se + th[4]), rtol=1e-14)


def test_moments_match_naive_gpy():
    rng = np.random.default_rng(0)
    for _ in range(200):
        tau, v, y = np.exp(rng.normal(0, 1)), rng.normal(0, 2), int(rng.integers(0, 2))
        lz, mu, s2, z = orc.moments_match(1.0 if y == 1 else -1.0, tau, v)
        if abs(z) > 8:
            continue
        Z0, mu0, s20 = orc.moments_match_gpy_naive(y, tau, v)
        assert np.isclose(np.exp(lz), Z0, rtol=1e-12) and np.isclose(mu, mu0, rtol=1e-10) and np.isclose(s2, s20, rtol=1e-9)
    # tails stay finite where the naive formula has already clamped / overflowed
    lz, mu, s2, _ = orc.moments_match(-1.0, 1.0, 60.0)
    assert np.isfinite(lz) and np.isfinite(mu) and s2 > 0


def test_sequential_and_parallel_ep_share_the_fixed_point(case):
    X, y, prog, theta = case
    a = orc.score(prog, theta, X, y, eps=1e-14, max_sweeps=500)
    b = orc.score(prog, theta, X, y, eps=1e-14, max_sweeps=500, ep="sequential", rng=np.random.default_rng(3))
    assert abs(a["nlml"] - b["nlml"]) <= 1e-11 * abs(a["nlml"])
    assert np.allclose(a["grad"], b["grad"], rtol=1e-6, atol=1e-8)
    # GPy's default epsilon: log Z (stationary in the sites) already agrees to the north_star tolerance
    c = orc.score(prog, theta, X, y)
    d = orc.score(prog, theta, X, y, ep="sequential", rng=np.random.default_rng(4))
    assert abs(c["nlml"] - d["nlml"]) <= 1e-6 * abs(c["nlml"])


def test_two_log_z_forms_agree(case):
    X, y, prog, theta = case
    r = orc.score(prog, theta, X, y, eps=1e-12, max_sweeps=300)
    assert np.isclose(-orc.ep_log_marginal_gpy18(r["K"], y, r["tau"], r["v"]), r["nlml"], rtol=1e-12)


def test_nlml_gradient_fd(case):
    X, y, prog, theta = case
    g = orc.score(prog, theta, X, y, eps=1e-14, max_sweeps=500)["grad"]
    for j in range(prog.n_theta):
        h = 1e-5
        tp, tm = theta.copy(), theta.copy()
        tp[j] += h
        tm[j] -= h
        fd = (orc.score(prog, tp, X, y, eps=1e-14, max_sweeps=500)["nlml"] -
              orc.score(prog, tm, X, y, eps=1e-14, max_sweeps=500)["nlml"]) / (2 * h)
        assert np.isclose(g[j], fd, rtol=2e-6, atol=1e-7), (j, g[j], fd)


def test_frozen_sites_gradient_is_gaussian_objective_gradient(case):
    """2016-GPy semantics: with sites frozen the returned gradient is d(-log N(mu~ | 0, K + S~^-1))/d theta."""
    X, y, prog, theta = case
    r0 = orc.score(prog, theta, X, y)
    f = lambda th: orc.score(prog, th, X, y, tau0=r0["tau"], v0=r0["v"], frozen_sites=True)
    g = f(theta)["grad"]
    for j in (0, 3, 7):
        h = 1e-6
        tp, tm = theta.copy(), theta.copy()
        tp[j] += h
        tm[j] -= h
        fd = (f(tp)["nlml_gauss"] - f(tm)["nlml_gauss"]) / (2 * h)
        assert np.isclose(g[j], fd, rtol=1e-5, atol=1e-7)


def test_transforms_and_bic():
    tr = orc.Transforms(np.array([0, 1, 2]), np.array([0.0, 0.5, 0.0]), np.array([0.0, 4.0, 0.0]))
    th = np.array([0.3, 7.0, -2.0])
    x = tr.to_x(th)
    assert np.allclose(tr.to_theta(x), [0.3, 2.25, -2.0])              # out-of-bounds start -> midpoint
    h = 1e-6
    for j in range(3):
        e = np.zeros(3)
        e[j] = h
        fd = (tr.to_theta(x + e)[j] - tr.to_theta(x - e)[j]) / (2 * h)
        assert np.isclose(tr.gradfactor(tr.to_theta(x))[j], fd, rtol=1e-6)
    assert orc.bic(10.0, 3, 100) == pytest.approx(20.0 + 3 * np.log(100))
    prog = orc.KernelProgram.from_tree(TREES["se0xse1+per3+se2"])
    assert orc.n_effective_params(prog) == 3 + 3 + 2                   # product shares one scale


def test_predict_against_dense_formula(case):
    X, y, prog, theta = case
    r = orc.score(prog, theta, X[:120], y[:120])
    p, mu, var = orc.predict(prog, theta, X[:120], X[120:], r["tau"], r["v"])
    K = r["K"]
    Ks = orc.kern_K(prog, theta, X[:120], X[120:])
    Kss = orc.kern_K(prog, theta, X[120:])
    Sinv = np.diag(1.0 / r["tau"])
    mu2 = Ks.T @ np.linalg.solve(K + Sinv, r["v"] / r["tau"])
    var2 = np.diag(Kss - Ks.T @ np.linalg.solve(K + Sinv, Ks))
    assert np.allclose(mu, mu2, rtol=1e-8) and np.allclose(var, var2, rtol=1e-7, atol=1e-10)
    assert np.all((p > 0) & (p < 1))


@pytest.mark.parametrize("name", sorted(f[:-4] for f in os.listdir(GOLD) if f.endswith(".npz")))
def test_golden_fixtures(name):
    """Regression pin: the oracle reproduces the committed vectors (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=True)
    prog = orc.KernelProgram.from_tree(eval(str(g["tree"])))
    r = orc.score(prog, g["theta"], g["X"], g["y"])
    assert r["sweeps"] == int(g["sweeps"]) and r["status"] == int(g["status"])
    assert np.isclose(r["nlml"], float(g["nlml"]), rtol=1e-10)
    assert np.isclose(r["bic"], float(g["bic"]), rtol=1e-10)
    assert np.allclose(r["grad"], g["grad"], rtol=1e-7, atol=1e-9)
    assert np.allclose(r["tau"], g["tau"], rtol=1e-8, atol=1e-12)
    p, _, _ = orc.predict(prog, g["theta"], g["X"], g["Xstar"], r["tau"], r["v"])
    assert np.allclose(p, g["pstar"], rtol=1e-9, atol=1e-12)


# ---- Laplace approximation (R&W alg. 3.1 / 5.1) -------------------------------------------------------------------
@pytest.mark.parametrize("name", ["se0xse1", "se0xse1+per3+se2", "mix"])
def test_laplace_gradient_fd_and_site_form(name):
    from tests.util import TREES, random_theta, synth
    X, y = synth(90, 4, 3)
    prog = orc.KernelProgram.from_tree(TREES[name])
    th = random_theta(prog, np.random.default_rng(2))
    r = orc.score_laplace(prog, th, X, y)
    assert r["status"] == 0 and r["sweeps"] <= 15
    g = np.zeros_like(th)
    for j in range(len(th)):
        h = 1e-5 * max(1.0, abs(th[j]))
        tp, tm = th.copy(), th.copy()
        tp[j] += h
        tm[j] -= h
        g[j] = (orc.score_laplace(prog, tp, X, y)["nlml"] - orc.score_laplace(prog, tm, X, y)["nlml"]) / (2 * h)
    assert np.all(np.abs(g - r["grad"]) <= 1e-6 * np.maximum(np.abs(g), 1e-3 * np.abs(g).max()))
    # the mode is a fixed point of the Newton map, and the (tau = W, v = W f + grad) site form reproduces the posterior
    e = orc.ep_log_marginal(r["K"], y, r["tau"], r["v"])
    assert np.allclose(e["mu"], r["f"], atol=1e-9) and np.allclose(e["alpha"], r["alpha"], atol=1e-9)
    assert np.allclose(e["sig_diag"], r["sig_diag"], rtol=1e-9)
    # Laplace and EP approximate the same marginal likelihood: same ball park, Laplace (mode-based) is the larger NLML
    ep = orc.score(prog, th, X, y, damping=0.0)
    assert 0 < r["nlml"] - ep["nlml"] < 0.05 * ep["nlml"]


def test_probit_derivatives_fd():
    f = np.linspace(-4, 4, 41)
    for ypm in (1.0, -1.0):
        lp, d1, W, d3 = orc.probit_derivatives(ypm, f)
        h = 1e-5
        lpp, d1p, Wp, _ = orc.probit_derivatives(ypm, f + h)
        lpm, d1m, Wm, _ = orc.probit_derivatives(ypm, f - h)
        assert np.allclose((lpp - lpm) / (2 * h), d1, rtol=1e-7, atol=1e-9)
        assert np.allclose((d1p - d1m) / (2 * h), -W, rtol=1e-7, atol=1e-9)
        assert np.allclose(-(Wp - Wm) / (2 * h), d3, rtol=1e-6, atol=1e-8)


def test_parallel_ep_posterior_reuse_rule():
    """oracle.ep_parallel: sites whose residual is already EP_REUSE_FACTOR below the tolerance are returned untouched
    (kept = True, the factorisation at those sites is final); otherwise the converged update is applied as in GPy."""
    X, y = synth(150, 2, 3)
    prog = orc.KernelProgram.from_tree(TREES["se0xse1"])
    K = orc.kern_K(prog, np.array([1.2, 0.9, 1.0, 1.1]), X)
    tight = orc.ep_parallel(K, y, eps=1e-20, max_sweeps=400, damping=0.0)
    r = orc.ep_parallel(K, y, eps=1e-6, damping=0.0, tau0=tight.tau_t, v0=tight.v_t)
    assert r.converged and r.kept and r.sweeps == 1
    assert np.array_equal(r.tau_t, tight.tau_t) and np.array_equal(r.v_t, tight.v_t)
    cold = orc.ep_parallel(K, y, eps=1e-6, damping=0.0)
    assert cold.converged and cold.sweeps >= 2
    # either way the returned sites satisfy the tolerance, and the log marginal agrees with the fixed point to 1e-8
    lt = orc.ep_log_marginal(K, y, tight.tau_t, tight.v_t)["log_z"]
    lc = orc.ep_log_marginal(K, y, cold.tau_t, cold.v_t)["log_z"]
    assert abs(lc - lt) <= 1e-8 * abs(lt)
