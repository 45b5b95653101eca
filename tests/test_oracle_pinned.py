"""Independent pin for the oracle's kernel arithmetic (reference call sites gpckernel.py:535-571).

GPy cannot be installed, and the reference holds no numeric known-answer test, so ``oracle/gpc_oracle.py`` is a
restatement.  scikit-learn IS in the image and is independent code: its ``gaussian_process.kernels`` implement the same
covariance functions (RBF = GPy.kern.RBF; ExpSineSquared with length_scale = 2 l = GPy.kern.StdPeriodic's 1/2-form;
ConstantKernel = GPy.kern.Bias; DotProduct on shifted inputs = the LIN extension; Sum / Product = GPy.kern.Add / Prod)
with analytic gradients in log-parameter space.  ``kern_K`` and ``kern_dK`` must agree with them to rounding.
"""
import numpy as np
import pytest

sk = pytest.importorskip("sklearn.gaussian_process.kernels")

from oracle import gpc_oracle as orc
from tests.util import TREES, random_theta, synth

CONFIG_TREES = dict(TREES)
CONFIG_TREES["cfg4:se0xse1+per2+se3"] = ("+", ("*", ("SE", 0), ("SE", 1)), ("PER", 2), ("SE", 3))
CONFIG_TREES["cfg3:se5x(lin2+se3)"] = ("*", ("SE", 5), ("+", ("LIN", 2), ("SE", 3)))
CONFIG_TREES["cfg2:per1xse0+se2"] = ("+", ("*", ("PER", 1), ("SE", 0)), ("SE", 2))


def sklearn_leaf(t, theta, x):
    """(K, [dK/dtheta_j]) of one leaf on the single input column x (n x 1), all of it computed by scikit-learn."""
    if t == orc.LEAF_SE:
        var, ell = theta
        k = sk.ConstantKernel(var) * sk.RBF(ell)
        K, G = k(x, eval_gradient=True)                    # d/d log var, d/d log ell
        return K, [G[:, :, 0] / var, G[:, :, 1] / ell]
    if t == orc.LEAF_PER:
        var, per, ell = theta
        k = sk.ConstantKernel(var) * sk.ExpSineSquared(length_scale=2.0 * ell, periodicity=per)
        K, G = k(x, eval_gradient=True)                    # d/d log var, d/d log length_scale, d/d log periodicity
        return K, [G[:, :, 0] / var, G[:, :, 2] / per, G[:, :, 1] / ell]
    if t == orc.LEAF_CONST:
        (var,) = theta
        K, G = sk.ConstantKernel(var)(x, eval_gradient=True)
        return K, [G[:, :, 0] / var]
    if t == orc.LEAF_LIN:
        var, loc = theta
        k = sk.ConstantKernel(var) * sk.DotProduct(sigma_0=1e-150)
        K, G = k(x - loc, eval_gradient=True)
        # the location gradient has no scikit-learn analogue: -var (x + x' - 2 loc), checked by central differences
        h = 1e-6
        Kp = (sk.ConstantKernel(var) * sk.DotProduct(sigma_0=1e-150))(x - (loc + h))
        Km = (sk.ConstantKernel(var) * sk.DotProduct(sigma_0=1e-150))(x - (loc - h))
        return K, [G[:, :, 0] / var, (Kp - Km) / (2 * h)]
    raise ValueError(t)


def sklearn_program(prog, theta, X):
    """K and dK/dtheta of the composite by the sum / product rules over scikit-learn's leaf matrices."""
    vals, grads = [], []
    for t, d, o in zip(prog.leaf_type, prog.leaf_dim, prog.leaf_off):
        K, G = sklearn_leaf(t, theta[o:o + orc.LEAF_NPARAM[t]], X[:, [d]])
        vals.append(K)
        grads.append(G)
    n = X.shape[0]
    K = np.zeros((n, n))
    dK = [np.zeros((n, n)) for _ in range(prog.n_theta)]
    for term in prog.terms:
        K += np.prod([vals[l] for l in term], axis=0)
        for l in term:
            rest = np.prod([vals[m] for m in term if m != l] + [np.ones((n, n))], axis=0)
            for j, g in enumerate(grads[l]):
                dK[prog.leaf_off[l] + j] += rest * g
    return K, dK


@pytest.mark.parametrize("name", sorted(CONFIG_TREES))
def test_kernel_matrix_and_gradients_match_scikit_learn(name):
    prog = orc.KernelProgram.from_tree(CONFIG_TREES[name])
    X, _ = synth(60, 8, 3)
    theta = random_theta(prog, np.random.default_rng(11))
    K_ref, dK_ref = sklearn_program(prog, theta, X)
    K = orc.kern_K(prog, theta, X)
    assert np.abs(K - K_ref).max() <= 1e-12 * max(1.0, np.abs(K_ref).max())
    dK = orc.kern_dK(prog, theta, X)
    lin_loc = {o + 1 for t, o in zip(prog.leaf_type, prog.leaf_off) if t == orc.LEAF_LIN}
    for j in range(prog.n_theta):
        tol = 1e-7 if j in lin_loc else 1e-12          # the LIN location partial is pinned by central differences only
        assert np.abs(dK[j] - dK_ref[j]).max() <= tol * max(1.0, np.abs(dK_ref[j]).max()), (name, j)


def test_se_ard_product_entirely_in_scikit_learn():
    """configs[0]'s kernel SE0 x SE1 as ONE scikit-learn object (Product of a constant and an anisotropic RBF), value and
    gradient evaluated by scikit-learn's own Product rule -- no combination code of ours in the loop."""
    prog = orc.KernelProgram.from_tree(TREES["se0xse1"])
    X, _ = synth(80, 2, 5)
    theta = np.array([1.7, 0.6, 0.8, 1.9])               # [var0, ell0, var1, ell1]
    k = sk.ConstantKernel(theta[0] * theta[2]) * sk.RBF([theta[1], theta[3]])
    K_ref, G = k(X, eval_gradient=True)                  # d/d log c, d/d log ell0, d/d log ell1
    assert np.abs(orc.kern_K(prog, theta, X) - K_ref).max() <= 1e-13
    dK = orc.kern_dK(prog, theta, X)
    assert np.abs(dK[0] - G[:, :, 0] / theta[0]).max() <= 1e-12
    assert np.abs(dK[2] - G[:, :, 0] / theta[2]).max() <= 1e-12
    assert np.abs(dK[1] - G[:, :, 1] / theta[1]).max() <= 1e-12
    assert np.abs(dK[3] - G[:, :, 2] / theta[3]).max() <= 1e-12


def test_sum_of_rbf_entirely_in_scikit_learn():
    """SE0 + SE1 through scikit-learn's Sum: each RBF is restricted to its dimension by an (effectively) infinite length
    scale on the other one."""
    prog = orc.KernelProgram.from_tree(TREES["se0+se1"])
    X, _ = synth(70, 2, 6)
    theta = np.array([0.9, 1.3, 2.1, 0.5])
    big = 1e150
    k = sk.ConstantKernel(theta[0]) * sk.RBF([theta[1], big]) + sk.ConstantKernel(theta[2]) * sk.RBF([big, theta[3]])
    K_ref, G = k(X, eval_gradient=True)   # log c0, log ell00, log ell01, log c1, log ell10, log ell11
    assert np.abs(orc.kern_K(prog, theta, X) - K_ref).max() <= 1e-13
    dK = orc.kern_dK(prog, theta, X)
    for j, (gi, scale) in enumerate([(0, theta[0]), (1, theta[1]), (3, theta[2]), (5, theta[3])]):
        assert np.abs(dK[j] - G[:, :, gi] / scale).max() <= 1e-12, j


def test_gp_log_marginal_matches_scikit_learn_regressor():
    """The Gaussian part of the oracle's linear algebra (jitchol, log det, quadratic form) against scikit-learn's
    GaussianProcessRegressor.log_marginal_likelihood on the EP pseudo-targets: with the sites frozen, GPy-2016's objective
    is exactly the log marginal of a GP regression with per-point noise 1 / tau_i on targets nu_i / tau_i
    (``nlml_gauss`` of the engine)."""
    gpr = pytest.importorskip("sklearn.gaussian_process")
    prog = orc.KernelProgram.from_tree(TREES["se0xse1"])
    X, y = synth(120, 2, 9)
    theta = np.array([1.4, 0.9, 1.0, 1.2])
    o = orc.score(prog, theta, X, y)
    tau, nu = o["tau"], o["v"]
    k = sk.ConstantKernel(theta[0] * theta[2], constant_value_bounds="fixed") * sk.RBF([theta[1], theta[3]], length_scale_bounds="fixed")
    m = gpr.GaussianProcessRegressor(kernel=k, alpha=1.0 / tau, optimizer=None).fit(X, nu / tau)
    ref = -m.log_marginal_likelihood()
    assert abs(o["nlml_gauss"] - ref) <= 1e-9 * abs(ref)


# ---- the EP recursion against the EXACT evidence and posterior moments on tiny problems ---------------------------
# For the probit likelihood the marginal likelihood is an orthant probability,
#     Z = int prod_i Phi(y_i f_i) N(f | 0, K) df = P(g <= 0),  g ~ N(0, I + Y K Y),  Y = diag(y)
# (write Phi(y_i f_i) = P(eps_i <= y_i f_i) with eps ~ N(0, I) independent of f), which SciPy integrates directly
# (Genz).  EP is an approximation, but for probit GP classification a famously tight one (Kuss & Rasmussen, JMLR 2005:
# log Z_EP within ~1e-2 nats of the truth).  This is not a rounding-level pin; it is an independent check that the
# recursion, the moment matching and the log Z_EP formula of the oracle (EP.inference / _log_Z_tilde as reached from
# gpckernel.py:245-246) compute what they claim to, which no amount of self-consistency between the oracle and the
# CUDA engine can show.
@pytest.mark.parametrize("n,seed,ell,var", [(3, 0, 0.7, 1.0), (4, 1, 1.5, 2.0), (5, 2, 0.4, 0.6), (6, 3, 1.0, 4.0)])
def test_ep_log_evidence_against_the_exact_orthant_probability(n, seed, ell, var):
    from scipy.stats import multivariate_normal
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, 1))
    y = np.where(rng.random(n) > 0.5, 1.0, -1.0)
    prog = orc.KernelProgram.from_tree(("SE", 0))
    K = orc.kern_K(prog, np.array([var, ell]), x)
    C = np.eye(n) + (y[:, None] * K) * y[None, :]
    Z = multivariate_normal(mean=np.zeros(n), cov=C, allow_singular=False).cdf(np.zeros(n))
    exact = float(np.log(Z))
    for run in (lambda: orc.ep_parallel(K, y, eps=1e-10), lambda: orc.ep_sequential(K, y, eps=1e-10, rng=np.random.default_rng(0))):
        res = run()
        logz = orc.ep_log_marginal(K, y, res.tau_t, res.v_t)["log_z"]
        assert abs(logz - exact) < 3e-3, (n, logz, exact)     # measured: 4e-4 ... 1e-3 nats
    # the two schedules share their fixed point
    a, b = orc.ep_parallel(K, y, eps=1e-12), orc.ep_sequential(K, y, eps=1e-12, rng=np.random.default_rng(1))
    assert np.allclose(a.tau_t, b.tau_t, rtol=1e-6, atol=1e-9) and np.allclose(a.v_t, b.v_t, rtol=1e-6, atol=1e-9)


def test_ep_posterior_mean_against_importance_sampling():
    """E[f | y] under the exact posterior, by self-normalised importance sampling from the prior (n = 4, 400 k draws:
    Monte-Carlo error ~3e-3), against the mean of EP's Gaussian approximation -- known to be accurate to a few 1e-3."""
    rng = np.random.default_rng(5)
    n = 4
    x = rng.normal(size=(n, 1))
    y = np.array([1.0, -1.0, 1.0, 1.0])
    prog = orc.KernelProgram.from_tree(("SE", 0))
    K = orc.kern_K(prog, np.array([1.5, 0.8]), x)
    f = rng.multivariate_normal(np.zeros(n), K, size=400_000)
    from scipy.special import ndtr
    w = np.prod(ndtr(f * y[None, :]), axis=1)
    mean_exact = (w[:, None] * f).sum(0) / w.sum()
    res = orc.ep_parallel(K, y, eps=1e-10)
    _, _, mu, _ = orc.posterior_from_sites(K, res.tau_t, res.v_t)
    assert np.allclose(mu, mean_exact, atol=8e-3), (mu, mean_exact)      # measured: 1.5e-3 (Monte-Carlo error included)
