"""CPU oracle for the AutoGPC candidate-scoring hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker* for the CUDA engine in ``autogpc_b200/csrc``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product path never does: ``autogpc_b200`` fails loudly when its CUDA library is missing.

PARITY UNPINNED for the EP recursion (no golden vector of the reference or of GPy exists to pin it to rounding); the
kernel functions (kern_K, kern_dK) and the Gaussian log-marginal are pinned against scikit-learn (independent code) in
``tests/test_oracle_pinned.py``, which also anchors the EP recursion on ground truth: log Z_EP against the exact probit
evidence (an orthant probability, integrated by SciPy) and EP's posterior mean against importance sampling, on n = 3 .. 6
problems (4e-4 .. 1e-3 nats: EP's own approximation error).  The reference (``/root/reference``,
charles92/autogpc) contains none of this arithmetic: it
delegates to GPy (SheffieldML/GPy, un-vendored and un-pinned: ``README.md:14-19``; era ~GPy 0.8-1.0, Python 2.7),
paramz and SciPy, none of which exist in the build image, and none of the reference's own tests asserts a numerical
value (``gpckerneltest.py:88-107,121`` are structural).  What follows is therefore a NumPy/SciPy FP64 restatement of
GPy's *published* algorithms, anchored on the reference's call sites:

  ==========================  =========================================================================
  reference call site          restated here as
  ==========================  =========================================================================
  gpckernel.py:514-574         kernel tree -> theta order / variance = sf**2 / active dim  (KernelProgram)
  gpckernel.py:245 (ctor)      GPy ``EP.inference``: K, EP sites, log-marginal, dL_dK      (score)
  gpckernel.py:246 optimize    paramz transforms + SciPy L-BFGS-B                          (optimize)
  gpckernel.py:371             ``log_likelihood``                                          (score()['log_z'])
  gpckernel.py:832             ``predict`` -> probit predictive mean                        (predict)
  flexible_function.py:685     BIC = 2 nll + p_eff ln n                                    (bic)
  ==========================  =========================================================================

Self-consistency checks that stand in for the missing golden vectors live in ``tests/test_oracle.py``:
finite-difference gradients, sequential-EP (GPy-literal site loop) == parallel-EP fixed point, the Rasmussen &
Williams (3.65)+(3.73) log Z_EP == the GPy>=1.8 ``_ep_marginal + log_Z_tilde`` form, and the erfcx-stabilised probit
moments == the naive GPy formulas for |z| < 8.

Formulas marked [GPy] restate SheffieldML/GPy ``kern/src/{rbf,stationary,standard_periodic,static,add,prod}.py``,
``inference/latent_function_inference/expectation_propagation.py``, ``likelihoods/bernoulli.py`` and
``util/linalg.py``; [paramz] restates ``paramz/transformations.py``; [R&W] is Rasmussen & Williams, GPML, ch. 3.6.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy import linalg as sla
from scipy import special as sps

# leaf opcodes -- shared with include/autogpc_b200.h (AGPC_LEAF_*)
LEAF_SE, LEAF_PER, LEAF_LIN, LEAF_CONST = 0, 1, 2, 3
LEAF_NPARAM = {LEAF_SE: 2, LEAF_PER: 3, LEAF_LIN: 2, LEAF_CONST: 1}

# per-item status codes -- shared with include/autogpc_b200.h (AGPC_ST_*)
ST_OK, ST_JITTER, ST_NOT_PD, ST_EP_NOT_CONVERGED, ST_NONFINITE = 0, 1, 2, 3, 4


# --------------------------------------------------------------------------------------------------------------
# kernel program: sum over terms of products over leaves (the distributed form of the GPy Add/Prod tree)
# --------------------------------------------------------------------------------------------------------------
@dataclass
class KernelProgram:
    """Flat encoding of a composite kernel.

    ``leaf_type[l]``/``leaf_dim[l]``/``leaf_off[l]``: opcode, input dimension and offset into theta of leaf ``l``;
    leaves are numbered in the depth-first order of ``gpss2gpy`` (gpckernel.py:514-574) so theta is laid out exactly
    as GPy's ``param_array``: RBF [variance, lengthscale]; StdPeriodic [variance, period, lengthscale];
    Bias [variance]; Linear (extension, flexible_function.py:1284-1343) [variance, location].
    ``terms``: K = sum_t prod_{l in terms[t]} k_l.  A GPy ``Prod`` over an ``Add`` is distributed, so a leaf may
    appear in several terms; its parameters are still single entries of theta.
    """
    leaf_type: List[int]
    leaf_dim: List[int]
    leaf_off: List[int]
    terms: List[List[int]]
    n_theta: int

    @staticmethod
    def from_tree(tree) -> "KernelProgram":
        """tree: nested tuples ('SE', d) | ('PER', d) | ('LIN', d) | ('CONST',) | ('+', a, b, ...) | ('*', a, ...)."""
        lt, ld, lo = [], [], []
        off = 0
        names = {"SE": LEAF_SE, "PER": LEAF_PER, "LIN": LEAF_LIN, "CONST": LEAF_CONST}

        def walk(node):
            nonlocal off
            op = node[0]
            if op in names:
                idx = len(lt)
                lt.append(names[op])
                ld.append(int(node[1]) if len(node) > 1 else 0)
                lo.append(off)
                off += LEAF_NPARAM[names[op]]
                return [[idx]]
            parts = [walk(ch) for ch in node[1:]]
            if op == "+":
                return [t for p in parts for t in p]
            if op == "*":
                acc = [[]]
                for p in parts:
                    acc = [a + t for a in acc for t in p]
                return acc
            raise ValueError("unknown node %r" % (op,))

        terms = walk(tree)
        return KernelProgram(lt, ld, lo, terms, off)


def _leaf_values(prog: KernelProgram, theta, X, X2, want_grad: bool):
    """Per-leaf kernel matrix and its partials w.r.t. that leaf's own parameters.  [GPy] kern.K / gradients."""
    vals, grads = [], []
    for t, d, o in zip(prog.leaf_type, prog.leaf_dim, prog.leaf_off):
        if t == LEAF_CONST:
            var = theta[o]
            k = np.full((X.shape[0], X2.shape[0]), var)                       # [GPy] static.py Bias.K
            g = [np.ones_like(k)] if want_grad else None
        else:
            x, x2 = X[:, d][:, None], X2[:, d][None, :]
            r = x - x2
            if t == LEAF_SE:
                var, ell = theta[o], theta[o + 1]
                e = np.exp(-0.5 * (r / ell) ** 2)                              # [GPy] rbf.py K_of_r
                k = var * e
                g = [e, k * r * r / ell ** 3] if want_grad else None            # stationary.py update_gradients_full
            elif t == LEAF_PER:
                var, per, ell = theta[o], theta[o + 1], theta[o + 2]
                b = np.pi * r / per
                s, c = np.sin(b), np.cos(b)
                e = np.exp(-0.5 * (s / ell) ** 2)                              # [GPy] standard_periodic.py K
                k = var * e
                g = [e, k * s * c * (b / per) / ell ** 2, k * s * s / ell ** 3] if want_grad else None
            elif t == LEAF_LIN:
                var, loc = theta[o], theta[o + 1]
                k = var * (x - loc) * (x2 - loc)                               # flexible_function.py:1284 (covLinear)
                g = [(x - loc) * (x2 - loc), -var * (x + x2 - 2.0 * loc)] if want_grad else None
            else:
                raise ValueError(t)
        vals.append(k)
        grads.append(g)
    return vals, grads


def kern_K(prog: KernelProgram, theta, X, X2=None):
    """K(X, X2) for the composite kernel.  [GPy] add.py / prod.py K."""
    X2 = X if X2 is None else X2
    vals, _ = _leaf_values(prog, np.asarray(theta, float), X, X2, False)
    K = np.zeros((X.shape[0], X2.shape[0]))
    for term in prog.terms:
        p = np.ones_like(K)
        for l in term:
            p = p * vals[l]
        K += p
    return K


def kern_dK(prog: KernelProgram, theta, X):
    """Materialised dK/dtheta_j, j = 0..P-1 (list of n x n arrays).  [GPy] Prod/Add.update_gradients_full chain."""
    theta = np.asarray(theta, float)
    vals, grads = _leaf_values(prog, theta, X, X, True)
    out = [None] * prog.n_theta
    for l in range(len(prog.leaf_type)):
        coef = np.zeros((X.shape[0], X.shape[0]))
        for term in prog.terms:
            if l in term:
                p = np.ones_like(coef)
                for m in term:
                    if m != l:
                        p = p * vals[m]
                coef += p
        for j, g in enumerate(grads[l]):
            out[prog.leaf_off[l] + j] = coef * g
    return out


def kern_grad(prog: KernelProgram, theta, X, dL_dK):
    """grad_j = sum_ik dL_dK_ik dK_ik/dtheta_j  ([GPy] update_gradients_full)."""
    return np.array([np.sum(dL_dK * dk) for dk in kern_dK(prog, theta, X)])


# --------------------------------------------------------------------------------------------------------------
# probit likelihood pieces  ([GPy] likelihoods/bernoulli.py moments_match_ep, predictive_mean)
# --------------------------------------------------------------------------------------------------------------
_SQRT2 = math.sqrt(2.0)
_SQRT_2_OVER_PI = math.sqrt(2.0 / math.pi)


def labels_pm1(y):
    """{0,1} or {-1,+1} labels -> +-1 (bernoulli.py: Y==1 -> +1; Y in {0,-1} -> -1)."""
    y = np.asarray(y).reshape(-1)
    return np.where(y == 1, 1.0, -1.0)


def pdf_over_cdf(z):
    """N(z) = phi(z)/Phi(z) = sqrt(2/pi) / erfcx(-z/sqrt2): the overflow-free form used on the device."""
    return _SQRT_2_OVER_PI / sps.erfcx(-z / _SQRT2)


def log_cdf(z):
    """ln Phi(z); erfc branch for z > -5, erfcx branch below (same split as the device function)."""
    z = np.asarray(z, float)
    t = -z / _SQRT2
    with np.errstate(divide="ignore", invalid="ignore"):
        hi = np.log(0.5 * sps.erfc(t))
        lo = np.log(0.5 * sps.erfcx(np.maximum(t, 0.0))) - t * t
    return np.where(z > -5.0, hi, lo)


def moments_match(ypm, tau_cav, v_cav):
    """Probit tilted moments.  z = y v/sqrt(tau^2+tau); returns (lnZhat, mu_hat, sigma2_hat, z)."""
    den = np.sqrt(tau_cav * tau_cav + tau_cav)
    z = ypm * v_cav / den
    N = pdf_over_cdf(z)
    mu_hat = v_cav / tau_cav + ypm * N / den
    sigma2_hat = 1.0 / tau_cav - N * (z + N) / (tau_cav * tau_cav + tau_cav)
    return log_cdf(z), mu_hat, sigma2_hat, z


def moments_match_gpy_naive(y01, tau_i, v_i):
    """Literal GPy formula (clamps Phi at 1e-15) -- only used to validate ``moments_match`` for moderate z."""
    sign = 1.0 if y01 == 1 else -1.0
    z = sign * v_i / np.sqrt(tau_i ** 2 + tau_i)
    Z_hat = 0.5 * sps.erfc(-z / _SQRT2)
    Z_hat = 1e-15 if Z_hat == 0 else Z_hat
    phi = math.exp(-0.5 * z * z) / math.sqrt(2 * math.pi)
    mu_hat = v_i / tau_i + sign * phi / (Z_hat * np.sqrt(tau_i ** 2 + tau_i))
    sigma2_hat = 1.0 / tau_i - (phi / ((tau_i ** 2 + tau_i) * Z_hat)) * (z + phi / Z_hat)
    return Z_hat, mu_hat, sigma2_hat


# --------------------------------------------------------------------------------------------------------------
# posterior from sites  ([GPy] EP end-of-sweep recompute; R&W eq. 3.53 / 3.67-3.68)
# --------------------------------------------------------------------------------------------------------------
def jitchol(A, maxtries=5):
    """[GPy] util/linalg.py jitchol: retry with jitter = mean(diag) * 1e-6 * 10**k, k < maxtries."""
    try:
        return sla.cholesky(A, lower=True, check_finite=False), 0.0
    except sla.LinAlgError:
        pass
    dg = np.diag(A)
    if np.any(dg <= 0.0) or not np.all(np.isfinite(dg)):
        raise sla.LinAlgError("not pd: non-positive diagonal elements")
    jitter = dg.mean() * 1e-6
    for _ in range(maxtries):
        try:
            return sla.cholesky(A + np.eye(A.shape[0]) * jitter, lower=True, check_finite=False), jitter
        except sla.LinAlgError:
            jitter *= 10
    raise sla.LinAlgError("not positive definite, even with jitter.")


def posterior_from_sites(K, tau_t, v_t):
    """B = I + S^1/2 K S^1/2 = L L^T;  diag(Sigma), mu = Sigma v~.  Returns (L, sig_diag, mu, jitter)."""
    n = K.shape[0]
    if not np.any(tau_t != 0.0):
        return np.eye(n), np.diag(K).copy(), K @ v_t, 0.0
    s = np.sqrt(tau_t)
    B = np.eye(n) + s[:, None] * K * s[None, :]
    L, jit = jitchol(B)
    V = sla.solve_triangular(L, s[:, None] * K, lower=True, check_finite=False)    # [GPy] dtrtrs(L, Sroot_tilde_K)
    sig_diag = np.diag(K) - np.einsum("ij,ij->j", V, V)                               # diag(K - V^T V)
    w = K @ v_t
    mu = w - V.T @ (V @ v_t)
    return L, sig_diag, mu, jit


# --------------------------------------------------------------------------------------------------------------
# EP  ([GPy] expectation_propagation.py)
# --------------------------------------------------------------------------------------------------------------
@dataclass
class EPResult:
    tau_t: np.ndarray
    v_t: np.ndarray
    sweeps: int
    converged: bool
    jitter: float = 0.0
    kept: bool = False      # parallel EP only: converged without a final update (posterior reuse)


def ep_sequential(K, y, eps=1e-6, max_sweeps=100, rng: Optional[np.random.Generator] = None, eta=1.0, delta=1.0):
    """GPy-literal sequential EP: random site order, rank-1 (DSYR) Sigma update and full ``mu = Sigma v~`` inside
    the site loop, Cholesky recompute at the end of every sweep, stop when mean(d tau^2) <= eps and
    mean(d v^2) <= eps measured between sweeps (>= 2 sweeps)."""
    rng = np.random.default_rng(0) if rng is None else rng
    ypm = labels_pm1(y)
    n = K.shape[0]
    mu = np.zeros(n)
    Sigma = K.copy()
    tau_t, v_t = np.zeros(n), np.zeros(n)
    tau_old = v_old = None
    tau_diff = v_diff = eps + 1.0
    it = 0
    jit_used = 0.0
    while (tau_diff > eps or v_diff > eps) and it < max_sweeps:
        for i in rng.permutation(n):
            tau_cav = 1.0 / Sigma[i, i] - eta * tau_t[i]
            v_cav = mu[i] / Sigma[i, i] - eta * v_t[i]
            _, mu_hat, s2_hat, _ = moments_match(ypm[i], tau_cav, v_cav)
            d_tau = delta / eta * (1.0 / s2_hat - 1.0 / Sigma[i, i])
            d_v = delta / eta * (mu_hat / s2_hat - mu[i] / Sigma[i, i])
            tau_t[i] += d_tau
            v_t[i] += d_v
            ci = d_tau / (1.0 + d_tau * Sigma[i, i])
            si = Sigma[:, i].copy()
            Sigma -= ci * np.outer(si, si)                                   # DSYR(Sigma, Sigma[:, i], -ci)
            mu = Sigma @ v_t
        s = np.sqrt(tau_t)
        B = np.eye(n) + s[:, None] * K * s[None, :]
        L, jit = jitchol(B)
        jit_used = max(jit_used, jit)
        V = sla.solve_triangular(L, s[:, None] * K, lower=True, check_finite=False)
        Sigma = K - V.T @ V
        mu = Sigma @ v_t
        if it > 0:
            tau_diff = np.mean(np.square(tau_t - tau_old))
            v_diff = np.mean(np.square(v_t - v_old))
        tau_old, v_old = tau_t.copy(), v_t.copy()
        it += 1
    return EPResult(tau_t, v_t, it, not (tau_diff > eps or v_diff > eps), jit_used)


TAU_FLOOR_REL = 1e-10   # tau~_i >= 1e-10 / K_ii: probit sites are > 0 in exact arithmetic; the floor keeps the
                        # device's diag(Sigma)_i = (1 - (B^-1)_ii) / tau~_i well defined (effect on log Z < 1e-10 / site)


def ep_parallel(K, y, eps=1e-6, max_sweeps=100, damping=1.0, tau0=None, v0=None):
    """Parallel EP (the schedule the CUDA engine runs; same fixed point as ``ep_sequential``).

    Every sweep: (1) posterior marginals diag(Sigma), mu from the *current* sites (no factorisation needed while
    all tau~ are 0: Sigma = K, mu = 0); (2) all n cavity / moment-match / site updates from those marginals,
    new = old + damping * delta.  Convergence as in GPy: mean(d tau^2) <= eps and mean(d v^2) <= eps, tested from
    the second sweep on.  Returns the sites *after* the last update (the caller recomputes the posterior), or -- when
    the residual at the current sites is already ``EP_REUSE_FACTOR`` below the tolerance -- the current sites untouched.
    """
    ypm = labels_pm1(y)
    n = K.shape[0]
    tau_t = np.zeros(n) if tau0 is None else np.array(tau0, float)
    v_t = np.zeros(n) if v0 is None else np.array(v0, float)
    it, converged, jit_used, kept = 0, False, 0.0, False
    # GPy measures the change between two sweeps (so >= 2 sweeps); the residual is an absolute fixed-point error, so
    # warm-started sites that are already converged may stop after one sweep
    min_sweeps = 2 if tau0 is None else 1
    floor = TAU_FLOOR_REL / np.diag(K)
    while it < max_sweeps:
        _, sig, mu, jit = posterior_from_sites(K, tau_t, v_t)
        jit_used = max(jit_used, jit)
        tau_cav = 1.0 / sig - tau_t
        v_cav = mu / sig - v_t
        _, mu_hat, s2_hat, _ = moments_match(ypm, tau_cav, v_cav)
        f_tau = 1.0 / s2_hat - 1.0 / sig            # full (undamped) step = the fixed-point residual
        f_v = mu_hat / s2_hat - mu / sig
        d = ep_damping(damping, it + 1)
        new_tau = np.maximum(tau_t + d * f_tau, floor)
        r_tau = np.maximum(tau_t + f_tau, floor) - tau_t
        it += 1
        # convergence is judged on the undamped residual, so damping never loosens the stopping rule
        # (identical to GPy's test for damping = 1)
        res_tau, res_v = np.mean(r_tau * r_tau), np.mean(f_v * f_v)
        if it >= min_sweeps and res_tau <= eps and res_v <= eps:
            converged = True
            # posterior reuse (engine: ep.cu EP_REUSE_FACTOR): the residual belongs to the CURRENT sites; when it is
            # already 10x below the tolerance they are kept as the final sites, so the factorisation just computed is
            # the final one.  Otherwise the update is applied as in GPy and the caller re-factorises.
            if res_tau <= EP_REUSE_FACTOR * eps and res_v <= EP_REUSE_FACTOR * eps:
                kept = True
                break
            tau_t, v_t = new_tau, v_t + d * f_v
            break
        tau_t, v_t = new_tau, v_t + d * f_v
    return EPResult(tau_t, v_t, it, converged, jit_used, kept)


EP_AUTO_FREE_SWEEPS, EP_AUTO_DAMPING = 3, 0.8
EP_REUSE_FACTOR = 0.1


def ep_damping(damping, sweep: int) -> float:
    """Step size of sweep ``sweep`` (1-based).  ``damping > 0``: constant (1.0 = GPy's delta).  ``damping <= 0`` /
    None: the engine's automatic schedule -- undamped for the first 3 sweeps, 0.8 afterwards.  Undamped *parallel* EP
    falls into a limit cycle on strongly correlated priors (large variance x long lengthscale: ~2 % of the evaluations
    of a search) where GPy's sequential EP converges; 0.8 restores convergence to the same fixed point at the cost of
    at most one extra sweep on the easy cases."""
    if damping is None or damping <= 0:
        return 1.0 if sweep <= EP_AUTO_FREE_SWEEPS else EP_AUTO_DAMPING
    return float(damping)


# --------------------------------------------------------------------------------------------------------------
# log marginal likelihood, dL/dK  ([GPy] EP.inference + exact_gaussian_inference; [R&W] 3.65, 3.73-3.75, 5.9-style)
# --------------------------------------------------------------------------------------------------------------
def ep_log_marginal(K, y, tau_t, v_t):
    """Full EP log Z (R&W 3.65 with 3.73-3.75) at the given sites, plus everything the gradient needs.

    Returns dict(log_z, log_z_gauss, alpha, Winv, L, mu, sig_diag, jitter).
    ``log_z_gauss`` is the 2016-era GPy objective -1/2 (n ln 2pi + ln|K+S~^-1| + mu~^T alpha) (sites frozen).
    """
    ypm = labels_pm1(y)
    n = K.shape[0]
    L, sig, mu, jit = posterior_from_sites(K, tau_t, v_t)
    s = np.sqrt(tau_t)
    tau_cav = 1.0 / sig - tau_t
    v_cav = mu / sig - v_t
    z = ypm * v_cav / np.sqrt(tau_cav * tau_cav + tau_cav)
    ln_phi = log_cdf(z)
    tsum = tau_cav + tau_t
    site = (ln_phi + 0.5 * np.log1p(tau_t / tau_cav) - 0.5 * v_t * v_t / tsum
            + 0.5 * v_cav * ((tau_t / tau_cav) * v_cav - 2.0 * v_t) / tsum)
    log_z = -np.sum(np.log(np.diag(L))) + 0.5 * float(v_t @ mu) + float(np.sum(site))
    # alpha = (K + S~^-1)^-1 mu~ = v~ - S^1/2 B^-1 S^1/2 K v~ ; W^-1 = S^1/2 B^-1 S^1/2          (R&W alg. 3.5/5.2)
    w = K @ v_t
    zv = sla.cho_solve((L, True), s * w, check_finite=False)
    alpha = v_t - s * zv
    Binv = sla.cho_solve((L, True), np.eye(n), check_finite=False)
    Winv = s[:, None] * Binv * s[None, :]
    # 2016-compat Gaussian-only value: ln|K+S~^-1| = ln|B| - sum ln tau~ ;  mu~^T alpha = sum (v~/tau~) alpha
    with np.errstate(divide="ignore", invalid="ignore"):
        mu_t = np.where(tau_t > 0, v_t / tau_t, 0.0)
        ln_tau = np.where(tau_t > 0, np.log(tau_t), 0.0)
    log_z_gauss = -0.5 * (n * math.log(2 * math.pi) + 2.0 * np.sum(np.log(np.diag(L))) - np.sum(ln_tau)
                          + float(mu_t @ alpha))
    return dict(log_z=log_z, log_z_gauss=log_z_gauss, alpha=alpha, Winv=Winv, L=L, mu=mu, sig_diag=sig, jitter=jit,
                z=z)


def ep_log_marginal_gpy18(K, y, tau_t, v_t):
    """GPy>=1.8 form: Gaussian marginal of the pseudo-targets + sum ln Z~_i.  Must equal ``ep_log_marginal``."""
    ypm = labels_pm1(y)
    n = K.shape[0]
    _, sig, mu, _ = posterior_from_sites(K, tau_t, v_t)
    tau_cav = 1.0 / sig - tau_t
    v_cav = mu / sig - v_t
    z = ypm * v_cav / np.sqrt(tau_cav * tau_cav + tau_cav)
    ln_zhat = log_cdf(z)
    mu_t = v_t / tau_t
    mu_cav = v_cav / tau_cav
    s2_cav = 1.0 / tau_cav
    s2_t = 1.0 / tau_t
    ln_z_tilde = np.sum(ln_zhat + 0.5 * np.log(2 * np.pi) + 0.5 * np.log(s2_t + s2_cav)
                        + 0.5 * (mu_t - mu_cav) ** 2 / (s2_cav + s2_t))
    Ky = K + np.diag(s2_t)
    Lk = sla.cholesky(Ky, lower=True)
    a = sla.cho_solve((Lk, True), mu_t)
    gauss = -0.5 * (n * np.log(2 * np.pi) + 2 * np.sum(np.log(np.diag(Lk))) + mu_t @ a)
    return float(gauss + ln_z_tilde)


def n_effective_params(prog: KernelProgram) -> int:
    """flexible_function.py:69-74 (leaf = len(param_vector)), :1554 (Sum = sum), :1647 (Product = sum - (k-1)),
    evaluated on the distributed form's *tree* equivalent: every product term shares one scale."""
    # effective params of the original tree are supplied by the host; this helper covers flat programs
    # (sum of products of distinct leaves), where the two agree.
    seen = set()
    total = 0
    for term in prog.terms:
        for l in term:
            if l not in seen:
                seen.add(l)
                total += LEAF_NPARAM[prog.leaf_type[l]]
        total -= len(term) - 1
    return total


def bic(nlml: float, p_eff: int, n: int) -> float:
    """flexible_function.py:685-687."""
    return 2.0 * nlml + p_eff * math.log(n)


def score(prog: KernelProgram, theta, X, y, eps=1e-6, max_sweeps=100, damping=1.0, p_eff: Optional[int] = None,
          ep="parallel", rng=None, tau0=None, v0=None, frozen_sites=False):
    """One candidate scoring = one ``score_batch`` item of the C-ABI: K, EP, factorise, NLML, grad, BIC.

    ``frozen_sites=True`` skips the site updates and evaluates at (tau0, v0) (2016-GPy optimisation semantics)."""
    theta = np.asarray(theta, float)
    X = np.asarray(X, float)
    n = X.shape[0]
    out = dict(status=ST_OK, sweeps=0)
    try:
        K = kern_K(prog, theta, X)
        if frozen_sites:
            res = EPResult(np.array(tau0, float), np.array(v0, float), 0, True)
        elif ep == "parallel":
            res = ep_parallel(K, y, eps, max_sweeps, damping, tau0, v0)
        else:
            res = ep_sequential(K, y, eps, max_sweeps, rng)
        lm = ep_log_marginal(K, y, res.tau_t, res.v_t)
    except sla.LinAlgError:
        P = prog.n_theta
        return dict(status=ST_NOT_PD, sweeps=0, nlml=float("inf"), nlml_gauss=float("inf"), grad=np.zeros(P),
                    bic=float("inf"), tau=None, v=None)
    dL_dK = 0.5 * (np.outer(lm["alpha"], lm["alpha"]) - lm["Winv"])           # [GPy] exact_gaussian_inference
    g = kern_grad(prog, theta, X, dL_dK)
    nlml = -lm["log_z"]
    p_eff = n_effective_params(prog) if p_eff is None else p_eff
    st = ST_OK
    if max(res.jitter, lm["jitter"]) > 0:
        st = ST_JITTER
    if not res.converged:
        st = ST_EP_NOT_CONVERGED
    if not (np.isfinite(nlml) and np.all(np.isfinite(g))):
        st = ST_NONFINITE
    out.update(status=st, sweeps=res.sweeps, nlml=nlml, nlml_gauss=-lm["log_z_gauss"], grad=-g, dlogz=g,
               bic=bic(nlml, p_eff, n), tau=res.tau_t, v=res.v_t, alpha=lm["alpha"], Winv=lm["Winv"], K=K,
               mu=lm["mu"], sig_diag=lm["sig_diag"], L=lm["L"])
    return out


# --------------------------------------------------------------------------------------------------------------
# Laplace approximation ([R&W] Algorithms 3.1 and 5.1) -- north_star's "(or Laplace)".  No reference behaviour exists:
# AutoGPC never selects it (GPClassification defaults to EP); flexible_function.py:2021-2028 only names it.
# --------------------------------------------------------------------------------------------------------------
LAPLACE_TOL, LAPLACE_MAX_ITER, LAPLACE_MAX_HALVINGS = 1e-10, 50, 10


def probit_derivatives(ypm, f):
    """log p(y|f) = log Phi(y f) and its first three derivatives in f.  With z = y f, h = phi(z)/Phi(z):
    d1 = y h,  d2 = -h (z + h) = -W,  d3 = y h ((z + h)(z + 2h) - 1)."""
    z = ypm * f
    h = pdf_over_cdf(z)
    return log_cdf(z), ypm * h, h * (z + h), ypm * h * ((z + h) * (z + 2.0 * h) - 1.0)


def laplace_mode(K, y, tol=LAPLACE_TOL, max_iter=LAPLACE_MAX_ITER, f0=None, a0=None):
    """Newton iteration of [R&W] Alg. 3.1 with step halving on psi(f) = -1/2 a^T f + sum log p(y|f), a = K^-1 f.
    Returns dict(f, a, W, grad, d3, psi, L, iters, converged, jitter); L = chol(I + W^1/2 K W^1/2) at the final f."""
    ypm = labels_pm1(y)
    n = K.shape[0]
    f = np.zeros(n) if f0 is None else np.array(f0, float)
    a = np.zeros(n) if a0 is None else np.array(a0, float)
    lp, d1, W, _ = probit_derivatives(ypm, f)
    psi = -0.5 * float(a @ f) + float(np.sum(lp))
    it, converged, jit_used = 0, False, 0.0
    while it < max_iter:
        sW = np.sqrt(W)
        B = np.eye(n) + sW[:, None] * K * sW[None, :]
        L, jit = jitchol(B)
        jit_used = max(jit_used, jit)
        b = W * f + d1
        a_new = b - sW * sla.cho_solve((L, True), sW * (K @ b), check_finite=False)
        f_new = K @ a_new
        step, ok = 1.0, False
        for _ in range(LAPLACE_MAX_HALVINGS + 1):
            a_t = a + step * (a_new - a)
            f_t = f + step * (f_new - f)
            psi_t = -0.5 * float(a_t @ f_t) + float(np.sum(log_cdf(ypm * f_t)))
            if psi_t >= psi - 1e-12 * abs(psi):
                ok = True
                break
            step *= 0.5
        it += 1
        d_psi = psi_t - psi
        a, f, psi = a_t, f_t, psi_t
        lp, d1, W, _ = probit_derivatives(ypm, f)
        if ok and abs(d_psi) <= tol * max(1.0, abs(psi)):
            converged = True
            break
    lp, d1, W, d3 = probit_derivatives(ypm, f)
    sW = np.sqrt(W)
    L, jit = jitchol(np.eye(n) + sW[:, None] * K * sW[None, :])
    return dict(f=f, a=a, W=W, grad=d1, d3=d3, psi=-0.5 * float(a @ f) + float(np.sum(lp)), L=L, iters=it,
                converged=converged, jitter=max(jit_used, jit))


def score_laplace(prog: KernelProgram, theta, X, y, p_eff: Optional[int] = None, tol=LAPLACE_TOL,
                  max_iter=LAPLACE_MAX_ITER):
    """NLML = -log q(y|X, theta) of the Laplace approximation, its gradient ([R&W] Alg. 5.1: explicit part + the
    implicit part through the mode), BIC, and the posterior in the engine's site form: tau = W, v = W f + grad, for
    which the EP posterior formulas (and ``predict``) return the Laplace posterior N(f, (K^-1 + W)^-1)."""
    theta = np.asarray(theta, float)
    X = np.asarray(X, float)
    n = X.shape[0]
    try:
        K = kern_K(prog, theta, X)
        m = laplace_mode(K, y, tol, max_iter)
    except sla.LinAlgError:
        P = prog.n_theta
        return dict(status=ST_NOT_PD, sweeps=0, nlml=float("inf"), nlml_gauss=float("inf"), grad=np.zeros(P),
                    bic=float("inf"), tau=None, v=None)
    W, a, f, L = m["W"], m["a"], m["f"], m["L"]
    sW = np.sqrt(W)
    log_q = m["psi"] - float(np.sum(np.log(np.diag(L))))
    Binv = sla.cho_solve((L, True), np.eye(n), check_finite=False)
    R = sW[:, None] * Binv * sW[None, :]
    sig_diag = (1.0 - np.diag(Binv)) / W                     # diag((K^-1 + W)^-1) = (1 - B^-1_ii) / W_i
    # d log q / d theta_j = 1/2 a^T C a - 1/2 tr(R C) + s2^T (I - K R) C grad,   C = dK/dtheta_j          ([R&W] 5.22-5.24)
    # with s2_i = 1/2 Sigma_ii d3_i (d3 = third derivative of log p); checked by finite differences in tests/test_oracle.py
    s2 = 0.5 * sig_diag * m["d3"]
    u = s2 - R @ (K @ s2)
    G = 0.5 * (np.outer(a, a) - R) + 0.5 * (np.outer(u, m["grad"]) + np.outer(m["grad"], u))
    g = kern_grad(prog, theta, X, G)                          # d log q / d theta
    nlml = -log_q
    p_eff = n_effective_params(prog) if p_eff is None else p_eff
    st = ST_OK
    if m["jitter"] > 0:
        st = ST_JITTER
    if not m["converged"]:
        st = ST_EP_NOT_CONVERGED
    if not (np.isfinite(nlml) and np.all(np.isfinite(g))):
        st = ST_NONFINITE
    return dict(status=st, sweeps=m["iters"], nlml=nlml, nlml_gauss=nlml, grad=-g, dlogz=g, bic=bic(nlml, p_eff, n),
                tau=W, v=W * f + m["grad"], alpha=a, f=f, K=K, sig_diag=sig_diag, L=L)


def predict(prog: KernelProgram, theta, X, Xstar, tau_t, v_t, K=None):
    """[GPy] GP.predict: mu* = K*^T alpha, v* = k** - diag(K*^T W^-1 K*), p* = Phi(mu*/sqrt(1+v*)).  gpckernel.py:832."""
    theta = np.asarray(theta, float)
    K = kern_K(prog, theta, X) if K is None else K
    n = K.shape[0]
    s = np.sqrt(tau_t)
    L, _, _, _ = posterior_from_sites(K, tau_t, v_t)
    w = K @ v_t
    alpha = v_t - s * sla.cho_solve((L, True), s * w, check_finite=False)
    Ks = kern_K(prog, theta, X, Xstar)                                         # n x m
    mu_s = Ks.T @ alpha
    V = sla.solve_triangular(L, s[:, None] * Ks, lower=True, check_finite=False)
    kss = np.array([kern_K(prog, theta, Xstar[i:i + 1], Xstar[i:i + 1])[0, 0] for i in range(Xstar.shape[0])])
    var_s = kss - np.einsum("ij,ij->j", V, V)
    p = 0.5 * sps.erfc(-(mu_s / np.sqrt(1.0 + var_s)) / _SQRT2)
    return p, mu_s, var_s


# --------------------------------------------------------------------------------------------------------------
# paramz transforms + L-BFGS-B driver  ([paramz] transformations.py Logexp / Logistic; optimization.py opt_lbfgsb)
# --------------------------------------------------------------------------------------------------------------
_LIM = 36.0


def logexp_f(x):
    x = np.asarray(x, float)
    return np.where(x > _LIM, x, np.log1p(np.exp(np.clip(x, -np.log(np.finfo(float).max), _LIM))))


def logexp_finv(f):
    f = np.asarray(f, float)
    return np.where(f > _LIM, f, np.log(np.expm1(f)))


def logexp_gradfactor(f):
    f = np.asarray(f, float)
    return np.where(f > _LIM, 1.0, -np.expm1(-f))


def logistic_f(x, lo, hi):
    return lo + (hi - lo) / (1.0 + np.exp(-x))


def logistic_finv(f, lo, hi):
    f = np.clip(f, lo + 1e-10 * (hi - lo), hi - 1e-10 * (hi - lo))
    return np.log((f - lo) / (hi - f))


def logistic_gradfactor(f, lo, hi):
    return (f - lo) * (hi - f) / (hi - lo)


@dataclass
class Transforms:
    """Per-theta constraint: kind 0 = positive (Logexp), 1 = bounded [lo, hi] (Logistic), 2 = free."""
    kind: np.ndarray
    lo: np.ndarray
    hi: np.ndarray

    def to_theta(self, x):
        x = np.asarray(x, float)
        th = np.array(x, float)
        p, b = self.kind == 0, self.kind == 1
        th[p] = logexp_f(x[p])
        th[b] = logistic_f(x[b], self.lo[b], self.hi[b])
        return th

    def to_x(self, theta):
        theta = np.array(theta, float)
        b = self.kind == 1
        # paramz: a start value outside [lo, hi] is reset to the midpoint
        bad = b & ((theta < self.lo) | (theta > self.hi))
        theta[bad] = 0.5 * (self.lo[bad] + self.hi[bad])
        x = theta.copy()
        p = self.kind == 0
        x[p] = logexp_finv(theta[p])
        x[b] = logistic_finv(theta[b], self.lo[b], self.hi[b])
        return x

    def gradfactor(self, theta):
        gf = np.ones_like(theta)
        p, b = self.kind == 0, self.kind == 1
        gf[p] = logexp_gradfactor(theta[p])
        gf[b] = logistic_gradfactor(theta[b], self.lo[b], self.hi[b])
        return gf


def optimize(prog, theta0, X, y, transforms: Transforms, max_evals=1000, p_eff=None, eps=1e-6, max_sweeps=100,
             damping=1.0, warm_start=True):
    """``m.optimize()`` (gpckernel.py:246): SciPy L-BFGS-B (m=10, factr=1e7, pgtol=1e-5, <=1000 evaluations) on the
    unconstrained vector; objective = -log Z_EP, <=10 linear-algebra failures tolerated (-> +inf)."""
    from scipy.optimize import fmin_l_bfgs_b
    state = dict(fails=0, evals=0, tau=None, v=None, last=None)

    def fg(x):
        theta = transforms.to_theta(x)
        state["evals"] += 1
        r = score(prog, theta, X, y, eps, max_sweeps, damping, p_eff,
                  tau0=state["tau"] if warm_start else None, v0=state["v"] if warm_start else None)
        if r["status"] in (ST_NOT_PD, ST_NONFINITE):
            state["fails"] += 1
            if state["fails"] > 10:
                raise sla.LinAlgError("too many failures")
            return 1e300, np.zeros_like(x)
        if warm_start:
            state["tau"], state["v"] = r["tau"], r["v"]
        state["last"] = r
        return r["nlml"], r["grad"] * transforms.gradfactor(theta)

    x0 = transforms.to_x(theta0)
    x, f, info = fmin_l_bfgs_b(fg, x0, maxfun=max_evals, m=10, factr=1e7, pgtol=1e-5)
    theta = transforms.to_theta(x)
    final = score(prog, theta, X, y, eps, max_sweeps, damping, p_eff)
    final.update(theta=theta, evals=state["evals"], lbfgs=info)
    return final
