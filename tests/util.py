"""Shared test helpers: synthetic data of SURVEY 8(d), program conversions."""
import numpy as np

from oracle import gpc_oracle as orc


def synth(N, D, seed=0):
    """X ~ U(-3,3)^{N x D}; latent = sin(2x0)cos(x1) + 0.5 x2 + sin(2 pi x3/1.5) + 0.3 x4 x5; balanced {0,1} labels."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, (N, D))
    f = np.sin(2 * X[:, 0])
    if D > 1:
        f = f * np.cos(X[:, 1])
    if D > 2:
        f = f + 0.5 * X[:, 2]
    if D > 3:
        f = f + np.sin(2 * np.pi * X[:, 3] / 1.5)
    if D > 5:
        f = f + 0.3 * X[:, 4] * X[:, 5]
    f = f + 0.3 * rng.standard_normal(N)
    y = (f > np.median(f)).astype(float)
    return X, y


TREES = {
    "se0": ("SE", 0),
    "se0xse1": ("*", ("SE", 0), ("SE", 1)),
    "se0+se1": ("+", ("SE", 0), ("SE", 1)),
    "per0": ("PER", 0),
    "se0xse1+per3+se2": ("+", ("*", ("SE", 0), ("SE", 1)), ("PER", 3), ("SE", 2)),
    "lin2x(se0+c)": ("*", ("LIN", 2), ("+", ("SE", 0), ("CONST",))),
    "mix": ("+", ("*", ("SE", 0), ("SE", 1)), ("PER", 3), ("*", ("LIN", 2), ("+", ("SE", 0), ("CONST",)))),
    "const": ("CONST",),
}


def random_theta(prog, rng):
    """Plausible constrained-space hyper-parameters for a program."""
    th = np.zeros(prog.n_theta)
    for t, o in zip(prog.leaf_type, prog.leaf_off):
        th[o] = np.exp(rng.normal(0, 0.5))                       # variance
        if t == orc.LEAF_SE:
            th[o + 1] = np.exp(rng.normal(0, 0.4))               # lengthscale
        elif t == orc.LEAF_PER:
            th[o + 1] = np.exp(rng.normal(0.3, 0.3))             # period
            th[o + 2] = np.exp(rng.normal(0, 0.3))               # lengthscale
        elif t == orc.LEAF_LIN:
            th[o + 1] = rng.normal(0, 1.0)                       # location
    return th


def to_engine_program(prog, p_eff=None):
    from autogpc_b200.engine import Program
    return Program.build(prog.leaf_type, prog.leaf_dim, prog.leaf_off, prog.terms, prog.n_theta,
                         orc.n_effective_params(prog) if p_eff is None else p_eff)
