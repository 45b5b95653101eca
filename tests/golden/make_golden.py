"""Generates the golden vectors under tests/golden/ from the CPU oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference holds no numerical fixtures for this path, so
these pin the oracle against itself across edits and give the GPU parity tests hermetic inputs: the five
BASELINE.json configs at reduced n, plus edge cases."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gpc_oracle as orc  # noqa: E402
from tests.util import random_theta, synth  # noqa: E402

CASES = {
    # BASELINE config 1: SE-ARD on 2-D (reduced n)
    "cfg1_se_ard_2d": (("*", ("SE", 0), ("SE", 1)), 200, 2, 0),
    # config 2: one greedy expansion child on D=4
    "cfg2_expansion_child_d4": (("+", ("*", ("SE", 0), ("PER", 3)), ("LIN", 2)), 240, 4, 1),
    # config 3: depth-3 structure on D=8
    "cfg3_depth3_d8": (("+", ("*", ("SE", 0), ("SE", 1)), ("*", ("SE", 4), ("SE", 5)), ("PER", 3)), 256, 8, 2),
    # config 4: the large-n composite SE0 x SE1 + Per2 + SE3 (reduced n)
    "cfg4_composite": (("+", ("*", ("SE", 0), ("SE", 1)), ("PER", 2), ("SE", 3)), 300, 4, 3),
    # config 5: one of the batched small problems
    "cfg5_small_batched": (("*", ("SE", 0), ("+", ("SE", 1), ("CONST",))), 128, 2, 4),
    # edges: constant-only kernel, ragged n (not a multiple of the 128 tile), single leaf
    "edge_const": (("CONST",), 77, 2, 5),
    "edge_ragged_n129": (("SE", 1), 129, 3, 6),
}

if __name__ == "__main__":
    out = os.path.dirname(os.path.abspath(__file__))
    for name, (tree, N, D, seed) in CASES.items():
        X, y = synth(N + 25, D, seed)
        prog = orc.KernelProgram.from_tree(tree)
        theta = random_theta(prog, np.random.default_rng(seed + 50))
        r = orc.score(prog, theta, X[:N], y[:N])
        p, mu, var = orc.predict(prog, theta, X[:N], X[N:], r["tau"], r["v"])
        np.savez_compressed(os.path.join(out, name + ".npz"), tree=repr(tree), X=X[:N], y=y[:N], Xstar=X[N:],
                            theta=theta, nlml=r["nlml"], nlml_gauss=r["nlml_gauss"], bic=r["bic"], grad=r["grad"],
                            tau=r["tau"], nu=r["v"], sweeps=r["sweeps"], status=r["status"], pstar=p, mustar=mu,
                            varstar=var, K=orc.kern_K(prog, theta, X[:16]))
        print(name, "nlml=%.6f sweeps=%d status=%d" % (r["nlml"], r["sweeps"], r["status"]))
