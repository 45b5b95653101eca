"""How far is the engine's EP schedule (parallel updates, automatic damping, posterior reuse; restated by the oracle's
``ep="parallel"``) from GPy's literal sequential EP at GPy's own tolerance eps = 1e-6?  CPU only (the oracle); the numbers
go into DESIGN.md section 4.  For each BASELINE kernel: NLML and gradient of (a) sequential EP with two different site
orders, (b) the engine schedule, all at eps = 1e-6, against a tightly converged fixed point (eps = 1e-14)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gpc_oracle as orc
from tests.util import random_theta, synth

CASES = [("configs[0] SE0xSE1, D=2", ("*", ("SE", 0), ("SE", 1)), 500, 2),
         ("configs[1] child Per1xSE0+SE2, D=4", ("+", ("*", ("PER", 1), ("SE", 0)), ("SE", 2)), 800, 4),
         ("configs[3] SE0xSE1+Per2+SE3, D=4", ("+", ("*", ("SE", 0), ("SE", 1)), ("PER", 2), ("SE", 3)), 800, 4),
         ("configs[2] SE5x(Lin2+SE3), D=8", ("*", ("SE", 5), ("+", ("LIN", 2), ("SE", 3))), 800, 8)]
out = []
for name, tree, n, D in CASES:
    X, y = synth(n, D, 3)
    prog = orc.KernelProgram.from_tree(tree)
    th = random_theta(prog, np.random.default_rng(4))
    t0 = time.time()
    ref = orc.score(prog, th, X, y, eps=1e-14, max_sweeps=2000, damping=0.0)
    runs = {"engine schedule (parallel, auto-damped, reuse)": orc.score(prog, th, X, y, eps=1e-6, damping=0.0),
            "sequential, site order A (GPy-literal)": orc.score(prog, th, X, y, eps=1e-6, ep="sequential", rng=np.random.default_rng(1)),
            "sequential, site order B (GPy-literal)": orc.score(prog, th, X, y, eps=1e-6, ep="sequential", rng=np.random.default_rng(2))}
    gs = np.max(np.abs(ref["grad"]))
    rec = {"case": name, "n": n, "nlml_fixed_point": ref["nlml"], "seconds": None, "runs": {}}
    for k, o in runs.items():
        rec["runs"][k] = {"sweeps": int(o["sweeps"]), "nlml_rel_vs_fixed_point": float(abs(o["nlml"] - ref["nlml"]) / abs(ref["nlml"])),
                          "grad_rel_of_max_vs_fixed_point": float(np.max(np.abs(o["grad"] - ref["grad"])) / gs)}
    a, b = runs["sequential, site order A (GPy-literal)"], runs["engine schedule (parallel, auto-damped, reuse)"]
    rec["engine_vs_sequential_A"] = {"nlml_rel": float(abs(a["nlml"] - b["nlml"]) / abs(a["nlml"])),
                                     "grad_rel_of_max": float(np.max(np.abs(a["grad"] - b["grad"])) / gs)}
    c = runs["sequential, site order B (GPy-literal)"]
    rec["sequential_A_vs_B"] = {"nlml_rel": float(abs(a["nlml"] - c["nlml"]) / abs(a["nlml"])),
                                "grad_rel_of_max": float(np.max(np.abs(a["grad"] - c["grad"])) / gs)}
    rec["seconds"] = time.time() - t0
    out.append(rec)
    print(json.dumps(rec), flush=True)
