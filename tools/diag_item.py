"""Diagnostic for the two config-4 items that end NOT_PD on the GPU while the oracle converges (Lin x SE with K_ii ~ 1e6):
sites after k = 1..9 sweeps, GPU vs oracle.  Run under gpurun."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autogpc_b200.engine import Engine
from oracle import gpc_oracle as orc
from tests.util import to_engine_program
from bench import synth
X, y = synth(1024, 8, 0)
eng = Engine(0)
did = eng.dataset(X, y)
rows = [np.arange(1024, dtype=np.int32)]
for tree, th in [(("*", ("LIN", 1), ("SE", 0)), [684.5038292101426, -7.263262330270217, 35.11191103441057, 4.802092614294498]),
                 (("*", ("LIN", 6), ("SE", 0)), [2.419445817821909, -8.349288593702862, 454.90781683271, 2.610330648401011])]:
    prog = orc.KernelProgram.from_tree(tree)
    ep = to_engine_program(prog)
    th = np.array(th)
    K = orc.kern_K(prog, th, X)
    for k in range(1, 10):
        r = eng.score_batch(did, [ep], [th], rows, max_sweeps=k, damping=1.0)
        o = orc.ep_parallel(K, y, 1e-6, k, 1.0)
        tg, to = r["tau"][0], o.tau_t
        print(json.dumps({"tree": str(tree), "max_sweeps": k, "gpu_status": int(r["status"][0]), "gpu_sweeps": int(r["sweeps"][0]),
                          "gpu_tau_minmax": [float(np.nanmin(tg)), float(np.nanmax(tg))], "gpu_tau_nan": int(np.isnan(tg).sum()),
                          "orc_tau_minmax": [float(to.min()), float(to.max())], "orc_sweeps": int(o.sweeps),
                          "tau_rel_diff": float(np.nanmax(np.abs(tg - to)) / np.max(np.abs(to))),
                          "gpu_nlml": float(r["nlml"][0])}), flush=True)
eng.close()
