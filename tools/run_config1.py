"""BASELINE configs[0]: single SE-ARD (SE0 x SE1) GPClassification, synthetic 2-D, n = 500, one full L-BFGS-B
optimisation -- the engine through the GPy-shaped shim, and the CPU oracle's restatement of the same optimisation timed
beside it on the host cores (the oracle is used here as the CPU baseline / checker only).  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autogpc_b200 import flexible_function as ff, gpckernel, gpy_compat as GPy  # noqa: E402
from autogpc_b200.engine import default_engine  # noqa: E402
from autogpc_b200.gpcdata import GPCData  # noqa: E402
from oracle import gpc_oracle as orc  # noqa: E402
from tests.util import synth  # noqa: E402

X, y = synth(500, 2, 0)
d = GPCData(X, y.reshape(-1, 1))
eng = default_engine()


def fit():
    k = ff.ProductKernel([ff.SqExpKernel(0, 1.0, 1.0), ff.SqExpKernel(1, 1.5, 0.8)])
    m = GPy.models.GPClassification(X, y.reshape(-1, 1), kernel=gpckernel.gpss2gpy(k, data=d))
    th0 = m.kern.param_array.copy()
    m.optimize()
    return m, th0


fit()                                                     # warm-up (context, first launches)
l0 = eng.launches
t0 = time.time()
m, th0 = fit()
t_gpu = time.time() - t0
launches = eng.launches - l0

prog = orc.KernelProgram.from_tree(("*", ("SE", 0), ("SE", 1)))
lo, hi = d.getLengthscaleBounds()
tr = orc.Transforms(np.array([0, 1, 0, 1]), np.array([0, lo[0], 0, lo[1]]), np.array([0, hi[0], 0, hi[1]]))
t0 = time.time()
ref = orc.optimize(prog, th0, X, y, tr, damping=0.0)
t_cpu = time.time() - t0
print(json.dumps({"config": "configs[0] single SE-ARD (SE0xSE1), 2-D, n=500, one L-BFGS-B optimisation",
                  "gpu_seconds": t_gpu, "gpu_launches": launches, "gpu_nlml": -m.log_likelihood(),
                  "gpu_theta": m.kern.param_array.tolist(), "gpu_evals": m.optimization_info["evals"],
                  "cpu_oracle_seconds": t_cpu, "cpu_cores": os.cpu_count(), "cpu_nlml": ref["nlml"],
                  "cpu_theta": np.asarray(ref["theta"]).tolist(), "cpu_evals": ref["evals"],
                  "nlml_rel_diff": abs(-m.log_likelihood() - ref["nlml"]) / abs(ref["nlml"])}))
