"""Host L-BFGS driver (SciPy setulb stepped per model, optim.lbfgsb_lockstep) against the on-device optimiser
(agpc_optimize_batch) on the same models: wall time, evaluations, final NLML.  Run under gpurun.
usage: optim_speed.py [n_models] [n_points] [max_evals]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autogpc_b200 import engine as eng_mod
from autogpc_b200 import gpy_compat as GPy
from autogpc_b200 import optim

n_models = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
max_evals = int(sys.argv[3]) if len(sys.argv) > 3 else 50
D = 4
rng = np.random.default_rng(0)
X = rng.normal(size=(n + n // 4, D))
f = np.sin(1.3 * X[:, 0]) + 0.5 * X[:, 1] * X[:, 2] - 0.3 * X[:, 3]
y = np.where(f + 0.3 * rng.normal(size=f.shape) > 0, 1.0, -1.0)
e = eng_mod.Engine(0)
eng_mod.set_default_engine(e)
did = e.dataset(X, y)


def kernels(seed):
    r = np.random.default_rng(seed)
    out = []
    for i in range(n_models):
        v = lambda: float(np.exp(r.normal(0, 0.5)))
        kind = i % 4
        if kind == 0:
            k = GPy.kern.RBF(1, variance=v(), lengthscale=v(), active_dims=[i % D])
        elif kind == 1:
            k = GPy.kern.Prod([GPy.kern.RBF(1, variance=v(), lengthscale=v(), active_dims=[0]),
                               GPy.kern.RBF(1, variance=v(), lengthscale=v(), active_dims=[1])])
        elif kind == 2:
            k = GPy.kern.Add([GPy.kern.RBF(1, variance=v(), lengthscale=v(), active_dims=[2]),
                              GPy.kern.StdPeriodic(1, variance=v(), period=2.0 * v(), lengthscale=v(), active_dims=[0])])
        else:
            k = GPy.kern.Add([GPy.kern.Prod([GPy.kern.RBF(1, variance=v(), lengthscale=v(), active_dims=[0]),
                                             GPy.kern.RBF(1, variance=v(), lengthscale=v(), active_dims=[1])]),
                              GPy.kern.RBF(1, variance=v(), lengthscale=v(), active_dims=[3])])
        out.append(k)
    return out


rows = [np.sort(np.random.default_rng(100 + i).choice(X.shape[0], size=n, replace=False)).astype(np.int32) for i in range(n_models)]


def run(device):
    ms = [GPy.models.GPClassification(kernel=k, engine=e, _dataset=did, _rows=r, _infer=False) for k, r in zip(kernels(7), rows)]
    t0 = time.perf_counter()
    optim.optimize_models(ms, max_evals=max_evals, device=device)
    dt = time.perf_counter() - t0
    return dt, ms


run(False)  # warm-up (arena, staging)
th, mh = run(False)
td, md = run(True)
nl_h = np.array([m._nlml for m in mh])
nl_d = np.array([m._nlml for m in md])
ev_h = np.array([m.optimization_info["evals"] for m in mh])
ev_d = np.array([m.optimization_info["evals"] for m in md])
rel = np.abs(nl_h - nl_d) / np.maximum(1.0, np.abs(nl_h))
print(json.dumps({"models": n_models, "n": n, "max_evals": max_evals, "host_s": round(th, 3), "device_s": round(td, 3),
                  "speedup": round(th / td, 3), "evals_host_mean": float(ev_h.mean()), "evals_device_mean": float(ev_d.mean()),
                  "models_per_s_host": round(n_models / th, 1), "models_per_s_device": round(n_models / td, 1),
                  "nlml_rel_diff_median": float(np.median(rel)), "nlml_rel_diff_p90": float(np.quantile(rel, 0.9)),
                  "device_better_or_equal_frac": float(np.mean(nl_d <= nl_h + 1e-6 * np.abs(nl_h)))}))
os._exit(0)
