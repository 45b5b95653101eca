"""BASELINE configs[1]: one greedy gpcsearch expansion (SE/LIN/PER bases, + and x) on synthetic D=4 n=2000, every
candidate x 5 folds optimised in lock-step on the GPU and ranked by BIC.  Prints timing and the ranked structures."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autogpc_b200 import flexible_function as ff, gpckernel as gk
from autogpc_b200.engine import default_engine
from autogpc_b200.gpcdata import GPCData
from bench import synth

N, D = int(os.environ.get("N", 2000)), 4
max_evals = int(os.environ.get("MAX_EVALS", 40))
X, y = synth(N, D, 0)
data = GPCData(X, y.reshape(-1, 1), seed=0)
gk.set_rng(0)
eng = default_engine()
root = gk.GPCKernel(ff.NoneKernel(), data, depth=0)
depth1 = root.expand(base_kernels="SE,Lin,Per")
t0 = time.time(); l0 = eng.launches
gk.train_kernels(depth1, mode="full", n_folds=5, max_evals=max_evals)
t1 = time.time()
depth1.sort(key=lambda k: (k.bic(), k.kernel.structure()))
parent = depth1[0]
children = [k for k in parent.expand(base_kernels="SE,Lin,Per") if len(k.getActiveDims()) > 1]
t2 = time.time(); l1 = eng.launches
eng.profile_enable(True)
gk.train_kernels(children, mode="full", n_folds=5, max_evals=max_evals)
t3 = time.time()
prof = eng.profile_read()
eng.profile_enable(False)
children.sort(key=lambda k: (k.bic(), k.kernel.structure()))
print(json.dumps({"config": "configs[1] greedy expansion D=4 n=%d" % N, "max_evals": max_evals,
                  "depth1": {"candidates": len(depth1), "seconds": t1 - t0, "launches": l1 - l0,
                             "ranking": [(k.kernel.structure(), round(k.bic(), 3), round(k.error(), 4)) for k in depth1]},
                  "expansion": {"parent": parent.kernel.structure(), "candidates": len(children), "seconds": t3 - t2,
                                "launches": eng.launches - l1,
                                "candidates_per_s": len(children) / (t3 - t2),
                                "kernel_ms": {k: round(v["ms"], 1) for k, v in prof.items()},
                                "gpu_busy_fraction": sum(v["ms"] for k, v in prof.items() if k != "gemm") / (1e3 * (t3 - t2)),
                                "gemm_tflops": prof["gemm"]["work"] / (prof["gemm"]["ms"] * 1e-3) / 1e12,
                                "ranking": [(k.kernel.structure(), round(k.bic(), 3), round(k.error(), 4)) for k in children]}}))
