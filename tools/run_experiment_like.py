"""The reference's own experiment shape (gpcexperiment.py:50-61, Pima: N = 768, 4 dims, depth 4, beam 2, SE bases,
CV-error ranking, full L-BFGS-B budget) on synthetic data, through GPCExperiment on the GPU engine."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autogpc_b200 import gpckernel as gk
from autogpc_b200.engine import default_engine
from autogpc_b200.gpcexperiment import GPCExperiment
from bench import synth

X, y = synth(768, 8, 3)
gk.set_rng(0)
eng = default_engine()
t0 = time.time()
out = GPCExperiment(base_kernels="SE", criterion="error", n_folds=5, max_evals=int(os.environ.get("MAX_EVALS", 1000)),
                    verbose=False).pima(X, y, depth=4, width=2)
dt = time.time() - t0
print(json.dumps({"config": "reference experiment shape: N=768, dims (1,5,6,7), depth 4, beam 2, SE, CV error, max_evals=%s"
                            % os.environ.get("MAX_EVALS", 1000),
                  "seconds": dt, "launches": eng.launches,
                  "candidates_per_depth": [len(l) for l in out["history"]],
                  "best_per_depth": [(k.kernel.structure(), round(k.error(), 4), round(k.getNLML(), 2)) for k in out["best"][1:]],
                  "baseline_error": round(out["baseline"].error(), 4),
                  "best1d": [(k.kernel.structure(), round(k.error(), 4)) for k in out["best1d"]]}))
