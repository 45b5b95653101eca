"""BASELINE configs[2..4] end to end on the GPU engine (run under gpurun):
   search : full gpcsearch to depth 3 on synthetic D=8 N=8000 (5 folds, SE/LIN/PER, BIC), MAX_EVALS L-BFGS-B evaluations
   large  : one cold evaluation of SE0xSE1 + Per2 + SE3 at n = 32768 (EP + blocked Cholesky on one B200)
   small  : 4096 candidate x restart models at n = 1024, one cold evaluation each (batched potrf / trtri)
Prints one JSON line per mode."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autogpc_b200 import flexible_function as ff, gpckernel as gk, gpy_compat as GPy, sharding  # noqa: E402
from autogpc_b200.engine import Engine, Program, LEAF_SE, LEAF_PER, default_engine  # noqa: E402
from autogpc_b200.gpcdata import GPCData  # noqa: E402
from autogpc_b200.gpcsearch import GPCSearch  # noqa: E402
from bench import synth  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "search"

if mode == "search":
    import torch
    rank, world = sharding.world()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    N = int(os.environ.get("N", 8000))
    X, y = synth(N, 8, 0)
    data = GPCData(X, y.reshape(-1, 1), seed=0)
    gk.set_rng(0)
    s = GPCSearch(data, base_kernels="SE,Lin,Per", max_depth=3, beam_width=1, criterion="bic", n_folds=5,
                  max_evals=int(os.environ.get("MAX_EVALS", 12)))
    eng = default_engine()
    t0 = time.time()
    best, best1d = s.search()
    dt = time.time() - t0
    if rank == 0:
        print(json.dumps({"config": "configs[2] full gpcsearch depth 3, D=8 N=%d, %d GPU(s)" % (N, world), "seconds": dt,
                          "max_evals": s.maxEvals, "launches_rank0": eng.launches,
                          "candidates_per_depth": [len(lv) for lv in s.history],
                          "candidates_per_s": sum(len(lv) for lv in s.history) / dt,
                          "best_per_depth": [(k.kernel.structure(), round(k.bic(), 2), round(k.error(), 4)) for k in best[1:]],
                          "top3_per_depth": [[(k.kernel.structure(), round(k.bic(), 2)) for k in lv[:3]] for lv in s.history]}))
    if world > 1:
        dist.destroy_process_group()

elif mode == "large":
    n = int(os.environ.get("N", 32768))
    X, y = synth(n, 4, 0)
    eng = Engine(0)
    p = Program.build([LEAF_SE, LEAF_SE, LEAF_PER, LEAF_SE], [0, 1, 2, 3], [0, 2, 4, 7], [[0, 1], [2], [3]], 9)
    th = np.array([1.0, 1.2, 1.0, 0.9, 0.8, 1.5, 1.0, 0.6, 1.4])
    did = eng.dataset(X, y)
    rows = [np.arange(n, dtype=np.int32)]
    eng.profile_enable(True)
    t0 = time.time()
    r = eng.score_batch(did, [p], [th], rows)
    dt = time.time() - t0
    prof = eng.profile_read()
    g = prof["gemm"]
    out = {"config": "configs[3] single composite SE0xSE1+Per2+SE3, n=%d, one cold evaluation" % n, "seconds": dt,
           "ep_sweeps": int(r["sweeps"][0]), "status": int(r["status"][0]), "nlml": float(r["nlml"][0]),
           "grad_norm": float(np.linalg.norm(r["grad"][0])),
           "gemm_tflops": g["work"] / (g["ms"] * 1e-3) / 1e12, "kernel_ms": {k: v["ms"] for k, v in prof.items()},
           "kbuild_gbs": prof["kbuild"]["work"] / (prof["kbuild"]["ms"] * 1e-3) / 1e9}
    eng.profile_enable(False)
    # ---- correctness at size (no CPU oracle finishes here): (1) the gradient against a directional central difference
    #      of the NLML through re-converged EP; (2) the factorisation residuals of B = I + S^1/2 K S^1/2 at the final sites
    rng = np.random.default_rng(0)
    d = rng.standard_normal(len(th))
    d /= np.linalg.norm(d)
    h = 1e-4
    tight = dict(epsilon=1e-12, max_sweeps=200)
    rp = eng.score_batch(did, [p], [th * np.exp(h * d)], rows, **tight)
    rm = eng.score_batch(did, [p], [th * np.exp(-h * d)], rows, **tight)
    r0 = eng.score_batch(did, [p], [th], rows, **tight)
    fd = (rp["nlml"][0] - rm["nlml"][0]) / (2 * h)
    an = float(np.dot(r0["grad"][0, :len(th)] * th, d))   # d/dh NLML(theta exp(h d)) = sum_j g_j theta_j d_j
    out["directional_derivative"] = {"finite_difference": float(fd), "analytic": an, "rel_diff": float(abs(fd - an) / abs(an)),
                                     "h": h, "nlml_tight_vs_default_rel": float(abs(r0["nlml"][0] - r["nlml"][0]) / abs(r["nlml"][0]))}
    import torch
    from autogpc_b200.engine import Program as _P
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    Xd = torch.from_numpy(X).cuda()
    K = torch.empty(n, n, dtype=torch.float64, device="cuda")
    eng.build_K(p, th, Xd.data_ptr(), n, 4, K.data_ptr(), n, 0, st.cuda_stream)
    sroot = torch.from_numpy(np.sqrt(r0["tau"][0][:n])).cuda()
    K.mul_(sroot[:, None]).mul_(sroot[None, :])
    K.diagonal().add_(1.0)                                   # K now holds B
    A = K.clone()
    Mi = torch.empty_like(K)
    torch.cuda.synchronize()
    t0 = time.time()
    eng.potrf(A.data_ptr(), n, 1, Mi.data_ptr(), 0, 0, st.cuda_stream)
    torch.cuda.synchronize()
    t_f = time.time() - t0
    L = torch.tril(A)
    del A
    R = L @ L.T
    R.sub_(K)
    res1 = float(R.abs().max().item() / K.abs().max().item())
    del R, K
    Ml = torch.tril(Mi)
    del Mi
    R = Ml @ L
    R.diagonal().sub_(1.0)
    res2 = float(R.abs().max().item())
    out["factorisation_residuals"] = {"LLt_minus_B_over_B": res1, "ML_minus_I": res2, "potrf_trtri_seconds": t_f,
                                      "potrf_trtri_tflops": 2 * n ** 3 / 3.0 / t_f / 1e12}
    print(json.dumps(out))

elif mode == "small":
    n, n_struct, n_restart = 1024, 64, 64
    X, y = synth(n, 8, 0)
    data = GPCData(X, y.reshape(-1, 1), seed=0)
    gk.set_rng(0)
    root = gk.GPCKernel(ff.NoneKernel(), data, depth=0)
    d1 = root.expand(base_kernels="SE,Lin,Per")
    pool = d1 + [k for k in d1[0].expand(base_kernels="SE,Lin,Per") if len(k.getActiveDims()) > 1]
    pool = pool[:n_struct]
    progs, thetas = [], []
    for c in pool:
        for _ in range(n_restart):
            k = c.kernel.copy()
            gk.resetGpssParams(k, data=data)
            kern = gk.gpss2gpy(k, data=data)
            progs.append(GPy.compile_program(kern))
            thetas.append(kern.param_array)
    eng = Engine(0)
    did = eng.dataset(X, y)
    rows = [np.arange(n, dtype=np.int32)] * len(progs)
    eng.score_batch(did, progs[:64], thetas[:64], rows[:64])   # warm-up
    eng.profile_enable(True)
    t0 = time.time()
    r = eng.score_batch(did, progs, thetas, rows)
    dt = time.time() - t0
    prof = eng.profile_read()
    eng.profile_enable(False)
    rep = []
    for _ in range(3):   # steady state: arena and staging allocated, no per-launch events
        t0 = time.time()
        eng.score_batch(did, progs, thetas, rows)
        rep.append(time.time() - t0)
    g = prof["gemm"]
    print(json.dumps({"config": "configs[4] batched small: %d models (%d structures x %d restarts) at n=%d"
                                % (len(progs), len(pool), n_restart, n),
                      "seconds": dt, "models_per_s": len(progs) / dt, "seconds_repeat": rep,
                      "models_per_s_steady": len(progs) / min(rep), "ep_sweeps_mean": float(np.mean(r["sweeps"])),
                      "status_ok": int(np.sum(r["status"] == 0)),
                      "not_ok": [{"item": int(i), "structure": pool[i // n_restart].kernel.structure(), "status": int(r["status"][i]),
                                  "sweeps": int(r["sweeps"][i]), "nlml": float(r["nlml"][i]),
                                  "theta": [float(v) for v in thetas[i]]} for i in np.nonzero(r["status"] != 0)[0][:8]],
                      "gemm_tflops": g["work"] / (g["ms"] * 1e-3) / 1e12, "kernel_ms": {k: v["ms"] for k, v in prof.items()}}))
