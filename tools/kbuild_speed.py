"""K-build throughput (HBM-write roofline: 8 n^2 bytes) for the programs of the bench and of config 4.  Run under gpurun.
AGPC_KB_OCC=2|3 selects the occupancy variant of build_K_tiled_kernel."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autogpc_b200.engine import Engine, Program, LEAF_SE, LEAF_PER, LEAF_LIN
from oracle import gpc_oracle as orc

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
eng = Engine(0)
_stream = torch.cuda.Stream()
torch.cuda.set_stream(_stream)
st = _stream.cuda_stream
progs = {
    "SE0": Program.build([LEAF_SE], [0], [0], [[0]], 2),
    "SE0xSE1": Program.build([LEAF_SE, LEAF_SE], [0, 1], [0, 2], [[0, 1]], 4),
    "Lin2+SE3": Program.build([LEAF_LIN, LEAF_SE], [2, 3], [0, 2], [[0], [1]], 4),
    "Per2": Program.build([LEAF_PER], [2], [0], [[0]], 3),
    "SE0xSE1+PER2+SE3 (config 4)": Program.build([LEAF_SE, LEAF_SE, LEAF_PER, LEAF_SE], [0, 1, 2, 3], [0, 2, 4, 7], [[0, 1], [2], [3]], 9),
}
# correctness on a ragged size first
n0 = 1000
X0 = np.random.default_rng(1).uniform(-3, 3, (n0, 8))
Xd = torch.from_numpy(X0).cuda()
for name, p in progs.items():
    th = np.abs(np.random.default_rng(0).normal(1.0, 0.2, p.n_theta)) + 0.3
    ld = 1024
    K = torch.full((n0, ld), 7.0, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    eng.build_K(p, th, Xd.data_ptr(), n0, 8, K.data_ptr(), ld, 0, st)
    torch.cuda.synchronize()
    op = orc.KernelProgram(list(p.leaf_type[:p.n_leaves]), list(p.leaf_dim[:p.n_leaves]), list(p.leaf_off[:p.n_leaves]), p.terms(), p.n_theta)
    ref = orc.kern_K(op, th, X0)
    err = np.abs(K.cpu().numpy()[:, :n0] - ref).max()
    print(json.dumps({"check": name, "n": n0, "max_abs_err": err, "ok": bool(err < 1e-12)}), flush=True)


def timed(f, reps=7):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for n in (6400, 8192, 16384):
    X = torch.rand(n, 8, dtype=torch.float64, device="cuda") * 6 - 3
    K = torch.empty(n, n, dtype=torch.float64, device="cuda")
    for name, p in progs.items():
        th = np.abs(np.random.default_rng(0).normal(1.0, 0.2, p.n_theta)) + 0.3
        ms = timed(lambda: eng.build_K(p, th, X.data_ptr(), n, 8, K.data_ptr(), n, 0, st))
        gbs = 8.0 * n * n / ms / 1e6
        print(json.dumps({"block": "build_K", "occ": os.environ.get("AGPC_KB_OCC", "3"), "program": name, "n": n, "ms": round(ms, 4),
                          "gbs": round(gbs, 1), "frac_hbm": round(gbs / HBM, 3)}), flush=True)
    del X, K
eng.close()
