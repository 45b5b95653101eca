#!/usr/bin/env python
"""Headline benchmark of the candidate-scoring hot path (BASELINE.json metric, configs[2]).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a engine
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm: the oracle (NumPy/LAPACK restatement of the
                                                             # GPy path -- GPy itself cannot be installed)

Workload (weak scaling, synthetic): N = 8000 points, D = 8, 5 folds (n = 6400 training points per model).  The
candidate pool is what gpcsearch generates with SE/LIN/PER bases (depth-1 and depth-2 expansions): one block of 8
candidates, one per structure shape, and per additional GPU the same block with fresh random restarts (distinct models,
the same expected work per GPU); `--cands` candidates x 5 folds = 40 items per GPU, dealt over the ranks longest-first
on their measured cost (autogpc_b200/sharding.py).  One STEP = every item of the rank taken through one full cold-start
evaluation -- K build, parallel EP to epsilon = 1e-6, blocked FP64 Cholesky / triangular inverse per sweep, log Z_EP,
gradient, BIC -- followed (N > 1) by the NCCL all-gather of the (score, theta) records.  A candidate counts as scored
when its 5 fold models have been evaluated: value = candidates / s over all ranks.

`value`: inputs (dataset, theta, row lists) resident in HBM, device outputs, CUDA-event timed, max over ranks.
`e2e`  : the same step through the public host API (Engine.dataset + Engine.score_batch, the call GPCKernel.train
         makes) with pinned HOST buffers: X, y, theta, rows up; nlml, grad, bic, sites down -- inside the timed region.
`roofline`: the dominant kernel (gemm_nt_kernel: TMA + FP64 DMMA tile GEMM behind potrf / trtri / lauum), achieved =
         algorithmic flops (n^3/3 per potrf, trtri, lauum) / summed CUDA-event durations of its launches
         inside the timed region; peak = cuBLAS DGEMM measured live in this process (MEASURED_PEAKS.json has no FP64
         entry).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_POINTS, N_DIM, N_FOLDS = 8000, 8, 5
METRIC = "candidate kernels scored/sec at n=8k D=8"
UNIT = "candidates/s"


# ---------------------------------------------------------------------------------------------------------------
def synth(N, D, seed=0):
    """SURVEY 8(d) generator: X ~ U(-3,3)^{N x D}, analytic latent, balanced {0,1} labels."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, (N, D))
    f = np.sin(2 * X[:, 0]) * np.cos(X[:, 1]) + 0.5 * X[:, 2] + np.sin(2 * np.pi * X[:, 3] / 1.5)
    if D > 5:
        f = f + 0.3 * X[:, 4] * X[:, 5]
    f = f + 0.3 * rng.standard_normal(N)
    return X, (f > np.median(f)).astype(float)


def build_workload(world: int, cands_per_gpu: int, n_points: int = N_POINTS, fresh_restarts: bool = False):
    """Global item list [(candidate index, fold, gpy-shaped kernel, theta0)] in canonical order + the data."""
    from autogpc_b200 import flexible_function as ff
    from autogpc_b200 import gpckernel as gk
    from autogpc_b200.gpcdata import GPCData
    X, y = synth(n_points, N_DIM, 0)
    data = GPCData(X, y.reshape(-1, 1), seed=0)
    gk.set_rng(0)
    root = gk.GPCKernel(ff.NoneKernel(), data, depth=0)
    depth1 = root.expand(base_kernels="SE,Lin,Per")
    depth2 = [k for k in depth1[0].expand(base_kernels="SE,Lin,Per") if len(k.getActiveDims()) > 1]
    # One BLOCK of `cands_per_gpu` candidates is a stratified sample of what a search to depth 3 scores (24 / 42 / 72
    # candidates per level, i.e. mostly composite ones): 3 depth-1 and 5 depth-2 structures per 8, both strata shuffled
    # with a fixed seed (the expansions come grouped by base kernel: SE.., Lin.., Per..).
    prng = np.random.default_rng(0)
    d1 = [depth1[i] for i in prng.permutation(len(depth1))]
    d2 = [depth2[i] for i in prng.permutation(len(depth2))]
    import re

    def shape(c):   # structure with the dimensions removed: "Per", "SE*SE", "Lin+SE", ...
        return re.sub(r"\d+", "", c.kernel.structure())

    def shift_safe(c):   # a two-leaf kernel of one type on dimensions 4 apart maps onto itself under a shift by 4
        dims = [int(x) for x in re.findall(r"\d+", c.kernel.structure())]
        return len(dims) < 2 or (dims[0] - dims[1]) % N_DIM != N_DIM // 2

    def pick(pool, n, taken):
        out = []
        while len(out) < n:   # one candidate per shape class per round, so that dimension shifts never collide
            progress = False
            for c in pool:
                if len(out) < n and shape(c) not in taken and shift_safe(c):
                    taken.add(shape(c))
                    out.append(c)
                    progress = True
            if not progress:
                taken.clear()
        return out

    n1 = sum(1 for i in range(cands_per_gpu) if i % 8 in (0, 3, 6))
    p1, p2 = pick(d1, n1, set()), pick(d2, cands_per_gpu - n1, set())
    block = [(p1.pop(0) if i % 8 in (0, 3, 6) else p2.pop(0)) for i in range(cands_per_gpu)]

    # Weak scaling: every further GPU adds a REPLICA of the block -- the same structures from the same start points theta_0
    # (distinct item ids, scored independently: nothing in the engine or the host recognises equal items), so the work
    # per GPU is fixed by construction and the 1 -> N curve measures the system (dealing, exchange), not the draw.
    # `fresh_restarts` instead gives every block its own restarts from the one seeded stream (gpckernel.py:241-243 draws
    # one per fold): distinct models, but then the blocks differ in cost -- with this seed block 1 needs 4.28 EP sweeps per
    # item against 3.875 for block 0, +9.5 % work, which the driver's efficiency would book as a scaling loss
    # (profiles/r02_summary.md section 4).  (Round 1 shifted the input dimensions per block: 4.1-4.6 sweeps.)
    folds = data.kFoldIndices(N_FOLDS)

    def draw(block_cands, ci0):   # one random restart per (candidate, fold) (gpckernel.py:241-243), from the seeded stream
        out = []
        for ci, c in enumerate(block_cands):
            for f in range(N_FOLDS):
                k = c.kernel.copy()
                gk.resetGpssParams(k, data=data)
                out.append((ci0 + ci, f, gk.gpss2gpy(k, data=data)))
        return out

    cands = list(block)
    items = draw(block, 0)            # block 0 first: its restarts are the same draws whatever the number of GPUs
    for j in range(1, world):
        copies = [gk.GPCKernel(c.kernel.copy(), data, depth=c.depth) for c in block]
        cands += copies
        if fresh_restarts:
            items += draw(copies, j * len(block))
        else:
            items += [(ci + j * len(block), f, k) for (ci, f, k) in items[:len(block) * N_FOLDS]]
    return data, cands, folds, items


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, device: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        sm, mx, reasons = [], 0.0, set()
        for t, c in self.rows:
            if t0 <= t <= t1 and len(c) >= 7:
                try:
                    sm.append(float(c[0]))
                    mx = max(mx, float(c[1]))
                except ValueError:
                    continue
                for name, v in zip(self.NAMES, c[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


def measure_dgemm_peak(torch, seconds=2.0):
    """cuBLAS DGEMM 8192^3: best-of-5 burst and a back-to-back `seconds`-long loop (sustained) in TFLOP/s."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps = max(3, int(seconds * 1e3 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    sustained = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12, sustained


def oracle_item_seconds(data, folds, item):
    """One item through the CPU oracle (parallel-EP mode: BLAS/LAPACK-bound, all host threads)."""
    from oracle import gpc_oracle as orc  # the checker, used here only as the CPU baseline leg
    from autogpc_b200 import gpy_compat as GPy
    ci, f, kern = item
    eprog = GPy.compile_program(kern)
    prog = orc.KernelProgram(list(eprog.leaf_type[:eprog.n_leaves]), list(eprog.leaf_dim[:eprog.n_leaves]),
                             list(eprog.leaf_off[:eprog.n_leaves]), eprog.terms(), eprog.n_theta)
    tr = folds[f][0]
    y = data.Y.reshape(-1)
    t = time.perf_counter()
    o = orc.score(prog, kern.param_array, data.X[tr], y[tr], damping=0.0)
    return time.perf_counter() - t, o


def host_cores():
    """Host cores this process may run on (the BLAS pool is sized to this, whatever OMP_NUM_THREADS says)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def all_host_threads():
    """Context manager: BLAS / LAPACK pools of the CPU arm on every core of the box.  torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would otherwise pin the oracle to ONE thread (round-1 SCALE record)."""
    from threadpoolctl import threadpool_limits
    return threadpool_limits(limits=host_cores())


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the oracle on the box's host cores, one item of the same workload per step."""
    if rank != 0:
        return
    data, cands, folds, items = build_workload(args.blocks if args.blocks > 0 else world, args.cands, args.points, args.fresh_restarts)
    times = []
    with all_host_threads():
        for s in range(args.warmup + args.steps):
            dt, _ = oracle_item_seconds(data, folds, items[s % len(items)])
            if s >= args.warmup:
                times.append(dt)
        cores = host_threads()
    total = float(np.sum(times))
    value = (len(times) / N_FOLDS) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world, len(items)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "1 item (one fold model, n=%d, cold-start EP + NLML + gradient) per step; NumPy/"
                                   "LAPACK oracle in parallel-EP mode on %d host threads; GPy itself is not installable "
                                   "(SURVEY 8c)" % (len(folds[0][0]), cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world, n_items):
    return {"workload": "configs[2]: gpcsearch candidate scoring on synthetic D=8 N=%d, 5-fold (n=%d per model), "
                        "SE/LIN/PER bases; step = every candidate x fold item of the rank through one cold-start "
                        "evaluation (K build, parallel EP eps=1e-6, FP64 potrf+trtri per sweep, logZ, gradient, BIC)"
                        % (args.points, args.points - args.points // N_FOLDS),
            "candidates_per_gpu": args.cands, "folds": N_FOLDS, "items_total": n_items,
            "blocks": ("one block of %d candidates" % args.cands if world == 1 else
                       ("%d blocks with their own random restarts (unequal work per block)" % world if args.fresh_restarts else
                        "%d replicas of the one-GPU block (same structures, same start points, scored independently): "
                        "fixed work per GPU" % world)),
            "parallelism": ("one lock-step batch on 1 GPU" if world == 1 else
                            ("items dealt longest-first over %d GPUs on the cost measured in the previous step (EP sweeps); "
                             % world if args.deal == "lpt" else
                             "chunks of %d items claimed dynamically by %d GPUs from a shared counter; " % (args.chunk, world))
                            + "one all-gather of (score, theta) records per step"),
            "l2": "inputs larger than L2 (per-step working set >> 126 MB: %d x 5 n^2 FP64 matrices per rank)" % (args.cands * N_FOLDS)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cands", type=int, default=8, help="candidate structures per GPU (x5 folds = items per step)")
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--deal", default="lpt", choices=["lpt", "dynamic"],
                    help="N > 1: cost-aware static dealing re-dealt from measured cost (lpt) or chunks claimed from a shared counter")
    ap.add_argument("--chunk", type=int, default=8, help="items per chunk with --deal dynamic")
    ap.add_argument("--train-evals", type=int, default=50,
                    help="objective-evaluation budget per fold of the train_e2e leg (SURVEY 8d: 5 x 50 = 250 per candidate); 0 = skip")
    ap.add_argument("--fresh-restarts", action="store_true",
                    help="N > 1: every block draws its own restarts (distinct models, unequal work per GPU) instead of replicas")
    ap.add_argument("--blocks", type=int, default=0,
                    help="diagnostic: number of candidate blocks in the workload (default: one per GPU = weak scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernels-only", action="store_true",
                    help="profiling aid (ncu passes): skip the DGEMM peak probe, the e2e leg and the CPU baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # stdout must carry exactly ONE JSON line: native libraries (NCCL prints its version banner there) get stderr
    # instead, and the line goes to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    import torch
    import torch.distributed as dist
    from autogpc_b200 import gpy_compat as GPy
    from autogpc_b200 import sharding
    from autogpc_b200.engine import Engine, EPOptions, Program

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- autogpc_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    data, cands, folds, items = build_workload(args.blocks if args.blocks > 0 else world, args.cands, args.points, args.fresh_restarts)
    eng = Engine(local_rank)
    n_items = len(items)
    progs = [GPy.compile_program(it[2]) for it in items]
    thetas = [it[2].param_array for it in items]
    rows = [folds[it[1]][0] for it in items]
    stride = max(p.n_theta for p in progs)
    n_max = max(len(r) for r in rows)
    # Work distribution (weak scaling: 40 items per GPU on average).  One GPU: the 40 items are ONE lock-step batch.
    # Several GPUs, `--deal lpt` (default): every rank runs its share as one lock-step batch; the shares are dealt
    # longest-processing-time-first (sharding.deal_items) on a cost estimate per item -- a-priori (structure type) for the
    # first step, then the MEASURED cost of the previous step (EP sweeps taken, exchanged with the records anyway), so the
    # ranks finish together and nobody idles at the all-gather.  `--deal dynamic`: chunks of `--chunk` consecutive items
    # claimed from a shared counter (sharding.WorkQueue); measured slower here because 8-item batches run the
    # factorisations ~20 % less efficiently than 40-item ones (profiles/r02_summary.md).
    # Every rank keeps the (small) inputs of ALL items resident so that any item can run anywhere.
    B = n_items
    th_h = np.zeros((B, stride))
    for b, t in enumerate(thetas):
        th_h[b, :len(t)] = t
    th_all = torch.from_numpy(th_h).to(dev)
    # results of every item this rank has scored (device path), by global item id
    nlml_d = torch.zeros(B, dtype=torch.float64, device=dev)
    grad_d = torch.zeros((B, stride), dtype=torch.float64, device=dev)
    bic_d = torch.zeros_like(nlml_d)
    sweeps_d = torch.zeros(B, dtype=torch.int32, device=dev)
    status_d = torch.zeros(B, dtype=torch.int32, device=dev)
    opt = EPOptions(1e-6, 0.0, 100, 0, 0, 0)   # damping 0 = the engine's automatic schedule (production default)
    did = eng.dataset(data.X, data.Y)
    # one explicit stream for torch, the timing events and the engine's launches (a NULL stream would make the engine use
    # its own non-blocking stream, which torch.cuda.Event on the default stream does not see)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    stream = side.cuda_stream
    rank_busy_ms = []   # per step: this rank's time from the start of the step until its last batch finished
    units_run = [0]

    def item_prior(i):
        p = progs[i]
        types = [p.leaf_type[l] for l in range(p.n_leaves)]
        return sharding.item_cost(len(rows[i]), p.n_leaves, sum(1 for t in types if t == 1))

    costs = [item_prior(i) for i in range(n_items)]
    state = {"mine": sharding.shard_indices(n_items, rank, world, costs) if world > 1 else list(range(n_items))}
    chunks = [list(range(lo, min(lo + max(1, args.chunk), n_items))) for lo in range(0, n_items, max(1, args.chunk))]
    preps = {}

    def prepare(ids):
        """Compact device-resident inputs / outputs of one lock-step batch (cached: assignments repeat)."""
        key = tuple(ids)
        if key not in preps:
            nb = len(ids)
            off = np.zeros(nb + 1, dtype=np.int64)
            off[1:] = np.cumsum([len(rows[i]) for i in ids])
            idx = torch.tensor(ids, dtype=torch.long, device=dev)
            preps[key] = dict(
                nb=nb, idx=idx, off=off, prog=(Program * nb)(*[progs[i] for i in ids]),
                th=th_all.index_select(0, idx).contiguous(),
                rows=torch.from_numpy(np.concatenate([rows[i] for i in ids]).astype(np.int32)).to(dev),
                tau=torch.zeros((nb, n_max), dtype=torch.float64, device=dev),
                nu=torch.zeros((nb, n_max), dtype=torch.float64, device=dev),
                nlml=torch.zeros(nb, dtype=torch.float64, device=dev), bic=torch.zeros(nb, dtype=torch.float64, device=dev),
                gauss=torch.zeros(nb, dtype=torch.float64, device=dev),
                grad=torch.zeros((nb, stride), dtype=torch.float64, device=dev),
                sweeps=torch.zeros(nb, dtype=torch.int32, device=dev), status=torch.zeros(nb, dtype=torch.int32, device=dev))
            while len(preps) > 64:
                preps.pop(next(iter(preps)))
        return preps[key]

    def run_unit_device(ids):
        p = prepare(ids)
        eng.score_batch_dev(did, p["prog"], p["nb"], p["th"].data_ptr(), stride, p["rows"].data_ptr(), p["off"], opt,
                            p["tau"].data_ptr(), p["nu"].data_ptr(), n_max, p["nlml"].data_ptr(), p["grad"].data_ptr(),
                            p["bic"].data_ptr(), p["gauss"].data_ptr(), p["sweeps"].data_ptr(), p["status"].data_ptr(), stream)
        nlml_d.index_copy_(0, p["idx"], p["nlml"])
        bic_d.index_copy_(0, p["idx"], p["bic"])
        grad_d.index_copy_(0, p["idx"], p["grad"])
        sweeps_d.index_copy_(0, p["idx"], p["sweeps"])
        status_d.index_copy_(0, p["idx"], p["status"])

    def my_units(run_unit):
        """Runs this rank's work of one step; returns the item ids it scored."""
        if world == 1 or args.deal == "lpt":
            if state["mine"]:
                run_unit(state["mine"])
                units_run[0] += 1
            return list(state["mine"])
        q = sharding.WorkQueue(len(chunks), tag="bench")
        done = []
        while True:
            c = q.claim()
            if c is None:
                return done
            run_unit(chunks[c])
            units_run[0] += 1
            done += chunks[c]

    def exchange(ids, nlml_np, bic_np, status_np, sweeps_np):
        """All-gather of the (score, theta) records of the items this rank scored; re-deals the next step's shares."""
        if world == 1:
            return
        local = np.zeros((len(ids), sharding.RECORD_WIDTH))
        for j, i in enumerate(ids):   # the cv-error slot carries the EP sweep count: the measured cost of the item
            local[j] = sharding.pack_record(i, nlml_np[i], bic_np[i], float(sweeps_np[i]), int(status_np[i]), thetas[i])
        table = sharding.allgather_records(local, n_items, per_rank_max=n_items)
        if args.deal == "lpt":
            # cost model of one item: S sweeps = S - 1 factorisations + inverses, the final one, lauum -> ~ 2 S + 1 units
            # of n^3 / 3; then calibrated per rank: the items of a rank are scaled so that they add up to the time that
            # rank actually took (the late sweeps of a batch run with few items left, which no per-item model sees)
            model = np.array([(2.0 * table[i, 3] + 1.0) * float(len(rows[i])) ** 3 for i in range(n_items)])
            busy = torch.zeros(world, dtype=torch.float64, device=dev)
            busy[rank] = rank_busy_ms[-1] if rank_busy_ms else 0.0
            dist.all_reduce(busy)
            own = torch.full((n_items,), -1.0, dtype=torch.float64, device=dev)
            own[torch.tensor(ids, dtype=torch.long, device=dev)] = float(rank)
            dist.all_reduce(own, op=dist.ReduceOp.MAX)
            owner = own.cpu().numpy().astype(int)
            busy = busy.cpu().numpy()
            if np.all(busy > 0) and np.all(owner >= 0):
                for rr in range(world):
                    sel = owner == rr
                    if sel.any():
                        model[sel] *= busy[rr] / model[sel].sum()
            # re-deal only when it pays: the measured times carry ~1 % noise, and a new assignment every step means new
            # batch compositions every step (fresh device buffers, no steady state).  Keep the current shares unless the
            # longest-first dealing on the calibrated costs predicts a step at least 2 % shorter.
            new = sharding.deal_items(list(model), world)
            if np.all(owner >= 0):
                cur_max = max(float(model[owner == rr].sum()) for rr in range(world))
                new_max = max(float(model[new[rr]].sum()) if new[rr] else 0.0 for rr in range(world))
                if cur_max <= 1.02 * new_max:
                    return
            state["mine"] = new[rank]

    def step_device():
        t0 = time.perf_counter()
        ids = my_units(run_unit_device)
        if world > 1:
            torch.cuda.synchronize()
            rank_busy_ms.append(1e3 * (time.perf_counter() - t0))
            if os.environ.get("AGPC_BENCH_DEBUG"):
                sw = sweeps_d.cpu().numpy()[ids]
                sys.stderr.write("[bench rank %d] step: %d items, busy %.1f ms, sweeps sum %d max %d\n"
                                 % (rank, len(ids), rank_busy_ms[-1], int(sw.sum()), int(sw.max())))
            exchange(ids, nlml_d.cpu().numpy(), bic_d.cpu().numpy(), status_d.cpu().numpy(), sweeps_d.cpu().numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak_burst, peak_sustained = (0.0, 0.0) if args.kernels_only else measure_dgemm_peak(torch)
    sampler = ClockSampler(local_rank) if rank == 0 else None

    # per-launch event profiling is on during warm-up as well: the engine creates its event pairs on first use, and
    # those cudaEventCreate calls (2 per launch, ~7.7 k per step) belong to warm-up, not to the timed region
    eng.profile_enable(True)
    for _ in range(args.warmup):
        step_device()
    barrier()
    eng.profile_enable(True)   # resets the counters and re-uses the pool
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.time()
    e0.record()
    step_ev[0].record()
    for k in range(args.steps):
        step_device()
        step_ev[k + 1].record()
    e1.record()
    barrier()
    t_wall1 = time.time()
    busy_device = list(rank_busy_ms[-args.steps:])   # (the host-path leg below appends its own step times)
    ms_each = [step_ev[k].elapsed_time(step_ev[k + 1]) for k in range(args.steps)]
    clocks = None
    if sampler is not None:   # the samples of the timed region are in; stop nvidia-smi before the host-path leg
        clocks = sampler.window(t_wall0, t_wall1)
        sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = eng.launches - launches0
    prof = eng.profile_read()
    eng.profile_enable(False)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    scored = nlml_d != 0
    sweeps_mean = float(sweeps_d[scored].float().mean().item())
    # lock-step batches pay for their slowest item: the distribution of EP sweeps over this rank's items of the last step
    sweeps_hist = {int(k): int(v) for k, v in zip(*np.unique(sweeps_d[scored].cpu().numpy(), return_counts=True))}
    status_ok = int(((status_d == 0) & scored).sum().item())
    n_cands_total = len(cands)
    value = n_cands_total * args.steps / (ms * 1e-3)

    # ---- e2e: the host API with pinned host buffers, copies inside the timed region ----
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    Xp, yp = pin(data.X), pin(data.Y.reshape(-1))
    thetas_p = [pin(t) for t in thetas]
    rows_p = [pin(r.astype(np.int32)) for r in rows]

    host_ms = {"dataset_upload": 0.0, "score_batch": 0.0, "dataset_free": 0.0, "exchange": 0.0}

    def step_host():
        t0 = time.perf_counter()
        d = eng.dataset(Xp, yp)
        t1 = time.perf_counter()
        r = {"nlml": np.full(n_items, np.nan), "bic": np.full(n_items, np.nan), "status": np.zeros(n_items, dtype=np.int32),
             "sweeps": np.zeros(n_items, dtype=np.int32)}

        def run_unit_host(ids):
            rc = eng.score_batch(d, [progs[i] for i in ids], [thetas_p[i] for i in ids], [rows_p[i] for i in ids], damping=0.0)
            for k in r:
                r[k][ids] = rc[k]
        ids = my_units(run_unit_host)
        t2 = time.perf_counter()
        if world > 1:
            rank_busy_ms.append(1e3 * (t2 - t0))   # the re-deal in exchange() calibrates on the time THIS path took
            if os.environ.get("AGPC_BENCH_DEBUG"):
                sys.stderr.write("[bench rank %d] host step: %d items, busy %.1f ms, sweeps sum %d\n"
                                 % (rank, len(ids), rank_busy_ms[-1], int(r["sweeps"][ids].sum())))
        eng.dataset_free(d)
        t3 = time.perf_counter()
        exchange(ids, r["nlml"], r["bic"], r["status"], r["sweeps"])
        r["done"] = ids
        t4 = time.perf_counter()
        for k, v in zip(host_ms, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            host_ms[k] += 1e3 * v
        return r

    Bm = n_items // world   # items per rank and step (on average, with dynamic dealing)
    h2d = Xp.nbytes + yp.nbytes + Bm * stride * 8 + Bm * n_max * 4 + Bm * 8 + Bm * 784
    d2h = Bm * (8 * 3 + 4 * 2 + stride * 8) + 2 * Bm * n_max * 8
    e2e_steps = max(1, min(args.steps, 5))
    if args.kernels_only:
        if rank == 0:
            emit({"kernels_only": True, "ms_per_step": ms / args.steps, "gpu_launches": int(launches),
                  "ep_sweeps_mean": sweeps_mean, "ep_sweeps_hist_rank0": sweeps_hist,
                  "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()}})
        eng.dataset_free(did)
        eng.close()
        if world > 1:
            dist.destroy_process_group()
        return
    r = step_host()  # warm-up of the host path (staging allocations)
    barrier()
    for k in host_ms:
        host_ms[k] = 0.0
    sampler2 = ClockSampler(local_rank) if rank == 0 else None
    if sampler2 is not None:
        time.sleep(0.3)
    barrier()   # rank 0 alone slept: without this the other ranks' clocks would include waiting for it at the first exchange
    t_e2e0 = time.time()
    t0 = time.perf_counter()
    e2e_each = []
    for _ in range(e2e_steps):
        t_s = time.perf_counter()
        r = step_host()
        e2e_each.append(1e3 * (time.perf_counter() - t_s))
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_clocks = sampler2.window(t_e2e0, time.time()) if sampler2 is not None else None
    if sampler2 is not None:
        sampler2.stop()
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = n_cands_total * e2e_steps / e2e_s
    mine_ids = list(r["done"])
    same = bool(np.allclose(r["nlml"][mine_ids], nlml_d.cpu().numpy()[mine_ids], rtol=1e-9, atol=0, equal_nan=True))


    # ---- train_e2e (SURVEY 8d "candidate scored"): the SAME candidates through GPCKernel.train semantics -- 5 folds, one
    #      random restart each, L-BFGS-B with a fixed budget of `--train-evals` objective evaluations per fold, warm-started
    #      EP, held-out error by predict -- via the public API (gpckernel.train_kernels, the call GPCSearch.search makes).
    train = None
    if args.train_evals > 0:
        from autogpc_b200 import engine as eng_mod
        from autogpc_b200 import gpckernel as gk
        eng_mod.set_default_engine(eng)
        gk.set_rng(1)
        l0 = eng.launches
        barrier()
        t0 = time.perf_counter()
        gk.train_kernels(cands, mode="full", n_folds=N_FOLDS, max_evals=args.train_evals, shard=(rank, world) if world > 1 else None)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        evals = sum(int(m.optimization_info.get("evals", 0)) for c in cands for m in getattr(c, "foldModels", [])
                    if getattr(m, "optimization_info", None))
        train = {"value": len(cands) / dt, "unit": UNIT, "seconds": dt, "evals_budget_per_fold": args.train_evals,
                 "folds": N_FOLDS, "candidates": len(cands), "evaluations_this_rank": evals,
                 "gpu_launches": int(eng.launches - l0),
                 "optimizer": "device (agpc_optimize_batch, csrc/optim.cu)" if os.environ.get("AGPC_DEVICE_OPT", "") != "0"
                              else "host (SciPy setulb stepped per model, optim.py)",
                 "ranking": [(c.kernel.structure(), round(float(c.bic()), 3), round(float(c.error()), 4))
                             for c in sorted(cands, key=lambda k: (k.bic(), k.kernel.structure()))][:8],
                 "note": "gpckernel.train_kernels(mode='full', n_folds=5, max_evals=%d): every (candidate, fold) model "
                         "optimised in lock-step by L-BFGS-B from a random restart with warm-started EP, then predict on the "
                         "held-out fold; host buffers, wall clock, max over ranks" % args.train_evals}

    busy = None
    if world > 1:
        mine_ms = float(np.mean(busy_device)) if busy_device else 0.0
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = mine_ms
        dist.all_reduce(t)
        busy = {"per_rank_ms_per_step": [round(x, 1) for x in t.tolist()], "min": float(t.min()), "max": float(t.max()),
                "note": "time from the start of a step until the rank's last claimed chunk finished (device path, mean over "
                        "the timed steps); the step ends with the all-gather, i.e. at the max"}

    if rank == 0:
        g = prof["gemm"]
        oz = prof["oz_gemm"]
        total_kernel_ms = sum(v["ms"] for k, v in prof.items() if k != "gemm" and not k.startswith("oz_")) or 1.0
        fp64_eq = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        chol_ms = g["ms"] + prof["potf2"]["ms"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        if oz["ms"] > 0.5 * g["ms"]:
            # dominant kernel: the int8 tcgen05 GEMM.  achieved = int8 operations it EXECUTED (S(S+1)/2 integer MMAs per
            # FP64 product, counted per launch from its tile / K-range rules) / its event time; peak = 2 x the measured
            # dense bf16 rate (kind::i8 issues at twice the kind::f16 rate on tcgen05), sustained figure: the kernel runs
            # inside a long step
            bf16 = peaks.get("bf16_tflops_sustained") or 1400.0
            peak = 2.0 * bf16
            achieved = oz["work"] / (oz["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "ozaki_gemm_stream_kernel<S> (persistent CTAs, tcgen05.mma kind::i8, rotating "
                                                 "accumulator slots in TMEM, bulk-copy fed; sliced-integer FP64 GEMM behind potrf "
                                                 "updates, trtri, lauum)",
                    "achieved": achieved, "peak": peak, "unit": "TOP/s (int8)", "frac": achieved / peak,
                    "traffic": 1.566e9,
                    "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the 5 launches of this kernel "
                                    "in profiles/r02_final_stream_kb_ncu_raw.csv (ncu --set full of this command with --kernels-only; "
                                    "1.35-2.03 GB per launch: operand slices read once + the C tiles read and written once)",
                    "peak_source": "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (%s); the same kernel's MMA loop on "
                                   "resident operands sustains 4.4-4.5 POP/s at N >= 128 and 2.9 POP/s at its own N = 64 tile "
                                   "(TMEM holds S x 64 accumulator columns), profiles/r02_i8mma_probe.txt"
                                   % ("measured" if peaks else "fallback 1400"),
                    "executed_int8_ops": oz["work"], "kernel_ms": oz["ms"], "launches": oz["launches"],
                    "share_of_kernel_time": oz["ms"] / total_kernel_ms,
                    "fp64_equivalent_tflops": fp64_eq,
                    "fp64_equivalent_note": "algorithmic FP64 flops of potrf + trtri + lauum (n^3/3 each) / event time of ALL "
                                            "tile-GEMM launches (int8 and DMMA); the FP64 DMMA pipe peaks at %.1f TFLOP/s "
                                            "(cuBLAS DGEMM measured live)" % peak_sustained}
        else:
            roof = {"bound": "tensor", "kernel": "gemm_nt_kernel (TMA + FP64 DMMA; AGPC_OZAKI=0)", "achieved": fp64_eq,
                    "peak": peak_sustained, "unit": "TFLOP/s", "frac": fp64_eq / peak_sustained if peak_sustained else None,
                    "traffic": None, "peak_source": "cuBLAS DGEMM 8192^3 measured live (MEASURED_PEAKS.json has no FP64 entry)",
                    "algorithmic_flops": g["work"], "kernel_ms": g["ms"], "launches": g["launches"],
                    "share_of_kernel_time": g["ms"] / total_kernel_ms}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, world, len(items)),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "matches_device_path": same,
                    "host_ms_per_step": {k: v / e2e_steps for k, v in host_ms.items()},
                    "ms_each_step": e2e_each, "clocks": e2e_clocks},
            "ms_each_step": ms_each,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "fp64_dgemm_peak_tflops": {"sustained": peak_sustained, "burst": peak_burst},
            "cholesky": {"tflops": g["work"] / (chol_ms * 1e-3) / 1e12 if chol_ms > 0 else None,
                         "note": "potrf+trtri+lauum algorithmic flops / (tile GEMM + potf2_inv) event time",
                         "phases_tflops": {ph: (prof[k]["work"] / (((prof[k]["ms"] + (prof["potf2"]["ms"] if ph == "potrf" else 0.0))
                                                                   or 1.0) * 1e-3) / 1e12)
                                           for ph, k in (("potrf", "gemm_potrf"), ("trtri", "gemm_trtri"), ("lauum", "gemm_lauum"))},
                         "phases_note": "n^3/3 flops each; potrf includes potf2_inv time (with look-ahead the two overlap, "
                                        "so the sum overstates its wall time)"},
            "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
            "evals_per_s": len(items) * args.steps / (ms * 1e-3),
            "ep_sweeps_mean": sweeps_mean, "ep_sweeps_hist_rank0": sweeps_hist, "items_ok": status_ok, "items_per_gpu": n_items // world,
            "batches_this_rank_per_step": units_run[0] / float(args.warmup + args.steps + e2e_steps + 1),
            "dealing": "one batch" if world == 1 else args.deal,
            "int8_slices": int(os.environ.get("AGPC_OZAKI", "7") or 0),
        }
        if busy is not None:
            line["rank_busy"] = busy
        if train is not None:
            line["train_e2e"] = train
        kb = prof["kbuild"]
        if kb["ms"] > 0:
            line["kbuild_gbs"] = kb["work"] / (kb["ms"] * 1e-3) / 1e9
            line["kbuild_frac_of_hbm"] = line["kbuild_gbs"] / (peaks.get("hbm_gbs") or 6650.0)
        sl = prof["slice"]
        if sl["ms"] > 0:
            line["slice_gbs_read"] = sl["work"] / (sl["ms"] * 1e-3) / 1e9
        if not args.no_cpu_baseline:
            # CPU arm on a bounded sample: three of the step's items (one single-leaf, one periodic, one composite) through
            # the oracle on every host core; the first one is the timed baseline, all three are the parity check at scale
            import re as _re
            gpu = {"nlml": nlml_d.cpu().numpy(), "bic": bic_d.cpu().numpy(), "grad": grad_d.cpu().numpy()}
            done_ids = set(int(i) for i in np.nonzero(gpu["nlml"])[0])   # items this rank scored on the device path
            structs = [cands[it[0]].kernel.structure() for it in items]
            pick = []
            for want in (lambda st: not _re.search(r"[+*]", st) and "Per" not in st, lambda st: "Per" in st,
                         lambda st: bool(_re.search(r"[+*]", st)) and "Per" not in st):
                for i in sorted(done_ids):
                    if want(structs[i]) and i not in pick:
                        pick.append(i)
                        break
            par, dt0, o0 = [], None, None
            with all_host_threads():
                cores = host_threads()
                for i in pick:
                    dt, o = oracle_item_seconds(data, folds, items[i])
                    if dt0 is None:
                        dt0, o0 = dt, o
                    P = progs[i].n_theta
                    gerr = float(np.max(np.abs(gpu["grad"][i, :P] - o["grad"])) / max(np.max(np.abs(o["grad"])), 1e-300))
                    par.append({"item": int(i), "structure": structs[i], "n": len(rows[i]), "cpu_seconds": dt,
                                "nlml_rel": float(abs(gpu["nlml"][i] - o["nlml"]) / abs(o["nlml"])),
                                "bic_rel": float(abs(gpu["bic"][i] - o["bic"]) / abs(o["bic"])),
                                "grad_rel_of_max": gerr, "sweeps_cpu": int(o["sweeps"]), "sweeps_gpu": int(sweeps_d[i].item())})
            line["cpu_baseline"] = {"value": (1.0 / N_FOLDS) / dt0, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "1 of the step's %d items (one fold model, n=%d, %d EP sweeps), NumPy/"
                                              "LAPACK oracle, parallel-EP mode, %d host threads: %.1f s"
                                              % (n_items, len(rows[pick[0]]), o0["sweeps"], cores, dt0)}
            line["parity_at_scale"] = {"items": par, "max_nlml_rel": max(p["nlml_rel"] for p in par),
                                       "max_bic_rel": max(p["bic_rel"] for p in par),
                                       "max_grad_rel": max(p["grad_rel_of_max"] for p in par),
                                       "note": "GPU (device path, last timed step) vs the CPU oracle on the same items at the "
                                               "benchmark size; gradient error relative to the largest gradient component"}
        emit(line)
    eng.dataset_free(did)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
