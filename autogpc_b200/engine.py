"""ctypes binding of the C ABI in ``include/autogpc_b200.h`` (``libautogpc_b200.so``).

This is the only place the Python host touches the native engine.  There is no CPU fallback: if the library is
missing or cannot create a context on an sm_100 device, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

MAX_LEAVES, MAX_TERMS, MAX_TERM_LEN, MAX_THETA = 16, 16, 8, 48
LEAF_SE, LEAF_PER, LEAF_LIN, LEAF_CONST = 0, 1, 2, 3
LEAF_NPARAM = {LEAF_SE: 2, LEAF_PER: 3, LEAF_LIN: 2, LEAF_CONST: 1}
ST_OK, ST_JITTER, ST_NOT_PD, ST_EP_NOT_CONVERGED, ST_NONFINITE = 0, 1, 2, 3, 4
INFER_EP, INFER_LAPLACE = 0, 1

# AGPC_LIB overrides the library path (used by tools/ to A/B kernel variants); the default is the in-tree build
_LIB_PATH = os.environ.get("AGPC_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libautogpc_b200.so")


class EngineError(RuntimeError):
    pass


class Program(C.Structure):
    """Mirror of ``agpc_program``: K = sum_t prod_{l in term t} k_l over a leaf table (theta order = gpss2gpy)."""
    _fields_ = [("n_leaves", C.c_int32), ("n_terms", C.c_int32), ("n_theta", C.c_int32), ("p_eff", C.c_int32),
                ("leaf_type", C.c_int32 * MAX_LEAVES), ("leaf_dim", C.c_int32 * MAX_LEAVES),
                ("leaf_off", C.c_int32 * MAX_LEAVES), ("term_len", C.c_int32 * MAX_TERMS),
                ("term_leaf", (C.c_int32 * MAX_TERM_LEN) * MAX_TERMS)]

    @staticmethod
    def build(leaf_type: Sequence[int], leaf_dim: Sequence[int], leaf_off: Sequence[int],
              terms: Sequence[Sequence[int]], n_theta: int, p_eff: Optional[int] = None) -> "Program":
        if len(leaf_type) > MAX_LEAVES or len(terms) > MAX_TERMS or n_theta > MAX_THETA:
            raise EngineError("kernel program exceeds engine limits (%d leaves, %d terms, %d theta)"
                              % (len(leaf_type), len(terms), n_theta))
        p = Program()
        p.n_leaves, p.n_terms, p.n_theta = len(leaf_type), len(terms), n_theta
        for i, (t, d, o) in enumerate(zip(leaf_type, leaf_dim, leaf_off)):
            p.leaf_type[i], p.leaf_dim[i], p.leaf_off[i] = int(t), int(d), int(o)
        for i, term in enumerate(terms):
            if len(term) > MAX_TERM_LEN:
                raise EngineError("product term too long")
            p.term_len[i] = len(term)
            for j, l in enumerate(term):
                p.term_leaf[i][j] = int(l)
        if p_eff is None:
            seen, p_eff = set(), 0
            for term in terms:
                for l in term:
                    if l not in seen:
                        seen.add(l)
                        p_eff += LEAF_NPARAM[int(leaf_type[l])]
                p_eff -= len(term) - 1
        p.p_eff = int(p_eff)
        return p

    def terms(self) -> List[List[int]]:
        return [[self.term_leaf[t][j] for j in range(self.term_len[t])] for t in range(self.n_terms)]


class EPOptions(C.Structure):
    _fields_ = [("epsilon", C.c_double), ("damping", C.c_double), ("max_sweeps", C.c_int32), ("warm_start", C.c_int32),
                ("inference", C.c_int32), ("reserved", C.c_int32)]


class LBFGSOptions(C.Structure):
    """Mirror of ``agpc_lbfgs_options`` (defaults = paramz opt_lbfgsb / SciPy fmin_l_bfgs_b)."""
    _fields_ = [("m", C.c_int32), ("max_evals", C.c_int32), ("maxiter", C.c_int32), ("maxls", C.c_int32),
                ("factr", C.c_double), ("pgtol", C.c_double)]


KIND_FREE, KIND_POSITIVE, KIND_BOUNDED = 0, 1, 2
OPT_TASKS = {0: "running", 1: "CONVERGENCE: NORM_OF_PROJECTED_GRADIENT_<=_PGTOL",
             2: "CONVERGENCE: REL_REDUCTION_OF_F_<=_FACTR*EPSMCH", 3: "STOP: TOTAL NO. of f AND g EVALUATIONS EXCEEDS LIMIT",
             4: "STOP: TOTAL NO. of ITERATIONS REACHED LIMIT", 5: "ABNORMAL_TERMINATION_IN_LNSRCH",
             6: "too many failed evaluations"}

_lib = None


def lib():
    """Loads the native library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise EngineError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(autogpc_b200 has no CPU fallback)" % _LIB_PATH)
        L = C.CDLL(_LIB_PATH)
        L.agpc_last_error.restype = C.c_char_p
        L.agpc_launch_count.restype = C.c_int64
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def _lp(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64)) if a is not None else None


class Engine:
    """One context = one device + one stream + one workspace arena (``agpc_create``)."""

    def __init__(self, device: int = 0, workspace_limit_bytes: int = 0):
        self._L = lib()
        self._ctx = C.c_void_p()
        rc = self._L.agpc_create(int(device), C.byref(self._ctx))
        if rc != 0:
            raise EngineError("agpc_create(device=%d) failed with %d: an sm_100 (B200) device is required" % (device, rc))
        self.device = device
        if workspace_limit_bytes:
            self._check(self._L.agpc_set_workspace_limit(self._ctx, C.c_uint64(workspace_limit_bytes)))

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.agpc_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise EngineError("autogpc_b200 error %d: %s" % (rc, self._L.agpc_last_error(self._ctx).decode()))

    @property
    def ctx(self):
        return self._ctx

    @property
    def launches(self) -> int:
        return int(self._L.agpc_launch_count(self._ctx))

    # ---- dataset ----
    def dataset(self, X: np.ndarray, y: np.ndarray) -> int:
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.ascontiguousarray(np.asarray(y).reshape(-1), dtype=np.float64)
        if X.ndim != 2 or y.shape[0] != X.shape[0]:
            raise EngineError("X must be N x D and y of length N")
        did = C.c_int32(-1)
        self._check(self._L.agpc_dataset_create(self._ctx, _dp(X), _dp(y), X.shape[0], X.shape[1], C.byref(did)))
        return did.value

    def dataset_free(self, did: int):
        self._check(self._L.agpc_dataset_destroy(self._ctx, int(did)))

    # ---- scoring ----
    def score_batch(self, dataset: int, programs: Sequence[Program], thetas: Sequence[np.ndarray],
                    rows: Sequence[np.ndarray], epsilon: float = 1e-6, damping: float = 1.0, max_sweeps: int = 100,
                    tau: Optional[Sequence[np.ndarray]] = None, nu: Optional[Sequence[np.ndarray]] = None,
                    want_grad: bool = True, frozen_sites: bool = False, inference: str = "ep") -> dict:
        """Scores B items; item b = (programs[b], thetas[b], rows[b]).  Returns a dict of NumPy arrays
        (nlml, grad[B x P_max], bic, nlml_gauss, sweeps, status) and lists tau/nu of per-item site vectors."""
        B = len(programs)
        if not (B == len(thetas) == len(rows)) or B == 0:
            raise EngineError("programs, thetas and rows must have the same non-zero length")
        stride = max(int(p.n_theta) for p in programs)
        th = np.zeros((B, stride))
        for b, t in enumerate(thetas):
            t = np.asarray(t, float).reshape(-1)
            if t.shape[0] != programs[b].n_theta:
                raise EngineError("theta length %d != program n_theta %d" % (t.shape[0], programs[b].n_theta))
            th[b, :t.shape[0]] = t
        row_off = np.zeros(B + 1, dtype=np.int64)
        row_off[1:] = np.cumsum([len(r) for r in rows])
        rows_cat = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.int32) for r in rows]))
        n_max = int(max(len(r) for r in rows))
        warm = tau is not None and nu is not None
        if frozen_sites and not warm:
            raise EngineError("frozen_sites needs tau and nu")
        tau_a = np.zeros((B, n_max))
        nu_a = np.zeros((B, n_max))
        if warm:
            for b in range(B):
                tau_a[b, :len(rows[b])] = tau[b]
                nu_a[b, :len(rows[b])] = nu[b]
        if inference not in ("ep", "laplace"):
            raise EngineError("inference must be 'ep' or 'laplace'")
        lap = inference == "laplace"
        opt = EPOptions(float(epsilon), float(damping), 0 if frozen_sites else int(max_sweeps),
                        1 if (warm and not lap) else 0, INFER_LAPLACE if lap else INFER_EP, 0)
        progs = (Program * B)(*programs)
        nlml = np.zeros(B)
        grad = np.zeros((B, stride))
        bic = np.zeros(B)
        gauss = np.zeros(B)
        sweeps = np.zeros(B, dtype=np.int32)
        status = np.zeros(B, dtype=np.int32)
        self._check(self._L.agpc_score_batch(self._ctx, int(dataset), B, progs, _dp(th), stride, _ip(rows_cat),
                                             _lp(row_off), C.byref(opt), _dp(tau_a), _dp(nu_a), n_max, _dp(nlml),
                                             _dp(grad) if want_grad else None, _dp(bic), _dp(gauss), _ip(sweeps),
                                             _ip(status)))
        return dict(nlml=nlml, grad=grad, bic=bic, nlml_gauss=gauss, sweeps=sweeps, status=status,
                    tau=[tau_a[b, :len(rows[b])].copy() for b in range(B)],
                    nu=[nu_a[b, :len(rows[b])].copy() for b in range(B)])

    def optimize_batch(self, dataset: int, programs: Sequence[Program], thetas: Sequence[np.ndarray],
                       kinds: Sequence[np.ndarray], lows: Sequence[np.ndarray], highs: Sequence[np.ndarray],
                       rows: Sequence[np.ndarray], epsilon: float = 1e-6, damping: float = 1.0, max_sweeps: int = 100,
                       max_evals: int = 1000, m: int = 10, factr: float = 1e7, pgtol: float = 1e-5, maxls: int = 20,
                       maxiter: int = 15000) -> dict:
        """``agpc_optimize_batch``: m.optimize() for B models in lock-step with the optimiser state on the device.
        kinds[b][i] in (KIND_FREE, KIND_POSITIVE, KIND_BOUNDED) is the paramz constraint of parameter i of item b
        (lows / highs only matter for KIND_BOUNDED).  Returns score_batch's dict at the optimum plus theta, evals,
        iters and task (index into OPT_TASKS)."""
        B = len(programs)
        if not (B == len(thetas) == len(rows) == len(kinds) == len(lows) == len(highs)) or B == 0:
            raise EngineError("programs, thetas, kinds, lows, highs and rows must have the same non-zero length")
        stride = max(int(p.n_theta) for p in programs)
        th = np.zeros((B, stride))
        kd = np.zeros((B, stride), dtype=np.int32)
        lo = np.zeros((B, stride))
        hi = np.ones((B, stride))
        for b in range(B):
            n = int(programs[b].n_theta)
            t = np.asarray(thetas[b], float).reshape(-1)
            if t.shape[0] != n:
                raise EngineError("theta length %d != program n_theta %d" % (t.shape[0], n))
            th[b, :n] = t
            kd[b, :n] = np.asarray(kinds[b], dtype=np.int32).reshape(-1)
            lo[b, :n] = np.asarray(lows[b], float).reshape(-1)
            hi[b, :n] = np.asarray(highs[b], float).reshape(-1)
        row_off = np.zeros(B + 1, dtype=np.int64)
        row_off[1:] = np.cumsum([len(r) for r in rows])
        rows_cat = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.int32) for r in rows]))
        n_max = int(max(len(r) for r in rows))
        ep = EPOptions(float(epsilon), float(damping), int(max_sweeps), 0, INFER_EP, 0)
        lb = LBFGSOptions(int(m), int(max_evals), int(maxiter), int(maxls), float(factr), float(pgtol))
        progs = (Program * B)(*programs)
        tau_a, nu_a = np.zeros((B, n_max)), np.zeros((B, n_max))
        nlml, bic, gauss = np.zeros(B), np.zeros(B), np.zeros(B)
        grad = np.zeros((B, stride))
        sweeps, status = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
        evals, iters, task = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
        self._check(self._L.agpc_optimize_batch(self._ctx, int(dataset), B, progs, _dp(th), stride, _ip(kd), _dp(lo), _dp(hi),
                                                _ip(rows_cat), _lp(row_off), C.byref(ep), C.byref(lb), _dp(tau_a), _dp(nu_a),
                                                n_max, _dp(nlml), _dp(grad), _dp(bic), _dp(gauss), _ip(sweeps), _ip(status),
                                                _ip(evals), _ip(iters), _ip(task)))
        return dict(theta=[th[b, :int(programs[b].n_theta)].copy() for b in range(B)], nlml=nlml, grad=grad, bic=bic,
                    nlml_gauss=gauss, sweeps=sweeps, status=status, evals=evals, iters=iters, task=task,
                    tau=[tau_a[b, :len(rows[b])].copy() for b in range(B)],
                    nu=[nu_a[b, :len(rows[b])].copy() for b in range(B)])

    def score_batch_dev(self, dataset: int, progs_array, n_items: int, theta_ptr: int, theta_stride: int, rows_ptr: int,
                        row_off: np.ndarray, opt: "EPOptions", tau_ptr: int, nu_ptr: int, site_stride: int,
                        nlml_ptr: int, grad_ptr: int, bic_ptr: int, gauss_ptr: int, sweeps_ptr: int, status_ptr: int,
                        stream: int = 0):
        """``agpc_score_batch_dev``: theta / rows / sites / outputs are DEVICE pointers (e.g. ``tensor.data_ptr()``);
        ``progs_array`` is a ctypes ``Program`` array and ``row_off`` a host int64 array of n_items + 1 offsets."""
        vp = lambda p: C.c_void_p(p) if p else None
        self._check(self._L.agpc_score_batch_dev(self._ctx, int(dataset), int(n_items), progs_array, vp(theta_ptr),
                                                 int(theta_stride), vp(rows_ptr), _lp(row_off), C.byref(opt),
                                                 vp(tau_ptr), vp(nu_ptr), int(site_stride), vp(nlml_ptr), vp(grad_ptr),
                                                 vp(bic_ptr), vp(gauss_ptr), vp(sweeps_ptr), vp(status_ptr),
                                                 vp(stream)))

    # ---- per-launch event timing (bench.py roofline) ----
    CLASSES = ("kbuild", "gemm_potrf", "potf2", "matvec", "ep", "grad", "formB", "misc", "gemm_trtri", "gemm_lauum",
               "slice", "oz_potrf", "oz_trtri", "oz_lauum")

    def profile_enable(self, on: bool = True):
        self._check(self._L.agpc_profile_enable(self._ctx, 1 if on else 0))

    def profile_read(self) -> dict:
        n = len(self.CLASSES)
        ms, work, cnt = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int64)
        self._check(self._L.agpc_profile_read(self._ctx, _dp(ms), _dp(work), _lp(cnt)))
        out = {c: dict(ms=float(ms[i]), work=float(work[i]), launches=int(cnt[i])) for i, c in enumerate(self.CLASSES)}
        # "gemm" = the tile GEMM over all three phases (the phases never overlap in time: a plain sum is the union)
        parts = [out[k] for k in ("gemm_potrf", "gemm_trtri", "gemm_lauum")]
        out["gemm"] = {k: sum(p[k] for p in parts) for k in ("ms", "work", "launches")}
        # the int8 tensor-core GEMM (csrc/ozaki.cu) over the three phases: event time, launches, EXECUTED int8 operations
        parts = [out[k] for k in ("oz_potrf", "oz_trtri", "oz_lauum")]
        out["oz_gemm"] = {k: sum(p[k] for p in parts) for k in ("ms", "work", "launches")}
        return out

    def predict_batch(self, dataset: int, programs: Sequence[Program], thetas: Sequence[np.ndarray],
                      rows: Sequence[np.ndarray], tau: Sequence[np.ndarray], nu: Sequence[np.ndarray],
                      test_rows: Sequence[np.ndarray], want_latent: bool = False):
        """p* = Phi(mu*/sqrt(1+v*)) for dataset rows ``test_rows[b]`` under item b's posterior."""
        B = len(programs)
        stride = max(int(p.n_theta) for p in programs)
        th = np.zeros((B, stride))
        for b, t in enumerate(thetas):
            th[b, :len(t)] = np.asarray(t, float).reshape(-1)
        row_off = np.zeros(B + 1, dtype=np.int64)
        row_off[1:] = np.cumsum([len(r) for r in rows])
        rows_cat = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.int32) for r in rows]))
        t_off = np.zeros(B + 1, dtype=np.int64)
        t_off[1:] = np.cumsum([len(r) for r in test_rows])
        t_cat = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.int32) for r in test_rows]))
        n_max = int(max(len(r) for r in rows))
        tau_a = np.zeros((B, n_max))
        nu_a = np.zeros((B, n_max))
        for b in range(B):
            tau_a[b, :len(rows[b])] = tau[b]
            nu_a[b, :len(rows[b])] = nu[b]
        progs = (Program * B)(*programs)
        total = int(t_off[-1])
        p = np.zeros(total)
        mu = np.zeros(total) if want_latent else None
        var = np.zeros(total) if want_latent else None
        status = np.zeros(B, dtype=np.int32)
        self._check(self._L.agpc_predict_batch(self._ctx, int(dataset), B, progs, _dp(th), stride, _ip(rows_cat),
                                               _lp(row_off), _dp(tau_a), _dp(nu_a), n_max, _ip(t_cat), _lp(t_off),
                                               _dp(p), _dp(mu), _dp(var), _ip(status)))
        split = lambda a: [a[t_off[b]:t_off[b + 1]] for b in range(B)]
        if want_latent:
            return split(p), split(mu), split(var), status
        return split(p), status

    def predict_points(self, dataset: int, program: Program, theta, rows, tau, nu, Xstar, want_gradients: bool = False):
        """(p*, mu*, var*[, d mu*/d x*]) at arbitrary points Xstar (m x D) for one item trained on ``rows``."""
        Xs = np.ascontiguousarray(Xstar, dtype=np.float64)
        m, D = Xs.shape
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        th = np.ascontiguousarray(theta, dtype=np.float64)
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        nu = np.ascontiguousarray(nu, dtype=np.float64)
        p, mu, var = np.zeros(m), np.zeros(m), np.zeros(m)
        g = np.zeros((m, D)) if want_gradients else None
        self._check(self._L.agpc_predict_points(self._ctx, int(dataset), C.byref(program), _dp(th), _ip(rows), len(rows),
                                                _dp(tau), _dp(nu), _dp(Xs), m, _dp(p), _dp(mu), _dp(var), _dp(g)))
        return (p, mu, var, g) if want_gradients else (p, mu, var)

    # ---- building blocks on device tensors (micro-benchmarks, unit parity tests) ----
    def build_K(self, program: Program, theta, X_dev_ptr: int, n: int, D: int, K_dev_ptr: int, ld: int,
                dK_dev_ptr: int = 0, stream: int = 0):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        self._check(self._L.agpc_build_K(self._ctx, C.byref(program), _dp(th), C.c_void_p(X_dev_ptr), n, D,
                                         C.c_void_p(K_dev_ptr), C.c_void_p(dK_dev_ptr) if dK_dev_ptr else None,
                                         C.c_int64(ld), C.c_void_p(stream) if stream else None))

    def kernel_grad(self, program: Program, theta, X_dev_ptr: int, n: int, D: int, G_dev_ptr: int, ld: int):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        out = np.zeros(program.n_theta)
        self._check(self._L.agpc_kernel_grad(self._ctx, C.byref(program), _dp(th), C.c_void_p(X_dev_ptr), n, D,
                                             C.c_void_p(G_dev_ptr), C.c_int64(ld), _dp(out), None))
        return out

    def potrf(self, A_dev_ptr: int, n: int, batch: int = 1, Minv_ptr: int = 0, Uinv_ptr: int = 0, info_ptr: int = 0,
              stream: int = 0):
        self._check(self._L.agpc_potrf(self._ctx, C.c_void_p(A_dev_ptr), n, batch,
                                       C.c_void_p(Minv_ptr) if Minv_ptr else None,
                                       C.c_void_p(Uinv_ptr) if Uinv_ptr else None,
                                       C.c_void_p(info_ptr) if info_ptr else None,
                                       C.c_void_p(stream) if stream else None))

    def gemm_nt(self, A_ptr: int, B_ptr: int, C_ptr: int, M: int, N: int, K: int, alpha=1.0, beta=0.0, batch=1,
                stream: int = 0):
        self._check(self._L.agpc_gemm_nt(self._ctx, C.c_void_p(A_ptr), C.c_void_p(B_ptr), C.c_void_p(C_ptr), M, N, K,
                                         C.c_double(alpha), C.c_double(beta), batch,
                                         C.c_void_p(stream) if stream else None))


    def jitchol(self, A_dev_ptr: int, n: int, batch: int = 1, jitter_ptr: int = 0, info_ptr: int = 0, stream: int = 0):
        """``agpc_jitchol``: in-place Cholesky with GPy's jitter ladder (``util.linalg.jitchol``)."""
        self._check(self._L.agpc_jitchol(self._ctx, C.c_void_p(A_dev_ptr), n, batch,
                                         C.c_void_p(jitter_ptr) if jitter_ptr else None,
                                         C.c_void_p(info_ptr) if info_ptr else None,
                                         C.c_void_p(stream) if stream else None))

    def gemm_nt_i8x(self, A_ptr: int, B_ptr: int, C_ptr: int, M: int, N: int, K: int, alpha=1.0, beta=0.0, batch=1,
                    n_slices: int = 7, stream: int = 0):
        """``agpc_gemm_nt_i8x``: the same product through int8 digit slices on tcgen05 (csrc/ozaki.cu)."""
        self._check(self._L.agpc_gemm_nt_i8x(self._ctx, C.c_void_p(A_ptr), C.c_void_p(B_ptr), C.c_void_p(C_ptr), M, N, K,
                                             C.c_double(alpha), C.c_double(beta), batch, n_slices,
                                             C.c_void_p(stream) if stream else None))


_default: Optional[Engine] = None


def default_engine() -> Engine:
    """Process-wide engine on this rank's GPU (LOCAL_RANK under torchrun, else device 0)."""
    global _default
    if _default is None:
        _default = Engine(int(os.environ.get("LOCAL_RANK", "0")))
    return _default


def set_default_engine(e) -> None:
    global _default
    _default = e
