"""GPU parity: the CUDA engine (through the C ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from oracle import gpc_oracle as orc
from tests.util import TREES, random_theta, synth, to_engine_program

pytestmark = pytest.mark.gpu

NLML_RTOL = 1e-6      # north_star: NLML and gradients within 1e-6 relative in FP64
GRAD_RTOL = 1e-6


def _score_both(engine, tree, N, D, seed, rows=None, **kw):
    X, y = synth(N, D, seed)
    prog = orc.KernelProgram.from_tree(tree)
    theta = random_theta(prog, np.random.default_rng(seed + 100))
    rows = np.arange(N, dtype=np.int32) if rows is None else rows
    did = engine.dataset(X, y)
    try:
        g = engine.score_batch(did, [to_engine_program(prog)], [theta], [rows], **kw)
    finally:
        engine.dataset_free(did)
    o = orc.score(prog, theta, X[rows], y[rows], eps=kw.get("epsilon", 1e-6), max_sweeps=kw.get("max_sweeps", 100),
                  damping=kw.get("damping", 1.0))
    return g, o


@pytest.mark.parametrize("name", list(TREES))
@pytest.mark.parametrize("N", [200, 384])
def test_score_matches_oracle(engine, name, N):
    g, o = _score_both(engine, TREES[name], N, 4, seed=3)
    assert g["status"][0] == o["status"]
    assert g["sweeps"][0] == o["sweeps"]
    assert abs(g["nlml"][0] - o["nlml"]) <= 1e-9 * abs(o["nlml"])
    assert abs(g["nlml_gauss"][0] - o["nlml_gauss"]) <= 1e-8 * abs(o["nlml_gauss"])
    assert abs(g["bic"][0] - o["bic"]) <= 1e-9 * abs(o["bic"])
    P = len(o["grad"])
    scale = np.maximum(np.abs(o["grad"]), 1e-3 * np.abs(o["grad"]).max())
    assert np.all(np.abs(g["grad"][0, :P] - o["grad"]) <= 1e-7 * scale)
    assert np.allclose(g["tau"][0], o["tau"], rtol=1e-7, atol=1e-12)
    assert np.allclose(g["nu"][0], o["v"], rtol=1e-7, atol=1e-10)


def test_score_matches_oracle_beyond_the_lookahead_threshold(engine):
    """n = 1500 / 1300 (12 and 11 panels): two-level blocking, the look-ahead streams and -- with two items -- the thin
    128 x 32 GEMM tiles are all on the path; same bar as the small cases."""
    N = 1500
    X, y = synth(N, 4, 17)
    prog = orc.KernelProgram.from_tree(TREES["se0xse1+per3+se2"])
    th = random_theta(prog, np.random.default_rng(3))
    rows = [np.arange(N, dtype=np.int32), np.arange(100, 1400, dtype=np.int32)]
    did = engine.dataset(X, y)
    g = engine.score_batch(did, [to_engine_program(prog)] * 2, [th, th * 0.9], rows, damping=0.0)
    engine.dataset_free(did)
    for b, (theta, r) in enumerate(((th, rows[0]), (th * 0.9, rows[1]))):
        o = orc.score(prog, theta, X[r], y[r], damping=0.0)
        assert g["status"][b] == o["status"] == 0 and g["sweeps"][b] == o["sweeps"]
        assert abs(g["nlml"][b] - o["nlml"]) <= 1e-9 * abs(o["nlml"])
        scale = np.maximum(np.abs(o["grad"]), 1e-3 * np.abs(o["grad"]).max())
        assert np.all(np.abs(g["grad"][b, :prog.n_theta] - o["grad"]) <= 1e-7 * scale)
        assert np.allclose(g["tau"][b], o["tau"], rtol=1e-7, atol=1e-12)


def test_score_vs_sequential_ep(engine):
    """Against the GPy-literal sequential EP (random site order, rank-1 updates): the north_star tolerance.

    At GPy's default epsilon = 1e-6 the sites are only converged to ~1e-3 RMS, so log Z (stationary in the sites)
    agrees to 1e-6 but gradients of two *different* EP runs -- including two GPy runs with different site
    permutations -- differ by ~1e-4; gradient parity is therefore asserted at the common fixed point (tight eps)."""
    X, y = synth(300, 4, 7)
    prog = orc.KernelProgram.from_tree(TREES["se0xse1+per3+se2"])
    theta = random_theta(prog, np.random.default_rng(5))
    rows = [np.arange(300, dtype=np.int32)]
    did = engine.dataset(X, y)
    g = engine.score_batch(did, [to_engine_program(prog)], [theta], rows)
    gt = engine.score_batch(did, [to_engine_program(prog)], [theta], rows, epsilon=1e-14, max_sweeps=500)
    engine.dataset_free(did)
    o = orc.score(prog, theta, X, y, ep="sequential", rng=np.random.default_rng(1))
    assert abs(g["nlml"][0] - o["nlml"]) <= NLML_RTOL * abs(o["nlml"])
    ot = orc.score(prog, theta, X, y, eps=1e-14, max_sweeps=500, ep="sequential", rng=np.random.default_rng(1))
    assert abs(gt["nlml"][0] - ot["nlml"]) <= NLML_RTOL * abs(ot["nlml"])
    scale = np.maximum(np.abs(ot["grad"]), 1e-2 * np.abs(ot["grad"]).max())
    assert np.all(np.abs(gt["grad"][0, :prog.n_theta] - ot["grad"]) <= GRAD_RTOL * scale)


def test_auto_damping_converges_where_undamped_parallel_ep_cycles(engine):
    """Strongly correlated prior (large variance x long lengthscale): undamped parallel EP limit-cycles (status 3 after
    max_sweeps), GPy's sequential EP converges; the automatic damping schedule must reach the sequential fixed point."""
    from autogpc_b200.gpcdata import GPCData
    X, y = synth(150, 3, 8)
    tr = GPCData(X, y.reshape(-1, 1), seed=1).kFoldIndices(3)[0][0]   # the fold on which a search run hit the cycle
    rows = [tr]
    Xr, yr = X[tr], y[tr]
    for tree, theta in ((("SE", 2), np.array([11.1748701, 1.25597298])),
                        (("*", ("SE", 0), ("SE", 2)), np.array([26.69141447, 2.30175307, 11.5477121, 1.34074968])),
                        (("*", ("PER", 0), ("SE", 2)), np.array([3.5139884, 0.71536986, 8.36953579, 3.38838001, 1.21044942]))):
        prog = orc.KernelProgram.from_tree(tree)
        ep = to_engine_program(prog)
        did = engine.dataset(X, y)
        und = engine.score_batch(did, [ep], [theta], rows, damping=1.0, max_sweeps=60)
        auto = engine.score_batch(did, [ep], [theta], rows, damping=0.0)
        tight = engine.score_batch(did, [ep], [theta], rows, damping=0.0, epsilon=1e-13, max_sweeps=400)
        engine.dataset_free(did)
        o_auto = orc.score(prog, theta, Xr, yr, damping=0.0)
        o_seq = orc.score(prog, theta, Xr, yr, ep="sequential", eps=1e-13, max_sweeps=400, rng=np.random.default_rng(3))
        assert und["status"][0] == 3                                   # the failure mode this schedule exists for
        assert auto["status"][0] == 0 and auto["sweeps"][0] == o_auto["sweeps"] and auto["sweeps"][0] <= 25
        assert abs(auto["nlml"][0] - o_auto["nlml"]) <= 1e-9 * abs(o_auto["nlml"])
        scale = np.maximum(np.abs(o_auto["grad"]), 1e-3 * np.abs(o_auto["grad"]).max())
        assert np.all(np.abs(auto["grad"][0, :prog.n_theta] - o_auto["grad"]) <= 1e-6 * scale)
        assert tight["status"][0] == 0
        assert abs(tight["nlml"][0] - o_seq["nlml"]) <= NLML_RTOL * abs(o_seq["nlml"])   # same fixed point as GPy's EP
        assert abs(auto["nlml"][0] - o_seq["nlml"]) <= 2e-5 * abs(o_seq["nlml"])          # at GPy's own epsilon


@pytest.mark.parametrize("name", ["se0xse1", "se0xse1+per3+se2", "mix", "const"])
def test_laplace_matches_oracle(engine, name):
    """north_star's "(or Laplace)": Newton mode finding + R&W alg. 5.1 gradient on the same kernels (potrf / trtri / lauum,
    fused contraction with the implicit rank-2 term) against oracle.score_laplace; predict through the site form."""
    X, y = synth(330, 4, 12)
    prog = orc.KernelProgram.from_tree(TREES[name])
    th = random_theta(prog, np.random.default_rng(8))
    tr, te = np.arange(0, 290, dtype=np.int32), np.arange(290, 330, dtype=np.int32)
    ep = to_engine_program(prog)
    did = engine.dataset(X, y)
    g = engine.score_batch(did, [ep, ep], [th, th * 1.1], [tr, tr[:200]], inference="laplace", epsilon=1e-10, max_sweeps=50)
    p, mu, var, _ = engine.predict_batch(did, [ep], [th], [tr], [g["tau"][0]], [g["nu"][0]], [te], want_latent=True)
    engine.dataset_free(did)
    for b, (theta, rows) in enumerate(((th, tr), (th * 1.1, tr[:200]))):
        o = orc.score_laplace(prog, theta, X[rows], y[rows])
        assert g["status"][b] == o["status"] == 0 and g["sweeps"][b] == o["sweeps"]
        assert abs(g["nlml"][b] - o["nlml"]) <= 1e-9 * abs(o["nlml"])
        assert abs(g["bic"][b] - o["bic"]) <= 1e-9 * abs(o["bic"])
        scale = np.maximum(np.abs(o["grad"]), 1e-3 * np.abs(o["grad"]).max())
        assert np.all(np.abs(g["grad"][b, :prog.n_theta] - o["grad"]) <= 1e-7 * scale)
        assert np.allclose(g["tau"][b], o["tau"], rtol=1e-8, atol=1e-14)
        assert np.allclose(g["nu"][b], o["v"], rtol=1e-8, atol=1e-10)
    o = orc.score_laplace(prog, th, X[tr], y[tr])
    po, muo, varo = orc.predict(prog, th, X[tr], X[te], o["tau"], o["v"])
    assert np.allclose(mu[0], muo, rtol=1e-8, atol=1e-10) and np.allclose(var[0], varo, rtol=1e-7, atol=1e-10)
    assert np.allclose(p[0], po, rtol=1e-8, atol=1e-12)


def test_batch_mixed_items_and_folds(engine):
    """Several candidates x folds in one lock-step batch, different n per item, different programs."""
    N, D = 700, 4
    X, y = synth(N, D, 11)
    rng = np.random.default_rng(0)
    perm = rng.permutation(N)
    folds = np.array_split(perm, 5)
    progs, thetas, rows = [], [], []
    for name in ["se0xse1", "se0+se1", "per0", "mix"]:
        prog = orc.KernelProgram.from_tree(TREES[name])
        for f in range(5):
            tr = np.sort(np.concatenate([folds[j] for j in range(5) if j != f])).astype(np.int32)
            if f == 2:
                tr = tr[:-7]           # ragged sizes inside one batch
            progs.append(prog)
            thetas.append(random_theta(prog, rng))
            rows.append(tr)
    did = engine.dataset(X, y)
    g = engine.score_batch(did, [to_engine_program(p) for p in progs], thetas, rows)
    engine.dataset_free(did)
    for b, (p, th, r) in enumerate(zip(progs, thetas, rows)):
        o = orc.score(p, th, X[r], y[r])
        assert g["status"][b] == o["status"]
        assert abs(g["nlml"][b] - o["nlml"]) <= 1e-8 * abs(o["nlml"]), (b, g["nlml"][b], o["nlml"])
        scale = np.maximum(np.abs(o["grad"]), 1e-3 * np.abs(o["grad"]).max())
        assert np.all(np.abs(g["grad"][b, :p.n_theta] - o["grad"]) <= 1e-6 * scale)


def test_warm_start_and_frozen_sites(engine):
    X, y = synth(256, 2, 2)
    prog = orc.KernelProgram.from_tree(TREES["se0xse1"])
    th = np.array([1.3, 0.9, 0.8, 1.2])
    rows = [np.arange(256, dtype=np.int32)]
    did = engine.dataset(X, y)
    ep = to_engine_program(prog)
    cold = engine.score_batch(did, [ep], [th], rows)
    th2 = th * 1.02
    warm = engine.score_batch(did, [ep], [th2], rows, tau=cold["tau"], nu=cold["nu"])
    cold2 = engine.score_batch(did, [ep], [th2], rows)
    assert warm["sweeps"][0] <= cold2["sweeps"][0]
    assert abs(warm["nlml"][0] - cold2["nlml"][0]) <= 1e-6 * abs(cold2["nlml"][0])
    frozen = engine.score_batch(did, [ep], [th2], rows, tau=cold["tau"], nu=cold["nu"], frozen_sites=True)
    engine.dataset_free(did)
    o = orc.score(prog, th2, X, y, tau0=cold["tau"][0], v0=cold["nu"][0], frozen_sites=True)
    assert frozen["sweeps"][0] == 0
    assert abs(frozen["nlml_gauss"][0] - o["nlml_gauss"]) <= 1e-9 * abs(o["nlml_gauss"])
    assert np.allclose(frozen["grad"][0, :4], o["grad"], rtol=1e-7, atol=1e-9)


def test_predict_matches_oracle(engine):
    N, D = 500, 4
    X, y = synth(N, D, 4)
    prog = orc.KernelProgram.from_tree(TREES["se0xse1+per3+se2"])
    th = random_theta(prog, np.random.default_rng(9))
    tr = np.arange(0, 400, dtype=np.int32)
    te = np.arange(400, 500, dtype=np.int32)
    did = engine.dataset(X, y)
    ep = to_engine_program(prog)
    g = engine.score_batch(did, [ep], [th], [tr])
    p, mu, var, st = engine.predict_batch(did, [ep], [th], [tr], g["tau"], g["nu"], [te], want_latent=True)
    engine.dataset_free(did)
    po, muo, varo = orc.predict(prog, th, X[tr], X[te], g["tau"][0], g["nu"][0])
    assert np.allclose(mu[0], muo, rtol=1e-8, atol=1e-10)
    assert np.allclose(var[0], varo, rtol=1e-7, atol=1e-10)
    assert np.allclose(p[0], po, rtol=1e-8, atol=1e-12)
    err_g = np.mean((p[0] - 0.5) * (y[te] - 0.5) < 0)
    err_o = np.mean((po - 0.5) * (y[te] - 0.5) < 0)
    assert err_g == err_o


@pytest.mark.parametrize("name", ["se0xse1+per3+se2", "mix"])
def test_predict_points_and_mean_gradients(engine, name):
    """m.predict(Xnew) / m._raw_predict / m.predictive_gradients (gpckernel.py:832, 431; gpcplot.py:110-114) at points that
    are not dataset rows: p*, mu*, v* against the oracle, d mu*/d x* against central differences of the oracle's mean."""
    X, y = synth(260, 4, 6)
    prog = orc.KernelProgram.from_tree(TREES[name])
    th = random_theta(prog, np.random.default_rng(12))
    tr = np.arange(200, dtype=np.int32)
    Xs = np.random.default_rng(1).uniform(-2.5, 2.5, (37, 4))
    ep = to_engine_program(prog)
    did = engine.dataset(X, y)
    g = engine.score_batch(did, [ep], [th], [tr])
    p, mu, var, dmu = engine.predict_points(did, ep, th, tr, g["tau"][0], g["nu"][0], Xs, want_gradients=True)
    engine.dataset_free(did)
    po, muo, varo = orc.predict(prog, th, X[tr], Xs, g["tau"][0], g["nu"][0])
    assert np.allclose(mu, muo, rtol=1e-8, atol=1e-10) and np.allclose(var, varo, rtol=1e-7, atol=1e-10)
    assert np.allclose(p, po, rtol=1e-8, atol=1e-12)
    h = 1e-6
    for d in range(4):
        e = np.zeros(4)
        e[d] = h
        fd = (orc.predict(prog, th, X[tr], Xs + e, g["tau"][0], g["nu"][0])[1] -
              orc.predict(prog, th, X[tr], Xs - e, g["tau"][0], g["nu"][0])[1]) / (2 * h)
        assert np.allclose(dmu[:, d], fd, rtol=1e-5, atol=1e-7), d


def test_gpy_shim_on_the_engine(engine):
    """The GPy-shaped surface on the real engine: constructor = one inference, optimize() lowers the objective, predict /
    _raw_predict / predictive_gradients shapes as AutoGPC consumes them."""
    from autogpc_b200 import engine as eng_mod
    from autogpc_b200 import gpy_compat as GPy
    X, y = synth(180, 2, 5)
    eng_mod.set_default_engine(engine)
    try:
        k = GPy.kern.RBF(1, variance=1.0, lengthscale=1.0, active_dims=np.array([0])) * \
            GPy.kern.RBF(1, variance=1.0, lengthscale=1.0, active_dims=np.array([1]))
        k.parts[0]["lengthscale"].constrain_bounded(0.05, 12.0, warning=False)
        m = GPy.models.GPClassification(X, y.reshape(-1, 1), kernel=k)
        prog = orc.KernelProgram.from_tree(("*", ("SE", 0), ("SE", 1)))
        nl0 = -m.log_likelihood()
        assert nl0 == pytest.approx(orc.score(prog, m.kern.param_array, X, y, damping=0.0)["nlml"], rel=1e-9)
        m.optimize(max_iters=40)
        assert -m.log_likelihood() < nl0
        Phi, _ = m.predict(X[:9])
        mu, var = m._raw_predict(X[:9])
        dmu, _ = m.predictive_gradients(X[:9])
        assert Phi.shape == (9, 1) and mu.shape == (9, 1) and var.shape == (9, 1) and dmu.shape == (9, 2, 1)
        assert np.all((Phi > 0) & (Phi < 1)) and np.all(var > 0)
        lap = GPy.models.GPClassification(X, y.reshape(-1, 1), kernel=m.kern.copy(),
                                          inference_method=GPy.inference.latent_function_inference.Laplace())
        assert -lap.log_likelihood() == pytest.approx(orc.score_laplace(prog, lap.kern.param_array, X, y)["nlml"], rel=1e-9)
    finally:
        eng_mod.set_default_engine(None)


def test_device_resident_entry_point_matches_host_entry_point(engine):
    """agpc_score_batch_dev (theta / rows / sites / outputs in HBM, caller's stream) against agpc_score_batch."""
    torch = _torch()
    from autogpc_b200.engine import EPOptions, Program
    X, y = synth(300, 4, 2)
    names = ["se0xse1", "mix", "per0"]
    oprogs = [orc.KernelProgram.from_tree(TREES[n]) for n in names]
    progs = [to_engine_program(p) for p in oprogs]
    thetas = [random_theta(p, np.random.default_rng(i)) for i, p in enumerate(oprogs)]
    rows = [np.arange(0, 260, dtype=np.int32), np.arange(20, 300, dtype=np.int32), np.arange(0, 300, 2, dtype=np.int32)]
    did = engine.dataset(X, y)
    host = engine.score_batch(did, progs, thetas, rows, damping=0.0)
    B, stride, n_max = 3, max(p.n_theta for p in progs), max(len(r) for r in rows)
    th = np.zeros((B, stride))
    for b, t in enumerate(thetas):
        th[b, :len(t)] = t
    off = np.zeros(B + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in rows])
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        th_d = torch.from_numpy(th).cuda()
        rows_d = torch.from_numpy(np.concatenate(rows)).cuda()
        tau_d = torch.zeros((B, n_max), dtype=torch.float64, device="cuda")
        nu_d = torch.zeros_like(tau_d)
        nlml_d = torch.zeros(B, dtype=torch.float64, device="cuda")
        grad_d = torch.zeros((B, stride), dtype=torch.float64, device="cuda")
        bic_d, gauss_d = torch.zeros_like(nlml_d), torch.zeros_like(nlml_d)
        sw_d = torch.zeros(B, dtype=torch.int32, device="cuda")
        st_d = torch.zeros(B, dtype=torch.int32, device="cuda")
        engine.score_batch_dev(did, (Program * B)(*progs), B, th_d.data_ptr(), stride, rows_d.data_ptr(), off,
                               EPOptions(1e-6, 0.0, 100, 0, 0, 0), tau_d.data_ptr(), nu_d.data_ptr(), n_max,
                               nlml_d.data_ptr(), grad_d.data_ptr(), bic_d.data_ptr(), gauss_d.data_ptr(), sw_d.data_ptr(),
                               st_d.data_ptr(), st.cuda_stream)
    st.synchronize()
    engine.dataset_free(did)
    assert np.array_equal(nlml_d.cpu().numpy(), host["nlml"]) and np.array_equal(grad_d.cpu().numpy(), host["grad"])
    assert np.array_equal(sw_d.cpu().numpy(), host["sweeps"]) and np.array_equal(st_d.cpu().numpy(), host["status"])
    for b in range(B):
        assert np.array_equal(tau_d[b, :len(rows[b])].cpu().numpy(), host["tau"][b])


def test_gemm_nt_building_block(engine):
    torch = _torch()
    torch.manual_seed(0)
    for (M, N, K, batch, alpha, beta) in [(128, 128, 16, 1, 1.0, 0.0), (256, 384, 160, 2, -1.0, 1.0), (512, 256, 1024, 3, 0.5, 0.25)]:
        A = torch.randn(batch, M, K, dtype=torch.float64, device="cuda")
        Bm = torch.randn(batch, N, K, dtype=torch.float64, device="cuda")
        C = torch.randn(batch, M, N, dtype=torch.float64, device="cuda")
        ref = alpha * A @ Bm.transpose(1, 2) + beta * C
        torch.cuda.synchronize()
        engine.gemm_nt(A.data_ptr(), Bm.data_ptr(), C.data_ptr(), M, N, K, alpha, beta, batch)
        torch.cuda.synchronize()
        assert (C - ref).abs().max().item() <= 1e-13 * ref.abs().max().item() * K ** 0.5


def test_profile_counters_book_the_launches(engine):
    """agpc_profile_enable / agpc_profile_read: per-class event time, algorithmic work and launch counts."""
    X, y = synth(300, 2, 1)
    prog = to_engine_program(orc.KernelProgram.from_tree(TREES["se0xse1"]))
    did = engine.dataset(X, y)
    engine.profile_enable(True)
    l0 = engine.launches
    engine.score_batch(did, [prog], [np.ones(4)], [np.arange(300, dtype=np.int32)])
    prof = engine.profile_read()
    engine.profile_enable(False)
    engine.dataset_free(did)
    assert sum(v["launches"] for k, v in prof.items() if k != "gemm" and not k.startswith("oz_")) <= engine.launches - l0
    assert prof["kbuild"]["launches"] == 1 and prof["kbuild"]["work"] == pytest.approx(8.0 * 300 * 300)
    assert prof["potf2"]["launches"] >= 3 and prof["gemm"]["launches"] > 0 and prof["gemm"]["ms"] > 0
    assert prof["gemm"]["work"] == pytest.approx(prof["gemm_potrf"]["work"] + prof["gemm_trtri"]["work"] + prof["gemm_lauum"]["work"])
    assert prof["gemm_lauum"]["work"] == pytest.approx(300.0 ** 3 / 3.0)


def test_chunked_batches_equal_one_chunk(engine):
    """A workspace limit that holds ~3 items forces a 10-item batch through several lock-step chunks (the path a
    360-item search depth takes at n = 6400): results must be bit-identical to the single-chunk run."""
    from autogpc_b200.engine import Engine
    X, y = synth(400, 4, 14)
    rng = np.random.default_rng(3)
    progs, thetas, rows = [], [], []
    for i, name in enumerate(["se0xse1", "per0", "mix", "se0+se1", "se0xse1+per3+se2"] * 2):
        prog = orc.KernelProgram.from_tree(TREES[name])
        progs.append(to_engine_program(prog))
        thetas.append(random_theta(prog, rng))
        rows.append(np.sort(rng.choice(400, size=300 + 7 * i, replace=False)).astype(np.int32))
    did = engine.dataset(X, y)
    one = engine.score_batch(did, progs, thetas, rows, damping=0.0)
    engine.dataset_free(did)
    per_item = 5 * 384 * 384 * 8 + 64 * 1024          # five padded n x n matrices + vectors
    saved = {k: os.environ.get(k) for k in engine.test_env}
    os.environ.update(engine.test_env)          # same kernel-selection rule as the fixture's engine
    try:
        small = Engine(0, workspace_limit_bytes=int(3.5 * per_item))
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    try:
        did2 = small.dataset(X, y)
        many = small.score_batch(did2, progs, thetas, rows, damping=0.0)
        small.dataset_free(did2)
    finally:
        small.close()
    assert np.array_equal(one["status"], many["status"]) and np.array_equal(one["sweeps"], many["sweeps"])
    assert np.array_equal(one["nlml"], many["nlml"]) and np.array_equal(one["grad"], many["grad"])
    for a, b in zip(one["tau"], many["tau"]):
        assert np.array_equal(a, b)


def test_not_pd_is_per_item_status(engine):
    """A hopeless item (negative variance) must not abort the batch (gpckernel.py:196-199 semantics)."""
    X, y = synth(200, 2, 1)
    prog = orc.KernelProgram.from_tree(TREES["se0xse1"])
    good = np.array([1.0, 1.0, 1.0, 1.0])
    bad = np.array([-50.0, 1.0, 1.0, 1.0])
    rows = [np.arange(200, dtype=np.int32)] * 2
    did = engine.dataset(X, y)
    ep = to_engine_program(prog)
    g = engine.score_batch(did, [ep, ep], [good, bad], rows)
    engine.dataset_free(did)
    o = orc.score(prog, good, X, y)
    assert g["status"][0] == 0 and abs(g["nlml"][0] - o["nlml"]) <= 1e-9 * abs(o["nlml"])
    assert g["status"][1] in (2, 4) and not np.isfinite(g["nlml"][1])


# ---- committed golden fixtures (tests/golden/*.npz, generated by tests/golden/make_golden.py from the oracle) ----
import ast
import glob
import os

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_fixtures_through_c_abi(engine, path):
    """The five BASELINE configs at reduced n + edge cases (constant-only kernel, n = 129 not a tile multiple)."""
    g = np.load(path)
    prog = orc.KernelProgram.from_tree(ast.literal_eval(str(g["tree"])))
    X, y, N = g["X"], g["y"], g["X"].shape[0]
    Xall = np.vstack([X, g["Xstar"]])
    yall = np.concatenate([y, np.zeros(len(g["Xstar"]))])
    did = engine.dataset(Xall, yall)
    ep = to_engine_program(prog)
    rows = [np.arange(N, dtype=np.int32)]
    r = engine.score_batch(did, [ep], [g["theta"]], rows)
    p, mu, var, _ = engine.predict_batch(did, [ep], [g["theta"]], rows, r["tau"], r["nu"],
                                         [np.arange(N, len(Xall), dtype=np.int32)], want_latent=True)
    engine.dataset_free(did)
    assert r["status"][0] == int(g["status"]) and r["sweeps"][0] == int(g["sweeps"])
    assert abs(r["nlml"][0] - g["nlml"]) <= 1e-9 * abs(g["nlml"])
    assert abs(r["bic"][0] - g["bic"]) <= 1e-9 * abs(g["bic"])
    assert abs(r["nlml_gauss"][0] - g["nlml_gauss"]) <= 1e-8 * abs(g["nlml_gauss"])
    scale = np.maximum(np.abs(g["grad"]), 1e-3 * np.abs(g["grad"]).max())
    assert np.all(np.abs(r["grad"][0, :prog.n_theta] - g["grad"]) <= 1e-7 * scale)
    assert np.allclose(r["tau"][0], g["tau"], rtol=1e-7, atol=1e-12)
    assert np.allclose(p[0], g["pstar"], rtol=1e-8, atol=1e-12)
    assert np.allclose(mu[0], g["mustar"], rtol=1e-8, atol=1e-10)
    assert np.allclose(var[0], g["varstar"], rtol=1e-7, atol=1e-10)


# ---- building blocks on device buffers (agpc_build_K / agpc_kernel_grad / agpc_potrf / agpc_gemm_nt) --------------
def _torch():
    import torch
    return torch


@pytest.mark.parametrize("name", ["se0xse1+per3+se2", "mix", "const"])
@pytest.mark.parametrize("n", [130, 512])
def test_build_K_dK_and_fused_gradient(engine, name, n):
    """kern.K, materialised dK/dtheta and the fused <G, dK/dtheta> contraction against the oracle (abs 1e-13)."""
    torch = _torch()
    X, _ = synth(n, 4, 21)
    prog = orc.KernelProgram.from_tree(TREES[name])
    th = random_theta(prog, np.random.default_rng(4))
    ep = to_engine_program(prog)
    Xd = torch.from_numpy(X).cuda()
    ld = n + (n & 1)
    K = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
    dK = torch.zeros((prog.n_theta, n, ld), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()   # the engine's stream is not ordered with torch's: buffers must be ready before the call
    engine.build_K(ep, th, Xd.data_ptr(), n, 4, K.data_ptr(), ld, dK.data_ptr())
    torch.cuda.synchronize()
    Ko = orc.kern_K(prog, th, X)
    dKo = orc.kern_dK(prog, th, X)
    assert np.abs(K.cpu().numpy()[:, :n] - Ko).max() <= 1e-13 * max(1.0, np.abs(Ko).max())
    for q in range(prog.n_theta):
        assert np.abs(dK[q].cpu().numpy()[:, :n] - dKo[q]).max() <= 1e-12 * max(1.0, np.abs(dKo[q]).max())
    K2 = torch.zeros_like(K)
    torch.cuda.synchronize()
    engine.build_K(ep, th, Xd.data_ptr(), n, 4, K2.data_ptr(), ld)      # K-only path (symmetric tiles)
    torch.cuda.synchronize()
    assert np.abs(K2.cpu().numpy()[:, :n] - Ko).max() <= 1e-13 * max(1.0, np.abs(Ko).max())
    assert torch.equal(K2[:, :n], K2[:, :n].T.contiguous())             # exactly symmetric
    G = np.random.default_rng(5).standard_normal((n, n))
    Gd = torch.from_numpy(G).cuda()
    torch.cuda.synchronize()
    g = engine.kernel_grad(ep, th, Xd.data_ptr(), n, 4, Gd.data_ptr(), n)
    go = orc.kern_grad(prog, th, X, G)
    assert np.all(np.abs(g - go) <= 1e-10 * np.maximum(np.abs(go), 1e-3 * np.abs(go).max() + 1e-30))


@pytest.mark.parametrize("n,batch", [(128, 1), (384, 3), (1152, 2), (2048, 1)])
def test_potrf_trtri_against_lapack(engine, n, batch):
    """jitchol / dtrtri stand-ins: L and L^-1 against NumPy's LAPACK on the same matrices."""
    torch = _torch()
    rng = np.random.default_rng(n)
    G = rng.standard_normal((batch, n, n))
    S = G @ G.transpose(0, 2, 1) / n + np.eye(n)
    A = torch.from_numpy(S).cuda().contiguous()
    Mi, Ui = torch.zeros_like(A), torch.zeros_like(A)
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()   # see test_build_K_dK_and_fused_gradient
    engine.potrf(A.data_ptr(), n, batch, Mi.data_ptr(), Ui.data_ptr(), info.data_ptr())
    torch.cuda.synchronize()
    assert info.cpu().tolist() == [0] * batch
    for b in range(batch):
        Lr = np.linalg.cholesky(S[b])
        assert np.abs(np.tril(A[b].cpu().numpy()) - Lr).max() <= 1e-12
        Wr = np.linalg.inv(Lr)
        assert np.abs(np.tril(Mi[b].cpu().numpy()) - Wr).max() <= 1e-11
        assert np.abs(np.triu(Ui[b].cpu().numpy()) - Wr.T).max() <= 1e-11


@pytest.mark.parametrize("n,batch", [(1024, 1), (1152, 1), (1152, 5), (2304, 2), (2304, 9), (1280, 17)])
def test_potrf_paths_around_the_scheduling_thresholds(engine, n, batch):
    """8 vs 9 panels (look-ahead off / on), thin 128 x 32 tiles vs full tiles (CTA-count threshold), partial outer
    blocks; every path must give the same factor.  Checked through L L^T = B and M L = I on random SPD matrices."""
    torch = _torch()
    torch.manual_seed(n + batch)
    G = torch.randn(batch, n, n, dtype=torch.float64, device="cuda")
    S = G @ G.transpose(1, 2) / n + torch.eye(n, dtype=torch.float64, device="cuda")
    A = S.clone()
    Mi, Ui = torch.zeros_like(A), torch.zeros_like(A)
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    for _ in range(2):                                   # twice: scratch regions and events are reused across calls
        A.copy_(S)
        torch.cuda.synchronize()
        engine.potrf(A.data_ptr(), n, batch, Mi.data_ptr(), Ui.data_ptr(), info.data_ptr())
        torch.cuda.synchronize()
    assert info.cpu().tolist() == [0] * batch
    L = torch.tril(A)
    eye = torch.eye(n, dtype=torch.float64, device="cuda")
    assert (L @ L.transpose(1, 2) - S).abs().max().item() <= 1e-12 * n ** 0.5 * S.abs().max().item()
    assert (torch.tril(Mi) @ L - eye).abs().max().item() <= 1e-11
    assert torch.equal(torch.triu(Ui), torch.tril(Mi).transpose(1, 2).contiguous())


def test_large_n_roundtrip_properties(engine):
    """Size-independent properties at a size the oracle would take minutes for: L L^T = B, M L = I (n = 4096)."""
    torch = _torch()
    n = 4096
    torch.manual_seed(0)
    G = torch.randn(n, n, dtype=torch.float64, device="cuda")
    S = G @ G.T / n + torch.eye(n, dtype=torch.float64, device="cuda")
    A = S.clone()
    Mi, Ui = torch.zeros_like(A), torch.zeros_like(A)
    torch.cuda.synchronize()
    engine.potrf(A.data_ptr(), n, 1, Mi.data_ptr(), Ui.data_ptr(), 0)
    torch.cuda.synchronize()
    L = torch.tril(A)
    assert (L @ L.T - S).abs().max().item() <= 1e-12 * S.abs().max().item() * n ** 0.5
    assert (torch.tril(Mi) @ L - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max().item() <= 1e-11
    assert torch.equal(torch.triu(Ui), torch.tril(Mi).T.contiguous())


# ---- selection parity: the same search on the CUDA engine and on the oracle-backed engine -----------------------
def test_search_selects_identical_structures_as_oracle(engine):
    """north_star: identical selected kernel structure and ordering per search depth.

    Two parts.  (1) Twin search: the same GPCSearch driven by the CUDA engine and by the oracle-backed engine
    (tests/fake_engine.OracleEngine; host, grammar, restarts and L-BFGS-B shared) must select the same kernel at every
    depth.  (2) Replay: L-BFGS-B amplifies last-digit differences of the objective, so the *full* ranking is compared at
    identical hyper-parameters -- every candidate the GPU search scored is re-scored by the oracle at the GPU's theta*
    (BIC of the median fold model; CV error over all fold models) and the oracle's ranking of those scores must be the
    GPU's ranking, position by position."""
    from autogpc_b200 import engine as eng_mod
    from autogpc_b200 import gpckernel
    from autogpc_b200.gpcdata import GPCData
    from autogpc_b200.gpcsearch import GPCSearch
    from tests.fake_engine import OracleEngine, _to_oracle

    n_points, folds = 150, 3
    X, y = synth(n_points, 3, 8)

    def run(e, criterion):
        eng_mod.set_default_engine(e)
        gpckernel.set_rng(5)
        d = GPCData(X, y.reshape(-1, 1), seed=1)
        s = GPCSearch(d, base_kernels="SE,Lin,Per", max_depth=2, beam_width=1, criterion=criterion, n_folds=folds,
                      max_evals=60)
        best, _ = s.search()
        return s, [k.kernel.structure() for k in best]

    def oracle_fold(m):
        prog = _to_oracle(m._program)
        th = m.kern.param_array
        o = orc.score(prog, th, X[m._rows], y[m._rows], p_eff=m._program.p_eff, damping=0.0)   # cold EP at theta*
        p, _, _ = orc.predict(prog, th, X[m._rows], X[m._test_rows], o["tau"], o["v"])
        margin = np.abs(p - 0.5)
        err = float(np.mean((p - 0.5) * (y[m._test_rows] - 0.5) < 0))
        return o, err, int(np.sum(margin < 1e-5))

    try:
        for criterion in ("bic", "error"):
            sg, bg = run(engine, criterion)
            so, bo = run(OracleEngine(), criterion)
            assert bg == bo, (criterion, bg, bo)                                  # (1) identical selection per depth
            for lvl in sg.history:                                                # (2) replay at the GPU's theta*
                gpu_scores, orc_scores, names = [], [], []
                for k in lvl:
                    names.append(k.kernel.structure())
                    gpu_scores.append(sg.score(k))
                    if criterion == "bic":
                        o, _, _ = oracle_fold(k.model)
                        assert abs(o["bic"] - k.model.bic) <= 1e-5 * abs(o["bic"]), (names[-1], o["bic"], k.model.bic)
                        orc_scores.append(o["bic"])
                    else:
                        errs, shaky = [], 0
                        for m in k.foldModels:
                            o, e, nshaky = oracle_fold(m)
                            assert abs(o["nlml"] - m._nlml) <= 1e-5 * abs(o["nlml"])   # warm- vs cold-started EP, eps 1e-6
                            errs.append(e)
                            shaky += nshaky
                        orc_scores.append(float(np.mean(errs)))
                        assert abs(orc_scores[-1] - k.errorRate) <= (shaky + 1e-9) / (len(errs) * 1.0 * n_points / folds) + 1e-12, names[-1]
                order_g = sorted(range(len(names)), key=lambda i: (gpu_scores[i], names[i]))
                order_o = sorted(range(len(names)), key=lambda i: (orc_scores[i], names[i]))
                assert order_g == list(range(len(names)))                        # history is stored in ranked order
                for a, b in zip(order_g, order_o):                                # same ranking, modulo exact-ish ties
                    assert a == b or abs(orc_scores[a] - orc_scores[b]) <= 1e-5 * max(1.0, abs(orc_scores[a])), \
                        (criterion, names[a], names[b], orc_scores[a], orc_scores[b])
    finally:
        eng_mod.set_default_engine(None)


# ---- edge cases: tiny / degenerate inputs, limits, API misuse -------------------------------------------------------
def test_edge_cases_tiny_n_single_dimension_and_one_class(engine):
    """n = 3 (far below one 128-tile), D = 1, and a fold whose labels are all one class."""
    rng = np.random.default_rng(0)
    X = rng.uniform(-2, 2, (40, 1))
    y = (X[:, 0] > 0).astype(float)
    prog = orc.KernelProgram.from_tree(("+", ("SE", 0), ("CONST",)))
    ep = to_engine_program(prog)
    th = np.array([1.2, 0.7, 0.5])
    rows = [np.array([0, 1, 2], dtype=np.int32), np.arange(40, dtype=np.int32),
            np.nonzero(y == 1)[0].astype(np.int32)]
    did = engine.dataset(X, y)
    g = engine.score_batch(did, [ep] * 3, [th] * 3, rows)
    engine.dataset_free(did)
    for b, r in enumerate(rows):
        o = orc.score(prog, th, X[r], y[r])
        assert g["status"][b] == o["status"] == 0 and g["sweeps"][b] == o["sweeps"]
        assert abs(g["nlml"][b] - o["nlml"]) <= 1e-9 * max(1.0, abs(o["nlml"]))
        assert np.allclose(g["grad"][b, :3], o["grad"], rtol=1e-6, atol=1e-9)


def test_api_misuse_is_an_error_code_not_a_crash(engine):
    from autogpc_b200.engine import EngineError, Program, LEAF_SE
    X, y = synth(50, 2, 0)
    prog = to_engine_program(orc.KernelProgram.from_tree(TREES["se0xse1"]))
    th = np.ones(4)
    did = engine.dataset(X, y)
    with pytest.raises(EngineError):            # row index outside the dataset
        engine.score_batch(did, [prog], [th], [np.array([0, 1, 50], dtype=np.int32)])
    with pytest.raises(EngineError):            # unknown dataset
        engine.score_batch(did + 1000, [prog], [th], [np.arange(10, dtype=np.int32)])
    with pytest.raises(EngineError):            # theta of the wrong length
        engine.score_batch(did, [prog], [np.ones(3)], [np.arange(10, dtype=np.int32)])
    with pytest.raises(EngineError):            # leaf on a dimension the data does not have
        bad = Program.build([LEAF_SE], [5], [0], [[0]], 2)
        engine.score_batch(did, [bad], [np.ones(2)], [np.arange(10, dtype=np.int32)])
    with pytest.raises(EngineError):            # more leaves than AGPC_MAX_LEAVES
        Program.build([LEAF_SE] * 17, [0] * 17, list(range(0, 34, 2)), [[i] for i in range(16)], 34)
    ok = engine.score_batch(did, [prog], [th], [np.arange(50, dtype=np.int32)])   # the context is still usable
    engine.dataset_free(did)
    assert ok["status"][0] == 0 and np.isfinite(ok["nlml"][0])


# ---- randomized sweep: random grammar-shaped programs, sizes and folds in one batch -----------------------------------
def _random_tree(rng, D, depth):
    if depth == 0 or rng.random() < 0.35:
        kind = rng.choice(["SE", "PER", "LIN", "CONST"], p=[0.45, 0.25, 0.2, 0.1])
        return ("CONST",) if kind == "CONST" else (str(kind), int(rng.integers(D)))
    op = "+" if rng.random() < 0.5 else "*"
    return (op,) + tuple(_random_tree(rng, D, depth - 1) for _ in range(int(rng.integers(2, 4))))


def test_randomized_programs_batch(engine):
    """24 random sum / product trees (up to 3 levels, shared leaves after distribution included), random theta, random
    ragged row subsets, scored in ONE batch with the production EP schedule: NLML, gradient, BIC and sites against the oracle."""
    rng = np.random.default_rng(2024)
    D = 5
    X, y = synth(420, D, 31)
    progs, oprogs, thetas, rows = [], [], [], []
    while len(progs) < 24:
        prog = orc.KernelProgram.from_tree(_random_tree(rng, D, 3))
        if len(prog.leaf_type) > 16 or len(prog.terms) > 16 or max(len(t) for t in prog.terms) > 8 or prog.n_theta > 48:
            continue
        if all(t == orc.LEAF_CONST for t in prog.leaf_type) and len(prog.leaf_type) > 1:
            continue
        oprogs.append(prog)
        progs.append(to_engine_program(prog))
        thetas.append(random_theta(prog, rng))
        n = int(rng.integers(60, 400))
        rows.append(np.sort(rng.choice(420, size=n, replace=False)).astype(np.int32))
    did = engine.dataset(X, y)
    g = engine.score_batch(did, progs, thetas, rows, damping=0.0)
    engine.dataset_free(did)
    for b, (prog, th, r) in enumerate(zip(oprogs, thetas, rows)):
        o = orc.score(prog, th, X[r], y[r], damping=0.0)
        assert g["status"][b] == o["status"], (b, g["status"][b], o["status"])
        if o["status"] != 0:
            continue
        assert g["sweeps"][b] == o["sweeps"], b
        assert abs(g["nlml"][b] - o["nlml"]) <= 1e-8 * abs(o["nlml"]), (b, g["nlml"][b], o["nlml"])
        assert abs(g["bic"][b] - o["bic"]) <= 1e-8 * abs(o["bic"])
        scale = np.maximum(np.abs(o["grad"]), 1e-3 * np.abs(o["grad"]).max() + 1e-12)
        assert np.all(np.abs(g["grad"][b, :prog.n_theta] - o["grad"]) <= 1e-6 * scale), b
        assert np.allclose(g["tau"][b], o["tau"], rtol=1e-6, atol=1e-12), b


def test_baseline_config0_single_se_ard_n500(engine):
    """BASELINE.json configs[0]: one SE-ARD (SE0 x SE1, gpckerneltest.py:53,80-81 spelling) GPClassification on synthetic
    2-D data, n = 500, no folds, one full L-BFGS-B optimisation (gpckernel.py:245-246) -- through the GPy-shaped shim on
    the engine against the oracle's restatement of the same optimisation: NLML and gradient at theta0 and at theta*."""
    from autogpc_b200 import engine as eng_mod, flexible_function as ff, gpckernel
    from autogpc_b200 import gpy_compat as GPy
    from autogpc_b200.gpcdata import GPCData
    X, y = synth(500, 2, 0)
    d = GPCData(X, y.reshape(-1, 1))
    prog = orc.KernelProgram.from_tree(("*", ("SE", 0), ("SE", 1)))
    eprog = to_engine_program(prog)
    eng_mod.set_default_engine(engine)
    try:
        k = ff.ProductKernel([ff.SqExpKernel(0, 1.0, 1.0), ff.SqExpKernel(1, 1.5, 0.8)])
        m = GPy.models.GPClassification(X, y.reshape(-1, 1), kernel=gpckernel.gpss2gpy(k, data=d))
        th0 = m.kern.param_array.copy()
        m.optimize()
        th1 = m.kern.param_array.copy()
        nl1 = -m.log_likelihood()
    finally:
        eng_mod.set_default_engine(None)
    lo, hi = d.getLengthscaleBounds()
    tr = orc.Transforms(np.array([0, 1, 0, 1]), np.array([0, lo[0], 0, lo[1]]), np.array([0, hi[0], 0, hi[1]]))
    rows = [np.arange(500, dtype=np.int32)]
    did = engine.dataset(X, y)
    for th in (th0, th1):                     # cold evaluations on both sides at theta0 and at the engine's theta*
        g = engine.score_batch(did, [eprog], [th], rows, damping=0.0)     # the shim's (production) EP schedule
        o = orc.score(prog, th, X, y, damping=0.0)
        assert g["status"][0] == 0 and o["status"] == 0 and g["sweeps"][0] == o["sweeps"]
        assert abs(g["nlml"][0] - o["nlml"]) <= 1e-9 * abs(o["nlml"])
        # at theta* the gradient is the ~1e-4 remainder of O(|nlml|) terms that cancel: absolute floor from that scale
        assert np.all(np.abs(g["grad"][0, :4] - o["grad"]) <= 1e-6 * np.abs(o["grad"]) + 1e-9 * abs(o["nlml"]))
        assert abs(g["bic"][0] - (2 * o["nlml"] + 3 * np.log(500))) <= 1e-8 * abs(g["bic"][0])
    engine.dataset_free(did)
    ref = orc.optimize(prog, th0, X, y, tr, damping=0.0)
    assert nl1 == pytest.approx(ref["nlml"], rel=1e-6)       # both optimisers stop at the same optimum
    assert np.allclose(th1, ref["theta"], rtol=1e-4)


def test_wide_inputs_odd_dimension_count(engine):
    """D = 33 input columns (odd, wider than a warp; the X panels of a tile are staged per dimension): kernels on the
    first, a middle and the last column, with and without materialised dK."""
    rng = np.random.default_rng(77)
    N, D = 333, 33
    X = rng.uniform(-3, 3, (N, D))
    y = (np.sin(X[:, 0]) + 0.5 * X[:, 16] * X[:, 32] + 0.3 * rng.standard_normal(N) > 0).astype(float)
    prog = orc.KernelProgram.from_tree(("+", ("*", ("SE", 0), ("PER", 16)), ("LIN", 32), ("SE", 32)))
    th = random_theta(prog, rng)
    did = engine.dataset(X, y)
    rows = [np.arange(N, dtype=np.int32), np.arange(5, 300, dtype=np.int32)]
    g = engine.score_batch(did, [to_engine_program(prog)] * 2, [th, th * 1.1], rows, damping=0.0)
    engine.dataset_free(did)
    for b, t in enumerate((th, th * 1.1)):
        o = orc.score(prog, t, X[rows[b]], y[rows[b]], damping=0.0)
        assert g["status"][b] == o["status"] == 0 and g["sweeps"][b] == o["sweeps"]
        assert abs(g["nlml"][b] - o["nlml"]) <= 1e-9 * abs(o["nlml"])
        scale = np.maximum(np.abs(o["grad"]), 1e-3 * np.abs(o["grad"]).max())
        assert np.all(np.abs(g["grad"][b, :prog.n_theta] - o["grad"]) <= 1e-6 * scale)


def test_posterior_reuse_keeps_converged_sites(engine):
    """EP posterior reuse (ep.cu EP_REUSE_FACTOR / oracle.ep_parallel): warm-started at the fixed point the residual at
    the supplied sites is far below the tolerance, so the evaluation stops after one sweep, returns the sites bit-for-bit
    and skips the final re-factorisation -- and still reports the same NLML / gradient as a re-factorised evaluation
    (frozen sites) at those sites.  A batch mixes such an item with one that has to iterate."""
    X, y = synth(300, 3, 12)
    prog = orc.KernelProgram.from_tree(TREES["se0xse1"])
    ep = to_engine_program(prog)
    th = np.array([1.4, 0.8, 0.9, 1.1])
    rows = [np.arange(300, dtype=np.int32)]
    did = engine.dataset(X, y)
    tight = engine.score_batch(did, [ep], [th], rows, epsilon=1e-18, max_sweeps=300, damping=0.0)
    l0 = engine.launches
    again = engine.score_batch(did, [ep], [th], rows, tau=tight["tau"], nu=tight["nu"], damping=0.0)
    l_reuse = engine.launches - l0
    l0 = engine.launches
    frozen = engine.score_batch(did, [ep], [th], rows, tau=tight["tau"], nu=tight["nu"], frozen_sites=True)
    l_frozen = engine.launches - l0
    assert again["status"][0] == 0 and again["sweeps"][0] == 1
    assert np.array_equal(again["tau"][0], tight["tau"][0]) and np.array_equal(again["nu"][0], tight["nu"][0])
    assert abs(again["nlml"][0] - frozen["nlml"][0]) <= 1e-12 * abs(frozen["nlml"][0])
    assert np.allclose(again["grad"][0, :4], frozen["grad"][0, :4], rtol=1e-10, atol=1e-12)
    assert l_reuse <= l_frozen + 8          # one factorisation in either call: the converged sweep's is the final one
    o = orc.score(prog, th, X, y, damping=0.0, tau0=tight["tau"][0], v0=tight["nu"][0])
    assert o["sweeps"] == 1 and np.array_equal(o["tau"], tight["tau"][0])
    assert abs(again["nlml"][0] - o["nlml"]) <= 1e-10 * abs(o["nlml"])
    # mixed batch: item 0 reuses, item 1 (different theta, same warm sites) must iterate and re-factorise
    mixed = engine.score_batch(did, [ep, ep], [th, th * 1.3], rows * 2, tau=[tight["tau"][0]] * 2, nu=[tight["nu"][0]] * 2,
                               damping=0.0)
    engine.dataset_free(did)
    assert mixed["sweeps"][0] == 1 and mixed["sweeps"][1] > 1
    assert np.array_equal(mixed["tau"][0], tight["tau"][0])
    o1 = orc.score(prog, th * 1.3, X, y, damping=0.0, tau0=tight["tau"][0], v0=tight["nu"][0])
    assert mixed["sweeps"][1] == o1["sweeps"] and abs(mixed["nlml"][1] - o1["nlml"]) <= 1e-9 * abs(o1["nlml"])
    assert abs(mixed["nlml"][0] - again["nlml"][0]) <= 1e-12 * abs(again["nlml"][0])


def test_cumulate_additive_kernels_on_engine_matches_oracle_backed_run(engine):
    """Second caller of the engine (gpcreport.py:657-680): un-optimised full-data inferences of cum + k for
    all remaining components in one batch (no gradient), training errors by one predict_batch -- same component order,
    errors and NLMLs as the identical host code running on the oracle-backed stand-in engine."""
    from autogpc_b200 import engine as eng_mod, flexible_function as ff, gpckernel
    from autogpc_b200.gpcdata import GPCData
    from tests.fake_engine import OracleEngine
    X, y = synth(220, 3, 6)
    out = []
    for e in (engine, OracleEngine()):
        eng_mod.set_default_engine(e)
        try:
            gpckernel.set_rng(4)
            d = GPCData(X, y.reshape(-1, 1))
            parent = gpckernel.GPCKernel(ff.SumKernel([ff.SqExpKernel(0, 1.0, 1.2), ff.SqExpKernel(1, 1.5, 0.7),
                                                       ff.ProductKernel([ff.SqExpKernel(2, 0.8, 1.0),
                                                                         ff.SqExpKernel(0, 1.0, 2.0)])]), d, depth=3)
            kers, cums = gpckernel.cumulateAdditiveKernels(parent.toSummands())
            out.append(([k.kernel.structure() for k in kers], [c.kernel.structure() for c in cums],
                        [c.error() for c in cums], [c.getNLML() for c in cums]))
        finally:
            eng_mod.set_default_engine(None)
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1]
    assert np.allclose(out[0][2], out[1][2]) and np.allclose(out[0][3], out[1][3], rtol=1e-8)


@pytest.mark.gpu
def test_jitchol_building_block_matches_gpy_ladder(engine):
    """agpc_jitchol against the oracle's restatement of GPy util/linalg.py jitchol (reached from gpckernel.py:245-246):
    a healthy matrix needs no jitter; a numerically rank-deficient one with a slightly negative spectrum is recovered
    with exactly jitter = mean(diag) 1e-6 10^k for the same k; a non-positive diagonal fails at once."""
    torch = _torch()
    rng = np.random.default_rng(5)
    n = 256
    G = rng.standard_normal((n, n))
    good = G @ G.T / n + np.eye(n)
    W = rng.standard_normal((n, 8))
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    low = W @ W.T                                            # rank 8 ...
    bad1 = low - 3e-6 * np.mean(np.diag(low)) * (Q[:, :4] @ Q[:, :4].T)     # ... minus a few 3e-6-relative directions
    bad2 = low - 5e-4 * np.mean(np.diag(low)) * (Q[:, :4] @ Q[:, :4].T)     # needs three decades of jitter
    neg = good.copy()
    neg[7, 7] = -1.0
    mats = [good, bad1, bad2, neg]
    A = torch.from_numpy(np.stack(mats)).cuda()
    jit = torch.zeros(len(mats), dtype=torch.float64, device="cuda")
    info = torch.zeros(len(mats), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    engine.jitchol(A.data_ptr(), n, len(mats), jit.data_ptr(), info.data_ptr())
    torch.cuda.synchronize()
    jit, info, L = jit.cpu().numpy(), info.cpu().numpy(), np.tril(A.cpu().numpy())
    for b, Mx in enumerate(mats[:3]):
        Lo, jo = orc.jitchol(Mx)
        assert info[b] == 0
        assert jit[b] == pytest.approx(jo, rel=1e-12, abs=0.0) and (jo > 0) == (b > 0)
        assert np.abs(L[b] @ L[b].T - (Mx + jo * np.eye(n))).max() <= 1e-12 * np.abs(Mx).max()
    assert jit[2] > 50 * jit[1]
    assert info[3] == orc.ST_NOT_PD
    with pytest.raises(Exception):
        orc.jitchol(neg)


@pytest.mark.gpu
def test_jitter_ladder_inside_score_batch():
    """The same ladder inside score_batch (jitchol of B = I + S^1/2 K S^1/2, gpckernel.py:245-246; what the bare except
    at :196-199 tolerates).  B is built from PSD kernels and bounded below by I, so it cannot be made to fail reliably
    (an indefiniteness at rounding level passes LAPACK's dpotrf too); the failure is injected: AGPC_FAULT_NOT_PD=2 makes
    the plain attempt and the first jittered attempt of item 0 report NOT_PD, so the ladder must succeed on its second
    rung (jitter = mean(diag B) 1e-5), flag AGPC_ST_JITTER, move the NLML by a relative ~1e-5 and leave item 1 alone.
    With the ladder switched off the same fault leaves the item NOT_PD."""
    import subprocess
    import sys
    code = r'''
import json, os, sys
import numpy as np
sys.path.insert(0, %r)
from autogpc_b200.engine import Engine
from oracle import gpc_oracle as orc
from tests.util import TREES, synth, to_engine_program
X, y = synth(300, 2, 17)
prog = orc.KernelProgram.from_tree(TREES["se0xse1"])
ep = to_engine_program(prog)
th = [np.array([1.3, 0.9, 0.8, 1.2]), np.array([0.7, 1.1, 1.5, 0.6])]
rows = [np.arange(300, dtype=np.int32)] * 2
e = Engine(0)
d = e.dataset(X, y)
r = e.score_batch(d, [ep, ep], th, rows)
print(json.dumps({"status": r["status"].tolist(), "nlml": r["nlml"].tolist(), "sweeps": r["sweeps"].tolist()}))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def run(env):
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        return json.loads(out.stdout.strip().splitlines()[-1])
    import json
    clean = run({})
    hit = run({"AGPC_FAULT_NOT_PD": "2"})
    off = run({"AGPC_FAULT_NOT_PD": "2", "AGPC_JITTER": "0"})
    assert clean["status"] == [0, 0]
    assert hit["status"] == [orc.ST_JITTER, 0] and hit["nlml"][1] == clean["nlml"][1]
    rel = abs(hit["nlml"][0] - clean["nlml"][0]) / abs(clean["nlml"][0])
    assert 0 < rel < 1e-3
    assert off["status"][0] == orc.ST_NOT_PD and off["status"][1] == 0 and off["nlml"][1] == clean["nlml"][1]


@pytest.mark.gpu
def test_gemm_nt_i8x_building_block(engine):
    """The sliced-integer tensor-core product (csrc/ozaki.cu, tcgen05 kind::i8) against torch FP64: with S = 7 digit slices
    it must pass the FP64 DMMA kernel's own bound (test_gemm_nt_building_block) with two orders of magnitude to spare --
    also when the rows differ by e^(+-9) in magnitude and some are exactly zero -- and S = 6 stays below 1e-12.  A
    non-finite operand entry poisons its row of C, as it would in FP64."""
    torch = _torch()
    torch.manual_seed(3)
    for S, tol in ((7, 1e-15), (6, 1e-12)):
        for (M, N, K, batch, alpha, beta) in [(128, 128, 32, 1, 1.0, 0.0), (256, 384, 160, 2, -1.0, 1.0), (512, 256, 1024, 3, 0.5, 0.25)]:
            A = torch.randn(batch, M, K, dtype=torch.float64, device="cuda")
            Bm = torch.randn(batch, N, K, dtype=torch.float64, device="cuda")
            C = torch.randn(batch, M, N, dtype=torch.float64, device="cuda")
            ref = alpha * A @ Bm.transpose(1, 2) + beta * C
            torch.cuda.synchronize()
            engine.gemm_nt_i8x(A.data_ptr(), Bm.data_ptr(), C.data_ptr(), M, N, K, alpha, beta, batch, S)
            torch.cuda.synchronize()
            assert (C - ref).abs().max().item() <= tol * ref.abs().max().item() * K ** 0.5, (S, M, N, K)
    # rows of very different magnitude, exact zero rows: error against the natural scale of each entry
    A = torch.randn(2, 256, 512, dtype=torch.float64, device="cuda") * torch.exp(3.0 * torch.randn(2, 256, 1, dtype=torch.float64, device="cuda"))
    Bm = torch.randn(2, 128, 512, dtype=torch.float64, device="cuda")
    Bm[:, 5, :] = 0.0
    C = torch.zeros(2, 256, 128, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    engine.gemm_nt_i8x(A.data_ptr(), Bm.data_ptr(), C.data_ptr(), 256, 128, 512, 1.0, 0.0, 2, 7)
    torch.cuda.synchronize()
    ref = A @ Bm.transpose(1, 2)
    scale = A.abs().amax(2, keepdim=True) * Bm.abs().amax(2, keepdim=True).transpose(1, 2) * 512 ** 0.5 + 1e-300
    assert ((C - ref).abs() / scale).max().item() <= 3e-14
    assert (C[:, :, 5] == 0).all()
    A[0, 7, 100] = float("nan")
    engine.gemm_nt_i8x(A.data_ptr(), Bm.data_ptr(), C.data_ptr(), 256, 128, 512, 1.0, 0.0, 2, 7)
    torch.cuda.synchronize()
    assert torch.isnan(C[0, 7]).all() and torch.isfinite(C[0, 8]).all() and torch.isfinite(C[1]).all()


@pytest.mark.gpu
def test_int8_and_fp64_factorisation_paths_agree():
    """potrf + trtri through the int8 tensor-core path (left-looking, statically scaled slices; the default) and through
    the FP64 DMMA path (AGPC_OZAKI=0) on the same matrices: both within 1e-12 of LAPACK, and of each other."""
    import json
    import subprocess
    import sys
    code = r'''
import json, os, sys
import numpy as np, torch
sys.path.insert(0, %r)
from autogpc_b200.engine import Engine
e = Engine(0)
rng = np.random.default_rng(2)
out = {}
for n, batch in ((640, 2), (2304, 1)):
    G = rng.standard_normal((batch, n, n))
    S = G @ G.transpose(0, 2, 1) / n + np.eye(n) * (1.0 + 50.0 * rng.random((batch, n, 1)))
    S = 0.5 * (S + S.transpose(0, 2, 1))
    A = torch.from_numpy(S.copy()).cuda()
    Mi = torch.zeros_like(A)
    Ui = torch.zeros_like(A)
    torch.cuda.synchronize()
    e.potrf(A.data_ptr(), n, batch, Mi.data_ptr(), Ui.data_ptr(), 0)
    torch.cuda.synchronize()
    L = np.tril(A.cpu().numpy())
    Lref = np.linalg.cholesky(S)
    Minv = np.linalg.inv(Lref)
    out[str(n)] = [float(np.abs(L - Lref).max() / np.abs(Lref).max()), float(np.abs(np.tril(Mi.cpu().numpy()) - Minv).max()),
                   float(np.abs(np.triu(Ui.cpu().numpy()) - Minv.transpose(0, 2, 1)).max()), float(L.sum()), float(np.tril(Mi.cpu().numpy()).sum())]
print(json.dumps(out))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for oz in ("7", "0"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, AGPC_OZAKI=oz), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res[oz] = json.loads(r.stdout.strip().splitlines()[-1])
    for n in ("640", "2304"):
        for oz in ("7", "0"):
            assert res[oz][n][0] <= 1e-12 and res[oz][n][1] <= 1e-12 and res[oz][n][2] <= 1e-12, (oz, n, res[oz][n])
        assert abs(res["7"][n][3] - res["0"][n][3]) <= 1e-10 * abs(res["0"][n][3])
        assert abs(res["7"][n][4] - res["0"][n][4]) <= 1e-10 * abs(res["0"][n][4])


@pytest.mark.gpu
def test_device_optimizer_matches_host_driver(engine):
    """m.optimize() (gpckernel.py:246) with the L-BFGS state on the device (agpc_optimize_batch) against the
    host driver that steps SciPy's own L-BFGS-B (optim.lbfgsb_lockstep).  Same algorithm and stopping rules, different
    rounding, so the trajectories agree to optimiser tolerance, not bit for bit: every model must end at the same
    optimum (NLML to 1e-4 relative -- factr = 1e7 stops at 2e-9 relative decrease per iteration --, hyper-parameters to
    2 %), with a comparable number of evaluations, and leave consistent sites behind (re-scoring at theta* from the
    returned sites reproduces the returned NLML)."""
    from autogpc_b200 import engine as eng_mod
    from autogpc_b200 import gpy_compat as GPy
    from autogpc_b200 import optim

    X, y = synth(260, 3, 31)
    eng_mod.set_default_engine(engine)
    did = engine.dataset(X, y)
    rng = np.random.default_rng(12)

    def kernels():
        k = []
        k.append(GPy.kern.RBF(1, variance=1.3, lengthscale=0.7, active_dims=[0]))
        k.append(GPy.kern.Prod([GPy.kern.RBF(1, variance=0.8, lengthscale=1.5, active_dims=[0]),
                                GPy.kern.RBF(1, variance=1.1, lengthscale=0.9, active_dims=[1])]))
        k.append(GPy.kern.Add([GPy.kern.RBF(1, variance=0.6, lengthscale=2.0, active_dims=[2]),
                               GPy.kern.Bias(3, variance=0.5)]))
        per = GPy.kern.StdPeriodic(1, variance=1.0, period=2.5, lengthscale=1.2, active_dims=[1])
        per["period"].constrain_bounded(0.5, 6.0, warning=False)      # a Logistic-transformed parameter
        k.append(GPy.kern.Add([per, GPy.kern.RBF(1, variance=0.9, lengthscale=1.1, active_dims=[0])]))
        k.append(GPy.kern.Linear(1, variance=0.7, location=0.2, active_dims=[0]))    # a free (untransformed) parameter
        return k

    folds = [np.sort(rng.choice(260, size=200 + 10 * i, replace=False)).astype(np.int32) for i in range(5)]

    def models():
        return [GPy.models.GPClassification(kernel=k, engine=engine, _dataset=did, _rows=r, _infer=False)
                for k, r in zip(kernels(), folds)]

    host, dev = models(), models()
    optim.optimize_models(host, max_evals=200, device=False)
    optim.optimize_models(dev, max_evals=200, device=True)
    for i, (h, d) in enumerate(zip(host, dev)):
        assert d.optimization_info["device"] and d.status in (0, 1), (i, d.status)
        assert d._nlml == pytest.approx(h._nlml, rel=1e-4, abs=1e-4), (i, h._nlml, d._nlml, h.optimization_info, d.optimization_info)
        assert np.allclose(d.kern.param_array, h.kern.param_array, rtol=2e-2, atol=1e-3), (i, h.kern.param_array, d.kern.param_array)
        assert d.optimization_info["warnflag"] == h.optimization_info["warnflag"], (i, h.optimization_info, d.optimization_info)
        assert abs(d.optimization_info["evals"] - h.optimization_info["evals"]) <= max(6, 0.5 * h.optimization_info["evals"])
        # the sites that came back belong to theta*: a warm-started re-evaluation stays put
        again = engine.score_batch(did, [d._program], [d.kern.param_array], [d._rows], tau=[d._tau], nu=[d._nu], **d.ep)
        assert again["nlml"][0] == pytest.approx(d._nlml, rel=1e-7)
    engine.dataset_free(did)


@pytest.mark.gpu
def test_device_optimizer_keeps_search_selection(engine):
    """north_star for N1: routing gpcsearch's training through the on-device optimiser must not change which structure is
    selected at any depth (same twin-search set-up as test_search_selects_identical_structures_as_oracle)."""
    from autogpc_b200 import engine as eng_mod
    from autogpc_b200 import gpckernel
    from autogpc_b200.gpcdata import GPCData
    from autogpc_b200.gpcsearch import GPCSearch

    X, y = synth(150, 3, 8)

    def run(device):
        old = os.environ.get("AGPC_DEVICE_OPT")
        os.environ["AGPC_DEVICE_OPT"] = "1" if device else "0"
        try:
            eng_mod.set_default_engine(engine)
            gpckernel.set_rng(5)
            d = GPCData(X, y.reshape(-1, 1), seed=1)
            s = GPCSearch(d, base_kernels="SE,Lin,Per", max_depth=2, beam_width=1, criterion="bic", n_folds=3, max_evals=60)
            best, _ = s.search()
            return [k.kernel.structure() for k in best], [s.score(k) for k in best]
        finally:
            if old is None:
                os.environ.pop("AGPC_DEVICE_OPT", None)
            else:
                os.environ["AGPC_DEVICE_OPT"] = old

    names_h, scores_h = run(False)
    names_d, scores_d = run(True)
    assert names_h == names_d, (names_h, names_d)
    assert np.allclose(scores_h, scores_d, rtol=1e-3), (scores_h, scores_d)


@pytest.mark.gpu
@pytest.mark.parametrize("max_evals", [1, 3, 7])
def test_device_optimizer_stopping_rules_match_scipy(engine, max_evals):
    """SciPy's maxfun semantics (the run stops at the first NEW iterate after more than maxfun evaluations, so it may
    overshoot inside a line search) and its warnflag, reproduced by the device loop: same number of objective evaluations,
    same warnflag, same final objective as the SciPy-stepped host driver under a tight budget."""
    from autogpc_b200 import engine as eng_mod
    from autogpc_b200 import gpy_compat as GPy
    from autogpc_b200 import optim

    X, y = synth(180, 2, 44)
    eng_mod.set_default_engine(engine)
    did = engine.dataset(X, y)
    rows = [np.arange(150, dtype=np.int32), np.arange(20, 180, dtype=np.int32)]

    def models():
        ks = [GPy.kern.RBF(1, variance=2.0, lengthscale=0.3, active_dims=[0]),
              GPy.kern.Add([GPy.kern.RBF(1, variance=0.5, lengthscale=1.7, active_dims=[1]), GPy.kern.Bias(2, variance=1.5)])]
        return [GPy.models.GPClassification(kernel=k, engine=engine, _dataset=did, _rows=r, _infer=False) for k, r in zip(ks, rows)]

    host, dev = models(), models()
    optim.optimize_models(host, max_evals=max_evals, device=False)
    optim.optimize_models(dev, max_evals=max_evals, device=True)
    for h, d in zip(host, dev):
        assert d.optimization_info["evals"] == h.optimization_info["evals"], (h.optimization_info, d.optimization_info)
        assert d.optimization_info["warnflag"] == h.optimization_info["warnflag"], (h.optimization_info, d.optimization_info)
        assert d._nlml == pytest.approx(h._nlml, rel=1e-9)
        assert np.allclose(d.kern.param_array, h.kern.param_array, rtol=1e-7)
    engine.dataset_free(did)


@pytest.mark.gpu
def test_slicing_shortcuts_do_not_change_a_single_bit():
    """trtri's slice launches take their row scales from maxima that the producers of the data collect (T: the epilogue of
    the product that writes it; U11 / M22: running row maxima of U and M) instead of a pass of their own over the data.
    Those maxima are exact, so L^-1 and L^-T must come out bit-identical with the shortcuts switched off (AGPC_OZ_ROWMAX=0);
    the split of the given-scale slice launches over several CTAs (AGPC_OZ_SLICE_SPLIT=0) must not change anything either."""
    import hashlib
    import json
    import subprocess
    import sys
    code = r'''
import hashlib, json, os, sys
import numpy as np, torch
sys.path.insert(0, %r)
from autogpc_b200.engine import Engine
e = Engine(0)
rng = np.random.default_rng(7)
out = {}
for n, batch in ((1536, 2), (2432, 1)):
    G = rng.standard_normal((batch, n, n))
    S = G @ G.transpose(0, 2, 1) / n + np.eye(n) * (1.0 + 30.0 * rng.random((batch, n, 1)))
    S = 0.5 * (S + S.transpose(0, 2, 1))
    A = torch.from_numpy(S.copy()).cuda()
    Mi = torch.zeros_like(A)
    Ui = torch.zeros_like(A)
    torch.cuda.synchronize()
    e.potrf(A.data_ptr(), n, batch, Mi.data_ptr(), Ui.data_ptr(), 0)
    torch.cuda.synchronize()
    h = hashlib.sha256()
    for t in (A, Mi, Ui):
        h.update(np.tril(t.cpu().numpy()).tobytes() if t is not Ui else np.triu(t.cpu().numpy()).tobytes())
    Minv = np.linalg.inv(np.linalg.cholesky(S))
    out[str(n)] = [h.hexdigest(), float(np.abs(np.tril(Mi.cpu().numpy()) - Minv).max())]
print(json.dumps(out))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for tag, env in (("default", {}), ("two-pass", {"AGPC_OZ_ROWMAX": "0"}), ("no-split", {"AGPC_OZ_SLICE_SPLIT": "0"})):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res[tag] = json.loads(r.stdout.strip().splitlines()[-1])
    for n in ("1536", "2432"):
        assert res["default"][n][1] < 1e-12
        assert res["default"][n][0] == res["two-pass"][n][0] == res["no-split"][n][0], (n, res)
