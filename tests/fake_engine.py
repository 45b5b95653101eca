"""Oracle-backed stand-in for ``autogpc_b200.engine.Engine`` so the host logic (GPy shim, lock-step optimiser, search,
sharding) can be exercised on a CPU-only box.  Lives under tests/ -- the product never imports the oracle."""
import numpy as np

from oracle import gpc_oracle as orc


def _to_oracle(program):
    L = program.n_leaves
    return orc.KernelProgram([program.leaf_type[i] for i in range(L)], [program.leaf_dim[i] for i in range(L)],
                             [program.leaf_off[i] for i in range(L)], program.terms(), program.n_theta)


class OracleEngine:
    device = -1

    def __init__(self):
        self._ds = {}
        self._next = 0
        self.launches = 0
        self.calls = 0
        self.items = 0

    def dataset(self, X, y):
        self._ds[self._next] = (np.asarray(X, float), np.asarray(y, float).reshape(-1))
        self._next += 1
        return self._next - 1

    def dataset_free(self, did):
        self._ds.pop(did, None)

    def score_batch(self, dataset, programs, thetas, rows, epsilon=1e-6, damping=1.0, max_sweeps=100, tau=None, nu=None,
                    want_grad=True, frozen_sites=False, inference="ep"):
        X, y = self._ds[dataset]
        B = len(programs)
        self.calls += 1
        self.items += B
        stride = max(p.n_theta for p in programs)
        out = dict(nlml=np.zeros(B), grad=np.zeros((B, stride)), bic=np.zeros(B), nlml_gauss=np.zeros(B),
                   sweeps=np.zeros(B, dtype=np.int32), status=np.zeros(B, dtype=np.int32), tau=[], nu=[])
        for b in range(B):
            if inference == "laplace":
                r = orc.score_laplace(_to_oracle(programs[b]), thetas[b], X[rows[b]], y[rows[b]], p_eff=programs[b].p_eff,
                                      tol=epsilon, max_iter=max_sweeps)
            else:
                r = orc.score(_to_oracle(programs[b]), thetas[b], X[rows[b]], y[rows[b]], eps=epsilon,
                              max_sweeps=max_sweeps, damping=damping, p_eff=programs[b].p_eff,
                              tau0=None if tau is None else tau[b], v0=None if nu is None else nu[b],
                              frozen_sites=frozen_sites)
            out["status"][b], out["sweeps"][b] = r["status"], r["sweeps"]
            out["nlml"][b], out["bic"][b], out["nlml_gauss"][b] = r["nlml"], r["bic"], r["nlml_gauss"]
            out["grad"][b, :programs[b].n_theta] = r["grad"]
            n = len(rows[b])
            out["tau"].append(r["tau"] if r["tau"] is not None else np.zeros(n))
            out["nu"].append(r["v"] if r["v"] is not None else np.zeros(n))
        return out

    def predict_batch(self, dataset, programs, thetas, rows, tau, nu, test_rows, want_latent=False):
        X, _ = self._ds[dataset]
        ps, mus, vs = [], [], []
        for b in range(len(programs)):
            p, mu, var = orc.predict(_to_oracle(programs[b]), thetas[b], X[rows[b]], X[test_rows[b]], tau[b], nu[b])
            ps.append(p); mus.append(mu); vs.append(var)
        st = np.zeros(len(programs), dtype=np.int32)
        return (ps, mus, vs, st) if want_latent else (ps, st)

    def predict_points(self, dataset, program, theta, rows, tau, nu, Xstar, want_gradients=False):
        X, _ = self._ds[dataset]
        prog = _to_oracle(program)
        p, mu, var = orc.predict(prog, theta, X[rows], np.asarray(Xstar, float), tau, nu)
        if not want_gradients:
            return p, mu, var
        h = 1e-6
        g = np.zeros_like(np.asarray(Xstar, float))
        for d in range(g.shape[1]):
            e = np.zeros(g.shape[1]); e[d] = h
            g[:, d] = (orc.predict(prog, theta, X[rows], Xstar + e, tau, nu)[1] -
                       orc.predict(prog, theta, X[rows], Xstar - e, tau, nu)[1]) / (2 * h)
        return p, mu, var, g
