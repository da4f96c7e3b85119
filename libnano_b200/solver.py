"""Device-resident L-BFGS and CGD (nb200_lbfgs_minimize, nb200_cgd_minimize): libnano's solver_lbfgs_t with its default line search
(src/solver/lbfgs.cpp:25-124, src/lsearchk/cgdescent.cpp, src/lsearch0/quadratic.cpp) keeping x, g, the (s, y)
history, the two-loop recursion and the trial points in HBM. libnano's own solvers still work unchanged through
`function(x, gx)`; this entry is for callers that want the PCIe round trip per evaluation gone."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .function import critical

STATUS = {0: "max_iters", 1: "gradient_test", 4: "failed"}


CGD_BETAS = {"hs": 0, "fr": 1, "pr": 2, "cd": 3, "ls": 4, "dy": 5, "n": 6, "dycd": 7, "dyhs": 8, "frpr": 9}


def _device_minimize(function, x0, invoke):
    """Shared plumbing of the device-resident line-search solvers: resolves the rank-local objective and the
    cross-rank exchange (fused peers / all-reduce callback / none), calls `invoke(handle, x_ptr, n, n_global,
    callback, fx, gtest, stats, status)` and packs the result like solver_state_t."""
    from .sharded import ShardedFunction

    x = np.ascontiguousarray(x0, dtype=np.float64).copy()
    critical(x.ndim == 1 and x.size == function.size(), "solver: incompatible initial point (", x.size,
             " dimensions), expecting ", function.size(), " dimensions!")

    callback, n_global = None, 0
    local = function
    if isinstance(function, ShardedFunction):
        shard = function._shard
        local = shard.local
        if function.world_size() > 1 and getattr(shard, "fused", False):
            n_global = function._n_global  # the all-reduce happens inside the kernel (peer memory): no callback
        elif function.world_size() > 1:
            import torch.distributed as dist

            n_global = function._n_global

            def _allreduce(_user, ptr, length, _stream):
                try:
                    dist.all_reduce(shard._view_for(ptr, length), group=function._group)
                    return 0
                except Exception:  # pragma: no cover - surfaced as a solver failure
                    return 1

            callback = _lib.ALLREDUCE_FN(_allreduce)

    function.clear_statistics()  # solver_t::minimize (src/solver.cpp:107)
    fx, gtest = ctypes.c_double(), ctypes.c_double()
    stats = (ctypes.c_int64 * 3)()
    status = ctypes.c_int()
    _lib.check(invoke(local.handle, x.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), x.size, int(n_global),
                      ctypes.cast(callback, ctypes.c_void_p) if callback else None, ctypes.byref(fx),
                      ctypes.byref(gtest), stats, ctypes.byref(status)))
    function._fcalls += int(stats[0])
    function._gcalls += int(stats[1])
    return {"x": x, "fx": fx.value, "gradient_test": gtest.value, "status": STATUS.get(status.value, "failed"),
            "fcalls": int(stats[0]), "gcalls": int(stats[1]), "iterations": int(stats[2])}


def lbfgs(function, x0, epsilon: float = 1e-8, max_evals: int = 1000, history: int = 50):
    """Minimise a GPU objective (LinearFunction, SyntheticFunction or sharded.ShardedFunction) from x0.

    Defaults are the reference's (src/solver.cpp:45-51: epsilon 1e-8, max_evals 1000; lbfgs.cpp:12: history 50).
    Returns a dict like solver_state_t: x, fx, gradient_test, status, fcalls, gcalls, iterations."""
    return _device_minimize(function, x0, lambda handle, x, n, n_global, callback, fx, gtest, stats, status:
                            _lib.lib().nb200_lbfgs_minimize(handle, x, n, float(epsilon), int(max_evals), int(history),
                                                            n_global, callback, None, fx, gtest, stats, status))


def cgd(function, x0, beta: str = "pr", epsilon: float = 1e-8, max_evals: int = 1000, orthotest: float = 0.1,
        eta: float = 0.01):
    """Device-resident non-linear conjugate gradient descent, libnano's solver ids cgd-<beta> (src/solver/cgd.cpp;
    "pr" is the one registered as plain "cgd"): same skeleton and line search as lbfgs() with tolerance (1e-4, 1e-1);
    the beta formula, the direction update and the restart test run as one kernel on the device."""
    critical(beta in CGD_BETAS, "solver: unknown cgd variant '", beta, "', expecting one of ", sorted(CGD_BETAS))
    return _device_minimize(function, x0, lambda handle, x, n, n_global, callback, fx, gtest, stats, status:
                            _lib.lib().nb200_cgd_minimize(handle, x, n, CGD_BETAS[beta], float(epsilon),
                                                          int(max_evals), float(orthotest), float(eta), n_global,
                                                          callback, None, fx, gtest, stats, status))


OSGA_STATUS = {0: "max_iters", 1: "gradient_test", 2: "specific_test", 3: "value_test", 4: "failed"}


def osga(function, x0, epsilon: float = 1e-8, max_evals: int = 1000, patience: int = 100, lambda_: float = 0.9,
         alpha_max: float = 0.7, kappas=(0.1, 1.1)):
    """OSGA (solver_osga_t::do_minimize, src/solver/osga.cpp:59-147), the non-smooth driver BASELINE config 5 names.

    A GPU objective (LinearFunction, SyntheticFunction, a device group, or a ShardedFunction on one rank / with fused
    peers) runs nb200_osga_minimize: every vector of the method stays in HBM. Anything else that is callable as
    `function(x, gx)` — a ShardedFunction over the NCCL two-phase path, a CPU double in the tests — takes osga_host, the
    same algorithm with numpy vectors."""
    from .function import _DeviceFunction
    from .sharded import ShardedFunction

    local, n_global = function, 0
    if isinstance(function, ShardedFunction):
        shard = function._shard
        local = getattr(shard, "local", None)
        if function.world_size() > 1:
            n_global = function._n_global
            if not getattr(shard, "fused", False):
                local = None
    if not isinstance(local, _DeviceFunction):
        return osga_host(function, x0, epsilon, max_evals, patience, lambda_, alpha_max, kappas)

    x = np.ascontiguousarray(x0, dtype=np.float64).copy()
    critical(x.ndim == 1 and x.size == function.size(), "solver: incompatible initial point (", x.size,
             " dimensions), expecting ", function.size(), " dimensions!")
    function.clear_statistics()
    fx, eta = ctypes.c_double(), ctypes.c_double()
    stats = (ctypes.c_int64 * 3)()
    status = ctypes.c_int()
    _lib.check(_lib.lib().nb200_osga_minimize(
        local.handle, x.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), x.size, float(epsilon), int(max_evals),
        int(patience), float(lambda_), float(alpha_max), float(kappas[0]), float(kappas[1]),
        float(function.strong_convexity()), int(n_global), ctypes.byref(fx), ctypes.byref(eta), stats,
        ctypes.byref(status)))
    function._fcalls += int(stats[0])
    function._gcalls += int(stats[1])
    return {"x": x, "fx": fx.value, "eta": eta.value, "status": OSGA_STATUS.get(status.value, "failed"),
            "fcalls": int(stats[0]), "gcalls": int(stats[1]), "iterations": int(stats[2])}


def osga_host(function, x0, epsilon: float = 1e-8, max_evals: int = 1000, patience: int = 100, lambda_: float = 0.9,
              alpha_max: float = 0.7, kappas=(0.1, 1.1)):
    """Host-driven OSGA (solver_osga_t::do_minimize, src/solver/osga.cpp:59-147) over `function(x, gx)`.

    The non-smooth driver BASELINE config 5 names (hinge + lasso / elastic-net). Per iteration: one f(x, g) and one
    value-only f(x') through the host-buffer call — exactly how libnano's OSGA drives a function_t — plus O(n) host
    vector algebra (n = D + 1, negligible next to the pass over X). Defaults: src/solver.cpp:45-51, osga.cpp:49-51."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = function.size()
    critical(x0.ndim == 1 and x0.size == n, "solver: incompatible initial point (", x0.size, " dimensions), expecting ",
             n, " dimensions!")
    kappa_prime, kappa = kappas
    miu = function.strong_convexity() / 2.0
    epsilon0 = 1e-15

    # proxy_t (osga.cpp:8-39): Q(z) = Q0 + |z - z0|^2 / 2
    z0 = x0.copy()
    Q0 = 0.5 * np.sqrt(np.linalg.norm(z0) + np.finfo(float).eps)

    def Q(z):
        return Q0 + 0.5 * float(np.dot(z - z0, z - z0))

    def E(gamma, h):
        beta = gamma + float(np.dot(h, z0))
        hh = float(np.dot(h, h))
        root = np.sqrt(beta * beta + 2.0 * Q0 * hh)
        return (-beta + root) / (2.0 * Q0) if beta <= 0.0 else hh / (beta + root)

    def U(gamma, h):
        return z0 - h / E(gamma, h)

    function.clear_statistics()
    g = np.empty(n)
    best_x, best_f = x0.copy(), float(function(x0, g))      # solver_state_t keeps track of the best state
    best_g = g.copy()
    history_df, history_dx = [], []
    track_x, track_f = best_x.copy(), best_f

    def track_update():                                       # solver_track_t::update (track.cpp:11-22)
        nonlocal track_x, track_f
        history_df.append(track_f - best_f)
        history_dx.append(float(np.max(np.abs(track_x - best_x))) / max(1.0, float(np.max(np.abs(track_x)))))
        track_x, track_f = best_x.copy(), best_f

    def value_test():                                         # track.cpp:24-59
        size = len(history_df)
        last, dd = size, np.finfo(float).max
        for it in range(size, 0, -1):
            if history_df[it - 1] > 0.0:
                dd, last = max(history_dx[it - 1], history_df[it - 1]), it - 1
                break
        if last == size:
            return 0.0 if size >= patience else dd
        return dd if last + patience >= size else 0.0

    h = best_g - miu * (best_x - z0)
    gamma = best_f - miu * Q(best_x) - float(np.dot(h, best_x))
    u = U(gamma - best_f, h)
    eta = E(gamma - best_f, h) - miu
    alpha, xb, fb = alpha_max, best_x.copy(), best_f
    status, iterations = "max_iters", 0

    while function.fcalls() + function.gcalls() < max_evals:
        if float(np.max(np.abs(best_g))) < epsilon0:
            status = "gradient_test"
            break
        x = xb + alpha * (u - xb)
        f = float(function(x, g))
        gq = g - miu * (x - z0)
        h_hat = h + alpha * (gq - h)
        gamma_hat = gamma + alpha * (f - miu * Q(x) - float(np.dot(gq, x)) - gamma)
        xb_prime, fb_prime = (x, f) if f < fb else (xb, fb)
        u_prime = U(gamma_hat - fb_prime, h_hat)
        x_prime = xb + alpha * (u_prime - xb)
        f_prime = float(function(x_prime))
        xb_hat, fb_hat = (x_prime, f_prime) if f_prime < fb_prime else (xb_prime, fb_prime)
        if np.isfinite(fb_hat) and best_f > fb_hat:          # state.update_if_better (gradient kept)
            best_x, best_f = xb_hat.copy(), fb_hat
        u_hat = U(gamma_hat - fb_hat, h_hat)
        eta_hat = E(gamma_hat - fb_hat, h_hat) - miu
        iterations += 1

        valid = np.isfinite(best_f) and bool(np.all(np.isfinite(best_x)))
        track_update()
        if not valid:
            status = "failed"
            break
        if eta_hat < epsilon:
            status, eta = "specific_test", eta_hat
            break
        track_update()                                        # done_value_test updates the history once more
        if value_test() < epsilon:
            status, eta = "value_test", eta_hat
            break

        R = (eta - eta_hat) / (lambda_ * alpha * eta)
        alpha = alpha * np.exp(-kappa) if R < 1.0 else min(alpha * np.exp(kappa_prime * (R - 1.0)), alpha_max)
        if eta_hat < eta:
            h, u, eta, gamma = h_hat, u_hat, eta_hat, gamma_hat
        xb, fb = np.array(xb_hat, copy=True), fb_hat

    return {"x": best_x, "fx": best_f, "eta": float(eta), "status": status, "fcalls": function.fcalls(),
            "gcalls": function.gcalls(), "iterations": iterations}


def _simplex_qp(K, h):
    """argmin_a  a.K.a / 2 - h.a  over the unit simplex (K symmetric positive semi-definite, m x m, m <= ~100): the
    DUAL of the bundle's quadratic program (bundle.cpp:109-168 builds the PRIMAL in n + 1 variables and hands it to
    libnano's interior-point solver), one variable per cutting plane. Host algebra of the library
    (nb200_simplex_qp, csrc/host_qp.cu: Mehrotra predictor-corrector, exact zeros for the inactive planes)."""
    K = np.ascontiguousarray(K, dtype=np.float64)
    h = np.ascontiguousarray(h, dtype=np.float64)
    alphas = np.empty(h.size, dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    status = _lib.lib().nb200_simplex_qp(K.ctypes.data_as(dp), h.ctypes.data_as(dp), h.size, alphas.ctypes.data_as(dp))
    critical(status == 0, "solver: the bundle's quadratic program is not convex (indefinite metric)")
    return alphas


def _rank_one(A, v, alpha):
    """A += alpha v v' in place for a symmetric C-contiguous A (no n x n temporary)."""
    from scipy.linalg.blas import dger

    dger(alpha, v, v, a=A.T, overwrite_a=True)


class _Bundle:
    """The bundle of cutting planes of libnano's proximal bundle solvers (bundle_t, src/solver/bundle/bundle.cpp):
    rows (g_j, h_j) with h_j = f_j + g_j.(x - x_j) relative to the stability centre x; a full bundle first drops the
    planes with zero multipliers, then the two smallest ones in favour of the aggregate plane."""

    def __init__(self, x, gx, fx, max_size, apply_inverse_metric=None):
        n = x.size
        self.G = np.zeros((max_size + 1, n))
        self.H = np.zeros(max_size + 1)
        self.size = 0
        self.x, self.gx, self.fx = x.copy(), gx.copy(), float(fx)
        self.alphas = np.zeros(0)
        self.solution_x = x.copy()
        self.solution_g = np.zeros(n)
        # K = G M^-1 G' of the current planes under the current metric, kept up to date INCREMENTALLY: a new plane costs
        # one M^-1 g (O(n^2)) and one G z (O(m n)), a rank-one change of the metric O(m n + m^2) — instead of the
        # O(m n^2) product per trial point that rebuilding it took (3 ms at n = 513 with a full bundle, against a 9.7 ms
        # pass over X at config 5). K only feeds the quadratic program; the proximal point itself is formed from the
        # metric directly (x - t M^-1 G'a), so rounding in K never separates y from the aggregate gradient it implies.
        self._inverse_metric = apply_inverse_metric if apply_inverse_metric is not None else (lambda v: v)
        self.K = np.zeros((max_size + 1, max_size + 1))
        self._updates = 0
        self._append(x, gx, fx, True)

    def capacity(self):
        return self.H.size

    def tol(self, epsilon):                                   # etol == gtol (bundle.cpp:66-74)
        return epsilon * (1.0 + abs(self.fx))

    def fhat(self, z):                                        # bundle.cpp:170-180
        return float(np.max(self.H[:self.size] + self.G[:self.size] @ (z - self.x)))

    def solve(self, apply_inverse_metric, t):
        """Proximal point y = argmin fhat(z) + (z - x).M.(z - x) / (2 t)  (bundle.cpp:95-168 without the level
        constraint, which RQB never sets). Through the dual:  y = x - t M^-1 G'a  with a = _simplex_qp(t G M^-1 G', h)."""
        m = self.size
        G, H = self.G[:m], self.H[:m]
        if self._updates >= 256:                              # bound the drift of the incremental updates
            self.refresh(apply_inverse_metric)
        K = self.K[:m, :m]
        self.alphas = _simplex_qp(t * (0.5 * (K + K.T)), H)
        self.solution_g = G.T @ self.alphas
        self.solution_x = self.x - t * apply_inverse_metric(self.solution_g)
        return self.solution_x

    def refresh(self, apply_inverse_metric=None):
        """Rebuild K from the planes and the metric."""
        if apply_inverse_metric is not None:
            self._inverse_metric = apply_inverse_metric
        m = self.size
        self.K[:m, :m] = self.G[:m] @ self._inverse_metric(self.G[:m].T)
        self._updates = 0

    def metric_rank_one(self, e, ev):
        """The inverse metric became W + e e' / ev (SR1 through Sherman-Morrison): K += (G e)(G e)' / ev."""
        m = self.size
        Ge = self.G[:m] @ e
        self.K[:m, :m] += np.multiply.outer(Ge / ev, Ge)
        self._updates += 1

    def metric_rescale(self, factor):
        """The inverse metric was multiplied by `factor` (scaled-identity metric)."""
        m = self.size
        self.K[:m, :m] *= factor
        self._updates += 1

    def _gram_append(self, row):
        """Plane `row` of G is new: its row / column of K."""
        k = self.G[:row + 1] @ self._inverse_metric(self.G[row])
        self.K[row, :row + 1] = k
        self.K[:row + 1, row] = k
        self._updates += 1

    def _remove(self, drop):
        keep = np.flatnonzero(~drop)
        self.G[:keep.size] = self.G[keep]
        self.H[:keep.size] = self.H[keep]
        self.K[:keep.size, :keep.size] = self.K[np.ix_(keep, keep)]
        self.alphas = self.alphas[keep]
        self.size = keep.size

    def _append(self, y, gy, fy, serious_step):               # bundle.cpp:247-275
        if self.size + 1 == self.capacity() and self.alphas.size == self.size:
            # delete_inactive(epsilon0) (bundle.cpp:182-188) - only once the bundle is full: the dual solve returns
            # exact zeros for every inactive plane, where the reference's interior-point multipliers stay above
            # 1e-15 for all but the far-away planes, so dropping zeros after every step would discard most of the
            # model each time
            self._remove(self.alphas < 1e-15)
            if self.size + 1 == self.capacity():              # delete_smallest(2) + aggregate (bundle.cpp:190-245)
                # the aggregate plane: (sum_j a_j g_j, fhat(y)); the reference writes (x - y) / tau for its gradient
                # (bundle.cpp:232), which equals this only under the identity metric - under RQB's variable metric
                # that plane is not a minorant of f, so the true aggregate is used here
                agg_h = self.fhat(self.solution_x) + float(self.solution_g @ (self.x - self.solution_x))
                agg_g = self.solution_g.copy()
                order = np.argsort(self.alphas, kind="stable")[:2]
                drop = np.zeros(self.size, dtype=bool)
                drop[order] = True
                self._remove(drop)
                self.G[self.size], self.H[self.size] = agg_g, agg_h
                self._gram_append(self.size)
                self.alphas = np.append(self.alphas, 1.0)
                self.size += 1
        if serious_step:
            self.H[:self.size] += self.G[:self.size] @ (y - self.x)
            self.x, self.gx, self.fx = y.copy(), gy.copy(), float(fy)
        self.G[self.size] = gy
        self.H[self.size] = fy + float(gy @ (self.x - y))
        self._gram_append(self.size)
        self.size += 1
        self.alphas = np.zeros(0)                             # multipliers are recomputed by the next solve

    def moveto(self, y, gy, fy):
        self._append(y, gy, fy, True)

    def append(self, y, gy, fy):
        self._append(y, gy, fy, False)


def rqb(function, x0, epsilon: float = 1e-8, max_evals: int = 1000, max_size: int = 100, prox_type: str = "sr1",
        tau_min: float = 1e-5, m1m2=(0.5, 0.9), m3: float = 1.0, m4: float = 1.0, interpol: float = 0.3,
        extrapol: float = 5.0, reference_init: bool = False):
    """Host-driven RQB, the reversal quasi-Newton proximal bundle solver BASELINE config 5 names next to OSGA
    (solver_rqb_t::do_minimize src/solver/rqb.cpp:23-71; curve search src/solver/bundle/csearch.cpp:40-153; metric
    src/solver/bundle/quasi.cpp; bundle src/solver/bundle/bundle.cpp) over `function(x, gx)`.

    One f(y, gy) through the host-buffer call per trial point, exactly how libnano's RQB drives a function_t; the
    bundle algebra stays on the host (at most max_size + 1 planes of n doubles). prox_type "sr1" keeps an n x n
    metric (n <= 8192); "miu" the scaled identity at any n. Defaults: src/solver.cpp:45-51, csearch.cpp:155-162,
    quasi.cpp:116-120, bundle.cpp:277-280.

    Same algorithm, different numerics behind the proximal point (the reference keeps its own RQB tests disabled,
    test/test_solver_bundle.cpp:11, so there is no reference trajectory to pin; tests/test_host_solvers.py checks
    the optimum against a linear program instead):
      * the quadratic program is solved in its dual - one variable per plane (_simplex_qp) - instead of the
        reference's interior-point solve in n + 1 variables, and the inverse metric is carried along by the matching
        rank-one formula (W += e e' / e.v) instead of being factorised per solve;
      * inactive planes are dropped only when the bundle is full, and the aggregate plane is the true aggregate
        (_Bundle._append);
      * the metric starts at I / tau0 (reference_init=True restores quasi.cpp:33's tau0 I); the SR1 update also
        requires positive curvature e.v, the scaled-identity update keeps positive candidates only and moves at most
        a decade per step (quasi_update);
      * the convexity tests of the curve search get a rounding-level slack (csearch.cpp:99), and a curve search whose
        bracket [tL, tR] has collapsed to rounding ends with the step its last trial point earned instead of bisecting
        until the evaluation budget is spent."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = function.size()
    critical(x0.ndim == 1 and x0.size == n, "solver: incompatible initial point (", x0.size, " dimensions), expecting ",
             n, " dimensions!")
    critical(prox_type in ("sr1", "miu"), "solver: unknown rqb metric '", prox_type, "', expecting sr1 or miu")
    critical(prox_type == "miu" or n <= 8192, "solver: rqb with the sr1 metric keeps two ", n, " x ", n,
             " matrices on the host, use prox_type='miu'")
    m1, m2 = m1m2
    eps0 = 1e-15

    function.clear_statistics()
    g = np.empty(n)
    x, fx = x0.copy(), float(function(x0, g))
    gx = g.copy()

    def evals():
        return function.fcalls() + function.gcalls()

    def valid():
        return np.isfinite(fx) and bool(np.all(np.isfinite(x))) and bool(np.all(np.isfinite(gx)))

    result_status, iterations = "max_iters", 0
    if not valid():
        return {"x": x, "fx": fx, "status": "failed", "fcalls": function.fcalls(), "gcalls": function.gcalls(),
                "iterations": 0}

    # quasi_t (quasi.cpp:9-41): M0 = tau0 I, history of (x, g, G) at the last two serious iterates
    gg = float(gx @ gx)
    tau0 = max(1.0, abs(fx)) / (5.0 * gg) if gg > 0.0 else np.inf
    tau0 = max(tau0, tau_min) if np.isfinite(tau0) else tau_min
    # The proximal term is (z - x).M.(z - x) / (2 t), so M is a curvature (1 / step size). quasi.cpp:33 starts it at
    # tau0 I, the step size of (6) itself; with that start the metric is ~|g|^-4 too stiff in every direction the
    # rank-one updates have not touched yet and the solver crawls (the reference keeps its own RQB tests disabled,
    # test/test_solver_bundle.cpp:11, test/test_linear_elastic_net.cpp:11). M0 = I / tau0 unless reference_init.
    miu = tau0 if reference_init else 1.0 / tau0
    M = np.eye(n) * miu if prox_type == "sr1" else None
    W = np.eye(n) / miu if prox_type == "sr1" else None
    xn1, gn1, Gn1 = x.copy(), gx.copy(), gx.copy()
    xn, gn, Gn = xn1, gn1, Gn1

    def apply_metric(v):
        return M @ v if M is not None else miu * v

    def apply_inverse_metric(v):
        return W @ v if W is not None else v / miu

    def quasi_update(y, gy, G, is_descent_step):              # quasi.cpp:43-114
        nonlocal xn, gn, Gn, xn1, gn1, Gn1, miu, M, W
        xn, gn, Gn = xn1, gn1, Gn1
        xn1, gn1, Gn1 = y.copy(), gy.copy(), G.copy()
        if not is_descent_step:
            return
        e = xn1 - xn
        if prox_type == "miu":
            v1 = Gn1 - Gn
            candidates = []
            with np.errstate(all="ignore"):
                for v in (v1, Gn1 - gn, gn1 - Gn, gn1 - gn):
                    value = 1.0 / (np.float64(v @ e) / np.float64(v1 @ v) + 1.0 / miu)
                    # the reference keeps any finite candidate (quasi.cpp:80-83); a negative one makes the metric
                    # indefinite and the quadratic program unbounded, so only positive ones are kept here
                    candidates.append(float(value) if np.isfinite(value) and value > 0.0 else np.finfo(float).max)
            if min(candidates) != np.finfo(float).max:
                previous = miu
                miu = min(max(min(candidates), 0.1 * miu), 10.0 * miu)    # safeguard: at most a decade per serious step
                bundle.metric_rescale(previous / miu)
            return
        v = gn1 - gn
        Me = M @ e
        denom = float(e @ (Me + v))
        ev = float(e @ v)
        # quasi.cpp:100-107 applies the update when |e.(Me + v)| is not tiny; M stays positive definite only if the
        # curvature e.v is positive as well (it is >= 0 for a convex f, up to rounding), so that is required too -
        # the safeguard the reference leaves as a TODO (quasi.cpp:97-98)
        # ... and, f being piecewise linear for the losses this solver is for (hinge, mae), e.v is often ~0 inside
        # one linear piece: the update would flatten M along e and blow up its inverse, so the curvature left along e
        # must stay above 1e-6 of the initial one
        eMe = float(e @ Me)
        curvature_left = eMe * ev / ((eMe + ev) * float(e @ e)) if ev > 0.0 and eMe > 0.0 else 0.0
        if (abs(denom) >= 1e-8 * float(np.linalg.norm(e)) * float(np.linalg.norm(Me + v))
                and curvature_left >= 1e-6 * miu):
            # in-place symmetric rank-one updates (BLAS dger on the transposed view: both matrices are symmetric)
            _rank_one(M, Me, -1.0 / denom)
            _rank_one(W, e, 1.0 / ev)                         # Sherman-Morrison: (M - Me Me'/c)^-1 = W + e e' / (c - e.Me)
            bundle.metric_rank_one(e, ev)

    bundle = _Bundle(x, gx, fx, max_size, apply_inverse_metric)
    y, gy = x.copy(), np.empty(n)

    while evals() < max_evals:
        # csearch_t::search (csearch.cpp:40-153)
        t, tL, tR = 1.0, 0.0, np.inf
        status = "budget"
        while evals() < max_evals:
            cx, cfx = bundle.x, bundle.fx
            try:
                y = bundle.solve(apply_inverse_metric, t)
            except RuntimeError:                              # indefinite metric: the reference's QP solver fails too
                status = "failed"
                break
            fy = float(function(y, gy))
            fyhat = bundle.fhat(y)
            # gyhat = M (x - y) / t (csearch.cpp:73) with y = x - t M^-1 G'a is the aggregate gradient G'a itself: taken
            # from the bundle instead of another n x n product (M and its inverse are kept consistent by the rank-one
            # updates, so the two differ by rounding only)
            gyhat = bundle.solution_g
            step = y - cx
            delta = cfx - fyhat + 0.5 * float(gyhat @ step)
            error = cfx - fy + float(gy @ step)
            epsil = cfx - fyhat + float(gyhat @ step)
            gnorm = float(np.linalg.norm(gyhat))
            tol = bundle.tol(epsilon)
            econv, gconv = epsil <= tol, gnorm <= tol
            if econv and gconv:                               # stopping criterion (35)
                status = "converged"
                break
            # delta and error are non-negative for a convex f in exact arithmetic; f itself is only pinned to 1e-12
            # relative (the parity bar) and the proximal point to the accuracy of the quadratic program, so the
            # reference's `< 0` tests (csearch.cpp:99) get that much slack (the stopping tolerance at most: a deficit
            # below it is no evidence of a non-convex function)
            slack = max(1e-12 * (1.0 + abs(cfx)), tol)
            if not np.isfinite(fy) or delta < -slack or error < -slack:
                status = "failed"
                break
            if fy < cfx - m1 * delta:                         # descent test (31)
                tL = t
                if float(gy @ step) >= -m2 * delta:           # test (34)
                    status = "descent_step"
                    break
                if not np.isfinite(tR) and (gconv or float(gyhat @ step) >= -m4 * epsil):   # cutting-plane test (36)
                    status = "cutting_plane_step"
                    break
            else:
                tR = t
                if tL < eps0 and error <= m3 * delta:         # null-step test (33)
                    status = "null_step"
                    break
            if np.isfinite(tR) and tR - tL <= 1e-10 * tR:
                # the bracket has collapsed onto one step size: f is flat to rounding between tL and tR (the objective is
                # only evaluated to ~1e-16 relative), no further trial can change the outcome of tests (31)-(36). The
                # last trial point decides: it either passed the descent test (t == tL) or it did not (a null step).
                status = "descent_step" if t == tL else "null_step"
                break
            t = (1.0 - interpol) * tL + interpol * tR if np.isfinite(tR) else t * extrapol

        iterations += 1
        if status == "budget":
            break
        if status == "failed" or not valid():                 # solver_t::done_specific_test (src/solver.cpp:175-182)
            result_status = "failed"
            break
        if status == "converged":
            result_status = "specific_test"
            break
        if status in ("descent_step", "cutting_plane_step"):
            quasi_update(y, gy, gyhat, status == "descent_step")
            bundle.moveto(y, gy, fy)
            x, gx, fx = y.copy(), gy.copy(), fy
        else:
            bundle.append(y, gy, fy)

    return {"x": x, "fx": fx, "status": result_status, "fcalls": function.fcalls(), "gcalls": function.gcalls(),
            "iterations": iterations}


def lbfgs_batch(function, x0s, l1s, l2s, epsilon: float = 1e-8, max_rounds: int = 1000, history: int = 20,
                c1: float = 1e-4, c2: float = 0.9):
    """K L-BFGS solves that share one pass over X per evaluation round (the tuner's trials on one fold,
    src/machine/tune.cpp:25-42, differ only in their regularisation values).

    Every round evaluates the current trial point of all the unfinished models with ONE `function.vgrad_batch`
    (nb200_vgrad_batch: the K weight vectors become the rows of one many-target problem), then each model advances its
    own two-loop recursion (src/solver/lbfgs.cpp:49-84) and its own line search on the host: bracketing / bisection to
    the weak Wolfe conditions (c1, c2), or CG-DESCENT's approximate ones at rounding level, from a unit step (1 / |g|
    for the first iteration). Smooth objectives only
    (l1 = 0). Stops per model on libnano's gradient test |g|_inf / max(1, |f|) < epsilon (src/solver/state.cpp:138-141).
    Returns a list of dicts like lbfgs(): x, fx, gradient_test, status, fcalls, iterations; plus "rounds" = the number
    of passes over X the whole batch took."""
    xs = np.array(x0s, dtype=np.float64, ndmin=2)
    K, n = xs.shape
    critical(n == function.size(), "solver: incompatible initial points (", n, " dimensions), expecting ",
             function.size(), " dimensions!")
    l1s = np.broadcast_to(np.asarray(l1s, dtype=np.float64), (K,)).copy()
    l2s = np.broadcast_to(np.asarray(l2s, dtype=np.float64), (K,)).copy()
    critical(bool(np.all(l1s == 0.0)), "solver: lbfgs_batch is for smooth objectives (l1 = 0)")

    def gradient_test(f, g):
        return float(np.max(np.abs(g))) / max(1.0, abs(f))

    fs, gs = function.vgrad_batch(xs, l1s, l2s)
    rounds = 1
    models = []
    for k in range(K):
        model = {"x": xs[k].copy(), "f": float(fs[k]), "g": gs[k].copy(), "S": [], "Y": [], "status": None,
                 "fcalls": 1, "iterations": 0}
        if not (np.isfinite(model["f"]) and np.all(np.isfinite(model["g"]))):
            model["status"] = "failed"
        elif gradient_test(model["f"], model["g"]) < epsilon:
            model["status"] = "gradient_test"
        models.append(model)

    def new_direction(model):
        # two-loop recursion over the (s, y) pairs, newest last
        q = model["g"].copy()
        alphas = []
        for s, y in zip(reversed(model["S"]), reversed(model["Y"])):
            alpha = float(s @ q) / float(s @ y)
            alphas.append(alpha)
            q -= alpha * y
        if model["S"]:
            s, y = model["S"][-1], model["Y"][-1]
            q *= float(s @ y) / float(y @ y)
        for (s, y), alpha in zip(zip(model["S"], model["Y"]), reversed(alphas)):
            beta = float(y @ q) / float(s @ y)
            q += (alpha - beta) * s
        d = -q
        dg = float(d @ model["g"])
        if not dg < 0.0:                                      # not a descent direction: restart from steepest descent
            model["S"], model["Y"] = [], []
            d = -model["g"]
            dg = float(d @ model["g"])
        model["d"], model["dg"] = d, dg
        model["t"] = 1.0 if model["S"] else 1.0 / max(float(np.linalg.norm(model["g"])), 1e-300)
        model["lo"], model["hi"] = 0.0, np.inf

    for model in models:
        if model["status"] is None:
            new_direction(model)

    while rounds < max_rounds:
        active = [k for k, model in enumerate(models) if model["status"] is None]
        if not active:
            break
        trial = np.stack([models[k]["x"] + models[k]["t"] * models[k]["d"] for k in active])
        fs, gs = function.vgrad_batch(trial, l1s[active], l2s[active])
        rounds += 1
        for row, k in enumerate(active):
            model = models[k]
            model["fcalls"] += 1
            f, g, t = float(fs[row]), gs[row], model["t"]
            finite = np.isfinite(f) and bool(np.all(np.isfinite(g)))
            gd = float(g @ model["d"]) if finite else np.nan
            armijo = finite and f <= model["f"] + c1 * t * model["dg"]
            # the approximate Wolfe conditions of CG-DESCENT (src/lsearchk/cgdescent.cpp, epsilon 1e-6): once the
            # decrease of f drowns in its rounding, the slope alone decides
            approximate = (finite and f <= model["f"] + 1e-6 * abs(model["f"])
                           and (2.0 * c1 - 1.0) * model["dg"] >= gd >= c2 * model["dg"])
            if not approximate and not (armijo and gd >= c2 * model["dg"]):
                if armijo:
                    model["lo"] = t                                          # curvature fails: grow
                else:
                    model["hi"] = t                                          # Armijo fails: shrink
                model["t"] = 0.5 * (model["lo"] + model["hi"]) if np.isfinite(model["hi"]) else 2.0 * model["lo"]
                collapsed = np.isfinite(model["hi"]) and model["hi"] - model["lo"] <= 1e-16 * max(1.0, model["hi"])
                if collapsed or model["t"] > 1e300:
                    model["status"] = "failed"                               # no Wolfe point in the bracket
                continue
            s_k, y_k = trial[row] - model["x"], g - model["g"]               # a Wolfe point: accept the step
            if float(s_k @ y_k) > 1e-16 * float(np.linalg.norm(s_k)) * float(np.linalg.norm(y_k)):
                model["S"].append(s_k)
                model["Y"].append(y_k)
                if len(model["S"]) > history:
                    model["S"].pop(0)
                    model["Y"].pop(0)
            model["x"], model["f"], model["g"] = trial[row].copy(), f, g.copy()
            model["iterations"] += 1
            if gradient_test(f, g) < epsilon:
                model["status"] = "gradient_test"
            else:
                new_direction(model)
    return [{"x": model["x"], "fx": model["f"], "gradient_test": gradient_test(model["f"], model["g"]),
             "status": model["status"] or "max_iters", "fcalls": model["fcalls"], "gcalls": model["fcalls"],
             "iterations": model["iterations"], "rounds": rounds} for model in models]
