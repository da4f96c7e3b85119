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


def osga(function, x0, epsilon: float = 1e-8, max_evals: int = 1000, patience: int = 100, lambda_: float = 0.9,
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
