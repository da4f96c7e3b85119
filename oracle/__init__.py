"""CPU ORACLE bindings (ctypes over oracle/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package. The product (libnano_b200/) never does: it fails loudly when its CUDA library is missing.

Every function here is a thin view of the C++ restatement in oracle.cpp; see oracle.h for the reference
file:line each one follows and for the list of reference known-answers the oracle is pinned against.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

LOSS_KINDS = {
    "mse": 0,
    "mae": 1,
    "cauchy": 2,
    "hinge": 3,
    "squared-hinge": 4,
    "classnll": 5,
    "savage": 6,
    "tangent": 7,
    "logistic": 8,
    "exponential": 9,
    "pinball": 10,
}

SYNTH_LOSSES = {"mse": 0, "mae": 1, "cauchy": 2, "hinge": 3, "logistic": 4}
SYNTH_REGRESSION = {"mse": True, "mae": True, "cauchy": True, "hinge": False, "logistic": False}


def loss_kind(loss_id: str) -> int:
    """Map a libnano loss id (src/loss.cpp:105-135) to the oracle's kind: s-/m- prefixes share value/vgrad."""
    base = loss_id[2:] if loss_id[:2] in ("s-", "m-") else loss_id
    return LOSS_KINDS[base]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (building the checker is not using it)."""
    sources = [os.path.join(_HERE, f) for f in ("oracle.cpp", "solvers.cpp", "oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in sources if os.path.exists(s)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), f"expected shape {shape}, got {a.shape}"
    return a


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_loss_value.argtypes = [ctypes.c_int, ctypes.c_double, _dp, _dp, _i64, _i64, _dp]
        L.oracle_loss_vgrad.argtypes = [ctypes.c_int, ctypes.c_double, _dp, _dp, _i64, _i64, _dp]
        L.oracle_loss_error.argtypes = [ctypes.c_int, _dp, _dp, _i64, _i64, _dp]
        L.oracle_predict.argtypes = [_dp, _dp, _dp, _i64, _i64, _i64, _dp]
        L.oracle_sum_reduce.argtypes = [_dp, _i64, _i64, _i64]
        for name in ("oracle_linear_vgrad", "oracle_linear_vgrad_fast"):
            fn = getattr(L, name)
            fn.restype = ctypes.c_double
            fn.argtypes = [_dp, _dp, _i64, _i64, _i64, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                           ctypes.c_double, _dp, _dp, _i64, ctypes.c_int]
        L.oracle_linear_vgrad_ld.restype = ctypes.c_double
        L.oracle_linear_vgrad_ld.argtypes = [_dp, _dp, _i64, _i64, _i64, ctypes.c_int, ctypes.c_double,
                                             ctypes.c_double, ctypes.c_double, _dp, _dp]
        L.oracle_linear_partial.argtypes = [_dp, _dp, _i64, _i64, _i64, ctypes.c_int, ctypes.c_double, _dp, ctypes.c_int,
                                            _i64, ctypes.c_int, _dp]
        L.oracle_linear_finish.restype = ctypes.c_double
        L.oracle_linear_finish.argtypes = [_dp, _i64, _i64, _i64, ctypes.c_double, ctypes.c_double, _dp, _dp]
        L.oracle_scalar_stats.argtypes = [_dp, _i64, _i64, ctypes.c_char_p, _dp]
        L.oracle_scale.argtypes = [_dp, _i64, _i64, _dp, ctypes.c_int]
        L.oracle_upscale.argtypes = [_dp, ctypes.c_int, _dp, ctypes.c_int, _i64, _i64, _dp, _dp]
        L.oracle_synth_samples.restype = _i64
        L.oracle_synth_samples.argtypes = [_i64, ctypes.c_double]
        L.oracle_synth_model.argtypes = [_i64, _i64, ctypes.c_uint64, _i64, ctypes.c_int, _dp, _dp, _dp, _dp]
        L.oracle_synth_vgrad.restype = ctypes.c_double
        L.oracle_synth_vgrad.argtypes = [_dp, _dp, ctypes.c_double, _i64, _i64, ctypes.c_int, ctypes.c_double,
                                         ctypes.c_double, _dp, _dp]
        L.oracle_hardware_threads.restype = ctypes.c_int
        L.oracle_gboost_vgrad.restype = ctypes.c_double
        L.oracle_gboost_vgrad.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, _dp, _i64, _i64,
                                          ctypes.POINTER(ctypes.c_int32), _i64, _dp, _dp, _dp, _dp, _i64]
        # restated drivers (solvers.cpp)
        L.oracle_solver_callback = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_void_p, _dp, _dp)
        L.oracle_lbfgs.restype = ctypes.c_int
        L.oracle_lbfgs.argtypes = [L.oracle_solver_callback, ctypes.c_void_p, _i64, _dp, ctypes.c_double, _i64, _i64,
                                   _dp, _dp, ctypes.POINTER(_i64)]
        L.oracle_cgd.restype = ctypes.c_int
        L.oracle_cgd.argtypes = [L.oracle_solver_callback, ctypes.c_void_p, _i64, _dp, ctypes.c_int, ctypes.c_double, _i64,
                                 ctypes.c_double, ctypes.c_double, _dp, _dp, ctypes.POINTER(_i64)]
        L.oracle_osga.restype = ctypes.c_int
        L.oracle_osga.argtypes = [L.oracle_solver_callback, ctypes.c_void_p, _i64, _dp, ctypes.c_double,
                                  ctypes.c_double, _i64, _i64, _dp, _dp, ctypes.POINTER(_i64)]
        _lib = L
    return _lib


# ----------------------------------------------------------------------------------------------------------------
# loss_t::value / loss_t::vgrad
# ----------------------------------------------------------------------------------------------------------------
def loss_value(loss_id: str, targets, outputs, param: float = 0.5):
    targets = _f64(targets)
    outputs = _f64(outputs)
    if targets.shape != outputs.shape:  # src/loss.cpp:10-16 (critical on dims mismatch)
        raise RuntimeError("critical check failed: loss: incompatible dimension-wise outputs and targets")
    N = targets.shape[0]
    T = int(np.prod(targets.shape[1:])) if targets.ndim > 1 else 1
    values = np.empty(N, dtype=np.float64)
    lib().oracle_loss_value(loss_kind(loss_id), param, _ptr(targets), _ptr(outputs), N, T, _ptr(values))
    return values


def loss_vgrad(loss_id: str, targets, outputs, param: float = 0.5):
    targets = _f64(targets)
    outputs = _f64(outputs)
    if targets.shape != outputs.shape:
        raise RuntimeError("critical check failed: loss: incompatible dimension-wise outputs and targets")
    N = targets.shape[0]
    T = int(np.prod(targets.shape[1:])) if targets.ndim > 1 else 1
    vgrads = np.empty_like(targets)
    lib().oracle_loss_vgrad(loss_kind(loss_id), param, _ptr(targets), _ptr(outputs), N, T, _ptr(vgrads))
    return vgrads


ERROR_ABSDIFF, ERROR_SCLASS, ERROR_MCLASS = 0, 1, 2


def error_kind(loss_id: str) -> int:
    """The terror policy behind a loss id (include/nano/loss/flatten.h:103-122): s- / m- prefix, else regression."""
    return ERROR_SCLASS if loss_id.startswith("s-") else ERROR_MCLASS if loss_id.startswith("m-") else ERROR_ABSDIFF


def loss_error(loss_id: str, targets, outputs, param: float = 0.5):
    """loss_t::error (src/loss.cpp:50-55): per-sample error; pinball's is its value (src/loss/pinball.cpp:20-23)."""
    if loss_id == "pinball":
        return loss_value(loss_id, targets, outputs, param)
    targets = _f64(targets)
    outputs = _f64(outputs)
    if targets.shape != outputs.shape:
        raise RuntimeError("critical check failed: loss: incompatible dimension-wise outputs and targets")
    N = targets.shape[0]
    T = int(np.prod(targets.shape[1:])) if targets.ndim > 1 else 1
    errors = np.empty(N, dtype=np.float64)
    lib().oracle_loss_error(error_kind(loss_id), _ptr(targets), _ptr(outputs), N, T, _ptr(errors))
    return errors


def evaluate(X, Y, loss_id: str, W, b, param: float = 0.5):
    """linear::evaluate (src/linear/util.cpp:30-49): [2, N] = per-sample (errors, loss values) of the predictions."""
    outputs = predict(X, W, b)
    Y = _f64(Y).reshape(outputs.shape)
    return np.stack([loss_error(loss_id, Y, outputs, param), loss_value(loss_id, Y, outputs, param)])


def predict(X, W, b):
    X = _f64(X)
    W = _f64(W)
    b = _f64(b)
    N, D = X.shape
    T = W.shape[0]
    assert W.shape == (T, D) and b.shape == (T,)
    out = np.empty((N, T), dtype=np.float64)
    lib().oracle_predict(_ptr(X), _ptr(W), _ptr(b), N, D, T, _ptr(out))
    return out


def sum_reduce(accumulators, samples: int):
    acc = _f64(accumulators).copy()
    count, length = acc.shape
    lib().oracle_sum_reduce(_ptr(acc), count, length, samples)
    return acc[0]


# ----------------------------------------------------------------------------------------------------------------
# objective (A): nano::linear::function_t
# ----------------------------------------------------------------------------------------------------------------
def linear_vgrad(X, Y, loss_id, x, l1=0.0, l2=0.0, want_grad=True, batch=100, threads=1, param=0.5,
                 mode="plain"):
    """f, g of linear::function_t::do_eval. mode: 'plain' (oracle), 'fast' (CPU baseline), 'ld' (long double)."""
    X = _f64(X)
    N, D = X.shape
    Y = _f64(Y).reshape(N, -1)
    T = Y.shape[1]
    x = _f64(x, ((D + 1) * T,))
    gx = np.empty_like(x) if want_grad else None
    kind = loss_kind(loss_id)
    if mode == "ld":
        fx = lib().oracle_linear_vgrad_ld(_ptr(X), _ptr(Y), N, D, T, kind, param, l1, l2, _ptr(x), _ptr(gx))
    else:
        fn = lib().oracle_linear_vgrad_fast if mode == "fast" else lib().oracle_linear_vgrad
        fx = fn(_ptr(X), _ptr(Y), N, D, T, kind, param, l1, l2, _ptr(x), _ptr(gx), batch, threads)
    return fx, gx


def linear_partial(X, Y, loss_id, x, want_grad=True, batch=100, threads=1, param=0.5):
    """Un-normalised [sum loss | sum dL | sum dL x^T] of the given rows (what one rank contributes)."""
    X = _f64(X)
    N, D = X.shape
    Y = _f64(Y).reshape(N, -1)
    T = Y.shape[1]
    x = _f64(x, ((D + 1) * T,))
    partial = np.zeros(1 + T + T * D, dtype=np.float64)
    lib().oracle_linear_partial(_ptr(X), _ptr(Y), N, D, T, loss_kind(loss_id), param, _ptr(x), int(want_grad), batch,
                                threads, _ptr(partial))
    return partial


def linear_finish(reduced, n_global, D, T, x, l1=0.0, l2=0.0, want_grad=True):
    reduced = _f64(reduced, (1 + T + T * D,))
    x = _f64(x, ((D + 1) * T,))
    gx = np.empty_like(x) if want_grad else None
    fx = lib().oracle_linear_finish(_ptr(reduced), n_global, D, T, l1, l2, _ptr(x), _ptr(gx))
    return fx, gx


# ----------------------------------------------------------------------------------------------------------------
# objective (B): function_{ridge,lasso,elasticnet}_t<loss> over linear_model_t
# ----------------------------------------------------------------------------------------------------------------
def synth_samples(dims: int, sratio: float = 10.0) -> int:
    return int(lib().oracle_synth_samples(dims, sratio))


def synth_model(dims: int, seed: int = 42, sratio: float = 10.0, modulo: int = 1, loss: str = "mse"):
    D = max(dims, 2)  # make_inputs (src/function/mlearn/ridge.cpp:13-16)
    N = synth_samples(dims, sratio)
    X = np.empty((N, D), dtype=np.float64)
    Y = np.empty(N, dtype=np.float64)
    w = np.empty(D, dtype=np.float64)
    b = np.empty(1, dtype=np.float64)
    lib().oracle_synth_model(N, D, seed, modulo, int(SYNTH_REGRESSION[loss]), _ptr(X), _ptr(Y), _ptr(w), _ptr(b))
    return X, Y, w, float(b[0])


def synth_vgrad(X, Y, bopt, loss, x, alpha1=0.0, alpha2=0.0, want_grad=True):
    X = _f64(X)
    N, D = X.shape
    Y = _f64(Y, (N,))
    x = _f64(x, (D,))
    gx = np.empty_like(x) if want_grad else None
    fx = lib().oracle_synth_vgrad(_ptr(X), _ptr(Y), bopt, N, D, SYNTH_LOSSES[loss], alpha1, alpha2, _ptr(x), _ptr(gx))
    return fx, gx


SCALINGS = {"none": 0, "mean": 1, "minmax": 2, "standard": 3}
STAT_NAMES = ("samples", "min", "max", "mean", "stdev", "div_range", "div_stdev")


def scalar_stats(values, enable=None):
    """scalar_stats_t over the columns of values[N, C] -> [7, C] array (rows: STAT_NAMES)."""
    values = _f64(values)
    N, C = values.shape
    stats = np.empty((7, C), dtype=np.float64)
    mask = None if enable is None else np.ascontiguousarray(enable, dtype=np.uint8).tobytes()
    lib().oracle_scalar_stats(_ptr(values), N, C, mask, _ptr(stats))
    return stats


def scale(values, stats, scaling: str):
    out = _f64(values).copy()
    N, C = out.shape
    lib().oracle_scale(_ptr(out), N, C, _ptr(_f64(stats, (7, C))), SCALINGS[scaling])
    return out


def upscale(istats, iscaling: str, tstats, tscaling: str, W, b):
    W = _f64(W).copy()
    b = _f64(b).copy()
    T, D = W.shape
    lib().oracle_upscale(_ptr(_f64(istats, (7, D))), SCALINGS[iscaling], _ptr(_f64(tstats, (7, T))), SCALINGS[tscaling],
                         D, T, _ptr(W), _ptr(b))
    return W, b


GBOOST_KINDS = {"bias": 0, "scale": 1, "grads": 2}


def gboost_vgrad(kind: str, loss_id: str, targets, x, cluster=None, groups=0, soutputs=None, woutputs=None,
                 want_grad=True, batch=100, param=0.5):
    """f, g of gboost::{bias,scale,grads}_function_t::do_eval (src/gboost/function.cpp) over targets[N, T]."""
    targets = _f64(targets)
    N, T = targets.shape
    n = {"bias": T, "scale": groups, "grads": N * T}[kind]
    x = _f64(x, (n,))
    gx = np.empty_like(x) if want_grad else None
    ids = None
    if kind == "scale":
        ids = np.ascontiguousarray(cluster, dtype=np.int32)
        soutputs, woutputs = _f64(soutputs, (N, T)), _f64(woutputs, (N, T))
    fx = lib().oracle_gboost_vgrad(GBOOST_KINDS[kind], loss_kind(loss_id), param, _ptr(targets), N, T,
                                   None if ids is None else ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), groups,
                                   _ptr(soutputs) if kind == "scale" else None,
                                   _ptr(woutputs) if kind == "scale" else None, _ptr(x), _ptr(gx), batch)
    return fx, gx


def hardware_threads() -> int:
    return int(lib().oracle_hardware_threads())


# ----------------------------------------------------------------------------------------------------------------
# restated drivers (solvers.cpp): L-BFGS + CG-DESCENT and OSGA over any `function(x, gx=None) -> fx` callable
# ----------------------------------------------------------------------------------------------------------------
SOLVER_STATUS = {0: "max_iters", 1: "gradient_test", 2: "value_test", 3: "specific_test", 4: "failed"}


def _make_callback(function, n):
    def trampoline(_ctx, x_ptr, g_ptr):
        x = np.ctypeslib.as_array(x_ptr, shape=(n,)).copy()
        if g_ptr:
            gx = np.empty(n, dtype=np.float64)
            fx = function(x, gx)
            ctypes.memmove(g_ptr, gx.ctypes.data, 8 * n)
            return float(fx)
        return float(function(x))

    return lib().oracle_solver_callback(trampoline)


def lbfgs(function, x0, epsilon=1e-8, max_evals=1000, history=50):
    """solver_lbfgs_t::minimize with the reference defaults (src/solver.cpp:45-51, src/solver/lbfgs.cpp:9-12)."""
    x = _f64(x0).copy()
    n = x.size
    fx, gtest = ctypes.c_double(), ctypes.c_double()
    stats = (_i64 * 3)()
    callback = _make_callback(function, n)
    status = lib().oracle_lbfgs(callback, None, n, _ptr(x), epsilon, max_evals, history, ctypes.byref(fx),
                                ctypes.byref(gtest), stats)
    return {"x": x, "fx": fx.value, "gradient_test": gtest.value, "status": SOLVER_STATUS[status],
            "fcalls": stats[0], "gcalls": stats[1], "iterations": stats[2]}


CGD_BETAS = {"hs": 0, "fr": 1, "pr": 2, "cd": 3, "ls": 4, "dy": 5, "n": 6, "dycd": 7, "dyhs": 8, "frpr": 9}


def cgd(function, x0, beta="pr", epsilon=1e-8, max_evals=1000, orthotest=0.1, eta=0.01):
    """solver_cgd_t::minimize, ids cgd-<beta> (src/solver/cgd.cpp:73-143; cgd-pr is libnano's default "cgd")."""
    x = _f64(x0).copy()
    n = x.size
    fx, gtest = ctypes.c_double(), ctypes.c_double()
    stats = (_i64 * 3)()
    callback = _make_callback(function, n)
    status = lib().oracle_cgd(callback, None, n, _ptr(x), CGD_BETAS[beta], epsilon, max_evals, orthotest, eta,
                              ctypes.byref(fx), ctypes.byref(gtest), stats)
    return {"x": x, "fx": fx.value, "gradient_test": gtest.value, "status": SOLVER_STATUS[status],
            "fcalls": stats[0], "gcalls": stats[1], "iterations": stats[2]}


def osga(function, x0, epsilon=1e-8, strong_convexity=0.0, max_evals=1000, patience=100):
    """solver_osga_t::minimize (src/solver/osga.cpp:59-147)."""
    x = _f64(x0).copy()
    n = x.size
    fx, eta = ctypes.c_double(), ctypes.c_double()
    stats = (_i64 * 3)()
    callback = _make_callback(function, n)
    status = lib().oracle_osga(callback, None, n, _ptr(x), epsilon, strong_convexity, max_evals, patience,
                               ctypes.byref(fx), ctypes.byref(eta), stats)
    return {"x": x, "fx": fx.value, "eta": eta.value, "status": SOLVER_STATUS[status], "fcalls": stats[0],
            "gcalls": stats[1], "iterations": stats[2]}
