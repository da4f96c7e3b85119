"""Device-resident L-BFGS (nb200_lbfgs_minimize): libnano's solver_lbfgs_t with its default line search
(src/solver/lbfgs.cpp:25-124, src/lsearchk/cgdescent.cpp, src/lsearch0/quadratic.cpp) keeping x, g, the (s, y)
history, the two-loop recursion and the trial points in HBM. libnano's own solvers still work unchanged through
`function(x, gx)`; this entry is for callers that want the PCIe round trip per evaluation gone."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .function import critical

STATUS = {0: "max_iters", 1: "gradient_test", 4: "failed"}


def lbfgs(function, x0, epsilon: float = 1e-8, max_evals: int = 1000, history: int = 50):
    """Minimise a GPU objective (LinearFunction, SyntheticFunction or sharded.ShardedFunction) from x0.

    Defaults are the reference's (src/solver.cpp:45-51: epsilon 1e-8, max_evals 1000; lbfgs.cpp:12: history 50).
    Returns a dict like solver_state_t: x, fx, gradient_test, status, fcalls, gcalls, iterations."""
    from .sharded import ShardedFunction

    x = np.ascontiguousarray(x0, dtype=np.float64).copy()
    critical(x.ndim == 1 and x.size == function.size(), "solver: incompatible initial point (", x.size,
             " dimensions), expecting ", function.size(), " dimensions!")

    callback, n_global = None, 0
    local = function
    if isinstance(function, ShardedFunction):
        shard = function._shard
        local = shard.local
        if function.world_size() > 1:
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
    _lib.check(_lib.lib().nb200_lbfgs_minimize(
        local.handle, x.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), x.size, float(epsilon), int(max_evals),
        int(history), int(n_global), ctypes.cast(callback, ctypes.c_void_p) if callback else None, None,
        ctypes.byref(fx), ctypes.byref(gtest), stats, ctypes.byref(status)))
    function._fcalls += int(stats[0])
    function._gcalls += int(stats[1])
    return {"x": x, "fx": fx.value, "gradient_test": gtest.value, "status": STATUS.get(status.value, "failed"),
            "fcalls": int(stats[0]), "gcalls": int(stats[1]), "iterations": int(stats[2])}
