"""Host-side mirror of the gradient-boosting cousins of the linear objective, over the C ABI:

    nano::gboost::bias_function_t    include/nano/gboost/function.h:50-76,  src/gboost/function.cpp:67-126
    nano::gboost::scale_function_t   include/nano/gboost/function.h:78-108, src/gboost/function.cpp:128-223
    nano::gboost::grads_function_t   include/nano/gboost/function.h:10-48,  src/gboost/function.cpp:8-64

Same contract as every function_t here: `f = function(x)` / `f = function(x, gx)` with host numpy vectors, size checks
that throw, call counters; smooth / convex follow the loss (function.cpp:16-17, 74-75, 141-142). The targets (and the
strong / weak learners' outputs and the cluster assignment for the scale function) are pinned in HBM once."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .function import Context, Function, _ptr, critical
from .loss import Loss

KINDS = {"bias": 0, "scale": 1, "grads": 2}


class _GBoostFunction(Function):
    def __init__(self, kind: str, ctx: Context, loss: Loss, targets, cluster=None, groups: int = 0, soutputs=None,
                 woutputs=None):
        targets = np.ascontiguousarray(targets, dtype=np.float64)
        critical(targets.ndim >= 1 and targets.shape[0] > 0, "gboost: no samples")
        targets = targets.reshape(targets.shape[0], -1)
        N, T = targets.shape
        ids = None
        if kind == "scale":
            ids = np.ascontiguousarray(cluster, dtype=np.int32)
            soutputs = np.ascontiguousarray(soutputs, dtype=np.float64).reshape(N, -1)
            woutputs = np.ascontiguousarray(woutputs, dtype=np.float64).reshape(N, -1)
            critical(ids.shape == (N,), "gboost: one cluster id per sample")
            critical(soutputs.shape == (N, T) and woutputs.shape == (N, T),
                     "gboost: strong and weak outputs must match the targets")
        handle = ctypes.c_void_p()
        _lib.check(_lib.lib().nb200_gboost_create(
            ctx.handle, KINDS[kind], loss.kind, loss.param, _ptr(targets), N, T,
            None if ids is None else ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), int(groups),
            _ptr(soutputs) if kind == "scale" else None, _ptr(woutputs) if kind == "scale" else None,
            ctypes.byref(handle)))
        self.handle = handle
        self._fx = ctypes.c_double()
        super().__init__(f"gboost-{kind}", int(_lib.lib().nb200_gboost_size(handle)))
        self._convex = loss.convex()
        self._smooth = loss.smooth()

    def do_eval(self, x, gx) -> float:
        _lib.check(_lib.lib().nb200_gboost_vgrad(self.handle, _ptr(x), x.size, _ptr(gx), ctypes.byref(self._fx)))
        return self._fx.value

    def launches(self) -> int:
        return int(_lib.lib().nb200_gboost_launches(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().nb200_gboost_release(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GBoostBiasFunction(_GBoostFunction):
    """f(x) = E[loss(target_i, x)], x of size T: the bias of a gradient-boosting model."""

    def __init__(self, ctx: Context, loss: Loss, targets):
        super().__init__("bias", ctx, loss, targets)


class GBoostScaleFunction(_GBoostFunction):
    """f(x) = E[loss(target_i, soutput_i + x[cluster_i] * woutput_i)], x of size `groups`: the line-search-like step of a
    boosting round; samples with cluster id -1 keep scale 0 and do not contribute to the gradient."""

    def __init__(self, ctx: Context, loss: Loss, targets, cluster, groups: int, soutputs, woutputs):
        super().__init__("scale", ctx, loss, targets, cluster, groups, soutputs, woutputs)


class GBoostGradsFunction(_GBoostFunction):
    """f(outputs) = E[loss(target_i, output_i)], x of size N * T; the gradient holds the per-sample loss gradients / N."""

    def __init__(self, ctx: Context, loss: Loss, targets):
        super().__init__("grads", ctx, loss, targets)


__all__ = ["GBoostBiasFunction", "GBoostScaleFunction", "GBoostGradsFunction"]
