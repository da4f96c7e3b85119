"""Mirror of nano::loss_t and its registry (include/nano/loss.h, src/loss.cpp:105-135): ids, convex/smooth flags and
the device kind each id evaluates with. value()/vgrad() run on the GPU through the C ABI."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

# device kinds (include/nb200.h: nb200_loss)
KIND_MSE, KIND_MAE, KIND_CAUCHY, KIND_HINGE, KIND_SQUARED_HINGE, KIND_CLASSNLL = 0, 1, 2, 3, 4, 5
KIND_SAVAGE, KIND_TANGENT, KIND_LOGISTIC, KIND_EXPONENTIAL, KIND_PINBALL = 6, 7, 8, 9, 10
KIND_SYNTH_CAUCHY, KIND_SYNTH_LOGISTIC = 11, 12

# error measures (include/nb200.h: nb200_error; include/nano/loss/error.h)
ERROR_ABSDIFF, ERROR_SCLASS, ERROR_MCLASS = 0, 1, 2

# base name -> (kind, convex, smooth): the `static constexpr` flags of include/nano/loss/*.h
_BASE = {
    "mse": (KIND_MSE, True, True),
    "mae": (KIND_MAE, True, False),
    "cauchy": (KIND_CAUCHY, False, True),
    "hinge": (KIND_HINGE, True, False),
    "squared-hinge": (KIND_SQUARED_HINGE, True, False),
    "classnll": (KIND_CLASSNLL, True, True),
    "savage": (KIND_SAVAGE, False, True),
    "tangent": (KIND_TANGENT, False, True),
    "logistic": (KIND_LOGISTIC, True, True),
    "exponential": (KIND_EXPONENTIAL, True, True),
    "pinball": (KIND_PINBALL, True, False),
}

# the 17 registered ids (src/loss.cpp:105-135)
LOSS_IDS = (
    "mae", "mse", "cauchy",
    "m-hinge", "s-hinge", "m-squared-hinge", "s-squared-hinge", "s-classnll",
    "m-savage", "s-savage", "m-tangent", "s-tangent", "m-logistic", "s-logistic",
    "s-exponential", "m-exponential", "pinball",
)


class Loss:
    """nano::loss_t: `value(targets, outputs)` -> (N,), `vgrad(targets, outputs)` -> targets.shape."""

    def __init__(self, loss_id: str):
        if loss_id not in LOSS_IDS:
            raise RuntimeError(f"critical check failed: loss: unknown id <{loss_id}>")
        base = loss_id[2:] if loss_id[:2] in ("s-", "m-") else loss_id
        self._id = loss_id
        self.kind, self._convex, self._smooth = _BASE[base]
        # the terror policy of flatten_loss_t (include/nano/loss/flatten.h:103-122): s- / m- prefix, else regression
        self.error_kind = {"s-": ERROR_SCLASS, "m-": ERROR_MCLASS}.get(loss_id[:2], ERROR_ABSDIFF)
        self._params = {"loss::pinball::alpha": 0.5} if base == "pinball" else {}

    @staticmethod
    def all():
        return LOSS_IDS

    def type_id(self) -> str:
        return self._id

    def convex(self) -> bool:
        return self._convex

    def smooth(self) -> bool:
        return self._smooth

    def clone(self) -> "Loss":
        other = Loss(self._id)
        other._params = dict(self._params)
        return other

    def parameter(self, name: str, value=None):
        """configurable_t::parameter: get, or set when `value` is given (range-checked like parameter_t)."""
        if name not in self._params:
            raise RuntimeError(f"critical check failed: configurable: unknown parameter <{name}>")
        if value is not None:
            if not (0.0 <= float(value) <= 1.0):  # src/loss/pinball.cpp:13 : 0 <= alpha <= 1
                raise RuntimeError(f"critical check failed: parameter {name}: {value} out of [0, 1]")
            self._params[name] = float(value)
        return self._params[name]

    @property
    def param(self) -> float:
        return self._params.get("loss::pinball::alpha", 0.0)

    # -- evaluation on the device ------------------------------------------------------------------------------
    def _check(self, targets, outputs):
        targets = np.ascontiguousarray(targets, dtype=np.float64)
        outputs = np.ascontiguousarray(outputs, dtype=np.float64)
        if targets.shape != outputs.shape:  # src/loss.cpp:10-16
            raise RuntimeError(f"critical check failed: loss: incompatible dimension-wise outputs and targets "
                               f"({outputs.shape}) vs. ({targets.shape})!")
        N = targets.shape[0] if targets.ndim else 0
        T = int(np.prod(targets.shape[1:])) if targets.ndim > 1 else 1
        return targets, outputs, N, T

    def value(self, ctx, targets, outputs):
        targets, outputs, N, T = self._check(targets, outputs)
        values = np.empty(N, dtype=np.float64)
        _lib.check(_lib.lib().nb200_loss_value(ctx.handle, self.kind, self.param, _ptr(targets), _ptr(outputs), N, T,
                                               _ptr(values)))
        return values

    def error(self, ctx, targets, outputs):
        """loss_t::error (src/loss.cpp:50-55): L1 distance (regression), 0-1 error of the arg-max label (s-),
        number of mis-predicted labels (m-); pinball reports its value (src/loss/pinball.cpp:20-23)."""
        if self.kind == KIND_PINBALL:
            return self.value(ctx, targets, outputs)
        targets, outputs, N, T = self._check(targets, outputs)
        errors = np.empty(N, dtype=np.float64)
        _lib.check(_lib.lib().nb200_loss_error(ctx.handle, self.error_kind, _ptr(targets), _ptr(outputs), N, T,
                                               _ptr(errors)))
        return errors

    def vgrad(self, ctx, targets, outputs):
        targets, outputs, N, T = self._check(targets, outputs)
        vgrads = np.empty_like(targets)
        _lib.check(_lib.lib().nb200_loss_vgrad(ctx.handle, self.kind, self.param, _ptr(targets), _ptr(outputs), N, T,
                                               _ptr(vgrads)))
        return vgrads


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
