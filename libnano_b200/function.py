"""Host-side mirror of the reference interface for the hot path, over the C ABI (include/nb200.h).

    nano::function_t                       include/nano/function.h:28-216, src/function.cpp:246-332   -> Function
    nano::linear::function_t               include/nano/linear/function.h:19-80, src/linear/function.cpp -> LinearFunction
    function_{ridge,lasso,elasticnet}_t    src/function/mlearn/{ridge,lasso,elasticnet}.cpp            -> SyntheticFunction
    flatten_iterator_t::cache_flatten      src/dataset/iterator.cpp:118-147                             -> Dataset

Same names, argument meaning and error behaviour (critical() -> RuntimeError) so that the parity tests read like the
reference's own tests. All numerics run in libnb200.so on the GPU; numpy only carries host buffers.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .loss import KIND_SYNTH_CAUCHY, KIND_SYNTH_LOGISTIC, Loss

FLAVOUR_LINEAR = 0
FLAVOUR_SYNTH = 1

SYNTH_CLASSIFICATION = 0
SYNTH_REGRESSION = 1
SYNTH_MULTICLASS = 2

_dp = ctypes.POINTER(ctypes.c_double)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def critical(condition: bool, *message) -> None:
    """nano::critical (include/nano/critical.h:13-29)."""
    if not condition:
        raise RuntimeError("critical check failed: " + "".join(str(m) for m in message))


class Context:
    """One CUDA device (nb200_ctx) — or, given a list of devices, a device GROUP: datasets pinned on it are split in
    contiguous row blocks over the devices and objectives on them evaluate on all of them from this one process (what
    libnano's thread pool does with the workers of one process, src/linear/function.cpp:53-57)."""

    def __init__(self, device=0):
        handle = ctypes.c_void_p()
        if isinstance(device, (list, tuple)):
            devices = (ctypes.c_int * len(device))(*[int(d) for d in device])
            _lib.check(_lib.lib().nb200_init_multi(len(device), devices, ctypes.byref(handle)))
            self.devices = [int(d) for d in device]
            self.device = self.devices[0]
        else:
            _lib.check(_lib.lib().nb200_init(device, ctypes.byref(handle)))
            self.devices = [int(device)]
            self.device = device
        self.handle = handle

    PEAKS = {"dmma_tflops": 0, "dfma_tflops": 1, "hbm_read_gbs": 2}

    def probe_peak(self, what: str) -> float:
        """A roofline denominator of this board measured now: 'dmma_tflops' (fp64 tensor pipe), 'dfma_tflops' (fp64 FMA
        pipe) or 'hbm_read_gbs' (read-only bulk-copy stream)."""
        value = ctypes.c_double()
        _lib.check(_lib.lib().nb200_probe_peak(self.handle, self.PEAKS[what], ctypes.byref(value)))
        return value.value

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().nb200_ctx_release(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Dataset:
    """The feature matrix X[N, D] and targets Y[N, T] pinned in HBM once per fit (what
    flatten_iterator_t::cache_flatten / cache_targets hold on the host, src/dataset/iterator.cpp:118-147,52-77)."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.handle = handle

    @classmethod
    def pin(cls, ctx: Context, X, Y) -> "Dataset":
        X = np.ascontiguousarray(X, dtype=np.float64)
        critical(X.ndim == 2, "dataset: inputs must be a (samples, features) matrix")
        N, D = X.shape
        critical(N > 0 and D > 0, "dataset: sizes must be positive")
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        critical(Y.shape[0] == N if Y.ndim else False, "dataset: inputs and targets disagree on the sample count")
        Y = Y.reshape(N, -1)
        handle = ctypes.c_void_p()
        _lib.check(_lib.lib().nb200_dataset_pin(ctx.handle, _ptr(X), _ptr(Y), N, D, Y.shape[1], ctypes.byref(handle)))
        return cls(ctx, handle)

    @classmethod
    def synth(cls, ctx: Context, kind: int, N: int, D: int, T: int = 1, seed: int = 42, row_offset: int = 0,
              noise: float = 0.1) -> "Dataset":
        handle = ctypes.c_void_p()
        _lib.check(_lib.lib().nb200_dataset_synth(ctx.handle, seed, kind, N, D, T, row_offset, noise,
                                                  ctypes.byref(handle)))
        return cls(ctx, handle)

    @property
    def samples(self) -> int:
        return int(_lib.lib().nb200_dataset_samples(self.handle))

    @property
    def inputs(self) -> int:
        return int(_lib.lib().nb200_dataset_inputs(self.handle))

    @property
    def targets(self) -> int:
        return int(_lib.lib().nb200_dataset_targets(self.handle))

    def read(self, row0: int = 0, rows: int | None = None):
        rows = self.samples - row0 if rows is None else rows
        X = np.empty((rows, self.inputs), dtype=np.float64)
        Y = np.empty((rows, self.targets), dtype=np.float64)
        _lib.check(_lib.lib().nb200_dataset_read(self.handle, row0, rows, _ptr(X), _ptr(Y)))
        return X, Y

    def read_targets(self, row0: int = 0, rows: int | None = None):
        """The targets of rows [row0, row0 + rows) without their inputs."""
        rows = self.samples - row0 if rows is None else rows
        Y = np.empty((rows, self.targets), dtype=np.float64)
        _lib.check(_lib.lib().nb200_dataset_read(self.handle, row0, rows, None, _ptr(Y)))
        return Y

    SCALINGS = {"none": 0, "mean": 1, "minmax": 2, "standard": 3}  # nano::scaling_type (dataset/scaling.h:10-16)

    def scale(self, inputs: str = "standard", targets: str = "none", enable_inputs=None, enable_targets=None):
        """flatten_iterator_t::scaling (src/dataset/stats.cpp:81-137,357-401) on the device, once: column statistics
        over the finite values, in-place scaling, non-finite -> 0. `enable_*`: per-column masks (False = categorical
        column / class target, left unscaled). Returns self."""
        critical(inputs in self.SCALINGS and targets in self.SCALINGS, "dataset: unhandled scaling type")

        def mask(values, size):
            if values is None:
                return None
            values = np.ascontiguousarray(values, dtype=np.uint8)
            critical(values.size == size, "dataset: scaling mask of the wrong size")
            return values.tobytes()

        _lib.check(_lib.lib().nb200_dataset_scale(self.handle, self.SCALINGS[inputs], self.SCALINGS[targets],
                                                  mask(enable_inputs, self.inputs), mask(enable_targets, self.targets)))
        return self

    def stats(self, which: str = "inputs"):
        """scalar_stats_t of the inputs or the targets (after scale()): dict of arrays."""
        size = self.inputs if which == "inputs" else self.targets
        names = ("samples", "min", "max", "mean", "stdev", "div_range", "div_stdev")
        arrays = [np.empty(size, dtype=np.float64) for _ in names]
        _lib.check(_lib.lib().nb200_dataset_stats(self.handle, 0 if which == "inputs" else 1, *[_ptr(a) for a in arrays]))
        return dict(zip(names, arrays))

    def upscale(self, W, b):
        """nano::upscale (src/dataset/stats.cpp:211-230): (W, b) fitted on the scaled data -> original units."""
        W = np.array(W, dtype=np.float64).reshape(self.targets, self.inputs)
        b = np.array(b, dtype=np.float64).reshape(self.targets)
        _lib.check(_lib.lib().nb200_upscale(self.handle, _ptr(W), _ptr(b)))
        return W, b

    def predict(self, W, b):
        """linear::predict (src/linear/util.cpp:6-21): outputs[N, T] = X W^T + b."""
        W = np.ascontiguousarray(W, dtype=np.float64).reshape(self.targets, self.inputs)
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(self.targets)
        out = np.empty((self.samples, self.targets), dtype=np.float64)
        _lib.check(_lib.lib().nb200_predict(self.ctx.handle, self.handle, _ptr(W), _ptr(b), _ptr(out)))
        return out

    def evaluate(self, loss: Loss, W, b):
        """linear::evaluate (src/linear/util.cpp:30-49): [2, N] = per-sample (errors, loss values) of the predictions
        X W^T + b against the pinned targets; the outputs never leave HBM."""
        W = np.ascontiguousarray(W, dtype=np.float64).reshape(self.targets, self.inputs)
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(self.targets)
        out = np.empty((2, self.samples), dtype=np.float64)
        _lib.check(_lib.lib().nb200_evaluate(self.ctx.handle, self.handle, loss.kind, loss.param, loss.error_kind,
                                             _ptr(W), _ptr(b), _ptr(out[0]), _ptr(out[1])))
        return out

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().nb200_dataset_release(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Function:
    """nano::function_t: `f = function(x)` / `f = function(x, gx)` with host numpy vectors.

    Keeps the non-virtual contract of function_t::operator() (src/function.cpp:246-272): size checks that throw,
    fcalls/gcalls counters, then dispatch to do_eval."""

    def __init__(self, type_id: str, size: int):
        self._type_id = type_id
        self._size = int(size)
        self._convex = False
        self._smooth = False
        self._strong_convexity = 0.0
        self._fcalls = 0
        self._gcalls = 0
        self._hcalls = 0

    # -- attributes ------------------------------------------------------------------------------------------------
    def type_id(self) -> str:
        return self._type_id

    def name(self, with_size: bool = True) -> str:
        return f"{self.do_name()}[{self.size()}D]" if with_size else self.do_name()

    def do_name(self) -> str:
        return self._type_id

    def size(self) -> int:
        return self._size

    def convex(self) -> bool:
        return self._convex

    def smooth(self) -> bool:
        return self._smooth

    def strong_convexity(self) -> float:
        return self._strong_convexity

    def fcalls(self) -> int:
        return self._fcalls

    def gcalls(self) -> int:
        return self._gcalls

    def hcalls(self) -> int:
        return self._hcalls

    def clear_statistics(self) -> None:
        self._fcalls = self._gcalls = self._hcalls = 0

    # -- evaluation ------------------------------------------------------------------------------------------------
    def __call__(self, x, gx=None, hx=None) -> float:
        x = np.asarray(x)
        critical(x.ndim == 1 and x.size == self.size(), "function: invalid input size, expecting (", self.size(),
                 ",), got (", x.size, ",) instead!")
        gsize = 0 if gx is None else np.asarray(gx).size
        critical(gsize == 0 or gsize == self.size(), "function: invalid gradient size, expecting (", self.size(),
                 ",) or empty, got (", gsize, ",) instead!")
        hsize = 0 if hx is None else np.asarray(hx).size
        critical(hsize == 0 or np.asarray(hx).shape == (self.size(), self.size()),
                 "function: invalid Hessian size, expecting (", self.size(), ", ", self.size(), ") or empty!")
        critical(hsize == 0 or self.smooth(), "function: cannot compute the Hessian for non-smooth functions!")
        critical(hsize == 0, "function: the Hessian is outside the GPU hot path (SURVEY.md §8 a14)")
        if gsize:
            critical(isinstance(gx, np.ndarray) and gx.dtype == np.float64 and gx.flags.c_contiguous and
                     gx.flags.writeable, "function: the gradient buffer must be a writable contiguous float64 array")
        self._fcalls += 1
        self._gcalls += 1 if gsize == self.size() else 0
        x = np.ascontiguousarray(x, dtype=np.float64)
        return self.do_eval(x, gx if gsize else None)

    def do_eval(self, x, gx) -> float:  # pragma: no cover - interface
        raise NotImplementedError

    def clone(self) -> "Function":  # pragma: no cover - interface
        raise NotImplementedError


class _DeviceFunction(Function):
    """A function_t whose do_eval is one nb200_vgrad call."""

    def __init__(self, type_id: str, dataset: Dataset, kind: int, loss_param: float, flavour: int, l1: float,
                 l2: float, fixed_bias=None, _clone_of=None):
        self._dataset = dataset
        handle = ctypes.c_void_p()
        if _clone_of is not None:
            _lib.check(_lib.lib().nb200_objective_clone(_clone_of, ctypes.byref(handle)))
        else:
            bias = None if fixed_bias is None else np.ascontiguousarray(fixed_bias, dtype=np.float64)
            _lib.check(_lib.lib().nb200_objective_create(dataset.handle, kind, loss_param, flavour, l1, l2, _ptr(bias),
                                                         ctypes.byref(handle)))
        self.handle = handle
        self._fx = ctypes.c_double()
        super().__init__(type_id, int(_lib.lib().nb200_objective_size(handle)))

    def do_eval(self, x, gx) -> float:
        _lib.check(_lib.lib().nb200_vgrad(self.handle, _ptr(x), x.size, _ptr(gx), ctypes.byref(self._fx)))
        return self._fx.value

    # -- many models in one pass over X ---------------------------------------------------------------
    def vgrad_batch(self, xs, l1s, l2s, want_grad: bool = True):
        """K single-target models sharing this objective's data and loss, differing in (x_k, l1_k, l2_k): returns
        (fs[K], gs[K, n] or None). What the tuner's grid of regularisation values costs the reference K passes over X
        (src/machine/tune.cpp:25-42) costs two here (one many-"target" problem on the FP64 tensor pipe)."""
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        critical(xs.ndim == 2 and xs.shape[1] == self.size(), "function: invalid input size, expecting (K, ",
                 self.size(), ")")
        K = xs.shape[0]
        l1s = np.ascontiguousarray(np.broadcast_to(np.asarray(l1s, dtype=np.float64), (K,)))
        l2s = np.ascontiguousarray(np.broadcast_to(np.asarray(l2s, dtype=np.float64), (K,)))
        gs = np.empty_like(xs) if want_grad else None
        fs = np.empty(K, dtype=np.float64)
        _lib.check(_lib.lib().nb200_vgrad_batch(self.handle, K, _ptr(xs), xs.shape[1], _ptr(l1s), _ptr(l2s), _ptr(gs),
                                                _ptr(fs)))
        self._fcalls += K
        self._gcalls += K if want_grad else 0
        return fs, gs

    # -- sample-sharded evaluation (one rank per GPU; see sharded.py) ------------------------------------------------
    def partial(self, x, want_grad: bool = True):
        """Phase 1: returns (device pointer, length) of the un-normalised [loss | gb | gW] sums."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        ptr = ctypes.c_void_p()
        length = ctypes.c_int64()
        _lib.check(_lib.lib().nb200_vgrad_partial(self.handle, _ptr(x), x.size, int(want_grad), ctypes.byref(ptr),
                                                  ctypes.byref(length)))
        return ptr.value, length.value

    def finish(self, n_global: int, gx=None) -> float:
        _lib.check(_lib.lib().nb200_vgrad_finish(self.handle, n_global, _ptr(gx), ctypes.byref(self._fx)))
        return self._fx.value

    # -- device-resident variants (x, gx, fx are raw device addresses; enqueue only) -----------------------------------
    def vgrad_device(self, x_ptr: int, gx_ptr: int | None, fx_ptr: int) -> None:
        _lib.check(_lib.lib().nb200_vgrad_device(self.handle, ctypes.c_void_p(x_ptr), self.size(),
                                                 ctypes.c_void_p(gx_ptr or 0), ctypes.c_void_p(fx_ptr)))

    def partial_device(self, x_ptr: int, want_grad: bool = True):
        ptr = ctypes.c_void_p()
        length = ctypes.c_int64()
        _lib.check(_lib.lib().nb200_vgrad_partial_device(self.handle, ctypes.c_void_p(x_ptr), self.size(),
                                                         int(want_grad), ctypes.byref(ptr), ctypes.byref(length)))
        return ptr.value, length.value

    def finish_device(self, x_ptr: int, n_global: int, gx_ptr: int | None, fx_ptr: int) -> None:
        _lib.check(_lib.lib().nb200_vgrad_finish_device(self.handle, ctypes.c_void_p(x_ptr), n_global,
                                                        ctypes.c_void_p(gx_ptr or 0), ctypes.c_void_p(fx_ptr)))

    # -- fused compute + collective (single-target kernel; peers' mailboxes mapped through CUDA IPC) -----------------
    def peer_export(self, world: int):
        """This rank's 64-byte IPC handle of its mailbox and the grid size (must agree across ranks)."""
        handle = ctypes.create_string_buffer(64)
        grid = ctypes.c_int64()
        _lib.check(_lib.lib().nb200_objective_peer_export(self.handle, world, handle, ctypes.byref(grid)))
        return handle.raw, grid.value

    def peer_connect(self, rank: int, world: int, handles: bytes) -> None:
        critical(len(handles) == 64 * world, "peers: expecting ", 64 * world, " bytes of handles")
        _lib.check(_lib.lib().nb200_objective_peer_connect(self.handle, rank, world, handles))

    def vgrad_sharded(self, x, n_global: int, gx=None) -> float:
        x = np.ascontiguousarray(x, dtype=np.float64)
        _lib.check(_lib.lib().nb200_vgrad_sharded(self.handle, _ptr(x), x.size, n_global, _ptr(gx),
                                                  ctypes.byref(self._fx)))
        return self._fx.value

    def vgrad_sharded_device(self, x_ptr: int, n_global: int, gx_ptr: int | None, fx_ptr: int) -> None:
        _lib.check(_lib.lib().nb200_vgrad_sharded_device(self.handle, ctypes.c_void_p(x_ptr), self.size(), n_global,
                                                         ctypes.c_void_p(gx_ptr or 0), ctypes.c_void_p(fx_ptr)))

    def peer_mode(self) -> int:
        """0: publish-only launches + stream wait; 1: the launch waits for its peers; -1: not connected."""
        return int(_lib.lib().nb200_objective_peer_mode(self.handle))

    def parts(self) -> int:
        """Devices this objective evaluates on (1 unless its dataset lives on a device group)."""
        return int(_lib.lib().nb200_objective_parts(self.handle))

    def part_stats(self, index: int, reset: bool = False):
        """(describe, total pass ms, passes, launches) of the objective on device `index` of the group."""
        part = ctypes.c_void_p(_lib.lib().nb200_objective_part(self.handle, index))
        critical(bool(part), "function: no such part")
        buffer = ctypes.create_string_buffer(512)
        _lib.check(_lib.lib().nb200_objective_describe(part, buffer, 512))
        total, count = ctypes.c_double(), ctypes.c_int64()
        _lib.check(_lib.lib().nb200_objective_pass_stats(part, ctypes.byref(total), ctypes.byref(count), int(reset)))
        return buffer.value.decode(), total.value, count.value, int(_lib.lib().nb200_objective_launches(part))

    def sync(self) -> None:
        _lib.check(_lib.lib().nb200_objective_sync(self.handle))

    def stream(self) -> int:
        return int(_lib.lib().nb200_objective_stream(self.handle) or 0)

    def set_stream(self, stream: int | None) -> None:
        """Order every launch of this objective on the given cudaStream_t (0 = the legacy default stream);
        None goes back to the objective's own stream."""
        _lib.check(_lib.lib().nb200_objective_set_stream(self.handle, ctypes.c_void_p(stream or 0),
                                                         int(stream is None)))

    # -- introspection -------------------------------------------------------------------------------------------------
    def launches(self) -> int:
        return int(_lib.lib().nb200_objective_launches(self.handle))

    def enable_timing(self, enabled: bool = True) -> None:
        _lib.check(_lib.lib().nb200_objective_enable_timing(self.handle, int(enabled)))

    def last_pass_ms(self) -> float:
        return float(_lib.lib().nb200_objective_last_pass_ms(self.handle))

    def pass_stats(self, reset: bool = False):
        total = ctypes.c_double()
        count = ctypes.c_int64()
        _lib.check(_lib.lib().nb200_objective_pass_stats(self.handle, ctypes.byref(total), ctypes.byref(count),
                                                         int(reset)))
        return total.value, count.value

    def pass_samples(self, part: int = 0):
        """Milliseconds of every timed pass since the last reset (of device `part` of a group)."""
        handle = ctypes.c_void_p(_lib.lib().nb200_objective_part(self.handle, part))
        critical(bool(handle), "function: no such part")
        count = ctypes.c_int64()
        _lib.check(_lib.lib().nb200_objective_pass_samples(handle, None, 0, ctypes.byref(count)))
        out = np.empty(count.value, dtype=np.float64)
        _lib.check(_lib.lib().nb200_objective_pass_samples(handle, _ptr(out), out.size, ctypes.byref(count)))
        return out

    def host_profile(self, reset: bool = False):
        """Host-side wall clock of the host-buffer call since the last reset: (stage_us, enqueue_us, wait_us, calls)."""
        a, b, c, calls = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        _lib.check(_lib.lib().nb200_objective_host_profile(self.handle, ctypes.byref(a), ctypes.byref(b),
                                                           ctypes.byref(c), ctypes.byref(calls), int(reset)))
        return a.value, b.value, c.value, calls.value

    def set_variant(self, variant: int) -> None:
        _lib.check(_lib.lib().nb200_objective_set_variant(self.handle, variant))

    def describe(self) -> str:
        buffer = ctypes.create_string_buffer(512)
        _lib.check(_lib.lib().nb200_objective_describe(self.handle, buffer, 512))
        return buffer.value.decode()

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib().nb200_objective_release(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LinearFunction(_DeviceFunction):
    """nano::linear::function_t (src/linear/function.cpp:21-168): the ERM objective of a linear model.

    x = [W (T x D row-major) | b (T)]; f = mean loss + l1 mean|W| + l2/2 mean(W^2)."""

    def __init__(self, dataset: Dataset, loss: Loss, l1reg: float = 0.0, l2reg: float = 0.0, _clone_of=None):
        self._loss = loss
        self._l1reg = float(l1reg)
        self._l2reg = float(l2reg)
        super().__init__("linear", dataset, loss.kind, loss.param, FLAVOUR_LINEAR, self._l1reg, self._l2reg,
                         _clone_of=_clone_of)
        self._isize = dataset.inputs
        self._tsize = dataset.targets
        # src/linear/function.cpp:34-36
        self._convex = loss.convex()
        self._smooth = loss.smooth() and self._l1reg <= 0.0
        self._strong_convexity = self._l2reg / float(self._isize * self._tsize)

    def weights(self, x):
        """include/nano/linear/function.h:30-43"""
        return np.asarray(x)[: self._isize * self._tsize].reshape(self._tsize, self._isize)

    def bias(self, x):
        """include/nano/linear/function.h:45-59"""
        return np.asarray(x)[self._isize * self._tsize:]

    def clone(self) -> "LinearFunction":
        other = LinearFunction(self._dataset, self._loss, self._l1reg, self._l2reg, _clone_of=self.handle)
        other._fcalls, other._gcalls, other._hcalls = self._fcalls, self._gcalls, self._hcalls
        return other


_SYNTH_LOSSES = {
    # basename: (device kind, convex, smooth, regression)  — src/function/mlearn/loss.h:11-75
    "mse": (0, True, True, True),
    "mae": (1, True, False, True),
    "cauchy": (KIND_SYNTH_CAUCHY, False, False, True),
    "hinge": (3, True, False, False),
    "logistic": (KIND_SYNTH_LOGISTIC, True, True, False),
}


class SyntheticFunction(_DeviceFunction):
    """function_{ridge,lasso,elasticnet}_t<loss> (src/function/mlearn/): the builtin synthetic linear-model test
    functions (`mse+ridge`, `hinge+lasso`, `logistic+elasticnet`, ... : 5 losses x 3 regularisers).

    Data as linear_model_t's constructor makes it (linear.cpp:5-45); x = w (D,), the bias is fixed."""

    def __init__(self, ctx: Context, loss: str = "mse", reg: str = "ridge", dims: int = 10, seed: int = 42,
                 alpha1: float = 1.0, alpha2: float = 1.0, sratio: float = 10.0, modulo: int = 1, _clone_of=None):
        critical(loss in _SYNTH_LOSSES, "function: unknown synthetic loss <", loss, ">")
        critical(reg in ("ridge", "lasso", "elasticnet"), "function: unknown regulariser <", reg, ">")
        # parameter ranges: ridge.cpp:35-38, lasso.cpp:35-38, elasticnet.cpp:36-40
        critical(0 <= seed <= 10000, "function::seed out of [0, 10000]")
        critical(0.0 <= alpha1 <= 1e8 and 0.0 <= alpha2 <= 1e8, "function: alpha out of [0, 1e+8]")
        critical(0.1 <= sratio <= 1e3, "function: sratio out of [0.1, 1e+3]")
        critical(1 <= modulo <= 100, "function: modulo out of [1, 100]")
        kind, convex, smooth, regression = _SYNTH_LOSSES[loss]
        self._ctx, self._loss_name, self._reg = ctx, loss, reg
        self._dims, self._seed, self._sratio, self._modulo = int(dims), int(seed), float(sratio), int(modulo)
        self._alpha1 = float(alpha1) if reg in ("lasso", "elasticnet") else 0.0
        self._alpha2 = float(alpha2) if reg in ("ridge", "elasticnet") else 0.0

        if _clone_of is None:
            D = max(self._dims, 2)                                      # make_inputs (ridge.cpp:13-16)
            N = int(max(self._sratio * float(self._dims), 10.0))        # make_samples (ridge.cpp:23-26)
            X = np.empty((N, D), dtype=np.float64)
            wopt = np.empty(D, dtype=np.float64)
            bopt = ctypes.c_double()
            _lib.check(_lib.lib().nb200_synth_linear_inputs(N, D, self._seed, self._modulo, _ptr(X), _ptr(wopt),
                                                            ctypes.byref(bopt)))
            # eval_outputs(m_woptimum) on the device, then the targets (linear.cpp:35-44)
            staging = Dataset.pin(ctx, X, np.zeros((N, 1)))
            outputs = staging.predict(wopt.reshape(1, D), np.array([bopt.value]))
            staging.close()
            Y = outputs if regression else np.sign((outputs - bopt.value) - 0.5)
            dataset = Dataset.pin(ctx, X, Y)
            self._woptimum, self._boptimum = wopt, bopt.value
            fixed_bias = np.array([bopt.value])
        else:
            dataset, fixed_bias = _clone_of._dataset, None
            self._woptimum, self._boptimum = _clone_of._woptimum, _clone_of._boptimum

        super().__init__(f"{loss}+{reg}", dataset, kind, 0.0, FLAVOUR_SYNTH, self._alpha1, self._alpha2,
                         fixed_bias=fixed_bias, _clone_of=None if _clone_of is None else _clone_of.handle)
        self._convex = convex
        if reg == "ridge":        # ridge.cpp:40-42
            self._smooth, self._strong_convexity = smooth, self._alpha2
        elif reg == "lasso":      # lasso.cpp:40-42
            self._smooth, self._strong_convexity = False, 0.0
        else:                     # elasticnet.cpp:42-44
            self._smooth, self._strong_convexity = (self._alpha1 == 0.0 and smooth), self._alpha2

    def do_name(self) -> str:
        parts = []
        if self._reg in ("lasso", "elasticnet"):
            parts.append(f"alpha1={self._alpha1:g}")
        if self._reg in ("ridge", "elasticnet"):
            parts.append(f"alpha2={self._alpha2:g}")
        parts += [f"sratio={self._sratio:g}", f"modulo={self._modulo}", f"seed={self._seed}"]
        return f"{self._type_id}[{','.join(parts)}]"

    def clone(self) -> "SyntheticFunction":
        other = SyntheticFunction(self._ctx, self._loss_name, self._reg, self._dims, self._seed,
                                  self._alpha1 if self._alpha1 else 1.0, self._alpha2 if self._alpha2 else 1.0,
                                  self._sratio, self._modulo, _clone_of=self)
        other._alpha1, other._alpha2 = self._alpha1, self._alpha2
        other._fcalls, other._gcalls, other._hcalls = self._fcalls, self._gcalls, self._hcalls
        return other

    def make(self, dims: int) -> "SyntheticFunction":
        """function_t::make(dims): same parameters, another dimension (ridge.cpp:82-91)."""
        return SyntheticFunction(self._ctx, self._loss_name, self._reg, dims, self._seed,
                                 self._alpha1 if self._reg != "ridge" else 1.0,
                                 self._alpha2 if self._reg != "lasso" else 1.0, self._sratio, self._modulo)

    @property
    def dataset(self) -> Dataset:
        return self._dataset


__all__ = [
    "Context", "Dataset", "Function", "LinearFunction", "SyntheticFunction", "Loss", "critical",
    "FLAVOUR_LINEAR", "FLAVOUR_SYNTH", "SYNTH_CLASSIFICATION", "SYNTH_REGRESSION", "SYNTH_MULTICLASS",
]
