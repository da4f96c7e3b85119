"""ctypes binding of libnb200.so (include/nb200.h). Loading fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnb200.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_i64 = ctypes.c_int64
c_void_pp = ctypes.POINTER(ctypes.c_void_p)

# every symbol include/nb200.h declares: (restype, argtypes)
SIGNATURES = {
    "nb200_last_error": (ctypes.c_char_p, []),
    "nb200_version": (ctypes.c_char_p, []),
    "nb200_device_count": (ctypes.c_int, []),
    "nb200_init": (ctypes.c_int, [ctypes.c_int, c_void_pp]),
    "nb200_init_multi": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_int), c_void_pp]),
    "nb200_ctx_devices": (ctypes.c_int, [ctypes.c_void_p]),
    "nb200_ctx_release": (None, [ctypes.c_void_p]),
    "nb200_dataset_pin": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, c_i64, c_i64, c_i64, c_void_pp]),
    "nb200_dataset_synth": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, c_i64, c_i64, c_i64, c_i64,
                                           ctypes.c_double, c_void_pp]),
    "nb200_dataset_read": (ctypes.c_int, [ctypes.c_void_p, c_i64, c_i64, c_double_p, c_double_p]),
    "nb200_dataset_scale": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p,
                                           ctypes.c_char_p]),
    "nb200_dataset_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int] + [c_double_p] * 7),
    "nb200_upscale": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p]),
    "nb200_dataset_samples": (c_i64, [ctypes.c_void_p]),
    "nb200_dataset_inputs": (c_i64, [ctypes.c_void_p]),
    "nb200_dataset_targets": (c_i64, [ctypes.c_void_p]),
    "nb200_dataset_retain": (None, [ctypes.c_void_p]),
    "nb200_dataset_release": (None, [ctypes.c_void_p]),
    "nb200_objective_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                              ctypes.c_double, ctypes.c_double, c_double_p, c_void_pp]),
    "nb200_objective_clone": (ctypes.c_int, [ctypes.c_void_p, c_void_pp]),
    "nb200_objective_release": (None, [ctypes.c_void_p]),
    "nb200_objective_size": (c_i64, [ctypes.c_void_p]),
    "nb200_vgrad": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_i64, c_double_p, c_double_p]),
    "nb200_vgrad_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_i64, ctypes.c_void_p, ctypes.c_void_p]),
    "nb200_objective_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "nb200_vgrad_partial": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_i64, ctypes.c_int, c_void_pp,
                                           ctypes.POINTER(c_i64)]),
    "nb200_vgrad_finish": (ctypes.c_int, [ctypes.c_void_p, c_i64, c_double_p, c_double_p]),
    "nb200_vgrad_partial_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_i64, ctypes.c_int, c_void_pp,
                                                  ctypes.POINTER(c_i64)]),
    "nb200_vgrad_finish_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_i64, ctypes.c_void_p,
                                                 ctypes.c_void_p]),
    "nb200_objective_peer_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                                   ctypes.POINTER(c_i64)]),
    "nb200_objective_peer_connect": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "nb200_vgrad_sharded": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_i64, c_i64, c_double_p, c_double_p]),
    "nb200_vgrad_sharded_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_i64, c_i64, ctypes.c_void_p,
                                                  ctypes.c_void_p]),
    "nb200_objective_peer_mode": (ctypes.c_int, [ctypes.c_void_p]),
    "nb200_objective_parts": (ctypes.c_int, [ctypes.c_void_p]),
    "nb200_objective_part": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int]),
    "nb200_objective_stream": (ctypes.c_void_p, [ctypes.c_void_p]),
    "nb200_objective_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "nb200_objective_launches": (c_i64, [ctypes.c_void_p]),
    "nb200_objective_enable_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "nb200_objective_last_pass_ms": (ctypes.c_double, [ctypes.c_void_p]),
    "nb200_objective_pass_stats": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.POINTER(c_i64), ctypes.c_int]),
    "nb200_objective_pass_samples": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_i64, ctypes.POINTER(c_i64)]),
    "nb200_probe_peak": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p]),
    "nb200_objective_host_profile": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_double_p, c_double_p,
                                                    ctypes.POINTER(c_i64), ctypes.c_int]),
    "nb200_objective_set_variant": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "nb200_objective_describe": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, c_i64]),
    "nb200_vgrad_batch": (ctypes.c_int, [ctypes.c_void_p, c_i64, c_double_p, c_i64, c_double_p, c_double_p, c_double_p,
                                         c_double_p]),
    "nb200_lbfgs_minimize": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_i64, ctypes.c_double, c_i64, c_i64, c_i64,
                                            ctypes.c_void_p, ctypes.c_void_p, c_double_p, c_double_p,
                                            ctypes.POINTER(c_i64), ctypes.POINTER(ctypes.c_int)]),
    "nb200_cgd_minimize": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_i64, ctypes.c_int, ctypes.c_double, c_i64,
                                          ctypes.c_double, ctypes.c_double, c_i64, ctypes.c_void_p, ctypes.c_void_p,
                                          c_double_p, c_double_p, ctypes.POINTER(c_i64), ctypes.POINTER(ctypes.c_int)]),
    "nb200_osga_minimize": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_i64, ctypes.c_double, c_i64, c_i64,
                                           ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                           ctypes.c_double, c_i64, c_double_p, c_double_p, ctypes.POINTER(c_i64),
                                           ctypes.POINTER(ctypes.c_int)]),
    "nb200_gboost_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, c_double_p, c_i64,
                                           c_i64, ctypes.POINTER(ctypes.c_int32), c_i64, c_double_p, c_double_p,
                                           c_void_pp]),
    "nb200_gboost_size": (c_i64, [ctypes.c_void_p]),
    "nb200_gboost_launches": (c_i64, [ctypes.c_void_p]),
    "nb200_gboost_vgrad": (ctypes.c_int, [ctypes.c_void_p, c_double_p, c_i64, c_double_p, c_double_p]),
    "nb200_gboost_release": (None, [ctypes.c_void_p]),
    "nb200_predict": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_double_p, c_double_p, c_double_p]),
    "nb200_loss_value": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, c_double_p, c_double_p, c_i64,
                                        c_i64, c_double_p]),
    "nb200_loss_error": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p, c_double_p, c_i64, c_i64,
                                        c_double_p]),
    "nb200_evaluate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                      c_double_p, c_double_p, c_double_p, c_double_p]),
    "nb200_simplex_qp": (ctypes.c_int, [c_double_p, c_double_p, c_i64, c_double_p]),
    "nb200_synth_linear_inputs": (ctypes.c_int, [c_i64, c_i64, ctypes.c_uint64, c_i64, c_double_p, c_double_p,
                                                 c_double_p]),
    "nb200_loss_vgrad": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, c_double_p, c_double_p, c_i64,
                                        c_i64, c_double_p]),
}

ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, c_i64, ctypes.c_void_p)

_lib = None


class Nb200Error(RuntimeError):
    """Raised for any non-zero status of the C ABI; mirrors nano::critical's std::runtime_error
    (include/nano/critical.h:13-29)."""

    def __init__(self, status: int, message: str):
        super().__init__(message or f"libnb200 failed with status {status}")
        self.status = status


def lib():
    """Load libnb200.so (built by libnano_b200.build). Raises if it is missing: the product path never falls back
    to a CPU implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Nb200Error(-1, f"{LIB_PATH} is missing: run `python -m libnano_b200.build` (needs nvcc); "
                                 "libnano_b200 has no CPU fallback")
        override = os.environ.get("NB200_LIBRARY")  # A/B runs of two builds on one GPU box (tools/): never a fallback
        L = ctypes.CDLL(override or LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            if override and not hasattr(L, name):
                continue  # an older build under test may predate a symbol
            fn = getattr(L, name)  # AttributeError here means the library does not match include/nb200.h
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise Nb200Error(status, lib().nb200_last_error().decode("utf-8", "replace"))
