#!/usr/bin/env python
"""Generates tests/golden/ref_loss_golden.json from the REFERENCE'S OWN loss code.

oracle/_ref/libref_loss.so (make -C oracle ref) is /root/reference/include/nano/loss/{mse,mae,cauchy,hinge,squared_hinge,
classnll,savage,tangent,logistic,exponential}.h compiled unmodified, in place, over the stand-in for the slice of Eigen's
array API they use (oracle/ref_shim/nano/tensor/eigen.h; Eigen3 itself is not installed in this image). This script runs
only where /root/reference exists; the JSON it writes is committed, and tests/ compare the oracle (CPU) and the CUDA
loss kernels (GPU) with it. Numbers are stored as float.hex() strings: bit-exact round trip.

    make -C oracle ref && python tests/golden/make_ref_golden.py
"""
import ctypes
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
KINDS = {"mse": 0, "mae": 1, "cauchy": 2, "hinge": 3, "squared-hinge": 4, "classnll": 5, "savage": 6, "tangent": 7,
         "logistic": 8, "exponential": 9}
CLASSIFICATION = {"hinge", "squared-hinge", "classnll", "savage", "tangent", "logistic", "exponential"}


def main():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_loss.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    for fn in (lib.ref_loss_value, lib.ref_loss_vgrad):
        fn.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int64, ctypes.c_int64, dp]
    rng = np.random.default_rng(20261020)
    cases = []
    for name, kind in KINDS.items():
        for T in (1, 3, 10):
            N = 24
            outputs = rng.normal(0.0, 2.0, (N, T))
            outputs[0] *= 10.0  # large margins: exp / log1p branches of logistic, saturation of savage / tangent
            outputs[1] *= 0.01
            if name in CLASSIFICATION:
                if T == 1:
                    targets = np.where(rng.uniform(-1, 1, (N, T)) < 0, -1.0, 1.0)
                else:
                    targets = -np.ones((N, T))
                    targets[np.arange(N), rng.integers(0, T, N)] = 1.0  # one-hot +-1 (include/nano/loss/class.h:51-58)
                if name == "exponential":
                    outputs = np.clip(outputs, -20.0, 20.0)
                # exactly on the hinge kink (t * o == 1) and exactly zero outputs
                outputs[2] = targets[2]
                outputs[3] = 0.0
            else:
                targets = rng.normal(0.0, 2.0, (N, T))
                outputs[2] = targets[2]  # zero residual: sign(0) == 0, |0|
                outputs[3] = targets[3] + 1e-300
            targets = np.ascontiguousarray(targets)
            outputs = np.ascontiguousarray(outputs)
            values = np.empty(N)
            vgrads = np.empty((N, T))
            assert lib.ref_loss_value(kind, targets.ctypes.data_as(dp), outputs.ctypes.data_as(dp), N, T,
                                      values.ctypes.data_as(dp)) == 0
            assert lib.ref_loss_vgrad(kind, targets.ctypes.data_as(dp), outputs.ctypes.data_as(dp), N, T,
                                      vgrads.ctypes.data_as(dp)) == 0
            assert np.all(np.isfinite(values)) and np.all(np.isfinite(vgrads)), (name, T)
            cases.append({"loss": name, "T": T, "N": N,
                          "targets": [float(v).hex() for v in targets.ravel()],
                          "outputs": [float(v).hex() for v in outputs.ravel()],
                          "values": [float(v).hex() for v in values],
                          "vgrads": [float(v).hex() for v in vgrads.ravel()]})
    # objective (B): src/function/mlearn/loss.cpp (fx is already divided by the sample count; T = 1)
    lib.ref_synth_loss.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int64, ctypes.c_int64, dp, dp]
    synth = []
    for name, kind in (("mse", 0), ("mae", 1), ("cauchy", 2), ("hinge", 3), ("logistic", 4)):
        N = 24
        outputs = rng.normal(0.0, 2.0, (N, 1))
        outputs[0] *= 10.0
        if name in ("hinge", "logistic"):
            targets = np.where(rng.uniform(-1, 1, (N, 1)) < 0, -1.0, 1.0)
            outputs[2] = targets[2]  # on the hinge kink
        else:
            targets = rng.normal(0.0, 2.0, (N, 1))
            outputs[2] = targets[2]  # zero residual
        fx = ctypes.c_double()
        gx = np.empty((N, 1))
        assert lib.ref_synth_loss(kind, outputs.ctypes.data_as(dp), targets.ctypes.data_as(dp), N, 1, ctypes.byref(fx),
                                  gx.ctypes.data_as(dp)) == 0
        assert np.isfinite(fx.value) and np.all(np.isfinite(gx))
        synth.append({"loss": name, "N": N, "outputs": [float(v).hex() for v in outputs.ravel()],
                      "targets": [float(v).hex() for v in targets.ravel()], "fx": float(fx.value).hex(),
                      "gx": [float(v).hex() for v in gx.ravel()]})

    # loss_t::error through the reference's error functors (include/nano/loss/error.h); its own stream, so the cases
    # above do not depend on this section
    lib.ref_loss_error.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int64, ctypes.c_int64, dp]
    erng = np.random.default_rng(20261021)
    errors = []
    for name, kind in (("absdiff", 0), ("sclass", 1), ("mclass", 2)):
        for T in (1, 3, 13):
            N = 32
            outputs = erng.normal(0.0, 1.0, (N, T))
            if name == "absdiff":
                targets = erng.normal(0.0, 1.0, (N, T))
                outputs[0] = targets[0]
            else:
                targets = np.where(erng.uniform(-1, 1, (N, T)) < 0, -1.0, 1.0)
                if name == "sclass" and T > 1:
                    targets = -np.ones((N, T))
                    targets[np.arange(N), erng.integers(0, T, N)] = 1.0
                outputs[0] = 0.0                                  # edges == 0 < epsilon; arg-max tie -> first label
                outputs[1] = targets[1] * 2.0 ** -52              # edges == epsilon exactly: not an error
                outputs[2] = targets[2] * 2.0 ** -53              # just below
                outputs[3] = outputs[3, 0]                        # all outputs equal (tie)
                outputs[4] = targets[4]
                outputs[5] = -targets[5]
            targets = np.ascontiguousarray(targets)
            outputs = np.ascontiguousarray(outputs)
            values = np.empty(N)
            assert lib.ref_loss_error(kind, targets.ctypes.data_as(dp), outputs.ctypes.data_as(dp), N, T,
                                      values.ctypes.data_as(dp)) == 0
            errors.append({"error": name, "T": T, "N": N, "targets": [float(v).hex() for v in targets.ravel()],
                           "outputs": [float(v).hex() for v in outputs.ravel()],
                           "errors": [float(v).hex() for v in values]})

    # the random stream of linear_model_t (seed 42 as in BASELINE config 1, and another seed)
    lib.ref_rng_draws.argtypes = [ctypes.c_uint64, ctypes.c_int64, dp]
    draws = {}
    for seed in (42, 7):
        out = np.empty(32)
        assert lib.ref_rng_draws(seed, out.size, out.ctypes.data_as(dp)) == 0
        draws[str(seed)] = [float(v).hex() for v in out]

    payload = {"rng_draws": draws, "synth": synth, "errors": errors, "source": "reference loss code compiled in place: include/nano/loss/*.h (cases) and "
                         "src/function/mlearn/loss.cpp (synth); oracle/ref_loss.cpp over oracle/ref_shim/ (stand-in for "
                         "the Eigen array / matrix vocabulary, sequential sums)",
               "generator": "tests/golden/make_ref_golden.py", "cases": cases}
    path = os.path.join(HERE, "ref_loss_golden.json")
    with open(path, "w") as handle:
        json.dump(payload, handle)
    print(path, len(cases), "cases +", len(synth), "synth cases +", len(errors), "error cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
