#!/usr/bin/env python
"""BASELINE config 1: the builtin synthetic test function mse+ridge (objective B: src/function/mlearn/ridge.cpp) at
D = 100 (sratio 100 -> N = 10000) and at bench_function's default D = 1024 (N = 10240); seed 42, alpha2 = 1.
These fit in L2 and are launch-latency bound: reports microseconds per f(x, g) call through the host-buffer API (what
app/bench_function.cpp:47-49 measures with nano::measure: minimum over trials, x = 0) next to the kernel time.
The CPU figure of the same call is the `cpu_baseline`-style number produced by tests/bench with the oracle; here only
the GPU side is timed (tools/ never import oracle/)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import libnano_b200 as nb

    ctx = nb.Context(0)
    for dims, sratio in ((100, 100.0), (1024, 10.0)):
        for loss, reg in (("mse", "ridge"), ("logistic", "elasticnet"), ("hinge", "lasso")):
            function = nb.SyntheticFunction(ctx, loss, reg, dims, seed=42, alpha1=1.0, alpha2=1.0, sratio=sratio)
            x, gx = np.zeros(dims), np.empty(dims)
            for _ in range(20):
                function(x, gx)
            function.enable_timing(True)
            best_call, best_value, kernel = 1e9, 1e9, []
            for _ in range(16):  # nano::measure: minimum over 16 trials (include/nano/core/chrono.h:81-114)
                t0 = time.perf_counter()
                for _ in range(50):
                    function(x, gx)
                best_call = min(best_call, (time.perf_counter() - t0) / 50)
                kernel.append(function.last_pass_ms())
                t0 = time.perf_counter()
                for _ in range(50):
                    function(x)
                best_value = min(best_value, (time.perf_counter() - t0) / 50)
            X, _ = function.dataset.read(0, 1)
            print(json.dumps({"function": function.name(), "samples": function.dataset.samples, "dims": dims,
                              "f(x,g)_us": round(best_call * 1e6, 2), "f(x)_us": round(best_value * 1e6, 2),
                              "kernel_us": round(min(kernel) * 1e3, 2), "plan": function.describe()}), flush=True)


if __name__ == "__main__":
    main()
