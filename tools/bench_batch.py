#!/usr/bin/env python
"""N3: K models in one pass over X versus K single-model evaluations (2^20 x 1024 logistic)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import libnano_b200 as nb

    N, D = 1 << 20, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    ctx = nb.Context(0)
    data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, N, D, seed=42)
    function = nb.LinearFunction(data, nb.Loss("s-logistic"), 0.0, 1e-2)
    rng = np.random.default_rng(0)
    g1 = np.empty(D + 1)
    for K in (2, 3, 4, 6, 8, 12, 16, 32, 64):
        xs = rng.uniform(-1, 1, (K, D + 1)) / np.sqrt(D)
        l2s = np.logspace(-6, 1, K)
        function.vgrad_batch(xs, 0.0, l2s)
        t0 = time.perf_counter()
        for _ in range(5):
            function.vgrad_batch(xs, 0.0, l2s)
        batch_ms = (time.perf_counter() - t0) / 5 * 1e3
        for k in range(K):
            function(xs[k], g1)
        t0 = time.perf_counter()
        for _ in range(2):
            for k in range(K):
                function(xs[k], g1)
        single_ms = (time.perf_counter() - t0) / 2 * 1e3
        print(json.dumps({"K": K, "D": D, "batch_ms": round(batch_ms, 3), "K_single_calls_ms": round(single_ms, 3),
                          "speedup": round(single_ms / batch_ms, 2),
                          "batch_tflops": round(4.0 * N * D * (K + K % 2) / batch_ms / 1e9, 2),
                          "evaluations_per_s": round(K / batch_ms * 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()
