#!/usr/bin/env python
"""BASELINE config 3: softmax cross-entropy linear model, N x 2048 x 100 fp64 (contraction-bound, FP64 tensor pipe).
Reports vgrad / value-only time, achieved TFLOP/s of the algorithmic flops (4*N*D*T resp. 2*N*D*T) and the cuBLAS
DGEMM yardstick for the same two contractions (torch.matmul fp64) on this GPU."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--rows", type=int, default=1 << 22)
    parser.add_argument("--features", type=int, default=2048)
    parser.add_argument("--targets", type=int, default=100)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--generic", action="store_true")
    parser.add_argument("--variant", type=int, default=-1, help="-1 heuristic, -3 generic, -4 tensor-pipe pair")
    parser.add_argument("--no-cublas", action="store_true")
    args = parser.parse_args()

    import libnano_b200 as nb

    N, D, T = args.rows, args.features, args.targets
    ctx = nb.Context(0)
    data = nb.Dataset.synth(ctx, nb.SYNTH_MULTICLASS, N, D, T=T, seed=42)
    function = nb.LinearFunction(data, nb.Loss("s-classnll"), 0.0, 1e-2)
    if args.generic:
        function.set_variant(-3)
    elif args.variant != -1:
        function.set_variant(args.variant)
    x = np.random.default_rng(0).uniform(-1, 1, function.size()) / np.sqrt(D)
    gx = np.empty_like(x)
    function(x, gx)
    function.enable_timing(True)
    result = {"rows": N, "features": D, "targets": T, "plan": function.describe()}
    for name, call, flops in (("vgrad", lambda: function(x, gx), 4.0 * N * D * T),
                              ("value", lambda: function(x), 2.0 * N * D * T)):
        times = []
        for _ in range(args.steps):
            fx = call()
            times.append(function.last_pass_ms())
        med = sorted(times)[len(times) // 2]
        result[name] = {"ms_med": round(med, 3), "ms_min": round(min(times), 3), "tflops": round(flops / med / 1e9, 2),
                        "x_gbs": round((2 if name == "vgrad" else 1) * 8.0 * N * D / med / 1e6, 1),
                        "x_once_gbs": round(8.0 * N * D / med / 1e6, 1), "f": fx}

    if not args.no_cublas:
        import torch

        rows = min(N, 1 << 20)
        A = torch.randn(rows, D, dtype=torch.float64, device="cuda")
        W = torch.randn(T, D, dtype=torch.float64, device="cuda")
        G = torch.randn(rows, T, dtype=torch.float64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for name, op, flops in (("cublas_forward", lambda: A @ W.t(), 2.0 * rows * D * T),
                                ("cublas_backward", lambda: G.t() @ A, 2.0 * rows * D * T)):
            op()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(5):
                e0.record()
                op()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            result[name] = {"rows": rows, "ms": round(best, 3), "tflops": round(flops / best / 1e9, 2)}
        S = 8192
        P, Q = torch.randn(S, S, dtype=torch.float64, device="cuda"), torch.randn(S, S, dtype=torch.float64, device="cuda")
        P @ Q
        torch.cuda.synchronize()
        e0.record()
        P @ Q
        e1.record()
        torch.cuda.synchronize()
        result["cublas_dgemm_8192"] = {"ms": round(e0.elapsed_time(e1), 3),
                                      "tflops": round(2.0 * S ** 3 / e0.elapsed_time(e1) / 1e9, 2)}
    print(json.dumps(result))


if __name__ == "__main__":
    main()
