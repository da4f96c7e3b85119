#!/usr/bin/env python
"""Multi-target one-pass / two-pass kernels: vgrad time per shape (2 GiB of X by default), burst (5 warm-up + 20 timed
launches after 1.5 s idle) and steady state (second half of --seconds of back-to-back evaluations).

    python tools/sweep_multi.py [--shapes "D,T,loss[,variant];..."] [--gib 2] [--seconds 1.5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

DEFAULT = "784,10,s-classnll;1024,8,s-classnll;512,16,s-classnll;256,10,s-classnll;100,10,s-classnll;1000,5,s-classnll;" \
          "1024,2,s-classnll;1024,4,s-classnll;512,4,s-classnll;2048,100,s-classnll"


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--shapes", default=DEFAULT)
    parser.add_argument("--gib", type=float, default=2.0)
    parser.add_argument("--seconds", type=float, default=1.5)
    args = parser.parse_args()

    import torch

    import libnano_b200 as nb

    ctx = nb.Context(0)
    for spec in args.shapes.split(";"):
        cells = spec.split(",")
        D, T, loss = int(cells[0]), int(cells[1]), cells[2]
        variant = int(cells[3]) if len(cells) > 3 else -1
        N = int(args.gib * 2 ** 30 / (8 * D))
        kind = nb.SYNTH_MULTICLASS if T > 1 else nb.SYNTH_CLASSIFICATION
        data = nb.Dataset.synth(ctx, kind, N, D, T=T, seed=42)
        function = nb.LinearFunction(data, nb.Loss(loss), 0.0, 1e-2)
        if variant != -1:
            function.set_variant(variant)
        function.set_stream(torch.cuda.current_stream().cuda_stream)
        n = function.size()
        x_dev = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, n) / np.sqrt(D)).cuda()
        g_dev = torch.empty(n, dtype=torch.float64, device="cuda")
        f_dev = torch.empty(1, dtype=torch.float64, device="cuda")

        def step():
            function.vgrad_device(x_dev.data_ptr(), g_dev.data_ptr(), f_dev.data_ptr())

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        time.sleep(1.5)
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            step()
        e1.record()
        torch.cuda.synchronize()
        burst_ms = e0.elapsed_time(e1) / 20
        batches, t0 = [], time.perf_counter()
        while time.perf_counter() - t0 < args.seconds:
            e0.record()
            for _ in range(10):
                step()
            e1.record()
            torch.cuda.synchronize()
            batches.append(e0.elapsed_time(e1) / 10)
        steady_ms = float(np.mean(batches[len(batches) // 2:]))
        x_bytes = 8.0 * N * D
        flops = 4.0 * N * D * T
        print(json.dumps({"D": D, "T": T, "rows": N, "loss": loss, "kernel": function.describe(),
                          "burst_ms": round(burst_ms, 4), "steady_ms": round(steady_ms, 4),
                          "x_once_gbs_burst": round(x_bytes / burst_ms / 1e6, 1),
                          "x_once_gbs_steady": round(x_bytes / steady_ms / 1e6, 1),
                          "tflops_burst": round(flops / burst_ms / 1e9, 2)}), flush=True)
        function.close()
        data.close()
        del x_dev, g_dev, f_dev
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
