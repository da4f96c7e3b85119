#!/usr/bin/env python
"""Probe: does the fused kernel run slower when it is launched onto an IDLE GPU (after a host synchronisation and a
gap) than back to back? Reports the CUDA-event kernel time and the wall time per call for a range of host gaps.

    python tools/probe_gap.py [--rows 1048576] [--features 1024] [--calls 100] [--torch-stream]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--rows", type=int, default=1 << 20)
    parser.add_argument("--features", type=int, default=1024)
    parser.add_argument("--calls", type=int, default=100)
    parser.add_argument("--torch-stream", action="store_true")
    parser.add_argument("--loss", default="s-logistic")
    args = parser.parse_args()

    import torch

    import libnano_b200 as nb

    ctx = nb.Context(0)
    data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, args.rows, args.features, seed=42, noise=0.1)
    function = nb.LinearFunction(data, nb.Loss(args.loss), 0.0, 1e-2)
    if args.torch_stream:
        function.set_stream(torch.cuda.current_stream().cuda_stream)
    n = function.size()
    x = np.random.default_rng(0).uniform(-1, 1, n) / np.sqrt(args.features)
    x_dev = torch.from_numpy(x).cuda()
    g_dev = torch.empty(n, dtype=torch.float64, device="cuda")
    f_dev = torch.empty(1, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    function.enable_timing(True)

    def launch():
        function.vgrad_device(x_dev.data_ptr(), g_dev.data_ptr(), f_dev.data_ptr())

    for _ in range(10):
        launch()
    function.sync()

    def report(label, wall_us, extra=None):
        total_ms, count = function.pass_stats(reset=True)
        line = {"what": label, "wall_us_per_call": round(wall_us, 1), "kernel_us": round(1e3 * total_ms / max(count, 1), 1),
                "torch_stream": bool(args.torch_stream), "plain_launch": os.environ.get("NB200_PLAIN_LAUNCH", "0"),
                "rows": args.rows, "features": args.features, "loss": args.loss}
        line.update(extra or {})
        print(json.dumps(line), flush=True)

    for depth in (args.calls, 2, 1):  # launches queued between two synchronisations
        function.pass_stats(reset=True)
        t0 = time.perf_counter()
        done = 0
        while done < args.calls:
            for _ in range(depth):
                launch()
            function.sync()
            done += depth
        report(f"queue depth {depth}", 1e6 * (time.perf_counter() - t0) / done)

    for gap_us in (0, 5, 10, 20, 40, 80, 200, 1000, 5000):
        function.pass_stats(reset=True)
        busy = 0.0
        t0 = time.perf_counter()
        for _ in range(args.calls):
            launch()
            function.sync()
            t1 = time.perf_counter()
            while (time.perf_counter() - t1) * 1e6 < gap_us:
                pass
            busy += time.perf_counter() - t1
        report(f"sync per call + {gap_us} us host gap", 1e6 * (time.perf_counter() - t0 - busy) / args.calls)


if __name__ == "__main__":
    main()
