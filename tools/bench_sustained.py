#!/usr/bin/env python
"""Sustained throughput of the fused single-target kernel: back-to-back evaluations for --seconds per configuration,
reporting the kernel time of the first batch (cool GPU, boost clocks) and of the steady state (second half of the
run; on this pool the board sits at its 1000 W power cap while streaming from HBM), with SM clock and power.

    python tools/bench_sustained.py [--seconds 3] [--configs "rows,features,loss,variant;..."]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

QUERY = "clocks.sm,power.draw.instant,temperature.gpu,clocks_event_reasons.sw_power_cap"
DEFAULT = "1048576,1024,s-logistic,-1;1048576,1024,s-logistic,21;1048576,1024,mse,-1;2097152,512,s-logistic,-1;" \
          "2097152,512,s-hinge,-1;4194304,256,s-logistic,-1;8388608,128,s-logistic,-1;524288,2048,s-logistic,-1"


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--seconds", type=float, default=3.0)
    parser.add_argument("--configs", default=DEFAULT)
    parser.add_argument("--value-only", action="store_true")
    args = parser.parse_args()

    import torch

    import libnano_b200 as nb

    samples = []
    proc = subprocess.Popen(["nvidia-smi", "--id=0", f"--query-gpu={QUERY}", "--format=csv,noheader,nounits", "-lms", "20"],
                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

    def pump():
        for line in proc.stdout:
            samples.append((time.perf_counter(), line.strip()))

    threading.Thread(target=pump, daemon=True).start()
    ctx = nb.Context(0)

    for spec in args.configs.split(";"):
        rows, features, loss, variant = spec.split(",")
        rows, features, variant = int(rows), int(features), int(variant)
        data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, rows, features, seed=42, noise=0.1)
        function = nb.LinearFunction(data, nb.Loss(loss), 0.0, 1e-2)
        if variant >= 0:
            function.set_variant(variant)
        n = function.size()
        x = np.random.default_rng(0).uniform(-1, 1, n) / np.sqrt(features)
        x_dev = torch.from_numpy(x).cuda()
        g_dev = torch.empty(n, dtype=torch.float64, device="cuda")
        f_dev = torch.empty(1, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        g_ptr = None if args.value_only else g_dev.data_ptr()
        for _ in range(3):
            function.vgrad_device(x_dev.data_ptr(), g_ptr, f_dev.data_ptr())
        function.sync()
        time.sleep(1.5)  # cool-down: every configuration starts from a similar state
        function.enable_timing(True)
        function.pass_stats(reset=True)
        # burst exactly as bench.py times it: 5 warm-up launches after the idle period, then 20 timed ones
        for _ in range(5):
            function.vgrad_device(x_dev.data_ptr(), g_ptr, f_dev.data_ptr())
        function.sync()
        function.pass_stats(reset=True)
        for _ in range(20):
            function.vgrad_device(x_dev.data_ptr(), g_ptr, f_dev.data_ptr())
        function.sync()
        burst = function.pass_samples()
        function.pass_stats(reset=True)
        time.sleep(1.5)
        batches = []
        t_start = time.perf_counter()
        while time.perf_counter() - t_start < args.seconds:
            for _ in range(25):
                function.vgrad_device(x_dev.data_ptr(), g_ptr, f_dev.data_ptr())
            function.sync()
            total_ms, count = function.pass_stats(reset=True)
            batches.append((time.perf_counter(), 1e3 * total_ms / max(count, 1)))
        t_end = time.perf_counter()
        half = len(batches) // 2
        steady = float(np.mean([b[1] for b in batches[half:]]))
        t_half = batches[half][0] if half < len(batches) else t_start
        cells = [l.split(",") for (t, l) in samples if t_half <= t <= t_end]
        sm = [float(c[0]) for c in cells if len(c) >= 2]
        pw = [float(c[1]) for c in cells if len(c) >= 2]
        bytes_alg = 8 * rows * features + 8 * rows + 16 * (features + 1)
        print(json.dumps({
            "rows": rows, "features": features, "loss": loss, "value_only": bool(args.value_only),
            "kernel": function.describe(), "burst_us_median": round(1e3 * float(np.median(burst)), 1),
            "burst_us_mean": round(1e3 * float(np.mean(burst)), 1),
            "burst_gbs": round(bytes_alg / float(np.mean(burst)) / 1e6, 1),
            "first_batch_us": round(batches[0][1], 1), "steady_us": round(steady, 1),
            "batches_us": [round(b[1], 1) for b in batches[:6]],
            "first_gbs": round(bytes_alg / batches[0][1] / 1e3, 1), "steady_gbs": round(bytes_alg / steady / 1e3, 1),
            "steady_sm_mhz": float(np.median(sm)) if sm else None, "steady_power_w": float(np.median(pw)) if pw else None,
            "launches": 25 * len(batches)}), flush=True)
        function.enable_timing(False)
        del function, data, x_dev, g_dev, f_dev
        torch.cuda.empty_cache()
    proc.terminate()


if __name__ == "__main__":
    main()
