#!/usr/bin/env python
"""Probe: kernel time, SM clock and board power of the fused kernel in alternating blocks of (a) back-to-back launches
and (b) one launch per host synchronisation, to tell a launch-onto-idle effect from power / thermal state.

    python tools/probe_power.py [--rows 1048576] [--features 1024] [--block-calls 150] [--blocks 6]
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

QUERY = "clocks.sm,clocks.mem,power.draw.instant,power.draw.average,temperature.gpu,temperature.memory," \
        "clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown"


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--rows", type=int, default=1 << 20)
    parser.add_argument("--features", type=int, default=1024)
    parser.add_argument("--block-calls", type=int, default=150)
    parser.add_argument("--blocks", type=int, default=6)
    parser.add_argument("--loss", default="s-logistic")
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
    data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, args.rows, args.features, seed=42, noise=0.1)
    function = nb.LinearFunction(data, nb.Loss(args.loss), 0.0, 1e-2)
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
    time.sleep(1.0)

    def summarise(t0, t1):
        cells = [l.split(",") for (t, l) in samples if t0 <= t <= t1]
        def col(i):
            out = []
            for c in cells:
                try:
                    out.append(float(c[i]))
                except (ValueError, IndexError):
                    pass
            return out
        sm, power, tg, tm = col(0), col(2), col(4), col(5)
        capped = sum(1 for c in cells if len(c) > 6 and c[6].strip().lower() == "active")
        return {"sm_mhz_min": min(sm) if sm else None, "sm_mhz_med": float(np.median(sm)) if sm else None,
                "power_w_med": float(np.median(power)) if power else None, "power_w_max": max(power) if power else None,
                "temp_gpu": max(tg) if tg else None, "temp_mem": max(tm) if tm else None,
                "samples": len(cells), "power_capped_samples": capped}

    for block in range(args.blocks):
        for mode in ("back-to-back", "sync per call"):
            function.pass_stats(reset=True)
            t0 = time.perf_counter()
            if mode == "back-to-back":
                for _ in range(args.block_calls):
                    launch()
                function.sync()
            else:
                for _ in range(args.block_calls):
                    launch()
                    function.sync()
            t1 = time.perf_counter()
            total_ms, count = function.pass_stats(reset=True)
            line = {"block": block, "mode": mode, "kernel_us": round(1e3 * total_ms / max(count, 1), 1),
                    "wall_us_per_call": round(1e6 * (t1 - t0) / args.block_calls, 1), "loss": args.loss}
            line.update(summarise(t0, t1))
            print(json.dumps(line), flush=True)
    proc.terminate()


if __name__ == "__main__":
    main()
