#!/usr/bin/env python
"""One process, a device group (nb200_init_multi): vgrad time of the single-target objective split over N devices,
through the host-buffer call and device-resident, beside the same rows on one device.

    python tools/bench_group.py [--devices N] [--rows-per-device R] [--features D]
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
    parser.add_argument("--devices", type=int, default=2)
    parser.add_argument("--rows-per-device", type=int, default=1 << 22)
    parser.add_argument("--features", type=int, default=1024)
    parser.add_argument("--steps", type=int, default=20)
    args = parser.parse_args()

    import libnano_b200 as nb

    D = args.features
    out = {"devices": args.devices, "rows_per_device": args.rows_per_device, "features": D}
    for name, devices in (("group", list(range(args.devices))), ("single", [0])):
        ctx = nb.Context(devices)
        N = args.rows_per_device * len(devices)
        data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, N, D, seed=42)
        function = nb.LinearFunction(data, nb.Loss("s-logistic"), 0.0, 1e-2)
        x = np.random.default_rng(0).uniform(-1, 1, D + 1) / np.sqrt(D)
        gx = np.empty_like(x)
        for _ in range(5):
            fx = function(x, gx)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fx = function(x, gx)
        ms = (time.perf_counter() - t0) / args.steps * 1e3
        out[name] = {"rows": N, "host_call_ms": round(ms, 4), "samples_features_per_s": N * D / (ms * 1e-3),
                     "f": fx, "plan": function.describe()}
        function.close()
        data.close()
        ctx.close()
    out["efficiency_vs_single"] = out["group"]["samples_features_per_s"] / (args.devices * out["single"]["samples_features_per_s"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
