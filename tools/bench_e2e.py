#!/usr/bin/env python
"""Where the host-buffer call spends its time: wall clock per f(x, gx) through each layer (ctypes -> nb200_vgrad,
LinearFunction.__call__, ShardedFunction.__call__), the host-side breakdown recorded inside nb200_vgrad
(stage / enqueue / wait) and the device-resident loop, for the bench shape and for a small L2-resident one.

    python tools/bench_e2e.py [--rows 1048576] [--features 1024] [--calls 200]
Environment knobs under test: NB200_RESULT_COPY=1 (device buffer + D2H copy instead of mapped-host stores).
"""
import argparse
import ctypes
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
    parser.add_argument("--calls", type=int, default=200)
    parser.add_argument("--torch-stream", action="store_true", help="run on torch's current stream as bench.py does")
    args = parser.parse_args()

    import libnano_b200 as nb
    from libnano_b200 import _lib
    from libnano_b200.function import _ptr

    ctx = nb.Context(0)
    data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, args.rows, args.features, seed=42, noise=0.1)
    function = nb.LinearFunction(data, nb.Loss("s-logistic"), 0.0, 1e-2)
    wrapped = None
    if args.torch_stream:
        from libnano_b200.sharded import ShardedFunction
        wrapped = ShardedFunction.from_local(function, args.rows)
    n = function.size()
    x = np.random.default_rng(0).uniform(-1, 1, n) / np.sqrt(args.features)
    gx = np.empty(n)
    fx = ctypes.c_double()
    L = _lib.lib()

    def timed(call, label):
        for _ in range(5):
            call()
        function.host_profile(reset=True)
        t0 = time.perf_counter()
        for _ in range(args.calls):
            call()
        wall_us = 1e6 * (time.perf_counter() - t0) / args.calls
        a, b, c, calls = function.host_profile(reset=True)
        out = {"what": label, "wall_us_per_call": round(wall_us, 2)}
        if calls:
            out.update(stage_us=round(a / calls, 2), enqueue_us=round(b / calls, 2), wait_us=round(c / calls, 2))
        return out

    lines = [timed(lambda: L.nb200_vgrad(function.handle, _ptr(x), n, _ptr(gx), ctypes.byref(fx)), "ctypes nb200_vgrad"),
             timed(lambda: L.nb200_vgrad(function.handle, _ptr(x), n, None, ctypes.byref(fx)), "ctypes nb200_vgrad value-only"),
             timed(lambda: function(x, gx), "LinearFunction.__call__")]
    if wrapped is not None:
        lines.append(timed(lambda: wrapped(x, gx), "ShardedFunction.__call__"))

    # device-resident: enqueue `calls` evaluations, one sync
    function.enable_timing(True)
    function.pass_stats(reset=True)
    import torch
    x_dev = torch.from_numpy(x).cuda()
    g_dev = torch.empty(n, dtype=torch.float64, device="cuda")
    f_dev = torch.empty(1, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for _ in range(5):
        function.vgrad_device(x_dev.data_ptr(), g_dev.data_ptr(), f_dev.data_ptr())
    function.sync()
    function.pass_stats(reset=True)
    t0 = time.perf_counter()
    for _ in range(args.calls):
        function.vgrad_device(x_dev.data_ptr(), g_dev.data_ptr(), f_dev.data_ptr())
    function.sync()
    wall_us = 1e6 * (time.perf_counter() - t0) / args.calls
    total_ms, count = function.pass_stats(reset=True)
    lines.append({"what": "device-resident loop", "wall_us_per_call": round(wall_us, 2),
                  "kernel_us": round(1e3 * total_ms / max(count, 1), 2)})
    # one evaluation + sync at a time (launch latency + kernel + wake-up, no copies)
    t0 = time.perf_counter()
    for _ in range(args.calls):
        function.vgrad_device(x_dev.data_ptr(), g_dev.data_ptr(), f_dev.data_ptr())
        function.sync()
    wall_us = 1e6 * (time.perf_counter() - t0) / args.calls
    total_ms, count = function.pass_stats(reset=True)
    lines.append({"what": "device-resident, sync per call", "wall_us_per_call": round(wall_us, 2),
                  "kernel_us": round(1e3 * total_ms / max(count, 1), 2)})
    function.enable_timing(False)
    for line in lines:
        line.update(rows=args.rows, features=args.features, result_copy=os.environ.get("NB200_RESULT_COPY", "0"),
                    torch_stream=bool(args.torch_stream), kernel=function.describe())
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
