#!/usr/bin/env python
"""GPU tuning helper: time every fused-kernel variant x stage geometry for one (rows, features) shape and print a
table (kernel ms from CUDA events on the launching stream, achieved GB/s of algorithmic bytes)."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(args):
    import numpy as np

    import libnano_b200 as nb

    ctx = nb.Context(0)
    data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, args.rows, args.features, seed=42)
    function = nb.LinearFunction(data, nb.Loss(args.loss), 0.0, 1e-2)
    try:
        function.set_variant(args.variant)
    except RuntimeError as e:
        print(json.dumps({"variant": args.variant, "error": str(e)}))
        return
    x = np.random.default_rng(0).uniform(-1, 1, function.size()) / np.sqrt(args.features)
    gx = np.empty_like(x)
    for _ in range(5):
        function(x, gx)
    function.enable_timing(True)
    function.pass_stats(reset=True)
    times = []
    for _ in range(args.steps):
        function(x, gx) if not args.value_only else function(x)
        times.append(function.last_pass_ms())
    bytes_alg = 8 * args.rows * args.features + 8 * args.rows + 16 * (args.features + 1)
    best, med = min(times), sorted(times)[len(times) // 2]
    print(json.dumps({"variant": args.variant, "stage_bytes": os.environ.get("NB200_STAGE_BYTES"),
                      "stages": os.environ.get("NB200_STAGES"), "ms_min": round(best, 4), "ms_med": round(med, 4),
                      "gbs_med": round(bytes_alg / med / 1e6, 1), "gbs_best": round(bytes_alg / best / 1e6, 1),
                      "plan": function.describe()}))


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--rows", type=int, default=1 << 20)
    parser.add_argument("--features", type=int, default=1024)
    parser.add_argument("--loss", default="s-logistic")
    parser.add_argument("--steps", type=int, default=20)
    parser.add_argument("--variant", type=int, default=None)
    parser.add_argument("--value-only", action="store_true")
    parser.add_argument("--variants", default="4,5,17,18,19,20,21,22")
    parser.add_argument("--stage-bytes", default="16384,32768,65536")
    parser.add_argument("--stages", default="8")
    args = parser.parse_args()
    if args.variant is not None:
        return one(args)
    for variant in args.variants.split(","):
        for stage_bytes in args.stage_bytes.split(","):
            for stages in args.stages.split(","):
                env = dict(os.environ, NB200_STAGE_BYTES=stage_bytes, NB200_STAGES=stages)
                cmd = [sys.executable, __file__, "--rows", str(args.rows), "--features", str(args.features), "--loss",
                       args.loss, "--steps", str(args.steps), "--variant", variant]
                if args.value_only:
                    cmd.append("--value-only")
                out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
                print(out.stdout.strip() or out.stderr.strip()[-300:], flush=True)


if __name__ == "__main__":
    main()
