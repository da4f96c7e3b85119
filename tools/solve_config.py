#!/usr/bin/env python
"""End-to-end solves of the BASELINE configs with the product's drivers (no oracle involved):
  --config 2   s-logistic 2^20 x 1024, L-BFGS (device-resident) to epsilon 1e-8, 1 GPU
  --config 4   the same objective sample-sharded, 2^23 rows per GPU by default (torchrun, one rank per GPU)
  --config 5   s-hinge 2^24 x 512 with lasso (l1 = 1e-2) and elastic net (l1 = l2 = 1e-2), OSGA (host-driven)
Prints one JSON line per solve on rank 0: status, f, calls, wall time, time per evaluation."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", type=int, required=True, choices=[2, 4, 5])
    parser.add_argument("--rows-per-gpu", type=int, default=None)
    parser.add_argument("--max-evals", type=int, default=None)
    parser.add_argument("--nonsmooth-solvers", default="osga", help="config 5: comma-separated subset of osga,rqb")
    args = parser.parse_args()

    import torch
    import torch.distributed as dist

    import libnano_b200 as nb
    from libnano_b200.sharded import ShardedFunction
    from libnano_b200.solver import lbfgs, osga, rqb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = nb.Context(local_rank)

    def report(name, result, seconds, extra):
        if rank == 0:
            evals = result["fcalls"]
            print(json.dumps({"solve": name, "world": world, "status": result["status"], "fx": result["fx"],
                              "gradient_test": result.get("gradient_test"), "eta": result.get("eta"),
                              "fcalls": result["fcalls"], "gcalls": result["gcalls"],
                              "iterations": result["iterations"], "seconds": round(seconds, 4),
                              "ms_per_fcall": round(1e3 * seconds / max(evals, 1), 4), **extra}), flush=True)

    if args.config in (2, 4):
        rows = args.rows_per_gpu or ((1 << 20) if args.config == 2 else (1 << 23))
        D = 1024
        data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, rows, D, seed=42, row_offset=rank * rows, noise=0.1)
        for l2 in (0.0, 1e-2):
            local = nb.LinearFunction(data, nb.Loss("s-logistic"), 0.0, l2)
            function = ShardedFunction.from_local(local, rows * world)
            fused = bool(world > 1 and function.connect_peers())  # all-reduce inside the kernel over NVLink peers
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            result = lbfgs(function, np.zeros(function.size()), epsilon=1e-8, max_evals=args.max_evals or 1000)
            seconds = time.perf_counter() - t0
            report(f"config{args.config} s-logistic l2={l2} lbfgs(device-resident)", result, seconds,
                   {"rows_per_gpu": rows, "features": D, "plan": local.describe(),
                    "collective": "fused peers" if world > 1 and fused else ("nccl" if world > 1 else "none")})
    else:
        rows = args.rows_per_gpu or (1 << 24)
        D = 512
        data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, rows, D, seed=42, row_offset=rank * rows, noise=0.1)
        for name, l1, l2 in (("lasso", 1e-2, 0.0), ("elastic-net", 1e-2, 1e-2)):
            local = nb.LinearFunction(data, nb.Loss("s-hinge"), l1, l2)
            function = ShardedFunction.from_local(local, rows * world)
            for solver_id in args.nonsmooth_solvers.split(","):
                solver = {"osga": osga, "rqb": rqb}[solver_id]
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                result = solver(function, np.zeros(function.size()), epsilon=1e-6, max_evals=args.max_evals or 5000)
                seconds = time.perf_counter() - t0
                report(f"config5 s-hinge {name} {solver_id}(host-driven)", result, seconds,
                       {"rows_per_gpu": rows, "features": D, "smooth": function.smooth(), "plan": local.describe()})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
