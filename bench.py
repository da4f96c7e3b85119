#!/usr/bin/env python
"""bench.py — the headline benchmark: full-batch value-and-gradient (vgrad) throughput of the logistic linear-model
objective, BASELINE.json configs[1]: 2^20 samples x 1024 features x 1 target, fp64, per GPU (weak scaling for N > 1:
every rank holds its own 2^20-row shard of ONE synthetic dataset, one all-reduce of the (loss, gradient) vector per
step). A "step" is one vgrad call = one pass of the hot path over the rank's rows.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rows-per-gpu R] [--features D] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0). Keys follow the driver contract:
  value      whole-job samples*features/s with x, g, f resident in HBM (device-timed, max over ranks)
  e2e        the same metric through the public host-buffer call `function(x, gx)` (H2D of x, D2H of g and f inside)
  roofline   HBM roofline of the fused kernel: algorithmic bytes / average kernel duration (CUDA events on the
             launching stream, inside the timed region) vs MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle's multi-threaded restatement of the reference path on the box's host cores (N = 1)
--impl reference times that CPU path alone (the reference itself cannot be built: Eigen3 is not installed).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "vgrad samples*features/s"
UNIT = "samples*features/s"
LOSS_ID = "s-logistic"
L2REG = 1e-2
SEED = 42
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=20)
    parser.add_argument("--warmup", type=int, default=5)
    parser.add_argument("--rows-per-gpu", type=int, default=1 << 20)
    parser.add_argument("--features", type=int, default=1024)
    parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
    parser.add_argument("--cpu-rows", type=int, default=1 << 17, help="rows of the bounded CPU sample")
    parser.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of cpu_baseline")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--settle-seconds", type=float, default=1.5,
                        help="idle time before the warm-up of each timed region: both regions start from the same "
                             "board state (streaming fp64 from HBM holds the board at its 1000 W power cap, so the "
                             "second region would otherwise inherit the first one's throttled clocks)")
    parser.add_argument("--sustained-seconds", type=float, default=0.0,
                        help="also run the device-resident loop back to back for this long and report the steady "
                             "state (second half) in a 'sustained' object")
    parser.add_argument("--variant", type=int, default=-1, help="fused-kernel variant (tuning), -1 = heuristic")
    parser.add_argument("--collective", default="auto", choices=["auto", "peer", "nccl"],
                        help="N > 1: 'peer' = all-reduce fused into the kernel over NVLink peer memory, 'nccl' = "
                             "two phases around torch.distributed.all_reduce, 'auto' = peer if it can be set up")
    return parser.parse_args()


def workload_name(rows, features):
    log2 = int(np.log2(rows)) if rows > 0 and (rows & (rows - 1)) == 0 else None
    rows_text = f"2^{log2}" if log2 is not None else str(rows)
    return f"{LOSS_ID} linear model, {rows_text} samples x {features} features x 1 target per GPU, fp64 vgrad"


def algorithmic_bytes(rows, features, targets=1):
    # 8*N*D (X once) + 8*N*T (Y) + 16*(D+1)*T (x in, g out)
    return 8 * rows * features + 8 * rows * targets + 16 * (features + 1) * targets


# ----------------------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.lines = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [l for (t, l) in self.lines if t_begin - 0.05 <= t <= t_end + 0.15] or [l for (_, l) in self.lines]
        sm, sm_max, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in rows:
            cells = [c.strip() for c in line.split(",")]
            if len(cells) < 9:
                continue
            try:
                sm.append(float(cells[1]))
                sm_max.append(float(cells[2]))
            except ValueError:
                continue
            for name, cell in zip(names, cells[5:9]):
                if cell.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(sm_max) if sm_max else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's multi-threaded restatement of linear::function_t::do_eval on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_sample(rows, features, X=None, Y=None):
    if X is None:
        rng = np.random.default_rng(SEED)
        X = rng.uniform(-1.0, 1.0, (rows, features))
        w = rng.uniform(-1.0, 1.0, features) / np.sqrt(features)
        Y = np.where(X @ w + 0.1 + 0.1 * rng.standard_normal(rows) < 0, -1.0, 1.0).reshape(rows, 1)
    return np.ascontiguousarray(X), np.ascontiguousarray(Y)


def time_cpu(X, Y, budget_seconds, min_calls=3, max_calls=100000):
    """Mirror of nano::measure (include/nano/core/chrono.h:81-114): repeated calls, minimum time; x = 0 as in
    app/bench_function.cpp:16. Returns (samples*features/s from the min, per-call seconds list, threads)."""
    import oracle

    oracle.build()
    threads = oracle.hardware_threads()
    rows, features = X.shape
    x = np.zeros(features + 1)
    oracle.linear_vgrad(X, Y, LOSS_ID, x, l2=L2REG, mode="fast", threads=threads)  # warm-up (page-in, thread start)
    times = []
    start = time.perf_counter()
    while len(times) < max_calls and (len(times) < min_calls or time.perf_counter() - start < budget_seconds):
        t0 = time.perf_counter()
        oracle.linear_vgrad(X, Y, LOSS_ID, x, l2=L2REG, mode="fast", threads=threads)
        times.append(time.perf_counter() - t0)
    return rows * features / min(times), times, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # rank 0 alone runs the CPU arm; the other ranks exit without work
    import oracle

    oracle.build()
    threads = oracle.hardware_threads()
    rows, features = args.cpu_rows, args.features
    X, Y = cpu_sample(rows, features)
    x = np.zeros(features + 1)
    for _ in range(max(args.warmup, 1)):
        oracle.linear_vgrad(X, Y, LOSS_ID, x, l2=L2REG, mode="fast", threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.linear_vgrad(X, Y, LOSS_ID, x, l2=L2REG, mode="fast", threads=threads)
    elapsed = time.perf_counter() - t0
    value = rows * features * args.steps / elapsed
    sample = (f"{rows} rows x {features} features per step (bounded sample of the {args.rows_per_gpu}-row workload; "
              "throughput is size-independent once X is out of cache)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.rows_per_gpu, args.features), "loss": LOSS_ID, "l2": L2REG,
                   "note": "CPU restatement (oracle/) of src/linear/function.cpp:44-168, batch=100, one std::thread "
                           "per host core; the reference's Eigen build is unavailable (Eigen3 not installed)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic(rows, features):
    """dram bytes per launch of the fused kernel from the committed ncu --set full capture, if it matches this shape."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            data = json.load(f)
        if data.get("rows") == rows and data.get("features") == features:
            return data.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


def run_b200(args):
    import torch
    import torch.distributed as dist

    import libnano_b200 as nb
    from libnano_b200.sharded import ShardedFunction

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    rows, features = args.rows_per_gpu, args.features
    n_global = rows * world
    ctx = nb.Context(local_rank)
    data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, rows, features, seed=SEED, row_offset=rank * rows, noise=0.1)
    local = nb.LinearFunction(data, nb.Loss(LOSS_ID), 0.0, L2REG)
    if args.variant >= 0:
        local.set_variant(args.variant)
    function = ShardedFunction.from_local(local, n_global)  # re-points `local` at torch's current stream
    n = function.size()
    collective = "none"
    if world > 1:
        collective = "nccl"
        if args.collective in ("auto", "peer"):
            try:
                if function.connect_peers():
                    collective = "peer"
            except Exception as error:  # noqa: BLE001 - fall back to the NCCL path, say so in the JSON line
                if args.collective == "peer":
                    raise
                collective = f"nccl (peer setup failed: {error})"
        flags = [None] * world
        dist.all_gather_object(flags, collective == "peer")
        if not all(flags):  # every rank must take the same path
            function._shard.fused = False
            collective = collective if collective != "peer" else "nccl"

    rng = np.random.default_rng(0)
    x = rng.uniform(-1.0, 1.0, n) / np.sqrt(features)
    gx = np.empty(n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- (1) device-resident loop: x, g, f stay in HBM -------------------------------------------------------
    x_dev = torch.from_numpy(x).cuda()
    g_dev = torch.empty(n, dtype=torch.float64, device="cuda")
    f_dev = torch.empty(1, dtype=torch.float64, device="cuda")

    def step_device():
        if world == 1:
            # one cooperative launch: pass over X, cross-CTA reduction, scaling + regularisation
            local.vgrad_device(x_dev.data_ptr(), g_dev.data_ptr(), f_dev.data_ptr())
        elif collective == "peer":
            # ONE launch per rank: pass over X + all-reduce over NVLink peer memory + scaling, no NCCL call
            local.vgrad_sharded_device(x_dev.data_ptr(), n_global, g_dev.data_ptr(), f_dev.data_ptr())
        else:
            ptr, length = local.partial_device(x_dev.data_ptr(), True)
            dist.all_reduce(function._shard._view_for(ptr, length))
            local.finish_device(x_dev.data_ptr(), n_global, g_dev.data_ptr(), f_dev.data_ptr())

    barrier()
    time.sleep(args.settle_seconds)
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    local.enable_timing(True)
    local.pass_stats(reset=True)
    launches0 = local.launches()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    start.record()
    for _ in range(args.steps):
        step_device()
    stop.record()
    barrier()
    t_end = time.perf_counter()
    device_ms = max_over_ranks(start.elapsed_time(stop))
    launches = local.launches() - launches0
    local.sync()
    pass_ms_total, pass_count = local.pass_stats(reset=True)
    local.enable_timing(False)
    f_device = float(f_dev.item())

    # ---- (2) end-to-end loop: the public call with HOST buffers ---------------------------------------------------
    barrier()
    time.sleep(args.settle_seconds)
    for _ in range(max(args.warmup, 3)):
        function(x, gx)
    barrier()
    start2, stop2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start2.record()
    for _ in range(args.steps):
        f_host = function(x, gx)
    stop2.record()
    barrier()
    e2e_ms = max_over_ranks(start2.elapsed_time(stop2))
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None

    sustained = None
    if args.sustained_seconds > 0:
        # back-to-back evaluations for a fixed wall time; the second half is the steady state under the power cap
        local.enable_timing(True)
        local.pass_stats(reset=True)
        batches = []
        t_sustained = time.perf_counter()
        while time.perf_counter() - t_sustained < args.sustained_seconds:
            for _ in range(25):
                step_device()
            barrier()
            local.sync()
            total_ms, count = local.pass_stats(reset=True)
            batches.append(total_ms / max(count, 1))
        local.enable_timing(False)
        steady_ms = max_over_ranks(float(np.mean(batches[len(batches) // 2:])))
        sustained = {"seconds": args.sustained_seconds, "launches": 25 * len(batches), "kernel_ms_first": batches[0],
                     "kernel_ms_steady": steady_ms}

    if not np.isfinite(f_host) or abs(f_host - f_device) > 1e-12 * (1 + abs(f_host)):
        raise SystemExit(f"bench.py: device-resident and host-buffer evaluations disagree: {f_device} vs {f_host}")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    work = float(n_global) * features * args.steps  # samples*features over all ranks
    value = work / (device_ms * 1e-3)
    e2e_value = work / (e2e_ms * 1e-3)
    kernel_ms = pass_ms_total / max(pass_count, 1)
    bytes_alg = algorithmic_bytes(rows, features)
    peak, peak_source = measured_hbm_peak()
    achieved = bytes_alg / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": device_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(rows, features), "loss": LOSS_ID, "l2": L2REG,
                   "rows_per_gpu": rows, "features": features, "targets": 1, "global_rows": n_global,
                   "parallelism": f"sample-sharded x{world}, one all-reduce of {n + 1} doubles per step"
                   if world > 1 else "single GPU",
                   "collective": {"peer": "fused into the kernel: peer-memory all-gather over NVLink + rank-ordered sum",
                                  "none": "none (single GPU)"}.get(collective, collective),
                   "l2_flush": f"inputs larger than L2: {8 * rows * features / 2 ** 30:.1f} GiB of X per GPU is "
                               "streamed every step (L2 is 126 MB)",
                   "kernel": local.describe()},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * (n + 1),
                "api": "libnano_b200.ShardedFunction.__call__(x, gx) (host buffers) -> " +
                       {"none": "nb200_vgrad", "peer": "nb200_vgrad_sharded"}.get(collective,
                                                                                 "nb200_vgrad_partial / finish")},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": recorded_traffic(rows, features),
                     "kernel": "vgrad_t1_kernel", "kernel_ms": kernel_ms, "launches_timed": int(pass_count),
                     "algorithmic_bytes_per_launch": bytes_alg, "peak_source": peak_source},
        "clocks": clocks,
        "f": f_host,
    }
    line["config"]["settle"] = (f"{args.settle_seconds} s idle before the warm-up of each timed region (device-resident "
                                "and end-to-end), so both start from the same board state")
    if sustained is not None:
        sustained["achieved_gbs"] = bytes_alg / (sustained["kernel_ms_steady"] * 1e-3) / 1e9
        sustained["frac"] = sustained["achieved_gbs"] / peak
        sustained["value"] = float(n_global) * features / (sustained["kernel_ms_steady"] * 1e-3)
        line["sustained"] = sustained

    if world == 1 and not args.no_cpu_baseline:
        cpu_rows = min(args.cpu_rows, rows)
        Xs, Ys = data.read(0, cpu_rows)
        cpu_value, cpu_times, threads = time_cpu(Xs, Ys, args.cpu_seconds)
        line["cpu_baseline"] = {
            "value": cpu_value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {cpu_rows} rows x {features} features of the same dataset, {len(cpu_times)} calls, "
                      f"min {min(cpu_times) * 1e3:.1f} ms (nano::measure takes the minimum); oracle/ restatement of "
                      "src/linear/function.cpp:44-168, batch=100, one std::thread per host core",
        }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
