#!/usr/bin/env python
"""bench.py — full-batch value-and-gradient (vgrad) throughput of libnano's linear-model objective on B200.

Main line (every N): BASELINE.json configs[3]'s shard — s-logistic, 2^23 samples x 1024 features x 1 target of fp64 per
GPU (64 GiB of X per GPU; at N = 8 that is literally the 64 Mi x 1024 objective of the target sentence; weak scaling:
every rank holds its own row block of ONE synthetic dataset, one exchange of the (loss, gradient) vector per step).
A "step" is one vgrad = one pass of the hot path over the rank's rows.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rows-per-gpu R] [--features D] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0):
  value        whole-job samples*features/s with x, g, f resident in HBM (device-timed, max over ranks)
  e2e          the same metric through the public host-buffer call `function(x, gx)` (H2D of x, D2H of g and f inside)
  roofline     HBM roofline of the fused kernel: algorithmic bytes / average kernel duration (CUDA events recorded by
               the library around every launch of the timed region, on the launching stream) vs MEASURED_PEAKS.json
  sustained    the steady state of >= 3 s of back-to-back steps (the board sits at its power cap while it streams fp64)
  parity       a second, small dataset through the SAME (sharded) path against the CPU oracle, in this run
  per_rank     N > 1: every rank's kernel-time distribution, step time, SM clock and power (names the scaling limiter)
  configs      N = 1: BASELINE configs 1, 2, 3 and 5, each with its own roofline, clocks and oracle parity
  cpu_baseline the CPU oracle's multi-threaded restatement of the reference path on the box's host cores (N = 1)
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
PARITY_BOUND = 1e-12       # BASELINE.json north_star: relative, fp64


def parse_args():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=20)
    parser.add_argument("--warmup", type=int, default=5)
    parser.add_argument("--rows-per-gpu", type=int, default=1 << 23)
    parser.add_argument("--features", type=int, default=1024)
    parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
    parser.add_argument("--cpu-rows", type=int, default=1 << 20,
                        help="rows of the bounded CPU sample (the config-2 dataset: 2^20 rows = 8 GiB of X)")
    parser.add_argument("--cpu-calls", type=int, default=16, help="timed calls of cpu_baseline (nano::measure uses 16)")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--no-configs", action="store_true", help="N = 1: skip the sub-benchmarks of configs 1/2/3/5")
    parser.add_argument("--settle-seconds", type=float, default=1.5,
                        help="idle time before the warm-up of each timed region: every region starts from the same "
                             "board state (streaming fp64 from HBM holds the board at its 1000 W power cap)")
    parser.add_argument("--sustained-seconds", type=float, default=3.0,
                        help="back-to-back steps for this long; the second half is reported as the steady state")
    parser.add_argument("--variant", type=int, default=-1, help="fused-kernel variant (tuning), -1 = heuristic")
    parser.add_argument("--collective", default="auto", choices=["auto", "peer", "nccl"],
                        help="N > 1: 'peer' = sums published into the peers' mailboxes from inside the kernel (NVLink), "
                             "'nccl' = two phases around torch.distributed.all_reduce, 'auto' = peer if it can be set up")
    parser.add_argument("--parity-rows", type=int, default=1 << 14, help="rows of the in-run parity dataset (all ranks)")
    parser.add_argument("--no-balance", action="store_true",
                        help="N > 1: keep equal row blocks instead of sizing them by each GPU's measured speed")
    return parser.parse_args()


def log2_text(rows):
    return f"2^{int(np.log2(rows))}" if rows > 0 and (rows & (rows - 1)) == 0 else str(rows)


def workload_config(args, world):
    """The workload-defining part of `config`: the same dict in the GPU arm and in the reference arm."""
    rows, features = args.rows_per_gpu, args.features
    return {
        "workload": f"{LOSS_ID} linear model, {log2_text(rows)} samples x {features} features x 1 target per GPU, fp64 "
                    "vgrad (BASELINE configs[3]'s shard; configs[1] is configs.cfg2)",
        "loss": LOSS_ID, "l1": 0.0, "l2": L2REG, "rows_per_gpu": rows, "features": features, "targets": 1,
        "global_rows": rows * world,
        "parallelism": f"sample-sharded x{world}, one exchange of {features + 2} doubles per step" if world > 1
                       else "single GPU",
        "l2_flush": f"inputs larger than L2: {8 * rows * features / 2 ** 30:.1f} GiB of X per GPU is streamed every "
                    "step (L2 is 126 MB)",
    }


def algorithmic_bytes(rows, features, targets=1):
    # 8*N*D (X once) + 8*N*T (Y) + 16*(D+1)*T (x in, g out)
    return 8 * rows * features + 8 * rows * targets + 16 * (features + 1) * targets


# ----------------------------------------------------------------------------------------------------------------
# clocks: sampled for the whole run, reported per timed window
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

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
            return self
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()
        time.sleep(0.25)
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def window(self, t_begin, t_end):
        """Median SM clock, max power and the throttle reasons seen in [t_begin, t_end] (at least one sample: a window
        shorter than the 100 ms period takes the samples around it)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        deadline = time.perf_counter() + 0.35
        while time.perf_counter() < deadline and not any(t >= t_end for t, _ in self.lines):
            time.sleep(0.02)
        rows = [l for (t, l) in self.lines if t_begin <= t <= t_end + 0.11]
        if not rows:
            rows = [l for (t, l) in self.lines if t_begin - 0.11 <= t <= t_end + 0.25] or [l for _, l in self.lines[-2:]]
        sm, sm_max, power, reasons = [], [], [], set()
        for line in rows:
            cells = [c.strip() for c in line.split(",")]
            if len(cells) < 9:
                continue
            try:
                sm.append(float(cells[1]))
                sm_max.append(float(cells[2]))
                power.append(float(cells[3]))
            except ValueError:
                continue
            for name, cell in zip(self.NAMES, cells[5:9]):
                if cell.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(sm_max) if sm_max else None,
                "power_w": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's multi-threaded restatement of linear::function_t::do_eval on the host cores
# ----------------------------------------------------------------------------------------------------------------
def host_sample(rows, features, seed=SEED):
    """The generating process of the GPU dataset (uniform features, noisy linear labels) drawn on the host, in
    parallel (numpy releases the GIL): 2^20 x 1024 doubles take about a second instead of ten."""
    from concurrent.futures import ThreadPoolExecutor

    X = np.empty((rows, features))
    w = np.random.default_rng(seed).uniform(-1.0, 1.0, features) / np.sqrt(features)
    Y = np.empty((rows, 1))
    chunk = max(1, rows // max(1, 4 * (os.cpu_count() or 1)))
    spans = [(lo, min(rows, lo + chunk)) for lo in range(0, rows, chunk)]

    def fill(index):
        lo, hi = spans[index]
        rng = np.random.default_rng([seed, index])
        rng.random(out=X[lo:hi].reshape(-1))
        X[lo:hi] *= 2.0
        X[lo:hi] -= 1.0
        Y[lo:hi, 0] = np.where(X[lo:hi] @ w + 0.1 + 0.1 * rng.standard_normal(hi - lo) < 0, -1.0, 1.0)

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as pool:
        list(pool.map(fill, range(len(spans))))
    return X, Y


def time_cpu(X, Y, calls, warmup=3):
    """Mirror of nano::measure (include/nano/core/chrono.h:81-114): repeated calls; x = 0 as in
    app/bench_function.cpp:16. Returns (per-call seconds, threads): the arms report the mean over the timed calls as
    `value` (what `--steps K` times) and the minimum (what nano::measure reports) beside it."""
    import oracle

    oracle.build()
    threads = oracle.hardware_threads()
    x = np.zeros(X.shape[1] + 1)
    for _ in range(warmup):  # page-in, thread start
        oracle.linear_vgrad(X, Y, LOSS_ID, x, l2=L2REG, mode="fast", threads=threads)
    times = []
    for _ in range(calls):
        t0 = time.perf_counter()
        oracle.linear_vgrad(X, Y, LOSS_ID, x, l2=L2REG, mode="fast", threads=threads)
        times.append(time.perf_counter() - t0)
    return times, threads


def cpu_baseline_entry(times, threads, rows, features, where, warmup=3):
    mean_s, min_s = float(np.mean(times)), float(np.min(times))
    return {
        "value": rows * features / mean_s, "unit": UNIT, "cores": threads, "kind": "port",
        "value_min_of_calls": rows * features / min_s,
        "sample": f"{where}: {rows} rows x {features} features per call, {len(times)} timed calls after {warmup} warm-up calls, "
                  f"mean {mean_s * 1e3:.1f} ms (= value), min {min_s * 1e3:.1f} ms (what nano::measure reports); oracle/ "
                  "restatement of src/linear/function.cpp:44-168, batch=100, one std::thread per host core; the "
                  "reference's Eigen build is unavailable (Eigen3 not installed)",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return 0  # rank 0 alone runs the CPU arm; the other ranks exit without work
    rows, features = min(args.cpu_rows, args.rows_per_gpu), args.features
    X, Y = host_sample(rows, features)
    times, threads = time_cpu(X, Y, args.steps, warmup=max(args.warmup, 1))
    entry = cpu_baseline_entry(times, threads, rows, features,
                               f"bounded sample of the {log2_text(args.rows_per_gpu)}-row workload (throughput is "
                               "size-independent once X is out of cache)", warmup=max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": entry["value"], "unit": UNIT, "n_gpus": max(args.gpus, world),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, max(args.gpus, world)),
        "cpu_baseline": entry,
        "e2e": {"value": entry["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs: a copy, read + write)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic(kernel_text, rows, features):
    """dram bytes per launch of this kernel on this shape from a committed `ncu --set full` capture (profiles/), or
    None: the figure is never extrapolated to another shape or kernel variant."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            captures = json.load(f)
    except Exception:
        return None
    for capture in captures if isinstance(captures, list) else [captures]:
        if capture.get("rows") == rows and capture.get("features") == features and \
                f"variant={capture.get('variant')} " in kernel_text + " ":
            return capture.get("dram_bytes_per_launch")
    return None


def summary(samples):
    samples = np.asarray(samples, dtype=np.float64)
    if samples.size == 0:
        return None
    return {"min": float(samples.min()), "median": float(np.median(samples)), "mean": float(samples.mean()),
            "max": float(samples.max()), "launches": int(samples.size)}


class Harness:
    """Timing plumbing shared by the main line and the sub-benchmarks."""

    def __init__(self, args, torch, dist, world, rank, sampler):
        self.args, self.torch, self.dist, self.world, self.rank, self.sampler = args, torch, dist, world, rank, sampler

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, value):
        if self.world > 1:
            t = self.torch.tensor([value], dtype=self.torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return value

    def region(self, step, local, steps, warmup, settle=None):
        """W untimed warm-up steps, then exactly `steps` timed ones bracketed by barrier + synchronize; CUDA events on
        the launching stream (the objective is ordered on torch's current stream); max over ranks."""
        torch = self.torch
        self.barrier()
        time.sleep(self.args.settle_seconds if settle is None else settle)
        for _ in range(max(warmup, 3)):
            step()
        self.barrier()
        local.enable_timing(True)
        local.pass_stats(reset=True)
        launches0 = local.launches()
        self.barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin = time.perf_counter()
        start.record()
        for _ in range(steps):
            step()
        stop.record()
        self.barrier()
        t_end = time.perf_counter()
        own_ms = start.elapsed_time(stop)
        local.sync()
        passes = local.pass_samples()
        local.pass_stats(reset=True)
        local.enable_timing(False)
        return {"ms": self.max_over_ranks(own_ms), "own_ms": own_ms, "passes": passes,
                "launches": local.launches() - launches0, "window": (t_begin, t_end)}

    def sustained(self, step, seconds, batch=25):
        """Back-to-back steps for `seconds`; per batch the device time of the batch (events); the second half of the
        batches is the steady state under the power cap."""
        torch = self.torch
        if seconds <= 0:
            return None
        self.barrier()
        per_step, t0 = [], time.perf_counter()
        while True:
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
            for _ in range(batch):
                step()
            stop.record()
            stop.synchronize()
            per_step.append(start.elapsed_time(stop) / batch)
            # every rank runs the same number of batches: rank 0's clock decides
            go = self.torch.tensor([1.0 if time.perf_counter() - t0 < seconds else 0.0], device="cuda")
            if self.world > 1:
                self.dist.broadcast(go, src=0)
            if float(go.item()) == 0.0:
                break
        t1 = time.perf_counter()
        steady = self.max_over_ranks(float(np.mean(per_step[len(per_step) // 2:])))
        return {"seconds": round(t1 - t0, 2), "steps": batch * len(per_step), "ms_per_step_first": per_step[0],
                "ms_per_step_steady": steady, "window": (t0 + (t1 - t0) / 2, t1)}


def oracle_parity(oracle, X, Y, loss_id, x, l1, l2, fx, gx):
    f_ref, g_ref = oracle.linear_vgrad(X, Y, loss_id, x, l1=l1, l2=l2, threads=oracle.hardware_threads())
    rel_f = abs(fx - f_ref) / (1.0 + abs(f_ref))
    rel_g = float(np.max(np.abs(gx - g_ref))) / (1.0 + float(np.max(np.abs(g_ref))))
    return {"rel_f": rel_f, "rel_g": rel_g, "bound": PARITY_BOUND, "ok": bool(rel_f <= PARITY_BOUND and rel_g <= PARITY_BOUND)}


def bench_streaming_config(h, nb, ctx, name, loss_id, rows, features, l1, l2, peak, peak_source, steps, warmup,
                           keep_dataset=False):
    """One single-target configuration on one GPU: burst + sustained throughput, roofline, clocks, and the SAME kernel
    plan on a small sample of the shape against the oracle."""
    import oracle

    torch = h.torch
    data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, rows, features, seed=SEED, noise=0.1)
    function = nb.LinearFunction(data, nb.Loss(loss_id), l1, l2)
    function.set_stream(torch.cuda.current_stream().cuda_stream)
    n = function.size()
    x = np.random.default_rng(0).uniform(-1.0, 1.0, n) / np.sqrt(features)
    x_dev = torch.from_numpy(x).cuda()
    g_dev = torch.empty(n, dtype=torch.float64, device="cuda")
    f_dev = torch.empty(1, dtype=torch.float64, device="cuda")

    def step():
        function.vgrad_device(x_dev.data_ptr(), g_dev.data_ptr(), f_dev.data_ptr())

    burst = h.region(step, function, steps, warmup)
    clocks = h.sampler.window(*burst["window"])
    steady = h.sustained(step, h.args.sustained_seconds)
    kernel = summary(burst["passes"])
    bytes_alg = algorithmic_bytes(rows, features)
    achieved = bytes_alg / (kernel["mean"] * 1e-3) / 1e9
    entry = {
        "workload": f"{loss_id} linear model, {log2_text(rows)} samples x {features} features x 1 target, fp64 vgrad, "
                    f"l1={l1:g} l2={l2:g}",
        "metric": METRIC, "value": rows * features * steps / (burst["ms"] * 1e-3), "unit": UNIT,
        "ms_per_step": burst["ms"] / steps, "steps": steps, "kernel": function.describe(),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic(function.describe(), rows, features), "kernel_ms": kernel,
                     "algorithmic_bytes_per_launch": bytes_alg, "peak_source": peak_source},
        "clocks": clocks, "f": float(f_dev.item()),
    }
    if steady is not None:
        window = steady.pop("window")
        steady["achieved_gbs"] = bytes_alg / (steady["ms_per_step_steady"] * 1e-3) / 1e9
        steady["frac"] = steady["achieved_gbs"] / peak
        steady["value"] = rows * features / (steady["ms_per_step_steady"] * 1e-3)
        steady["clocks"] = h.sampler.window(*window)
        entry["sustained"] = steady
    # parity: the same kernel plan (it depends on D and the loss, not on N) on the first rows of this dataset
    sample_rows = min(rows, 1 << 14)
    Xs, Ys = data.read(0, sample_rows)
    small = nb.LinearFunction(nb.Dataset.pin(ctx, Xs, Ys), nb.Loss(loss_id), l1, l2)
    gx = np.empty(n)
    fx = small(x, gx)
    entry["parity"] = dict(oracle_parity(oracle, Xs, Ys, loss_id, x, l1, l2, fx, gx), rows=sample_rows,
                           same_plan=small.describe().split(" grid=")[0] == function.describe().split(" grid=")[0])
    small.close()
    if keep_dataset:
        return entry, data, function, x
    function.close()
    data.close()
    return entry


def bench_config1(h, nb, ctx):
    """BASELINE configs[0]: the builtin synthetic test function mse+ridge (objective B, src/function/mlearn/ridge.cpp),
    seed 42, alpha2 = 1, at D = 100 (sratio 100 -> N = 10 000) and at bench_function's default D = 1024 (N = 10 240).
    L2-resident and launch-latency bound: microseconds per f(x, g) as app/bench_function.cpp:47-49 measures it
    (nano::measure: minimum over 16 trials, x = 0), beside the single-thread CPU port on the same data."""
    import oracle

    torch = h.torch
    out = []
    for dims, sratio in ((100, 100.0), (1024, 10.0)):
        function = nb.SyntheticFunction(ctx, "mse", "ridge", dims, seed=42, alpha2=1.0, sratio=sratio)
        N = function.dataset.samples
        x, gx = np.zeros(dims), np.empty(dims)
        for _ in range(20):
            function(x, gx)
        function.enable_timing(True)
        best_call = 1e9
        t_begin = time.perf_counter()
        for _ in range(16):
            t0 = time.perf_counter()
            for _ in range(50):
                fx = function(x, gx)
            best_call = min(best_call, (time.perf_counter() - t0) / 50)
        t_end = time.perf_counter()
        function.sync()
        kernel_ms = function.pass_samples()
        function.enable_timing(False)
        # device-resident calls back to back (no host round trip): what a device-resident solver pays per evaluation
        function.set_stream(torch.cuda.current_stream().cuda_stream)
        x_dev = torch.zeros(dims, dtype=torch.float64, device="cuda")
        g_dev = torch.empty(dims, dtype=torch.float64, device="cuda")
        f_dev = torch.empty(1, dtype=torch.float64, device="cuda")
        for _ in range(20):
            function.vgrad_device(x_dev.data_ptr(), g_dev.data_ptr(), f_dev.data_ptr())
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(200):
            function.vgrad_device(x_dev.data_ptr(), g_dev.data_ptr(), f_dev.data_ptr())
        stop.record()
        torch.cuda.synchronize()
        resident_us = start.elapsed_time(stop) / 200 * 1e3
        # the CPU port, one thread (the reference's bench_function is single-threaded: app/CMakeLists.txt:10-16)
        X, Y = function.dataset.read()
        bopt = function._boptimum
        cpu_best, f_ref, g_ref = 1e9, None, None
        for _ in range(16):
            t0 = time.perf_counter()
            f_ref, g_ref = oracle.synth_vgrad(X, Y[:, 0], bopt, "mse", x, alpha2=1.0)
            cpu_best = min(cpu_best, time.perf_counter() - t0)
        rel_f = abs(fx - f_ref) / (1 + abs(f_ref))
        rel_g = float(np.max(np.abs(gx - g_ref))) / (1 + float(np.max(np.abs(g_ref))))
        out.append({
            "function": function.name(), "samples": N, "dims": dims,
            "us_per_vgrad_host_call": best_call * 1e6, "us_per_vgrad_device_resident": resident_us,
            "kernel_us": float(np.min(kernel_ms)) * 1e3 if len(kernel_ms) else None,
            "value": N * dims / best_call, "unit": UNIT,
            "cpu_port_single_thread_us": cpu_best * 1e6, "cpu_port_value": N * dims / cpu_best,
            "parity": {"rel_f": rel_f, "rel_g": rel_g, "bound": PARITY_BOUND,
                       "ok": bool(rel_f <= PARITY_BOUND and rel_g <= PARITY_BOUND)},
            "kernel": function.describe(), "clocks": h.sampler.window(t_begin, t_end),
            "note": "L2-resident (8-84 MB), launch-latency bound: us per call is the metric, no roofline fraction",
        })
        function.close()
    return {"workload": "builtin synthetic mse+ridge test function (objective B), D=100 / N=10000 and D=1024 / N=10240, "
                        "seed 42, x=0; min over 16 trials like nano::measure", "cases": out}


def fp64_peaks(h, ctx):
    """The fp64 ceilings of this board, measured in this run: the DMMA / DFMA probes of the library and cuBLAS DGEMM."""
    torch = h.torch
    peaks = {"dmma_tflops": ctx.probe_peak("dmma_tflops"), "dfma_tflops": ctx.probe_peak("dfma_tflops"),
             "hbm_read_gbs": ctx.probe_peak("hbm_read_gbs")}
    S = 8192
    P = torch.randn(S, S, dtype=torch.float64, device="cuda")
    Q = torch.randn(S, S, dtype=torch.float64, device="cuda")
    P @ Q
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        P @ Q
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    peaks["cublas_dgemm_8192_tflops"] = 2.0 * S ** 3 / best / 1e9
    del P, Q
    torch.cuda.empty_cache()
    return peaks


def bench_handful_of_targets(h, nb, ctx, peaks, hbm_peak, steps=20, warmup=5):
    """Not a BASELINE config: the one-pass tensor-pipe kernels for a handful of targets, softmax over
    2 GiB of X per shape. Two ceilings per case: X read once from HBM, and the fp64 tensor pipe with the targets padded
    to whole m8n8k4 tiles (T = 10 occupies 16 columns)."""
    import oracle

    torch = h.torch
    cases = []
    for D, T in ((784, 10), (512, 16), (1024, 8)):
        N = (1 << 31) // (8 * D)
        data = nb.Dataset.synth(ctx, nb.SYNTH_MULTICLASS, N, D, T=T, seed=SEED)
        function = nb.LinearFunction(data, nb.Loss("s-classnll"), 0.0, L2REG)
        function.set_stream(torch.cuda.current_stream().cuda_stream)
        n = function.size()
        x = np.random.default_rng(0).uniform(-1.0, 1.0, n) / np.sqrt(D)
        x_dev = torch.from_numpy(x).cuda()
        g_dev = torch.empty(n, dtype=torch.float64, device="cuda")
        f_dev = torch.empty(1, dtype=torch.float64, device="cuda")

        def step():
            function.vgrad_device(x_dev.data_ptr(), g_dev.data_ptr(), f_dev.data_ptr())

        burst = h.region(step, function, steps, warmup, settle=0.5)
        ms = burst["ms"] / steps
        padded = 4.0 * N * D * ((T + 7) // 8 * 8)
        entry = {"features": D, "targets": T, "rows": N, "loss": "s-classnll", "kernel": function.describe(),
                 "ms": ms, "x_once_gbs": 8.0 * N * D / ms / 1e6, "frac_of_hbm_copy_peak": 8.0 * N * D / ms / 1e6 / hbm_peak,
                 "tflops": 4.0 * N * D * T / ms / 1e9, "padded_tflops": padded / ms / 1e9,
                 "frac_of_dmma_peak_padded": padded / ms / 1e9 / peaks["dmma_tflops"],
                 "clocks": h.sampler.window(*burst["window"])}
        sample_rows = 2048
        Xs, Ys = data.read(0, sample_rows)
        small = nb.LinearFunction(nb.Dataset.pin(ctx, Xs, Ys), nb.Loss("s-classnll"), 0.0, L2REG)
        gx = np.empty(n)
        fx = small(x, gx)
        entry["parity"] = dict(oracle_parity(oracle, Xs, Ys, "s-classnll", x, 0.0, L2REG, fx, gx), rows=sample_rows,
                               same_plan=small.describe().split(" ")[0] == function.describe().split(" ")[0])
        small.close()
        function.close()
        data.close()
        del x_dev, g_dev, f_dev
        torch.cuda.empty_cache()
        cases.append(entry)
    return {"workload": "softmax linear models with a handful of classes, 2 GiB of X per shape, fp64 vgrad (one pass "
                        "over X, both contractions on the fp64 tensor pipe)",
            "metric": "vgrad ms", "unit": "ms", "higher_is_better": False, "steps": steps, "cases": cases}


def bench_config3(h, nb, ctx, peaks, steps=5, warmup=3):
    """BASELINE configs[2]: softmax cross-entropy, 2^22 x 2048 x 100 fp64 — a dense contraction on the FP64 tensor pipe
    (mma.sync m8n8k4 f64; tcgen05 has no fp64 kind). flops_alg = 4*N*D*T per vgrad."""
    import oracle

    torch = h.torch
    N, D, T = 1 << 22, 2048, 100
    data = nb.Dataset.synth(ctx, nb.SYNTH_MULTICLASS, N, D, T=T, seed=SEED)
    function = nb.LinearFunction(data, nb.Loss("s-classnll"), 0.0, L2REG)
    function.set_stream(torch.cuda.current_stream().cuda_stream)
    n = function.size()
    x = np.random.default_rng(0).uniform(-1.0, 1.0, n) / np.sqrt(D)
    x_dev = torch.from_numpy(x).cuda()
    g_dev = torch.empty(n, dtype=torch.float64, device="cuda")
    f_dev = torch.empty(1, dtype=torch.float64, device="cuda")

    def step():
        function.vgrad_device(x_dev.data_ptr(), g_dev.data_ptr(), f_dev.data_ptr())

    burst = h.region(step, function, steps, warmup)
    kernel = summary(burst["passes"])
    flops = 4.0 * N * D * T
    achieved = flops / (kernel["mean"] * 1e-3) / 1e12
    entry = {
        "workload": "s-classnll (softmax cross-entropy) linear model, 2^22 samples x 2048 features x 100 classes, fp64 vgrad",
        "metric": "vgrad ms (and TFLOP/s of 4*N*D*T)", "value": burst["ms"] / steps, "unit": "ms", "steps": steps,
        "higher_is_better": False, "samples_features_per_s": N * D * steps / (burst["ms"] * 1e-3),
        "kernel": function.describe(),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["dmma_tflops"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["dmma_tflops"], "traffic": None, "kernel_ms": kernel,
                     "algorithmic_flops_per_launch": flops,
                     "peak_source": "in-run probe of the fp64 tensor pipe (nb200_probe_peak: mma.sync m8n8k4 f64 issue "
                                    "rate); MEASURED_PEAKS.json holds no fp64 figure",
                     "frac_of_cublas_dgemm": achieved / peaks["cublas_dgemm_8192_tflops"],
                     "pass": "pack W + forward (X W^T, loss, dL) + backward (dL^T X): X is read twice"},
        "clocks": h.sampler.window(*burst["window"]), "f": float(f_dev.item()),
    }
    sample_rows = 4096
    Xs, Ys = data.read(0, sample_rows)
    small = nb.LinearFunction(nb.Dataset.pin(ctx, Xs, Ys), nb.Loss("s-classnll"), 0.0, L2REG)
    gx = np.empty(n)
    fx = small(x, gx)
    entry["parity"] = dict(oracle_parity(oracle, Xs, Ys, "s-classnll", x, 0.0, L2REG, fx, gx), rows=sample_rows,
                           same_plan=small.describe().split(" ")[0] == function.describe().split(" ")[0])
    small.close()
    function.close()
    data.close()
    return entry


def run_b200(args):
    import torch
    import torch.distributed as dist

    import libnano_b200 as nb
    from libnano_b200.sharded import ShardedFunction, balanced_rows

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
    sampler = ClockSampler(local_rank).start()  # every rank watches its own board
    h = Harness(args, torch, dist, world, rank, sampler)

    rows, features = args.rows_per_gpu, args.features
    n_global = rows * world
    ctx = nb.Context(local_rank)

    def sharded_objective(rows_local, row_offset, n_total, seed, want_peers):
        data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, rows_local, features, seed=seed, row_offset=row_offset,
                                noise=0.1)
        local = nb.LinearFunction(data, nb.Loss(LOSS_ID), 0.0, L2REG)
        if args.variant >= 0:
            local.set_variant(args.variant)
        function = ShardedFunction.from_local(local, n_total)  # re-points `local` at torch's current stream
        how = "none"
        if world > 1:
            how = "nccl"
            if want_peers:
                try:
                    if function.connect_peers():
                        how = "peer"
                except Exception as error:  # noqa: BLE001 - fall back to the NCCL path, say so in the JSON line
                    if args.collective == "peer":
                        raise
                    how = f"nccl (peer setup failed: {error})"
            flags = [None] * world
            dist.all_gather_object(flags, how == "peer")
            if not all(flags):  # every rank must take the same path
                function._shard.fused = False
                how = how if how != "peer" else "nccl"
        return data, local, function, how

    data, local, function, collective = sharded_objective(rows, rank * rows, n_global, SEED,
                                                          args.collective in ("auto", "peer"))
    n = function.size()
    x = np.random.default_rng(0).uniform(-1.0, 1.0, n) / np.sqrt(features)
    gx = np.empty(n)

    # ---- (1) device-resident loop: x, g, f stay in HBM -------------------------------------------------------
    x_dev = torch.from_numpy(x).cuda()
    g_dev = torch.empty(n, dtype=torch.float64, device="cuda")
    f_dev = torch.empty(1, dtype=torch.float64, device="cuda")

    def make_step(local_, function_, n_total):
        def step_device():
            if world == 1:
                # one cooperative launch: pass over X, cross-CTA reduction, scaling + regularisation
                local_.vgrad_device(x_dev.data_ptr(), g_dev.data_ptr(), f_dev.data_ptr())
            elif collective == "peer":
                # the launch streams X AND publishes its sums into every peer's mailbox over NVLink; no NCCL call
                local_.vgrad_sharded_device(x_dev.data_ptr(), n_total, g_dev.data_ptr(), f_dev.data_ptr())
            else:
                ptr, length = local_.partial_device(x_dev.data_ptr(), True)
                dist.all_reduce(function_._shard._view_for(ptr, length))
                local_.finish_device(x_dev.data_ptr(), n_total, g_dev.data_ptr(), f_dev.data_ptr())
        return step_device

    step_device = make_step(local, function, n_global)

    # ---- (0) N > 1: size the row blocks by each GPU's measured speed ----------------------------------------------
    # the boards sit at the power cap at different clocks (their kernel times are 1.3 % apart on an 8 x B200 box) and a
    # sharded evaluation is as slow as its slowest rank; the partition is re-cut ONCE, before the timed regions, from a
    # calibration region identical to the timed one, and stays static afterwards (libnano_b200.sharded.balanced_rows)
    partition = {"rows": [rows] * world, "balanced": False}
    if world > 1 and not args.no_balance:
        calibration = h.region(step_device, local, args.steps, args.warmup)
        mine_ms = float(np.median(calibration["passes"]))
        measured = [None] * world
        dist.all_gather_object(measured, mine_ms)
        row0s, new_rows = balanced_rows(n_global, [ms / rows for ms in measured])
        partition = {"rows": new_rows, "balanced": True, "calibration_kernel_ms": measured}
        if max(abs(r - rows) for r in new_rows) > 1e-4 * rows:
            local.close()
            data.close()
            del function, local, data
            data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, new_rows[rank], features, seed=SEED,
                                    row_offset=row0s[rank], noise=0.1)
            local = nb.LinearFunction(data, nb.Loss(LOSS_ID), 0.0, L2REG)
            if args.variant >= 0:
                local.set_variant(args.variant)
            function = ShardedFunction.from_local(local, n_global)
            if collective == "peer":
                connected = [None] * world
                ok = False
                try:
                    ok = function.connect_peers()
                except Exception:  # noqa: BLE001
                    ok = False
                dist.all_gather_object(connected, ok)
                if not all(connected):
                    raise SystemExit("bench.py: the peers could not be re-connected after re-cutting the row blocks")
            step_device = make_step(local, function, n_global)
    my_rows = partition["rows"][rank]
    main = h.region(step_device, local, args.steps, args.warmup)
    f_device = float(f_dev.item())
    clocks = sampler.window(*main["window"])

    # ---- (2) end-to-end loop: the public call with HOST buffers ---------------------------------------------------
    holder = {}

    def step_host():
        holder["f"] = function(x, gx)

    e2e = h.region(step_host, local, args.steps, args.warmup)
    f_host = holder["f"]
    e2e_clocks = sampler.window(*e2e["window"])

    # ---- (3) steady state ------------------------------------------------------------------------------------------
    sustained = h.sustained(step_device, args.sustained_seconds)

    if not np.isfinite(f_host) or abs(f_host - f_device) > 1e-12 * (1 + abs(f_host)):
        raise SystemExit(f"bench.py: device-resident and host-buffer evaluations disagree: {f_device} vs {f_host}")
    kernel_text = local.describe()
    peer_mode = local.peer_mode()

    # ---- (4) per-rank evidence (what limits the 1 -> N curve) --------------------------------------------------------
    mine = {"rank": rank, "rows": my_rows, "kernel_ms": summary(main["passes"]), "step_ms": main["own_ms"] / args.steps,
            "clocks": clocks}
    per_rank = [mine]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)

    # ---- (5) parity in this run: a second, small dataset through the SAME path against the oracle ------------------
    local.close()
    data.close()
    del function, local, data
    small_rows = max(2, args.parity_rows // world)
    small_total = small_rows * world
    p_data, p_local, p_function, p_collective = sharded_objective(small_rows, rank * small_rows, small_total, SEED + 1,
                                                                  collective == "peer")
    gp = np.empty(n)
    fp = p_function(x, gp)
    parity = None
    if rank == 0:
        import oracle

        oracle.build()
        whole = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, small_total, features, seed=SEED + 1, noise=0.1)
        Xp, Yp = whole.read()
        whole.close()
        parity = dict(oracle_parity(oracle, Xp, Yp, LOSS_ID, x, 0.0, L2REG, fp, gp), rows=small_total,
                      features=features, ranks=world, collective=p_collective,
                      path={"none": "nb200_vgrad", "peer": "nb200_vgrad_sharded"}.get(p_collective, "partial / all_reduce / finish"))
    verdict = [parity["ok"] if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(verdict, src=0)
    p_local.close()
    p_data.close()
    if not verdict[0]:
        if rank == 0:
            print(json.dumps({"error": "parity against the oracle failed", "parity": parity}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return 1

    if rank != 0:
        sampler.stop()
        if world > 1:
            dist.destroy_process_group()
        return 0

    work = float(n_global) * features * args.steps  # samples*features over all ranks
    kernel = summary(main["passes"])
    bytes_alg = algorithmic_bytes(my_rows, features)  # rank 0's launch
    peak, peak_source = measured_hbm_peak()
    achieved = bytes_alg / (kernel["mean"] * 1e-3) / 1e9
    config = workload_config(args, world)
    line = {
        "metric": METRIC, "value": work / (main["ms"] * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main["ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "details": {
            "kernel": kernel_text,
            "collective": {"peer": "inside the launch: reduced column slices stored into every peer's mailbox over NVLink; "
                                   + ("the stream waits for the peers' ready words (cuStreamWaitValue64), a finish kernel "
                                      "folds the slots in rank order" if peer_mode == 0 else
                                      "the launch waits for the peers' flags and folds the slots in rank order"),
                           "none": "none (single GPU)"}.get(collective, collective),
            "settle": f"{args.settle_seconds} s idle before the warm-up of each timed region, so every region starts "
                      "from the same board state",
            "partition": partition,
        },
        "e2e": {"value": work / (e2e["ms"] * 1e-3), "unit": UNIT, "ms_per_step": e2e["ms"] / args.steps,
                "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * (n + 1),
                "api": "libnano_b200.ShardedFunction.__call__(x, gx) (host buffers) -> " +
                       {"none": "nb200_vgrad", "peer": "nb200_vgrad_sharded"}.get(collective,
                                                                                 "nb200_vgrad_partial / finish"),
                "clocks": e2e_clocks},
        "gpu_launches": int(main["launches"]),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic(kernel_text, my_rows, features), "kernel": "vgrad_t1_kernel",
                     "kernel_ms": kernel["mean"], "kernel_ms_distribution": kernel, "launches_timed": kernel["launches"],
                     "algorithmic_bytes_per_launch": bytes_alg, "peak_source": peak_source},
        "clocks": clocks,
        "parity": parity,
        "f": f_host,
    }
    if sustained is not None:
        window = sustained.pop("window")
        sustained["achieved_gbs"] = bytes_alg / (sustained["ms_per_step_steady"] * 1e-3) / 1e9
        sustained["frac"] = sustained["achieved_gbs"] / peak
        sustained["value"] = float(n_global) * features / (sustained["ms_per_step_steady"] * 1e-3)
        sustained["clocks"] = sampler.window(*window)
        line["sustained"] = sustained
    if world > 1:
        slowest = max(r["kernel_ms"]["median"] for r in per_rank)
        fastest = min(r["kernel_ms"]["median"] for r in per_rank)
        line["per_rank"] = per_rank
        line["scaling_limiter"] = {
            "kernel_ms_median_fastest_rank": fastest, "kernel_ms_median_slowest_rank": slowest,
            "straggler_spread_ms": slowest - fastest,
            "exchange_ms_beyond_slowest_kernel": main["ms"] / args.steps - slowest,
            "reading": "step = the slowest rank's pass over X + what the exchange adds; the spread between the ranks' "
                       "median kernel times is straggling (boards at the power cap run at different clocks), the "
                       "remainder is the publish / wait / fold of the exchange",
        }

    # ---- (6) N = 1: the other BASELINE configs, each with its own roofline, clocks and parity ------------------------
    if world == 1 and not args.no_configs:
        configs = {}
        peaks = fp64_peaks(h, ctx)
        line["peaks_in_run"] = dict(peaks, note="hbm_read_gbs: read-only stream through 1-D bulk copies (the single-"
                                                "target pass only reads); roofline.peak stays the driver's copy figure")
        configs["cfg1"] = bench_config1(h, nb, ctx)
        entry2, data2, function2, x2 = bench_streaming_config(
            h, nb, ctx, "cfg2", LOSS_ID, 1 << 20, 1024, 0.0, L2REG, peak, peak_source, args.steps, args.warmup,
            keep_dataset=True)
        if not args.no_cpu_baseline:
            # the CPU port on the FULL config-2 dataset (the very rows the GPU holds), and parity at the full shape
            import oracle

            cpu_rows = min(args.cpu_rows, 1 << 20)
            Xc, Yc = data2.read(0, cpu_rows)
            times, threads = time_cpu(Xc, Yc, args.cpu_calls)
            line["cpu_baseline"] = cpu_baseline_entry(times, threads, cpu_rows, 1024,
                                                      "the config-2 dataset read back from HBM (configs.cfg2)")
            if cpu_rows == 1 << 20:
                g2 = np.empty(function2.size())
                f2 = function2(x2, g2)
                entry2["parity_full_shape"] = dict(oracle_parity(oracle, Xc, Yc, LOSS_ID, x2, 0.0, L2REG, f2, g2),
                                                   rows=cpu_rows)
            del Xc, Yc
        function2.close()
        data2.close()
        configs["cfg2"] = entry2
        configs["cfg3"] = bench_config3(h, nb, ctx, peaks)
        configs["cfg5"] = bench_streaming_config(h, nb, ctx, "cfg5", "s-hinge", 1 << 24, 512, 1e-2, 0.0, peak,
                                                 peak_source, args.steps, args.warmup)
        configs["handful_of_targets"] = bench_handful_of_targets(h, nb, ctx, peaks, peak)
        line["configs"] = configs
        bad = [name for name, entry in configs.items()
               if not all(case.get("parity", {}).get("ok", True) for case in entry.get("cases", [entry]))
               or not entry.get("parity_full_shape", {}).get("ok", True)]
        if bad:
            print(json.dumps({"error": f"parity against the oracle failed in {bad}", "configs": configs}), flush=True)
            return 1
    elif world == 1 and not args.no_cpu_baseline:
        Xc, Yc = host_sample(min(args.cpu_rows, rows), features)
        times, threads = time_cpu(Xc, Yc, args.cpu_calls)
        line["cpu_baseline"] = cpu_baseline_entry(times, threads, Xc.shape[0], features, "host-generated sample")
    sampler.stop()
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
