"""Sample-sharded objective: one process per GPU, rows of (X, Y) split in contiguous blocks, ONE all-reduce of the
(loss, gradient) vector per evaluation.

The objective is a sum over independent samples; the reference already shards it — across the threads of one
process: per-thread accumulators (src/linear/function.cpp:51-57) folded by sum_reduce (include/nano/core/reduce.h:
23-31). Here each rank plays the role of one accumulator:

    phase 1  local pass over the rank's rows      -> un-normalised [sum loss | sum dL | sum dL x^T] in HBM
    all-reduce (sum) over ranks                    (torch.distributed: NCCL over NVLink on GPUs, gloo in CPU tests)
    phase 2  / N_global + regularisation           -> (f, g), identical on every rank

No other collective exists on this path: W and b are replicated (<= 1.6 MB), the regularisers are O(T*D).
"""
from __future__ import annotations

import numpy as np

from .function import Function, critical


def shard_rows(n_global: int, world_size: int, rank: int):
    """Contiguous, balanced row blocks: (row0, rows) of `rank`; the first n_global % world_size ranks hold one more."""
    critical(world_size > 0 and 0 <= rank < world_size, "sharding: invalid rank ", rank, " of ", world_size)
    base, extra = divmod(int(n_global), int(world_size))
    rows = base + (1 if rank < extra else 0)
    row0 = rank * base + min(rank, extra)
    return row0, rows


def balanced_rows(n_global: int, seconds_per_row):
    """Row counts proportional to each rank's measured speed (1 / seconds_per_row), summing to n_global exactly.

    The ranks of a sample-sharded evaluation finish together only if their shards match their speed: boards that sit
    at the power cap run at different clocks (1.3 % apart on an 8 x B200 box, profiles/r02_bench_8gpu.json), and the
    evaluation is as slow as its slowest rank. The partition stays STATIC — contiguous blocks in rank order, sizes
    fixed before the first evaluation — so results are reproducible for a given partition. Returns (row0s, rows)."""
    speeds = 1.0 / np.asarray(seconds_per_row, dtype=np.float64)
    critical(speeds.ndim == 1 and speeds.size > 0 and bool(np.all(np.isfinite(speeds))) and bool(np.all(speeds > 0)),
             "sharding: the measured times must be positive and finite")
    critical(n_global >= speeds.size, "sharding: every rank needs at least one sample")
    ideal = n_global * speeds / speeds.sum()
    rows = np.maximum(np.floor(ideal).astype(np.int64), 1)
    # hand the remainder out (or take an excess back) where the rounding error is largest; fixed tie-break by rank
    while rows.sum() != n_global:
        error = ideal - rows
        if rows.sum() < n_global:
            rows[int(np.argmax(error))] += 1
        else:
            candidates = np.where(rows > 1, error, np.inf)
            rows[int(np.argmin(candidates))] -= 1
    row0s = np.concatenate([[0], np.cumsum(rows)[:-1]])
    return [int(v) for v in row0s], [int(v) for v in rows]


class _CudaBufferView:
    """Zero-copy view of a device buffer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, length: int):
        self.__cuda_array_interface__ = {
            "shape": (int(length),),
            "typestr": "<f8",
            "data": (int(ptr), False),
            "version": 2,
            "strides": None,
        }


class DeviceShard:
    """Adapter of a LinearFunction / SyntheticFunction (the rank's local objective) to the sharded protocol.

    The local objective is re-pointed at torch's current CUDA stream so that phase 1, the NCCL all-reduce and phase 2
    are ordered on ONE stream without host synchronisation in between."""

    def __init__(self, local):
        import torch

        self.local = local
        self._torch = torch
        self._stream = torch.cuda.current_stream()
        local.set_stream(self._stream.cuda_stream)
        self._view = None
        self.fused = False  # True once connect_peers() mapped the peers' mailboxes

    def size(self) -> int:
        return self.local.size()

    def _view_for(self, ptr: int, length: int):
        """torch tensor aliasing the objective's reduced-sums buffer (cached: the buffer never moves)."""
        if self._view is None or self._view[0] != (ptr, length):
            tensor = self._torch.as_tensor(_CudaBufferView(ptr, length), device=self._torch.device("cuda"))
            self._view = ((ptr, length), tensor)
        return self._view[1]

    def partial(self, x, want_grad: bool):
        ptr, length = self.local.partial(x, want_grad)
        view = self._view_for(ptr, length)
        # value-only evaluations leave the gradient entries untouched: reduce [loss | gb] only
        return view if want_grad else view[: 1 + self.local._dataset.targets]

    def finish(self, n_global: int, gx):
        return self.local.finish(n_global, gx)

    def eval(self, x, gx):
        """Single-rank shortcut: the fused one-launch evaluation (no all-reduce to wait for)."""
        return self.local.do_eval(x, gx)

    def connect_peers(self, group=None) -> bool:
        """Set up the fused compute + collective path (connect_local_peers); False leaves the NCCL two-phase path."""
        self.fused = connect_local_peers(self.local, group)
        return self.fused


def connect_local_peers(local, group=None) -> bool:
    """Set up the fused compute + collective path of a rank-local objective: every rank maps every other rank's
    mailbox (CUDA IPC over NVLink), after which ONE kernel launch per rank does the pass over X AND publishes its sums
    into the peers' mailboxes. Returns False (nothing changed) when the objective is not on the single-target kernel or
    any rank failed. Collective: every rank of `group` must call it, in the same order for every objective. All ranks
    leave through a barrier, so no rank evaluates before every mailbox is mapped."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world < 2 or world > 8:
        return False

    def everyone(flag: bool) -> bool:
        """Every rank takes part in every collective, whatever happened locally: no rank is left waiting."""
        mine = torch.tensor([1 if flag else 0], dtype=torch.uint8, device="cuda")
        votes = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(votes, mine, group=group)
        return all(int(v.item()) == 1 for v in votes)

    # (1) export this rank's mailbox (fails, e.g., when the objective is not on the single-target kernel)
    handle, grid, ok = bytes(64), 0, True
    try:
        if not local.describe().startswith("t1"):
            raise RuntimeError("not the single-target kernel")
        handle, grid = local.peer_export(world)
    except Exception:  # noqa: BLE001 - reported through the vote below
        ok = False
    mine = torch.tensor(list(handle) + [grid % 256, grid // 256, int(ok)], dtype=torch.uint8, device="cuda")
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    blobs = [bytes(t.cpu().tolist()) for t in gathered]
    if not all(b[66] == 1 for b in blobs) or not all(b[64:66] == blobs[0][64:66] for b in blobs):
        return False  # some rank could not export, or the ranks run different grids: keep the NCCL path

    # (2) map the peers' mailboxes (CUDA IPC); only if every rank succeeded is the fused path switched on
    try:
        local.peer_connect(rank, world, b"".join(b[:64] for b in blobs))
    except Exception:  # noqa: BLE001
        ok = False
    fused = everyone(ok)
    dist.barrier(group=group)
    return fused


class ShardedFunction(Function):
    """function_t over ALL samples of a dataset whose rows are spread over the ranks of a process group.

    `shard` is the rank-local evaluator (DeviceShard in production; tests drive the same logic with a CPU double
    over gloo). Every rank must call the function with the same x (the solver state is replicated)."""

    def __init__(self, shard, n_global: int, group=None, type_id: str = "linear", convex=False, smooth=False,
                 strong_convexity: float = 0.0):
        import torch.distributed as dist

        super().__init__(type_id, shard.size())
        critical(n_global > 0, "sharded function: the global sample count must be positive")
        self._shard = shard
        self._n_global = int(n_global)
        self._group = group
        self._dist = dist
        self._convex, self._smooth, self._strong_convexity = bool(convex), bool(smooth), float(strong_convexity)

    @classmethod
    def from_local(cls, local, n_global: int, group=None) -> "ShardedFunction":
        return cls(DeviceShard(local), n_global, group, local.type_id(), local.convex(), local.smooth(),
                   local.strong_convexity())

    def world_size(self) -> int:
        return self._dist.get_world_size(self._group) if self._dist.is_initialized() else 1

    def connect_peers(self) -> bool:
        """Switch to the fused compute + collective kernel (peer memory over NVLink) when the shard supports it."""
        return bool(getattr(self._shard, "connect_peers", lambda group=None: False)(self._group))

    def do_eval(self, x, gx) -> float:
        if self.world_size() == 1 and hasattr(self._shard, "eval"):
            return self._shard.eval(x, gx)
        if getattr(self._shard, "fused", False):
            return self._shard.local.vgrad_sharded(x, self._n_global, gx)
        buffer = self._shard.partial(x, gx is not None)
        if self.world_size() > 1:
            self._dist.all_reduce(buffer, op=self._dist.ReduceOp.SUM, group=self._group)
        return self._shard.finish(self._n_global, gx)


def broadcast_x(x: np.ndarray, src: int = 0, group=None) -> np.ndarray:
    """Replicate the solver state from `src` (used once at start; iterates stay identical because (f, g) are)."""
    import torch
    import torch.distributed as dist

    tensor = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl":
            tensor = tensor.cuda()
            dist.broadcast(tensor, src=src, group=group)
            tensor = tensor.cpu()
        else:
            dist.broadcast(tensor, src=src, group=group)
    return tensor.numpy()
