"""N > 1 path on CPU: world_size-2 (and 3) `gloo` process groups drive libnano_b200.sharded.ShardedFunction — the
same host logic that runs over NCCL on GPUs — with a CPU double for the rank-local evaluator (the oracle's
partial / finish halves), and compare with the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import make_problem, rel_f, rel_g


class OracleShard:
    """Test double of DeviceShard: rank-local sums from the oracle, buffer as a CPU torch tensor."""

    def __init__(self, oracle, X, Y, loss_id, l1, l2):
        self.oracle, self.X, self.Y, self.loss_id, self.l1, self.l2 = oracle, X, Y, loss_id, l1, l2
        self.D, self.T = X.shape[1], Y.shape[1]
        self.buffer = torch.zeros(1 + self.T + self.T * self.D, dtype=torch.float64)
        self.x = None

    def size(self):
        return (self.D + 1) * self.T

    def partial(self, x, want_grad):
        self.x = x.copy()
        partial = self.oracle.linear_partial(self.X, self.Y, self.loss_id, x, want_grad=want_grad)
        self.buffer.copy_(torch.from_numpy(partial))
        return self.buffer if want_grad else self.buffer[: 1 + self.T]

    def finish(self, n_global, gx):
        fx, g = self.oracle.linear_finish(self.buffer.numpy(), n_global, self.D, self.T, self.x, self.l1, self.l2,
                                          want_grad=gx is not None)
        if gx is not None:
            gx[:] = g
        return fx


def _worker(rank, world, port, N, D, T, loss_id, out):
    import oracle
    from libnano_b200.sharded import ShardedFunction, broadcast_x, shard_rows

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, Y, x = make_problem(77, N, D, T, loss_id != "mse")
        l1, l2 = 0.05, 0.2
        row0, rows = shard_rows(N, world, rank)
        shard = OracleShard(oracle, X[row0:row0 + rows], Y[row0:row0 + rows], loss_id, l1, l2)
        function = ShardedFunction(shard, N, convex=True, smooth=False, strong_convexity=l2 / (D * T))
        # rank 0 owns the solver state; everyone evaluates the same x
        x = broadcast_x(x if rank == 0 else np.zeros_like(x), src=0)
        gx = np.empty_like(x)
        fx = function(x, gx)
        fv = function(x)
        f_ref, g_ref = oracle.linear_vgrad(X, Y, loss_id, x, l1, l2)
        ok = rel_f(fx, f_ref) <= 1e-13 and rel_g(gx, g_ref) <= 1e-13 and rel_f(fv, f_ref) <= 1e-13
        ok = ok and (function.fcalls(), function.gcalls()) == (2, 1) and function.world_size() == world
        # every rank ends with identical bits (the all-reduced vector is identical, phase 2 is deterministic)
        gathered = [torch.zeros(x.size + 1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(np.concatenate([gx, [fx]])))
        ok = ok and all(torch.equal(gathered[0], g) for g in gathered)
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,N,D,T,loss_id", [(2, 1001, 37, 1, "s-logistic"), (2, 400, 16, 3, "s-classnll"),
                                                 (3, 500, 8, 1, "mse")])
def test_sharded_function_over_gloo(oracle, world, N, D, T, loss_id):
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, D, T, loss_id, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert list(out) == [1] * world


def _solver_worker(rank, world, port, N, D, solver_id, out):
    """The host-driven non-smooth drivers over a sharded objective: every rank runs the same host algebra on the same
    all-reduced (f, g), so the ranks stay in lock-step without any further exchange."""
    import oracle
    from libnano_b200 import solver as nb_solver
    from libnano_b200.sharded import ShardedFunction, shard_rows
    from test_host_solvers import OracleFunction, informative_problem

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, Y = informative_problem(4, N, D, True)
        l1, l2 = 1e-2, 1e-2
        row0, rows = shard_rows(N, world, rank)
        shard = OracleShard(oracle, X[row0:row0 + rows], Y[row0:row0 + rows], "s-hinge", l1, l2)
        function = ShardedFunction(shard, N, convex=True, smooth=False, strong_convexity=l2 / D)
        x0 = np.zeros(D + 1)
        if solver_id == "rqb":
            result = nb_solver.rqb(function, x0, epsilon=1e-8, max_evals=4000)
            whole = nb_solver.rqb(OracleFunction(oracle, X, Y, "s-hinge", l1=l1, l2=l2), x0, epsilon=1e-8, max_evals=4000)
            ok = result["status"] == "specific_test" and abs(result["fx"] - whole["fx"]) <= 1e-7 * (1 + abs(whole["fx"]))
        else:
            result = nb_solver.osga(function, x0, epsilon=1e-6, max_evals=600)
            ok = result["fx"] < function(x0) and result["iterations"] > 10
        # lock-step: identical iterates and call counts on every rank
        packed = np.concatenate([result["x"], [result["fx"], result["fcalls"], result["iterations"]]])
        gathered = [torch.zeros(packed.size, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(packed))
        ok = ok and all(torch.equal(gathered[0], g) for g in gathered)
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("solver_id", ["rqb", "osga"])
def test_host_driven_solvers_stay_in_lock_step_over_gloo(oracle, solver_id):
    # BASELINE config 5 sharded: hinge + elastic net over 2 ranks, the drivers replicated on every rank
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    port = _free_port()
    procs = [ctx.Process(target=_solver_worker, args=(r, world, port, 600, 8, solver_id, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert list(out) == [1] * world


def test_balanced_rows_follow_the_measured_speeds():
    from libnano_b200.sharded import balanced_rows

    row0s, rows = balanced_rows(1 << 26, [9.4255, 9.5263, 9.5089, 9.5467, 9.4695, 9.4956, 9.4408, 9.5233])
    assert sum(rows) == 1 << 26 and row0s[0] == 0
    assert all(a + b == c for a, b, c in zip(row0s, rows, row0s[1:]))
    assert rows[0] == max(rows) and rows[3] == min(rows)              # the fastest rank holds the most rows
    times = [r * t for r, t in zip(rows, [9.4255, 9.5263, 9.5089, 9.5467, 9.4695, 9.4956, 9.4408, 9.5233])]
    assert max(times) / min(times) - 1.0 < 1e-6                        # everyone finishes together
    assert balanced_rows(10, [1.0, 1.0, 1.0]) == ([0, 4, 7], [4, 3, 3])
    assert balanced_rows(3, [1.0, 1e-9, 1.0])[1] == [1, 1, 1]
    with pytest.raises(RuntimeError, match="positive and finite"):
        balanced_rows(100, [1.0, 0.0])
    with pytest.raises(RuntimeError, match="at least one sample"):
        balanced_rows(1, [1.0, 1.0])
