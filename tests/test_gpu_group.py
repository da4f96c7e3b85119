"""GPU tests of the device group (one process, several devices behind the C ABI: nb200_init_multi) and of concurrent
callers.

libnano is one process that shards linear::function_t::do_eval over the workers of its thread pool
(src/linear/function.cpp:53-57, per-worker accumulators folded by sum_reduce, include/nano/core/reduce.h:23-31) and
evaluates folds x trials concurrently on separate function objects (src/machine/tune.cpp:25-42, clones per thread in
app/bench_solver.cpp:296). A group context plays the pool's part with devices. A group may name the same device more
than once: the per-device sums then travel by copies on that device (the two-phase path), which is how a one-GPU box
covers the group logic; with two or more GPUs the fused peer-memory path and its bit-equality with the two-phase path
are checked as well, and the torchrun checks of tools/check_sharded.py run as a sub-process.
"""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from conftest import PARITY, make_problem, rel_f, rel_g

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_count():
    from libnano_b200 import _lib

    return int(_lib.lib().nb200_device_count())


def group_devices(parts):
    """`parts` devices for a group: distinct ones when the box has them, else device 0 repeated."""
    count = device_count()
    return list(range(parts)) if count >= parts else [0] * parts


@pytest.mark.parametrize("parts", [1, 2, 3])
@pytest.mark.parametrize("N,D,T,loss_id,l1,l2", [
    (5001, 1024, 1, "s-logistic", 0.0, 1e-2),   # the fused single-target kernel on every device
    (4097, 512, 1, "s-hinge", 1e-2, 1e-2),      # config 5's objective
    (3000, 784, 10, "s-classnll", 0.0, 1e-2),   # one-pass tensor-pipe kernel
    (2500, 96, 5, "s-classnll", 0.0, 0.0),      # two-pass tensor-pipe pair
    (1999, 33, 3, "mse", 0.0, 1e-3),            # shape-generic kernels
])
def test_group_objective_matches_the_oracle(oracle, parts, N, D, T, loss_id, l1, l2):
    import libnano_b200 as nb

    X, Y, x = make_problem(11, N, D, T, classification=loss_id != "mse")
    group = nb.Context(group_devices(parts))
    data = nb.Dataset.pin(group, X, Y)
    assert (data.samples, data.inputs, data.targets) == (N, D, T)
    Xb, Yb = data.read(N // 3, N // 2)  # a range that crosses the devices' row blocks
    assert np.array_equal(Xb, X[N // 3:N // 3 + N // 2]) and np.array_equal(Yb, Y[N // 3:N // 3 + N // 2])

    function = nb.LinearFunction(data, nb.Loss(loss_id), l1, l2)
    assert function.parts() == parts and function.size() == (D + 1) * T
    gx = np.empty_like(x)
    fx = function(x, gx)
    f_ref, g_ref = oracle.linear_vgrad(X, Y, loss_id, x, l1=l1, l2=l2)
    assert rel_f(fx, f_ref) <= PARITY and rel_g(gx, g_ref) <= PARITY
    assert rel_f(function(x), f_ref) <= PARITY          # value-only call
    g2 = np.empty_like(x)
    assert function(x, g2) == fx and np.array_equal(g2, gx)  # fixed summation order: same bits every time
    assert function.launches() >= parts
    # the same partition evaluated by one objective per block + a host-side fold in block order gives the same bits
    clone = function.clone()
    g3 = np.empty_like(x)
    assert clone(x, g3) == fx and np.array_equal(g3, gx)
    for i in range(parts):
        text, _, _, launches = function.part_stats(i)
        assert launches > 0 and text
    group.close()


def test_group_equals_single_device_to_rounding(ctx):
    import libnano_b200 as nb

    N, D = 20000, 1024
    whole = nb.LinearFunction(nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, N, D, seed=7), nb.Loss("s-logistic"), 0.0, 1e-2)
    group = nb.Context(group_devices(2))
    data = nb.Dataset.synth(group, nb.SYNTH_CLASSIFICATION, N, D, seed=7)  # the same rows, generated per device
    Xw, Yw = whole._dataset.read()
    Xg, Yg = data.read()
    assert np.array_equal(Xw, Xg) and np.array_equal(Yw, Yg)
    split = nb.LinearFunction(data, nb.Loss("s-logistic"), 0.0, 1e-2)
    x = np.random.default_rng(3).uniform(-1, 1, D + 1) / np.sqrt(D)
    g1, g2 = np.empty_like(x), np.empty_like(x)
    f1, f2 = whole(x, g1), split(x, g2)
    assert rel_f(f2, f1) <= 1e-13 and rel_g(g2, g1) <= 1e-13   # shard invariance
    group.close()


def test_group_scaling_predict_evaluate_and_solve(ctx, oracle):
    import libnano_b200 as nb
    from libnano_b200.solver import lbfgs

    N, D = 6000, 64
    rng = np.random.default_rng(5)
    X = rng.normal(3.0, 2.0, (N, D))
    X[rng.uniform(size=X.shape) < 0.01] = np.nan        # missing values
    w = rng.normal(size=D)
    Y = np.where(np.nan_to_num(X) @ w > np.median(np.nan_to_num(X) @ w), 1.0, -1.0).reshape(N, 1)

    single = nb.Dataset.pin(ctx, X, Y).scale("standard", "none")
    group_ctx = nb.Context(group_devices(3))
    split = nb.Dataset.pin(group_ctx, X, Y).scale("standard", "none")
    for name, values in single.stats("inputs").items():
        assert np.allclose(split.stats("inputs")[name], values, rtol=1e-12, atol=1e-12), name
    Xs, _ = single.read()
    Xg, _ = split.read()
    assert np.allclose(Xs, Xg, rtol=1e-12, atol=1e-12) and np.all(np.isfinite(Xg))

    W, b = rng.normal(size=(1, D)), rng.normal(size=1)
    assert np.allclose(split.predict(W, b), single.predict(W, b), rtol=1e-12, atol=1e-12)
    loss = nb.Loss("s-logistic")
    assert np.allclose(split.evaluate(loss, W, b), single.evaluate(loss, W, b), rtol=1e-12, atol=1e-12)

    # the device-resident L-BFGS: solver state on the first device, every evaluation on all devices
    f_single = nb.LinearFunction(single, loss, 0.0, 1e-2)
    f_group = nb.LinearFunction(split, loss, 0.0, 1e-2)
    a = lbfgs(f_single, np.zeros(D + 1), epsilon=1e-9)
    c = lbfgs(f_group, np.zeros(D + 1), epsilon=1e-9)
    assert a["status"] == c["status"] == "gradient_test"
    assert abs(a["fx"] - c["fx"]) <= 1e-10 * (1 + abs(a["fx"])) and np.max(np.abs(a["x"] - c["x"])) <= 1e-5
    group_ctx.close()


def test_group_rejects_the_cross_process_protocol(ctx):
    import libnano_b200 as nb

    X, Y, x = make_problem(2, 600, 32, 1, True)
    group = nb.Context(group_devices(2))
    function = nb.LinearFunction(nb.Dataset.pin(group, X, Y), nb.Loss("s-logistic"))
    with pytest.raises(nb.Nb200Error, match="not available on a device group"):
        function.partial(x)
    with pytest.raises(nb.Nb200Error, match="not available on a device group"):
        function.peer_export(2)
    with pytest.raises(nb.Nb200Error, match="at least one sample"):
        nb.Dataset.pin(nb.Context(group_devices(3)), X[:2], Y[:2])
    group.close()


@pytest.mark.skipif("device_count() < 2")
def test_fused_group_is_bit_identical_to_the_two_phase_group(oracle, monkeypatch):
    """Two real devices: the launches publish their sums into the first device's mailbox over NVLink; the fold order
    is the device order in both paths, so the bits agree."""
    import libnano_b200 as nb

    world = min(device_count(), 8)
    N, D = 300001, 1024
    results = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("NB200_GROUP_FUSED", fused)
        group = nb.Context(list(range(world)))
        data = nb.Dataset.synth(group, nb.SYNTH_CLASSIFICATION, N, D, seed=7)
        function = nb.LinearFunction(data, nb.Loss("s-logistic"), 0.0, 1e-2)
        assert ("two-phase" in function.describe()) == (fused == "0"), function.describe()
        x = np.random.default_rng(1).uniform(-1, 1, D + 1) / np.sqrt(D)
        gx = np.empty_like(x)
        fx = function(x, gx)
        for _ in range(5):  # mailbox parity / epochs
            g2 = np.empty_like(x)
            assert function(x, g2) == fx and np.array_equal(g2, gx)
        fv = function(x)
        results[fused] = (fx, gx.copy(), fv)
        if fused == "1":
            X, Y = data.read()
            f_ref, g_ref = oracle.linear_vgrad(X, Y, "s-logistic", x, l2=1e-2, threads=oracle.hardware_threads())
            assert rel_f(fx, f_ref) <= PARITY and rel_g(gx, g_ref) <= PARITY
        group.close()
    assert results["1"][0] == results["0"][0] and np.array_equal(results["1"][1], results["0"][1])
    assert results["1"][2] == results["0"][2]


def test_two_host_threads_two_clones_one_device(ctx, oracle):
    """ml::tune evaluates folds x trials concurrently on separate function objects (src/machine/tune.cpp:25-42);
    app/bench_solver.cpp:296 clones per thread. Distinct objectives on one dataset are concurrent-safe."""
    import libnano_b200 as nb

    X, Y, x0 = make_problem(21, 30000, 1024, 1, True)
    data = nb.Dataset.pin(ctx, X, Y)
    base = nb.LinearFunction(data, nb.Loss("s-logistic"), 0.0, 1e-2)
    clones = [base.clone(), base.clone()]
    points = [x0, -0.5 * x0]
    expected = [oracle.linear_vgrad(X, Y, "s-logistic", p, l2=1e-2) for p in points]
    failures = []

    def work(function, x, f_ref, g_ref):
        try:
            for _ in range(40):
                gx = np.empty_like(x)
                fx = function(x, gx)
                if rel_f(fx, f_ref) > PARITY or rel_g(gx, g_ref) > PARITY:
                    failures.append((fx, f_ref))
                    return
        except Exception as error:  # noqa: BLE001
            failures.append(error)

    threads = [threading.Thread(target=work, args=(clones[i], points[i], *expected[i])) for i in range(2)]
    for thread in threads:
        thread.start()
    for thread in threads:
        thread.join()
    assert not failures, failures


def test_two_host_threads_on_group_clones(oracle):
    """The same with group objectives: both threads fan out over all devices; launches are ordered by one lock."""
    import libnano_b200 as nb

    X, Y, x0 = make_problem(22, 20000, 512, 1, True)
    group = nb.Context(group_devices(2))
    base = nb.LinearFunction(nb.Dataset.pin(group, X, Y), nb.Loss("s-hinge"), 1e-2, 0.0)
    clones = [base.clone(), base.clone()]
    points = [x0, 0.25 * x0]
    expected = [oracle.linear_vgrad(X, Y, "s-hinge", p, l1=1e-2) for p in points]
    failures = []

    def work(function, x, f_ref, g_ref):
        try:
            for _ in range(25):
                gx = np.empty_like(x)
                fx = function(x, gx)
                if rel_f(fx, f_ref) > PARITY or rel_g(gx, g_ref) > PARITY:
                    failures.append((fx, f_ref))
                    return
        except Exception as error:  # noqa: BLE001
            failures.append(error)

    threads = [threading.Thread(target=work, args=(clones[i], points[i], *expected[i])) for i in range(2)]
    for thread in threads:
        thread.start()
    for thread in threads:
        thread.join()
    assert not failures, failures
    group.close()


@pytest.mark.skipif("device_count() < 2")
@pytest.mark.parametrize("mode", ["--fused", ""])
def test_sharded_processes_under_torchrun(mode):
    """One process per GPU (torchrun, NCCL rendezvous): sharded == whole dataset, identical bits on every rank, the
    sharded L-BFGS solve, and — fused — two connected objectives evaluated in opposite order on two ranks."""
    world = min(device_count(), 8)
    command = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_sharded.py")]
    if mode:
        command.append(mode)
    done = subprocess.run(command, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert done.returncode == 0, done.stdout[-4000:] + done.stderr[-4000:]
    assert '"ok": true' in done.stdout


def test_linear_model_fit_on_a_device_group():
    """linear_t::fit (src/linear.cpp:79-116) with every dataset of the fit split over the devices of one process."""
    import libnano_b200 as nb
    from test_linear_model import linear_datasource

    X, Y, W, b = linear_datasource(4, 3000, 64, noise=0.05)
    group = nb.Context(group_devices(2))
    model = nb.LinearModel("ridge")
    result = model.fit(group, X, Y, "mse", folds=2)
    assert result.trials() >= 5 and np.all(np.isfinite(result.valid))
    assert np.max(np.abs(model.weights() - W)) < 0.05 and np.max(np.abs(model.bias() - b)) < 0.05
    outputs = model.predict(group, X)
    assert np.max(np.abs(outputs - (X @ model.weights().T + model.bias()))) < 1e-10
    group.close()
