"""End-to-end on the GPU: the restated libnano drivers (oracle/solvers.cpp: L-BFGS + CG-DESCENT, OSGA) minimise the
GPU objective and the CPU oracle objective from the same start; the optima agree within the solver's epsilon
(BASELINE.json north_star). Mirrors test/test_linear_function.cpp:51-65,152-188 (minimize_noreg: lbfgs + cgdescent,
epsilon 1e-10, recovers the generating weights) and the non-smooth cases driven by OSGA."""
import numpy as np
import pytest

from conftest import make_problem

pytestmark = pytest.mark.gpu


def oracle_objective(oracle, X, Y, loss_id, l1, l2):
    def function(x, gx=None):
        fx, g = oracle.linear_vgrad(X, Y, loss_id, x, l1, l2, want_grad=gx is not None)
        if gx is not None:
            gx[:] = g
        return fx

    return function


def test_minimize_noreg_recovers_ground_truth(ctx, oracle):
    # noise-free linear targets: the unregularised mse objective has its minimum (0) at the generating (W, b)
    import libnano_b200 as nb

    rng = np.random.default_rng(0)
    N, D, T = 50, 4, 1
    X = rng.uniform(-1, 1, (N, D))
    W, b = rng.uniform(-1, 1, (T, D)), rng.uniform(-1, 1, T)
    Y = X @ W.T + b
    function = nb.LinearFunction(nb.Dataset.pin(ctx, X, Y), nb.Loss("mse"), 0.0, 0.0)
    result = oracle.lbfgs(function, np.zeros(function.size()), epsilon=1e-10, max_evals=10000)
    assert result["status"] == "gradient_test"
    assert abs(result["fx"]) < 1e-7 and function.fcalls() > 5
    assert np.max(np.abs(function.weights(result["x"]) - W)) < 1e-7
    assert np.max(np.abs(function.bias(result["x"]) - b)) < 1e-7
    # call counters seen by the solver == the function's own (test/fixture/solver.h:178-294)
    assert (result["fcalls"], result["gcalls"]) == (function.fcalls(), function.gcalls())


@pytest.mark.parametrize("N,D,T,loss_id,l2", [(20000, 64, 1, "s-logistic", 1e-2), (5000, 32, 3, "s-classnll", 1e-2),
                                              (8000, 100, 1, "mse", 0.0)])
def test_lbfgs_reaches_the_same_optimum_on_gpu_and_cpu(ctx, oracle, N, D, T, loss_id, l2):
    import libnano_b200 as nb

    X, Y, _ = make_problem(1, N, D, T, loss_id != "mse")
    gpu = nb.LinearFunction(nb.Dataset.pin(ctx, X, Y), nb.Loss(loss_id), 0.0, l2)
    cpu = oracle_objective(oracle, X, Y, loss_id, 0.0, l2)
    x0 = np.zeros(gpu.size())
    epsilon = 1e-8
    r_gpu = oracle.lbfgs(gpu, x0, epsilon=epsilon, max_evals=1000)
    r_cpu = oracle.lbfgs(cpu, x0, epsilon=epsilon, max_evals=1000)
    assert r_gpu["status"] == r_cpu["status"] == "gradient_test"
    assert r_gpu["gradient_test"] < epsilon and r_cpu["gradient_test"] < epsilon
    assert abs(r_gpu["fx"] - r_cpu["fx"]) <= epsilon * (1 + abs(r_cpu["fx"]))
    # strongly convex (l2 > 0) => the minimiser is unique: the iterates end within the solver's epsilon of each other
    if l2 > 0:
        mu = l2 / (D * T)
        assert np.max(np.abs(r_gpu["x"] - r_cpu["x"])) <= 10 * epsilon / mu
    assert abs(r_gpu["fcalls"] - r_cpu["fcalls"]) <= max(3, r_cpu["fcalls"] // 4)
    # the GPU optimum is a stationary point of the CPU objective too
    g = np.empty_like(x0)
    f = cpu(r_gpu["x"], g)
    assert np.max(np.abs(g)) / max(1.0, abs(f)) < 10 * epsilon


@pytest.mark.parametrize("loss_id,l1,l2", [("s-hinge", 1e-2, 0.0), ("s-hinge", 1e-2, 1e-2), ("mae", 0.0, 1e-2)])
def test_osga_on_nonsmooth_objectives(ctx, oracle, loss_id, l1, l2):
    # BASELINE config 5 in small: hinge + lasso / elastic-net are non-smooth => OSGA through the GPU vgrad
    import libnano_b200 as nb

    N, D = 4000, 24
    X, Y, _ = make_problem(2, N, D, 1, loss_id != "mae")
    gpu = nb.LinearFunction(nb.Dataset.pin(ctx, X, Y), nb.Loss(loss_id), l1, l2)
    assert not gpu.smooth() and gpu.convex()
    cpu = oracle_objective(oracle, X, Y, loss_id, l1, l2)
    x0 = np.zeros(gpu.size())
    kwargs = dict(epsilon=1e-6, strong_convexity=gpu.strong_convexity(), max_evals=5000, patience=100)
    r_gpu = oracle.osga(gpu, x0, **kwargs)
    r_cpu = oracle.osga(cpu, x0, **kwargs)
    assert r_gpu["status"] in ("specific_test", "value_test", "max_iters")
    f0 = cpu(x0)
    assert r_gpu["fx"] < f0 and r_cpu["fx"] < f0  # progress from the start (test_linear_function.cpp l1/l2 variants)
    # same driver, same objective up to 1e-12 => optima within OSGA's (loose) precision of each other
    assert abs(r_gpu["fx"] - r_cpu["fx"]) <= 1e-4 * (1 + abs(r_cpu["fx"]))
    assert abs(cpu(r_gpu["x"]) - r_gpu["fx"]) <= 1e-12 * (1 + abs(r_gpu["fx"]))
    # one f(x, g) + one f(x') per iteration (src/solver/osga.cpp:101,112)
    assert r_gpu["fcalls"] == 2 * r_gpu["iterations"] + 1 and r_gpu["gcalls"] == r_gpu["iterations"] + 1


# ----------------------------------------------------------------------------------------------------------------
# device-resident L-BFGS (nb200_lbfgs_minimize): same algorithm, vectors never leave HBM
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,D,T,loss_id,l2", [(20000, 64, 1, "s-logistic", 1e-2), (5000, 128, 4, "s-classnll", 1e-2),
                                              (8000, 100, 1, "mse", 0.0), (3000, 33, 2, "mse", 1e-3)])
def test_device_resident_lbfgs_matches_the_host_driven_solve(ctx, oracle, N, D, T, loss_id, l2):
    import libnano_b200 as nb
    from libnano_b200.solver import lbfgs

    X, Y, _ = make_problem(1, N, D, T, loss_id != "mse")
    gpu = nb.LinearFunction(nb.Dataset.pin(ctx, X, Y), nb.Loss(loss_id), 0.0, l2)
    x0 = np.zeros(gpu.size())
    epsilon = 1e-8
    r_dev = lbfgs(gpu, x0, epsilon=epsilon, max_evals=1000)
    assert r_dev["status"] == "gradient_test" and r_dev["gradient_test"] < epsilon
    assert (gpu.fcalls(), gpu.gcalls()) == (r_dev["fcalls"], r_dev["gcalls"]) and r_dev["fcalls"] == r_dev["gcalls"]

    # (a) the restated CPU driver over the CPU oracle objective: same optimum within the solver's epsilon
    r_cpu = oracle.lbfgs(oracle_objective(oracle, X, Y, loss_id, 0.0, l2), x0, epsilon=epsilon, max_evals=1000)
    assert r_cpu["status"] == "gradient_test"
    assert abs(r_dev["fx"] - r_cpu["fx"]) <= epsilon * (1 + abs(r_cpu["fx"]))
    if l2 > 0:
        assert np.max(np.abs(r_dev["x"] - r_cpu["x"])) <= 10 * epsilon / (l2 / (D * T))
    assert abs(r_dev["fcalls"] - r_cpu["fcalls"]) <= max(3, r_cpu["fcalls"] // 4)

    # (b) the returned point is what it claims: f and the gradient test re-evaluated through the host-buffer call
    g = np.empty_like(x0)
    f = gpu(r_dev["x"], g)
    assert abs(f - r_dev["fx"]) <= 1e-12 * (1 + abs(f))
    assert abs(np.max(np.abs(g)) / max(1.0, abs(f)) - r_dev["gradient_test"]) <= 1e-12


def test_device_resident_lbfgs_contract(ctx):
    import libnano_b200 as nb
    from libnano_b200.solver import lbfgs

    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (50, 4))
    W, b = rng.uniform(-1, 1, (1, 4)), rng.uniform(-1, 1, 1)
    function = nb.LinearFunction(nb.Dataset.pin(ctx, X, X @ W.T + b), nb.Loss("mse"), 0.0, 0.0)
    result = lbfgs(function, np.zeros(5), epsilon=1e-10, max_evals=10000)
    assert result["status"] == "gradient_test" and abs(result["fx"]) < 1e-7
    assert np.max(np.abs(result["x"] - np.concatenate([W.ravel(), b]))) < 1e-7  # test_linear_function.cpp:152-188
    # restarting at the optimum converges at the first gradient test (the line search still probes a few steps:
    # |f - f0| < epsilon1 triggers the "step too small" growth of src/lsearchk.cpp:61-69)
    again = lbfgs(function, result["x"], epsilon=1e-6)
    assert again["status"] == "gradient_test" and again["iterations"] <= 1 and again["fcalls"] <= 40
    # exactly at a stationary point the early exit of lbfgs.cpp:34-38 costs one evaluation
    exact = nb.LinearFunction(nb.Dataset.pin(ctx, np.eye(4), np.zeros((4, 1))), nb.Loss("mse"), 0.0, 0.0)
    at_zero = lbfgs(exact, np.zeros(5))
    assert at_zero["status"] == "gradient_test" and at_zero["fcalls"] == 1 and at_zero["fx"] == 0.0
    # max_evals is honoured (lbfgs.cpp:45)
    short = lbfgs(function, np.zeros(5), epsilon=1e-14, max_evals=6)
    assert short["status"] in ("max_iters", "gradient_test") and short["fcalls"] + short["gcalls"] <= 6 + 8
    with pytest.raises(RuntimeError, match="incompatible initial point"):
        lbfgs(function, np.zeros(4))


def test_device_resident_osga_matches_the_restated_driver(ctx, oracle):
    # nb200_osga_minimize (every vector in HBM), libnano_b200.solver.osga_host (numpy vectors) and oracle/solvers.cpp
    # restate the same algorithm independently: same GPU objective, same start => the same trajectory up to rounding
    import libnano_b200 as nb
    from libnano_b200.solver import osga, osga_host

    N, D = 3000, 16
    X, Y, _ = make_problem(5, N, D, 1, True)
    for l1, l2 in ((1e-2, 0.0), (1e-2, 1e-2)):
        gpu = nb.LinearFunction(nb.Dataset.pin(ctx, X, Y), nb.Loss("s-hinge"), l1, l2)
        x0 = np.zeros(gpu.size())
        mine = osga(gpu, x0, epsilon=1e-6, max_evals=2000)
        ref = oracle.osga(gpu.clone(), x0, epsilon=1e-6, strong_convexity=gpu.strong_convexity(), max_evals=2000)
        assert mine["status"] == ref["status"]
        assert abs(mine["iterations"] - ref["iterations"]) <= max(2, ref["iterations"] // 20)
        assert abs(mine["fx"] - ref["fx"]) <= 1e-6 * (1 + abs(ref["fx"]))
        assert mine["fcalls"] == 2 * mine["iterations"] + 1 and mine["gcalls"] == mine["iterations"] + 1
        assert gpu.fcalls() == mine["fcalls"] and gpu.gcalls() == mine["gcalls"]
        host = osga_host(gpu.clone(), x0, epsilon=1e-6, max_evals=2000)
        assert host["status"] == mine["status"]
        assert abs(host["iterations"] - mine["iterations"]) <= max(2, mine["iterations"] // 20)
        assert abs(host["fx"] - mine["fx"]) <= 1e-6 * (1 + abs(mine["fx"]))
        assert np.max(np.abs(host["x"] - mine["x"])) <= 1e-3


def test_device_resident_osga_on_a_device_group(oracle):
    # the same solve with the samples split over a device group (one device serving twice on a one-GPU box)
    import libnano_b200 as nb
    from libnano_b200.solver import osga
    from test_gpu_group import group_devices

    X, Y, _ = make_problem(6, 5000, 32, 1, True)
    whole_ctx, group_ctx = nb.Context(0), nb.Context(group_devices(2))
    whole = nb.LinearFunction(nb.Dataset.pin(whole_ctx, X, Y), nb.Loss("s-hinge"), 1e-2, 1e-2)
    split = nb.LinearFunction(nb.Dataset.pin(group_ctx, X, Y), nb.Loss("s-hinge"), 1e-2, 1e-2)
    a = osga(whole, np.zeros(33), epsilon=1e-6, max_evals=3000)
    b = osga(split, np.zeros(33), epsilon=1e-6, max_evals=3000)
    assert a["status"] == b["status"] and abs(a["fx"] - b["fx"]) <= 1e-6 * (1 + abs(a["fx"]))
    with pytest.raises(RuntimeError, match="incompatible initial point"):
        osga(split, np.zeros(5))
    group_ctx.close()
    whole_ctx.close()


@pytest.mark.parametrize("loss_id,l1,l2", [("s-hinge", 1e-2, 0.0), ("s-hinge", 1e-2, 1e-2), ("mae", 1e-2, 0.0)])
def test_host_driven_rqb_on_nonsmooth_objectives(ctx, oracle, loss_id, l1, l2):
    # BASELINE config 5 in small, second driver: the RQB proximal bundle solver through the GPU vgrad. Same driver
    # over the CPU oracle objective (and, for lasso, the linear program's exact optimum) is the check.
    import libnano_b200 as nb
    from libnano_b200.solver import rqb
    from test_host_solvers import OracleFunction, informative_problem, lasso_linear_program

    N, D = 600, 8
    X, Y = informative_problem(4, N, D, loss_id != "mae")
    gpu = nb.LinearFunction(nb.Dataset.pin(ctx, X, Y), nb.Loss(loss_id), l1, l2)
    cpu = OracleFunction(oracle, X, Y, loss_id, l1=l1, l2=l2)
    x0 = np.zeros(gpu.size())
    epsilon = 1e-8
    r_gpu = rqb(gpu, x0, epsilon=epsilon, max_evals=4000)
    r_cpu = rqb(cpu, x0, epsilon=epsilon, max_evals=4000)
    assert r_gpu["status"] == "specific_test" and r_cpu["status"] == "specific_test", (r_gpu, r_cpu)
    assert abs(r_gpu["fx"] - r_cpu["fx"]) <= 10 * epsilon * (1 + abs(r_cpu["fx"]))
    assert abs(cpu(r_gpu["x"]) - r_gpu["fx"]) <= 1e-12 * (1 + abs(r_gpu["fx"]))
    assert r_gpu["fcalls"] == r_gpu["gcalls"] == gpu.fcalls() and gpu.launches() > 0
    if l2 == 0.0:
        f_star = lasso_linear_program(X, Y, loss_id, l1)
        assert f_star - 1e-10 <= r_gpu["fx"] <= f_star + 10 * epsilon * (1 + abs(f_star))


# ----------------------------------------------------------------------------------------------------------------
# device-resident CGD (nb200_cgd_minimize): the ten beta formulas of src/solver/cgd.cpp on the same skeleton
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("beta", ["hs", "fr", "pr", "cd", "ls", "dy", "n", "dycd", "dyhs", "frpr"])
def test_device_resident_cgd_matches_the_restated_driver(ctx, oracle, beta):
    import libnano_b200 as nb
    from libnano_b200.solver import cgd

    N, D, l2 = 6000, 48, 1e-2
    X, Y, _ = make_problem(2, N, D, 1, True)
    gpu = nb.LinearFunction(nb.Dataset.pin(ctx, X, Y), nb.Loss("s-logistic"), 0.0, l2)
    x0 = np.zeros(gpu.size())
    epsilon = 1e-8
    r_dev = cgd(gpu, x0, beta=beta, epsilon=epsilon, max_evals=2000)
    assert r_dev["status"] == "gradient_test" and r_dev["gradient_test"] < epsilon, r_dev
    assert (gpu.fcalls(), gpu.gcalls()) == (r_dev["fcalls"], r_dev["gcalls"])
    # the restated CPU driver over the CPU oracle objective: same optimum within the solver's epsilon, similar effort
    r_cpu = oracle.cgd(oracle_objective(oracle, X, Y, "s-logistic", 0.0, l2), x0, beta=beta, epsilon=epsilon,
                       max_evals=2000)
    assert r_cpu["status"] == "gradient_test"
    assert abs(r_dev["fx"] - r_cpu["fx"]) <= epsilon * (1 + abs(r_cpu["fx"]))
    assert np.max(np.abs(r_dev["x"] - r_cpu["x"])) <= 10 * epsilon / (l2 / D)
    assert abs(r_dev["fcalls"] - r_cpu["fcalls"]) <= max(4, r_cpu["fcalls"] // 3)
    # the returned point is what it claims
    g = np.empty_like(x0)
    f = gpu(r_dev["x"], g)
    assert abs(f - r_dev["fx"]) <= 1e-12 * (1 + abs(f))


def test_device_resident_cgd_contract(ctx):
    import libnano_b200 as nb
    from libnano_b200.solver import cgd

    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (50, 4))
    W, b = rng.uniform(-1, 1, (1, 4)), rng.uniform(-1, 1, 1)
    function = nb.LinearFunction(nb.Dataset.pin(ctx, X, X @ W.T + b), nb.Loss("mse"), 0.0, 0.0)
    result = cgd(function, np.zeros(5), epsilon=1e-10, max_evals=10000)
    assert result["status"] == "gradient_test" and abs(result["fx"]) < 1e-7
    assert np.max(np.abs(result["x"] - np.concatenate([W.ravel(), b]))) < 1e-7  # test_linear_function.cpp:152-188
    at_optimum = cgd(function, np.concatenate([W.ravel(), b]))
    assert at_optimum["status"] == "gradient_test" and at_optimum["fcalls"] == 1  # cgd.cpp:95-99
    with pytest.raises(RuntimeError, match="unknown cgd variant"):
        cgd(function, np.zeros(5), beta="xyz")
    with pytest.raises(RuntimeError, match="incompatible initial point"):
        cgd(function, np.zeros(4))
    # many targets go through the tensor-pipe / generic evaluation paths under the same driver
    Xm, Ym, _ = make_problem(3, 4000, 128, 4, True)
    multi = nb.LinearFunction(nb.Dataset.pin(ctx, Xm, Ym), nb.Loss("s-classnll"), 0.0, 1e-2)
    solved = cgd(multi, np.zeros(multi.size()), epsilon=1e-8)
    assert solved["status"] == "gradient_test"
