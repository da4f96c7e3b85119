"""CPU tests (no GPU) of the restated drivers in oracle/solvers.cpp: L-BFGS, the ten CGD variants and OSGA minimise
the oracle's own linear objective. Mirrors test/test_linear_function.cpp:51-65,152-188 (minimize_noreg: the solver
recovers the generating weights to 1e-7 with epsilon 1e-10) and test/test_solver_*: every smooth solver reaches the
same optimum of a strongly convex problem."""
import numpy as np
import pytest

from conftest import make_problem


def objective(oracle, X, Y, loss_id, l1=0.0, l2=0.0):
    def function(x, gx=None):
        fx, g = oracle.linear_vgrad(X, Y, loss_id, x, l1, l2, want_grad=gx is not None)
        if gx is not None:
            gx[:] = g
        return fx

    return function


def noise_free_problem(seed=0, N=50, D=4, T=1):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (N, D))
    W, b = rng.uniform(-1, 1, (T, D)), rng.uniform(-1, 1, T)
    return X, X @ W.T + b, W, b


CGD_BETAS = ("hs", "fr", "pr", "cd", "ls", "dy", "n", "dycd", "dyhs", "frpr")


def test_lbfgs_minimize_noreg_recovers_ground_truth(oracle):
    X, Y, W, b = noise_free_problem()
    result = oracle.lbfgs(objective(oracle, X, Y, "mse"), np.zeros(5), epsilon=1e-10, max_evals=10000)
    assert result["status"] == "gradient_test" and result["gradient_test"] < 1e-10
    assert np.max(np.abs(result["x"] - np.concatenate([W.ravel(), b]))) < 1e-7


@pytest.mark.parametrize("beta", CGD_BETAS)
def test_cgd_minimize_noreg_recovers_ground_truth(oracle, beta):
    # every beta formula of src/solver/cgd.cpp:7-71 (cgd-hs ... cgd-frpr) with the restarts of cgd.cpp:121-126
    X, Y, W, b = noise_free_problem()
    result = oracle.cgd(objective(oracle, X, Y, "mse"), np.zeros(5), beta=beta, epsilon=1e-10, max_evals=10000)
    assert result["status"] == "gradient_test" and result["gradient_test"] < 1e-10, result
    assert np.max(np.abs(result["x"] - np.concatenate([W.ravel(), b]))) < 1e-7
    assert result["fcalls"] == result["gcalls"] > result["iterations"] > 1


@pytest.mark.parametrize("loss_id,T", [("s-logistic", 1), ("s-classnll", 3)])
def test_smooth_solvers_agree_on_a_strongly_convex_problem(oracle, loss_id, T):
    X, Y, _ = make_problem(3, 400, 12, T, True)
    function = objective(oracle, X, Y, loss_id, 0.0, 1e-1)
    x0 = np.zeros((12 + 1) * T)
    reference = oracle.lbfgs(function, x0, epsilon=1e-9)
    assert reference["status"] == "gradient_test"
    for beta in CGD_BETAS:
        result = oracle.cgd(function, x0, beta=beta, epsilon=1e-9, max_evals=5000)
        assert result["status"] == "gradient_test", (beta, result)
        assert abs(result["fx"] - reference["fx"]) <= 1e-9 * (1 + abs(reference["fx"]))
        assert np.max(np.abs(result["x"] - reference["x"])) < 1e-5, beta


def test_cgd_early_exit_and_budget(oracle):
    X, Y, W, b = noise_free_problem()
    function = objective(oracle, X, Y, "mse")
    at_optimum = oracle.cgd(function, np.concatenate([W.ravel(), b]))
    assert at_optimum["status"] == "gradient_test" and at_optimum["fcalls"] == 1  # cgd.cpp:95-99
    short = oracle.cgd(function, np.zeros(5), epsilon=1e-14, max_evals=6)
    assert short["status"] == "max_iters" and short["fcalls"] + short["gcalls"] <= 6 + 2 * 10


def test_osga_decreases_a_nonsmooth_objective(oracle):
    X, Y, _ = make_problem(5, 300, 8, 1, True)
    function = objective(oracle, X, Y, "s-hinge", 1e-2, 1e-2)
    x0 = np.zeros(9)
    f0 = function(x0)
    result = oracle.osga(function, x0, epsilon=1e-6, strong_convexity=1e-2 / 8, max_evals=2000)
    assert result["fx"] < f0 and result["status"] in ("value_test", "specific_test", "max_iters")
