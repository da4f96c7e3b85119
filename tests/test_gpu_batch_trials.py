"""GPU tests of the callers of nb200_vgrad_batch: solver.lbfgs_batch (K L-BFGS solves sharing one pass over X per round) and
LinearModel.fit with batched trials (the default for smooth objectives). Their host logic is pinned on the CPU over an
oracle double (tests/test_host_solvers.py, tests/test_linear_model.py); here the combination runs on the device. (They
were non-strict xfail until their first B200 run: 4 / 4 XPASS in GPUTEST_r01.json and in round 2's own runs.)"""
import numpy as np
import pytest

from conftest import make_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,D,K", [(3000, 256, 8), (1000, 100, 3), (2048, 512, 31)])
def test_batched_lbfgs_matches_the_device_resident_solves(ctx, N, D, K):
    import libnano_b200 as nb
    from libnano_b200.solver import lbfgs, lbfgs_batch

    X, Y, _ = make_problem(N + K, N, D, 1, True)
    data = nb.Dataset.pin(ctx, X, Y)
    l2s = np.logspace(-4, 2, K)
    function = nb.LinearFunction(data, nb.Loss("s-logistic"), 0.0, 0.0)
    results = lbfgs_batch(function, np.zeros((K, D + 1)), 0.0, l2s, epsilon=1e-8)
    assert results[0]["rounds"] < sum(r["fcalls"] for r in results)
    for l2, result in zip(l2s, results):
        assert result["status"] == "gradient_test", result
        single = lbfgs(nb.LinearFunction(data, nb.Loss("s-logistic"), 0.0, l2), np.zeros(D + 1), epsilon=1e-9)
        assert abs(result["fx"] - single["fx"]) <= 1e-8 * (1 + abs(single["fx"]))


def test_ridge_fit_with_batched_trials_on_the_gpu(ctx):
    from libnano_b200 import LinearModel
    from libnano_b200.solver import lbfgs
    from test_linear_model import linear_datasource

    X, Y, W, b = linear_datasource(2, 4000, 256, noise=0.05)
    sequential, batched = LinearModel("ridge"), LinearModel("ridge")
    solver = lambda f, x0: lbfgs(f, x0, epsilon=1e-9, max_evals=5000)  # noqa: E731
    r_seq = sequential.fit(ctx, X, Y, "mse", folds=2, solver=solver, batch_trials=False)
    r_bat = batched.fit(ctx, X, Y, "mse", folds=2, solver=solver, batch_trials=True)
    assert abs(r_bat.value(r_bat.optimum_trial()) - r_seq.value(r_seq.optimum_trial())) < 1e-6
    assert np.max(np.abs(batched.weights() - sequential.weights())) < 1e-4
    assert np.max(np.abs(batched.weights() - W)) < 0.05


def test_batched_trials_are_the_default_for_smooth_objectives(ctx):
    from libnano_b200 import LinearModel
    from test_linear_model import linear_datasource

    X, Y, W, b = linear_datasource(3, 3000, 128, noise=0.05)
    model = LinearModel("ridge")
    result = model.fit(ctx, X, Y, "mse", folds=2)          # no solver given: the trials of a fold advance together
    assert result.trials() >= 5 and np.max(np.abs(model.weights() - W)) < 0.05
    plain = LinearModel("ordinary")
    result = plain.fit(ctx, X, Y, "mse", folds=3)          # no hyper-parameters: still one cross-validated trial
    assert result.trials() == 1 and result.params.shape == (1, 0) and result.optimum_trial() == 0
    assert np.all(np.isfinite(result.valid)) and np.all(np.isfinite(result.train)) and result.value(0) < 0.1
    assert np.max(np.abs(plain.weights() - W)) < 0.05
