"""N1: what flatten_iterator_t does before the objective sees the data — column statistics, scaling,
NaN -> 0 — done once in HBM (nb200_dataset_scale), and nano::upscale of the fitted parameters.
Mirrors test/test_linear_function.cpp:152-188 (minimize_noreg recovers the generating weights) for every scaling."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SCALINGS = ("none", "mean", "minmax", "standard")


@pytest.mark.parametrize("N,D,T", [(1000, 13, 2), (20000, 257, 1), (5000, 1024, 3)])
def test_device_statistics_and_scaling_match_the_oracle(ctx, oracle, N, D, T):
    import libnano_b200 as nb

    rng = np.random.default_rng(N + D)
    X = rng.normal(1.5, 3.0, (N, D)) * rng.uniform(0.1, 10.0, D)
    X[rng.uniform(size=(N, D)) < 0.02] = np.nan
    X[:, 0] = -2.0  # constant column
    Y = rng.normal(0.0, 2.0, (N, T))
    enable_inputs = np.ones(D, dtype=np.uint8)
    enable_inputs[1] = 0
    for iscaling, tscaling in (("standard", "none"), ("minmax", "mean"), ("mean", "standard"), ("none", "minmax")):
        data = nb.Dataset.pin(ctx, X, Y).scale(iscaling, tscaling, enable_inputs=enable_inputs)
        istats, tstats = oracle.scalar_stats(X, enable_inputs), oracle.scalar_stats(Y)
        for which, ref in (("inputs", istats), ("targets", tstats)):
            got = data.stats(which)
            for k, name in enumerate(oracle.STAT_NAMES):
                assert np.allclose(got[name], ref[k], rtol=1e-12, atol=1e-12), (which, name)
        Xd, Yd = data.read()
        Xr, Yr = oracle.scale(X, istats, iscaling), oracle.scale(Y, tstats, tscaling)
        assert np.all(np.isfinite(Xd)) and np.all(Xd[~np.isfinite(X)] == 0.0)
        assert np.max(np.abs(Xd - Xr) / (1 + np.abs(Xr))) <= 1e-12 and np.max(np.abs(Yd - Yr) / (1 + np.abs(Yr))) <= 1e-12
    with pytest.raises(RuntimeError):
        data.scale("standard")  # once per dataset


@pytest.mark.parametrize("scaling", SCALINGS)
def test_fit_on_scaled_data_then_upscale_recovers_the_ground_truth(ctx, oracle, scaling):
    import libnano_b200 as nb
    from libnano_b200.solver import lbfgs

    rng = np.random.default_rng(1)
    N, D, T = 500, 6, 2
    X = rng.normal(2.0, 3.0, (N, D)) * np.array([1.0, 10.0, 0.1, 5.0, 1.0, 2.0])
    W, b = rng.uniform(-1, 1, (T, D)), rng.uniform(-1, 1, T)
    Y = X @ W.T + b
    data = nb.Dataset.pin(ctx, X, Y).scale(scaling, scaling)
    function = nb.LinearFunction(data, nb.Loss("mse"), 0.0, 0.0)
    result = lbfgs(function, np.zeros(function.size()), epsilon=1e-10, max_evals=10000)
    assert result["status"] == "gradient_test" and abs(result["fx"]) < 1e-7
    Wu, bu = data.upscale(function.weights(result["x"]), function.bias(result["x"]))
    assert np.max(np.abs(Wu - W)) < 1e-6 and np.max(np.abs(bu - b)) < 1e-5
    # the oracle's upscale gives the same numbers from the same statistics
    istats = np.stack([data.stats("inputs")[n] for n in oracle.STAT_NAMES])
    tstats = np.stack([data.stats("targets")[n] for n in oracle.STAT_NAMES])
    Wo, bo = oracle.upscale(istats, scaling, tstats, scaling, function.weights(result["x"]), function.bias(result["x"]))
    assert np.max(np.abs(Wo - Wu)) <= 1e-13 * (1 + np.max(np.abs(Wu))) and np.max(np.abs(bo - bu)) <= 1e-12
