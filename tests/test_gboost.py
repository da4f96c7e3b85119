"""CPU tests of the oracle's restatement of the gradient-boosting cousins (gboost::bias_function_t, scale_function_t,
grads_function_t, src/gboost/function.cpp), pinned on the reference's own known answers for them
(test/test_gboost_function.cpp:123-218): the fixture's data (targets = outputs + scale[group] * woutputs, only odd
samples assigned to a cluster), function sizes, f(0) against the closed form for the mse loss, the known optima (column
means / the generating scales / the targets) and the finite-difference gradient check of test/fixture/function.h:39-57."""
import numpy as np
import pytest

from conftest import EPSILON2


def fixture(samples=100, tsize=2, groups=3, seed=0):
    """fixture_datasource_t (test/test_gboost_function.cpp:12-101)."""
    rng = np.random.default_rng(seed)
    scale = rng.uniform(1.1, 2.2, groups)
    outputs = rng.uniform(-1, 1, (samples, tsize))
    woutputs = rng.uniform(-1, 1, (samples, tsize))
    group = np.arange(samples) % groups
    targets = outputs + scale[group][:, None] * woutputs
    cluster = np.where(np.arange(samples) % 2 > 0, group, -1).astype(np.int32)  # :41-52
    return scale, outputs, woutputs, targets, cluster


def fd_gradient(function, x, delta=1e-6):
    g = np.empty_like(x)
    for i in range(x.size):
        xp, xn = x.copy(), x.copy()
        xp[i] += delta
        xn[i] -= delta
        g[i] = (function(xp) - function(xn)) / (2 * delta)
    return g


@pytest.mark.parametrize("rows", [(0, 100), (10, 60), (0, 50), (10, 100)])
def test_bias_function(oracle, rows):
    _, _, _, targets, _ = fixture(100)
    T = targets[rows[0]:rows[1]]
    f0, g0 = oracle.gboost_vgrad("bias", "mse", T, np.zeros(2))
    assert g0.size == 2                                                     # function.size() == tsize (:142)
    assert abs(f0 - np.mean(0.5 * np.sum(T ** 2, axis=1))) <= 1e-12         # check_value (:112-118)
    assert np.allclose(g0, -T.mean(axis=0), atol=1e-14)
    solved = oracle.lbfgs(lambda x, gx: _eval(oracle, "bias", "mse", T, x, gx), np.zeros(2), epsilon=1e-8)
    assert np.allclose(solved["x"], T.mean(axis=0), atol=1e-6)              # check_optimum (:105-110)


def _eval(oracle, kind, loss, targets, x, gx, **extra):
    f, g = oracle.gboost_vgrad(kind, loss, targets, x, want_grad=gx is not None, **extra)
    if gx is not None:
        gx[:] = g
    return f


@pytest.mark.parametrize("rows", [(0, 50), (10, 40), (0, 40), (10, 50)])
def test_scale_function(oracle, rows):
    scale, outputs, woutputs, targets, cluster = fixture(50)
    lo, hi = rows
    extra = dict(cluster=cluster[lo:hi], groups=3, soutputs=outputs[lo:hi], woutputs=woutputs[lo:hi])
    f0, g0 = oracle.gboost_vgrad("scale", "mse", targets[lo:hi], np.zeros(3), **extra)
    assert g0.size == 3                                                     # function.size() == groups (:175)
    assert abs(f0 - np.mean(0.5 * np.sum((targets[lo:hi] - outputs[lo:hi]) ** 2, axis=1))) <= 1e-12
    solved = oracle.lbfgs(lambda x, gx: _eval(oracle, "scale", "mse", targets[lo:hi], x, gx, **extra), np.zeros(3),
                          epsilon=1e-8)
    assert np.allclose(solved["x"], scale, atol=1e-6)                       # check_optimum(function, scale) (:181)
    # samples outside every cluster keep scale 0 whatever x is
    x = np.array([5.0, -3.0, 2.0])
    f_all = oracle.gboost_vgrad("scale", "mse", targets[lo:hi], x, want_grad=False, **extra)[0]
    outs = outputs[lo:hi] + np.where(cluster[lo:hi] < 0, 0.0, x[np.maximum(cluster[lo:hi], 0)])[:, None] * woutputs[lo:hi]
    assert abs(f_all - np.mean(0.5 * np.sum((targets[lo:hi] - outs) ** 2, axis=1))) <= 1e-12


def test_grads_function(oracle):
    _, _, _, targets, _ = fixture(10)
    x0 = np.zeros(20)
    f0, g0 = oracle.gboost_vgrad("grads", "mse", targets, x0)
    assert g0.size == 20 and abs(f0 - np.mean(0.5 * np.sum(targets ** 2, axis=1))) <= 1e-12
    assert np.allclose(g0, (x0 - targets.reshape(-1)) / 10.0, atol=1e-15)   # per-sample gradients / N (:36)
    solved = oracle.lbfgs(lambda x, gx: _eval(oracle, "grads", "mse", targets, x, gx), x0, epsilon=1e-9)
    assert np.allclose(solved["x"], targets.reshape(-1), atol=1e-6)         # check_optimum(function, targets) (:216)


@pytest.mark.parametrize("loss_id", ["mse", "cauchy", "s-logistic", "s-classnll", "s-exponential", "s-squared-hinge"])
@pytest.mark.parametrize("kind", ["bias", "scale", "grads"])
def test_gradients_match_finite_differences(oracle, kind, loss_id):
    scale, outputs, woutputs, targets, cluster = fixture(40, tsize=3, groups=4, seed=3)
    if loss_id != "mse" and loss_id != "cauchy":
        targets = np.where(targets > 0, 1.0, -1.0)
    extra = dict(cluster=cluster, groups=4, soutputs=outputs, woutputs=woutputs) if kind == "scale" else {}
    n = {"bias": 3, "scale": 4, "grads": 120}[kind]
    rng = np.random.default_rng(1)
    for _ in range(3):
        x = rng.uniform(-1, 1, n)
        _, g = oracle.gboost_vgrad(kind, loss_id, targets, x, **extra)
        approx = fd_gradient(lambda v: oracle.gboost_vgrad(kind, loss_id, targets, v, want_grad=False, **extra)[0], x)
        assert np.max(np.abs(g - approx)) <= 10 * EPSILON2 * (1 + np.max(np.abs(g)))


def test_batch_size_does_not_change_the_sums_beyond_rounding(oracle):
    _, outputs, woutputs, targets, cluster = fixture(257, tsize=2, groups=3, seed=5)
    extra = dict(cluster=cluster, groups=3, soutputs=outputs, woutputs=woutputs)
    x = np.array([0.3, -0.7, 1.9])
    results = [oracle.gboost_vgrad("scale", "cauchy", targets, x, batch=b, **extra) for b in (1, 7, 100, 1000)]
    for f, g in results[1:]:
        assert abs(f - results[0][0]) <= 1e-14 and np.max(np.abs(g - results[0][1])) <= 1e-14
