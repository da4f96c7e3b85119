"""N3: many (l1, l2) / parameter vectors in ONE pass over X — nb200_vgrad_batch against K separate
evaluations (GPU) and against the oracle."""
import numpy as np
import pytest

from conftest import PARITY, is_classification, make_problem, rel_f, rel_g

pytestmark = pytest.mark.gpu


# (one-pass tensor-pipe kernel: K <= 16 where W fits in shared memory and D >= 200; two-pass pair: K >= 6 elsewhere;
#  K single passes otherwise)
@pytest.mark.parametrize("N,D,K", [(3000, 128, 4), (5000, 256, 7), (2049, 1024, 16), (700, 384, 1), (1000, 100, 3),
                                   (4000, 512, 5), (3001, 1024, 8), (2500, 784, 2), (1500, 512, 16), (999, 300, 11),
                                   (6000, 640, 31), (1200, 512, 9), (1300, 256, 17), (800, 1024, 24), (900, 512, 25)])
@pytest.mark.parametrize("loss_id", ["s-logistic", "mse", "s-hinge", "mae"])
def test_batched_models_match_separate_evaluations(ctx, oracle, N, D, K, loss_id):
    import libnano_b200 as nb

    X, Y, _ = make_problem(N + D + K, N, D, 1, is_classification(loss_id))
    data = nb.Dataset.pin(ctx, X, Y)
    rng = np.random.default_rng(K)
    xs = rng.uniform(-1, 1, (K, D + 1))
    xs[:, :D] *= 2.0 / np.sqrt(D)
    l1s = rng.choice([0.0, 1e-2, 0.3], K)
    l2s = rng.choice([0.0, 1e-3, 0.5], K)
    function = nb.LinearFunction(data, nb.Loss(loss_id), 0.0, 0.0)
    fs, gs = function.vgrad_batch(xs, l1s, l2s)
    fv, none = function.vgrad_batch(xs, l1s, l2s, want_grad=False)
    assert none is None and function.fcalls() == 2 * K and function.gcalls() == K
    for k in range(K):
        f_ref, g_ref = oracle.linear_vgrad(X, Y, loss_id, xs[k], l1=l1s[k], l2=l2s[k])
        assert rel_f(fs[k], f_ref) <= PARITY and rel_g(gs[k], g_ref) <= PARITY, (k, fs[k], f_ref)
        assert rel_f(fv[k], f_ref) <= PARITY
        single = nb.LinearFunction(data, nb.Loss(loss_id), l1s[k], l2s[k])
        g1 = np.empty(D + 1)
        f1 = single(xs[k], g1)
        assert rel_f(fs[k], f1) <= 1e-13 and rel_g(gs[k], g1) <= 1e-13
    # same input => same bits
    f2, g2 = function.vgrad_batch(xs, l1s, l2s)
    assert np.array_equal(f2, fs) and np.array_equal(g2, gs)


def test_batched_regularisation_grid_like_the_tuner(ctx, oracle):
    # the tuner's grid (src/linear/util.cpp:75-86: 31 values of 1e-8 .. 1e+7) at one x: only the reg terms differ
    import libnano_b200 as nb

    N, D = 4000, 256
    X, Y, x = make_problem(3, N, D, 1, True)
    grid = np.array([1e-8, 3e-8, 1e-7, 3e-7, 1e-6, 3e-6, 1e-5, 3e-5, 1e-4, 3e-4, 1e-3, 3e-3, 1e-2, 3e-2, 1e-1, 3e-1, 1e0,
                     3e0, 1e1, 3e1, 1e2, 3e2, 1e3, 3e3, 1e4, 3e4, 1e5, 3e5, 1e6, 3e6, 1e7])
    function = nb.LinearFunction(nb.Dataset.pin(ctx, X, Y), nb.Loss("s-logistic"), 0.0, 0.0)
    fs, gs = function.vgrad_batch(np.tile(x, (grid.size, 1)), 0.0, grid)
    f0, g0 = oracle.linear_vgrad(X, Y, "s-logistic", x)
    W = x[:D]
    for k, l2 in enumerate(grid):
        assert abs(fs[k] - (f0 + 0.5 * l2 * np.mean(W * W))) <= 1e-12 * (1 + abs(fs[k]))
        assert np.max(np.abs(gs[k][:D] - (g0[:D] + l2 * W / D))) <= 1e-12 * (1 + np.max(np.abs(gs[k])))
        assert abs(gs[k][D] - g0[D]) <= 1e-13
    with pytest.raises(RuntimeError):
        function.vgrad_batch(np.zeros((2, D)), 0.0, 0.0)  # wrong n
