"""linear_t::fit mirrored over the GPU objective (libnano_b200/linear.py). Mirrors the reference's model
tests (test/test_linear_ordinary.cpp, test_linear_ridge.cpp, test_linear_lasso.cpp: fit on a synthetic linear
datasource with some irrelevant features, then the weights and bias are recovered and the predictions match the
targets). CPU tests run the same orchestration over an oracle-backed backend (test infrastructure) and the host-only
pieces (splitter, tuner); GPU tests run the product path."""
import numpy as np
import pytest

from libnano_b200 import linear as nb_linear
from libnano_b200 import solver as nb_solver
from test_host_solvers import OracleFunction


def linear_datasource(seed, N, D, T=1, relevant_modulo=2, noise=0.0, classification=False):
    """test/fixture/datasource/linear.h in small: targets = W x + b with only every `relevant_modulo`-th feature used"""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (N, D)) * rng.uniform(0.5, 20.0, D) + rng.uniform(-3, 3, D)    # badly scaled on purpose
    W = rng.uniform(-1, 1, (T, D))
    W[:, np.arange(D) % relevant_modulo != 0] = 0.0
    b = rng.uniform(-1, 1, T)
    Y = X @ W.T + b + noise * rng.normal(size=(N, T))
    if classification:
        Y = np.where(Y - np.median(Y, axis=0) < 0, -1.0, 1.0)
    return X, Y, W, b


# ----------------------------------------------------------------------------------------------------------------
# an oracle-backed backend with the interface of linear._DeviceBackend
# ----------------------------------------------------------------------------------------------------------------
class OracleDataset:
    def __init__(self, oracle, X, Y):
        self.oracle, self.X, self.Y = oracle, np.array(X, dtype=np.float64), np.array(Y, dtype=np.float64)
        self.istats = self.tstats = None
        self.iscaling = self.tscaling = "none"

    @property
    def inputs(self):
        return self.X.shape[1]

    @property
    def targets(self):
        return self.Y.shape[1]

    def scale(self, inputs="standard", targets="none"):
        self.istats, self.tstats = self.oracle.scalar_stats(self.X), self.oracle.scalar_stats(self.Y)
        self.iscaling, self.tscaling = inputs, targets
        self.X = self.oracle.scale(self.X, self.istats, inputs)
        self.Y = self.oracle.scale(self.Y, self.tstats, targets)
        return self

    def upscale(self, W, b):
        if self.istats is None:
            return np.array(W), np.array(b)
        return self.oracle.upscale(self.istats, self.iscaling, self.tstats, self.tscaling, W, b)

    def evaluate(self, loss, W, b):
        return self.oracle.evaluate(self.X, self.Y, loss.type_id(), W, b)

    def predict(self, W, b):
        return self.oracle.predict(self.X, W, b)


class OracleLinearFunction(OracleFunction):
    def weights(self, x):
        D, T = self._X.shape[1], self._Y.shape[1]
        return np.asarray(x)[: D * T].reshape(T, D)

    def bias(self, x):
        D, T = self._X.shape[1], self._Y.shape[1]
        return np.asarray(x)[D * T:]


    def vgrad_batch(self, xs, l1s, l2s, want_grad=True):
        # the product evaluates the K models in one pass over X (nb200_vgrad_batch); the double one by one
        type(self).batch_calls = getattr(type(self), "batch_calls", 0) + 1
        fs, gs = np.empty(len(xs)), np.empty_like(xs)
        for k, x in enumerate(xs):
            fs[k], gs[k] = self._oracle.linear_vgrad(self._X, self._Y, self._loss_id, x, float(l1s[k]), float(l2s[k]))
        return fs, (gs if want_grad else None)


class OracleBackend:
    def __init__(self, oracle):
        self.oracle = oracle

    def pin(self, X, Y):
        return OracleDataset(self.oracle, X, Y)

    def function(self, data, loss, l1reg, l2reg):
        return OracleLinearFunction(self.oracle, data.X, data.Y, loss.type_id(), l1=l1reg, l2=l2reg,
                                    smooth=loss.smooth() and l1reg <= 0.0)


def cpu_solver(oracle):
    def solve(function, x0):
        if function.smooth():
            result = oracle.lbfgs(lambda x, gx=None: function.do_eval(np.asarray(x), gx), x0, epsilon=1e-10,
                                  max_evals=5000)
            assert result["status"] == "gradient_test"
            return result
        return nb_solver.rqb(function, x0, epsilon=1e-9, max_evals=4000)

    return solve


# ----------------------------------------------------------------------------------------------------------------
# host-only pieces
# ----------------------------------------------------------------------------------------------------------------
def test_param_space_is_the_reference_grid():
    grid = nb_linear.PARAM_GRID       # src/linear/util.cpp:75-86
    assert len(grid) == 31 and grid[0] == 1e-8 and grid[1] == 3e-8 and grid[15] == 3e-1 and grid[-1] == 1e7
    assert all(a < b for a, b in zip(grid, grid[1:]))


@pytest.mark.parametrize("count,folds", [(10, 2), (11, 3), (100, 5), (5, 5)])
def test_kfold_split_partitions_the_samples(count, folds):
    samples = np.arange(100, 100 + count)
    splits = nb_linear.kfold_split(samples, folds, seed=7)
    assert len(splits) == folds
    seen = []
    for fold, (train, valid) in enumerate(splits):
        assert np.array_equal(np.sort(np.concatenate([train, valid])), samples)      # a partition, both parts sorted
        assert np.all(np.diff(train) > 0) and np.all(np.diff(valid) > 0)
        assert valid.size == (count // folds if fold + 1 < folds else count - (folds - 1) * (count // folds))
        seen.append(valid)
    assert np.array_equal(np.sort(np.concatenate(seen)), samples)                    # every sample validates once
    again = nb_linear.kfold_split(samples, folds, seed=7)
    assert all(np.array_equal(a[1], b[1]) for a, b in zip(splits, again))
    with pytest.raises(RuntimeError):
        nb_linear.kfold_split(samples, count + 1)


def test_local_search_neighbourhood():
    # src/tuner/util.cpp:60-83
    assert nb_linear.local_search((0,), (30,), (15,), 2) == [(13,), (15,), (17,)]
    assert nb_linear.local_search((0,), (30,), (0,), 1) == [(0,), (1,)]
    assert len(nb_linear.local_search((0, 0), (30, 30), (15, 15), 4)) == 9
    assert len(nb_linear.local_search((0, 0), (30, 30), (30, 0), 1)) == 4


@pytest.mark.parametrize("optimum", [(0,), (30,), (11,), (22,)])
def test_tuner_finds_the_grid_optimum_of_a_convex_response(optimum):
    grid = nb_linear.PARAM_GRID
    calls = []

    def callback(params):
        calls.extend(tuple(p) for p in params)
        return [abs(np.log10(p[0]) - np.log10(grid[optimum[0]])) for p in params]

    steps = nb_linear.tune([grid], callback)
    assert steps[0][0] == optimum and steps[0][2] == 0.0
    assert len(calls) == len(set(calls)) == len(steps) <= 31                 # no grid point is evaluated twice
    assert calls[0] == (grid[15],)                                           # starts from the middle of the grid
    assert [s[2] for s in steps] == sorted(s[2] for s in steps)


def test_tuner_two_parameters_and_budget():
    grid = nb_linear.PARAM_GRID
    target = (grid[9], grid[20])
    steps = nb_linear.tune([grid, grid], lambda params: [abs(np.log10(p[0] / target[0])) + abs(np.log10(p[1] / target[1]))
                                                         for p in params])
    assert steps[0][0] == (9, 20)
    capped = nb_linear.tune([grid, grid], lambda params: [float(p[0] + p[1]) for p in params], max_evals=12)
    assert len(capped) <= 12 + 9
    with pytest.raises(RuntimeError, match="invalid value"):
        nb_linear.tune([grid], lambda params: [np.nan for _ in params])
    with pytest.raises(RuntimeError, match="at least one parameter space"):
        nb_linear.tune([], lambda params: [])


def test_feature_importance_and_sparsity_ratio():
    # src/linear/util.cpp:51-73 (test/test_linear_util.cpp)
    W = np.array([[1.0, -2.0, 0.0, 1e-9], [0.5, 0.0, 0.0, -1e-9]])
    importance = nb_linear.feature_importance(W)
    assert np.allclose(importance, [1.5, 2.0, 0.0, 2e-9], rtol=0, atol=1e-18)
    assert nb_linear.sparsity_ratio(importance) == 0.5 and nb_linear.sparsity_ratio(importance, 1e-12) == 0.25
    with pytest.raises(RuntimeError):
        nb_linear.sparsity_ratio(np.array([-1.0]))


def test_reference_known_answers_of_the_host_pieces():
    # test/test_linear_util.cpp:82-110 (13 flatten columns of 4 features) and test/test_tuner.cpp:224-272
    columns = [0, 1, 1, 1, 1, 1, 2, 3, 3, 3, 3, 3, 3]
    assert np.array_equal(nb_linear.feature_importance([[1, 0, -1, 0, 0, 1, -2, -3, 0, 0, 1, 1, 1]], columns),
                          [1, 2, 2, 6])
    assert np.array_equal(nb_linear.feature_importance([[0, 0, 0, 0, 0, 1, -2, -3, 0, 0, 0, 0, 0]], columns),
                          [0, 1, 2, 3])
    importance = np.array([0.0, 1e-9, 1e-3, 1e-1, 1e-7, 1e-7, 1e-1, 1e+1])
    for threshold, expected in ((1e-10, 1 / 8), (1e-8, 2 / 8), (1e-6, 4 / 8), (1e-2, 5 / 8)):
        assert nb_linear.sparsity_ratio(importance, threshold) == expected
    low, high, middle = (0, 0), (9, 6), (5, 3)
    assert nb_linear.local_search(low, high, low, 1) == [(0, 0), (0, 1), (1, 0), (1, 1)]
    assert nb_linear.local_search(low, high, high, 2) == [(7, 4), (7, 6), (9, 4), (9, 6)]
    assert nb_linear.local_search(low, high, middle, 1) == [(4, 2), (4, 3), (4, 4), (5, 2), (5, 3), (5, 4), (6, 2),
                                                            (6, 3), (6, 4)]


@pytest.mark.parametrize("seed", [42, 11, 122])
def test_kfold_reference_case(seed):
    # test/test_splitter.cpp:60-88: 25 samples, 5 folds -> 20 / 5, disjoint validation parts that cover the samples
    splits = nb_linear.kfold_split(np.arange(25), 5, seed)
    assert [(train.size, valid.size) for train, valid in splits] == [(20, 5)] * 5
    assert np.array_equal(np.sort(np.concatenate([valid for _, valid in splits])), np.arange(25))
    for train, valid in splits:
        assert not set(train) & set(valid)


@pytest.mark.parametrize("seed", [42, 11, 122])
def test_random_split_reference_case(seed):
    # test/test_splitter.cpp:90-108: 30 samples, 5 folds, 80 % -> 24 / 6, each fold a partition
    splits = nb_linear.random_split(np.arange(30), 5, seed)
    assert [(train.size, valid.size) for train, valid in splits] == [(24, 6)] * 5
    for train, valid in splits:
        assert np.array_equal(np.sort(np.concatenate([train, valid])), np.arange(30))
        assert np.all(np.diff(train) > 0) and np.all(np.diff(valid) > 0)


@pytest.mark.parametrize("splitter", ["k-fold", "random"])
def test_splitters_are_consistent(splitter):
    # test/test_splitter.cpp:110-140: same seed, same splits; another seed, other splits
    split = nb_linear.SPLITTERS[splitter]
    a, b, c = split(np.arange(21), 5, 42), split(np.arange(21), 5, 42), split(np.arange(21), 5, 11)
    assert all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(a, b))
    assert any(not np.array_equal(x[1], y[1]) for x, y in zip(a, c))


def test_evaluate_is_zero_for_the_generating_model(oracle):
    # test/test_linear_util.cpp:63-80: mse errors and values vanish (1e-12) at the datasource's own weights and bias
    rng = np.random.default_rng(9)
    X, W, b = rng.uniform(-1, 1, (20, 13)), rng.uniform(-1, 1, (3, 13)), rng.uniform(-1, 1, 3)
    values = oracle.evaluate(X, X @ W.T + b, "mse", W, b)
    assert values.shape == (2, 20) and np.max(np.abs(values)) < 1e-12


# test/test_tuner.cpp:13-26: param1 on a linear scale, param2 on a log10 scale
TUNER_SPACES = (nb_linear.ParamSpace((0.0, 0.1, 0.2, 0.3, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0), "linear"),
                nb_linear.ParamSpace((1e-3, 1e-2, 1e-1, 1e+0, 1e+1, 1e+2, 1e+3), "log10"))


def evaluate_ll(x, y, x0, y0):      # test/test_tuner.cpp:28-31
    return (x - x0) ** 2 + (y - y0) ** 2 + 0.5


def evaluate_10(x, y, x0, y0):      # test/test_tuner.cpp:33-36
    return (x - x0) ** 2 + (np.log10(y) - np.log10(y0) - (1.0 + x) / (1.0 + x0) + 1.0) ** 2 + 0.5


@pytest.mark.parametrize("tuner,evaluator", [("local-search", evaluate_ll), ("surrogate", evaluate_10)])
def test_tuners_find_every_grid_optimum(tuner, evaluator):
    # check_optimize (test/test_tuner.cpp:38-80): wherever the optimum sits on the 10 x 7 grid, the tuner ends on it,
    # its steps are sorted and no grid point is evaluated twice
    for ix0, x0 in enumerate(TUNER_SPACES[0]):
        for iy0, y0 in enumerate(TUNER_SPACES[1]):
            steps = nb_linear.tune(TUNER_SPACES, lambda params: [evaluator(p[0], p[1], x0, y0) for p in params],
                                   tuner=tuner)
            assert [s[2] for s in steps] == sorted(s[2] for s in steps)
            assert len({s[0] for s in steps}) == len(steps)
            assert steps[0][0] == (ix0, iy0) and tuple(steps[0][1]) == (x0, y0) and steps[0][2] == 0.5


@pytest.mark.parametrize("tuner", ["local-search", "surrogate"])
def test_tuners_fail_like_the_reference(tuner):
    # test/test_tuner.cpp:384-410
    with pytest.raises(RuntimeError, match="at least one parameter space"):
        nb_linear.tune([], lambda params: [], tuner=tuner)
    with pytest.raises(RuntimeError, match="invalid value"):
        nb_linear.tune(TUNER_SPACES, lambda params: [np.nan] * len(params), tuner=tuner)
    with pytest.raises(RuntimeError, match="unknown id"):
        nb_linear.tune(TUNER_SPACES, lambda params: [0.0] * len(params), tuner="bayesian")


def test_param_space_contract():
    # test/test_tuner.cpp:125-214
    for values, scale in (((), "log10"), ((1.0,), "log10"), ((1.0, 1.0), "linear"), ((2.0, 1.0), "linear"),
                          ((0.0, 1.0), "log10"), ((-1.0, 1.0), "log10")):
        with pytest.raises(RuntimeError):
            nb_linear.ParamSpace(values, scale)
    log_space, linear_space = TUNER_SPACES[1], TUNER_SPACES[0]
    assert log_space.to_surrogate(1e-3) == -3.0 and log_space.to_surrogate(1e+2) == 2.0
    assert linear_space.to_surrogate(0.5) == 0.5 and linear_space.to_surrogate(1.0) == 1.0
    with pytest.raises(RuntimeError, match="outside the parameter grid range"):
        log_space.to_surrogate(1e+4)
    assert [log_space.closest_grid_point_from_surrogate(v) for v in (-9.0, -2.6, -2.4, 0.1, 2.9, 7.0)] == [0, 0, 1, 3, 6, 6]
    assert [linear_space.closest_grid_point_from_surrogate(v) for v in (-1.0, 0.04, 0.41, 0.95, 3.0)] == [0, 0, 4, 8, 9]


@pytest.mark.parametrize("p,q", [((1.0,), (1.0, -2.0, 1.0)), ((1.0, -2.0), (5.0, -2.0, 4.0, 1.0, 0.0, 1.0)),
                                 ((0.1, 1.0), (1.0, 0.0, -2.0, 1.0, -0.2, 1.01))])
def test_quadratic_surrogate_known_minimisers(p, q):
    # test/test_tuner.cpp:313-335: the surrogate with coefficients q has its minimum, 0, at p
    Q, c, d = nb_linear.quadratic_surrogate_model(q)
    optimum = np.linalg.solve(Q, -c)
    assert np.allclose(optimum, p, rtol=0, atol=1e-12)
    assert abs(optimum @ (0.5 * Q @ optimum + c) + d) < 1e-12
    assert abs(float((nb_linear.quadratic_surrogate_features([p]) @ np.asarray(q))[0])) < 1e-12    # same model


def test_quadratic_surrogate_fit_recovers_the_coefficients():
    # test/test_tuner.cpp:337-368
    q1 = np.array([1.0, 0.5, -1.0])
    p1 = np.arange(-3.0, 4.0).reshape(7, 1)
    y1 = q1[0] + q1[1] * p1[:, 0] + q1[2] * p1[:, 0] ** 2
    assert np.allclose(np.linalg.lstsq(nb_linear.quadratic_surrogate_features(p1), y1, rcond=None)[0], q1, atol=1e-12)
    q2 = np.array([1.0, 0.5, 1.5, 2.0, -1.0, -1.0])
    p2 = np.array([[a, b] for a in range(-2, 3) for b in range(-2, 3)], dtype=np.float64)
    y2 = q2[0] + q2[1] * p2[:, 0] + q2[2] * p2[:, 1] + q2[3] * p2[:, 0] ** 2 + q2[4] * p2[:, 0] * p2[:, 1] \
        + q2[5] * p2[:, 1] ** 2
    assert np.allclose(np.linalg.lstsq(nb_linear.quadratic_surrogate_features(p2), y2, rcond=None)[0], q2, atol=1e-12)


def test_ridge_fit_with_the_surrogate_tuner(oracle):
    X, Y, W, b = linear_datasource(2, 100, 5)
    model = nb_linear.LinearModel("ridge", "standard")
    result = model.fit(None, X, Y, "mse", solver=cpu_solver(oracle), folds=2, tuner="surrogate",
                       _backend=OracleBackend(oracle))
    assert result.optimum_params()["l2reg"] <= 1e-4 and np.max(np.abs(model.weights() - W)) < 1e-3


def test_fit_result_bookkeeping():
    result = nb_linear.FitResult(("l1reg", "l2reg"), 2)
    result.add([[1.0, 1.0], [10.0, 1.0]])
    result.valid[0, :, 0], result.valid[1, :, 0] = (0.5, 0.3), (0.1, 0.2)
    result.add([[9.0, 2.0]])
    result.valid[2, :, 0] = (0.4, 0.4)
    assert result.trials() == 3 and result.closest_trial(np.array([9.0, 2.0]), 2) == 1    # result.cpp:82-100
    assert result.optimum_trial() == 1 and result.optimum_params() == {"l1reg": 10.0, "l2reg": 1.0}
    with pytest.raises(RuntimeError, match="unknown model id"):
        nb_linear.LinearModel("bayesian")


# ----------------------------------------------------------------------------------------------------------------
# the orchestration over the oracle backend (CPU)
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scaling", ["none", "mean", "minmax", "standard"])
def test_ordinary_fit_recovers_the_generating_model(oracle, scaling):
    # test/test_linear_ordinary.cpp: mse + lbfgs, weights and bias to 1e-6 whatever the scaling
    X, Y, W, b = linear_datasource(1, 120, 6, T=2)
    model = nb_linear.LinearModel("ordinary", scaling)
    result = model.fit(None, X, Y, "mse", solver=cpu_solver(oracle), _backend=OracleBackend(oracle))
    # no hyper-parameters: ml::tune still runs one parameter-free trial over the folds (src/machine/tune.cpp:48-52)
    assert result.trials() == 1 and result.optimum_params() == {} and result.optimum_trial() == 0
    assert result.train.shape == (1, 5, 2) and np.all(np.isfinite(result.train)) and np.all(np.isfinite(result.valid))
    assert result.value(0) < 1e-5
    assert np.max(np.abs(model.weights() - W)) < 1e-6 and np.max(np.abs(model.bias() - b)) < 1e-6
    assert np.max(result.refit["errors"]) < 1e-5 and result.refit["state"]["status"] == "gradient_test"
    outputs = model.predict(None, X, _backend=OracleBackend(oracle))
    assert np.max(np.abs(outputs - Y)) < 1e-5
    values = model.evaluate(None, X, Y, "mse", _backend=OracleBackend(oracle))
    assert np.array_equal(values, np.stack([result.refit["errors"], result.refit["values"]]))
    importance = nb_linear.feature_importance(model.weights())
    assert nb_linear.sparsity_ratio(importance, 1e-5) == 0.5          # every second feature is irrelevant


def test_ridge_fit_tunes_l2reg_and_recovers_the_model(oracle):
    # test/test_linear_ridge.cpp: the tuner settles on a negligible l2reg for noise-free targets
    X, Y, W, b = linear_datasource(2, 100, 5)
    model = nb_linear.LinearModel("ridge", "standard")
    result = model.fit(None, X, Y, "mse", solver=cpu_solver(oracle), folds=2, _backend=OracleBackend(oracle))
    assert result.param_names == ("l2reg",) and 3 <= result.trials() <= 31
    assert result.optimum_params()["l2reg"] <= 1e-4
    assert np.all(np.isfinite(result.valid)) and np.all(np.isfinite(result.train))
    assert np.max(np.abs(model.weights() - W)) < 1e-3 and np.max(np.abs(model.bias() - b)) < 1e-2
    # a heavier regulariser validates worse than the optimum
    heavy = int(np.argmax(result.params[:, 0]))
    assert result.value(heavy) > result.value(result.optimum_trial())


def test_ridge_fit_with_batched_trials_matches_the_sequential_fit(oracle):
    # N3 inside the tuner: the trials of each fold advance together (solver.lbfgs_batch over vgrad_batch)
    X, Y, W, b = linear_datasource(2, 200, 6, noise=0.05)
    sequential, batched = nb_linear.LinearModel("ridge"), nb_linear.LinearModel("ridge")
    r_seq = sequential.fit(None, X, Y, "mse", solver=cpu_solver(oracle), folds=2, _backend=OracleBackend(oracle))
    OracleLinearFunction.batch_calls = 0
    r_bat = batched.fit(None, X, Y, "mse", solver=cpu_solver(oracle), folds=2, batch_trials=True,
                        _backend=OracleBackend(oracle))
    assert OracleLinearFunction.batch_calls > 0
    # same trials up to ties between validation errors that differ by less than the solvers' precision (1e-8)
    common = min(r_bat.trials(), r_seq.trials())
    assert common >= 5 and np.array_equal(r_bat.params[:5], r_seq.params[:5])
    for trial in range(r_bat.trials()):
        match = np.flatnonzero(r_seq.params[:, 0] == r_bat.params[trial, 0])
        if match.size:
            assert np.max(np.abs(r_bat.valid[trial] - r_seq.valid[match[0]])) < 1e-6
    assert abs(r_bat.value(r_bat.optimum_trial()) - r_seq.value(r_seq.optimum_trial())) < 1e-7
    assert np.max(np.abs(batched.weights() - sequential.weights())) < 1e-5
    # a non-smooth objective falls back to one solve per trial
    OracleLinearFunction.batch_calls = 0
    nb_linear.LinearModel("lasso").fit(None, X, Y, "mse", solver=cpu_solver(oracle), folds=2, tuner_max_evals=4,
                                       batch_trials=True, _backend=OracleBackend(oracle))
    assert OracleLinearFunction.batch_calls == 0


@pytest.mark.parametrize("kind,loss_id", [("lasso", "mse"), ("lasso", "mae"), ("elastic_net", "mae")])
def test_regularised_fits_pass_the_reference_checks(oracle, kind, loss_id):
    # test/test_linear_lasso.cpp / test_linear_elastic_net.cpp (disabled upstream "until an efficient optimizer ... is
    # available"): weights and bias recovered to 1e-6, relevant features important (> 1e-1), irrelevant ones out (< 1e-6)
    X, Y, W, b = linear_datasource(3, 100, 4)
    model = nb_linear.LinearModel(kind, "standard")
    kwargs = {"tuner_max_evals": 12} if kind == "elastic_net" else {}
    result = model.fit(None, X, Y, loss_id, solver=cpu_solver(oracle), folds=2, _backend=OracleBackend(oracle), **kwargs)
    assert result.refit["state"]["status"] in ("specific_test", "gradient_test")
    assert np.max(np.abs(model.weights() - W)) < 1e-6 and np.max(np.abs(model.bias() - b)) < 1e-6
    importance = nb_linear.feature_importance(model.weights())
    relevant = np.abs(W[0]) > 0
    assert np.all(importance[relevant] > 1e-1) and np.all(importance[~relevant] < 1e-6)
    assert nb_linear.sparsity_ratio(importance) == 0.5                      # check_importance (fixture/linear.h:134-158)


def test_lasso_fit_with_a_nonsmooth_loss(oracle):
    # test/test_linear_lasso.cpp with mae: RQB drives the non-smooth objective; irrelevant features stay out
    X, Y, W, b = linear_datasource(3, 80, 4)
    model = nb_linear.LinearModel("lasso", "standard")
    result = model.fit(None, X, Y, "mae", solver=cpu_solver(oracle), folds=2, tuner_max_evals=8,
                       _backend=OracleBackend(oracle))
    assert result.param_names == ("l1reg",) and result.trials() >= 3
    assert np.max(np.abs(model.weights() - W)) < 1e-2 and abs(model.bias()[0] - b[0]) < 5e-2
    assert np.mean(result.refit["errors"]) < 1e-2


# ----------------------------------------------------------------------------------------------------------------
# GPU: the product path
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_evaluate_is_zero_for_the_generating_model(ctx):
    # test/test_linear_util.cpp:63-80 on the device
    import libnano_b200 as nb

    rng = np.random.default_rng(9)
    X, W, b = rng.uniform(-1, 1, (20, 13)), rng.uniform(-1, 1, (3, 13)), rng.uniform(-1, 1, 3)
    values = nb.Dataset.pin(ctx, X, X @ W.T + b).evaluate(nb.Loss("mse"), W, b)
    assert values.shape == (2, 20) and np.max(np.abs(values)) < 1e-12      # |outputs| <= 14: rounding is ~1e-15


@pytest.mark.gpu
@pytest.mark.parametrize("scaling", ["none", "standard"])
def test_gpu_ordinary_fit_recovers_the_generating_model(ctx, scaling):
    X, Y, W, b = linear_datasource(1, 2000, 24, T=2)
    model = nb_linear.LinearModel("ordinary", scaling)
    result = model.fit(ctx, X, Y, "mse", solver=lambda f, x0: nb_solver.lbfgs(f, x0, epsilon=1e-10, max_evals=5000))
    assert result.refit["state"]["status"] == "gradient_test"
    assert np.max(np.abs(model.weights() - W)) < 1e-6 and np.max(np.abs(model.bias() - b)) < 1e-6
    assert np.max(np.abs(model.predict(ctx, X) - Y)) < 1e-5 and np.max(result.refit["errors"]) < 1e-5


@pytest.mark.gpu
def test_gpu_ridge_fit_matches_the_oracle_backed_fit(ctx, oracle):
    # same orchestration, same splits: the product path and the oracle backend pick the same l2reg and agree on the
    # validation statistics and the model
    X, Y, W, b = linear_datasource(2, 400, 8, noise=0.05)
    gpu, cpu = nb_linear.LinearModel("ridge"), nb_linear.LinearModel("ridge")
    r_gpu = gpu.fit(ctx, X, Y, "mse", folds=2, solver=lambda f, x0: nb_solver.lbfgs(f, x0, epsilon=1e-10, max_evals=5000))
    r_cpu = cpu.fit(None, X, Y, "mse", folds=2, solver=cpu_solver(oracle), _backend=OracleBackend(oracle))
    assert r_gpu.trials() == r_cpu.trials() and np.array_equal(r_gpu.params, r_cpu.params)
    assert r_gpu.optimum_params() == r_cpu.optimum_params()
    assert np.max(np.abs(r_gpu.valid - r_cpu.valid) / (1 + np.abs(r_cpu.valid))) < 1e-6
    assert np.max(np.abs(gpu.weights() - cpu.weights())) < 1e-6 and np.max(np.abs(gpu.bias() - cpu.bias())) < 1e-6
    assert np.max(np.abs(gpu.weights() - W)) < 0.05


@pytest.mark.gpu
def test_gpu_elastic_net_classification_with_the_default_solvers(ctx):
    # s-hinge + elastic net is non-smooth: default_solver picks the host-driven RQB over the CUDA vgrad; the tuner
    # minimises the validation 0-1 error (device-side linear::evaluate)
    X, Y, W, b = linear_datasource(4, 600, 6, classification=True)
    model = nb_linear.LinearModel("elastic_net", "standard")
    result = model.fit(ctx, X, Y, "s-hinge", folds=2, tuner_max_evals=6)
    assert result.param_names == ("l1reg", "l2reg") and result.trials() >= 1
    assert np.all((result.valid[:, :, 0] >= 0.0) & (result.valid[:, :, 0] <= 1.0))
    assert np.mean(result.refit["errors"]) < 0.1                              # 0-1 error of the refit model
    predictions = np.sign(model.predict(ctx, X))
    assert np.mean(predictions != Y) == np.mean(result.refit["errors"])
