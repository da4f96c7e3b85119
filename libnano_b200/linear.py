"""Mirror of nano::linear_t::fit (src/linear.cpp:30-48,79-116) over the GPU objective — SURVEY §8f row N1: the caller
either side of the hot path. One fit = pin the selected samples in HBM, scale them there (flatten_iterator_t::scaling),
minimise linear::function_t with a solver that drives the CUDA vgrad, upscale (W, b) back to the original units
(nano::upscale) and evaluate the per-sample (error, loss) on the device (linear::evaluate). Around it: the
hyper-parameter tuning of ml::tune (src/machine/tune.cpp:8-58) — k-fold splits (src/splitter/kfold.cpp), the coarse
grid + local search of the tuner (src/tuner.cpp:17-42, src/tuner/local.cpp, src/tuner/util.cpp:60-118), warm starts
from the closest earlier trial (src/machine/result.cpp:82-100) — and the final refit on all samples.

The four registered models (src/linear.cpp:124-139) differ in which regularisers they tune (src/linear/{ordinary,
lasso,ridge,elastic_net}.cpp). What is NOT mirrored: serialization and the file loggers."""
from __future__ import annotations

import itertools

import numpy as np

from .function import Dataset, LinearFunction, critical
from .loss import ERROR_ABSDIFF, Loss

# linear::make_param_space (src/linear/util.cpp:75-86)
PARAM_GRID = tuple(float(f"{m}e{e}") for e in range(-8, 7) for m in (1, 3)) + (1e7,)
# make_param_spaces of the registered ids
MODELS = {"ordinary": (), "lasso": ("l1reg",), "ridge": ("l2reg",), "elastic_net": ("l1reg", "l2reg")}


def feature_importance(weights, column2feature=None):
    """linear::feature_importance (src/linear/util.cpp:51-63): the weight magnitude per feature, cumulated over the
    targets and over the flatten columns of the feature (`column2feature[column]`, dataset_t::column2feature; one
    column per feature when omitted)."""
    per_column = np.abs(np.atleast_2d(np.asarray(weights, dtype=np.float64))).sum(axis=0)
    if column2feature is None:
        return per_column
    column2feature = np.asarray(column2feature, dtype=np.int64)
    critical(column2feature.size == per_column.size and column2feature.min() >= 0,
             "linear: one feature index per flatten column is needed")
    importance = np.zeros(int(column2feature.max()) + 1)
    np.add.at(importance, column2feature, per_column)
    return importance


def sparsity_ratio(importance, threshold: float = 1e-6) -> float:
    """linear::sparsity_ratio (src/linear/util.cpp:65-73): the fraction of the features below the threshold."""
    importance = np.asarray(importance, dtype=np.float64)
    critical(threshold > 0.0 and importance.size > 0 and importance.min() >= 0.0, "linear: invalid feature importance")
    return float(np.count_nonzero(importance < threshold)) / float(importance.size)


def kfold_split(samples, folds: int = 5, seed: int = 42):
    """kfold_splitter_t::split (src/splitter/kfold.cpp:11-45): shuffle, cut into `folds` validation chunks (the last
    one takes the remainder), sort both parts. The shuffle is numpy's, not libstdc++'s std::shuffle(mt19937_64)."""
    samples = np.array(samples, dtype=np.int64)
    critical(folds >= 2 and samples.size >= folds, "splitter: at least two folds and one sample per fold are needed")
    shuffled = np.random.default_rng(seed).permutation(samples)
    chunk = samples.size // folds
    splits = []
    for fold in range(folds):
        begin = fold * chunk
        end = begin + chunk if fold + 1 < folds else samples.size
        valid = np.sort(shuffled[begin:end])
        train = np.sort(np.concatenate([shuffled[:begin], shuffled[end:]]))
        splits.append((train, valid))
    return splits


def random_split(samples, folds: int = 5, seed: int = 42, train_per: int = 80):
    """random_splitter_t::split (src/splitter/random.cpp:13-46): `folds` independent shuffles, the first train_per
    percent train, the rest validate; both parts sorted. (numpy's shuffle, as in kfold_split.)"""
    samples = np.array(samples, dtype=np.int64)
    critical(10 <= train_per <= 90, "splitter: train_per out of [10, 90]")
    train_size = (train_per * samples.size + 50) // 100          # idiv: rounds to nearest (numeric.h)
    rng = np.random.default_rng(seed)
    splits = []
    for _ in range(folds):
        samples = rng.permutation(samples)
        splits.append((np.sort(samples[:train_size]), np.sort(samples[train_size:])))
    return splits


SPLITTERS = {"k-fold": kfold_split, "random": random_split}


def local_search(min_igrid, max_igrid, src_igrid, radius: int):
    """src/tuner/util.cpp:60-83: the 3^n neighbours at +-radius around src (itself included), inside the grid."""
    igrids = []
    for offsets in itertools.product((-1, 0, 1), repeat=len(src_igrid)):
        igrid = tuple(s + o * radius for s, o in zip(src_igrid, offsets))
        if all(lo <= i <= hi for lo, i, hi in zip(min_igrid, igrid, max_igrid)):
            igrids.append(igrid)
    return igrids


class ParamSpace(tuple):
    """param_space_t (src/tuner/space.cpp): a sorted grid of distinct values searched on a linear or a log10 scale. A
    plain tuple of grid values is a log10 space (what linear::make_param_space builds)."""

    def __new__(cls, values, scale: str = "log10"):
        values = tuple(float(v) for v in values)
        critical(len(values) >= 2, "parameter space: at least two grid values must be given!")
        critical(all(a < b for a, b in zip(values, values[1:])), "parameter space: the grid values must be sorted and "
                 "distinct!")
        critical(scale in ("linear", "log10"), "parameter space: unknown scale")
        critical(scale != "log10" or values[0] >= np.finfo(float).eps,
                 "parameter space: the grid values must be strictly positive if using the logarithmic scale!")
        self = super().__new__(cls, values)
        self.scale = scale
        return self

    def to_surrogate(self, value: float) -> float:
        critical(self[0] <= value <= self[-1], "parameter space: cannot map value (", value,
                 ") outside the parameter grid range [", self[0], ",", self[-1], "]!")
        return (value - self[0]) / (self[-1] - self[0]) if self.scale == "linear" else float(np.log10(value))

    def closest_grid_point_from_surrogate(self, value: float) -> int:
        return int(np.argmin([abs(value - self.to_surrogate(v)) for v in self]))      # the first of equal distances


def quadratic_surrogate_features(p):
    """quadratic_surrogate_fit_t (src/tuner/surrogate.cpp:8-41): [1 | p_i | p_i p_j (i <= j)] per sample."""
    p = np.atleast_2d(np.asarray(p, dtype=np.float64))
    size = p.shape[1]
    columns = [np.ones(p.shape[0])] + [p[:, i] for i in range(size)]
    columns += [p[:, i] * p[:, j] for i in range(size) for j in range(i, size)]
    return np.stack(columns, axis=1)


def quadratic_surrogate_model(q):
    """quadratic_surrogate_t (src/tuner/surrogate.cpp:73-100): q -> (Q, c, d) with value x.Q.x / 2 + c.x + d."""
    q = np.asarray(q, dtype=np.float64)
    size = int(np.sqrt(2 * q.size)) - 1
    critical(size > 0 and q.size == (size + 1) * (size + 2) // 2, "tuner: invalid surrogate model")
    Q = np.zeros((size, size))
    k = size + 1
    for i in range(size):
        for j in range(i, size):
            Q[i, j] = Q[j, i] = q[k]
            k += 1
    Q[np.diag_indices(size)] *= 2.0
    return Q, q[1:size + 1].copy(), float(q[0])


def _surrogate_candidate(spaces, steps):
    """One iteration of surrogate_tuner_t::do_optimize (src/tuner/surrogate.cpp:139-186): fit the quadratic to every
    evaluated step (least squares = the mse fit the reference runs L-BFGS on, minimum norm while under-determined),
    minimise it, return the closest grid point. A surrogate that is not convex has no minimiser (the reference's
    L-BFGS then stops wherever its budget ends): the grid point with the smallest surrogate value is taken instead."""
    p = np.array([[space.to_surrogate(v) for space, v in zip(spaces, step[1])] for step in steps])
    y = np.array([step[2] for step in steps])
    q = np.linalg.lstsq(quadratic_surrogate_features(p), y, rcond=None)[0]
    critical(np.all(np.isfinite(q)), "tuner: failed to fit the surrogate model!")
    Q, c, _ = quadratic_surrogate_model(q)
    try:
        np.linalg.cholesky(Q)
        optimum = np.linalg.solve(Q, -c)
        return tuple(space.closest_grid_point_from_surrogate(x) for space, x in zip(spaces, optimum))
    except np.linalg.LinAlgError:
        grids = [np.array([space.to_surrogate(v) for v in space]) for space in spaces]
        best, best_value = None, np.inf
        for igrid in itertools.product(*[range(len(space)) for space in spaces]):
            x = np.array([grid[i] for grid, i in zip(grids, igrid)])
            value = float(x @ (0.5 * Q @ x + c))
            if value < best_value:
                best, best_value = igrid, value
        return best


def tune(spaces, callback, max_evals: int = 100, tuner: str = "local-search"):
    """tuner_t::optimize (src/tuner.cpp:17-42) with the local-search tuner (src/tuner/local.cpp:16-31) or the
    quadratic-surrogate tuner (src/tuner/surrogate.cpp:125-186): `callback(params[trials, n_params]) -> values[trials]`;
    returns the steps [(igrid, params, value)] sorted by value."""
    critical(len(spaces) > 0, "tuner: at least one parameter space is needed!")
    critical(tuner in ("local-search", "surrogate"), "tuner: unknown id <", tuner, ">")
    spaces = [space if isinstance(space, ParamSpace) else ParamSpace(space) for space in spaces]
    min_igrid = tuple(0 for _ in spaces)
    max_igrid = tuple(len(space) - 1 for space in spaces)
    avg_igrid = tuple(len(space) // 2 for space in spaces)
    steps = []

    def evaluate(igrids):                                      # src/tuner/util.cpp:85-118
        known = {step[0] for step in steps}
        igrids = [igrid for igrid in igrids if igrid not in known]
        if not igrids:
            return False
        params = np.array([[spaces[p][i] for p, i in enumerate(igrid)] for igrid in igrids], dtype=np.float64)
        values = np.asarray(callback(params), dtype=np.float64)
        for igrid, param, value in zip(igrids, params, values):
            critical(np.isfinite(value), "tuner: invalid value (", value, ") detected for parameters (", param, ")!")
            steps.append((igrid, param, float(value)))
        steps.sort(key=lambda step: step[2])                   # stable, like std::sort on the value only
        return True

    evaluate([avg_igrid])
    radius = 2
    while steps and len(steps) < max_evals // 2:               # coarse grid around the running optimum
        if not evaluate(local_search(min_igrid, max_igrid, steps[0][0], radius)):
            break
        radius *= 2
    while steps and len(steps) < max_evals:                    # around the optimum / the surrogate's minimiser, radius 1
        centre = steps[0][0] if tuner == "local-search" else _surrogate_candidate(spaces, steps)
        if not evaluate(local_search(min_igrid, max_igrid, centre, 1)):
            break
    return steps


class _DeviceBackend:
    """The product path: datasets pinned and scaled in HBM, the CUDA objective, device-side evaluation."""

    def __init__(self, ctx):
        self.ctx = ctx

    def pin(self, X, Y):
        return Dataset.pin(self.ctx, X, Y)

    @staticmethod
    def function(data, loss, l1reg, l2reg):
        return LinearFunction(data, loss, l1reg, l2reg)


def default_solver(function, x0):
    """The reference's tests fit smooth objectives with lbfgs and non-smooth ones with rqb
    (test/test_linear_ridge.cpp:24): device-resident L-BFGS, host-driven RQB."""
    from . import solver

    if function.smooth():
        return solver.lbfgs(function, x0, epsilon=1e-8, max_evals=1000)
    return solver.rqb(function, x0, epsilon=1e-8, max_evals=2000)


class FitResult:
    """ml::result_t in small: per trial the parameters, the per-fold mean (error, loss) on both splits, the optimum
    and the statistics of the final refit."""

    def __init__(self, param_names, folds):
        self.param_names = tuple(param_names)
        self.folds = folds
        self.params = np.zeros((0, len(self.param_names)))
        self.train = np.zeros((0, folds, 2))                   # [trial, fold, (error, loss)] means
        self.valid = np.zeros((0, folds, 2))
        self.extras = []                                       # [trial][fold] -> solver solution (warm starts)
        self.refit = None

    def trials(self) -> int:
        return self.params.shape[0]

    def add(self, params):
        params = np.asarray(params, dtype=np.float64)
        # (a model without hyper-parameters still has trials: rows of zero columns)
        params = params.reshape(-1, len(self.param_names)) if self.param_names else params.reshape(params.shape[0], 0)
        count = params.shape[0]
        self.params = np.concatenate([self.params, params])
        self.train = np.concatenate([self.train, np.full((count, self.folds, 2), np.nan)])
        self.valid = np.concatenate([self.valid, np.full((count, self.folds, 2), np.nan)])
        self.extras.extend([[None] * self.folds for _ in range(count)])

    def closest_trial(self, params, max_trials: int) -> int:  # src/machine/result.cpp:82-100
        best, best_distance = 0, np.finfo(float).max
        for trial in range(max_trials):
            distance = float(np.linalg.norm(self.params[trial] - params))
            if distance < best_distance:
                best, best_distance = trial, distance
        return best

    def value(self, trial: int) -> float:                      # mean over folds of the mean validation error
        return float(np.mean(self.valid[trial, :, 0]))

    def optimum_trial(self) -> int:                            # src/machine/result.cpp:64-80
        return int(np.argmin([self.value(trial) for trial in range(self.trials())]))

    def optimum_params(self) -> dict:
        if not self.param_names:
            return {}
        return dict(zip(self.param_names, self.params[self.optimum_trial()]))


class LinearModel:
    """nano::linear_t: `fit` tunes the model's regularisation parameters by cross-validation and refits on all the
    samples; `predict` and `evaluate` run on the device as well. `kind`: ordinary | lasso | ridge | elastic_net."""

    def __init__(self, kind: str = "ordinary", scaling: str = "standard"):
        critical(kind in MODELS, "linear: unknown model id <", kind, ">")
        critical(scaling in Dataset.SCALINGS, "linear: unhandled scaling type")
        self.kind, self.scaling = kind, scaling
        self._weights = self._bias = None

    def weights(self):
        return self._weights

    def bias(self):
        return self._bias

    def make_param_spaces(self):
        return [PARAM_GRID for _ in MODELS[self.kind]]

    def _regularisers(self, params):
        values = dict(zip(MODELS[self.kind], params))
        return float(values.get("l1reg", 0.0)), float(values.get("l2reg", 0.0))

    def _prepare(self, backend, loss, X, Y):
        """The scaled copy of the samples the objective reads (flatten_iterator_t::scaling + cache_flatten): pinned
        and scaled in HBM once, shared by every trial on these samples."""
        data = backend.pin(X, Y)
        # class targets are never scaled (their statistics are disabled); regression targets follow the inputs
        return data.scale(self.scaling, self.scaling if loss.error_kind == ERROR_ABSDIFF else "none")

    def _fit_once(self, backend, loss, solver, data, params, x0=None):
        """::fit (src/linear.cpp:30-48) on a prepared dataset -> (W, b in original units, the solver's result)."""
        l1reg, l2reg = self._regularisers(params)
        function = backend.function(data, loss, l1reg, l2reg)
        start = np.zeros(function.size()) if x0 is None else np.asarray(x0, dtype=np.float64)
        state = solver(function, start)
        W, b = data.upscale(function.weights(state["x"]), function.bias(state["x"]))
        return W, b, state

    def fit(self, ctx, X, Y, loss, samples=None, solver=None, folds: int = 5, seed: int = 42,
            tuner_max_evals: int = 100, tuner: str = "local-search", splitter: str = "k-fold",
            batch_trials: bool | None = None, _backend=None) -> FitResult:
        """linear_t::fit (src/linear.cpp:79-116). X[N, D], Y[N, T] host arrays; `loss`: Loss or its id; `solver`:
        callable (function, x0) -> dict with "x" (default_solver); `tuner`: "local-search" or "surrogate" (the reference's
        default, src/machine/params.cpp:11); `batch_trials`: solve the trials the tuner asks for together on each fold
        with solver.lbfgs_batch - ONE pass over X per round for all of them (nb200_vgrad_batch) - when the objective
        is smooth (l1reg = 0, smooth loss, one target). None (the default) batches whenever that applies and no custom
        `solver` was given; returns the FitResult, keeps (weights, bias).

        A model without hyper-parameters ("ordinary") is still cross-validated once over the folds, like ml::tune does
        (its callback runs with one parameter-free trial, src/machine/tune.cpp:48-52), so the result always carries
        train / validation statistics."""
        backend = _backend or _DeviceBackend(ctx)
        loss = Loss(loss) if isinstance(loss, str) else loss
        if batch_trials is None:
            batch_trials = solver is None
        solver = solver or default_solver
        X = np.ascontiguousarray(X, dtype=np.float64)
        Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(X.shape[0], -1)
        samples = np.arange(X.shape[0]) if samples is None else np.asarray(samples, dtype=np.int64)
        spaces = self.make_param_spaces()
        result = FitResult(MODELS[self.kind], folds)

        if folds > 0:
            critical(splitter in SPLITTERS, "splitter: unknown id <", splitter, ">")
            splits = SPLITTERS[splitter](samples, folds, seed)
            # per fold, pinned once for all the trials: the scaled training samples (what the objective reads) and the
            # splits unscaled (the upscaled model is evaluated on them). Class targets are never scaled, so there the
            # training statistics come from the scaled copy and the model in the scaled space (the outputs are the
            # same numbers: upscale is the inverse map) and the unscaled training split is not pinned a second time.
            scaled_sets = [self._prepare(backend, loss, X[train], Y[train]) for train, _ in splits]
            targets_scaled = loss.error_kind == ERROR_ABSDIFF and self.scaling != "none"
            train_sets = [backend.pin(X[train], Y[train]) for train, _ in splits] if targets_scaled else None
            valid_sets = [backend.pin(X[valid], Y[valid]) for _, valid in splits]

            def store(trial, fold, W, b, x):
                result.extras[trial][fold] = x
                if train_sets is not None:
                    result.train[trial, fold] = train_sets[fold].evaluate(loss, W, b).mean(axis=1)
                else:
                    isize, tsize = scaled_sets[fold].inputs, scaled_sets[fold].targets
                    x = np.asarray(x)
                    result.train[trial, fold] = scaled_sets[fold].evaluate(
                        loss, x[:isize * tsize].reshape(tsize, isize), x[isize * tsize:]).mean(axis=1)
                result.valid[trial, fold] = valid_sets[fold].evaluate(loss, W, b).mean(axis=1)

            def batched_callback(new_params, old_trials):
                """All the new trials of a fold advance together: one pass over X per round (solver.lbfgs_batch)."""
                from . import solver as nb_solver

                regularisers = [self._regularisers(params) for params in new_params]
                for fold, data in enumerate(scaled_sets):
                    function = backend.function(data, loss, 0.0, 0.0)
                    x0s = np.zeros((len(new_params), function.size()))
                    for trial, params in enumerate(new_params):
                        if old_trials > 0:
                            x0s[trial] = result.extras[result.closest_trial(params, old_trials)][fold]
                    states = nb_solver.lbfgs_batch(function, x0s, 0.0, [l2 for _, l2 in regularisers], epsilon=1e-8)
                    for trial, state in enumerate(states):
                        W, b = data.upscale(function.weights(state["x"]), function.bias(state["x"]))
                        store(old_trials + trial, fold, W, b, state["x"])

            def callback(new_params):                          # ml::tune's tuner_callback (tune.cpp:19-46)
                old_trials = result.trials()
                result.add(new_params)
                # (nb200_vgrad_batch runs on one device: a device group solves its trials one by one)
                one_device = len(getattr(ctx, "devices", [0])) == 1
                batchable = (batch_trials and one_device and loss.smooth() and Y.shape[1] == 1
                             and all(self._regularisers(params)[0] == 0.0 for params in new_params))
                if batchable:
                    batched_callback(new_params, old_trials)
                    return [result.value(trial) for trial in range(old_trials, result.trials())]
                for trial, params in enumerate(new_params):
                    closest = result.closest_trial(params, old_trials) if old_trials > 0 else None
                    for fold in range(len(splits)):
                        # warm start from the closest earlier trial of this fold (tune.cpp:31-37; its solution in the
                        # scaled space, where the reference passes the upscaled weights, src/linear.cpp:16-28)
                        x0 = result.extras[closest][fold] if closest is not None else None
                        W, b, state = self._fit_once(backend, loss, solver, scaled_sets[fold], params, x0)
                        store(old_trials + trial, fold, W, b, state["x"])
                return [result.value(trial) for trial in range(old_trials, result.trials())]

            if spaces:
                tune(spaces, callback, tuner_max_evals, tuner)
            else:
                callback(np.zeros((1, 0)))                     # no hyper-parameters: one trial over the folds
            params = result.params[result.optimum_trial()]

        # refit with the optimum hyper-parameters on all the given samples (linear.cpp:99-112)
        W, b, state = self._fit_once(backend, loss, solver, self._prepare(backend, loss, X[samples], Y[samples]), params)
        values = backend.pin(X[samples], Y[samples]).evaluate(loss, W, b)
        self._weights, self._bias = W, b
        result.refit = {"state": state, "errors": values[0], "values": values[1], "params": params}
        return result

    def evaluate(self, ctx, X, Y, loss, _backend=None):
        """learner_t::evaluate (src/learner.cpp:79-93): [2, N] = per-sample (errors, loss values) of the fitted model."""
        critical(self._weights is not None, "linear: the model is not fitted")
        backend = _backend or _DeviceBackend(ctx)
        loss = Loss(loss) if isinstance(loss, str) else loss
        X = np.ascontiguousarray(X, dtype=np.float64)
        Y = np.ascontiguousarray(Y, dtype=np.float64).reshape(X.shape[0], -1)
        return backend.pin(X, Y).evaluate(loss, self._weights, self._bias)

    def predict(self, ctx, X, _backend=None):
        """linear_t::do_predict (src/linear.cpp:118-128): outputs[N, T] = X W^T + b on the device."""
        critical(self._weights is not None, "linear: the model is not fitted")
        backend = _backend or _DeviceBackend(ctx)
        X = np.ascontiguousarray(X, dtype=np.float64)
        return backend.pin(X, np.zeros((X.shape[0], self._bias.size))).predict(self._weights, self._bias)
