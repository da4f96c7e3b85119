"""app/bench_linear.cpp over the GPU path: nested cross-validation of a linear model on a synthetic linear datasource.
For each outer fold: fit (and tune) on the training samples with LinearModel.fit — every pass over the data is a CUDA
kernel (scaling, vgrad, evaluate) — then evaluate on the held-out samples. One JSON line per outer fold with the
columns of the reference's table (optimum params, train / valid / refit / test error) plus timings, and one line with
the throughput of linear::evaluate on its own.

    python tools/bench_linear.py --linear ridge --loss mse --rows 262144 --features 256"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--linear", default="ridge", choices=["ordinary", "lasso", "ridge", "elastic_net"])
    parser.add_argument("--loss", default="mse")
    parser.add_argument("--rows", type=int, default=1 << 18)
    parser.add_argument("--features", type=int, default=256)
    parser.add_argument("--outer-folds", type=int, default=2)
    parser.add_argument("--folds", type=int, default=2)
    parser.add_argument("--tuner-max-evals", type=int, default=100)
    parser.add_argument("--tuner", default="local-search", choices=["local-search", "surrogate"])
    parser.add_argument("--splitter", default="k-fold", choices=["k-fold", "random"])
    parser.add_argument("--batch-trials", action="store_true",
                        help="smooth objectives: the trials of a fold advance together, one pass over X per round")
    parser.add_argument("--noise", type=float, default=0.1)
    args = parser.parse_args()

    import libnano_b200 as nb
    from libnano_b200 import linear

    ctx = nb.Context(0)
    rng = np.random.default_rng(42)
    N, D = args.rows, args.features
    X = rng.uniform(-1.0, 1.0, (N, D)) * rng.uniform(0.5, 20.0, D)
    W = rng.uniform(-1.0, 1.0, (1, D))
    W[:, np.arange(D) % 2 != 0] = 0.0                             # every second feature is irrelevant
    Y = X @ W.T + 0.3 + args.noise * rng.normal(size=(N, 1))
    loss = nb.Loss(args.loss)
    if loss.error_kind != 0:
        Y = np.where(Y - np.median(Y) < 0, -1.0, 1.0)

    # linear::evaluate on its own: predict + error + value over the pinned samples
    data = nb.Dataset.pin(ctx, X, Y)
    data.evaluate(loss, W, np.array([0.3]))
    t0 = time.perf_counter()
    repeats = 5
    for _ in range(repeats):
        data.evaluate(loss, W, np.array([0.3]))
    seconds = (time.perf_counter() - t0) / repeats
    print(json.dumps({"evaluate": f"{N} x {D}", "ms_per_call": round(1e3 * seconds, 3),
                      "gbs_of_X": round(8e-9 * N * D / seconds, 1),
                      "note": "wall clock of nb200_evaluate: W, b up, three kernels, 2 N doubles back through the pinned staging buffer"}),
          flush=True)
    data.close()

    for fold, (train, valid) in enumerate(linear.SPLITTERS[args.splitter](np.arange(N), args.outer_folds, seed=42)):
        model = nb.LinearModel(args.linear, "standard")
        t0 = time.perf_counter()
        result = model.fit(ctx, X, Y, loss, samples=train, folds=args.folds, tuner_max_evals=args.tuner_max_evals,
                           tuner=args.tuner, splitter=args.splitter, batch_trials=args.batch_trials)
        fit_seconds = time.perf_counter() - t0
        test = model.evaluate(ctx, X[valid], Y[valid], loss)
        optimum = result.optimum_trial() if result.trials() else None
        importance = linear.feature_importance(model.weights())
        print(json.dumps({
            "fold": f"{fold + 1}/{args.outer_folds}", "linear": args.linear, "loss": args.loss,
            "optimum_params": {k: float(v) for k, v in result.optimum_params().items()},
            "train_error": None if optimum is None else float(np.mean(result.train[optimum, :, 0])),
            "valid_error": None if optimum is None else result.value(optimum),
            "refit_error": float(np.mean(result.refit["errors"])), "test_error": float(np.mean(test[0])),
            "trials": result.trials(), "inner_folds": args.folds, "tuner": args.tuner, "splitter": args.splitter, "batch_trials": args.batch_trials, "fit_seconds": round(fit_seconds, 3),
            "refit_status": result.refit["state"]["status"], "refit_fcalls": result.refit["state"]["fcalls"],
            "max_weight_error": float(np.max(np.abs(model.weights() - W))) if loss.error_kind == 0 else None,
            "sparsity_ratio_1e-2": linear.sparsity_ratio(importance, 1e-2),
            "train_rows": int(train.size), "features": D}), flush=True)


if __name__ == "__main__":
    main()
