"""loss_t::error — the reference's known-answer tables (test/test_loss.cpp:150-269: single_class, single_label_multi_class,
multi_label_multi_class, regression), the golden vectors its own error functors produced here (include/nano/loss/error.h
compiled in place, tests/golden/make_ref_golden.py) and linear::evaluate (src/linear/util.cpp:30-49). The CPU oracle is
pinned on both; the CUDA kernels are checked against the same tables, vectors and the oracle."""
import json
import os

import numpy as np
import pytest

from conftest import CLASSIFICATION_LOSSES, REGRESSION_LOSSES, make_problem

HERE = os.path.dirname(os.path.abspath(__file__))
S_LOSSES = ("s-hinge", "s-squared-hinge", "s-classnll", "s-savage", "s-tangent", "s-logistic", "s-exponential")
M_LOSSES = ("m-hinge", "m-squared-hinge", "m-savage", "m-tangent", "m-logistic", "m-exponential")
ERROR_LOSS = {"absdiff": "mse", "sclass": "s-classnll", "mclass": "m-logistic"}


def class_target(n_labels, *indices):
    """include/nano/loss/class.h:51-58: -1 everywhere, +1 at the given labels (out-of-range indices are ignored)"""
    target = -np.ones(n_labels)
    for index in indices:
        if 0 <= index < n_labels:
            target[index] = 1.0
    return target


def reference_tables():
    """(loss ids, targets, outputs, expected errors) of test/test_loss.cpp:150-251"""
    single = (S_LOSSES,
              np.array([class_target(1), class_target(1, 0), class_target(1), class_target(1, 0)]),
              np.array([class_target(1), class_target(1, 0), class_target(1, 0), class_target(1)]),
              [0.0, 0.0, 1.0, 1.0])
    outputs = np.array([class_target(13, 11), class_target(13, 12), class_target(13, 11), class_target(13)])
    outputs[2, 7] = 2.0
    multi_class = (S_LOSSES, np.array([class_target(13, 11)] * 4), outputs, [0.0, 1.0, 1.0, 1.0])
    multi_label = (M_LOSSES, np.array([class_target(13, 7, 9)] * 6),
                   np.array([class_target(13, 7, 9), class_target(13), class_target(13, 5), class_target(13, 7),
                             class_target(13, 5, 9), class_target(13, 7, 9, 11)]),
                   [0.0, 2.0, 3.0, 1.0, 2.0, 1.0])
    return {"single_class": single, "single_label_multi_class": multi_class, "multi_label_multi_class": multi_label}


def golden_errors():
    with open(os.path.join(HERE, "golden", "ref_loss_golden.json")) as handle:
        payload = json.load(handle)
    unhex = lambda values, shape: np.array([float.fromhex(v) for v in values]).reshape(shape)  # noqa: E731
    return [(c["error"], c["T"], unhex(c["targets"], (c["N"], c["T"])), unhex(c["outputs"], (c["N"], c["T"])),
             unhex(c["errors"], (c["N"],))) for c in payload["errors"]]


TABLES = reference_tables()
GOLDEN = golden_errors()
GOLDEN_IDS = [f"{c[0]}-T{c[1]}" for c in GOLDEN]


# ----------------------------------------------------------------------------------------------------------------
# CPU: the oracle against the reference
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("table", sorted(TABLES))
def test_oracle_reproduces_the_reference_error_tables(oracle, table):
    loss_ids, targets, outputs, expected = TABLES[table]
    for loss_id in loss_ids:
        assert np.array_equal(oracle.loss_error(loss_id, targets, outputs), expected), loss_id


def test_oracle_regression_error_is_zero_on_the_targets(oracle):
    # test/test_loss.cpp:253-269
    target = np.random.default_rng(0).uniform(-1, 1, (3, 4))
    for loss_id in ("mae", "mse", "cauchy"):
        assert np.max(np.abs(oracle.loss_error(loss_id, target, target.copy()))) < 1e-15


@pytest.mark.parametrize("name,T,targets,outputs,errors", GOLDEN, ids=GOLDEN_IDS)
def test_oracle_reproduces_the_reference_error_functors(oracle, name, T, targets, outputs, errors):
    got = oracle.loss_error(ERROR_LOSS[name], targets, outputs)
    if name == "absdiff":
        assert np.max(np.abs(got - errors) / (1 + errors)) <= 4e-16
    else:
        assert np.array_equal(got, errors)      # counts: exact


def test_golden_error_cases_cover_the_edges():
    assert sorted({c[0] for c in GOLDEN}) == ["absdiff", "mclass", "sclass"] and len(GOLDEN) == 9
    for name, T, targets, outputs, errors in GOLDEN:
        if name != "absdiff":
            assert errors[1] == 0.0 or (name == "sclass" and T > 1)   # target * output == epsilon is not an error
            assert set(np.unique(errors)) <= set(np.arange(T + 1.0))


def test_oracle_pinball_error_is_its_value(oracle):
    rng = np.random.default_rng(1)
    targets, outputs = rng.normal(size=(7, 2)), rng.normal(size=(7, 2))
    for alpha in (0.2, 0.5, 0.8):
        assert np.array_equal(oracle.loss_error("pinball", targets, outputs, alpha),
                              oracle.loss_value("pinball", targets, outputs, alpha))


def test_oracle_evaluate_matches_predict_then_loss(oracle):
    X, Y, x = make_problem(3, 50, 7, 3, True)
    W, b = x[:21].reshape(3, 7), x[21:]
    values = oracle.evaluate(X, Y, "s-classnll", W, b)
    outputs = X @ W.T + b
    assert values.shape == (2, 50)
    assert np.array_equal(values[0], (Y[np.arange(50), np.argmax(outputs, axis=1)] <= 0).astype(float))
    assert np.max(np.abs(values[1] - oracle.loss_value("s-classnll", Y, outputs))) <= 1e-13


def test_error_kind_of_every_registered_id():
    from libnano_b200 import LOSS_IDS, Loss
    from libnano_b200.loss import ERROR_ABSDIFF, ERROR_MCLASS, ERROR_SCLASS

    for loss_id in LOSS_IDS:
        expected = ERROR_SCLASS if loss_id.startswith("s-") else ERROR_MCLASS if loss_id.startswith("m-") else ERROR_ABSDIFF
        assert Loss(loss_id).error_kind == expected


# ----------------------------------------------------------------------------------------------------------------
# GPU: the CUDA kernels through the C ABI
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("table", sorted(TABLES))
def test_cuda_error_kernel_reproduces_the_reference_tables(ctx, table):
    import libnano_b200 as nb

    loss_ids, targets, outputs, expected = TABLES[table]
    for loss_id in loss_ids:
        assert np.array_equal(nb.Loss(loss_id).error(ctx, targets, outputs), expected), loss_id


@pytest.mark.gpu
@pytest.mark.parametrize("name,T,targets,outputs,errors", GOLDEN, ids=GOLDEN_IDS)
def test_cuda_error_kernel_matches_the_reference_error_functors(ctx, name, T, targets, outputs, errors):
    import libnano_b200 as nb

    got = nb.Loss(ERROR_LOSS[name]).error(ctx, targets, outputs)
    if name == "absdiff":
        assert np.max(np.abs(got - errors) / (1 + errors)) <= 1e-15
    else:
        assert np.array_equal(got, errors)


@pytest.mark.gpu
@pytest.mark.parametrize("loss_id", CLASSIFICATION_LOSSES + REGRESSION_LOSSES)
@pytest.mark.parametrize("N,T", [(1, 1), (257, 1), (1000, 5), (3, 100)])
def test_cuda_error_kernel_matches_the_oracle(ctx, oracle, loss_id, N, T):
    import libnano_b200 as nb

    if loss_id == "s-classnll" and T == 1:
        pytest.skip("classnll needs two labels")
    rng = np.random.default_rng(N * 131 + T)
    outputs = rng.normal(size=(N, T))
    if loss_id in REGRESSION_LOSSES:
        targets = rng.normal(size=(N, T))
    elif loss_id.startswith("s-") and T > 1:
        targets = -np.ones((N, T))
        targets[np.arange(N), rng.integers(0, T, N)] = 1.0
    else:
        targets = np.where(rng.uniform(-1, 1, (N, T)) < 0, -1.0, 1.0)
    got = nb.Loss(loss_id).error(ctx, targets, outputs)
    ref = oracle.loss_error(loss_id, targets, outputs)
    assert np.max(np.abs(got - ref) / (1 + np.abs(ref))) <= 1e-15, loss_id


@pytest.mark.gpu
def test_cuda_error_contract(ctx):
    import libnano_b200 as nb

    loss = nb.Loss("s-hinge")
    assert loss.error(ctx, np.zeros((0, 3)), np.zeros((0, 3))).shape == (0,)
    with pytest.raises(RuntimeError, match="incompatible dimension-wise"):
        loss.error(ctx, np.zeros((4, 3)), np.zeros((4, 2)))
    pinball = nb.Loss("pinball")
    pinball.parameter("loss::pinball::alpha", 0.2)
    rng = np.random.default_rng(2)
    targets, outputs = rng.normal(size=(9, 2)), rng.normal(size=(9, 2))
    assert np.array_equal(pinball.error(ctx, targets, outputs), pinball.value(ctx, targets, outputs))


@pytest.mark.gpu
@pytest.mark.parametrize("N,D,T,loss_id", [(1000, 33, 1, "s-logistic"), (5000, 64, 7, "s-classnll"),
                                           (777, 20, 4, "m-hinge"), (2048, 100, 3, "mse"), (500, 16, 1, "pinball"),
                                           (1, 5, 2, "mae")])
def test_evaluate_matches_the_oracle(ctx, oracle, N, D, T, loss_id):
    # linear::evaluate over the pinned dataset: outputs stay in HBM, (errors, values) come back
    import libnano_b200 as nb

    classification = loss_id[:2] in ("s-", "m-")
    X, Y, x = make_problem(N + D, N, D, T, classification)
    if loss_id.startswith("m-"):
        Y = np.where(np.random.default_rng(5).uniform(-1, 1, (N, T)) < 0, -1.0, 1.0)
    W, b = x[: D * T].reshape(T, D), x[D * T:]
    data = nb.Dataset.pin(ctx, X, Y)
    got = data.evaluate(nb.Loss(loss_id), W, b)
    ref = oracle.evaluate(X, Y, loss_id, W, b)
    assert got.shape == (2, N)
    if classification:
        # a 0-1 error can only differ where an edge / a margin between two outputs is at rounding level
        outputs = X @ W.T + b
        margin = np.min(np.abs(outputs), axis=1) if T == 1 or loss_id.startswith("m-") else \
            np.sort(outputs, axis=1)[:, -1] - np.sort(outputs, axis=1)[:, -2]
        safe = margin > 1e-12
        assert safe.mean() > 0.99 and np.array_equal(got[0][safe], ref[0][safe])
    else:
        assert np.max(np.abs(got[0] - ref[0]) / (1 + np.abs(ref[0]))) <= 1e-12
    assert np.max(np.abs(got[1] - ref[1]) / (1 + np.abs(ref[1]))) <= 1e-12
    # the statistics the tuner consumes (src/machine/result.cpp:136-148): mean error / mean loss
    assert abs(got[0].mean() - ref[0].mean()) <= 1e-12 * (1 + abs(ref[0].mean()))
