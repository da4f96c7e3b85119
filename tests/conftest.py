"""pytest configuration: `gpu` marker + shared fixtures.

`-m "not gpu"` runs on the CPU container (oracle vs the reference's known answers, host logic, ABI surface);
`-m gpu` runs on a B200 and is the parity suite proper: every GPU test calls the CUDA path through the C ABI and
checks it against the CPU oracle (oracle/), never the other way round.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    import oracle as _oracle

    _oracle.build()
    return _oracle


@pytest.fixture(scope="session")
def ctx():
    """One device context for the whole GPU session. Fails (not skips) when the CUDA library or device is missing."""
    import libnano_b200 as nb

    context = nb.Context(0)
    yield context
    context.close()


# tolerances: the reference's precision constants (include/nano/core/numeric.h:92-117) and the north-star bound
EPSILON0 = 1e-15
EPSILON1 = 1e-10
EPSILON2 = 1e-8
PARITY = 1e-12


def rel_f(f, f_ref):
    """|f - f_ref| / (1 + |f_ref|)  (SURVEY.md §8d parity protocol)"""
    return abs(f - f_ref) / (1.0 + abs(f_ref))


def rel_g(g, g_ref):
    """|g - g_ref|_inf / (1 + |g_ref|_inf)"""
    g = np.asarray(g)
    g_ref = np.asarray(g_ref)
    return float(np.max(np.abs(g - g_ref))) / (1.0 + float(np.max(np.abs(g_ref)))) if g.size else 0.0


def make_problem(seed, N, D, T, classification):
    """Seeded dense problem of the shape linear::function_t sees after cache_flatten."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1.0, 1.0, (N, D))
    if classification:
        if T == 1:
            Y = np.where(rng.uniform(-1, 1, (N, 1)) < 0, -1.0, 1.0)
        else:
            Y = -np.ones((N, T))
            Y[np.arange(N), rng.integers(0, T, N)] = 1.0
    else:
        Y = rng.uniform(-1.0, 1.0, (N, T))
    # weights scaled so that the outputs stay O(1) for any D (exp-based losses would otherwise overflow fp64)
    x = rng.uniform(-1.0, 1.0, (D + 1) * T)
    x[: D * T] *= 2.0 / np.sqrt(D)
    return X, Y, x


CLASSIFICATION_LOSSES = ("s-logistic", "m-logistic", "s-hinge", "m-hinge", "s-classnll", "s-squared-hinge",
                         "s-savage", "s-tangent", "s-exponential", "m-exponential")
REGRESSION_LOSSES = ("mse", "mae", "cauchy", "pinball")
NORTH_STAR_LOSSES = ("mse", "mae", "s-logistic", "s-hinge", "s-classnll")


def is_classification(loss_id):
    return loss_id not in REGRESSION_LOSSES
