"""GPU parity tests of the gradient-boosting cousins: nb200_gboost_* against the oracle's
restatement of gboost::bias_function_t / scale_function_t / grads_function_t (src/gboost/function.cpp), at the
reference's own fixture (test/test_gboost_function.cpp) and at larger ragged shapes; optima through the solver."""
import numpy as np
import pytest

from conftest import PARITY, rel_f, rel_g
from test_gboost import fixture

pytestmark = pytest.mark.gpu

LOSSES = ["mse", "mae", "cauchy", "pinball", "s-logistic", "s-hinge", "s-classnll", "s-squared-hinge", "s-savage",
          "s-tangent", "s-exponential"]


def make(nb, ctx, kind, loss_id, targets, extra):
    loss = nb.Loss(loss_id)
    if kind == "bias":
        return nb.GBoostBiasFunction(ctx, loss, targets)
    if kind == "grads":
        return nb.GBoostGradsFunction(ctx, loss, targets)
    return nb.GBoostScaleFunction(ctx, loss, targets, extra["cluster"], extra["groups"], extra["soutputs"],
                                  extra["woutputs"])


@pytest.mark.parametrize("loss_id", LOSSES)
@pytest.mark.parametrize("kind", ["bias", "scale", "grads"])
@pytest.mark.parametrize("N,T,groups", [(100, 2, 3), (4099, 1, 7), (3001, 10, 5), (777, 40, 2)])
def test_matches_the_oracle(ctx, oracle, kind, loss_id, N, T, groups):
    import libnano_b200 as nb

    _, outputs, woutputs, targets, cluster = fixture(N, tsize=T, groups=groups, seed=N + T)
    if loss_id not in ("mse", "mae", "cauchy", "pinball"):
        targets = np.where(targets > 0, 1.0, -1.0)
        if loss_id == "s-classnll" and T > 1:
            targets = -np.ones((N, T))
            targets[np.arange(N), np.arange(N) % T] = 1.0
    extra = dict(cluster=cluster, groups=groups, soutputs=outputs, woutputs=woutputs) if kind == "scale" else {}
    function = make(nb, ctx, kind, loss_id, targets, extra)
    n = {"bias": T, "scale": groups, "grads": N * T}[kind]
    assert function.size() == n and function.type_id() == f"gboost-{kind}"
    rng = np.random.default_rng(2)
    for x in (np.zeros(n), rng.uniform(-1, 1, n)):
        gx = np.empty(n)
        fx = function(x, gx)
        f_ref, g_ref = oracle.gboost_vgrad(kind, loss_id, targets, x, **extra)
        assert rel_f(fx, f_ref) <= PARITY and rel_g(gx, g_ref) <= PARITY, (kind, loss_id, N, T)
        assert rel_f(function(x), f_ref) <= PARITY   # value-only call
    assert function.launches() > 0
    assert (function.fcalls(), function.gcalls()) == (4, 2)


def test_known_optima_of_the_reference_fixture(ctx, oracle):
    """test/test_gboost_function.cpp:123-218: the bias converges to the column means of the targets, the scales to the
    generating scales, the outputs to the targets (L-BFGS, 100 x epsilon)."""
    import libnano_b200 as nb

    scale, outputs, woutputs, targets, cluster = fixture(100)
    loss = nb.Loss("mse")
    bias = nb.GBoostBiasFunction(ctx, loss, targets)
    solved = oracle.lbfgs(bias, np.zeros(2), epsilon=1e-8)
    assert solved["status"] == "gradient_test" and np.allclose(solved["x"], targets.mean(axis=0), atol=1e-6)
    scales = nb.GBoostScaleFunction(ctx, loss, targets, cluster, 3, outputs, woutputs)
    solved = oracle.lbfgs(scales, np.zeros(3), epsilon=1e-8)
    assert solved["status"] == "gradient_test" and np.allclose(solved["x"], scale, atol=1e-6)
    grads = nb.GBoostGradsFunction(ctx, loss, targets[:10])
    solved = oracle.lbfgs(grads, np.zeros(20), epsilon=1e-9)
    assert np.allclose(solved["x"], targets[:10].reshape(-1), atol=1e-6)


def test_contract_errors(ctx):
    import libnano_b200 as nb

    _, outputs, woutputs, targets, cluster = fixture(50)
    function = nb.GBoostBiasFunction(ctx, nb.Loss("mse"), targets)
    with pytest.raises(RuntimeError, match="invalid input size"):
        function(np.zeros(3))
    assert function.fcalls() == 0
    with pytest.raises(nb.Nb200Error, match="cluster id out of range"):
        nb.GBoostScaleFunction(ctx, nb.Loss("mse"), targets, cluster, 2, outputs, woutputs)
    with pytest.raises(RuntimeError, match="must match the targets"):
        nb.GBoostScaleFunction(ctx, nb.Loss("mse"), targets, cluster, 3, outputs[:, :1], woutputs)
    assert function.convex() and function.smooth()
    assert not nb.GBoostBiasFunction(ctx, nb.Loss("mae"), targets).smooth()
