"""GPU tests at scale and for the sharded path: synthetic HBM-resident datasets, two-phase (partial / finish)
evaluation, logical shards on one device, and size-independent properties at the BASELINE sizes."""

import numpy as np
import pytest

from conftest import PARITY, make_problem, rel_f, rel_g

pytestmark = pytest.mark.gpu


# -- host restatement of the counter-based generator (libnano_b200/csrc/synth.cuh) -------------------------------
def philox4x32_10(c, k):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = [np.asarray(v, dtype=np.uint64) for v in c]
    k0, k1 = np.uint64(k[0]), np.uint64(k[1])
    mask = np.uint64(0xFFFFFFFF)
    for r in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        if r < 9:
            k0, k1 = (k0 + np.uint64(W0)) & mask, (k1 + np.uint64(W1)) & mask
    return c0, c1, c2, c3


def synth_inputs_host(seed, row0, rows, D):
    e = np.arange(row0 * D, (row0 + rows) * D, dtype=np.uint64)
    c = e >> np.uint64(1)
    zero = np.zeros_like(c)
    r = philox4x32_10((c & np.uint64(0xFFFFFFFF), c >> np.uint64(32), zero, zero),
                      (seed & 0xFFFFFFFF, seed >> 32))
    odd = (e & np.uint64(1)).astype(bool)
    lo = np.where(odd, r[2], r[0])
    hi = np.where(odd, r[3], r[1])
    bits = (hi << np.uint64(32)) | lo
    u = (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return (2.0 * u - 1.0).reshape(rows, D)


@pytest.mark.parametrize("D", [64, 101, 1024])
def test_synth_inputs_bit_exact_and_shardable(ctx, D):
    import libnano_b200 as nb

    N = 777
    whole = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, N, D, seed=42)
    X, Y = whole.read()
    assert np.array_equal(X, synth_inputs_host(42, 0, N, D))
    assert set(np.unique(Y)) <= {-1.0, 1.0} and 0.2 < np.mean(Y > 0) < 0.8
    # any row range of the same dataset can be generated on its own (how ranks shard it)
    for row0, rows in ((0, 100), (100, 333), (433, 344)):
        part = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, rows, D, seed=42, row_offset=row0)
        Xp, Yp = part.read()
        assert np.array_equal(Xp, X[row0:row0 + rows]) and np.array_equal(Yp, Y[row0:row0 + rows])
    other, _ = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, N, D, seed=43).read()
    assert not np.array_equal(other, X)


def test_synth_kinds(ctx):
    import libnano_b200 as nb

    reg = nb.Dataset.synth(ctx, nb.SYNTH_REGRESSION, 500, 32, seed=7, noise=0.0)
    X, Y = reg.read()
    # noise-free regression targets are linear in X: solve for w*, b* and check the residual
    A = np.hstack([X, np.ones((500, 1))])
    sol, *_ = np.linalg.lstsq(A, Y[:, 0], rcond=None)
    assert np.max(np.abs(A @ sol - Y[:, 0])) < 1e-12 and abs(sol[-1] - 0.1) < 1e-12
    mc = nb.Dataset.synth(ctx, nb.SYNTH_MULTICLASS, 400, 48, T=10, seed=7)
    _, Ym = mc.read()
    assert np.all(np.sum(Ym > 0, axis=1) == 1) and set(np.unique(Ym)) == {-1.0, 1.0}
    assert len(np.unique(np.argmax(Ym, axis=1))) > 3
    with pytest.raises(RuntimeError):
        nb.Dataset.synth(ctx, nb.SYNTH_MULTICLASS, 10, 4, T=1)


# -- two-phase evaluation == one-shot evaluation; logical shards on one device ------------------------------------
def _read_device(ptr, length):
    import torch

    from libnano_b200.sharded import _CudaBufferView

    torch.cuda.synchronize()
    return torch.as_tensor(_CudaBufferView(ptr, length), device="cuda").cpu().numpy().copy()


@pytest.mark.parametrize("N,D,T,loss_id", [(4001, 1024, 1, "s-logistic"), (1000, 64, 4, "s-classnll"),
                                           (999, 100, 1, "s-hinge")])
def test_partial_finish_and_logical_shards(ctx, oracle, N, D, T, loss_id):
    import libnano_b200 as nb
    from libnano_b200.sharded import shard_rows

    X, Y, x = make_problem(3, N, D, T, True)
    l1, l2 = 0.05, 0.1
    whole = nb.LinearFunction(nb.Dataset.pin(ctx, X, Y), nb.Loss(loss_id), l1, l2)
    g_ref = np.empty_like(x)
    f_ref = whole(x, g_ref)

    # (1) partial + finish on the whole dataset == nb200_vgrad
    ptr, length = whole.partial(x, True)
    assert length == 1 + T + T * D
    g = np.empty_like(x)
    f = whole.finish(N, g)
    # same sums; only the order of the O(D) regulariser sums differs between the fused tail and finish_kernel
    assert rel_f(f, f_ref) <= 1e-15 and rel_g(g, g_ref) <= 1e-15

    # (2) k logical shards, partial sums added on the host in rank order, finished with N_global
    for k in (2, 3, 8):
        total = np.zeros(length)
        shards = []
        for rank in range(k):
            row0, rows = shard_rows(N, k, rank)
            fn = nb.LinearFunction(nb.Dataset.pin(ctx, X[row0:row0 + rows], Y[row0:row0 + rows]), nb.Loss(loss_id),
                                   l1, l2)
            ptr, length_k = fn.partial(x, True)
            fn.sync()
            total += _read_device(ptr, length_k)
            shards.append(fn)
        # write the reduced vector back into shard 0's buffer and finish there
        import torch
        from libnano_b200.sharded import _CudaBufferView

        ptr, _ = shards[0].partial(x, True)
        shards[0].sync()
        torch.as_tensor(_CudaBufferView(ptr, length), device="cuda").copy_(torch.from_numpy(total))
        torch.cuda.synchronize()
        g = np.empty_like(x)
        f = shards[0].finish(N, g)
        assert rel_f(f, f_ref) <= 1e-13 and rel_g(g, g_ref) <= 1e-13  # shard invariance

    fo, go = oracle.linear_vgrad(X, Y, loss_id, x, l1=l1, l2=l2)
    assert rel_f(f_ref, fo) <= PARITY and rel_g(g_ref, go) <= PARITY


def test_sharded_function_single_rank(ctx):
    """ShardedFunction with world size 1 (no process group) is the local function."""
    import libnano_b200 as nb
    from libnano_b200.sharded import ShardedFunction

    X, Y, x = make_problem(6, 2000, 128, 1, True)
    local = nb.LinearFunction(nb.Dataset.pin(ctx, X, Y), nb.Loss("s-logistic"), 0.0, 1e-2)
    g_ref = np.empty_like(x)
    f_ref = local(x, g_ref)
    sharded = ShardedFunction.from_local(local.clone(), 2000)
    g = np.empty_like(x)
    assert sharded(x, g) == f_ref and np.array_equal(g, g_ref)
    assert sharded(x) == f_ref
    assert (sharded.fcalls(), sharded.gcalls()) == (2, 1)
    assert sharded.smooth() and sharded.convex() and sharded.strong_convexity() == 1e-2 / 128


# -- BASELINE sizes: size-independent properties (the oracle only sees a bounded sub-sample) ----------------------
@pytest.mark.parametrize("N,D,loss_id", [(1 << 20, 1024, "s-logistic"), (1 << 21, 512, "s-hinge")])
def test_full_size_properties(ctx, oracle, N, D, loss_id):
    import libnano_b200 as nb
    from libnano_b200.sharded import shard_rows

    data = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, N, D, seed=42, noise=0.1)
    l2 = 1e-2
    function = nb.LinearFunction(data, nb.Loss(loss_id), 0.0, l2)
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, D + 1) / np.sqrt(D)
    gx = np.empty_like(x)
    fx = function(x, gx)
    assert np.isfinite(fx) and np.all(np.isfinite(gx))
    assert function(x) == fx or rel_f(function(x), fx) <= 1e-15  # value-only path

    # (a) the objective is a mean over samples: the mean of shard objectives (l2 = 0 part) recombines to the whole
    k = 4
    f_sum, g_sum = 0.0, np.zeros_like(x)
    for rank in range(k):
        row0, rows = shard_rows(N, k, rank)
        part = nb.Dataset.synth(ctx, nb.SYNTH_CLASSIFICATION, rows, D, seed=42, row_offset=row0, noise=0.1)
        fn = nb.LinearFunction(part, nb.Loss(loss_id), 0.0, 0.0)
        g = np.empty_like(x)
        f_sum += fn(x, g) * rows
        g_sum += g * rows
        if rank == 1:
            # (b) oracle on a bounded sub-sample of THIS shard (rows read back from HBM)
            Xs, Ys = part.read(1000, 4096)
            sub = nb.LinearFunction(nb.Dataset.pin(ctx, Xs, Ys), nb.Loss(loss_id), 0.0, 0.0)
            gs = np.empty_like(x)
            fs = sub(x, gs)
            fo, go = oracle.linear_vgrad(Xs, Ys, loss_id, x)
            assert rel_f(fs, fo) <= PARITY and rel_g(gs, go) <= PARITY
    whole0 = nb.LinearFunction(data, nb.Loss(loss_id), 0.0, 0.0)
    g0 = np.empty_like(x)
    f0 = whole0(x, g0)
    assert rel_f(f_sum / N, f0) <= 1e-13 and rel_g(g_sum / N, g0) <= 1e-13

    # (c) regulariser is additive: f(l2) - f(0) == 0.5 * l2 * mean(W^2)
    W = x[:D]
    assert abs((fx - f0) - 0.5 * l2 * np.mean(W * W)) <= 1e-14 * (1 + abs(fx))

    # (d) directional derivative: (f(x + h d) - f(x - h d)) / 2h == g . d (smooth loss only)
    if function.smooth():
        d = rng.uniform(-1, 1, D + 1)
        d /= np.linalg.norm(d)
        h = 1e-5
        fd = (function(x + h * d) - function(x - h * d)) / (2 * h)
        assert abs(fd - float(gx @ d)) <= 1e-8 * (1 + abs(fd))

    # (e) determinism at size
    g2 = np.empty_like(x)
    assert function(x, g2) == fx and np.array_equal(g2, gx)


@pytest.mark.parametrize("N,D,T,path", [(1 << 18, 2048, 100, "dmma"), (1 << 20, 784, 10, "fl-pipe"), (1 << 21, 512, 2, "ft"),
                                        (1 << 20, 1024, 8, "fl-pipe"), (1 << 20, 1000, 4, "ft"),
                                        (1 << 20, 512, 16, "fl NT8=2 NS8=8 NSETS=12"), (1 << 20, 768, 8, "fd")])
def test_full_size_properties_many_targets(ctx, oracle, N, D, T, path):
    # the multi-target paths at sizes the oracle cannot hold: BASELINE config 3's width (tensor-pipe pair) and the
    # one-pass kernels; same size-independent properties as above
    import libnano_b200 as nb
    from libnano_b200.sharded import shard_rows

    data = nb.Dataset.synth(ctx, nb.SYNTH_MULTICLASS, N, D, T=T, seed=42)
    l2 = 1e-2
    function = nb.LinearFunction(data, nb.Loss("s-classnll"), 0.0, l2)
    assert function.describe().startswith(path), function.describe()
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (D + 1) * T) / np.sqrt(D)
    gx = np.empty_like(x)
    fx = function(x, gx)
    assert np.isfinite(fx) and np.all(np.isfinite(gx))
    assert rel_f(function(x), fx) <= 1e-15

    # (a) shards recombine to the whole; (b) the oracle on a bounded sub-sample of one shard
    k = 3
    f_sum, g_sum = 0.0, np.zeros_like(x)
    for rank in range(k):
        row0, rows = shard_rows(N, k, rank)
        part = nb.Dataset.synth(ctx, nb.SYNTH_MULTICLASS, rows, D, T=T, seed=42, row_offset=row0)
        fn = nb.LinearFunction(part, nb.Loss("s-classnll"), 0.0, 0.0)
        g = np.empty_like(x)
        f_sum += fn(x, g) * rows
        g_sum += g * rows
        if rank == 1:
            Xs, Ys = part.read(777, 1500)
            sub = nb.LinearFunction(nb.Dataset.pin(ctx, Xs, Ys), nb.Loss("s-classnll"), 0.0, 0.0)
            gs = np.empty_like(x)
            fs = sub(x, gs)
            fo, go = oracle.linear_vgrad(Xs, Ys, "s-classnll", x)
            assert rel_f(fs, fo) <= PARITY and rel_g(gs, go) <= PARITY
    whole0 = nb.LinearFunction(data, nb.Loss("s-classnll"), 0.0, 0.0)
    g0 = np.empty_like(x)
    f0 = whole0(x, g0)
    assert rel_f(f_sum / N, f0) <= 1e-13 and rel_g(g_sum / N, g0) <= 1e-13

    # (c) additive regulariser; (d) directional derivative; (e) determinism; (f) the other kernels agree at size
    W = x[: T * D]
    assert abs((fx - f0) - 0.5 * l2 * np.mean(W * W)) <= 1e-14 * (1 + abs(fx))
    d = rng.uniform(-1, 1, x.size)
    d /= np.linalg.norm(d)
    h = 1e-5
    fd = (function(x + h * d) - function(x - h * d)) / (2 * h)
    assert abs(fd - float(gx @ d)) <= 1e-8 * (1 + abs(fd))
    g2 = np.empty_like(x)
    assert function(x, g2) == fx and np.array_equal(g2, gx)
    if path != "dmma":
        function.set_variant(-4)
        g3 = np.empty_like(x)
        f3 = function(x, g3)
        assert rel_f(f3, fx) <= 1e-13 and rel_g(g3, gx) <= 1e-13

