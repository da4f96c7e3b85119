"""CPU tests of the boundary and the host logic (no compute calls: there is no GPU here).

* libnb200.so loads and exports every symbol include/nb200.h declares, and the ctypes table matches the header;
* the product fails LOUDLY without a GPU (no CPU fallback) and never touches oracle/;
* host mirror of function_t / loss_t: registry, flags, size checks, counters (reference: src/function.cpp:246-272,
  src/loss.cpp:105-135, test/test_function_factory.cpp:53-90).
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nb200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"NB200_API[^;(]*?\b(nb200_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def library():
    from libnano_b200 import build

    return ctypes.CDLL(build.build())


def test_library_exports_every_declared_symbol(library):
    symbols = header_symbols()
    assert len(symbols) >= 30
    for name in symbols:
        assert hasattr(library, name), f"libnb200.so does not export {name}"
    exported = subprocess.run(["nm", "-D", "--defined-only", library._name], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (nb200_\w+)", exported)))
    assert exported == symbols, "exported nb200_* symbols and include/nb200.h disagree"


def test_ctypes_table_matches_header():
    from libnano_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_library_is_sm100a_native(library):
    # the cubin inside is sm_100a and the hot kernel uses the bulk-copy engine + mbarrier (SASS: UBLKCP / SYNCS)
    out = subprocess.run(["cuobjdump", "-lelf", library._name], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", library._name], capture_output=True, text=True).stdout
    assert "vgrad_t1_kernel" in sass and "UBLKCP" in sass and "SYNCS.PHASECHK" in sass and "DFMA" in sass
    # every kernel family is in the library, and the many-target ones run on the FP64 tensor pipe (DMMA)
    for family in ("vgrad_ft_kernel", "vgrad_fd_kernel", "mt_forward_kernel", "mt_backward_kernel", "gen_forward_kernel",
                   "lbfgs_two_loop_kernel", "cgd_direction_kernel", "column_stats_kernel"):
        assert family in sass, family
    assert "DMMA.8x8x4" in sass and "USETMAXREG" in sass


def test_register_hand_over_matches_the_allocation(library):
    # the fused kernel re-distributes registers between its producer and consumer warpgroups (setmaxnreg); asking
    # for more than the CTA's pool would dead-lock, so pin the per-thread launch allocation ptxas chose
    out = subprocess.run(["cuobjdump", "-res-usage", library._name], capture_output=True, text=True).stdout
    found = re.findall(r"vgrad_t1_kernelILi\d+ELi\d+ELi\d+ELi(\d+)ELi\d+ELi\d+ELb[01]E\w*:\s*\n\s*REG:(\d+)", out)
    assert len(found) >= 40
    for ncw, regs in found:
        threads = (int(ncw) + 4) * 32
        launch = (65536 // threads) // 8 * 8
        if launch < 232:  # the producer warpgroup hands registers over: the pool is threads x launch registers
            assert int(regs) == launch, (ncw, regs)
        else:  # four consumer warps: every thread already owns more than setmaxnreg could give
            assert int(regs) <= 255 and int(regs) * threads <= 65536 + 255, (ncw, regs)


def test_lane_loss_pipeline_register_requests_fit_the_pool(library):
    # vgrad_fp_kernel (vgrad_fl.cuh): consumers 208 / 192 registers, the producer, loss and value warps 80 / 64; the
    # pool setmaxnreg moves registers in is threads x the launch allocation — a request beyond it blocks for ever
    # (measured the hard way: a 21-minute hang on the B200)
    out = subprocess.run(["cuobjdump", "-res-usage", library._name], capture_output=True, text=True).stdout
    found = re.findall(r"vgrad_fp_kernelILi(\d+)ELi\d+ELb[01]E\w*:\s*\n\s*REG:(\d+)", out)
    assert len(found) >= 20
    for nt8, regs in found:
        threads, consumer, other = (384, 208, 80) if nt8 == "1" else (512, 192, 64)
        assert 256 * consumer + (threads - 256) * other <= threads * int(regs), (nt8, regs)


def test_hot_kernels_keep_their_code_generation(library):
    # guards against silent code-generation regressions of the kind found in round 1 (one translation unit under
    # -split-compile: adding unrelated kernels changed mt_backward_kernel's SASS and cost config 3 15 %): the hot
    # kernels of each family must not spill beyond what was measured, and the backward kernel keeps its size
    out = subprocess.run(["cuobjdump", "-res-usage", library._name], capture_output=True, text=True).stdout
    stack = dict(re.findall(r"Function (\w+):\s*\n\s*REG:\d+ STACK:(\d+)", out))
    limits = {
        "_ZN5nb20015vgrad_t1_kernelILi2ELi16ELi1ELi4ELi2ELi1ELb1EEEvNS_9t1_paramsE": 32,   # config 2 (variant 27)
        "_ZN5nb20015vgrad_t1_kernelILi2ELi8ELi2ELi8ELi4ELi2ELb1EEEvNS_9t1_paramsE": 128,   # config 4 shards (37)
        "_ZN5nb20015vgrad_t1_kernelILi2ELi8ELi1ELi8ELi4ELi2ELb1EEEvNS_9t1_paramsE": 128,   # config 5 (33)
        "_ZN5nb20018mt_backward_kernelILi13EEEvNS_18mt_backward_paramsE": 0,               # config 3
        "_ZN5nb20017mt_forward_kernelILi13ELb1EEEvNS_17mt_forward_paramsE": 0,
        "_ZN5nb20021mt_forward_tma_kernelILi13ELb1EEEvNS_17mt_forward_paramsE14CUtensorMap_st": 16,
        "_ZN5nb20015vgrad_fd_kernelILi2ELi13ELi1ELb1EEEvNS_9fd_paramsE": 32,               # 784 x 10
        "_ZN5nb20015vgrad_fd_kernelILi1ELi32ELi2ELb1EEEvNS_9fd_paramsE": 0,                # 1024 x 8
        "_ZN5nb20015vgrad_ft_kernelILi2ELi2ELi8ELi1ELb1EEEvNS_9ft_paramsE": 0,             # 1024 x 2
        "_ZN5nb20015vgrad_fr_kernelILi2ELi4ELi1ELb1EEEvNS_9fd_paramsE": 32,                # 256 x 10 (row-split)
        "_ZN5nb20015vgrad_fr_kernelILi2ELi2ELi2ELb1EEEvNS_9fd_paramsE": 32,                # 100 x 10
        "_ZN5nb20015vgrad_fr_kernelILi1ELi4ELi1ELb1EEEvNS_9fd_paramsE": 0,                 # 256 x 2..8
        "_ZN5nb20015vgrad_fp_kernelILi2ELi13ELb1EEEvNS_9fd_paramsE": 64,                   # 784 x 10 (lane-loss pipeline)
        "_ZN5nb20015vgrad_fp_kernelILi1ELi16ELb1EEEvNS_9fd_paramsE": 0,                    # 1024 x 8
        "_ZN5nb20015vgrad_fl_kernelILi2ELi8ELi2ELb1EEEvNS_9fd_paramsE": 0,                 # 512 x 16 (two barriers)
    }
    for name, limit in limits.items():
        assert name in stack, name
        assert int(stack[name]) <= limit, (name, stack[name])
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN5nb20018mt_backward_kernelILi13EEEvNS_18mt_backward_paramsE",
                           library._name], capture_output=True, text=True).stdout
    instructions = len(re.findall(r"^\s+/\*[0-9a-f]{4,5}\*/", sass, flags=re.M))
    # 1.5 k when compiled on its own (a predicate-free loop for full stages + the masked loop for the last one);
    # the regressed one-translation-unit build of round 1 had 2.1 k for the masked loop alone
    assert 1200 <= instructions <= 1900, instructions
    # the many-target forward is fed by the tensor-map engine
    sass = subprocess.run(["cuobjdump", "-sass", "-fun",
                           "_ZN5nb20021mt_forward_tma_kernelILi13ELb1EEEvNS_17mt_forward_paramsE14CUtensorMap_st",
                           library._name], capture_output=True, text=True).stdout
    assert "UTMALDG.2D" in sass and "DMMA.8x8x4" in sass
    # the lane-loss pipeline: bulk copies into the ring, mbarrier hand-overs, the fp64 tensor pipe, register hand-over —
    # and the straight-line exponential (no call into the library routine's slow path from the softmax step)
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN5nb20015vgrad_fp_kernelILi2ELi13ELb1EEEvNS_9fd_paramsE",
                           library._name], capture_output=True, text=True).stdout
    for mnemonic in ("UBLKCP", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "DMMA.8x8x4", "USETMAXREG"):
        assert mnemonic in sass, mnemonic


def test_no_gpu_means_loud_failure_not_fallback(library):
    from libnano_b200 import Context, Nb200Error, _lib

    lib = _lib.lib()
    assert lib.nb200_version() == b"0.1.0"
    if lib.nb200_device_count() > 0:
        pytest.skip("a GPU is visible: the no-device behaviour is not observable here")
    handle = ctypes.c_void_p()
    status = lib.nb200_init(0, ctypes.byref(handle))
    assert status != 0 and not handle.value
    assert b"CUDA" in lib.nb200_last_error() or b"no CUDA device" in lib.nb200_last_error()
    with pytest.raises(Nb200Error):
        Context(0)


def test_product_never_imports_the_oracle():
    # the oracle is test infrastructure: nothing under libnano_b200/ or include/ may reference it
    offenders = []
    for base in ("libnano_b200", "include"):
        for folder, _, files in os.walk(os.path.join(ROOT, base)):
            for name in files:
                if name.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    text = open(os.path.join(folder, name)).read()
                    if re.search(r"^\s*(import|from)\s+oracle\b|liboracle|oracle\.h|oracle/", text, flags=re.M):
                        offenders.append(os.path.join(folder, name))
    assert not offenders, offenders


def test_argument_validation_without_a_device(library):
    lib = library
    lib.nb200_last_error.restype = ctypes.c_char_p
    # null handles are rejected before any CUDA call
    assert lib.nb200_objective_size(None) == 0
    assert lib.nb200_dataset_samples(None) == 0
    fx = ctypes.c_double()
    lib.nb200_vgrad.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    assert lib.nb200_vgrad(None, None, 0, None, ctypes.byref(fx)) == 1
    assert b"critical check failed" in lib.nb200_last_error()
    # host-side generator of the synthetic test functions needs no device
    X = np.empty((10, 2))
    w = np.empty(2)
    b = ctypes.c_double()
    dp = ctypes.POINTER(ctypes.c_double)
    lib.nb200_synth_linear_inputs.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int64, dp, dp,
                                              dp]
    assert lib.nb200_synth_linear_inputs(10, 2, 42, 1, X.ctypes.data_as(dp), w.ctypes.data_as(dp),
                                         ctypes.byref(b)) == 0
    assert X[0, 0] == 13930160852258120406 / 2.0 ** 64 and abs(w.sum() - 1.0) < 1e-15 and -0.5 <= b.value < 0.5
    assert lib.nb200_synth_linear_inputs(0, 2, 42, 1, None, None, None) == 1


def test_synth_inputs_match_oracle_stream(oracle, library):
    # product-side generator and the oracle's restatement of linear_model_t's constructor draw the same stream
    dp = ctypes.POINTER(ctypes.c_double)
    library.nb200_synth_linear_inputs.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int64, dp,
                                                  dp, dp]
    for dims, seed, modulo in ((5, 1, 1), (37, 42, 3)):
        X_ref, _, w_ref, b_ref = oracle.synth_model(dims, seed=seed, modulo=modulo, loss="mse")
        X = np.empty_like(X_ref)
        w = np.empty_like(w_ref)
        b = ctypes.c_double()
        assert library.nb200_synth_linear_inputs(X.shape[0], X.shape[1], seed, modulo, X.ctypes.data_as(dp),
                                                 w.ctypes.data_as(dp), ctypes.byref(b)) == 0
        assert np.array_equal(X, X_ref) and np.array_equal(w, w_ref) and b.value == b_ref


def test_loss_registry_mirrors_reference():
    from libnano_b200 import LOSS_IDS, Loss

    assert len(LOSS_IDS) == 17  # src/loss.cpp:105-135
    flags = {"mse": (True, True), "mae": (True, False), "cauchy": (False, True), "s-hinge": (True, False),
             "m-hinge": (True, False), "s-squared-hinge": (True, False), "s-classnll": (True, True),
             "s-savage": (False, True), "s-tangent": (False, True), "s-logistic": (True, True),
             "m-logistic": (True, True), "s-exponential": (True, True), "pinball": (True, False)}
    for loss_id, (convex, smooth) in flags.items():
        loss = Loss(loss_id)
        assert (loss.convex(), loss.smooth()) == (convex, smooth) and loss.type_id() == loss_id
    assert Loss("s-logistic").kind == Loss("m-logistic").kind  # the prefix only changes error()
    pinball = Loss("pinball")
    assert pinball.parameter("loss::pinball::alpha") == 0.5
    pinball.parameter("loss::pinball::alpha", 0.2)
    assert pinball.clone().param == 0.2
    with pytest.raises(RuntimeError):
        pinball.parameter("loss::pinball::alpha", 1.5)
    with pytest.raises(RuntimeError):
        Loss("no-such-loss")


def test_function_contract_on_the_host():
    """function_t::operator() bookkeeping (src/function.cpp:246-272) with a stub do_eval."""
    from libnano_b200 import Function

    class Sphere(Function):
        def __init__(self):
            super().__init__("sphere", 3)
            self._convex = self._smooth = True

        def do_eval(self, x, gx):
            if gx is not None:
                gx[:] = 2 * x
            return float(x @ x)

    function = Sphere()
    x, gx = np.array([1.0, 2.0, 3.0]), np.empty(3)
    assert function(x) == 14.0 and function(x, gx) == 14.0 and np.array_equal(gx, 2 * x)
    assert (function.fcalls(), function.gcalls(), function.hcalls()) == (2, 1, 0)
    assert function.name() == "sphere[3D]" and function.name(False) == "sphere"
    for bad in (np.zeros(2), np.zeros((3, 1))):
        with pytest.raises(RuntimeError, match="invalid input size"):
            function(bad)
    with pytest.raises(RuntimeError, match="invalid gradient size"):
        function(x, np.empty(2))
    with pytest.raises(RuntimeError):
        function(x, np.empty(3, dtype=np.float32))
    assert (function.fcalls(), function.gcalls()) == (2, 1)  # failed calls are not counted
    function.clear_statistics()
    assert function.fcalls() == 0


def test_shard_rows_partition():
    from libnano_b200.sharded import shard_rows

    for n, world in ((10, 3), (1 << 20, 8), (7, 8), (64 << 20, 8), (5, 1)):
        blocks = [shard_rows(n, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and sum(rows for _, rows in blocks) == n
        for (a0, ar), (b0, _) in zip(blocks, blocks[1:]):
            assert a0 + ar == b0
        assert max(r for _, r in blocks) - min(r for _, r in blocks) <= 1
    with pytest.raises(RuntimeError):
        shard_rows(10, 2, 2)
