/*
 * nb200.h — C ABI of libnb200.so: the B200-native (sm_100a) value-and-gradient evaluation of libnano's
 * linear-model objective. Plain pointers and sizes only; no exceptions cross this boundary; every entry point
 * returns 0 on success and a non-zero nb200_status otherwise (message via nb200_last_error()).
 *
 * What it replaces in the reference (/root/reference, all CPU/Eigen):
 *   nano::function_t::operator()(x, gx)           src/function.cpp:246-272   -> nb200_vgrad
 *   nano::linear::function_t (ctor, do_eval)      src/linear/function.cpp:21-168 -> nb200_objective_create / nb200_vgrad
 *   nano::linear::predict                         src/linear/util.cpp:6-21   -> nb200_predict
 *   nano::loss_t::value / vgrad                   src/loss.cpp:57-71         -> nb200_loss_value / nb200_loss_vgrad
 *   accumulator_t + sum_reduce                    src/linear/accumulator.cpp, include/nano/core/reduce.h:23-31
 *                                                 -> nb200_vgrad_partial + (all-reduce) + nb200_vgrad_finish
 *   flatten_iterator_t::cache_flatten/targets     src/dataset/iterator.cpp:118-147,52-77 -> nb200_dataset_pin
 *   linear_model_t + function_{ridge,lasso,elasticnet}_t   src/function/mlearn/   -> flavour NB200_FLAVOUR_SYNTH
 *
 * Threading contract (same as the reference's `const do_eval` + `mutable` scratch): one call in
 * flight per nb200_obj; distinct objects (clones) sharing one nb200_data may be used concurrently.
 */
#ifndef NB200_H
#define NB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NB200_API __attribute__((visibility("default")))
#else
#define NB200_API
#endif

typedef struct nb200_ctx  nb200_ctx;  /* one CUDA device + stream                                  */
typedef struct nb200_data nb200_data; /* a feature matrix X[N,D] and targets Y[N,T] pinned in HBM  */
typedef struct nb200_obj  nb200_obj;  /* an objective over a pinned dataset (loss, reg, scratch)   */

typedef enum nb200_status
{
    NB200_OK               = 0,
    NB200_ERR_INVALID      = 1, /* bad argument (size mismatch, unknown loss, null pointer)          */
    NB200_ERR_CUDA         = 2, /* CUDA runtime error (no device, out of memory, launch failure)     */
    NB200_ERR_UNSUPPORTED  = 3, /* well-formed request outside the implemented path (e.g. Hessian)   */
} nb200_status;

/* loss kinds. libnano ids (src/loss.cpp:105-135) map onto them with the s-/m- prefix dropped, because the prefix
 * only changes error(), never value()/vgrad() (include/nano/loss/flatten.h:103-122). */
typedef enum nb200_loss
{
    NB200_LOSS_MSE            = 0,  /* include/nano/loss/mse.h:19-29                                 */
    NB200_LOSS_MAE            = 1,  /* include/nano/loss/mae.h:19-29                                 */
    NB200_LOSS_CAUCHY         = 2,  /* include/nano/loss/cauchy.h                                    */
    NB200_LOSS_HINGE          = 3,  /* include/nano/loss/hinge.h:19-29                               */
    NB200_LOSS_SQUARED_HINGE  = 4,  /* include/nano/loss/squared_hinge.h                             */
    NB200_LOSS_CLASSNLL       = 5,  /* include/nano/loss/classnll.h:19-37                            */
    NB200_LOSS_SAVAGE         = 6,  /* include/nano/loss/savage.h                                    */
    NB200_LOSS_TANGENT        = 7,  /* include/nano/loss/tangent.h                                   */
    NB200_LOSS_LOGISTIC       = 8,  /* include/nano/loss/logistic.h:19-41                            */
    NB200_LOSS_EXPONENTIAL    = 9,  /* include/nano/loss/exponential.h                               */
    NB200_LOSS_PINBALL        = 10, /* src/loss/pinball.cpp:24-48 (loss_param = loss::pinball::alpha) */
    NB200_LOSS_SYNTH_CAUCHY   = 11, /* src/function/mlearn/loss.cpp:50-63 (no 1/2, gradient x2)      */
    NB200_LOSS_SYNTH_LOGISTIC = 12, /* src/function/mlearn/loss.cpp:97-125 (un-stabilised gradient)  */
    NB200_LOSS_COUNT          = 13
} nb200_loss;

/* error measures behind loss_t::error (include/nano/loss/error.h:11-70): which one a libnano loss id uses is its
 * prefix (include/nano/loss/flatten.h:103-122): none -> absdiff, "s-" -> sclass, "m-" -> mclass. */
typedef enum nb200_error
{
    NB200_ERROR_ABSDIFF = 0, /* regression: L1 distance between target and output                            */
    NB200_ERROR_SCLASS  = 1, /* single-label: 0-1 error of the arg-max label (of the sign when T == 1)        */
    NB200_ERROR_MCLASS  = 2, /* multi-label: number of labels with target * output < epsilon                  */
    NB200_ERROR_COUNT   = 3
} nb200_error;

typedef enum nb200_flavour
{
    /* nano::linear::function_t (src/linear/function.cpp): x = [W (T x D row-major) | b (T)], n = (D+1)*T,
     * f = mean loss + l1*mean|W| + 0.5*l2*mean(W^2), gW += l1*sign(W)/(T*D) + l2*W/(T*D). */
    NB200_FLAVOUR_LINEAR = 0,
    /* function_{ridge,lasso,elasticnet}_t over linear_model_t (src/function/mlearn/): T = 1, x = w (n = D), the
     * bias is FIXED (fixed_bias), f = mean loss + l1*|x|_1 + 0.5*l2*|x|^2, g += l1*sign(x) + l2*x (un-normalised;
     * ridge.cpp:69-79, lasso.cpp:69-75, elasticnet.cpp:74-85). */
    NB200_FLAVOUR_SYNTH = 1
} nb200_flavour;

typedef enum nb200_synth_kind
{
    /* X_ij ~ U[-1,1) from Philox4x32-10 (key = seed, counter = global element index),
     * w*_j ~ U[-1,1)/sqrt(D), b* = 0.1, y_i = sign(x_i . w* + b* + noise * n_i) in {-1,+1}, T = 1 */
    NB200_SYNTH_CLASSIFICATION = 0,
    /* same X, w*, b*; y_i = x_i . w* + b* + noise * n_i, T = 1 */
    NB200_SYNTH_REGRESSION = 1,
    /* same X; W*[T,D] ~ U[-1,1)/sqrt(D); targets are -1 everywhere except +1 at argmax_t(x_i . W*_t)
     * (include/nano/loss/class.h:51-58), T >= 2 */
    NB200_SYNTH_MULTICLASS = 2
} nb200_synth_kind;

/* ---- library / device ---------------------------------------------------------------------------------------- */

/* Thread-local message describing the last failure on the calling thread ("" if none). */
NB200_API const char* nb200_last_error(void);

/* Library version "major.minor.patch". */
NB200_API const char* nb200_version(void);

/* Number of visible CUDA devices (0 if none / no driver); never fails. */
NB200_API int nb200_device_count(void);

/* Bind a context to CUDA device `device` (creates a non-blocking stream). Fails loudly (NB200_ERR_CUDA) when
 * there is no usable GPU: there is no CPU fallback. */
NB200_API int nb200_init(int device, nb200_ctx** out_ctx);

/* A device GROUP: one process, several GPUs. libnano is one process that shards do_eval over the
 * worker threads of its pool (src/linear/function.cpp:53-57) and fits one model per process (src/linear.cpp:30-48);
 * a group context gives that process every GPU of the box. Datasets pinned (or generated) on it are split in
 * contiguous balanced row blocks over `devices` (the first N % ndev devices hold one row more), objectives created
 * on them evaluate on every device from the calling host thread, and every entry point that takes an nb200_obj /
 * nb200_data / nb200_ctx accepts the group handle (exceptions: the cross-process protocol nb200_vgrad_partial /
 * finish / peer_* / sharded, nb200_objective_set_stream and nb200_vgrad_batch return NB200_ERR_UNSUPPORTED).
 * Single-target objectives on distinct devices with a peer path publish their reduced sums straight into the first
 * device's mailbox (NVLink) from inside the launch; everything else gathers the per-device sums by peer copies. The
 * fold runs in device order: results are bit-identical to the one-process-per-GPU path on the same partition.
 * x / gx / fx of nb200_vgrad_device live on devices[0]. */
NB200_API int nb200_init_multi(int ndev, const int* devices, nb200_ctx** out_ctx);
/* devices of a context: 1 for nb200_init, ndev for nb200_init_multi */
NB200_API int nb200_ctx_devices(const nb200_ctx* ctx);
NB200_API void nb200_ctx_release(nb200_ctx* ctx);

/* ---- dataset: replaces flatten_iterator_t::cache_flatten/cache_targets ---------------------------------------- */

/* Copy X[N,D] and Y[N,T] (host, row-major, contiguous fp64) into HBM once per fit. */
NB200_API int nb200_dataset_pin(nb200_ctx* ctx, const double* X, const double* Y, int64_t N, int64_t D, int64_t T,
                                nb200_data** out_data);

/* Generate rows [row_offset, row_offset + N) of a synthetic dataset directly in HBM (no host staging; the big
 * BASELINE configs do not fit host RAM). The same (seed, kind, D, T, noise) with different row_offset gives
 * disjoint shards of ONE dataset, which is how ranks shard it. */
NB200_API int nb200_dataset_synth(nb200_ctx* ctx, uint64_t seed, int synth_kind, int64_t N, int64_t D, int64_t T,
                                  int64_t row_offset, double noise, nb200_data** out_data);

/* Read rows [row0, row0 + rows) back to the host (X and/or Y may be NULL). */
NB200_API int nb200_dataset_read(const nb200_data* data, int64_t row0, int64_t rows, double* X, double* Y);

NB200_API int64_t nb200_dataset_samples(const nb200_data* data);
NB200_API int64_t nb200_dataset_inputs(const nb200_data* data);
NB200_API int64_t nb200_dataset_targets(const nb200_data* data);

/* flatten_iterator_t::scaling on the device (src/dataset/stats.cpp:81-137,357-401): column statistics over the finite
 * values of the pinned inputs and targets, then in-place scaling (0 none, 1 mean, 2 minmax, 3 standard — the values of
 * nano::scaling_type, include/nano/dataset/scaling.h:10-16) with non-finite values (missing data) set to 0.
 * enable_*: optional byte masks, 0 = leave that column unscaled (categorical columns / class targets). Call once, before
 * creating objectives. nb200_dataset_stats returns the scalar_stats_t arrays (any pointer may be NULL; which: 0 inputs,
 * 1 targets); nb200_upscale is nano::upscale (stats.cpp:211-230): it maps (W [T, D], b [T]) fitted on the scaled data
 * back to the original units, in place. */
NB200_API int nb200_dataset_scale(nb200_data* data, int inputs_scaling, int targets_scaling,
                                  const unsigned char* enable_inputs, const unsigned char* enable_targets);
NB200_API int nb200_dataset_stats(const nb200_data* data, int which, double* samples, double* min, double* max,
                                  double* mean, double* stdev, double* div_range, double* div_stdev);
NB200_API int nb200_upscale(const nb200_data* data, double* W, double* b);

/* Datasets are reference counted: objectives retain the dataset they were created on. */
NB200_API void nb200_dataset_retain(nb200_data* data);
NB200_API void nb200_dataset_release(nb200_data* data);

/* ---- objective: replaces nano::linear::function_t / function_{ridge,lasso,elasticnet}_t ---------------------- */

/* fixed_bias: T doubles (host), used by NB200_FLAVOUR_SYNTH only (may be NULL otherwise). */
NB200_API int nb200_objective_create(nb200_data* data, int loss, double loss_param, int flavour, double l1,
                                     double l2, const double* fixed_bias, nb200_obj** out_obj);

/* clone(): shares the pinned dataset (retained), duplicates stream and scratch (src/linear/function.cpp:39-42). */
NB200_API int nb200_objective_clone(const nb200_obj* obj, nb200_obj** out_obj);
NB200_API void nb200_objective_release(nb200_obj* obj);

/* n: (D+1)*T for NB200_FLAVOUR_LINEAR, D for NB200_FLAVOUR_SYNTH. */
NB200_API int64_t nb200_objective_size(const nb200_obj* obj);

/* f(x) and, when gx != NULL, the gradient. x, gx, fx are HOST pointers (what libnano's solvers pass); the call
 * copies x to the device, evaluates, copies (gx, fx) back and returns when they are valid. `n` must equal
 * nb200_objective_size() (the reference throws on a size mismatch: src/function.cpp:248-252). Non-finite results
 * are returned as they are (solver_state_t::valid() decides, src/solver/state.cpp:170-174). */
NB200_API int nb200_vgrad(nb200_obj* obj, const double* x, int64_t n, double* gx, double* fx);

/* Same evaluation with x / gx / fx resident in HBM; enqueued on the objective's stream, no host sync
 * (call nb200_objective_sync before reading on the host or from another stream). */
NB200_API int nb200_vgrad_device(nb200_obj* obj, const double* x_dev, int64_t n, double* gx_dev, double* fx_dev);
NB200_API int nb200_objective_sync(nb200_obj* obj);

/* Sample-sharded evaluation (one rank per GPU): phase 1 leaves the UN-normalised local sums
 *   partial = [ sum_i loss_i | sum_i dL_i (T) | sum_i dL_i x_i^T (T*D) ]      (1 + T + T*D doubles, device)
 * in a device buffer owned by the objective (gradient entries are left untouched when want_grad == 0). The
 * caller all-reduces (sum) that buffer across ranks on `nb200_objective_stream()` or after
 * nb200_objective_sync(), then phase 2 divides by the GLOBAL sample count, adds the regularisation terms and
 * writes (gx, fx) to the host. With one rank the pair is equivalent to nb200_vgrad. */
NB200_API int nb200_vgrad_partial(nb200_obj* obj, const double* x, int64_t n, int want_grad, double** out_partial_dev,
                                  int64_t* out_partial_len);
NB200_API int nb200_vgrad_finish(nb200_obj* obj, int64_t n_global, double* gx, double* fx);

/* The same two phases with x / gx / fx resident in HBM (device-resident solvers; no host copy, no sync). */
NB200_API int nb200_vgrad_partial_device(nb200_obj* obj, const double* x_dev, int64_t n, int want_grad,
                                         double** out_partial_dev, int64_t* out_partial_len);
NB200_API int nb200_vgrad_finish_device(nb200_obj* obj, const double* x_dev, int64_t n_global, double* gx_dev,
                                        double* fx_dev);

/* Fused compute + collective for the single-target kernel: instead of all-reducing between two phases, the ONE launch
 * of every rank writes its locally reduced column slices into the peers' mailboxes over NVLink (peer memory mapped
 * through CUDA IPC), waits for the matching slices and folds them in rank order — identical bits on every rank, no
 * NCCL call, no second kernel. Setup once per objective: every rank exports a 64-byte handle (and its grid size,
 * which must agree across ranks), the handles are all-gathered by the host (world x 64 bytes, rank order), then
 * connected. Evaluate with nb200_vgrad_sharded(_device); all ranks must call it the same number of times. */
NB200_API int nb200_objective_peer_export(nb200_obj* obj, int world, void* handle_out, int64_t* grid_out);
NB200_API int nb200_objective_peer_connect(nb200_obj* obj, int rank, int world, const void* handles);
NB200_API int nb200_vgrad_sharded(nb200_obj* obj, const double* x, int64_t n, int64_t n_global, double* gx, double* fx);
NB200_API int nb200_vgrad_sharded_device(nb200_obj* obj, const double* x_dev, int64_t n, int64_t n_global,
                                         double* gx_dev, double* fx_dev);

/* How a connected objective meets its peers: 0 = the launch publishes its sums into the peers' mailboxes and leaves,
 * the STREAM waits for the peers' ready words (cuStreamWaitValue64: no SM is held, so evaluations of different
 * objectives can be enqueued in any order on different ranks without cross-waiting) and a finish kernel folds the
 * slots; 1 = the launch itself waits for the peers (one kernel; used when the driver offers no stream wait, or with
 * NB200_PEER_WAIT=1); -1 = not connected. Both kinds raise both signals, so ranks in different modes interoperate. */
NB200_API int nb200_objective_peer_mode(const nb200_obj* obj);

/* The per-device objectives behind a group objective (1 and the objective itself for a plain one): borrowed handles
 * for per-device introspection (nb200_objective_describe / pass_stats / last_pass_ms / launches). */
NB200_API int        nb200_objective_parts(const nb200_obj* obj);
NB200_API nb200_obj* nb200_objective_part(nb200_obj* obj, int index);

/* cudaStream_t of the objective, as an opaque pointer (so that callers can order collectives after phase 1);
 * set_stream re-points the objective at a caller-owned stream (a NULL handle is the legacy default stream), e.g. the
 * stream the caller's NCCL all-reduce is enqueued on, so that phase 1 -> all-reduce -> phase 2 need no host
 * synchronisation; use_own != 0 goes back to the objective's own stream. */
NB200_API void* nb200_objective_stream(const nb200_obj* obj);
NB200_API int   nb200_objective_set_stream(nb200_obj* obj, void* stream, int use_own);

/* Kernels launched by this objective since creation (bench.py's `gpu_launches`). */
NB200_API int64_t nb200_objective_launches(const nb200_obj* obj);

/* Device timing of the main pass (the kernel(s) that stream X), from CUDA events recorded on the objective's
 * stream around those launches. Off by default. last_pass_ms: most recent evaluation (< 0 if unavailable; syncs
 * the stream); pass_stats: sum and count over the evaluations since the last reset. */
NB200_API int    nb200_objective_enable_timing(nb200_obj* obj, int enabled);
NB200_API double nb200_objective_last_pass_ms(nb200_obj* obj);
NB200_API int    nb200_objective_pass_stats(nb200_obj* obj, double* total_ms, int64_t* count, int reset);

/* The passes since the last reset one by one (milliseconds, at most `capacity` of them; *count = how many there are):
 * the distribution behind pass_stats, e.g. to tell a straggling device from a slow collective. A group objective
 * reports its first device; the others through nb200_objective_part. */
NB200_API int nb200_objective_pass_samples(nb200_obj* obj, double* pass_ms, int64_t capacity, int64_t* count);

/* The roofline denominators of THIS board, measured in the run that quotes them (a few milliseconds each):
 * the issue rate of the fp64 tensor pipe (mma.sync m8n8k4 f64 — tcgen05 has no fp64 kind), of the fp64 FMA pipe, and
 * a read-only HBM stream through 1-D bulk copies (the single-target pass only reads; the usual copy figure counts
 * read + write). */
#define NB200_PEAK_DMMA_TFLOPS 0
#define NB200_PEAK_DFMA_TFLOPS 1
#define NB200_PEAK_HBM_READ_GBS 2
NB200_API int nb200_probe_peak(nb200_ctx* ctx, int kind, double* value);

/* Host-side wall clock of nb200_vgrad since the last reset, microseconds summed over `calls` calls: staging x
 * (+ enqueueing its H2D copy), enqueueing the evaluation, waiting for the stream and handing (g, f) over. */
NB200_API int nb200_objective_host_profile(nb200_obj* obj, double* stage_us, double* enqueue_us, double* wait_us,
                                           int64_t* calls, int reset);

/* Tuning knob: variant id of the fused T == 1 kernel (see DESIGN.md "kernel variants"), -1 = heuristic,
 * -2 = force the shape-generic kernels, -3 = shape-generic kernels for T >= 2 only, -4 = the two-pass tensor-pipe
 * pair for T >= 2 even where a one-pass kernel applies, -5 / -6 = the one-pass tensor-pipe / CUDA-core kernel. */
NB200_API int nb200_objective_set_variant(nb200_obj* obj, int variant);

/* Human-readable launch plan (kernel variant, grid, stages, shared memory) into buffer[size]. */
NB200_API int nb200_objective_describe(const nb200_obj* obj, char* buffer, int64_t size);

/* Many models in ONE pass over X (the tuner's grid of (l1, l2) values and cross-validation trials
 * re-read the same X once per trial in the reference, src/machine/tune.cpp:25-42). K single-target models that share
 * the objective's dataset and loss and differ in x_k, l1_k, l2_k: x [K][n], l1 [K], l2 [K] -> gx [K][n] (nullable),
 * fx [K], all host. Runs as one many-"target" problem on the FP64 tensor pipe: ONE read of X per 8 models for K <= 24
 * (one-pass kernel, D >= 200), two reads of X for larger batches; shapes outside those kernels (odd D, D < 200 with
 * fewer than 6 models) fall back to K passes of the single-model kernel. */
NB200_API int nb200_vgrad_batch(nb200_obj* obj, int64_t K, const double* x, int64_t n, const double* l1,
                                const double* l2, double* gx, double* fx);

/* ---- device-resident solver (BASELINE north_star item 3) --------------------------------------------------------- */

/* Sum-all-reduce `length` doubles at `device_buffer` across the ranks of a sample-sharded job, enqueued on
 * `cuda_stream` (e.g. ncclAllReduce, or torch.distributed.all_reduce on a tensor aliasing the buffer). 0 = success. */
typedef int (*nb200_allreduce_fn)(void* user, double* device_buffer, int64_t length, void* cuda_stream);

/* nano::solver_lbfgs_t::minimize (src/solver/lbfgs.cpp:25-124; lsearch0 = quadratic, lsearchk = cgdescent, tolerance
 * (1e-4, 0.9)) with x, g, the (s, y) history, the two-loop recursion and the line-search trial points resident in
 * HBM: per evaluation only four scalars reach the host. x: in = x0, out = the returned state's x (host, n doubles).
 * n_global = 0 and allreduce = NULL for a single rank; otherwise every rank calls this with the same x0 and the
 * callback all-reduces the un-normalised sums between the two phases. stats = {fcalls, gcalls, iterations};
 * status: 0 max_evals reached, 1 converged (gradient test < epsilon), 4 failed (non-finite state / line search). */
NB200_API int nb200_lbfgs_minimize(nb200_obj* obj, double* x, int64_t n, double epsilon, int64_t max_evals,
                                   int64_t history, int64_t n_global, nb200_allreduce_fn allreduce,
                                   void* allreduce_user, double* fx, double* gradient_test, int64_t* stats,
                                   int* status);

/* nano::solver_cgd_t::minimize (src/solver/cgd.cpp:86-143; ids cgd-hs, cgd-fr, cgd-pr (libnano's "cgd"), cgd-cd,
 * cgd-ls, cgd-dy, cgd-n, cgd-dycd, cgd-dyhs, cgd-frpr = beta 0..9; tolerance (1e-4, 1e-1), lsearch0 = quadratic,
 * lsearchk = cgdescent) on the same device-resident skeleton: x, g, the previous gradient and the direction stay in
 * HBM, the beta formula, the update d = -g + beta d and the restart test (cgd.cpp:121-126) are one kernel.
 * orthotest = solver::cgd::orthotest (default 0.1), eta = solver::cgdN::eta (default 0.01; beta 6 only).
 * Other arguments, stats and status as nb200_lbfgs_minimize. */
#define NB200_CGD_HS 0
#define NB200_CGD_FR 1
#define NB200_CGD_PR 2
#define NB200_CGD_CD 3
#define NB200_CGD_LS 4
#define NB200_CGD_DY 5
#define NB200_CGD_N 6
#define NB200_CGD_DYCD 7
#define NB200_CGD_DYHS 8
#define NB200_CGD_FRPR 9
NB200_API int nb200_cgd_minimize(nb200_obj* obj, double* x, int64_t n, int beta, double epsilon, int64_t max_evals,
                                 double orthotest, double eta, int64_t n_global, nb200_allreduce_fn allreduce,
                                 void* allreduce_user, double* fx, double* gradient_test, int64_t* stats, int* status);

/* nano::solver_osga_t::do_minimize (src/solver/osga.cpp:59-147; the non-smooth driver: hinge / mae losses, lasso and
 * elastic-net regularisation) with z0, xb, x, x', g, h and h_hat resident in HBM: per iteration one f(x, g), one
 * value-only f(x'), two small kernels for the O(n) algebra (u = U(gamma, h) is never materialised) and a handful of
 * scalars through mapped memory; the parameter update (alpha, eta, gamma) and the stopping tests (eta < epsilon; the
 * value / patience test of src/solver/track.cpp:24-59) run on the host. lambda, alpha_max, kappa_prime, kappa: the
 * solver::osga::* parameters (defaults 0.9, 0.7, 0.1, 1.1); strong_convexity: function_t::strong_convexity().
 * x: in = x0, out = the best point. Sharded: connected peers (n_global = all samples) or a device group.
 * stats = {fcalls, gcalls, iterations}; status: 0 max_evals, 1 gradient test, 2 eta < epsilon, 3 value test, 4 failed. */
NB200_API int nb200_osga_minimize(nb200_obj* obj, double* x, int64_t n, double epsilon, int64_t max_evals,
                                  int64_t patience, double lambda, double alpha_max, double kappa_prime, double kappa,
                                  double strong_convexity, int64_t n_global, double* fx, double* eta, int64_t* stats,
                                  int* status);

/* ---- gradient-boosting cousins of the linear objective ---------------------------------------------
 * The same per-sample loss value / vgrad and accumulate-then-sum_reduce structure, without a feature matrix:
 *   NB200_GBOOST_BIAS   gboost::bias_function_t  (src/gboost/function.cpp:67-126):  f(x) = mean_i loss(y_i, x), n = T
 *   NB200_GBOOST_SCALE  gboost::scale_function_t (src/gboost/function.cpp:128-223): f(x) = mean_i loss(y_i, s_i + x[c_i] w_i),
 *                       n = groups; cluster[i] = group of sample i or -1 (scale 0, no gradient contribution)
 *   NB200_GBOOST_GRADS  gboost::grads_function_t (src/gboost/function.cpp:8-64):    f(o) = mean_i loss(y_i, o_i), n = N*T,
 *                       gradient = the per-sample loss gradients / N
 * targets / soutputs / woutputs: [N, T] host, row-major, gathered for the samples of the iterator (what
 * targets_iterator_t::loop serves); cluster / groups / soutputs / woutputs are used by NB200_GBOOST_SCALE only.
 * nb200_gboost_vgrad has the contract of nb200_vgrad (host x / gx / fx, gx nullable, size checked). */
typedef struct nb200_gboost nb200_gboost;
#define NB200_GBOOST_BIAS 0
#define NB200_GBOOST_SCALE 1
#define NB200_GBOOST_GRADS 2
NB200_API int nb200_gboost_create(nb200_ctx* ctx, int kind, int loss, double loss_param, const double* targets, int64_t N,
                                  int64_t T, const int32_t* cluster, int64_t groups, const double* soutputs,
                                  const double* woutputs, nb200_gboost** out_obj);
NB200_API int64_t nb200_gboost_size(const nb200_gboost* obj);
NB200_API int64_t nb200_gboost_launches(const nb200_gboost* obj);
NB200_API int     nb200_gboost_vgrad(nb200_gboost* obj, const double* x, int64_t n, double* gx, double* fx);
NB200_API void    nb200_gboost_release(nb200_gboost* obj);

/* ---- building blocks exposed on their own (reference: linear::predict, loss_t::value / loss_t::vgrad) -------- */

/* outputs[N,T] (host) = X * W^T + b over the pinned dataset; W[T,D], b[T] host. */
NB200_API int nb200_predict(nb200_ctx* ctx, const nb200_data* data, const double* W, const double* b,
                            double* outputs);

/* per-sample loss values[N] / gradients vgrads[N,T] for host targets/outputs[N,T]. */
NB200_API int nb200_loss_value(nb200_ctx* ctx, int loss, double loss_param, const double* targets,
                               const double* outputs, int64_t N, int64_t T, double* values);
NB200_API int nb200_loss_vgrad(nb200_ctx* ctx, int loss, double loss_param, const double* targets,
                               const double* outputs, int64_t N, int64_t T, double* vgrads);

/* per-sample errors[N] (loss_t::error, src/loss.cpp:50-55) for host targets/outputs[N,T]; error_kind: nb200_error. */
NB200_API int nb200_loss_error(nb200_ctx* ctx, int error_kind, const double* targets, const double* outputs, int64_t N,
                               int64_t T, double* errors);

/* linear::evaluate (src/linear/util.cpp:30-49) over the pinned dataset: outputs = X W^T + b stay in HBM, the
 * per-sample errors[N] and loss values[N] (either may be NULL) come back to the host. W[T,D], b[T] host. The
 * pinball loss reports its value as its error (src/loss/pinball.cpp:20-23) whatever error_kind says. */
NB200_API int nb200_evaluate(nb200_ctx* ctx, const nb200_data* data, int loss, double loss_param, int error_kind,
                             const double* W, const double* b, double* errors, double* values);

/* Host-side algebra of the proximal bundle drivers (no device needed): alpha[m] = argmin a.K.a / 2 - h.a over the unit
 * simplex, K[m,m] symmetric positive semi-definite. This is the dual of the quadratic program libnano's bundle solvers
 * build in n + 1 variables (src/solver/bundle/bundle.cpp:109-168): one variable per cutting plane. Inactive planes get
 * an exact zero. NB200_ERR_INVALID if K is not positive semi-definite (or an argument is null / non-finite). */
NB200_API int nb200_simplex_qp(const double* K, const double* h, int64_t m, double* alpha);

/* Host-side data of the builtin synthetic test functions: the random part of the linear_model_t constructor
 * (src/function/mlearn/linear.cpp:15-33): rng = std::mt19937_64(seed), u ~ uniform_real_distribution(0, 1);
 * X[samples, inputs] row-major, then wopt[inputs], then bopt = u - 0.5; wopt is divided by its sum and the columns
 * with j % modulo != 0 are zeroed. (The targets follow from nb200_predict.) */
NB200_API int nb200_synth_linear_inputs(int64_t samples, int64_t inputs, uint64_t seed, int64_t modulo, double* X,
                                        double* wopt, double* bopt);

#ifdef __cplusplus
}
#endif

#endif /* NB200_H */
