/*
 * oracle.h — C interface of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a std-only C++ restatement of libnano's linear-model value-and-gradient path
 * (SURVEY.md §8). It exists to CHECK the CUDA path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. Nothing under libnano_b200/ may.
 *
 * Parity status: PINNED against the reference's own known answers for this path
 *   - pinball golden vectors              test/test_loss.cpp:271-307
 *   - sum_reduce / accumulator arithmetic test/test_linear_util.cpp:12-44
 *   - predict == W x + b                  test/test_linear_util.cpp:46-61
 *   - check_vgrad identity                test/test_linear_function.cpp:23-49
 *   - strong-convexity / size / flags     test/test_linear_function.cpp:70-150
 *   - FD-gradient < 1e-8, convexity, same-seed determinism (test/fixture/function.h:39-83,
 *     test/test_function_factory.cpp:172-253)
 * The reference itself cannot be compiled in this image (Eigen3 is an un-vendored, unpinned
 * external dependency that is not installed: CMakeLists.txt:75), so the Eigen-backed sums are
 * restated with plain loops; see DESIGN.md "Oracle".
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* loss kinds: the s-/m- prefixes of the reference only change error(), never value()/vgrad()
 * (include/nano/loss/flatten.h:103-122), so they share a kind. */
enum oracle_loss_kind
{
    ORACLE_LOSS_MSE           = 0,  /* include/nano/loss/mse.h:19-29            */
    ORACLE_LOSS_MAE           = 1,  /* include/nano/loss/mae.h:19-29            */
    ORACLE_LOSS_CAUCHY        = 2,  /* include/nano/loss/cauchy.h               */
    ORACLE_LOSS_HINGE         = 3,  /* include/nano/loss/hinge.h:19-29          */
    ORACLE_LOSS_SQUARED_HINGE = 4,  /* include/nano/loss/squared_hinge.h        */
    ORACLE_LOSS_CLASSNLL      = 5,  /* include/nano/loss/classnll.h:19-37       */
    ORACLE_LOSS_SAVAGE        = 6,  /* include/nano/loss/savage.h               */
    ORACLE_LOSS_TANGENT       = 7,  /* include/nano/loss/tangent.h              */
    ORACLE_LOSS_LOGISTIC      = 8,  /* include/nano/loss/logistic.h:19-41       */
    ORACLE_LOSS_EXPONENTIAL   = 9,  /* include/nano/loss/exponential.h          */
    ORACLE_LOSS_PINBALL       = 10, /* src/loss/pinball.cpp:24-48               */
    ORACLE_LOSS_COUNT         = 11
};

/* objective (B) losses: src/function/mlearn/loss.cpp (already divided by N) */
enum oracle_synth_loss
{
    ORACLE_SYNTH_MSE      = 0,
    ORACLE_SYNTH_MAE      = 1,
    ORACLE_SYNTH_CAUCHY   = 2,
    ORACLE_SYNTH_HINGE    = 3,
    ORACLE_SYNTH_LOGISTIC = 4
};

/* loss_t::value / loss_t::vgrad over (N, T) row-major targets/outputs (src/loss.cpp:57-71) */
void oracle_loss_value(int kind, double param, const double* targets, const double* outputs, int64_t N, int64_t T,
                       double* values);
void oracle_loss_vgrad(int kind, double param, const double* targets, const double* outputs, int64_t N, int64_t T,
                       double* vgrads);

/* loss_t::error over (N, T) (include/nano/loss/flatten.h:62-67 -> include/nano/loss/error.h:11-70).
 * error_kind: 0 absdiff (regression ids: L1 distance), 1 sclass (s- ids: 0-1 error of the arg-max label, or of the sign
 * when T == 1), 2 mclass (m- ids: number of labels with target * output < epsilon). pinball's error is its value
 * (src/loss/pinball.cpp:20-23). */
void oracle_loss_error(int error_kind, const double* targets, const double* outputs, int64_t N, int64_t T,
                       double* errors);

/* linear::predict (src/linear/util.cpp:6-21): outputs[N,T] = X[N,D] * W[T,D]^T + b[T] */
void oracle_predict(const double* X, const double* W, const double* b, int64_t N, int64_t D, int64_t T,
                    double* outputs);

/* sum_reduce (include/nano/core/reduce.h:23-31) over `count` accumulators of `len` doubles each,
 * laid out contiguously; result (sum / samples) lands in the first accumulator. */
void oracle_sum_reduce(double* accumulators, int64_t count, int64_t len, int64_t samples);

/* objective (A) linear::function_t::do_eval (src/linear/function.cpp:44-168).
 * x = [W (T x D row-major) | b (T)], gx nullable (value only). `batch` rows per task
 * (linear::batch, default 100), `threads` per-thread accumulators (batches are dealt round-robin so
 * the result is deterministic for a fixed (batch, threads)). */
double oracle_linear_vgrad(const double* X, const double* Y, int64_t N, int64_t D, int64_t T, int kind, double param,
                           double l1, double l2, const double* x, double* gx, int64_t batch, int threads);

/* the two halves of oracle_linear_vgrad, split where a sample-sharded evaluation all-reduces: `partial` receives the
 * UN-normalised [sum loss | sum dL (T) | sum dL x^T (T*D)] of the given rows; `finish` divides the (summed) vector by
 * the global sample count and adds the regularisation terms (src/linear/function.cpp:124-167). */
void oracle_linear_partial(const double* X, const double* Y, int64_t N, int64_t D, int64_t T, int kind, double param,
                           const double* x, int want_grad, int64_t batch, int threads, double* partial);
double oracle_linear_finish(const double* reduced, int64_t n_global, int64_t D, int64_t T, double l1, double l2,
                            const double* x, double* gx);

/* the same objective evaluated with long-double accumulation everywhere ("truth" for error bounds) */
double oracle_linear_vgrad_ld(const double* X, const double* Y, int64_t N, int64_t D, int64_t T, int kind,
                              double param, double l1, double l2, const double* x, double* gx);

/* fast CPU baseline of objective (A): same arithmetic as oracle_linear_vgrad but the row dot products are
 * SIMD-reassociated (what Eigen's packet GEMV does); used by bench.py as the host-CPU arm only. */
double oracle_linear_vgrad_fast(const double* X, const double* Y, int64_t N, int64_t D, int64_t T, int kind,
                                double param, double l1, double l2, const double* x, double* gx, int64_t batch,
                                int threads);

/* objective (B) data: linear_model_t ctor (src/function/mlearn/linear.cpp:5-45).
 * X[N,D], Y[N], wopt[D], bopt[1]; T = 1 (src/function/mlearn/ridge.cpp:18-21). */
void oracle_synth_model(int64_t samples, int64_t inputs, uint64_t seed, int64_t modulo, int regression, double* X,
                        double* Y, double* wopt, double* bopt);

/* objective (B) evaluation: linear_model_t::eval<tloss> (src/function/mlearn/linear.h:21-38) followed by the
 * regulariser of ridge.cpp:63-80 / lasso.cpp:63-76 / elasticnet.cpp:67-86 (alpha1 = 0 and/or alpha2 = 0 select). */
double oracle_synth_vgrad(const double* X, const double* Y, double bopt, int64_t N, int64_t D, int synth_loss,
                          double alpha1, double alpha2, const double* x, double* gx);

/* flatten_iterator_t's view of the data: scalar_stats_t update/done (src/dataset/stats.cpp:81-137),
 * scalar_stats_t::scale + nan2zero (:357-401, :8-22) and nano::upscale (:211-230).
 * stats: [7][C] = samples, min, max, mean, stdev, div_range, div_stdev; scaling: 0 none, 1 mean, 2 minmax, 3 standard */
void oracle_scalar_stats(const double* values, int64_t N, int64_t C, const unsigned char* enable, double* stats);
void oracle_scale(double* values, int64_t N, int64_t C, const double* stats, int scaling);
void oracle_upscale(const double* istats, int iscaling, const double* tstats, int tscaling, int64_t D, int64_t T,
                    double* W, double* b);

/* make_samples (src/function/mlearn/ridge.cpp:23-26) */
int64_t oracle_synth_samples(int64_t dims, double sratio);

/* gboost::bias_function_t / scale_function_t / grads_function_t::do_eval (src/gboost/function.cpp:26-50, 85-126, 151-223)
 * over targets[N,T] (+ cluster[N], soutputs / woutputs[N,T] for the scale function); kind: 0 bias (n = T), 1 scale
 * (n = groups), 2 grads (n = N * T); gx nullable; `batch` rows per targets_iterator_t::loop range. */
double oracle_gboost_vgrad(int kind, int loss, double param, const double* targets, int64_t N, int64_t T,
                           const int32_t* cluster, int64_t groups, const double* soutputs, const double* woutputs,
                           const double* x, double* gx, int64_t batch);

int oracle_hardware_threads(void);

#ifdef __cplusplus
}
#endif
