// oracle.cpp — CPU ORACLE for libnano's linear-model value+gradient path.
//
// TEST INFRASTRUCTURE ONLY. Nothing under libnano_b200/ (the product) may include, link or load this file;
// it is the checker for the CUDA path (tests/, __graft_entry__.smoke()) and the host-CPU arm of bench.py.
//
// It restates — with plain loops, no Eigen — the arithmetic of the reference files cited at each function.
// The reference's GEMV/GEMM/sum kernels live in Eigen3 (unpinned, not installed here: CMakeLists.txt:75), whose
// packetised reduction order is unspecified, so parity with the reference is tolerance-based (1e-12 relative),
// see oracle.h for the list of reference known-answers this oracle is pinned against.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <random>
#include <thread>
#include <vector>

namespace
{
// Eigen's array sign(): (a > 0) - (a < 0), so sign(0) == 0.
template <class S>
inline S sign_of(const S a)
{
    return static_cast<S>((a > S(0)) - (a < S(0)));
}

// Eigen's array max(scalar): plain compare.
template <class S>
inline S max_of(const S a, const S b)
{
    return (a < b) ? b : a;
}

template <class S>
inline S square(const S a)
{
    return a * a;
}

// ---------------------------------------------------------------------------------------------------------------
// per-sample loss value over the T outputs of one sample
// ---------------------------------------------------------------------------------------------------------------
template <class S>
S loss_value(const int kind, const S param, const double* t, const S* o, const int64_t T)
{
    S value = S(0);
    switch (kind)
    {
    case ORACLE_LOSS_MSE: // include/nano/loss/mse.h:19-22 : 0.5 * (output - target).square().sum()
        for (int64_t k = 0; k < T; ++k)
        {
            value += square(o[k] - S(t[k]));
        }
        return S(0.5) * value;

    case ORACLE_LOSS_MAE: // include/nano/loss/mae.h:19-22 : (output - target).abs().sum()
        for (int64_t k = 0; k < T; ++k)
        {
            value += std::fabs(o[k] - S(t[k]));
        }
        return value;

    case ORACLE_LOSS_CAUCHY: // include/nano/loss/cauchy.h : 0.5 * ((target - output).square() + 1).log().sum()
        for (int64_t k = 0; k < T; ++k)
        {
            value += std::log(square(S(t[k]) - o[k]) + S(1));
        }
        return S(0.5) * value;

    case ORACLE_LOSS_HINGE: // include/nano/loss/hinge.h:19-22 : (1 - target * output).max(0).sum()
        for (int64_t k = 0; k < T; ++k)
        {
            value += max_of(S(1) - S(t[k]) * o[k], S(0));
        }
        return value;

    case ORACLE_LOSS_SQUARED_HINGE: // include/nano/loss/squared_hinge.h : (1 - t*o).max(0).square().sum()
        for (int64_t k = 0; k < T; ++k)
        {
            value += square(max_of(S(1) - S(t[k]) * o[k], S(0)));
        }
        return value;

    case ORACLE_LOSS_CLASSNLL: // include/nano/loss/classnll.h:19-25
    {
        // log(sum(exp(o - omax))) - 0.5 * sum((1 + t) * o) + omax
        S omax = o[0];
        for (int64_t k = 1; k < T; ++k)
        {
            omax = max_of(omax, o[k]);
        }
        S sumexp = S(0), dot = S(0);
        for (int64_t k = 0; k < T; ++k)
        {
            sumexp += std::exp(o[k] - omax);
            dot += (S(1) + S(t[k])) * o[k];
        }
        return std::log(sumexp) - S(0.5) * dot + omax;
    }

    case ORACLE_LOSS_SAVAGE: // include/nano/loss/savage.h : (1 / (1 + exp(t*o)).square()).sum()
        for (int64_t k = 0; k < T; ++k)
        {
            value += S(1) / square(S(1) + std::exp(S(t[k]) * o[k]));
        }
        return value;

    case ORACLE_LOSS_TANGENT: // include/nano/loss/tangent.h : (2 * atan(t*o) - 1).square().sum()
        for (int64_t k = 0; k < T; ++k)
        {
            value += square(S(2) * std::atan(S(t[k]) * o[k]) - S(1));
        }
        return value;

    case ORACLE_LOSS_LOGISTIC: // include/nano/loss/logistic.h:19-28 (branch at x < 1 kept as is)
        for (int64_t k = 0; k < T; ++k)
        {
            const S x = -S(t[k]) * o[k];
            value += (x < S(1)) ? std::log1p(std::exp(x)) : (x + std::log1p(std::exp(-x)));
        }
        return value;

    case ORACLE_LOSS_EXPONENTIAL: // include/nano/loss/exponential.h : (-t*o).exp().sum()
        for (int64_t k = 0; k < T; ++k)
        {
            value += std::exp(-S(t[k]) * o[k]);
        }
        return value;

    case ORACLE_LOSS_PINBALL: // src/loss/pinball.cpp:24-35 : alpha*max(t-o,0) + (1-alpha)*max(o-t,0)
        for (int64_t k = 0; k < T; ++k)
        {
            value += param * max_of(S(t[k]) - o[k], S(0)) + (S(1) - param) * max_of(o[k] - S(t[k]), S(0));
        }
        return value;

    default: return std::nan("");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// per-sample d(loss)/d(output) over the T outputs of one sample
// ---------------------------------------------------------------------------------------------------------------
template <class S>
void loss_vgrad(const int kind, const S param, const double* t, const S* o, const int64_t T, S* g)
{
    switch (kind)
    {
    case ORACLE_LOSS_MSE: // mse.h:24-29
        for (int64_t k = 0; k < T; ++k)
        {
            g[k] = o[k] - S(t[k]);
        }
        break;

    case ORACLE_LOSS_MAE: // mae.h:24-29
        for (int64_t k = 0; k < T; ++k)
        {
            g[k] = sign_of(o[k] - S(t[k]));
        }
        break;

    case ORACLE_LOSS_CAUCHY: // cauchy.h vgrad : (o - t) / (1 + (o - t)^2)
        for (int64_t k = 0; k < T; ++k)
        {
            const S d = o[k] - S(t[k]);
            g[k]      = d / (S(1) + square(d));
        }
        break;

    case ORACLE_LOSS_HINGE: // hinge.h:24-29 : -t * (sign(1 - t*o) + 1) * 0.5  (=> -t/2 exactly on the kink)
        for (int64_t k = 0; k < T; ++k)
        {
            g[k] = -S(t[k]) * (sign_of(S(1) - S(t[k]) * o[k]) + S(1)) * S(0.5);
        }
        break;

    case ORACLE_LOSS_SQUARED_HINGE: // squared_hinge.h vgrad : -t * max(1 - t*o, 0) * 2
        for (int64_t k = 0; k < T; ++k)
        {
            g[k] = -S(t[k]) * max_of(S(1) - S(t[k]) * o[k], S(0)) * S(2);
        }
        break;

    case ORACLE_LOSS_CLASSNLL: // classnll.h:27-37 : softmax(o) - 0.5 * (1 + t)
    {
        S omax = o[0];
        for (int64_t k = 1; k < T; ++k)
        {
            omax = max_of(omax, o[k]);
        }
        S sumexp = S(0);
        for (int64_t k = 0; k < T; ++k)
        {
            g[k] = std::exp(o[k] - omax);
            sumexp += g[k];
        }
        for (int64_t k = 0; k < T; ++k)
        {
            g[k] = g[k] / sumexp - S(0.5) * (S(1) + S(t[k]));
        }
        break;
    }

    case ORACLE_LOSS_SAVAGE: // savage.h vgrad : -2t / ((1 + e^{to}) * (2 + e^{to} + e^{-to}))
        for (int64_t k = 0; k < T; ++k)
        {
            const S epos = std::exp(S(t[k]) * o[k]);
            const S eneg = std::exp(-S(t[k]) * o[k]);
            g[k]         = S(-2) * S(t[k]) / ((S(1) + epos) * (S(2) + epos + eneg));
        }
        break;

    case ORACLE_LOSS_TANGENT: // tangent.h vgrad : 4 t (2 atan(to) - 1) / (1 + (to)^2)
        for (int64_t k = 0; k < T; ++k)
        {
            const S to = S(t[k]) * o[k];
            g[k]       = S(4) * S(t[k]) * (S(2) * std::atan(to) - S(1)) / (S(1) + square(to));
        }
        break;

    case ORACLE_LOSS_LOGISTIC: // logistic.h:30-41
        for (int64_t k = 0; k < T; ++k)
        {
            const S x = -S(t[k]) * o[k];
            const S s = (x < S(1)) ? (std::exp(x) / (S(1) + std::exp(x))) : (S(1) / (S(1) + std::exp(-x)));
            g[k]      = -S(t[k]) * s;
        }
        break;

    case ORACLE_LOSS_EXPONENTIAL: // exponential.h vgrad : -t * exp(-t*o)
        for (int64_t k = 0; k < T; ++k)
        {
            g[k] = -S(t[k]) * std::exp(-S(t[k]) * o[k]);
        }
        break;

    case ORACLE_LOSS_PINBALL: // pinball.cpp:37-48 : -alpha + 0.5 * (1 - sign(t - o))
        for (int64_t k = 0; k < T; ++k)
        {
            g[k] = -param + S(0.5) * (S(1) - sign_of(S(t[k]) - o[k]));
        }
        break;

    default:
        for (int64_t k = 0; k < T; ++k)
        {
            g[k] = std::nan("");
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// dot products: the plain sequential one (oracle proper) and a SIMD-reassociated one (CPU baseline only)
// ---------------------------------------------------------------------------------------------------------------
inline double dot_plain(const double* a, const double* b, const int64_t n)
{
    double s = 0.0;
    for (int64_t j = 0; j < n; ++j)
    {
        s += a[j] * b[j];
    }
    return s;
}

// linear::accumulator_t (include/nano/linear/accumulator.h:11-50) without the Hessian buffers:
// [ fx | gb (T) | gw (T*D) ] packed in one vector so that sum_reduce is one loop.
struct accumulator_t
{
    accumulator_t(const int64_t D, const int64_t T)
        : m_data(static_cast<size_t>(1 + T + T * D), 0.0)
    {
    }

    double& fx() { return m_data[0]; }

    double* gb() { return m_data.data() + 1; }

    double* gw(const int64_t T) { return m_data.data() + 1 + T; }

    std::vector<double> m_data;
};

// the body of the lambda in linear::function_t::do_eval (src/linear/function.cpp:53-72) for one batch of rows
template <bool fast>
void eval_batch(const double* X, const double* Y, const int64_t D, const int64_t T, const int kind, const double param,
                const double* W, const double* b, const int64_t begin, const int64_t end, const bool has_grad,
                accumulator_t& acc, std::vector<double>& outputs, std::vector<double>& vgrads,
                std::vector<double>& gwbatch)
{
    const int64_t rows = end - begin;

    // linear::predict (src/linear/util.cpp:19-20): outputs = inputs * W^T, then += bias per row
    for (int64_t i = 0; i < rows; ++i)
    {
        const double* xi = X + (begin + i) * D;
        for (int64_t t = 0; t < T; ++t)
        {
            double s = 0.0;
            if constexpr (fast)
            {
#pragma omp simd reduction(+ : s)
                for (int64_t j = 0; j < D; ++j)
                {
                    s += xi[j] * W[t * D + j];
                }
            }
            else
            {
                s = dot_plain(xi, W + t * D, D);
            }
            outputs[static_cast<size_t>(i * T + t)] = s + b[t];
        }
    }

    // m_loss.value(...) then accumulator.m_fx += m_loss_fx.sum() (src/linear/function.cpp:60-62)
    double fsum = 0.0;
    for (int64_t i = 0; i < rows; ++i)
    {
        fsum += loss_value<double>(kind, param, Y + (begin + i) * T, outputs.data() + i * T, T);
    }
    acc.fx() += fsum;

    if (has_grad)
    {
        // m_loss.vgrad(...) (src/linear/function.cpp:66)
        for (int64_t i = 0; i < rows; ++i)
        {
            loss_vgrad<double>(kind, param, Y + (begin + i) * T, outputs.data() + i * T, T, vgrads.data() + i * T);
        }

        // m_gb += gmatrix.colwise().sum() (src/linear/function.cpp:69)
        double* gb = acc.gb();
        for (int64_t t = 0; t < T; ++t)
        {
            double s = 0.0;
            for (int64_t i = 0; i < rows; ++i)
            {
                s += vgrads[static_cast<size_t>(i * T + t)];
            }
            gb[t] += s;
        }

        // m_gw += gmatrix^T * inputs (src/linear/function.cpp:70): the batch product, then the accumulation
        std::fill(gwbatch.begin(), gwbatch.end(), 0.0);
        for (int64_t i = 0; i < rows; ++i)
        {
            const double* xi = X + (begin + i) * D;
            for (int64_t t = 0; t < T; ++t)
            {
                const double d  = vgrads[static_cast<size_t>(i * T + t)];
                double*      gt = gwbatch.data() + t * D;
                for (int64_t j = 0; j < D; ++j)
                {
                    gt[j] += d * xi[j];
                }
            }
        }
        double* gw = acc.gw(T);
        for (int64_t k = 0; k < T * D; ++k)
        {
            gw[k] += gwbatch[static_cast<size_t>(k)];
        }
    }
}

// regularisation + output of linear::function_t::do_eval (src/linear/function.cpp:127-143,158-167)
double finish_linear(const double* reduced /* [fx|gb|gw] already / N */, const int64_t D, const int64_t T,
                     const double l1, const double l2, const double* x, double* gx)
{
    const double* W     = x;
    const int64_t wsize = T * D;

    if (gx != nullptr)
    {
        for (int64_t k = 0; k < wsize; ++k)
        {
            gx[k] = reduced[1 + T + k];
        }
        for (int64_t t = 0; t < T; ++t)
        {
            gx[wsize + t] = reduced[1 + t];
        }
        if (l1 > 0.0)
        {
            for (int64_t k = 0; k < wsize; ++k)
            {
                gx[k] += l1 * sign_of(W[k]) / static_cast<double>(wsize);
            }
        }
        if (l2 > 0.0)
        {
            for (int64_t k = 0; k < wsize; ++k)
            {
                gx[k] += l2 * W[k] / static_cast<double>(wsize);
            }
        }
    }

    double fx = reduced[0];
    if (l1 > 0.0)
    {
        double s = 0.0;
        for (int64_t k = 0; k < wsize; ++k)
        {
            s += std::fabs(W[k]);
        }
        fx += l1 * (s / static_cast<double>(wsize));
    }
    if (l2 > 0.0)
    {
        const double sq = std::sqrt(l2);
        double       s  = 0.0;
        for (int64_t k = 0; k < wsize; ++k)
        {
            s += square(sq * W[k]);
        }
        fx += 0.5 * (s / static_cast<double>(wsize));
    }
    return fx;
}

// the loop + sum_reduce's additions (without the division): un-normalised [fx | gb | gw] over all accumulators
template <bool fast>
std::vector<double> linear_accumulate(const double* X, const double* Y, const int64_t N, const int64_t D,
                                      const int64_t T, const int kind, const double param, const double* x,
                                      const bool has_grad, int64_t batch, int threads)
{
    batch   = std::max<int64_t>(batch, 1);
    threads = std::max(threads, 1);

    const double* W = x;         // include/nano/linear/function.h:30-43
    const double* b = x + T * D; // include/nano/linear/function.h:45-59

    // m_accumulators(concurrency) + clear() (src/linear/function.cpp:29,51)
    std::vector<accumulator_t> accumulators(static_cast<size_t>(threads), accumulator_t{D, T});

    // flatten_iterator_t::loop -> pool.map(N, batch, op) (src/dataset/iterator.cpp:149-158,
    // include/nano/core/parallel.h:214-244). The reference hands batches to whichever worker is free; here batch
    // k goes to accumulator k % threads so that a (batch, threads) pair always gives the same bits.
    const int64_t nbatches = (N + batch - 1) / batch;
    const auto    worker   = [&](const int tnum)
    {
        std::vector<double> outputs(static_cast<size_t>(batch * T));
        std::vector<double> vgrads(static_cast<size_t>(batch * T));
        std::vector<double> gwbatch(static_cast<size_t>(T * D));
        for (int64_t k = tnum; k < nbatches; k += threads)
        {
            const int64_t begin = k * batch;
            const int64_t end   = std::min(begin + batch, N);
            eval_batch<fast>(X, Y, D, T, kind, param, W, b, begin, end, has_grad,
                             accumulators[static_cast<size_t>(tnum)], outputs, vgrads, gwbatch);
        }
    };

    if (threads == 1)
    {
        worker(0);
    }
    else
    {
        std::vector<std::thread> pool;
        pool.reserve(static_cast<size_t>(threads));
        for (int t = 0; t < threads; ++t)
        {
            pool.emplace_back(worker, t);
        }
        for (auto& th : pool)
        {
            th.join();
        }
    }

    // sum_reduce(m_accumulators, N): accumulator0 += accumulators[i] (include/nano/core/reduce.h:25-29)
    auto& acc0 = accumulators[0].m_data;
    for (size_t i = 1; i < accumulators.size(); ++i)
    {
        for (size_t k = 0; k < acc0.size(); ++k)
        {
            acc0[k] += accumulators[i].m_data[k];
        }
    }
    return std::move(acc0);
}

template <bool fast>
double linear_vgrad(const double* X, const double* Y, const int64_t N, const int64_t D, const int64_t T,
                    const int kind, const double param, const double l1, const double l2, const double* x, double* gx,
                    const int64_t batch, const int threads)
{
    auto acc0 = linear_accumulate<fast>(X, Y, N, D, T, kind, param, x, gx != nullptr, batch, threads);
    for (auto& v : acc0) // accumulator0 /= samples (include/nano/core/reduce.h:30)
    {
        v /= static_cast<double>(N);
    }
    return finish_linear(acc0.data(), D, T, l1, l2, x, gx);
}
} // namespace

extern "C" {

void oracle_loss_value(const int kind, const double param, const double* targets, const double* outputs,
                       const int64_t N, const int64_t T, double* values)
{
    // flatten_loss_t::do_value (include/nano/loss/flatten.h:69-75)
    for (int64_t i = 0; i < N; ++i)
    {
        values[i] = loss_value<double>(kind, param, targets + i * T, outputs + i * T, T);
    }
}

void oracle_loss_vgrad(const int kind, const double param, const double* targets, const double* outputs,
                       const int64_t N, const int64_t T, double* vgrads)
{
    // flatten_loss_t::do_vgrad (include/nano/loss/flatten.h:80-86)
    for (int64_t i = 0; i < N; ++i)
    {
        loss_vgrad<double>(kind, param, targets + i * T, outputs + i * T, T, vgrads + i * T);
    }
}

void oracle_loss_error(const int error_kind, const double* targets, const double* outputs, const int64_t N,
                       const int64_t T, double* errors)
{
    constexpr double epsilon = std::numeric_limits<double>::epsilon();
    for (int64_t i = 0; i < N; ++i)
    {
        const double* t = targets + i * T;
        const double* o = outputs + i * T;
        double        error = 0.0;
        if (error_kind == 0)
        {
            // absdiff_t (error.h:11-22): (target - output).abs().sum()
            for (int64_t k = 0; k < T; ++k)
            {
                error += std::fabs(t[k] - o[k]);
            }
        }
        else if (error_kind == 1 && T > 1)
        {
            // sclass_t, multi-class (error.h:55-61): the first maximum of the outputs must be a positive target
            int64_t idx = 0;
            for (int64_t k = 1; k < T; ++k)
            {
                idx = (o[k] > o[idx]) ? k : idx;
            }
            error = (t[idx] > 0.0) ? 0.0 : 1.0;
        }
        else
        {
            // mclass_t (error.h:28-40) and binary sclass_t (error.h:63-67): count of target * output < epsilon
            for (int64_t k = 0; k < T; ++k)
            {
                error += (t[k] * o[k] < epsilon) ? 1.0 : 0.0;
            }
        }
        errors[i] = error;
    }
}

void oracle_predict(const double* X, const double* W, const double* b, const int64_t N, const int64_t D,
                    const int64_t T, double* outputs)
{
    // src/linear/util.cpp:19-20
    for (int64_t i = 0; i < N; ++i)
    {
        for (int64_t t = 0; t < T; ++t)
        {
            outputs[i * T + t] = dot_plain(X + i * D, W + t * D, D) + b[t];
        }
    }
}

void oracle_sum_reduce(double* accumulators, const int64_t count, const int64_t len, const int64_t samples)
{
    // include/nano/core/reduce.h:23-31 with accumulator_t::operator+= / operator/= (src/linear/accumulator.cpp:27-47)
    for (int64_t i = 1; i < count; ++i)
    {
        for (int64_t k = 0; k < len; ++k)
        {
            accumulators[k] += accumulators[i * len + k];
        }
    }
    for (int64_t k = 0; k < len; ++k)
    {
        accumulators[k] /= static_cast<double>(samples);
    }
}

double oracle_linear_vgrad(const double* X, const double* Y, const int64_t N, const int64_t D, const int64_t T,
                           const int kind, const double param, const double l1, const double l2, const double* x,
                           double* gx, const int64_t batch, const int threads)
{
    return linear_vgrad<false>(X, Y, N, D, T, kind, param, l1, l2, x, gx, batch, threads);
}

double oracle_linear_vgrad_fast(const double* X, const double* Y, const int64_t N, const int64_t D, const int64_t T,
                                const int kind, const double param, const double l1, const double l2,
                                const double* x, double* gx, const int64_t batch, const int threads)
{
    return linear_vgrad<true>(X, Y, N, D, T, kind, param, l1, l2, x, gx, batch, threads);
}

double oracle_linear_vgrad_ld(const double* X, const double* Y, const int64_t N, const int64_t D, const int64_t T,
                              const int kind, const double param, const double l1, const double l2, const double* x,
                              double* gx)
{
    // same formulas as src/linear/function.cpp:44-168, every sum and every loss in long double: the "truth" the
    // fp64 paths (oracle and CUDA) are bounded against.
    using ld = long double;

    const double* W = x;
    const double* b = x + T * D;

    std::vector<ld> o(static_cast<size_t>(T)), d(static_cast<size_t>(T));
    std::vector<ld> gw(static_cast<size_t>(T * D), 0.0L), gb(static_cast<size_t>(T), 0.0L);
    ld              fsum = 0.0L;

    for (int64_t i = 0; i < N; ++i)
    {
        const double* xi = X + i * D;
        for (int64_t t = 0; t < T; ++t)
        {
            ld s = 0.0L;
            for (int64_t j = 0; j < D; ++j)
            {
                s += static_cast<ld>(xi[j]) * static_cast<ld>(W[t * D + j]);
            }
            o[static_cast<size_t>(t)] = s + static_cast<ld>(b[t]);
        }
        fsum += loss_value<ld>(kind, static_cast<ld>(param), Y + i * T, o.data(), T);
        if (gx != nullptr)
        {
            loss_vgrad<ld>(kind, static_cast<ld>(param), Y + i * T, o.data(), T, d.data());
            for (int64_t t = 0; t < T; ++t)
            {
                gb[static_cast<size_t>(t)] += d[static_cast<size_t>(t)];
                for (int64_t j = 0; j < D; ++j)
                {
                    gw[static_cast<size_t>(t * D + j)] += d[static_cast<size_t>(t)] * static_cast<ld>(xi[j]);
                }
            }
        }
    }

    const ld wsize = static_cast<ld>(T * D);
    const ld n     = static_cast<ld>(N);
    if (gx != nullptr)
    {
        for (int64_t k = 0; k < T * D; ++k)
        {
            ld g = gw[static_cast<size_t>(k)] / n;
            if (l1 > 0.0)
            {
                g += static_cast<ld>(l1) * sign_of(static_cast<ld>(W[k])) / wsize;
            }
            if (l2 > 0.0)
            {
                g += static_cast<ld>(l2) * static_cast<ld>(W[k]) / wsize;
            }
            gx[k] = static_cast<double>(g);
        }
        for (int64_t t = 0; t < T; ++t)
        {
            gx[T * D + t] = static_cast<double>(gb[static_cast<size_t>(t)] / n);
        }
    }

    ld fx = fsum / n;
    if (l1 > 0.0)
    {
        ld s = 0.0L;
        for (int64_t k = 0; k < T * D; ++k)
        {
            s += std::fabs(static_cast<ld>(W[k]));
        }
        fx += static_cast<ld>(l1) * s / wsize;
    }
    if (l2 > 0.0)
    {
        ld s = 0.0L;
        for (int64_t k = 0; k < T * D; ++k)
        {
            s += square(static_cast<ld>(W[k]));
        }
        fx += 0.5L * static_cast<ld>(l2) * s / wsize;
    }
    return static_cast<double>(fx);
}

void oracle_linear_partial(const double* X, const double* Y, const int64_t N, const int64_t D, const int64_t T,
                           const int kind, const double param, const double* x, const int want_grad,
                           const int64_t batch, const int threads, double* partial)
{
    const auto acc0 = linear_accumulate<false>(X, Y, N, D, T, kind, param, x, want_grad != 0, batch, threads);
    std::memcpy(partial, acc0.data(), acc0.size() * sizeof(double));
}

double oracle_linear_finish(const double* reduced, const int64_t n_global, const int64_t D, const int64_t T,
                            const double l1, const double l2, const double* x, double* gx)
{
    std::vector<double> scaled(reduced, reduced + 1 + T + T * D);
    for (auto& v : scaled)
    {
        v /= static_cast<double>(n_global);
    }
    return finish_linear(scaled.data(), D, T, l1, l2, x, gx);
}

void oracle_scalar_stats(const double* values, const int64_t N, const int64_t C, const unsigned char* enable,
                         double* stats /* [7][C]: samples, min, max, mean, stdev, div_range, div_stdev */)
{
    // update + done of src/dataset/stats.cpp:81-137 (sequential over samples, finite values only)
    const double epsilon = 1e-8; // epsilon2
    double* samples = stats;
    double* vmin    = stats + C;
    double* vmax    = stats + 2 * C;
    double* mean    = stats + 3 * C;
    double* stdev   = stats + 4 * C;
    double* drange  = stats + 5 * C;
    double* dstdev  = stats + 6 * C;
    for (int64_t c = 0; c < C; ++c)
    {
        samples[c] = 0.0;
        vmin[c]    = std::numeric_limits<double>::max();
        vmax[c]    = std::numeric_limits<double>::lowest();
        mean[c]    = 0.0;
        stdev[c]   = 0.0;
    }
    for (int64_t i = 0; i < N; ++i)
    {
        for (int64_t c = 0; c < C; ++c)
        {
            const double value = values[i * C + c];
            if (std::isfinite(value))
            {
                samples[c] += 1.0;
                mean[c] += value;
                stdev[c] += value * value;
                vmin[c] = std::min(vmin[c], value);
                vmax[c] = std::max(vmax[c], value);
            }
        }
    }
    for (int64_t c = 0; c < C; ++c)
    {
        if (const double n = samples[c]; n > 1.0)
        {
            stdev[c]  = std::sqrt((stdev[c] - mean[c] * mean[c] / n) / (n - 1.0));
            mean[c]   = mean[c] / n;
            drange[c] = 1.0 / std::max(vmax[c] - vmin[c], epsilon);
            dstdev[c] = 1.0 / std::max(stdev[c], epsilon);
        }
        else
        {
            if (n == 0.0)
            {
                vmin[c] = vmax[c] = mean[c] = 0.0;
            }
            stdev[c]  = 0.0;
            drange[c] = 1.0;
            dstdev[c] = 1.0;
        }
        if (enable != nullptr && enable[c] == 0)
        {
            vmin[c] = vmax[c] = mean[c] = stdev[c] = 0.0;
            drange[c] = dstdev[c] = 1.0;
        }
    }
}

void oracle_scale(double* values, const int64_t N, const int64_t C, const double* stats, const int scaling)
{
    // scalar_stats_t::scale (src/dataset/stats.cpp:357-401) + nan2zero (:8-22)
    const double* vmin   = stats + C;
    const double* mean   = stats + 3 * C;
    const double* drange = stats + 5 * C;
    const double* dstdev = stats + 6 * C;
    for (int64_t i = 0; i < N; ++i)
    {
        for (int64_t c = 0; c < C; ++c)
        {
            double& value = values[i * C + c];
            switch (scaling)
            {
            case 1: value = (value - mean[c]) * drange[c]; break;
            case 2: value = (value - vmin[c]) * drange[c]; break;
            case 3: value = (value - mean[c]) * dstdev[c]; break;
            default: break;
            }
            if (!std::isfinite(value))
            {
                value = 0.0;
            }
        }
    }
}

void oracle_upscale(const double* istats, const int iscaling, const double* tstats, const int tscaling, const int64_t D,
                    const int64_t T, double* W, double* b)
{
    // nano::upscale (src/dataset/stats.cpp:211-230) with make_scaling (:46-78)
    const auto affine = [](const double* stats, const int64_t C, const int scaling, std::vector<double>& w,
                           std::vector<double>& off)
    {
        w.assign(static_cast<size_t>(C), 1.0);
        off.assign(static_cast<size_t>(C), 0.0);
        for (int64_t c = 0; c < C; ++c)
        {
            const double vmin = stats[C + c], mean = stats[3 * C + c], drange = stats[5 * C + c],
                         dstdev = stats[6 * C + c];
            switch (scaling)
            {
            case 1: w[c] = drange; off[c] = -mean * drange; break;
            case 2: w[c] = drange; off[c] = -vmin * drange; break;
            case 3: w[c] = dstdev; off[c] = -mean * dstdev; break;
            default: break;
            }
        }
    };
    std::vector<double> fw, fb, tw, tb;
    affine(istats, D, iscaling, fw, fb);
    affine(tstats, T, tscaling, tw, tb);
    for (int64_t t = 0; t < T; ++t)
    {
        b[t] = (dot_plain(W + t * D, fb.data(), D) + b[t] - tb[t]) / tw[t];
        for (int64_t j = 0; j < D; ++j)
        {
            W[t * D + j] = W[t * D + j] / tw[t] * fw[j];
        }
    }
}

int64_t oracle_synth_samples(const int64_t dims, const double sratio)
{
    // make_samples (src/function/mlearn/ridge.cpp:23-26)
    return static_cast<int64_t>(std::max(sratio * static_cast<double>(dims), 10.0));
}

void oracle_synth_model(const int64_t samples, const int64_t inputs, const uint64_t seed, const int64_t modulo,
                        const int regression, double* X, double* Y, double* wopt, double* bopt)
{
    // linear_model_t ctor (src/function/mlearn/linear.cpp:5-45) with outputs == 1 (ridge.cpp:18-21);
    // rng = std::mt19937_64(seed) (src/core/random.cpp:6-17), udist = uniform_real_distribution<double>(0, 1).
    std::mt19937_64                        rng(seed);
    std::uniform_real_distribution<double> udist(0.0, 1.0);

    for (int64_t k = 0; k < samples * inputs; ++k) // m_inputs.full(...), row-major
    {
        X[k] = udist(rng);
    }
    for (int64_t j = 0; j < inputs; ++j) // m_woptimum.full(...)
    {
        wopt[j] = udist(rng);
    }
    bopt[0] = udist(rng) - 0.5; // m_boptimum.full(...)

    double wsum = 0.0; // m_woptimum.row(o) /= m_woptimum.row(o).sum()
    for (int64_t j = 0; j < inputs; ++j)
    {
        wsum += wopt[j];
    }
    for (int64_t j = 0; j < inputs; ++j)
    {
        wopt[j] /= wsum;
    }
    for (int64_t j = 0; j < inputs; ++j) // columns with i % modulo != 0 are zeroed
    {
        if (j % modulo != 0)
        {
            wopt[j] = 0.0;
        }
    }

    for (int64_t i = 0; i < samples; ++i) // eval_outputs(m_woptimum) (linear.cpp:52-56)
    {
        const double o = dot_plain(X + i * inputs, wopt, inputs) + bopt[0];
        Y[i]           = regression ? o : sign_of((o - bopt[0]) - 0.5); // linear.cpp:37-44
    }
}

double oracle_synth_vgrad(const double* X, const double* Y, const double bopt, const int64_t N, const int64_t D,
                          const int synth_loss, const double alpha1, const double alpha2, const double* x, double* gx)
{
    // linear_model_t::eval<tloss> (src/function/mlearn/linear.h:21-38): outputs, then gx, then fx
    std::vector<double> o(static_cast<size_t>(N)), g(static_cast<size_t>(N));
    for (int64_t i = 0; i < N; ++i) // eval_outputs (linear.cpp:52-56)
    {
        o[static_cast<size_t>(i)] = dot_plain(X + i * D, x, D) + bopt;
    }

    const double samples = static_cast<double>(N);
    double       fx      = 0.0;
    switch (synth_loss)
    {
    case ORACLE_SYNTH_MSE: // src/function/mlearn/loss.cpp:5-18
        for (int64_t i = 0; i < N; ++i)
        {
            const double delta = o[i] - Y[i];
            fx += delta * delta;
            g[i] = delta;
        }
        fx = 0.5 * fx / samples;
        break;

    case ORACLE_SYNTH_MAE: // loss.cpp:30-43
        for (int64_t i = 0; i < N; ++i)
        {
            const double delta = o[i] - Y[i];
            fx += std::fabs(delta);
            g[i] = sign_of(delta);
        }
        fx = fx / samples;
        break;

    case ORACLE_SYNTH_CAUCHY: // loss.cpp:50-63 (no 0.5 factor here, unlike loss_t's cauchy)
        for (int64_t i = 0; i < N; ++i)
        {
            const double delta = o[i] - Y[i];
            fx += std::log(delta * delta + 1.0);
            g[i] = 2.0 * delta / (1.0 + delta * delta);
        }
        fx = fx / samples;
        break;

    case ORACLE_SYNTH_HINGE: // loss.cpp:77-90 : edges = -o*y ; max(1 + edges, 0) ; -y * (sign(1+edges)*0.5 + 0.5)
        for (int64_t i = 0; i < N; ++i)
        {
            const double edge = -o[i] * Y[i];
            fx += max_of(1.0 + edge, 0.0);
            g[i] = -Y[i] * (sign_of(1.0 + edge) * 0.5 + 0.5);
        }
        fx = fx / samples;
        break;

    case ORACLE_SYNTH_LOGISTIC: // loss.cpp:97-125 (gx is the un-stabilised -y*e/(1+e))
        for (int64_t i = 0; i < N; ++i)
        {
            const double xx = -o[i] * Y[i];
            fx += (xx < 1.0) ? std::log1p(std::exp(xx)) : (xx + std::log1p(std::exp(-xx)));
            const double edge = std::exp(-o[i] * Y[i]);
            g[i]              = (-Y[i] * edge) / (1.0 + edge);
        }
        fx = fx / samples;
        break;

    default: return std::nan("");
    }

    if (gx != nullptr)
    {
        // eval_grad: gw = G^T X / samples (linear.cpp:58-67)
        for (int64_t j = 0; j < D; ++j)
        {
            gx[j] = 0.0;
        }
        for (int64_t i = 0; i < N; ++i)
        {
            const double  d  = g[static_cast<size_t>(i)];
            const double* xi = X + i * D;
            for (int64_t j = 0; j < D; ++j)
            {
                gx[j] += d * xi[j];
            }
        }
        for (int64_t j = 0; j < D; ++j)
        {
            gx[j] /= samples;
        }
        // ridge.cpp:69-72, lasso.cpp:69-72, elasticnet.cpp:74-77
        for (int64_t j = 0; j < D; ++j)
        {
            gx[j] += alpha1 * sign_of(x[j]) + alpha2 * x[j];
        }
    }

    // ridge.cpp:79 : 0.5 * (sqrt(alpha2) * x).squaredNorm() ; lasso.cpp:75 : alpha1 * x.lpNorm<1>()
    double l1norm = 0.0, sqnorm = 0.0;
    for (int64_t j = 0; j < D; ++j)
    {
        l1norm += std::fabs(x[j]);
        sqnorm += square(std::sqrt(alpha2) * x[j]);
    }
    fx += alpha1 * l1norm + 0.5 * sqnorm;
    return fx;
}

// gboost::{bias,scale,grads}_function_t::do_eval (src/gboost/function.cpp:26-50, 85-126, 151-223): targets_iterator_t
// serves the samples in batches; per batch the outputs are formed, loss_t::value / vgrad run over the batch and the
// accumulator sums values and gradients; sum_reduce divides by the sample count (src/gboost/accumulator.cpp:20-47).
// kind: 0 bias (n = T), 1 scale (n = groups), 2 grads (n = N * T). One accumulator (one worker).
double oracle_gboost_vgrad(const int kind, const int loss, const double param, const double* targets, const int64_t N,
                           const int64_t T, const int32_t* cluster, const int64_t groups, const double* soutputs,
                           const double* woutputs, const double* x, double* gx, const int64_t batch)
{
    const int64_t       n = (kind == 0) ? T : (kind == 1 ? groups : N * T);
    std::vector<double> outputs(static_cast<size_t>(batch * T)), vgrads(static_cast<size_t>(batch * T));
    std::vector<double> acc_g(static_cast<size_t>(kind == 2 ? 0 : n), 0.0);
    double              acc_f = 0.0;
    for (int64_t begin = 0; begin < N; begin += batch)
    {
        const int64_t end = std::min(N, begin + batch);
        for (int64_t i = begin; i < end; ++i)
        {
            for (int64_t t = 0; t < T; ++t)
            {
                double o;
                if (kind == 0)
                {
                    o = x[t]; // output = bias (function.cpp:102-103)
                }
                else if (kind == 1)
                {
                    const int32_t group = cluster[i];
                    const double  scale = (group < 0) ? 0.0 : x[group]; // function.cpp:173-177
                    o                   = soutputs[i * T + t] + scale * woutputs[i * T + t];
                }
                else
                {
                    o = x[i * T + t];
                }
                outputs[static_cast<size_t>((i - begin) * T + t)] = o;
            }
        }
        double batch_f = 0.0; // accumulator.m_loss_fx.sum()
        for (int64_t i = begin; i < end; ++i)
        {
            batch_f += loss_value<double>(loss, param, targets + i * T, outputs.data() + (i - begin) * T, T);
        }
        acc_f += batch_f;
        if (gx != nullptr)
        {
            for (int64_t i = begin; i < end; ++i)
            {
                double* g = vgrads.data() + (i - begin) * T;
                loss_vgrad<double>(loss, param, targets + i * T, outputs.data() + (i - begin) * T, T, g);
                if (kind == 2)
                {
                    for (int64_t t = 0; t < T; ++t)
                    {
                        gx[i * T + t] = g[t] / static_cast<double>(N); // grads.vector() / denom (function.cpp:36)
                    }
                }
            }
            if (kind == 0)
            {
                // m_gx += gmatrix.colwise().sum() (function.cpp:112-114): per batch, rows in order
                for (int64_t t = 0; t < T; ++t)
                {
                    double column = 0.0;
                    for (int64_t i = begin; i < end; ++i)
                    {
                        column += vgrads[static_cast<size_t>((i - begin) * T + t)];
                    }
                    acc_g[static_cast<size_t>(t)] += column;
                }
            }
            else if (kind == 1)
            {
                for (int64_t i = begin; i < end; ++i) // function.cpp:188-200
                {
                    const int32_t group = cluster[i];
                    if (group < 0)
                    {
                        continue;
                    }
                    acc_g[static_cast<size_t>(group)] += dot_plain(vgrads.data() + (i - begin) * T, woutputs + i * T, T);
                }
            }
        }
    }
    if (gx != nullptr && kind != 2)
    {
        for (int64_t k = 0; k < n; ++k)
        {
            gx[k] = acc_g[static_cast<size_t>(k)] / static_cast<double>(N);
        }
    }
    return acc_f / static_cast<double>(N);
}

int oracle_hardware_threads(void)
{
    const auto n = std::thread::hardware_concurrency();
    return n == 0 ? 1 : static_cast<int>(n);
}

} // extern "C"
