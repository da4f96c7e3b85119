// gboost.cuh — the gradient-boosting cousins of the linear objective: the same per-sample loss
// value / vgrad and the same accumulate-then-sum_reduce structure as linear::function_t::do_eval, without a feature
// matrix.
//
//   gboost::bias_function_t::do_eval   src/gboost/function.cpp:85-126   f(x) = mean_i loss(y_i, x),            n = T
//       g = mean_i dL(y_i, x)
//   gboost::scale_function_t::do_eval  src/gboost/function.cpp:151-223  f(x) = mean_i loss(y_i, s_i + x[c_i] w_i), n = K
//       g[k] = (1/N) sum_{i: c_i = k} dL_i . w_i ; samples outside every cluster (c_i < 0) get scale 0, no gradient
//   gboost::grads_function_t::do_eval  src/gboost/function.cpp:26-50    f(o) = mean_i loss(y_i, o_i),         n = N*T
//       g = dL / N  (the per-sample gradients the weak learners are fitted to)
//   accumulators: src/gboost/accumulator.cpp:20-47 (sum, then / samples), sum_reduce include/nano/core/reduce.h:23-31
//
// One streaming pass over targets (+ strong / weak outputs + cluster ids): a warp handles 32 / TP samples per
// iteration (TP = the power of two >= T, capped at 32), lane = (sample of the group, target); sums inside a sample are
// shuffles among its TP lanes. Per-block partials [sum loss | sum gradient], folded in block order.
#pragma once

#include "common.cuh"
#include "loss.cuh"

namespace nb200
{
constexpr int kGbWarps = 8;

enum gboost_kind : int
{
    GBOOST_BIAS  = 0,
    GBOOST_SCALE = 1,
    GBOOST_GRADS = 2
};

struct gboost_params
{
    const double* targets;  // [N, T]
    const double* soutputs; // [N, T] (scale)
    const double* woutputs; // [N, T] (scale)
    const int*    cluster;  // [N] group of each sample or -1 (scale)
    const double* x;        // [n] device
    double*       partials; // [blocks, 1 + n_out]; n_out = T (bias), K (scale), 0 (grads)
    double*       gx;       // grads: [N, T] written directly (already / N)
    int64_t       N;
    int           T;
    int           K;
    int           kind;
    int           loss;
    double        loss_param;
};

#ifdef __CUDACC__

// sum over the TP lanes of a sample (TP a power of two <= 32); every lane of the group gets the same bits
__device__ __forceinline__ double group_sum(double v, const int tp)
{
    for (int offset = tp >> 1; offset > 0; offset >>= 1)
    {
        v += __shfl_xor_sync(0xffffffffu, v, offset);
    }
    return v;
}

__device__ __forceinline__ double group_max(double v, const int tp)
{
    for (int offset = tp >> 1; offset > 0; offset >>= 1)
    {
        const double other = __shfl_xor_sync(0xffffffffu, v, offset);
        v                  = (v < other) ? other : v;
    }
    return v;
}

// dynamic shared memory: acc[kGbWarps][n_out] (scale: per-cluster sums; bias with T > 32: per-target sums), f[kGbWarps]
template <bool GRAD>
__global__ void __launch_bounds__(kGbWarps * 32) gboost_kernel(const gboost_params p)
{
    extern __shared__ __align__(16) unsigned char gb_smem[];
    const int T     = p.T;
    const int n_out = (p.kind == GBOOST_BIAS) ? T : (p.kind == GBOOST_SCALE ? p.K : 0);
    double*   acc_all = reinterpret_cast<double*>(gb_smem);
    double*   f_all   = acc_all + static_cast<int64_t>(kGbWarps) * n_out;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    double*   acc  = acc_all + static_cast<int64_t>(warp) * n_out;
    for (int k = lane; k < n_out; k += 32)
    {
        acc[k] = 0.0;
    }
    __syncwarp();

    int tp = 1;
    while (tp < T && tp < 32)
    {
        tp <<= 1;
    }
    const int per_iter = 32 / tp;  // samples per warp iteration
    const int sub      = lane % tp; // first target of this lane (then sub + tp, ... when T > 32)
    const int mine     = lane / tp; // sample of the iteration this lane belongs to

    double fsum = 0.0; // owner lanes (sub == 0) carry the loss sums of their samples
    double gacc = 0.0; // bias, T <= 32: this lane's target over its samples

    const int64_t groups      = (p.N + per_iter - 1) / per_iter;
    const int64_t warps_total = static_cast<int64_t>(gridDim.x) * kGbWarps;
    for (int64_t it = static_cast<int64_t>(blockIdx.x) * kGbWarps + warp; it < groups; it += warps_total)
    {
        const int64_t row   = it * per_iter + mine;
        const bool    valid = row < p.N;
        int           group = -1;
        double        scale = 0.0;
        if (p.kind == GBOOST_SCALE && valid)
        {
            group = p.cluster[row];
            scale = (group < 0) ? 0.0 : p.x[group]; // function.cpp:173-174
        }

        // outputs of this sample (this lane's targets), loss terms
        double rowsum = 0.0, contrib = 0.0;
        if (p.loss == LOSS_CLASSNLL)
        {
            // classnll.h:19-37 over the sample's T outputs (two sweeps: max, then exp sums)
            double omax = -INFINITY;
            for (int t = sub; t < T; t += tp)
            {
                if (valid)
                {
                    const int64_t k = row * T + t;
                    const double  o = (p.kind == GBOOST_BIAS) ? p.x[t]
                                      : (p.kind == GBOOST_GRADS ? p.x[k] : __dadd_rn(p.soutputs[k], __dmul_rn(scale, p.woutputs[k])));
                    omax            = max_of(omax, o);
                }
            }
            omax = group_max(omax, tp);
            double sumexp = 0.0, dot = 0.0;
            for (int t = sub; t < T; t += tp)
            {
                if (valid)
                {
                    const int64_t k = row * T + t;
                    const double  o = (p.kind == GBOOST_BIAS) ? p.x[t]
                                      : (p.kind == GBOOST_GRADS ? p.x[k] : __dadd_rn(p.soutputs[k], __dmul_rn(scale, p.woutputs[k])));
                    sumexp += exp(__dsub_rn(o, omax));
                    dot += __dmul_rn(__dadd_rn(1.0, p.targets[k]), o);
                }
            }
            sumexp = group_sum(sumexp, tp);
            dot    = group_sum(dot, tp);
            rowsum = __dadd_rn(__dsub_rn(log(sumexp), __dmul_rn(0.5, dot)), omax);
            if (GRAD)
            {
                for (int t = sub; t < T; t += tp)
                {
                    if (valid)
                    {
                        const int64_t k = row * T + t;
                        const double  o = (p.kind == GBOOST_BIAS) ? p.x[t]
                                          : (p.kind == GBOOST_GRADS ? p.x[k] : __dadd_rn(p.soutputs[k], __dmul_rn(scale, p.woutputs[k])));
                        const double  g = __dsub_rn(__ddiv_rn(exp(__dsub_rn(o, omax)), sumexp),
                                                    __dmul_rn(0.5, __dadd_rn(1.0, p.targets[k])));
                        if (p.kind == GBOOST_GRADS)
                        {
                            p.gx[k] = __ddiv_rn(g, static_cast<double>(p.N));
                        }
                        else if (p.kind == GBOOST_BIAS)
                        {
                            if (T <= 32)
                            {
                                gacc += g;
                            }
                            else
                            {
                                acc[t] += g; // one sample per iteration: target t belongs to this lane alone
                            }
                        }
                        else
                        {
                            contrib = fma(g, p.woutputs[k], contrib);
                        }
                    }
                }
            }
        }
        else
        {
            double sum = 0.0;
            for (int t = sub; t < T; t += tp)
            {
                if (valid)
                {
                    const int64_t k = row * T + t;
                    const double  o = (p.kind == GBOOST_BIAS) ? p.x[t]
                                      : (p.kind == GBOOST_GRADS ? p.x[k] : __dadd_rn(p.soutputs[k], __dmul_rn(scale, p.woutputs[k])));
                    double term, g;
                    loss_term<GRAD>(p.loss, p.loss_param, p.targets[k], o, term, g);
                    sum += term;
                    if (GRAD)
                    {
                        if (p.kind == GBOOST_GRADS)
                        {
                            p.gx[k] = __ddiv_rn(g, static_cast<double>(p.N));
                        }
                        else if (p.kind == GBOOST_BIAS)
                        {
                            if (T <= 32)
                            {
                                gacc += g;
                            }
                            else
                            {
                                acc[t] += g;
                            }
                        }
                        else
                        {
                            contrib = fma(g, p.woutputs[k], contrib);
                        }
                    }
                }
            }
            rowsum = loss_post(p.loss, group_sum(sum, tp));
        }
        if (valid && sub == 0)
        {
            fsum += rowsum;
        }
        if (GRAD && p.kind == GBOOST_SCALE)
        {
            // g[group] += dL . woutput (function.cpp:196-200); the samples of one iteration take turns, in order
            contrib = group_sum(contrib, tp);
            for (int r = 0; r < per_iter; ++r)
            {
                if (mine == r && sub == 0 && valid && group >= 0)
                {
                    acc[group] += contrib;
                }
                __syncwarp();
            }
        }
    }

    // fold the lanes of the warp, then the warps of the block, in a fixed order
    fsum = warp_allreduce_sum(fsum);
    if (GRAD && p.kind == GBOOST_BIAS && T <= 32)
    {
        for (int offset = tp; offset < 32; offset <<= 1) // lanes with the same target
        {
            gacc += __shfl_xor_sync(0xffffffffu, gacc, offset);
        }
        if (lane < T)
        {
            acc[lane] = gacc;
        }
    }
    if (lane == 0)
    {
        f_all[warp] = fsum;
    }
    __syncthreads();
    double* out = p.partials + static_cast<int64_t>(blockIdx.x) * (1 + n_out);
    if (threadIdx.x == 0)
    {
        double f = 0.0;
        for (int q = 0; q < kGbWarps; ++q)
        {
            f += f_all[q];
        }
        out[0] = f;
    }
    if (GRAD)
    {
        for (int k = threadIdx.x; k < n_out; k += blockDim.x)
        {
            double s = 0.0;
            for (int q = 0; q < kGbWarps; ++q)
            {
                s += acc_all[static_cast<int64_t>(q) * n_out + k];
            }
            out[1 + k] = s;
        }
    }
}

// out[0] = f = reduced[0] / N ; gx[k] = reduced[1 + k] / N  (accumulator_t::operator/=, accumulator.cpp:30-36)
__global__ void __launch_bounds__(256) gboost_finish_kernel(const double* __restrict__ reduced, const int64_t n_out,
                                                            const int64_t N, double* __restrict__ gx,
                                                            double* __restrict__ fx)
{
    const int64_t k       = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const double  samples = static_cast<double>(N);
    if (k == 0)
    {
        *fx = __ddiv_rn(reduced[0], samples);
    }
    if (gx != nullptr && k < n_out)
    {
        gx[k] = __ddiv_rn(reduced[1 + k], samples);
    }
}

#endif // __CUDACC__
} // namespace nb200
