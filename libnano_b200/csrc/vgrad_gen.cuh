// vgrad_gen.cuh — the shape-generic path (any D, any T >= 1): forward + loss, then the dL^T X reduction.
//
// Same reference call sites as vgrad_t1.cuh (src/linear/function.cpp:53-72, src/linear/util.cpp:19-20,
// include/nano/loss/flatten.h:69-86); used for multi-target models and for feature counts the register-resident
// single-target kernel does not cover. Two kernels, X read twice (the many-class case is contraction-bound, not
// HBM-bound: SURVEY.md §8d):
//   gen_forward_kernel   O = X W^T + b per row, loss value, dL written to HBM, per-block [loss | colsum(dL)]
//   gen_backward_kernel  gW = dL^T X over a (column tile, target tile, row split) grid, per-split partials
// Both produce partials that reduce_partials_kernel folds in a fixed order.
#pragma once

#include "common.cuh"
#include "loss.cuh"

namespace nb200
{
constexpr int kGenWarps = 8;  // warps per forward block
constexpr int kGenTB    = 4;  // targets per register block in the forward pass
constexpr int kBwdTB    = 8;  // targets per register block in the backward pass
constexpr int kBwdCols  = 128; // columns per backward block (one per thread)

struct gen_params
{
    const double* X;     // [N, D]
    const double* Y;     // [N, T]
    const double* W;     // [T, D] device
    const double* b;     // [T] device
    double*       dL;    // [N, T] scratch (nullptr for value-only)
    double*       partials_fb; // [blocks, 1 + T] : loss sum | dL column sums
    int64_t       N;
    int64_t       D;
    int64_t       T;
    int           loss;
    double        loss_param;
};

#ifdef __CUDACC__

// dynamic shared memory: outputs o_s[kGenWarps][T], then gb_s[kGenWarps][T], then fsum_s[kGenWarps]
template <bool GRAD>
__global__ void __launch_bounds__(kGenWarps * 32) gen_forward_kernel(const gen_params p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t T      = p.T;
    const int64_t D      = p.D;
    double*       o_all  = reinterpret_cast<double*>(smem);
    double*       gb_all = o_all + kGenWarps * T;
    double*       f_all  = gb_all + kGenWarps * T;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    double*   o_s  = o_all + warp * T;
    double*   gb_s = gb_all + warp * T;

    for (int64_t t = lane; t < T; t += 32)
    {
        gb_s[t] = 0.0;
    }
    double fsum = 0.0;

    const int64_t warps_total = static_cast<int64_t>(gridDim.x) * kGenWarps;
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * kGenWarps + warp; row < p.N; row += warps_total)
    {
        const double* xrow = p.X + row * D;
        const double* yrow = p.Y + row * T;

        // predict: kGenTB targets per sweep over the row
        for (int64_t t0 = 0; t0 < T; t0 += kGenTB)
        {
            double acc[kGenTB];
#pragma unroll
            for (int j = 0; j < kGenTB; ++j)
            {
                acc[j] = 0.0;
            }
            for (int64_t c = lane; c < D; c += 32)
            {
                const double xv = xrow[c];
#pragma unroll
                for (int j = 0; j < kGenTB; ++j)
                {
                    if (t0 + j < T)
                    {
                        acc[j] = fma(xv, p.W[(t0 + j) * D + c], acc[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kGenTB; ++j)
            {
                const double dot = warp_allreduce_sum(acc[j]);
                if (t0 + j < T && lane == 0)
                {
                    o_s[t0 + j] = dot + p.b[t0 + j];
                }
            }
        }
        __syncwarp();

        // loss value and dL for this sample
        if (p.loss == LOSS_CLASSNLL)
        {
            // classnll.h:19-37
            double omax = -INFINITY;
            for (int64_t t = lane; t < T; t += 32)
            {
                omax = max_of(omax, o_s[t]);
            }
            omax = warp_allreduce_max(omax);
            double sumexp = 0.0, dot = 0.0;
            for (int64_t t = lane; t < T; t += 32)
            {
                sumexp += exp(__dsub_rn(o_s[t], omax));
                dot += __dmul_rn(__dadd_rn(1.0, yrow[t]), o_s[t]);
            }
            sumexp = warp_allreduce_sum(sumexp);
            dot    = warp_allreduce_sum(dot);
            fsum += __dadd_rn(__dsub_rn(log(sumexp), __dmul_rn(0.5, dot)), omax);
            if (GRAD)
            {
                for (int64_t t = lane; t < T; t += 32)
                {
                    const double e = exp(__dsub_rn(o_s[t], omax));
                    const double g = __dsub_rn(__ddiv_rn(e, sumexp), __dmul_rn(0.5, __dadd_rn(1.0, yrow[t])));
                    p.dL[row * T + t] = g;
                    gb_s[t] += g;
                }
            }
        }
        else
        {
            double sum = 0.0;
            for (int64_t t = lane; t < T; t += 32)
            {
                double term, g;
                loss_term<GRAD>(p.loss, p.loss_param, yrow[t], o_s[t], term, g);
                sum += term;
                if (GRAD)
                {
                    p.dL[row * T + t] = g;
                    gb_s[t] += g;
                }
            }
            fsum += loss_post(p.loss, warp_allreduce_sum(sum));
        }
        __syncwarp();
    }

    if (lane == 0)
    {
        f_all[warp] = fsum;
    }
    __syncthreads();

    // fold the warps in warp order
    double* out = p.partials_fb + static_cast<int64_t>(blockIdx.x) * (1 + T);
    if (threadIdx.x == 0)
    {
        double f = 0.0;
        for (int q = 0; q < kGenWarps; ++q)
        {
            f += f_all[q];
        }
        out[0] = f;
    }
    if (GRAD)
    {
        for (int64_t t = threadIdx.x; t < T; t += blockDim.x)
        {
            double s = 0.0;
            for (int q = 0; q < kGenWarps; ++q)
            {
                s += gb_all[q * T + t];
            }
            out[1 + t] = s;
        }
    }
}

// grid: (ceil(D / kBwdCols), ceil(T / kBwdTB), splits). Thread = one column, kBwdTB targets, one row range.
// partials_gw[split][t][d]
__global__ void __launch_bounds__(kBwdCols) gen_backward_kernel(const double* __restrict__ X,
                                                                const double* __restrict__ dL, const int64_t N,
                                                                const int64_t D, const int64_t T,
                                                                double* __restrict__ partials_gw)
{
    const int64_t d      = static_cast<int64_t>(blockIdx.x) * kBwdCols + threadIdx.x;
    const int64_t t0     = static_cast<int64_t>(blockIdx.y) * kBwdTB;
    const int64_t splits = gridDim.z;
    const int64_t chunk  = (N + splits - 1) / splits;
    const int64_t begin  = static_cast<int64_t>(blockIdx.z) * chunk;
    const int64_t end    = min(begin + chunk, N);

    double acc[kBwdTB];
#pragma unroll
    for (int j = 0; j < kBwdTB; ++j)
    {
        acc[j] = 0.0;
    }

    if (d < D)
    {
        for (int64_t n = begin; n < end; ++n)
        {
            const double  xv = X[n * D + d];
            const double* g  = dL + n * T + t0;
#pragma unroll
            for (int j = 0; j < kBwdTB; ++j)
            {
                if (t0 + j < T)
                {
                    acc[j] = fma(g[j], xv, acc[j]);
                }
            }
        }
        double* out = partials_gw + static_cast<int64_t>(blockIdx.z) * T * D;
#pragma unroll
        for (int j = 0; j < kBwdTB; ++j)
        {
            if (t0 + j < T)
            {
                out[(t0 + j) * D + d] = acc[j];
            }
        }
    }
}

// outputs[N, T] = X W^T + b  (linear::predict, src/linear/util.cpp:6-21); one warp per row
__global__ void __launch_bounds__(256) predict_kernel(const double* __restrict__ X, const double* __restrict__ W,
                                                      const double* __restrict__ b, const int64_t N, const int64_t D,
                                                      const int64_t T, double* __restrict__ outputs)
{
    const int     lane        = threadIdx.x & 31;
    const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t warps_total = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = warp_global; row < N; row += warps_total)
    {
        const double* xrow = X + row * D;
        for (int64_t t = 0; t < T; ++t)
        {
            double acc = 0.0;
            for (int64_t c = lane; c < D; c += 32)
            {
                acc = fma(xrow[c], W[t * D + c], acc);
            }
            acc = warp_allreduce_sum(acc);
            if (lane == 0)
            {
                outputs[row * T + t] = acc + b[t];
            }
        }
    }
}

// elementwise loss_t::value / loss_t::vgrad over (N, T) (src/loss.cpp:57-71); one thread per sample
__global__ void __launch_bounds__(256) loss_rows_kernel(const int kind, const double param,
                                                        const double* __restrict__ targets,
                                                        const double* __restrict__ outputs, const int64_t N,
                                                        const int64_t T, double* __restrict__ values,
                                                        double* __restrict__ vgrads)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= N)
    {
        return;
    }
    const double* t = targets + i * T;
    const double* o = outputs + i * T;
    if (kind == LOSS_CLASSNLL)
    {
        double omax = o[0];
        for (int64_t k = 1; k < T; ++k)
        {
            omax = max_of(omax, o[k]);
        }
        double sumexp = 0.0, dot = 0.0;
        for (int64_t k = 0; k < T; ++k)
        {
            sumexp += exp(__dsub_rn(o[k], omax));
            dot += __dmul_rn(__dadd_rn(1.0, t[k]), o[k]);
        }
        if (values != nullptr)
        {
            values[i] = __dadd_rn(__dsub_rn(log(sumexp), __dmul_rn(0.5, dot)), omax);
        }
        if (vgrads != nullptr)
        {
            for (int64_t k = 0; k < T; ++k)
            {
                const double e    = exp(__dsub_rn(o[k], omax));
                vgrads[i * T + k] = __dsub_rn(__ddiv_rn(e, sumexp), __dmul_rn(0.5, __dadd_rn(1.0, t[k])));
            }
        }
    }
    else
    {
        double sum = 0.0;
        for (int64_t k = 0; k < T; ++k)
        {
            double term, g;
            loss_term<true>(kind, param, t[k], o[k], term, g);
            sum += term;
            if (vgrads != nullptr)
            {
                vgrads[i * T + k] = g;
            }
        }
        if (values != nullptr)
        {
            values[i] = loss_post(kind, sum);
        }
    }
}

// elementwise loss_t::error over (N, T) (include/nano/loss/flatten.h:62-67 -> include/nano/loss/error.h:11-70); one
// thread per sample. error_kind: 0 absdiff, 1 sclass, 2 mclass (include/nb200.h: nb200_error)
__global__ void __launch_bounds__(256) loss_error_kernel(const int error_kind, const double* __restrict__ targets,
                                                         const double* __restrict__ outputs, const int64_t N,
                                                         const int64_t T, double* __restrict__ errors)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= N)
    {
        return;
    }
    const double* t     = targets + i * T;
    const double* o     = outputs + i * T;
    double        error = 0.0;
    if (error_kind == 0)
    {
        for (int64_t k = 0; k < T; ++k)
        {
            error += fabs(__dsub_rn(t[k], o[k]));
        }
    }
    else if (error_kind == 1 && T > 1)
    {
        int64_t idx = 0;  // Eigen's maxCoeff keeps the first maximum
        for (int64_t k = 1; k < T; ++k)
        {
            idx = (o[k] > o[idx]) ? k : idx;
        }
        error = (t[idx] > 0.0) ? 0.0 : 1.0;
    }
    else
    {
        for (int64_t k = 0; k < T; ++k)
        {
            error += (__dmul_rn(t[k], o[k]) < 2.220446049250313e-16) ? 1.0 : 0.0;
        }
    }
    errors[i] = error;
}

#endif // __CUDACC__
} // namespace nb200
