// batch.cuh — SURVEY.md §8f N3: many parameter vectors in ONE pass over X.
//
// libnano's tuner evaluates a grid of regularisation values (31 per parameter: src/linear/util.cpp:75-86) times the
// folds of a cross-validation (src/machine/tune.cpp:25-42), every trial re-reading the same X. K single-target models
// that share (X, y, loss) and differ in (x_k, l1_k, l2_k) are evaluated here as ONE many-"target" problem on the
// FP64 tensor pipe (vgrad_mt.cuh): outputs[N, K] = X [w_1 .. w_K] + b, the loss of column k against the same y,
// gW_k = dL[:, k]^T X. The T = 1 GEMV becomes a bandwidth-amortised GEMM: two reads of X for K evaluations.
#pragma once

#include "common.cuh"
#include "loss.cuh"

namespace nb200
{
#ifdef __CUDACC__

// x_all [K][n] (n = D + 1 for LINEAR: [w | b]; n = D for SYNTH with the fixed bias) -> W [Kp][D], b [Kp]
// (rows K .. Kp-1 repeat model K-1: padding to an even count)
__global__ void __launch_bounds__(256) batch_pack_kernel(const int K, const int Kp, const int64_t D, const int64_t n,
                                                         const int flavour, const double* __restrict__ x_all,
                                                         const double* __restrict__ fixed_bias,
                                                         double* __restrict__ W, double* __restrict__ b)
{
    const int     k   = blockIdx.y;
    const int     src = min(k, K - 1);
    const double* x   = x_all + static_cast<int64_t>(src) * n;
    for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < D;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x)
    {
        W[static_cast<int64_t>(k) * D + j] = x[j];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        b[k] = (flavour == 0) ? x[D] : fixed_bias[0];
    }
}

// one block per model: the scaling + regularisation of finish_kernel (reduce.cuh) with per-model (l1, l2)
// red_fb = [total | dL sums (Kp) | loss sums (Kp)], red_gw = [Kp][D]; out [K][n + 1] = gradient, then the value
__global__ void __launch_bounds__(256) batch_finish_kernel(const int Kp, const int64_t D, const int64_t n,
                                                           const int64_t samples_i, const int flavour,
                                                           const double* __restrict__ red_fb,
                                                           const double* __restrict__ red_gw,
                                                           const double* __restrict__ x_all,
                                                           const double* __restrict__ l1s, const double* __restrict__ l2s,
                                                           const int want_grad, double* __restrict__ out)
{
    __shared__ double s_abs[256];
    __shared__ double s_sqr[256];
    const int     k       = blockIdx.x;
    const double* x       = x_all + static_cast<int64_t>(k) * n;
    double*       o       = out + static_cast<int64_t>(k) * (n + 1);
    const double  l1      = l1s[k];
    const double  l2      = l2s[k];
    const double  samples = static_cast<double>(samples_i);
    const double  wsize   = static_cast<double>(D);
    const double  root    = sqrt(l2);

    double a = 0.0, q = 0.0;
    for (int64_t j = threadIdx.x; j < D; j += blockDim.x)
    {
        const double w = x[j];
        if (l1 > 0.0 || l2 > 0.0)
        {
            a += fabs(w);
            const double sw = __dmul_rn(root, w);
            q               = __dadd_rn(q, __dmul_rn(sw, sw));
        }
        if (want_grad)
        {
            double g = __ddiv_rn(red_gw[static_cast<int64_t>(k) * D + j], samples);
            if (flavour == 0)
            {
                if (l1 > 0.0)
                {
                    g = __dadd_rn(g, __ddiv_rn(__dmul_rn(l1, sign_of(w)), wsize));
                }
                if (l2 > 0.0)
                {
                    g = __dadd_rn(g, __ddiv_rn(__dmul_rn(l2, w), wsize));
                }
            }
            else
            {
                g = __dadd_rn(g, __dadd_rn(__dmul_rn(l1, sign_of(w)), __dmul_rn(l2, w)));
            }
            o[j] = g;
        }
    }
    s_abs[threadIdx.x] = a;
    s_sqr[threadIdx.x] = q;
    __syncthreads();
    for (int half = 128; half > 0; half >>= 1)
    {
        if (threadIdx.x < half)
        {
            s_abs[threadIdx.x] += s_abs[threadIdx.x + half];
            s_sqr[threadIdx.x] += s_sqr[threadIdx.x + half];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        if (want_grad && flavour == 0)
        {
            o[D] = __ddiv_rn(red_fb[1 + k], samples); // gb
        }
        double fx = __ddiv_rn(red_fb[1 + Kp + k], samples);
        if (flavour == 0)
        {
            if (l1 > 0.0)
            {
                fx = __dadd_rn(fx, __dmul_rn(l1, __ddiv_rn(s_abs[0], wsize)));
            }
            if (l2 > 0.0)
            {
                fx = __dadd_rn(fx, __dmul_rn(0.5, __ddiv_rn(s_sqr[0], wsize)));
            }
        }
        else
        {
            fx = __dadd_rn(fx, __dadd_rn(__dmul_rn(l1, s_abs[0]), __dmul_rn(0.5, s_sqr[0])));
        }
        o[n] = fx;
    }
}

#endif // __CUDACC__
} // namespace nb200
