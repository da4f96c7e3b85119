// reduce.cuh — the tail of the evaluation: fixed-order reduction of per-CTA partials and the final scaling +
// regularisation.
//
// Replaces sum_reduce (include/nano/core/reduce.h:23-31, accumulator_t::operator+=, operator/=
// src/linear/accumulator.cpp:27-47) and the regularisation block of linear::function_t::do_eval
// (src/linear/function.cpp:127-143,158-167), resp. function_{ridge,lasso,elasticnet}_t::do_eval
// (src/function/mlearn/ridge.cpp:63-80, lasso.cpp:63-76, elasticnet.cpp:67-86).
//
// The reduced vector has the layout shared by every path (and all-reduced across ranks when sample-sharded):
//   [ sum_i loss_i | sum_i dL_i (T) | sum_i dL_i x_i^T (T*D, row-major) ]        un-normalised
#pragma once

#include "common.cuh"
#include "loss.cuh"

namespace nb200
{
#ifdef __CUDACC__

// out[c] = sum_p partials[p * stride + c], p ascending: the summation order depends only on the launch geometry
// of the producing kernel, never on timing (the reference's thread pool does not give that: SURVEY.md §8 a4).
__global__ void __launch_bounds__(128) reduce_partials_kernel(const double* __restrict__ partials, const int rows,
                                                              const int64_t stride, const int64_t cols,
                                                              double* __restrict__ out)
{
    const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= cols)
    {
        return;
    }
    const double* src = partials + c;
    double        sum = 0.0;
    int           p   = 0;
    for (; p + 4 <= rows; p += 4)
    {
        // four independent loads in flight, added in order
        const double a0 = src[static_cast<int64_t>(p) * stride];
        const double a1 = src[static_cast<int64_t>(p + 1) * stride];
        const double a2 = src[static_cast<int64_t>(p + 2) * stride];
        const double a3 = src[static_cast<int64_t>(p + 3) * stride];
        sum += a0;
        sum += a1;
        sum += a2;
        sum += a3;
    }
    for (; p < rows; ++p)
    {
        sum += src[static_cast<int64_t>(p) * stride];
    }
    out[c] = sum;
}

struct finish_params
{
    const double* reduced; // [1 + T + T*D] un-normalised sums (after the cross-rank all-reduce when sharded)
    const double* x;       // parameters on the device: [W | b] (LINEAR) or w (SYNTH)
    double*       gx;      // [n] gradient out (device) or nullptr
    double*       fx;      // 1 double out (device)
    int64_t       D;
    int64_t       T;
    int64_t       n_global; // samples over ALL ranks
    int           flavour;  // 0 LINEAR, 1 SYNTH
    double        l1;
    double        l2;
    // sample-sharded evaluation whose ranks PUBLISHED their sums into this rank's mailbox (vgrad_t1.cuh, peer_wait == 0):
    // mail [world][col_stride] holds one reduced vector per rank; they are folded here in rank order (identical bits on
    // every rank) and the all-reduced vector is also left in reduced_out. mail == nullptr: `reduced` is used as it is.
    const double* mail;
    int           world;
    int64_t       col_stride;
    double*       reduced_out;
};

// entry c of the (all-)reduced vector as seen by the thread that finishes it
__device__ __forceinline__ double finish_reduced_at(const finish_params& p, const int64_t c)
{
    if (p.mail == nullptr)
    {
        return p.reduced[c];
    }
    double sum = 0.0;
    for (int q = 0; q < p.world; ++q)
    {
        sum += __ldcg(p.mail + static_cast<int64_t>(q) * p.col_stride + c);
    }
    p.reduced_out[c] = sum;
    return sum;
}

// block 0 also produces the function value (needs a reduction over W for the regularisers)
__global__ void __launch_bounds__(256) finish_kernel(const finish_params p)
{
    const int64_t wsize   = p.T * p.D;
    const double  samples = static_cast<double>(p.n_global);
    const int64_t k       = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;

    if (p.gx != nullptr)
    {
        if (p.flavour == 0)
        {
            // gw = acc.m_gw / N, then += l1 * sign(w) / wsize, then += l2 * w / wsize (function.cpp:132-142);
            // gb = acc.m_gb / N
            if (k < wsize)
            {
                double g = __ddiv_rn(finish_reduced_at(p, 1 + p.T + k), samples);
                if (p.l1 > 0.0)
                {
                    g = __dadd_rn(g, __ddiv_rn(__dmul_rn(p.l1, sign_of(p.x[k])), static_cast<double>(wsize)));
                }
                if (p.l2 > 0.0)
                {
                    g = __dadd_rn(g, __ddiv_rn(__dmul_rn(p.l2, p.x[k]), static_cast<double>(wsize)));
                }
                p.gx[k] = g;
            }
            else if (k < wsize + p.T)
            {
                p.gx[k] = __ddiv_rn(finish_reduced_at(p, 1 + (k - wsize)), samples);
            }
        }
        else
        {
            // gw = G^T X / N (linear.cpp:66), then += alpha1 * sign(x) + alpha2 * x (elasticnet.cpp:76)
            if (k < wsize)
            {
                const double g   = __ddiv_rn(finish_reduced_at(p, 1 + p.T + k), samples);
                const double reg = __dadd_rn(__dmul_rn(p.l1, sign_of(p.x[k])), __dmul_rn(p.l2, p.x[k]));
                p.gx[k]          = __dadd_rn(g, reg);
            }
        }
    }

    if (blockIdx.x == 0)
    {
        // sum |w| and sum (sqrt(l2) * w)^2 with a fixed-shape tree
        __shared__ double s_abs[256];
        __shared__ double s_sqr[256];
        double            a = 0.0, q = 0.0;
        if (p.l1 > 0.0 || p.l2 > 0.0)
        {
            const double root = sqrt(p.l2);
            for (int64_t j = threadIdx.x; j < wsize; j += blockDim.x)
            {
                const double w = p.x[j];
                a += fabs(w);
                const double sw = __dmul_rn(root, w);
                q               = __dadd_rn(q, __dmul_rn(sw, sw));
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
            double fx = __ddiv_rn(finish_reduced_at(p, 0), samples);
            if (p.flavour == 0)
            {
                // fx += l1 * mean|w| ; fx += 0.5 * mean((sqrt(l2) * w)^2)   (function.cpp:158-166)
                if (p.l1 > 0.0)
                {
                    fx = __dadd_rn(fx, __dmul_rn(p.l1, __ddiv_rn(s_abs[0], static_cast<double>(wsize))));
                }
                if (p.l2 > 0.0)
                {
                    fx = __dadd_rn(fx, __dmul_rn(0.5, __ddiv_rn(s_sqr[0], static_cast<double>(wsize))));
                }
            }
            else
            {
                // fx += alpha1 * |x|_1 + 0.5 * |sqrt(alpha2) * x|^2   (elasticnet.cpp:85)
                fx = __dadd_rn(fx, __dadd_rn(__dmul_rn(p.l1, s_abs[0]), __dmul_rn(0.5, s_sqr[0])));
            }
            *p.fx = fx;
        }
    }
}

#endif // __CUDACC__
} // namespace nb200
