// synth.cuh — synthetic datasets generated directly in HBM (the big BASELINE configs exceed host RAM: SURVEY.md F9).
//
// Counter-based (Philox4x32-10), so any row range of the SAME dataset can be produced independently on any rank:
//   X[i, j]  = 2 * u(stream 0, counter = (i * D + j) / 2, half = (i * D + j) % 2) - 1            in [-1, 1)
//   W*[t, j] = (2 * u(stream 1, counter = t * D + j, half 0) - 1) / sqrt(D)
//   n_i      = Box-Muller normal from (stream 2, counter = i)
// with u = (53 high bits of two 32-bit outputs) * 2^-53. The role of the reference's synthetic generators
// (src/datasource/linear.cpp, capped at 1e6 samples; src/function/mlearn/linear.cpp:15-45) for sizes they cannot reach.
#pragma once

#include "common.cuh"

namespace nb200
{
struct philox4
{
    uint32_t r[4];
};

__host__ __device__ inline philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1)
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int round = 0; round < 10; ++round)
    {
        const uint64_t p0  = static_cast<uint64_t>(M0) * c0;
        const uint64_t p1  = static_cast<uint64_t>(M1) * c2;
        const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
        const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        if (round < 9)
        {
            k0 += W0;
            k1 += W1;
        }
    }
    return philox4{{c0, c1, c2, c3}};
}

__host__ __device__ inline double u53(const uint32_t lo, const uint32_t hi)
{
    const uint64_t bits = (static_cast<uint64_t>(hi) << 32) | lo;
    return static_cast<double>(bits >> 11) * (1.0 / 9007199254740992.0);
}

// uniform in [-1, 1) for global element index e of stream `stream`
__host__ __device__ inline double synth_uniform(const uint64_t seed, const uint32_t stream, const uint64_t e)
{
    const uint64_t c = e >> 1;
    const philox4  q = philox4x32_10(static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 32), 0u, stream,
                                     static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    const double   u = (e & 1) ? u53(q.r[2], q.r[3]) : u53(q.r[0], q.r[1]);
    return 2.0 * u - 1.0; // exact in fp64
}

#ifdef __CUDACC__

// X rows [row_offset, row_offset + N): one thread per PAIR of consecutive elements (one Philox call)
__global__ void __launch_bounds__(256) synth_inputs_kernel(const uint64_t seed, const int64_t N, const int64_t D,
                                                           const int64_t row_offset, double* __restrict__ X)
{
    const uint64_t first  = static_cast<uint64_t>(row_offset) * static_cast<uint64_t>(D);
    const uint64_t total  = static_cast<uint64_t>(N) * static_cast<uint64_t>(D);
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    // pairs are aligned on the GLOBAL element index so that shards agree with the unsharded dataset
    const uint64_t pair0 = first >> 1;
    const uint64_t pairs = ((first + total + 1) >> 1) - pair0;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < pairs; i += stride)
    {
        const uint64_t c = pair0 + i;
        const philox4  q = philox4x32_10(static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 32), 0u, 0u,
                                         static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
        const uint64_t e0 = 2 * c, e1 = 2 * c + 1;
        if (e0 >= first && e0 < first + total)
        {
            X[e0 - first] = 2.0 * u53(q.r[0], q.r[1]) - 1.0;
        }
        if (e1 >= first && e1 < first + total)
        {
            X[e1 - first] = 2.0 * u53(q.r[2], q.r[3]) - 1.0;
        }
    }
}

// W*[T, D] and b*[T] = 0.1
__global__ void __launch_bounds__(256) synth_weights_kernel(const uint64_t seed, const int64_t D, const int64_t T,
                                                            double* __restrict__ W, double* __restrict__ b)
{
    const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k < T * D)
    {
        const philox4 q = philox4x32_10(static_cast<uint32_t>(k), static_cast<uint32_t>(static_cast<uint64_t>(k) >> 32),
                                        0u, 1u, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
        W[k] = __ddiv_rn(2.0 * u53(q.r[0], q.r[1]) - 1.0, sqrt(static_cast<double>(D)));
    }
    if (k < T)
    {
        b[k] = 0.1;
    }
}

// in place: Y holds the clean outputs X W*^T + b* on entry, the targets on exit
// kind 0: y = sign(o + noise * n) in {-1, +1}; kind 1: y = o + noise * n; kind 2: -1 / +1 at the argmax (T >= 2)
__global__ void __launch_bounds__(256) synth_targets_kernel(const uint64_t seed, const int kind, const int64_t N,
                                                            const int64_t T, const int64_t row_offset,
                                                            const double noise, double* __restrict__ Y)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= N)
    {
        return;
    }
    const uint64_t row = static_cast<uint64_t>(row_offset + i);
    double*        y   = Y + i * T;
    if (kind == 2)
    {
        int64_t best = 0;
        for (int64_t t = 1; t < T; ++t)
        {
            if (y[t] > y[best])
            {
                best = t;
            }
        }
        for (int64_t t = 0; t < T; ++t)
        {
            y[t] = (t == best) ? 1.0 : -1.0;
        }
        return;
    }
    const philox4 q  = philox4x32_10(static_cast<uint32_t>(row), static_cast<uint32_t>(row >> 32), 0u, 2u,
                                     static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    const double  u1 = 1.0 - u53(q.r[0], q.r[1]); // (0, 1]
    const double  u2 = u53(q.r[2], q.r[3]);
    const double  n  = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    for (int64_t t = 0; t < T; ++t)
    {
        const double o = y[t] + noise * n;
        y[t]           = (kind == 0) ? ((o < 0.0) ? -1.0 : 1.0) : o;
    }
}

#endif // __CUDACC__
} // namespace nb200
