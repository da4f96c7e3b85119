// probe.cuh — the roofline denominators measured IN the run that quotes them (bench.py): the fp64 ceilings of this
// board (DMMA m8n8k4 issue rate = the tensor pipe the many-target kernels run on; DFMA = the CUDA-core pipe of the
// few-target kernels) and a read-only HBM stream (the single-target pass only reads; the driver's MEASURED_PEAKS.json
// figure is a copy: read + write).
#pragma once

#include "common.cuh"

namespace nb200
{
constexpr int kProbeIters = 4096;

#ifdef __CUDACC__

// 8 independent m8n8k4 accumulator chains per warp: 8 * 2 * 8*8*4 flops per iteration and warp
__global__ void __launch_bounds__(256) probe_dmma_kernel(double* out, const double a, const double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        c[i][0] = threadIdx.x;
        c[i][1] = i;
    }
    for (int it = 0; it < kProbeIters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        s += c[i][0] + c[i][1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 16 independent FMA chains per thread: 16 * 2 flops per iteration and thread
__global__ void __launch_bounds__(256) probe_dfma_kernel(double* out, const double a, const double b)
{
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        acc[i] = threadIdx.x + i;
    }
    for (int it = 0; it < kProbeIters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 16; ++i)
        {
            acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        s += acc[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// read-only stream: every CTA pulls 64 KiB tiles through a 2-deep ring of 1-D bulk copies (the way the single-target
// kernel reads X) and touches one word per tile, so the bytes really arrive
__global__ void __launch_bounds__(128) probe_read_kernel(const unsigned char* __restrict__ src, const int64_t tiles,
                                                         double* out)
{
    constexpr int kTile = 65536, kStages = 2;
    extern __shared__ __align__(1024) unsigned char probe_smem[];
    uint64_t*      bars   = reinterpret_cast<uint64_t*>(probe_smem);
    unsigned char* stages = probe_smem + 128;
    if (threadIdx.x == 0)
    {
        for (int s = 0; s < kStages; ++s)
        {
            mbar_init(bars + s, 1);
        }
        mbar_fence_init();
    }
    __syncthreads();
    double sum = 0.0;
    if (threadIdx.x == 0)
    {
        const uint64_t policy = l2_policy_evict_first();
        int64_t        next   = blockIdx.x;
        for (int s = 0; s < kStages && next < tiles; ++s, next += gridDim.x)
        {
            mbar_arrive_expect_tx(bars + s, kTile);
            bulk_copy_g2s(stages + s * kTile, src + next * kTile, kTile, bars + s, policy);
        }
        int      s     = 0;
        uint32_t phase = 0;
        for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x)
        {
            mbar_wait(bars + s, phase);
            sum += *reinterpret_cast<const double*>(stages + s * kTile + (tile & 31) * 8);
            if (next < tiles)
            {
                // the load above before the refill below (generic proxy -> async proxy, see mbar_release_stage)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive_expect_tx(bars + s, kTile);
                bulk_copy_g2s(stages + s * kTile, src + next * kTile, kTile, bars + s, policy);
                next += gridDim.x;
            }
            if (++s == kStages)
            {
                s = 0;
                phase ^= 1u;
            }
        }
        out[blockIdx.x] = sum;
    }
}

#endif // __CUDACC__
} // namespace nb200
