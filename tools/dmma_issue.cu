// dmma_issue.cu — how close can mma.sync.m8n8k4.f64 get to its issue peak with the warp count and operand feed of
// the many-target kernels (vgrad_mt.cuh)? Sweeps consumer warps per SM and compares register-only operands with
// operands re-loaded from shared memory per k-step the way the forward kernel does (2 A + NT8 B fragment loads per
// 2 * NT8 DMMAs).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_issue tools/dmma_issue.cu
#include <cuda_runtime.h>

#include <cstdio>

constexpr int kIters = 2048;

__device__ __forceinline__ void dmma(double& c0, double& c1, const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int NT8, bool SMEM>
__global__ void kernel(double* out, const double seed)
{
    extern __shared__ double tile[]; // [128 + NT8 * 8][20]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tig = lane & 3;
    for (int i = threadIdx.x; i < (128 + NT8 * 8) * 20; i += blockDim.x)
    {
        tile[i] = seed + 1e-9 * i;
    }
    __syncthreads();
    double acc[2][NT8][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < NT8; ++n)
        {
            acc[m][n][0] = lane;
            acc[m][n][1] = n;
        }
    const double* xa = tile + ((warp % 8) * 16 + g) * 20 + tig;
    const double* wb = tile + 128 * 20 + g * 20 + tig;
    double        a0 = seed, a1 = seed * 2, bv = seed * 3;
    for (int it = 0; it < kIters; ++it)
    {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
        {
            if (SMEM)
            {
                a0 = xa[kk * 4];
                a1 = xa[8 * 20 + kk * 4];
            }
#pragma unroll
            for (int n = 0; n < NT8; ++n)
            {
                if (SMEM)
                {
                    bv = wb[n * 8 * 20 + kk * 4];
                }
                dmma(acc[0][n][0], acc[0][n][1], a0, bv);
                dmma(acc[1][n][0], acc[1][n][1], a1, bv);
            }
        }
        if (SMEM && it == kIters - 1)
        {
            tile[threadIdx.x] = acc[0][0][0]; // keep the loads alive across iterations
        }
    }
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < NT8; ++n)
        {
            s += acc[m][n][0] + acc[m][n][1];
        }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT8, bool SMEM>
void run(const int sms, const int warps, double* out)
{
    const int smem = (128 + NT8 * 8) * 20 * 8;
    cudaFuncSetAttribute(kernel<NT8, SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep)
    {
        cudaEventRecord(e0);
        kernel<NT8, SMEM><<<sms, warps * 32, smem>>>(out, 1.0000001);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best)
        {
            best = ms;
        }
    }
    const double flops = static_cast<double>(sms) * warps * kIters * 4.0 * 2 * NT8 * 512.0;
    std::printf("{\"nt8\": %d, \"operands\": \"%s\", \"warps_per_sm\": %d, \"tflops\": %.2f, \"error\": \"%s\"}\n", NT8,
                SMEM ? "shared" : "registers", warps, flops / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, 0);
    double* out = nullptr;
    cudaMalloc(&out, sizeof(double) * prop.multiProcessorCount * 1024);
    for (const int warps : {4, 8, 12, 16})
    {
        run<13, false>(prop.multiProcessorCount, warps, out);
        run<13, true>(prop.multiProcessorCount, warps, out);
    }
    for (const int warps : {8, 16})
    {
        run<4, false>(prop.multiProcessorCount, warps, out);
        run<4, true>(prop.multiProcessorCount, warps, out);
        run<2, true>(prop.multiProcessorCount, warps, out);
    }
    return 0;
}
