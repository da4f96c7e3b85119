// fp64_peaks.cu — measures the fp64 ceilings of this B200 that SURVEY.md §8d asks for before the many-class (T = 100)
// kernels are designed: DFMA issue rate, DMMA (mma.sync f64) rate for the m8n8k4 and m16n8k16 shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_peaks tools/fp64_peaks.cu && /tmp/fp64_peaks
#include <cuda_runtime.h>

#include <cstdio>

constexpr int kIters = 4096;

__global__ void dfma_kernel(double* out, double a, double b)
{
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
    {
        acc[i] = threadIdx.x + i;
    }
    for (int it = 0; it < kIters; ++it)
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

__global__ void dmma884_kernel(double* out, double a, double b)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
        c[i][0] = threadIdx.x;
        c[i][1] = i;
    }
    for (int it = 0; it < kIters; ++it)
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

__global__ void dmma16816_kernel(double* out, double a, double b)
{
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            c[i][j] = threadIdx.x + i + j;
        }
    }
    for (int it = 0; it < kIters; ++it)
    {
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                         "{%4,%4,%4,%4,%4,%4,%4,%4}, {%5,%5,%5,%5}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
double time_kernel(K kernel, int blocks, int threads, double* out)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kernel<<<blocks, threads>>>(out, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep)
    {
        cudaEventRecord(e0);
        kernel<<<blocks, threads>>>(out, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best * 1e-3;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double*   out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    for (int warps : {4, 8, 16, 32})
    {
        const int    threads = warps * 32;
        const int    blocks  = sms * 2;
        const double total   = static_cast<double>(blocks) * threads;
        const double t_fma   = time_kernel(dfma_kernel, blocks, threads, out);
        const double t_884   = time_kernel(dmma884_kernel, blocks, threads, out);
        const double t_16816 = time_kernel(dmma16816_kernel, blocks, threads, out);
        printf("{\"warps_per_cta\": %d, \"ctas\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, "
               "\"dmma_m16n8k16_tflops\": %.2f}\n",
               warps, blocks, 2.0 * 16 * kIters * total / t_fma / 1e12,
               2.0 * 8 * 8 * 8 * 4 * kIters * (total / 32) / t_884 / 1e12,
               2.0 * 4 * 16 * 8 * 16 * kIters * (total / 32) / t_16816 / 1e12);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
    {
        printf("cuda error: %s\n", cudaGetErrorString(err));
        return 1;
    }
    return 0;
}
