// fp64_peaks.cu — measures the fp64 ceilings of this B200 that SURVEY.md §8d asks for before the many-class (T = 100)
// kernels are designed: DFMA issue rate, DMMA (mma.sync f64) rate for the m8n8k4 and m16n8k16 shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_peaks tools/fp64_peaks.cu && /tmp/fp64_peaks
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

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

// NVML through dlopen: board power and SM clock while a kernel runs
struct nvml_t
{
    void* dev{nullptr};
    int (*power)(void*, unsigned*){nullptr};
    int (*clock)(void*, int, unsigned*){nullptr};
    bool open()
    {
        void* lib = dlopen("libnvidia-ml.so.1", RTLD_NOW);
        if (!lib) return false;
        auto init   = (int (*)())dlsym(lib, "nvmlInit_v2");
        auto handle = (int (*)(unsigned, void**))dlsym(lib, "nvmlDeviceGetHandleByIndex_v2");
        power       = (int (*)(void*, unsigned*))dlsym(lib, "nvmlDeviceGetPowerUsage");
        clock       = (int (*)(void*, int, unsigned*))dlsym(lib, "nvmlDeviceGetClockInfo");
        return init && handle && power && clock && init() == 0 && handle(0, &dev) == 0;
    }
};

// back-to-back launches for `seconds`: the steady state (second half) is what a 100 ms contraction actually gets
template <class K>
void sustained(const char* name, K kernel, int blocks, int threads, double* out, double flops_per_launch,
               double seconds, nvml_t& nvml)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    std::vector<float>    times;
    std::vector<unsigned> watts, mhz;
    const auto start = std::chrono::steady_clock::now();
    while (std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count() < seconds)
    {
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) kernel<<<blocks, threads>>>(out, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        unsigned mw = 0, clk = 0;
        if (nvml.dev)
        {
            nvml.power(nvml.dev, &mw);
            nvml.clock(nvml.dev, 1, &clk);
        }
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        times.push_back(ms / 20);
        watts.push_back(mw);
        mhz.push_back(clk);
    }
    const size_t h = times.size() / 2;
    double t = 0, w = 0, c = 0;
    for (size_t i = h; i < times.size(); ++i) { t += times[i]; w += watts[i]; c += mhz[i]; }
    const double n = double(times.size() - h);
    printf("{\"sustained\": \"%s\", \"seconds\": %.1f, \"first_tflops\": %.2f, \"steady_tflops\": %.2f, "
           "\"steady_sm_mhz\": %.0f, \"steady_power_w\": %.0f}\n",
           name, seconds, flops_per_launch / (times[0] * 1e-3) / 1e12, flops_per_launch / (t / n * 1e-3) / 1e12, c / n,
           w / n / 1e3);
    fflush(stdout);
    cudaDeviceSynchronize();
    const auto c0 = std::chrono::steady_clock::now();
    while (std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count() < 1.5) {}
}

int main(int argc, char** argv)
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
    if (argc > 1 && std::strcmp(argv[1], "--sustained") == 0)
    {
        nvml_t nvml;
        if (!nvml.open()) fprintf(stderr, "NVML unavailable\n");
        const int    threads = 256, blocks = sms * 2;
        const double total   = static_cast<double>(blocks) * threads;
        sustained("dfma", dfma_kernel, blocks, threads, out, 2.0 * 16 * kIters * total, 3.0, nvml);
        sustained("dmma_m8n8k4", dmma884_kernel, blocks, threads, out, 2.0 * 8 * 8 * 8 * 4 * kIters * (total / 32), 3.0,
                  nvml);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
    {
        printf("cuda error: %s\n", cudaGetErrorString(err));
        return 1;
    }
    return 0;
}
