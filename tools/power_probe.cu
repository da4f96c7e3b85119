// power_probe.cu — why does the fused single-target kernel hit the 1000 W board power cap, and what would a leaner
// consumer loop buy? Streams a [N, 1024] fp64 matrix through the same producer ring as vgrad_t1_kernel (1-D bulk
// copies + mbarriers, 3 x 64 KiB stages, one persistent CTA per SM) with consumers of increasing work:
//   0 stream   : wait for the stage, release it (data movement only: HBM -> L2 -> shared memory)
//   1 lds      : + read the row into registers (128-bit LDS)
//   2 dot      : + dot with w (shared memory) + butterfly
//   3 dot+grad : + g += dl * x  (the arithmetic of the real kernel without the loss)
//   4 dot+grad, w in registers (no second LDS stream)
// each with a spinning or a backing-off (nanosleep) barrier wait. Every variant runs back to back for --seconds and
// reports the steady-state time per launch (second half of the run), SM clock and board power (NVML).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/power_probe tools/power_probe.cu -ldl
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x)                                                                                                       \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (x);                                                                                          \
        if (e_ != cudaSuccess)                                                                                         \
        {                                                                                                              \
            std::fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e_));                                       \
            std::exit(1);                                                                                              \
        }                                                                                                              \
    } while (0)

constexpr int D      = 1024;
constexpr int RPS    = 8; // rows per stage = consumer warps
constexpr int STAGES = 3;
constexpr int NCW    = 8;
constexpr int STAGE_BYTES = RPS * D * 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t tx)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(tx) : "memory");
}
template <int WAIT>
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity)
{
    uint32_t done = 0;
    while (true)
    {
        if constexpr (WAIT == 2)
        {
            // hardware-suspended wait with a long time hint (ns)
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(smem_u32(b)), "r"(parity), "r"(1000000u)
                         : "memory");
        }
        else
        {
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(smem_u32(b)), "r"(parity)
                         : "memory");
        }
        if (done) break;
        if constexpr (WAIT == 1) __nanosleep(200);
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, "
                 "[%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int WORK, int WAIT>
__global__ void __launch_bounds__((NCW + 1) * 32, 1)
    probe_kernel(const double* __restrict__ X, const double* __restrict__ w, double* __restrict__ out, int64_t num_tiles)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full  = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + STAGES;
    double*   w_s   = reinterpret_cast<double*>(smem + 1024);
    unsigned char* stages = smem + 1024 + D * 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0)
    {
        for (int s = 0; s < STAGES; ++s)
        {
            mbar_init(full + s, 1);
            mbar_init(empty + s, NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int j = threadIdx.x; j < D; j += blockDim.x) w_s[j] = w[j];
    __syncthreads();

    if (warp == NCW)
    {
        if (lane == 0)
        {
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            int s = 0;
            uint32_t phase = 1;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
            {
                mbar_wait<WAIT>(empty + s, phase);
                mbar_expect(full + s, STAGE_BYTES);
                bulk_g2s(stages + s * STAGE_BYTES, X + tile * RPS * D, STAGE_BYTES, full + s, policy);
                if (++s == STAGES) { s = 0; phase ^= 1; }
            }
        }
        return;
    }

    double g[32];
    double wr[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) g[k] = 0.0;
    if constexpr (WORK == 4)
    {
#pragma unroll
        for (int k = 0; k < 16; ++k)
        {
            const double2 wv = *reinterpret_cast<const double2*>(w_s + (k * 32 + lane) * 2);
            wr[2 * k] = wv.x; wr[2 * k + 1] = wv.y;
        }
    }
    double fsum = 0.0;
    int s = 0;
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
    {
        mbar_wait<WAIT>(full + s, phase);
        const double* row = reinterpret_cast<const double*>(stages + s * STAGE_BYTES) + warp * D;
        if constexpr (WORK == 0)
        {
            if (lane == 0) mbar_arrive(empty + s);
        }
        else
        {
            double x[32];
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int k = 0; k < 16; ++k)
            {
                const double2 xv = *reinterpret_cast<const double2*>(row + (k * 32 + lane) * 2);
                x[2 * k] = xv.x; x[2 * k + 1] = xv.y;
                if constexpr (WORK == 2 || WORK == 3)
                {
                    const double2 wv = *reinterpret_cast<const double2*>(w_s + (k * 32 + lane) * 2);
                    a0 = fma(xv.x, wv.x, a0); a1 = fma(xv.y, wv.y, a1);
                }
                if constexpr (WORK == 4)
                {
                    a0 = fma(xv.x, wr[2 * k], a0); a1 = fma(xv.y, wr[2 * k + 1], a1);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
            if constexpr (WORK == 1)
            {
#pragma unroll
                for (int k = 0; k < 32; ++k) fsum += x[k]; // sink (DADD, keeps the loads alive)
            }
            else
            {
                const double dot = warp_sum(a0 + a1);
                fsum += dot;
                if constexpr (WORK >= 3)
                {
                    const double dl = dot * 1e-3;
#pragma unroll
                    for (int k = 0; k < 32; ++k) g[k] = fma(dl, x[k], g[k]);
                }
            }
        }
        if (++s == STAGES) { s = 0; phase ^= 1; }
    }
    double total = fsum;
#pragma unroll
    for (int k = 0; k < 32; ++k) total += g[k];
    if (total == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = total; // never true; keeps the work alive
}

// ---- NVML through dlopen -------------------------------------------------------------------------------------------
struct nvml_t
{
    void* lib{nullptr};
    void* dev{nullptr};
    int (*power)(void*, unsigned*){nullptr};
    int (*clock)(void*, int, unsigned*){nullptr};
    int (*temp)(void*, int, unsigned*){nullptr};
    bool open()
    {
        lib = dlopen("libnvidia-ml.so.1", RTLD_NOW);
        if (!lib) return false;
        auto init   = (int (*)())dlsym(lib, "nvmlInit_v2");
        auto handle = (int (*)(unsigned, void**))dlsym(lib, "nvmlDeviceGetHandleByIndex_v2");
        power       = (int (*)(void*, unsigned*))dlsym(lib, "nvmlDeviceGetPowerUsage");
        clock       = (int (*)(void*, int, unsigned*))dlsym(lib, "nvmlDeviceGetClockInfo");
        temp        = (int (*)(void*, int, unsigned*))dlsym(lib, "nvmlDeviceGetTemperature");
        if (!init || !handle || !power || !clock) return false;
        if (init() != 0) return false;
        return handle(0, &dev) == 0;
    }
};

template <int WORK, int WAIT>
void run(const char* name, const double* X, const double* w, double* out, int64_t rows, double seconds, nvml_t& nvml,
         int sms)
{
    const int smem = 1024 + D * 8 + STAGES * STAGE_BYTES;
    CHECK(cudaFuncSetAttribute(probe_kernel<WORK, WAIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t tiles = rows / RPS;
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    const auto start = std::chrono::steady_clock::now();
    std::vector<float> times;
    std::vector<unsigned> powers, clocks;
    double elapsed = 0.0;
    while (elapsed < seconds)
    {
        CHECK(cudaEventRecord(e0));
        for (int i = 0; i < 20; ++i) probe_kernel<WORK, WAIT><<<sms, (NCW + 1) * 32, smem>>>(X, w, out, tiles);
        CHECK(cudaEventRecord(e1));
        // sample while the batch runs
        unsigned mw = 0, mhz = 0;
        if (nvml.dev)
        {
            nvml.power(nvml.dev, &mw);
            nvml.clock(nvml.dev, 1 /*NVML_CLOCK_SM*/, &mhz);
        }
        CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        times.push_back(ms / 20);
        powers.push_back(mw);
        clocks.push_back(mhz);
        elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    }
    // steady state = second half
    const size_t h = times.size() / 2;
    double t = 0, pw = 0, ck = 0;
    for (size_t i = h; i < times.size(); ++i) { t += times[i]; pw += powers[i]; ck += clocks[i]; }
    const double n = double(times.size() - h);
    const double us = 1e3 * t / n;
    unsigned tg = 0;
    if (nvml.dev && nvml.temp) nvml.temp(nvml.dev, 0, &tg);
    std::printf("{\"variant\": \"%s\", \"wait\": %d, \"first_us\": %.1f, \"steady_us\": %.1f, \"steady_gbs\": %.1f, "
                "\"sm_mhz\": %.0f, \"power_w\": %.0f, \"temp_c\": %u, \"batches\": %zu}\n",
                name, WAIT, 1e3 * times[0], us, rows * double(D) * 8 / us / 1e3, ck / n, pw / n / 1e3, tg, times.size());
    std::fflush(stdout);
    CHECK(cudaDeviceSynchronize());
    // cool-down so that every variant starts from a similar state
    {
        const auto c0 = std::chrono::steady_clock::now();
        while (std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count() < 1.5) {}
    }
}

__global__ void fill(double* X, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        X[i] = double((i * 2654435761ULL) & 0xffff) / 65536.0 - 0.5;
}

int main(int argc, char** argv)
{
    const double  seconds = argc > 1 ? std::atof(argv[1]) : 3.0;
    const int64_t rows    = 1 << 20;
    int sms = 0;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double *X, *w, *out;
    CHECK(cudaMalloc(&X, rows * D * 8 + 256));
    CHECK(cudaMalloc(&w, D * 8));
    CHECK(cudaMalloc(&out, 148 * 512 * 8));
    fill<<<1184, 256>>>(X, rows * D);
    fill<<<4, 256>>>(w, D);
    CHECK(cudaDeviceSynchronize());
    nvml_t nvml;
    if (!nvml.open()) std::fprintf(stderr, "NVML unavailable: power / clocks reported as 0\n");

    run<0, 0>("stream", X, w, out, rows, seconds, nvml, sms);
    run<0, 1>("stream", X, w, out, rows, seconds, nvml, sms);
    run<0, 2>("stream", X, w, out, rows, seconds, nvml, sms);
    run<1, 0>("lds", X, w, out, rows, seconds, nvml, sms);
    run<2, 0>("dot", X, w, out, rows, seconds, nvml, sms);
    run<3, 0>("dot+grad", X, w, out, rows, seconds, nvml, sms);
    run<3, 1>("dot+grad", X, w, out, rows, seconds, nvml, sms);
    run<3, 2>("dot+grad", X, w, out, rows, seconds, nvml, sms);
    run<4, 0>("dot+grad w-in-regs", X, w, out, rows, seconds, nvml, sms);
    run<4, 2>("dot+grad w-in-regs", X, w, out, rows, seconds, nvml, sms);
    return 0;
}
