// common.cuh — error handling and the sm_100a PTX building blocks (mbarrier, 1-D bulk async copy, warp reductions)
// shared by the kernels of libnb200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

namespace nb200
{
// ---------------------------------------------------------------------------------------------------------------
// host-side error plumbing: no exception crosses the C ABI; the message is kept per thread.
// ---------------------------------------------------------------------------------------------------------------
inline thread_local std::string g_last_error;

inline int fail(const int status, const std::string& message)
{
    g_last_error = message;
    return status;
}

#define NB200_CUDA_TRY(expr)                                                                                           \
    do                                                                                                                 \
    {                                                                                                                  \
        const cudaError_t nb200_err_ = (expr);                                                                         \
        if (nb200_err_ != cudaSuccess)                                                                                 \
        {                                                                                                              \
            return ::nb200::fail(2 /*NB200_ERR_CUDA*/, std::string("CUDA error in ") + #expr + ": " +                   \
                                                           cudaGetErrorName(nb200_err_) + " (" +                       \
                                                           cudaGetErrorString(nb200_err_) + ")");                      \
        }                                                                                                              \
    } while (0)

#define NB200_REQUIRE(cond, message)                                                                                   \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(cond))                                                                                                   \
        {                                                                                                              \
            return ::nb200::fail(1 /*NB200_ERR_INVALID*/, std::string("critical check failed: ") + (message));          \
        }                                                                                                              \
    } while (0)

inline int64_t round_up(const int64_t value, const int64_t multiple)
{
    return (value + multiple - 1) / multiple * multiple;
}

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* ptr)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(ptr));
}

// mbarrier (shared::cta, 64-bit): init / arrive / arrive+expect_tx / parity wait
__device__ __forceinline__ void mbar_init(uint64_t* bar, const uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init()
{
    // make the initialised barriers visible to the async proxy (the bulk-copy engine)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// A consumer warp hands a stage it has READ with ordinary shared-memory loads (generic proxy) back to the producer, whose
// refill is a bulk / tensor-map copy (async proxy). The cross-proxy fence orders every lane's loads before the refill
// that follows the arrival. Without it the arrival alone was measured NOT to be enough: in the many-target forward a
// refill overtook loads still in flight about once per 10^6 stage uses (one warp's 16 rows of one tile wrong, different
// rows on every run). Called by all 32 lanes; one arrival per warp.
__device__ __forceinline__ void mbar_release_stage(uint64_t* empty_bar, const int lane)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0)
    {
        mbar_arrive(empty_bar);
    }
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, const uint32_t tx_bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx_bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, const uint32_t parity)
{
    // try_wait suspends the thread in hardware up to a time limit; loop until the phase with `parity` completed
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "NB200_WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra NB200_DONE_%=;\n\t"
                 "bra NB200_WAIT_%=;\n\t"
                 "NB200_DONE_%=:\n\t"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}

// L2 eviction policy for data that is streamed exactly once (X tiles)
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}

// ... and for a dataset small enough to stay in the 126 MB L2 between evaluations (the solver re-reads X every call)
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}

// 1-D bulk asynchronous copy global -> shared (the TMA engine without a tensor map; SASS: UBLKCP).
// dst and src 16-byte aligned, bytes a multiple of 16; completion is signalled on `bar` as transaction bytes.
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, const uint32_t bytes,
                                              uint64_t* bar, const uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, "
                 "[%3], %4;" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

// 16-byte asynchronous copy global -> shared through the LSU (SASS LDGSTS): the right tool when a tile row is only
// 128 bytes long, where one bulk-copy request per row would be bound by the request rate of the copy engine
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

// the executing thread arrives on `bar` once all of its earlier cp.async copies have landed (the arrival must be
// part of the barrier's expected count: .noinc)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// named barrier over a subset of the CTA's warps
__device__ __forceinline__ void named_bar_sync(const uint32_t id, const uint32_t threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// butterfly all-reduce: every lane ends with the same bits (fp add is commutative, so both partners of each
// exchange compute the identical sum)
__device__ __forceinline__ double warp_allreduce_sum(double value)
{
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1)
    {
        value += __shfl_xor_sync(0xffffffffu, value, offset);
    }
    return value;
}

__host__ __device__ constexpr int log2_of(const int value)
{
    return value <= 1 ? 0 : 1 + log2_of(value / 2);
}

// Sum R per-lane values ACROSS the warp at once (R a power of two <= 32): recursive halving — at every step a lane
// keeps one half of its values and ships the other half to its partner — takes R - 1 exchanges plus a plain
// butterfly over the remaining log2(32 / R) offsets, against 5 * R exchanges for R separate butterflies. Returns
// the total of row (lane >> (5 - log2 R)); every lane of that group holds the same bits (both partners of an
// exchange add the same two numbers). values[] is clobbered.
template <int R>
__device__ __forceinline__ double warp_reduce_rows_from(double (&values)[R], const int lane, int offset)
{
#pragma unroll
    for (int n = R; n > 1; n /= 2)
    {
        const bool upper = (lane & offset) != 0;
#pragma unroll
        for (int j = 0; j < n / 2; ++j)
        {
            const double keep = upper ? values[j + n / 2] : values[j];
            const double send = upper ? values[j] : values[j + n / 2];
            values[j]         = keep + __shfl_xor_sync(0xffffffffu, send, offset);
        }
        offset >>= 1;
    }
    double total = values[0];
#pragma unroll
    for (; offset > 0; offset >>= 1)
    {
        total += __shfl_xor_sync(0xffffffffu, total, offset);
    }
    return total;
}

template <int R>
__device__ __forceinline__ double warp_reduce_rows(double (&values)[R], const int lane)
{
    return warp_reduce_rows_from<R>(values, lane, 16);
}

__device__ __forceinline__ double warp_allreduce_max(double value)
{
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1)
    {
        const double other = __shfl_xor_sync(0xffffffffu, value, offset);
        value              = (value < other) ? other : value;
    }
    return value;
}

#endif // __CUDACC__
} // namespace nb200
