// vgrad_mt.cuh — the many-target path (softmax / multi-label models, BASELINE config 3: 4M x 2048 x 100): the two
// dense contractions of linear::function_t::do_eval on the FP64 tensor pipe (mma.sync m8n8k4 f64, SASS DMMA.8x8x4).
//
//   mt_forward_kernel   O = X W^T + b  (src/linear/util.cpp:19-20), loss value + dL in the epilogue straight from the
//                       accumulator fragments (include/nano/loss/flatten.h:69-86): O is never written to HBM
//   mt_backward_kernel  gW = dL^T X    (src/linear/function.cpp:70) as a split-K GEMM over row ranges
//
// This case is contraction-bound (50 flop per byte of X at T = 100), so X is read twice (once per GEMM). tcgen05 has no
// fp64 kind; DMMA through mma.sync is the fp64 tensor path of sm_100a. Both kernels: persistent CTAs, a producer
// warpgroup feeding a ring of shared-memory stages behind mbarriers (padded row stride so that the 64-bit fragment
// loads are bank-conflict free), 8 consumer warps each owning 2 x NT8 (resp. NT8 x 2) m8n8 tiles. The backward tile
// rows are >= 800 bytes and travel as 1-D bulk copies; the forward tile rows are 128 bytes, where one bulk request per
// row is bound by the copy engine's request rate (measured: 9 TFLOP/s), so they travel as 16-byte cp.async chunks.
#pragma once

#include <cuda.h> // CUtensorMap (types only: the encoder is fetched from the driver at run time)

#include "common.cuh"
#include "loss.cuh"

namespace nb200
{
constexpr int kMtWarps    = 8;   // consumer warps
constexpr int kMtBM       = 128; // forward: rows per CTA tile (16 per warp)
constexpr int kMtBK       = 16;  // forward: features per stage
constexpr int kMtLd       = kMtBK + 4; // forward: shared row stride in doubles (8*row + 2*k mod 32 distinct per half-warp)
constexpr int kMtBwdCols  = 128; // backward: columns per CTA (2 n8 tiles per warp)
constexpr int kMtBwdRows  = 16;  // backward: rows per stage
constexpr int kMtBwdLd    = kMtBwdCols + 4; // backward: shared row stride of the X tile
constexpr int kMtMaxStages = 8;
constexpr int kMtProducerWarps = 4;   // a whole warpgroup, so that its registers can be handed to the consumers
constexpr int kMtThreads       = (kMtWarps + kMtProducerWarps) * 32;
constexpr int kMtProducerRegs  = 40;  // 384 threads x 168 registers at launch -> 128 x 40 + 256 x 232
constexpr int kMtConsumerRegs  = 232;
constexpr int kMtEpilogueRows     = 4;   // rows a warp evaluates together in the loss epilogue (independent fp64 chains)
constexpr int kMtEpilogueRowsTma  = 8;   // ... in the tensor-map-fed forward, whose accumulators are all dead by then
constexpr int kMtTmaHeader        = 1024; // barriers
constexpr int kMtXTileBytes       = kMtBM * kMtBK * 8; // 128 rows x 128 bytes, dense, 128-byte swizzled

struct mt_forward_params
{
    const double* X;  // [N, D]
    const double* Y;  // [N, T]
    const double* W;  // packed by mt_pack_weights_kernel: [D / 16][T][20] (k-chunk major, padded rows)
    const double* b;  // [T] device
    double*       dL; // [N, T + (T & 1)] (GRAD): even row stride, so that the backward moves rows as 16-byte copies
    double*       partials_fb; // [grid, 1 + T]
    int64_t       N;
    int           D; // even (rows start on 16-byte boundaries); the last k-chunk may be partial
    int           T;
    int           stages;
    int           loss;
    double        loss_param;
    // batched evaluation of K single-target models that share X and Y: the K weight vectors play
    // the role of the targets, every "target" is compared with the same y_i, the losses are kept apart per model
    int           y_broadcast;     // 1: Y is [N, 1] and is used for every column
    int           per_target_loss; // 1: partial row = [total | dL sums (T) | loss sums (T)]
    int           fb_stride;       // doubles per row of partials_fb: 1 + T, or 1 + 2 T with per_target_loss
};

struct mt_backward_params
{
    const double* X;           // [N, D]
    const double* dL;          // [N, T + (T & 1)]
    double*       partials_gw; // [splits, T, D]
    int64_t       N;
    int           D;      // even; the last column tile may be partial
    int           T;
    int           t_ld;   // shared row stride of the dL tile (>= T, = 4 mod 16)
    int           splits;
    int           stages;
};

#ifdef __CUDACC__

// W [T, D] -> Wp [D / kMtBK][T][kMtLd]: the W tile of one k-chunk becomes ONE contiguous block that already has the
// padded shared-memory row stride, so the forward producer moves it with a single bulk copy per stage
static __global__ void __launch_bounds__(256) mt_pack_weights_kernel(const double* __restrict__ W, const int T, const int D,
                                                              double* __restrict__ Wp)
{
    const int64_t total = static_cast<int64_t>((D + kMtBK - 1) / kMtBK) * T * kMtLd;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    {
        const int     j   = static_cast<int>(i % kMtLd);
        const int64_t r   = i / kMtLd;
        const int     t   = static_cast<int>(r % T);
        const int64_t kc  = r / T;
        const int64_t col = kc * kMtBK + j;
        Wp[i]             = (j < kMtBK && col < D) ? W[static_cast<int64_t>(t) * D + col] : 0.0;
    }
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// sum / max over the 4 lanes that hold one accumulator row (lanes g*4 + {0..3})
__device__ __forceinline__ double quad_sum(double v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

__device__ __forceinline__ double quad_max(double v)
{
    double o = __shfl_xor_sync(0xffffffffu, v, 1);
    v        = (v < o) ? o : v;
    o        = __shfl_xor_sync(0xffffffffu, v, 2);
    return (v < o) ? o : v;
}

// Loss epilogue of `row_count` consecutive rows whose outputs sit in shared memory (row stride TPs): value, dL and
// the per-lane column sums. Several rows are evaluated TOGETHER: one row at a time the epilogue is a chain of
// dependent fp64 instructions (max -> exp -> sum -> log, or the elementwise loss) plus the latency of the targets'
// global load, ~4900 cycles per row with the tensor pipe idle behind it (measured: 2.2 of the forward's 14.6 ms at
// config 3); interleaved rows overlap those latencies. Lanes stride the targets (t = lane + 32 k). The softmax takes R
// rows per step (its chain is the longest); the elementwise losses take 8 / TL rows per step in a loop that is NOT
// unrolled further — every element carries the switch over the loss family, and R x TL copies of it (44 000
// instructions per kernel) cost more in instruction fetch than the extra overlap returned.
template <int TP, bool GRAD, int R>
__device__ __forceinline__ void mt_epilogue_rows(const mt_forward_params& p, const double* orows, const int64_t row_first,
                                                 const int row_count, const int lane, double& fsum,
                                                 double (&gbacc)[(TP + 31) / 32], double (&lossacc)[(TP + 31) / 32])
{
    constexpr int TPs = TP + 2;
    constexpr int TL  = (TP + 31) / 32;
    const int     T     = p.T;
    const int     dl_ld = T + (T & 1);
    if (p.loss == LOSS_CLASSNLL && p.y_broadcast == 0) // (batched single-output models: elementwise)
    {
        for (int r0 = 0; r0 < row_count; r0 += R)
        {
            // classnll.h:19-37
            double y[R][TL];
            bool   live[R];
#pragma unroll
            for (int j = 0; j < R; ++j)
            {
                live[j]            = r0 + j < row_count;
                const double* yrow = p.Y + (row_first + r0 + (live[j] ? j : 0)) * T;
#pragma unroll
                for (int k = 0; k < TL; ++k)
                {
                    const int t = lane + 32 * k;
                    y[j][k]     = (t < T) ? yrow[t] : 0.0;
                }
            }
            double omax[R], sumexp[R], dot[R];
            double ex[R][TL];
#pragma unroll
            for (int j = 0; j < R; ++j)
            {
                const double* orow = orows + (r0 + (live[j] ? j : 0)) * TPs;
                omax[j]            = -INFINITY;
#pragma unroll
                for (int k = 0; k < TL; ++k)
                {
                    const int t = lane + 32 * k;
                    if (t < T)
                    {
                        omax[j] = max_of(omax[j], orow[t]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < R; ++j)
            {
                omax[j] = warp_allreduce_max(omax[j]);
            }
#pragma unroll
            for (int j = 0; j < R; ++j)
            {
                const double* orow = orows + (r0 + (live[j] ? j : 0)) * TPs;
                sumexp[j] = dot[j] = 0.0;
#pragma unroll
                for (int k = 0; k < TL; ++k)
                {
                    // (no branch: the R x TL exponentials of a lane are independent chains the compiler interleaves;
                    // targets past T enter as -inf -> 0, and y is 0 there)
                    const int    t  = lane + 32 * k;
                    const bool   in = t < T;
                    const double o  = orow[in ? t : 0];
                    ex[j][k]        = exp_nonpositive(in ? __dsub_rn(o, omax[j]) : -INFINITY);
                    sumexp[j] += ex[j][k];
                    dot[j] += in ? __dmul_rn(__dadd_rn(1.0, y[j][k]), o) : 0.0;
                }
            }
#pragma unroll
            for (int j = 0; j < R; ++j)
            {
                sumexp[j] = warp_allreduce_sum(sumexp[j]);
                dot[j]    = warp_allreduce_sum(dot[j]);
            }
            // every lane holds the R sums: lane j takes row j, so that the warp evaluates ONE logarithm and ONE
            // reciprocal per step instead of R logarithms and R x TL divisions per lane (each a chain of dependent fp64
            // instructions on the unit the DMMAs use), and the results go back by shuffles. exp(o - max) x (1 / sum)
            // differs from the reference's quotient by at most one rounding (parity bar: 1e-12).
            double mine = 1.0;
#pragma unroll
            for (int j = 0; j < R; ++j)
            {
                mine = (lane == j) ? sumexp[j] : mine;
            }
            const double my_log = log(mine);
            const double my_rcp = __ddiv_rn(1.0, mine);
#pragma unroll
            for (int j = 0; j < R; ++j)
            {
                const double lg  = __shfl_sync(0xffffffffu, my_log, j);
                const double rcp = __shfl_sync(0xffffffffu, my_rcp, j);
                if (live[j])
                {
                    fsum += __dadd_rn(__dsub_rn(lg, __dmul_rn(0.5, dot[j])), omax[j]);
                    if (GRAD)
                    {
                        const int64_t row = row_first + r0 + j;
#pragma unroll
                        for (int k = 0; k < TL; ++k)
                        {
                            const int t = lane + 32 * k;
                            if (t < T)
                            {
                                const double gv =
                                    __dsub_rn(__dmul_rn(ex[j][k], rcp), __dmul_rn(0.5, __dadd_rn(1.0, y[j][k])));
                                p.dL[row * dl_ld + t] = gv;
                                gbacc[k] += gv;
                            }
                        }
                    }
                }
            }
            if (GRAD && dl_ld != T && lane == 0)
            {
#pragma unroll
                for (int j = 0; j < R; ++j)
                {
                    if (live[j])
                    {
                        p.dL[(row_first + r0 + j) * dl_ld + T] = 0.0; // the pad column of an odd T
                    }
                }
            }
        }
        return;
    }

    constexpr int RE = (8 / TL) < 2 ? 2 : (8 / TL); // rows per step: eight independent elements per lane
#pragma unroll 1
    for (int r0 = 0; r0 < row_count; r0 += RE)
    {
        double sum[RE];
#pragma unroll
        for (int j = 0; j < RE; ++j)
        {
            const bool    live = r0 + j < row_count;
            const int64_t row  = row_first + r0 + (live ? j : 0);
            const double* orow = orows + (r0 + (live ? j : 0)) * TPs;
            const double* yrow = p.Y + row * (p.y_broadcast ? 1 : T);
            sum[j]             = 0.0;
#pragma unroll
            for (int k = 0; k < TL; ++k)
            {
                const int t = lane + 32 * k;
                if (t < T && live)
                {
                    const double y = p.y_broadcast ? yrow[0] : yrow[t];
                    double       term, gv;
                    loss_term<GRAD>(p.loss, p.loss_param, y, orow[t], term, gv);
                    sum[j] += term;
                    lossacc[k] += loss_post(p.loss, term); // one output per model: the factor applies per term
                    if (GRAD)
                    {
                        p.dL[row * dl_ld + t] = gv;
                        gbacc[k] += gv;
                    }
                }
            }
            if (GRAD && dl_ld != T && lane == 0 && live)
            {
                p.dL[row * dl_ld + T] = 0.0; // the pad column of an odd T
            }
        }
#pragma unroll
        for (int j = 0; j < RE; ++j)
        {
            sum[j] = warp_allreduce_sum(sum[j]);
            if (r0 + j < row_count)
            {
                fsum += loss_post(p.loss, sum[j]);
            }
        }
    }
}

template <int NT8, bool GRAD>
__global__ void __launch_bounds__(kMtThreads, 1) mt_forward_kernel(const mt_forward_params p)
{
    constexpr int TP = NT8 * 8; // padded targets

    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar = full_bar + kMtMaxStages;
    constexpr int TPs   = TP + 2; // row stride of the output tile: 16-byte fragment stores hit 8 distinct bank groups
    constexpr int kFold = 8 * TPs; // distance between the warps' fold rows (each inside that warp's 8 tile rows)
    static_assert(1 + 2 * TP <= kFold, "the fold row [loss | dL sums | per-target loss sums] must fit 8 tile rows");
    double*   otile     = reinterpret_cast<double*>(smem + 256); // [kMtBM][TPs] outputs of the current tile
    otile               = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(otile) + 127) / 128 * 128);
    double*   fold      = otile; // after the last tile the output tile is dead: it hosts the per-warp fold rows
    double*   stage0    = otile + (kMtBM / 2) * TPs; // 8 rows per warp at a time: the two m8 tiles take turns
    stage0              = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(stage0) + 127) / 128 * 128);
    constexpr int kStageDoubles = (kMtBM + TP) * kMtLd;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int g    = lane >> 2; // fragment row / column group
    const int tig  = lane & 3;  // thread in group

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < p.stages; ++s)
        {
            // every producer thread arrives when its X chunks landed; one more arrival carries the W tile's bytes
            mbar_init(full_bar + s, kMtProducerWarps * 32 + 1);
            mbar_init(empty_bar + s, kMtWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int     D      = p.D;
    const int     T      = p.T;
    const int     kchunks = (D + kMtBK - 1) / kMtBK; // the last one holds D - (kchunks - 1) * 16 features
    const int64_t tiles  = (p.N + kMtBM - 1) / kMtBM;
    const int     dl_ld  = T + (T & 1);

    if (warp >= kMtWarps)
    {
        // ===== producer warpgroup: 128 threads stream 16-byte chunks (cp.async); a tile row is only 128 bytes =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kMtProducerRegs));
        const int pt = threadIdx.x - kMtWarps * 32;
        constexpr int kChunks = kMtBK / 2; // 16-byte chunks per tile row
        uint64_t      policy_w;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy_w));
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x)
        {
            const int64_t row0 = tile * kMtBM;
            const int     rows = static_cast<int>(min(static_cast<int64_t>(kMtBM), p.N - row0));
            for (int kc = 0; kc < kchunks; ++kc, ++it)
            {
                const int      s     = it % p.stages;
                const uint32_t phase = static_cast<uint32_t>((it / p.stages) & 1);
                mbar_wait(empty_bar + s, phase ^ 1u);
                double*       xs   = stage0 + static_cast<int64_t>(s) * kStageDoubles;
                double*       ws   = xs + kMtBM * kMtLd;
                const double* xsrc = p.X + row0 * D + kc * kMtBK;
                if (pt == 0)
                {
                    // the packed W tile of this k-chunk: one contiguous bulk copy (re-read by every CTA: keep in L2)
                    const uint32_t bytes = static_cast<uint32_t>(T * kMtLd * 8);
                    mbar_arrive_expect_tx(full_bar + s, bytes);
                    bulk_copy_g2s(ws, p.W + static_cast<int64_t>(kc) * T * kMtLd, bytes, full_bar + s, policy_w);
                }
                // 16-byte chunks of this k-chunk that exist (D is even): the tail of the last k-chunk is not copied,
                // the consumers mask those columns
                const int valid = min(kChunks, (D - kc * kMtBK) / 2);
                for (int c = pt; c < rows * kChunks; c += kMtProducerWarps * 32)
                {
                    const int r = c / kChunks, j = c % kChunks;
                    if (j < valid)
                    {
                        cp_async_16(xs + r * kMtLd + j * 2, xsrc + static_cast<int64_t>(r) * D + j * 2);
                    }
                }
                cp_async_arrive(full_bar + s);
            }
        }
        return;
    }

    // ===== consumer warps: rows [warp*16, warp*16 + 16) of the tile, all targets =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kMtConsumerRegs));
    constexpr int TL = (TP + 31) / 32; // targets per lane in the row-wise epilogue (target t = lane + 32 * k)
    double gbacc[TL];                  // column sums of dL owned by this lane
    double lossacc[TL];                // per-target loss sums (batched single-target models)
#pragma unroll
    for (int k = 0; k < TL; ++k)
    {
        gbacc[k]   = 0.0;
        lossacc[k] = 0.0;
    }
    double fsum = 0.0;

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x)
    {
        double acc[2][NT8][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
        {
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
                acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
            }
        }

        for (int kc = 0; kc < kchunks; ++kc, ++it)
        {
            const int      s     = it % p.stages;
            const uint32_t phase = static_cast<uint32_t>((it / p.stages) & 1);
            mbar_wait(full_bar + s, phase);
            const double* xs = stage0 + static_cast<int64_t>(s) * kStageDoubles;
            const double* ws = xs + kMtBM * kMtLd;
            const double* xa = xs + (warp * 16 + g) * kMtLd + tig;
            const double* wb = ws + g * kMtLd + tig;
            const int kleft = D - kc * kMtBK; // features in this k-chunk (>= 16 except for the last one)
#pragma unroll
            for (int kk = 0; kk < kMtBK / 4; ++kk)
            {
                // columns past D were not copied: whatever the stage holds there must not reach the accumulators
                const bool   in = kk * 4 + tig < kleft;
                const double a0 = in ? xa[kk * 4] : 0.0;
                const double a1 = in ? xa[8 * kMtLd + kk * 4] : 0.0;
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt)
                {
                    const double bv = wb[nt * 8 * kMtLd + kk * 4];
                    dmma_m8n8k4(acc[0][nt][0], acc[0][nt][1], a0, bv);
                    dmma_m8n8k4(acc[1][nt][0], acc[1][nt][1], a1, bv);
                }
            }
            mbar_release_stage(empty_bar + s, lane);
        }

        // ----- epilogue: accumulator fragments (+ bias) -> this warp's 16 rows of the shared output tile, then a
        //       compact row-wise loop (lanes stride the targets: coalesced Y reads and dL writes, small code) -----
        double* orows = otile + (warp * 8) * TPs;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
        {
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
                const int col = nt * 8 + tig * 2;
                const double b0 = (col < T) ? p.b[col] : 0.0;
                const double b1 = (col + 1 < T) ? p.b[col + 1] : 0.0;
                *reinterpret_cast<double2*>(orows + g * TPs + col) =
                    make_double2(acc[mt][nt][0] + b0, acc[mt][nt][1] + b1);
            }
        __syncwarp();

        const int64_t row_first = tile * kMtBM + warp * 16 + mt * 8;
        const int     row_count = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(8), p.N - row_first)));
        mt_epilogue_rows<TP, GRAD, kMtEpilogueRows>(p, orows, row_first, row_count, lane, fsum, gbacc, lossacc);
        __syncwarp(); // the next m8 tile / the next tile's epilogue overwrites this warp's rows
        }
    }

    // ----- fold the 8 consumer warps in warp order -----
    double* mine = fold + warp * kFold;
    if (lane == 0)
    {
        mine[0] = fsum; // identical on every lane
    }
#pragma unroll
    for (int k = 0; k < TL; ++k)
    {
        const int t = lane + 32 * k;
        if (t < TP)
        {
            mine[1 + t]      = GRAD ? gbacc[k] : 0.0;
            mine[1 + TP + t] = lossacc[k];
        }
    }
    named_bar_sync(1, kMtWarps * 32);
    double*   out   = p.partials_fb + static_cast<int64_t>(blockIdx.x) * p.fb_stride;
    const int ncols = p.per_target_loss ? (1 + 2 * T) : (GRAD ? 1 + T : 1);
    for (int c = threadIdx.x; c < ncols; c += kMtWarps * 32)
    {
        const int src = (c <= T) ? c : (1 + TP + (c - 1 - T)); // [loss | dL sums | per-target loss sums]
        double    sum = 0.0;
#pragma unroll
        for (int q = 0; q < kMtWarps; ++q)
        {
            sum += fold[q * kFold + src];
        }
        out[c] = sum;
    }
}

// 2-D tiled copy global -> shared through a tensor map (the TMA engine proper; SASS UTMALDG): box = 16 features x 128
// rows of X, rows and columns past the matrix come back as zeros, completion as transaction bytes on `bar`
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tensor_map, const int c0, const int c1,
                                            uint64_t* bar, const uint64_t policy)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                 " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(smem_dst)),
                 "l"(tensor_map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

// The forward contraction fed by the tensor-map engine: ONE lane issues, per k-chunk, one 2-D tiled copy of the X tile
// (128 rows x 16 features = dense 128-byte rows under the 128-byte swizzle; rows / columns past the matrix arrive as
// zeros, so nothing is masked) and one bulk copy of the packed W tile — instead of 128 threads issuing 1024 16-byte
// cp.async chunks per stage. Fragment rows are permuted — lane group g reads tile row (g % 4) * 2 + g / 4 — so that
// with the swizzle (16-byte chunk index ^= row % 8) both half-warps of every 64-bit A-fragment load hit 16 distinct
// 8-byte slots. The epilogue stays on the consumer warps, each on its own 16 rows (no cross-warp synchronisation): a
// variant with the epilogue on dedicated warps, overlapped with the next tile's DMMAs, was measured SLOWER — the loss's
// fp64 instructions share the FP64 unit with the DMMAs, and behind a saturated DMMA queue every one of them waited
// ~130 cycles (profiles/r02_mt_forward_roles.jsonl).
template <int NT8, bool GRAD>
__global__ void __launch_bounds__(kMtThreads, 1)
    mt_forward_tma_kernel(const mt_forward_params p, const __grid_constant__ CUtensorMap x_map)
{
    constexpr int TP    = NT8 * 8;
    constexpr int TPs   = TP + 2;  // row stride of the output tile
    constexpr int TL    = (TP + 31) / 32;
    constexpr int kFold = 16 * TPs; // distance between the warps' fold rows (each inside that warp's 16 tile rows)
    static_assert(1 + 2 * TP <= kFold, "the fold row [loss | dL sums | per-target loss sums] must fit the warp's rows");

    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar = full_bar + kMtMaxStages;
    constexpr int kOtileBytes = (kMtBM * TPs * 8 + 1023) / 1024 * 1024;
    double*        otile  = reinterpret_cast<double*>(smem + kMtTmaHeader); // [128][TPs]: 16 rows per warp
    unsigned char* stage0 = smem + kMtTmaHeader + kOtileBytes;              // 1 KiB aligned (swizzle atom)
    const int      T      = p.T;
    const int      D      = p.D;
    const int      w_bytes = T * kMtLd * 8;
    // the W region holds all TP padded rows (the fragment loads of the rows past T read stale shared memory whose
    // products land in accumulator columns nobody looks at — it must still be this CTA's shared memory), and every
    // stage starts on a 1 KiB boundary: the swizzle is a function of the shared-memory ADDRESS bits
    constexpr int  stage_bytes = kMtXTileBytes + (TP * kMtLd * 8 + 1023) / 1024 * 1024;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < p.stages; ++s)
        {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, kMtWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int     kchunks = (D + kMtBK - 1) / kMtBK;
    const int64_t tiles   = (p.N + kMtBM - 1) / kMtBM;

    if (warp >= kMtWarps)
    {
        // ===== producer warpgroup: one lane, two copies per stage =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kMtProducerRegs));
        if (warp == kMtWarps && lane == 0)
        {
            const uint64_t policy_x = l2_policy_evict_first();
            const uint64_t policy_w = l2_policy_evict_last();
            int            s        = 0;
            uint32_t       phase    = 1; // the first lap passes at once
            for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x)
            {
                const int row0 = static_cast<int>(tile * kMtBM);
                for (int kc = 0; kc < kchunks; ++kc)
                {
                    mbar_wait(empty_bar + s, phase);
                    unsigned char* xs = stage0 + static_cast<int64_t>(s) * stage_bytes;
                    mbar_arrive_expect_tx(full_bar + s, static_cast<uint32_t>(kMtXTileBytes + w_bytes));
                    tma_load_2d(xs, &x_map, kc * kMtBK, row0, full_bar + s, policy_x);
                    bulk_copy_g2s(xs + kMtXTileBytes, p.W + static_cast<int64_t>(kc) * T * kMtLd,
                                  static_cast<uint32_t>(w_bytes), full_bar + s, policy_w);
                    if (++s == p.stages)
                    {
                        s = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
        return;
    }

    // ===== consumer warps: rows [warp*16, warp*16 + 16) of the tile, all targets =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kMtConsumerRegs));
    const int g   = lane >> 2;
    const int tig = lane & 3;
    const int rho = (g & 3) * 2 + (g >> 2); // tile row (within an m8 tile) served by this lane group
    // byte offset of this lane's A element in k-step kk: row * 128 + ((chunk ^ rho) << 4) + (tig & 1) * 8 with
    // chunk = 2 kk + (tig >> 1)  (rho == row % 8: the warp's rows start at a multiple of 8)
    const int a_row = (warp * 16 + rho) * 128 + (tig & 1) * 8;
    const int a_hi  = tig >> 1;

    double gbacc[TL], lossacc[TL];
#pragma unroll
    for (int k = 0; k < TL; ++k)
    {
        gbacc[k] = lossacc[k] = 0.0;
    }
    double fsum = 0.0;

    int      s     = 0;
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x)
    {
        double acc[2][NT8][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
        {
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
                acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
            }
        }

        for (int kc = 0; kc < kchunks; ++kc)
        {
            mbar_wait(full_bar + s, phase);
            const unsigned char* xs = stage0 + static_cast<int64_t>(s) * stage_bytes;
            const double*        wb = reinterpret_cast<const double*>(xs + kMtXTileBytes) + g * kMtLd + tig;
#pragma unroll
            for (int kk = 0; kk < kMtBK / 4; ++kk)
            {
                const int    off = a_row + (((2 * kk + a_hi) ^ rho) << 4);
                const double a0  = *reinterpret_cast<const double*>(xs + off);
                const double a1  = *reinterpret_cast<const double*>(xs + off + 8 * 128);
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt)
                {
                    const double bv = wb[nt * 8 * kMtLd + kk * 4];
                    dmma_m8n8k4(acc[0][nt][0], acc[0][nt][1], a0, bv);
                    dmma_m8n8k4(acc[1][nt][0], acc[1][nt][1], a1, bv);
                }
            }
            mbar_release_stage(empty_bar + s, lane);
            if (++s == p.stages)
            {
                s = 0;
                phase ^= 1u;
            }
        }

        // ----- epilogue: ALL accumulator fragments (+ bias) -> this warp's 16 rows of the shared output tile, then the
        //       row-wise loss over them, eight rows at a time (the accumulator registers are free by then) -----
        double* orows = otile + (warp * 16) * TPs;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
        {
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
                const int    col = nt * 8 + tig * 2;
                const double b0  = (col < T) ? p.b[col] : 0.0;
                const double b1  = (col + 1 < T) ? p.b[col + 1] : 0.0;
                *reinterpret_cast<double2*>(orows + (mt * 8 + rho) * TPs + col) =
                    make_double2(acc[mt][nt][0] + b0, acc[mt][nt][1] + b1);
            }
        }
        __syncwarp();
        const int64_t row_first = tile * kMtBM + warp * 16;
        const int     row_count =
            static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(16), p.N - row_first)));
        mt_epilogue_rows<TP, GRAD, kMtEpilogueRowsTma>(p, orows, row_first, row_count, lane, fsum, gbacc, lossacc);
        __syncwarp(); // the next tile's epilogue overwrites this warp's rows
    }

    // ----- fold the 8 consumer warps in warp order (the output tile is dead: it hosts the per-warp fold rows) -----
    double* mine = otile + warp * kFold;
    if (lane == 0)
    {
        mine[0] = fsum; // identical on every lane
    }
#pragma unroll
    for (int k = 0; k < TL; ++k)
    {
        const int t = lane + 32 * k;
        if (t < TP)
        {
            mine[1 + t]      = GRAD ? gbacc[k] : 0.0;
            mine[1 + TP + t] = lossacc[k];
        }
    }
    named_bar_sync(1, kMtWarps * 32);
    double*   out   = p.partials_fb + static_cast<int64_t>(blockIdx.x) * p.fb_stride;
    const int ncols = p.per_target_loss ? (1 + 2 * T) : (GRAD ? 1 + T : 1);
    for (int c = threadIdx.x; c < ncols; c += kMtWarps * 32)
    {
        const int src = (c <= T) ? c : (1 + TP + (c - 1 - T)); // [loss | dL sums | per-target loss sums]
        double    sum = 0.0;
#pragma unroll
        for (int q = 0; q < kMtWarps; ++q)
        {
            sum += otile[q * kFold + src];
        }
        out[c] = sum;
    }
}

// grid.x = (D / kMtBwdCols) * splits; CTA (ct, split) accumulates gW[:, ct*128 .. +128) over its row range.
template <int NT8>
__global__ void __launch_bounds__(kMtThreads, 1) mt_backward_kernel(const mt_backward_params p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar = full_bar + kMtMaxStages;
    double*   stage0    = reinterpret_cast<double*>(smem + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int g    = lane >> 2;
    const int tig  = lane & 3;

    const int T      = p.T;
    const int D      = p.D;
    const int t_ld   = p.t_ld;
    const int dl_ld  = T + (T & 1);
    const int ctiles = (D + kMtBwdCols - 1) / kMtBwdCols;
    const int ct     = blockIdx.x % ctiles;
    const int cols   = min(kMtBwdCols, D - ct * kMtBwdCols); // even; the last column tile may be partial
    const int split  = blockIdx.x / ctiles;

    // stage layout: dL tile [kMtBwdRows][t_ld] then X tile [kMtBwdRows][kMtBwdLd]; each part rounded to 128 bytes
    const int dl_doubles    = (kMtBwdRows * t_ld + 15) / 16 * 16;
    const int stage_doubles = dl_doubles + kMtBwdRows * kMtBwdLd;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < p.stages; ++s)
        {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, kMtWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // contiguous row range of this split, in whole stages
    const int64_t chunks_total = (p.N + kMtBwdRows - 1) / kMtBwdRows;
    const int64_t chunks_each  = (chunks_total + p.splits - 1) / p.splits;
    const int64_t chunk_begin  = min(chunks_total, split * chunks_each);
    const int64_t chunk_end    = min(chunks_total, chunk_begin + chunks_each);

    if (warp >= kMtWarps)
    {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kMtProducerRegs));
        if (warp != kMtWarps)
        {
            return;
        }
        const uint64_t policy_x = l2_policy_evict_first();
        uint64_t       policy_g;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy_g));
        int it = 0;
        for (int64_t chunk = chunk_begin; chunk < chunk_end; ++chunk, ++it)
        {
            const int      s     = it % p.stages;
            const uint32_t phase = static_cast<uint32_t>((it / p.stages) & 1);
            mbar_wait(empty_bar + s, phase ^ 1u);
            const int64_t row0 = chunk * kMtBwdRows;
            const int     rows = static_cast<int>(min(static_cast<int64_t>(kMtBwdRows), p.N - row0));
            double*       gs   = stage0 + static_cast<int64_t>(s) * stage_doubles;
            double*       xs   = gs + dl_doubles;
            if (lane == 0)
            {
                mbar_arrive_expect_tx(full_bar + s, static_cast<uint32_t>(rows * (dl_ld + cols) * 8));
            }
            __syncwarp();
            if (lane < rows)
            {
                bulk_copy_g2s(gs + lane * t_ld, p.dL + (row0 + lane) * dl_ld, dl_ld * 8, full_bar + s, policy_g);
                bulk_copy_g2s(xs + lane * kMtBwdLd, p.X + (row0 + lane) * D + ct * kMtBwdCols, cols * 8, full_bar + s,
                              policy_x);
            }
        }
        return;
    }

    // consumer warp: n8 tiles {2*warp, 2*warp + 1} of the 16 column tiles, all NT8 target tiles
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kMtConsumerRegs));
    double acc[NT8][2][2];
#pragma unroll
    for (int mt = 0; mt < NT8; ++mt)
    {
        acc[mt][0][0] = acc[mt][0][1] = acc[mt][1][0] = acc[mt][1][1] = 0.0;
    }

    int it = 0;
    for (int64_t chunk = chunk_begin; chunk < chunk_end; ++chunk, ++it)
    {
        const int      s     = it % p.stages;
        const uint32_t phase = static_cast<uint32_t>((it / p.stages) & 1);
        mbar_wait(full_bar + s, phase);
        const int     rows = static_cast<int>(min(static_cast<int64_t>(kMtBwdRows), p.N - chunk * kMtBwdRows));
        const double* gs   = stage0 + static_cast<int64_t>(s) * stage_doubles;
        const double* xs   = gs + dl_doubles;
        if (rows == kMtBwdRows)
        {
            // full stage: no predicates at all. The dL columns past T (the last target tile) are whatever the shared
            // memory holds — they only reach accumulator rows m >= T, which are never stored — and plain loads let the
            // compiler run them well ahead of the DMMAs that use them (the predicated form re-zeroed each register
            // right behind the DMMA reading it and kept only three fragments in flight).
#pragma unroll
            for (int kk = 0; kk < kMtBwdRows / 4; ++kk)
            {
                const int    k  = kk * 4 + tig;
                const double b0 = xs[k * kMtBwdLd + (2 * warp) * 8 + g];
                const double b1 = xs[k * kMtBwdLd + (2 * warp + 1) * 8 + g];
#pragma unroll
                for (int mt = 0; mt < NT8; ++mt)
                {
                    const double a = gs[k * t_ld + mt * 8 + g];
                    dmma_m8n8k4(acc[mt][0][0], acc[mt][0][1], b0, a);
                    dmma_m8n8k4(acc[mt][1][0], acc[mt][1][1], b1, a);
                }
            }
        }
        else
        {
#pragma unroll 1
            for (int kk = 0; kk < kMtBwdRows / 4; ++kk)
            {
                const int    k  = kk * 4 + tig; // row inside the stage
                const bool   in = k < rows;
                const double b0 = in ? xs[k * kMtBwdLd + (2 * warp) * 8 + g] : 0.0;
                const double b1 = in ? xs[k * kMtBwdLd + (2 * warp + 1) * 8 + g] : 0.0;
#pragma unroll
                for (int mt = 0; mt < NT8; ++mt)
                {
                    const int    m = mt * 8 + g;
                    const double a = (in && m < T) ? gs[k * t_ld + m] : 0.0;
                    dmma_m8n8k4(acc[mt][0][0], acc[mt][0][1], b0, a);
                    dmma_m8n8k4(acc[mt][1][0], acc[mt][1][1], b1, a);
                }
            }
        }
        mbar_release_stage(empty_bar + s, lane);
    }

    // The products are formed transposed — the X fragment is the A operand (reused by the NT8 DMMAs that follow, the
    // operand pattern of the forward, which the tensor pipe sustains at a higher rate), the dL fragment the B operand —
    // so lane (g, tig) holds gW^T[column (2 warp + j) * 8 + g][target 8 mt + 2 tig + {0, 1}].
    double* out = p.partials_gw + static_cast<int64_t>(split) * T * D;
#pragma unroll
    for (int mt = 0; mt < NT8; ++mt)
    {
#pragma unroll
        for (int j = 0; j < 2; ++j)
        {
            // columns past the partial tile multiplied whatever the stage held: they are simply not stored
            const int col = ct * kMtBwdCols + (2 * warp + j) * 8 + g;
#pragma unroll
            for (int e = 0; e < 2; ++e)
            {
                const int t = mt * 8 + tig * 2 + e;
                if (t < T && col < D)
                {
                    out[static_cast<int64_t>(t) * D + col] = acc[mt][j][e];
                }
            }
        }
    }
}

#endif // __CUDACC__
} // namespace nb200
