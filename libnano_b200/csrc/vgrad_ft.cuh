// vgrad_ft.cuh — FEW targets (2 <= T <= 10, D <= 2048): ONE pass over X per evaluation.
//
// BASELINE.json north_star: "these paths are HBM-bound when there are few targets". With a handful of targets the
// two contractions of linear::function_t::do_eval (outputs = X W^T + b, src/linear/util.cpp:19-20, and
// gW += dL^T X, src/linear/function.cpp:68-70) need 4 T flops per 8 bytes of X — below the fp64 ridge up to T ~ 10 —
// so the tensor-pipe pair of vgrad_mt.cuh, which reads X twice and round-trips dL through HBM, leaves most of the
// bandwidth unused (4 Mi x 784 x 10: 23.9 ms for 26 GB of X). This kernel is the single-target design (vgrad_t1.cuh)
// generalised to T targets: X is read from HBM once and from shared memory once.
//
//   * persistent CTAs, a producer warp streaming tiles of whole rows (+ their targets) into a ring of shared-memory
//     stages with 1-D bulk copies behind full / empty mbarriers, exactly as in vgrad_t1.cuh;
//   * all eight consumer warps share every row: warp w owns the columns [w * 64 KC, (w + 1) * 64 KC); its lanes keep
//     their slice of W (T x 2 KC doubles) and of the gradient accumulator (same size) in REGISTERS for the whole launch;
//   * one iteration = RW rows: the RW * T partial dots of the warp's column slice are reduced together by recursive
//     halving (warp_reduce_rows), exchanged between the eight warps through shared memory (one named barrier per
//     iteration, double-buffered slots) and summed in warp order, so that lane group q = r * T + t holds output (r, t);
//   * ONE lane per (row, target) evaluates the loss (elementwise losses: loss.cuh; classnll, classnll.h:19-37, with
//     shuffles among the T lanes of a row), dL goes through a per-warp scratch to every lane, and
//     g[t][:] += dL[r][t] * x[r][:] runs out of the registers that still hold the rows;
//   * the CTA's partial row [loss | sum dL (T) | sum dL x^T (T * D)] goes to HBM; the fold over CTAs and the final
//     scaling are the shared reduce_partials_kernel / finish_kernel (reduce.cuh).
#pragma once

#include "common.cuh"
#include "loss.cuh"

namespace nb200
{
constexpr int kFtWarps         = 8; // consumer warps (all of them share every row)
constexpr int kFtProducerWarps = 4; // one warpgroup: its registers are handed to the consumers
constexpr int kFtThreads       = (kFtWarps + kFtProducerWarps) * 32;
constexpr int kFtMaxStages     = 8;
constexpr int kFtExchange      = 128;                                  // [2][kFtWarps][32] doubles
constexpr int kFtScratch       = kFtExchange + 2 * kFtWarps * 32 * 8;  // [kFtWarps][32] doubles (dL of the iteration)
constexpr int kFtTail          = kFtScratch + kFtWarps * 32 * 8;       // [sets][2][32] doubles (loss / dL sums)
constexpr int kFtHeaderSize    = 8192;

struct ft_params
{
    const double* X;        // [N, D] row-major, padded by >= 16 bytes at the end
    const double* Y;        // [N, T], padded likewise
    const double* W;        // [T, D] device
    const double* b;        // [T] device
    double*       partials; // [grid, 1 + T + T * D]
    int64_t       N;
    int64_t       num_tiles;
    int           D;              // even
    int           rows_per_stage; // multiple of RW; rows_per_stage * T even
    int           stages;
    int           stage_stride; // bytes between stages (multiple of 128)
    int           y_offset;     // byte offset of the Y tile inside a stage (multiple of 128)
    int           loss;
    double        loss_param;
};

#ifdef __CUDACC__

__host__ __device__ constexpr int ft_pow2_ceil(const int value)
{
    int p = 1;
    while (p < value)
    {
        p *= 2;
    }
    return p;
}

template <int T, int KC, int RW, int NS, bool GRAD>
__global__ void __launch_bounds__(kFtThreads, 1) vgrad_ft_kernel(const ft_params p)
{
    // NS sets of WPS consumer warps take the iterations in turn (NS = 2: while one set is in its reduction / loss
    // chain the other one multiplies); the warps of a set share every row of their iterations
    constexpr int WPS = kFtWarps / NS;
    static_assert(NS == 1 || NS == 2, "one or two alternating warp sets");
    constexpr int PAIRS = RW * T;               // (row, target) pairs of one iteration
    constexpr int P     = ft_pow2_ceil(PAIRS);  // padded to a power of two for the halving reduction
    constexpr int SHIFT = 5 - log2_of(P);       // pair q lives in lanes [q << SHIFT, (q + 1) << SHIFT)
    constexpr int CPW   = 64 * KC;              // columns per warp (128-bit loads: 2 per lane and chunk)
    static_assert(P <= 32 && T >= 2, "rows per iteration x targets must fit one warp");

    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t*      full_bar   = reinterpret_cast<uint64_t*>(smem);
    uint64_t*      empty_bar  = full_bar + kFtMaxStages;
    double*        exchange   = reinterpret_cast<double*>(smem + kFtExchange);
    double*        scratch    = reinterpret_cast<double*>(smem + kFtScratch);
    double*        tail       = reinterpret_cast<double*>(smem + kFtTail);
    unsigned char* stage_base = smem + kFtHeaderSize;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int D    = p.D;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < p.stages; ++s)
        {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, kFtWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t rps = p.rows_per_stage;
    if (warp >= kFtWarps)
    {
        // ===== producer warpgroup: one lane feeds the ring =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(40));
        if (warp == kFtWarps && lane == 0)
        {
            const uint64_t policy = l2_policy_evict_first();
            int            s      = 0;
            uint32_t       phase  = 1;
            for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
            {
                mbar_wait(empty_bar + s, phase);
                const int64_t  row0    = tile * rps;
                const int64_t  rows    = min(rps, p.N - row0);
                const uint32_t bytes_x = static_cast<uint32_t>((rows * D * 8 + 15) / 16 * 16);
                const uint32_t bytes_y = static_cast<uint32_t>((rows * T * 8 + 15) / 16 * 16);
                unsigned char* stage   = stage_base + static_cast<int64_t>(s) * p.stage_stride;
                mbar_arrive_expect_tx(full_bar + s, bytes_x + bytes_y);
                bulk_copy_g2s(stage, p.X + row0 * D, bytes_x, full_bar + s, policy);
                bulk_copy_g2s(stage + p.y_offset, p.Y + row0 * T, bytes_y, full_bar + s, policy);
                if (++s == p.stages)
                {
                    s = 0;
                    phase ^= 1u;
                }
            }
        }
        return;
    }

    // ===== consumer warps =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(232));
    const int set   = warp / WPS;
    const int wis   = warp % WPS; // warp in set: which column slice
    const int cbase = wis * CPW;

    // this lane's slice of W and of the gradient accumulator
    double w[T][2 * KC];
    double g[T][2 * KC];
#pragma unroll
    for (int t = 0; t < T; ++t)
    {
#pragma unroll
        for (int k = 0; k < KC; ++k)
        {
            const int col   = cbase + (k * 32 + lane) * 2;
            const bool in   = col < D;
            w[t][2 * k]     = in ? p.W[static_cast<int64_t>(t) * D + col] : 0.0;
            w[t][2 * k + 1] = in ? p.W[static_cast<int64_t>(t) * D + col + 1] : 0.0;
            g[t][2 * k] = g[t][2 * k + 1] = 0.0;
        }
    }

    // the (row, target) pair this lane owns after the reduction: q = lane >> SHIFT, fixed for the whole launch
    const int    my_pair  = lane >> SHIFT;
    const bool   owner    = (lane & ((1 << SHIFT) - 1)) == 0 && my_pair < PAIRS;
    const int    my_r     = my_pair / T;
    const int    my_t     = my_pair - my_r * T;
    const double my_bias  = (my_pair < PAIRS) ? p.b[my_t] : 0.0;
    double       fsum     = 0.0; // loss terms of this lane's pair slot
    double       gbsum    = 0.0; // dL of this lane's pair slot (always target my_t)
    double*      my_dl    = scratch + warp * 32;

    int      stage_index = 0;
    uint32_t phase       = 0;
    int      parity      = 0;
    int      turn        = 0; // whose iteration comes next (iterations are dealt round-robin over the sets)
    double*  set_exchange = exchange + static_cast<int64_t>(set) * (2 * WPS * 32);
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
    {
        const int s = stage_index;
        mbar_wait(full_bar + s, phase);
        const uint32_t wrapped = (stage_index + 1 == p.stages) ? 1u : 0u;
        stage_index            = wrapped ? 0 : stage_index + 1;
        phase ^= wrapped;

        const unsigned char* stage = stage_base + static_cast<int64_t>(s) * p.stage_stride;
        const double*        xs    = reinterpret_cast<const double*>(stage);
        const double*        ys    = reinterpret_cast<const double*>(stage + p.y_offset);
        const int            rows  = static_cast<int>(min(rps, p.N - tile * rps));

        // the last iteration of this tile that belongs to this warp's set: the stage is released right after its
        // rows are in registers (at once when the set has no iteration in the tile)
        const int n_iter   = (rows + RW - 1) / RW;
        int       last_own = n_iter - 1;
        if constexpr (NS > 1)
        {
            last_own -= ((turn + n_iter - 1) % NS == set) ? 0 : 1; // NS == 2: the last or the one before
            if (last_own < 0)
            {
                if (lane == 0)
                {
                    mbar_arrive(empty_bar + s);
                }
            }
        }

        for (int i = 0, j = 0; i < rows; i += RW, ++j)
        {
            if constexpr (NS > 1)
            {
                const bool mine_now = turn == set;
                turn                = (turn + 1 == NS) ? 0 : turn + 1;
                if (!mine_now)
                {
                    continue;
                }
            }
            const int nrows = min(RW, rows - i);

            // RW row slices -> registers
            double x[RW][2 * KC];
#pragma unroll
            for (int k = 0; k < KC; ++k)
            {
                const int  col = cbase + (k * 32 + lane) * 2;
                const bool in  = col < D;
#pragma unroll
                for (int r = 0; r < RW; ++r)
                {
                    double2 xv = make_double2(0.0, 0.0);
                    if (in && r < nrows)
                    {
                        xv = *reinterpret_cast<const double2*>(xs + static_cast<int64_t>(i + r) * D + col);
                    }
                    x[r][2 * k]     = xv.x;
                    x[r][2 * k + 1] = xv.y;
                }
            }
            const double y = (owner && my_r < nrows) ? ys[(i + my_r) * T + my_t] : 0.0;

            // the stage can be refilled as soon as its last rows (and targets) are in registers
            if (j == last_own)
            {
                mbar_release_stage(empty_bar + s, lane);
            }

            // partial outputs of this warp's column slice for every (row, target) pair, reduced over the lanes. The
            // first halving step is fused with the dot products (pairs j and j + P / 2 are produced together and only
            // the kept half stays in registers); warp_reduce_rows_from does the rest.
            const auto dot = [&](const int q)
            {
                double a0 = 0.0, a1 = 0.0;
                if (q < PAIRS)
                {
                    const int r = q / T, t = q % T;
#pragma unroll
                    for (int k = 0; k < KC; ++k)
                    {
                        a0 = fma(x[r][2 * k], w[t][2 * k], a0);
                        a1 = fma(x[r][2 * k + 1], w[t][2 * k + 1], a1);
                    }
                }
                return a0 + a1;
            };
            double mine;
            if constexpr (P >= 2)
            {
                double     half[P / 2];
                const bool upper = (lane & 16) != 0;
#pragma unroll
                for (int j = 0; j < P / 2; ++j)
                {
                    const double lo = dot(j), hi = dot(j + P / 2);
                    const double keep = upper ? hi : lo;
                    const double send = upper ? lo : hi;
                    half[j]           = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
                mine = warp_reduce_rows_from<P / 2>(half, lane, 8);
            }

            // combine the eight column slices in warp order
            double* slot = set_exchange + static_cast<int64_t>(parity) * (WPS * 32);
            if (owner)
            {
                slot[wis * 32 + my_pair] = mine;
            }
            named_bar_sync(1 + set, WPS * 32);
            mine = 0.0;
            if (my_pair < PAIRS)
            {
#pragma unroll
                for (int q = 0; q < WPS; ++q)
                {
                    mine += slot[q * 32 + my_pair];
                }
            }
            parity ^= 1;

            // loss + dL of pair (my_r, my_t): one lane per pair (every warp evaluates the same pairs: no second
            // barrier, and the results are bit-identical across the warps)
            const double o     = mine + my_bias; // product first, then + bias (src/linear/util.cpp:19-20)
            const bool   valid = owner && my_r < nrows;
            double       dl    = 0.0;
            if (p.loss == LOSS_CLASSNLL)
            {
                // classnll.h:19-37 over the T lanes of this row (lanes of pair my_r * T + t, t = 0 .. T - 1)
                const int first_lane = (my_r * T) << SHIFT;
                double    omax       = -INFINITY;
#pragma unroll
                for (int t = 0; t < T; ++t)
                {
                    const double other = __shfl_sync(0xffffffffu, o, min(31, first_lane + (t << SHIFT)));
                    omax               = max_of(omax, other);
                }
                const double ex     = valid ? exp(__dsub_rn(o, omax)) : 0.0;
                double       sumexp = 0.0;
#pragma unroll
                for (int t = 0; t < T; ++t)
                {
                    sumexp += __shfl_sync(0xffffffffu, ex, min(31, first_lane + (t << SHIFT)));
                }
                if (valid)
                {
                    // per row: log(sum exp) - 0.5 * sum (1 + y) o + omax; the lane of target 0 carries the row terms
                    fsum += __dmul_rn(-0.5, __dmul_rn(__dadd_rn(1.0, y), o));
                    if (my_t == 0)
                    {
                        fsum += __dadd_rn(log(sumexp), omax);
                    }
                    if (GRAD)
                    {
                        dl = __dsub_rn(__ddiv_rn(ex, sumexp), __dmul_rn(0.5, __dadd_rn(1.0, y)));
                    }
                }
            }
            else if (valid)
            {
                double term;
                loss_term<GRAD>(p.loss, p.loss_param, y, o, term, dl);
                fsum += loss_post(p.loss, term); // the per-sample factor is linear: applied per term
            }

            if constexpr (GRAD)
            {
                gbsum += dl;
                // dL of every pair to every lane through this warp's scratch row
                __syncwarp();
                if ((lane & ((1 << SHIFT) - 1)) == 0)
                {
                    my_dl[lane >> SHIFT] = dl;
                }
                __syncwarp();
#pragma unroll
                for (int q = 0; q < PAIRS; ++q)
                {
                    const int    r   = q / T, t = q % T;
                    const double dlq = my_dl[q];
#pragma unroll
                    for (int k = 0; k < 2 * KC; ++k)
                    {
                        g[t][k] = fma(dlq, x[r][k], g[t][k]);
                    }
                }
            }
        }
    }

    // ----- this CTA's partial row: [loss | sum dL (T) | sum dL x^T (T * D)] -----
    double* out = p.partials + static_cast<int64_t>(blockIdx.x) * (1 + T + static_cast<int64_t>(T) * D);
    if constexpr (NS > 1)
    {
        // both sets cover the same columns: set 1 parks its gradient slices in the (now idle) stage area, set 0 adds
        named_bar_sync(15, kFtWarps * 32); // every issued copy has been consumed
        double* park = reinterpret_cast<double*>(stage_base);
        if (GRAD && set == 1)
        {
#pragma unroll
            for (int t = 0; t < T; ++t)
            {
#pragma unroll
                for (int k = 0; k < 2 * KC; ++k)
                {
                    park[(static_cast<int64_t>(wis) * T * 2 * KC + t * 2 * KC + k) * 32 + lane] = g[t][k];
                }
            }
        }
        if (wis == 0 && (lane & ((1 << SHIFT) - 1)) == 0)
        {
            tail[set * 64 + (lane >> SHIFT)]      = owner ? fsum : 0.0;
            tail[set * 64 + 32 + (lane >> SHIFT)] = owner ? gbsum : 0.0;
        }
        named_bar_sync(15, kFtWarps * 32);
        if (set == 1)
        {
            return;
        }
        if constexpr (GRAD)
        {
#pragma unroll
            for (int t = 0; t < T; ++t)
            {
#pragma unroll
                for (int k = 0; k < 2 * KC; ++k)
                {
                    g[t][k] += park[(static_cast<int64_t>(wis) * T * 2 * KC + t * 2 * KC + k) * 32 + lane];
                }
            }
        }
    }
    else if (warp == 0 && (lane & ((1 << SHIFT) - 1)) == 0)
    {
        tail[lane >> SHIFT]        = owner ? fsum : 0.0;
        tail[32 + (lane >> SHIFT)] = owner ? gbsum : 0.0;
    }
    if constexpr (GRAD)
    {
#pragma unroll
        for (int t = 0; t < T; ++t)
        {
#pragma unroll
            for (int k = 0; k < KC; ++k)
            {
                const int col = cbase + (k * 32 + lane) * 2;
                if (col < D)
                {
                    // 64-bit stores: the gradient block starts 1 + T doubles into the row (not 16-byte aligned)
                    double* dst = out + 1 + T + static_cast<int64_t>(t) * D + col;
                    dst[0]      = g[t][2 * k];
                    dst[1]      = g[t][2 * k + 1];
                }
            }
        }
    }
    if (warp == 0)
    {
        // the warps of a set carry identical loss / dL sums: warp 0 folds the pair slots of every set in order
        __syncwarp();
        if (lane == 0)
        {
            double f = 0.0;
            for (int a = 0; a < NS; ++a)
            {
                for (int q = 0; q < PAIRS; ++q)
                {
                    f += tail[a * 64 + q];
                }
            }
            out[0] = f;
        }
        if (GRAD && lane < T)
        {
            double b = 0.0;
            for (int a = 0; a < NS; ++a)
            {
                for (int r = 0; r < RW; ++r)
                {
                    b += tail[a * 64 + 32 + r * T + lane];
                }
            }
            out[1 + lane] = b;
        }
    }
}

#endif // __CUDACC__
} // namespace nb200
