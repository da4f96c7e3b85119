// vgrad_fl.cuh — a handful of targets (T <= 16) over rows of 256 .. 1024 features: ONE pass over X, both contractions
// on the FP64 tensor pipe, the forward split over FEATURES like vgrad_fd.cuh — and the loss evaluated ONCE, one
// (row, target) pair per lane.
//
// In vgrad_fd.cuh every warp ends the exchange with the whole 8 x TP output tile in accumulator-fragment layout and
// evaluates the loss of four (row, target) pairs per lane — all eight warps the SAME 8 rows, back to back: ncu showed
// 1244 warp instructions per tile and warp for 98 DMMAs, the tensor pipe 42.6 % active and the fixed-latency stall of
// the fp64 chain (exp / log / division) on top (profiles/r02_ncu_full_fd_784x10.csv). Here an ITERATION is TPI m8
// tiles (8 TPI rows, one ring stage per tile), and per iteration
//
//   * forward, split-K: warp w multiplies its column slice of all TPI tiles (the W fragments are loaded once for the
//     TPI tiles), O_w = X[:, slice_w] W[:, slice_w]^T, src/linear/util.cpp:19-20; the partial tiles go to shared
//     memory, ONE named barrier;
//   * loss: warp w owns rows [w TPI, (w + 1) TPI) of the iteration, lane (row, t) sums the eight partial outputs of ITS
//     pair in warp order, adds the bias and evaluates the loss ONCE (softmax: shuffles among the TP lanes of a row):
//     the chain is one element long instead of four, and it is not repeated by the other seven warps; dL meets in a
//     shared tile [8 TPI][kFdLdd], a second named barrier;
//   * backward, split-N: warp w accumulates gW[:, slice_w] += dL^T X[:, slice_w] over the 8 TPI rows in registers for
//     the whole launch (src/linear/function.cpp:68-70), the B fragments straight from the stages, which then go back
//     to the producer.
//
// No predicates in the two contraction loops: stage rows past the data hold zeros or stale (finite) rows — their
// outputs are masked in the loss step and their dL is 0 —, and W rows past T read whatever follows W in shared memory:
// they only reach output columns >= T, which the loss step never looks at (selects, not products).
// Shared memory: barriers | partial outputs [8 warps][8 TPI][TPS] | dL [8 TPI][kFdLdd] | W [T][ld] | stages.
// Same fd_params / launch code / partial rows as vgrad_fd.cuh (reduce_partials_kernel + finish_kernel fold them).
#pragma once

#include "vgrad_fd.cuh"

namespace nb200
{
constexpr int kFlMaxStages = 8;
constexpr int kFlExchange  = 128; // after full [8] | empty [8]

__host__ __device__ constexpr int fl_tps(const int nt8)
{
    return nt8 == 1 ? 8 : 8 * nt8 + 8; // = 8 (mod 16): the 128-bit stores of a quarter warp (rows g, g + 1) do not collide
}
__host__ __device__ constexpr int fl_dl_offset(const int nt8, const int tpi)
{
    // the exchange area doubles as the fold scratch [3][8 warps][32 lanes] at the end of the launch
    const int exchange = kFdWarps * kFdRows * tpi * fl_tps(nt8) * 8;
    return kFlExchange + (exchange < 3 * kFdWarps * 32 * 8 ? 3 * kFdWarps * 32 * 8 : exchange);
}
__host__ __device__ constexpr int fl_header(const int nt8, const int tpi)
{
    return (fl_dl_offset(nt8, tpi) + kFdRows * tpi * kFdLdd * 8 + 127) / 128 * 128;
}

#ifdef __CUDACC__

// sum / max over the TP lanes of a row (aligned groups of TP lanes)
template <int TP>
__device__ __forceinline__ double group_sum(double v)
{
#pragma unroll
    for (int offset = TP / 2; offset > 0; offset >>= 1)
    {
        v += __shfl_xor_sync(0xffffffffu, v, offset);
    }
    return v;
}
template <int TP>
__device__ __forceinline__ double group_max(double v)
{
#pragma unroll
    for (int offset = TP / 2; offset > 0; offset >>= 1)
    {
        const double o = __shfl_xor_sync(0xffffffffu, v, offset);
        v              = (v < o) ? o : v;
    }
    return v;
}

// fd_params: num_tiles = number of ITERATIONS (8 TPI rows each), slice = 8 NS8 columns per warp, ld = 4 (mod 8) >= D,
// stages >= TPI + 1, 512 bytes of slack behind the last stage
template <int NT8, int NS8, int TPI, bool GRAD>
__global__ void __launch_bounds__(kFdThreads, 1) vgrad_fl_kernel(const fd_params p)
{
    constexpr int TP  = 8 * NT8;
    constexpr int TPS = fl_tps(NT8);
    constexpr int RI  = kFdRows * TPI; // rows per iteration
    static_assert(TPI * TP <= 32, "one (row, target) pair per lane");

    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar = full_bar + kFlMaxStages;
    double*   exchange  = reinterpret_cast<double*>(smem + kFlExchange);
    double*   dl_tile   = reinterpret_cast<double*>(smem + fl_dl_offset(NT8, TPI));
    double*   w_s       = reinterpret_cast<double*>(smem + p.w_offset);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int g    = lane >> 2;
    const int tig  = lane & 3;
    const int D    = p.D;
    const int T    = p.T;
    const int ld   = p.ld;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < p.stages; ++s)
        {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, kFdWarps);
        }
        mbar_fence_init();
    }
    // the stages start as zeros (rows that are never copied and the pad columns [D, ld) multiply as zero), then W
    // (zero beyond D columns; only the T rows that exist are stored)
    {
        double2*  z = reinterpret_cast<double2*>(smem + p.stage_offset);
        const int n = p.stages * (p.stage_stride / 16);
        for (int i = threadIdx.x; i < n; i += kFdThreads)
        {
            z[i] = make_double2(0.0, 0.0);
        }
    }
    for (int i = threadIdx.x; i < T * ld; i += kFdThreads)
    {
        const int t = i / ld, c = i - t * ld;
        w_s[i]      = (c < D) ? p.W[static_cast<int64_t>(t) * D + c] : 0.0;
    }
    // make the generic-proxy stores above visible to the async proxy's later writes into the same stage rows
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const int64_t iterations = p.num_tiles;
    const int     ty         = p.y_broadcast ? 1 : T; // targets stored per row of Y

    if (warp >= kFdWarps)
    {
        // ===== producer warpgroup: one warp feeds the ring, one stage per 8-row tile, one bulk copy per row =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(40));
        if (warp != kFdWarps)
        {
            return;
        }
        const uint64_t policy = l2_policy_evict_first();
        int            s      = 0;
        uint32_t       phase  = 1;
        for (int64_t it = blockIdx.x; it < iterations; it += gridDim.x)
        {
            for (int j = 0; j < TPI; ++j)
            {
                mbar_wait(empty_bar + s, phase);
                const int64_t  row0    = (it * TPI + j) * kFdRows;
                const int      rows    = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(kFdRows), p.N - row0)));
                const uint32_t bytes_y = static_cast<uint32_t>((rows * ty * 8 + 15) / 16 * 16);
                unsigned char* stage   = smem + p.stage_offset + static_cast<int64_t>(s) * p.stage_stride;
                if (lane == 0)
                {
                    // (a tile past the last row carries no bytes: the arrival alone completes the phase)
                    mbar_arrive_expect_tx(full_bar + s, static_cast<uint32_t>(rows) * D * 8 + bytes_y);
                    if (rows > 0)
                    {
                        bulk_copy_g2s(stage + p.y_offset, p.Y + row0 * ty, bytes_y, full_bar + s, policy);
                    }
                }
                __syncwarp();
                if (lane < rows)
                {
                    bulk_copy_g2s(stage + static_cast<int64_t>(lane) * ld * 8, p.X + (row0 + lane) * D,
                                  static_cast<uint32_t>(D) * 8, full_bar + s, policy);
                }
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
    const int c0 = warp * p.slice; // first column of this warp's slice (forward: K range, backward: N range)
    // k-steps of this warp in the forward: the shared rows are ld (= 4 mod 8, >= D) doubles long, the columns [D, ld)
    // are zero in X and W, and the slices of the last warps may reach past ld (backward: those columns are never stored;
    // the launch adds 512 bytes of slack behind the last stage for their loads)
    const int kend = max(0, min(2 * NS8, (ld - c0) / 4));

    double gacc[NT8][NS8][2]; // gW[8 mt + g][c0 + 8 ns + 2 tig + {0, 1}]
#pragma unroll
    for (int mt = 0; mt < NT8; ++mt)
    {
#pragma unroll
        for (int ns = 0; ns < NS8; ++ns)
        {
            gacc[mt][ns][0] = gacc[mt][ns][1] = 0.0;
        }
    }
    // the (row, target) pair of this lane in the loss step
    const bool   lane_on = lane < TPI * TP;
    const int    lt      = lane % TP;                 // target
    const int    r_it    = warp * TPI + (lane_on ? lane / TP : 0); // row inside the iteration
    const int    r_tile  = r_it >> 3;                 // ... = its tile (ring stage) of the iteration
    const int    r_g     = r_it & 7;                  // ... and its row inside that tile
    const bool   t_ok    = lane_on && lt < T;
    const double my_bias = t_ok ? p.b[lt] : 0.0;
    double       fsum    = 0.0; // losses of this lane's rows (kept by the lanes with lt == 0)
    log_product  logs;          // softmax: the rows' sums of exponentials, multiplied up (loss.cuh)
    double       gbsum   = 0.0; // sum over this lane's rows of dL[row][lt]
    double       lsum    = 0.0; // per-target loss sums (batched models)

    int      stage_index = 0;
    uint32_t phase       = 0;
    for (int64_t it = blockIdx.x; it < iterations; it += gridDim.x)
    {
        const unsigned char* stage[TPI];
        int                  used[TPI];
#pragma unroll
        for (int j = 0; j < TPI; ++j)
        {
            used[j] = stage_index;
            mbar_wait(full_bar + stage_index, phase);
            stage[j] = smem + p.stage_offset + static_cast<int64_t>(stage_index) * p.stage_stride;
            const uint32_t wrapped = (stage_index + 1 == p.stages) ? 1u : 0u;
            stage_index            = wrapped ? 0 : stage_index + 1;
            phase ^= wrapped;
        }

        // ----- forward, split-K: partial outputs of this warp's column slice for the TPI tiles -----
        // (at least four independent accumulator chains per warp: a DMMA depends on the previous one of its chain)
        {
            constexpr int CH0 = (TPI * NT8 >= 4) ? 1 : 4 / (TPI * NT8);
            constexpr int CH  = CH0;
            double part[TPI][CH][NT8][2];
#pragma unroll
            for (int j = 0; j < TPI; ++j)
            {
#pragma unroll
                for (int c = 0; c < CH; ++c)
                {
#pragma unroll
                    for (int nt = 0; nt < NT8; ++nt)
                    {
                        part[j][c][nt][0] = part[j][c][nt][1] = 0.0;
                    }
                }
            }
            const int     xoff = g * ld + c0 + tig;
            const double* wb   = w_s + xoff;
            int           kk   = 0;
#pragma unroll 4
            for (; kk + CH <= kend; kk += CH) // this warp's k-steps, CH at a time
            {
#pragma unroll
                for (int c = 0; c < CH; ++c)
                {
                    double a[TPI];
#pragma unroll
                    for (int j = 0; j < TPI; ++j)
                    {
                        a[j] = reinterpret_cast<const double*>(stage[j])[xoff + (kk + c) * 4];
                    }
#pragma unroll
                    for (int nt = 0; nt < NT8; ++nt)
                    {
                        const double bv = wb[nt * 8 * ld + (kk + c) * 4];
#pragma unroll
                        for (int j = 0; j < TPI; ++j)
                        {
                            dmma_m8n8k4(part[j][c][nt][0], part[j][c][nt][1], a[j], bv);
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CH - 1; ++c) // (ld / 4 is odd: the last warp's range does not split evenly)
            {
                if (kk + c < kend)
                {
#pragma unroll
                    for (int nt = 0; nt < NT8; ++nt)
                    {
                        const double bv = wb[nt * 8 * ld + (kk + c) * 4];
#pragma unroll
                        for (int j = 0; j < TPI; ++j)
                        {
                            dmma_m8n8k4(part[j][c][nt][0], part[j][c][nt][1],
                                        reinterpret_cast<const double*>(stage[j])[xoff + (kk + c) * 4], bv);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < TPI; ++j)
            {
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt)
                {
                    double s0 = part[j][0][nt][0], s1 = part[j][0][nt][1];
#pragma unroll
                    for (int c = 1; c < CH; ++c) // chain order: fixed
                    {
                        s0 += part[j][c][nt][0];
                        s1 += part[j][c][nt][1];
                    }
                    *reinterpret_cast<double2*>(exchange + (warp * RI + j * kFdRows + g) * TPS + nt * 8 + tig * 2) =
                        make_double2(s0, s1);
                }
            }
        }
        named_bar_sync(1, kFdWarps * 32);

        // ----- loss + dL, one (row, target) pair per lane -----
        {
            const bool row_ok = lane_on && (it * TPI) * kFdRows + r_it < p.N;
            const bool el_ok  = row_ok && lt < T;
            double     o      = 0.0;
            if (lane_on)
            {
#pragma unroll
                for (int q = 0; q < kFdWarps; ++q) // warp order: fixed
                {
                    o += exchange[(q * RI + r_it) * TPS + lt];
                }
                o += my_bias; // product first, then + bias (src/linear/util.cpp:19-20)
            }
            const double* ys = reinterpret_cast<const double*>(stage[0] + p.y_offset);
#pragma unroll
            for (int j = 1; j < TPI; ++j)
            {
                ys = (r_tile == j) ? reinterpret_cast<const double*>(stage[j] + p.y_offset) : ys;
            }
            const double y  = el_ok ? (p.y_broadcast ? ys[r_g] : ys[r_g * T + lt]) : 0.0;
            double       dl = 0.0;
            if (p.loss == LOSS_CLASSNLL && p.y_broadcast == 0) // (batched single-output models: elementwise)
            {
                // classnll.h:19-37: the TP lanes of a group hold the whole row
                const double omax   = group_max<TP>(el_ok ? o : -INFINITY);
                const double ex     = exp_nonpositive(el_ok ? __dsub_rn(o, omax) : -INFINITY); // (straight-line)
                const double sumexp = group_sum<TP>(ex);
                const double dot    = group_sum<TP>(el_ok ? __dmul_rn(__dadd_rn(1.0, y), o) : 0.0);
                // (the logarithm of the value is taken once per kLogEvery rows: loss.cuh, log_product)
                const double grad = __dsub_rn(__ddiv_rn(ex, sumexp), __dmul_rn(0.5, __dadd_rn(1.0, y)));
                fsum += (row_ok && lt == 0) ? __dsub_rn(omax, __dmul_rn(0.5, dot)) : 0.0;
                logs.product *= (row_ok && lt == 0) ? sumexp : 1.0;
                log_product_flush(logs, 1, fsum);
                dl = (GRAD && el_ok) ? grad : 0.0;
            }
            else
            {
                double term = 0.0;
                if (el_ok)
                {
                    loss_term<GRAD>(p.loss, p.loss_param, y, o, term, dl);
                    lsum += loss_post(p.loss, term); // one output per model: the factor applies per term
                }
                const double rowsum = group_sum<TP>(term);
                if (row_ok && lt == 0)
                {
                    fsum += loss_post(p.loss, rowsum);
                }
            }
            if (GRAD && lane_on)
            {
                gbsum += dl;
                dl_tile[r_it * kFdLdd + lt] = dl; // rows past the data and targets past T carry dL = 0
            }
        }
        named_bar_sync(1, kFdWarps * 32); // dL is complete; the partial outputs may be overwritten

        if constexpr (GRAD)
        {
            // ----- backward, split-N: gW[:, slice] += dL^T X[:, slice]; k runs over the 8 TPI rows -----
#pragma unroll
            for (int kk = 0; kk < 2 * TPI; ++kk)
            {
                const int r = kk * 4 + tig; // row inside the iteration
                double    a[NT8];
#pragma unroll
                for (int mt = 0; mt < NT8; ++mt)
                {
                    a[mt] = dl_tile[r * kFdLdd + mt * 8 + g];
                }
                const double* src = reinterpret_cast<const double*>(stage[kk / 2]) + ((kk & 1) * 4 + tig) * ld + c0 + g;
#pragma unroll
                for (int ns = 0; ns < NS8; ++ns)
                {
                    const double bv = src[ns * 8];
#pragma unroll
                    for (int mt = 0; mt < NT8; ++mt)
                    {
                        dmma_m8n8k4(gacc[mt][ns][0], gacc[mt][ns][1], a[mt], bv);
                    }
                }
            }
        }
        // the stages go back behind one cross-proxy fence (common.cuh, mbar_release_stage)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0)
        {
#pragma unroll
            for (int j = 0; j < TPI; ++j)
            {
                mbar_arrive(empty_bar + used[j]);
            }
        }
    }

    // ----- this CTA's partial row: [loss | sum dL (T) | sum dL x^T (T * D)] -----
    const bool batched = p.partials_gw != nullptr;
    double*    out     = p.partials + static_cast<int64_t>(blockIdx.x) *
                                   (batched ? (1 + 2 * T) : (1 + T + static_cast<int64_t>(T) * D));
    double*    out_gw  = batched ? p.partials_gw + static_cast<int64_t>(blockIdx.x) * T * D : out + 1 + T;

    if constexpr (GRAD)
    {
#pragma unroll
        for (int mt = 0; mt < NT8; ++mt)
        {
            const int t = mt * 8 + g;
            if (t < T)
            {
#pragma unroll
                for (int ns = 0; ns < NS8; ++ns)
                {
                    const int col = c0 + ns * 8 + tig * 2;
                    if (col < D)
                    {
                        double* dst = out_gw + static_cast<int64_t>(t) * D + col; // 8-byte aligned only
                        dst[0]      = gacc[mt][ns][0];
                        dst[1]      = gacc[mt][ns][1];
                    }
                }
            }
        }
    }
    // loss / dL / per-target loss sums: every lane parks its sums, warp 0 folds them in (warp, row) order
    // (the last barrier of the loop is behind every read of the exchange area)
    fsum += log(logs.product); // (1 when nothing is pending)
    double* fold = exchange;   // [3][8 warps][32 lanes]
    fold[0 * kFdWarps * 32 + warp * 32 + lane] = (lane_on && lt == 0) ? fsum : 0.0;
    fold[1 * kFdWarps * 32 + warp * 32 + lane] = lane_on ? gbsum : 0.0;
    fold[2 * kFdWarps * 32 + warp * 32 + lane] = lane_on ? lsum : 0.0;
    named_bar_sync(1, kFdWarps * 32);
    if (warp == 0 && lane < TP)
    {
        double f = 0.0, gb = 0.0, ls = 0.0;
        for (int q = 0; q < kFdWarps; ++q)
        {
#pragma unroll
            for (int rw = 0; rw < TPI; ++rw)
            {
                f += fold[0 * kFdWarps * 32 + q * 32 + rw * TP];
                gb += fold[1 * kFdWarps * 32 + q * 32 + rw * TP + lane];
                ls += fold[2 * kFdWarps * 32 + q * 32 + rw * TP + lane];
            }
        }
        if (lane == 0)
        {
            out[0] = f;
        }
        if (lane < T)
        {
            if (GRAD || batched)
            {
                out[1 + lane] = GRAD ? gb : 0.0;
            }
            if (batched)
            {
                out[1 + T + lane] = ls;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// vgrad_fp_kernel — the same three steps as a PIPELINE over iterations of one 8-row tile, the loss on its OWN warps,
// no CTA-wide barrier in the loop.
//
// In vgrad_fl_kernel the tensor pipe idles while the loss chain runs: ~1.3 us per iteration whatever the width (two
// barriers, eight dependent additions, max / exp / sum / log / division, all latency). Moving the backward product of
// iteration i - 1 next to the loss of iteration i on the SAME warps did not help (measured: every warp still runs
// forward + backward + loss back to back, and with two warps per sub-partition nothing fills the pipe while both are
// in their chains). Here the eight consumer warps only multiply, and 2 NT8 more warps evaluate the losses next to them,
// one (row, target) pair per lane (the chain of one pair takes ~1.3 us, of two pairs per lane ~2.3 us, measured):
//
//   consumer warp, iteration i:  forward(i) -> partial outputs to shared memory (behind part_free(i - 1)) -> arrive on
//       part_ready -> backward(i - 1): A fragments from the dL tile of i - 1 (double-buffered, dl_ready(i - 1)), B
//       fragments = xb, this warp's column slice of the rows of i - 1, in registers -> xb(i) from the stage of i, which
//       goes back to the producer (a stage is held for one iteration only);
//   loss warp l (rows l 8 / LW ... of the tile), iteration i:  wait part_ready(i) -> lane (row, t) sums the eight
//       partial outputs of its (row, target) pair in warp order, reads its target from the stage -> arrive on
//       part_free and on the stage's empty barrier -> loss + dL (softmax: shuffles among the TP lanes of a row) -> dL
//       tile -> arrive on dl_ready. The softmax value (a logarithm per row) is not on the way to dL: the loss warps
//       leave (sum exp, dot, max) per row to one more warp, which turns them into the loss behind their backs.
//
// Every synchronisation is an mbarrier with the arrival early and the wait late, one phase per iteration (parity =
// iteration parity; every wait precedes the waiting warp's own next arrival on a barrier that could otherwise run two
// phases ahead). The dL buffer of i is rewritten in loss(i + 2), behind part_ready(i + 2), i.e. behind every
// consumer's backward(i) in program order; the partials are single-buffered behind part_free.
// Registers: consumers 208 / 192 (gacc + xb), the other warps 80 / 64 (setmaxnreg; the pool is what the launch
// allocated — a request beyond it blocks for ever).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kFpBars      = 2 * kFlMaxStages; // part_ready, part_free, dl_ready follow full [8] | empty [8]
constexpr int kFpExchange  = 256;
// loss warps: one (row, target) pair per lane, so 2 NT8 of them cover the 8 x 8 NT8 pairs of a tile; with two target
// tiles the CTA has a fourth warpgroup (one loss warp of it is used)
__host__ __device__ constexpr int fp_loss_warps(const int nt8)
{
    return 2 * nt8;
}
__host__ __device__ constexpr int fp_threads(const int nt8)
{
    return nt8 == 1 ? kFdThreads : 512; // (whole warpgroups: ptxas sizes the launch allocation for them)
}

__host__ __device__ constexpr int fp_dl_offset(const int nt8)
{
    return fl_dl_offset(nt8, 1) - kFlExchange + kFpExchange;
}
__host__ __device__ constexpr int fp_header(const int nt8)
{
    return (fp_dl_offset(nt8) + 2 * kFdRows * kFdLdd * 8 + kFdRows * 4 * 8 + 127) / 128 * 128;
}

template <int NT8, int NS8, bool GRAD>
__global__ void __launch_bounds__(fp_threads(NT8), 1) vgrad_fp_kernel(const fd_params p)
{
    constexpr int TP      = 8 * NT8;
    constexpr int TPS     = fl_tps(NT8);
    constexpr int LW      = fp_loss_warps(NT8);
    constexpr int RPP     = 32 / TP; // rows per loss warp: LW x RPP = 8
    constexpr int THREADS = fp_threads(NT8);
    // setmaxnreg moves registers inside the pool the launch allocated (THREADS x the launch count): 384 x 168 >=
    // 256 x 208 + 128 x 80, 512 x 128 = 256 x 192 + 256 x 64 — a request beyond the pool would block for ever
    constexpr int CONSUMER_REGS = NT8 == 1 ? 208 : 192;
    constexpr int OTHER_REGS    = NT8 == 1 ? 80 : 64;
    static_assert(LW * RPP == kFdRows && NT8 <= 2, "one (row, target) pair per lane of the loss warps");

    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar   = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar  = full_bar + kFlMaxStages;
    uint64_t* part_ready = full_bar + kFpBars;
    uint64_t* part_free  = part_ready + 1;
    uint64_t* dl_ready    = part_ready + 2;
    uint64_t* stats_ready = part_ready + 3;
    uint64_t* stats_free  = part_ready + 4;
    double*   exchange    = reinterpret_cast<double*>(smem + kFpExchange);          // [8 warps][8][TPS]
    double*   dl_tiles    = reinterpret_cast<double*>(smem + fp_dl_offset(NT8));    // [2][8][kFdLdd]
    double*   row_stats   = dl_tiles + 2 * kFdRows * kFdLdd;                        // [8][4]: sum exp, dot, max of a row
    double*   w_s        = reinterpret_cast<double*>(smem + p.w_offset);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int g    = lane >> 2;
    const int tig  = lane & 3;
    const int D    = p.D;
    const int T    = p.T;
    const int ld   = p.ld;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < p.stages; ++s)
        {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, kFdWarps + LW);
        }
        mbar_init(part_ready, kFdWarps);
        mbar_init(part_free, LW);
        mbar_init(dl_ready, LW);
        mbar_init(stats_ready, LW);
        mbar_init(stats_free, 1);
        mbar_fence_init();
    }
    {
        double2*  z = reinterpret_cast<double2*>(smem + p.stage_offset);
        const int n = p.stages * (p.stage_stride / 16);
        for (int i = threadIdx.x; i < n; i += THREADS)
        {
            z[i] = make_double2(0.0, 0.0);
        }
    }
    for (int i = threadIdx.x; i < T * ld; i += THREADS)
    {
        const int t = i / ld, c = i - t * ld;
        w_s[i]      = (c < D) ? p.W[static_cast<int64_t>(t) * D + c] : 0.0;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    const int64_t iterations = p.num_tiles;
    const int     ty         = p.y_broadcast ? 1 : T;
    const bool    batched    = p.partials_gw != nullptr;
    double*       out        = p.partials + static_cast<int64_t>(blockIdx.x) *
                                         (batched ? (1 + 2 * T) : (1 + T + static_cast<int64_t>(T) * D));
    double*       out_gw     = batched ? p.partials_gw + static_cast<int64_t>(blockIdx.x) * T * D : out + 1 + T;

    if (warp >= kFdWarps)
    {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(OTHER_REGS));
        if (warp == kFdWarps)
        {
            // ===== producer warp: feeds the ring, one stage per 8-row tile, one bulk copy per row =====
            const uint64_t policy = l2_policy_evict_first();
            int            s      = 0;
            uint32_t       phase  = 1;
            for (int64_t it = blockIdx.x; it < iterations; it += gridDim.x)
            {
                mbar_wait(empty_bar + s, phase);
                const int64_t  row0    = it * kFdRows;
                const int      rows    = static_cast<int>(min(static_cast<int64_t>(kFdRows), p.N - row0));
                const uint32_t bytes_y = static_cast<uint32_t>((rows * ty * 8 + 15) / 16 * 16);
                unsigned char* stage   = smem + p.stage_offset + static_cast<int64_t>(s) * p.stage_stride;
                if (lane == 0)
                {
                    mbar_arrive_expect_tx(full_bar + s, static_cast<uint32_t>(rows) * D * 8 + bytes_y);
                    bulk_copy_g2s(stage + p.y_offset, p.Y + row0 * ty, bytes_y, full_bar + s, policy);
                }
                __syncwarp();
                if (lane < rows)
                {
                    bulk_copy_g2s(stage + static_cast<int64_t>(lane) * ld * 8, p.X + (row0 + lane) * D,
                                  static_cast<uint32_t>(D) * 8, full_bar + s, policy);
                }
                if (++s == p.stages)
                {
                    s = 0;
                    phase ^= 1u;
                }
            }
            return;
        }
        const int  lw      = warp - kFdWarps - 1; // loss warp
        const bool softmax = p.loss == LOSS_CLASSNLL && p.y_broadcast == 0; // (batched single-output models: elementwise)
        if (lw > LW)
        {
            return;
        }
        double* fold = exchange; // [3][LW warps][32 lanes] + [8]: where the sums of this CTA meet at the end
        if (lw == LW)
        {
            // ===== value warp (softmax only): lane r < 8 turns the row statistics the loss warps leave behind into
            // the loss of row r — classnll.h:19-37: log(sum exp(o - max)) - 0.5 (1 + y) . o + max. The logarithm is not
            // on the way to dL, so it does not sit in the loss warps' chain =====
            double   fsum = 0.0;
            uint32_t par  = 0;
            for (int64_t it = blockIdx.x; softmax && it < iterations; it += gridDim.x)
            {
                mbar_wait(stats_ready, par);
                double sumexp = 1.0, dot = 0.0, omax = 0.0;
                if (lane < kFdRows)
                {
                    sumexp = row_stats[lane * 4 + 0];
                    dot    = row_stats[lane * 4 + 1];
                    omax   = row_stats[lane * 4 + 2];
                }
                __syncwarp();
                if (lane == 0)
                {
                    mbar_arrive(stats_free);
                }
                const double value = __dadd_rn(__dsub_rn(log(sumexp), __dmul_rn(0.5, dot)), omax);
                fsum += (lane < kFdRows && it * kFdRows + lane < p.N) ? value : 0.0;
                par ^= 1u;
            }
            named_bar_sync(2, (LW + 1) * 32);
            if (lane < kFdRows)
            {
                fold[3 * LW * 32 + lane] = fsum;
            }
            named_bar_sync(2, (LW + 1) * 32);
            return;
        }
        // ===== loss warps: rows RPP lw .. RPP lw + RPP - 1 of every tile, one (row, target) pair per lane =====
        const int    lt      = lane % TP;
        const int    lr      = lw * RPP + lane / TP;
        const double my_bias = lt < T ? p.b[lt] : 0.0;
        double       fsum    = 0.0; // losses of this lane's rows (kept by the lanes with lt == 0; elementwise losses)
        double       gbsum   = 0.0; // sum over this lane's rows of dL[row][lt]
        double       lsum    = 0.0; // per-target loss sums (batched models)
        int          s       = 0;
        uint32_t     phase   = 0;
        uint32_t     par     = 0;
        for (int64_t it = blockIdx.x; it < iterations; it += gridDim.x)
        {
            mbar_wait(full_bar + s, phase); // (the targets of the tile: this warp reads them itself)
            const double* ys = reinterpret_cast<const double*>(smem + p.stage_offset +
                                                               static_cast<int64_t>(s) * p.stage_stride + p.y_offset);
            mbar_wait(part_ready, par);
            const bool row_ok = it * kFdRows + lr < p.N;
            const bool el_ok  = row_ok && lt < T;
            double     o      = 0.0;
#pragma unroll
            for (int q = 0; q < kFdWarps; ++q) // warp order: fixed
            {
                o += exchange[(q * kFdRows + lr) * TPS + lt];
            }
            o += my_bias; // product first, then + bias (src/linear/util.cpp:19-20)
            const double y = el_ok ? (p.y_broadcast ? ys[lr] : ys[lr * T + lt]) : 0.0;
            // the partials and the stage's targets have been read
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0)
            {
                mbar_arrive(part_free);
                mbar_arrive(empty_bar + s);
            }
            double dl = 0.0, term = 0.0;
            if (softmax)
            {
                // the TP lanes of a group hold the whole row; dL = softmax - 0.5 (1 + y)
                const double omax   = group_max<TP>(el_ok ? o : -INFINITY);
                const double ex     = exp_nonpositive(el_ok ? __dsub_rn(o, omax) : -INFINITY); // (straight-line)
                const double sumexp = group_sum<TP>(ex);
                const double grad   = __dsub_rn(__ddiv_rn(ex, sumexp), __dmul_rn(0.5, __dadd_rn(1.0, y)));
                dl                  = (GRAD && el_ok) ? grad : 0.0;
                term                = el_ok ? __dmul_rn(__dadd_rn(1.0, y), o) : 0.0; // summed below: the dot product
                if (lt == 0)
                {
                    mbar_wait(stats_free, par ^ 1u); // the value warp has read the previous tile's statistics
                    row_stats[lr * 4 + 0] = sumexp;
                    row_stats[lr * 4 + 2] = omax;
                }
            }
            else if (el_ok)
            {
                loss_term<GRAD>(p.loss, p.loss_param, y, o, term, dl);
                lsum += loss_post(p.loss, term); // one output per model: the factor applies per term
            }
            if constexpr (GRAD)
            {
                gbsum += dl; // rows past the data and targets past T carry dL = 0
                dl_tiles[par * (kFdRows * kFdLdd) + lr * kFdLdd + lt] = dl;
                __syncwarp();
                if (lane == 0)
                {
                    mbar_arrive(dl_ready);
                }
            }
            // (behind the arrival: the consumers do not wait for the value)
            const double rowsum = group_sum<TP>(term);
            if (softmax)
            {
                if (lt == 0)
                {
                    row_stats[lr * 4 + 1] = rowsum;
                }
                __syncwarp();
                if (lane == 0)
                {
                    mbar_arrive(stats_ready);
                }
            }
            else
            {
                fsum += (row_ok && lt == 0) ? loss_post(p.loss, rowsum) : 0.0;
            }
            if (++s == p.stages)
            {
                s = 0;
                phase ^= 1u;
            }
            par ^= 1u;
        }
        // [loss | sum dL (T)] (+ per-target loss sums) of this CTA: the loss warps and the value warp meet in the
        // exchange area once all have read the last partials (the consumers do not write it again)
        named_bar_sync(2, (LW + 1) * 32);
        fold[(0 * LW + lw) * 32 + lane] = (lt == 0) ? fsum : 0.0;
        fold[(1 * LW + lw) * 32 + lane] = gbsum;
        fold[(2 * LW + lw) * 32 + lane] = lsum;
        named_bar_sync(2, (LW + 1) * 32);
        if (lw == 0 && lane < TP)
        {
            double f = 0.0, gb = 0.0, ls = 0.0;
#pragma unroll
            for (int q = 0; q < LW; ++q)
            {
#pragma unroll
                for (int rr = 0; rr < RPP; ++rr)
                {
                    f += fold[(0 * LW + q) * 32 + rr * TP];
                    gb += fold[(1 * LW + q) * 32 + rr * TP + lane];
                    ls += fold[(2 * LW + q) * 32 + rr * TP + lane];
                }
            }
#pragma unroll
            for (int r = 0; r < kFdRows; ++r) // (the softmax values, row order)
            {
                f += fold[3 * LW * 32 + r];
            }
            if (lane == 0)
            {
                out[0] = f;
            }
            if (lane < T)
            {
                if (GRAD || batched)
                {
                    out[1 + lane] = GRAD ? gb : 0.0;
                }
                if (batched)
                {
                    out[1 + T + lane] = ls;
                }
            }
        }
        return;
    }

    // ===== consumer warps: the two contractions =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS));
    const int c0   = warp * p.slice;
    const int kend = max(0, min(2 * NS8, (ld - c0) / 4)); // (see vgrad_fl_kernel)

    double gacc[NT8][NS8][2]; // gW[8 mt + g][c0 + 8 ns + 2 tig + {0, 1}]
    double xb[2][NS8];        // B fragments of the backward product: X[4 kk + tig][c0 + 8 ns + g] of the previous tile
#pragma unroll
    for (int mt = 0; mt < NT8; ++mt)
    {
#pragma unroll
        for (int ns = 0; ns < NS8; ++ns)
        {
            gacc[mt][ns][0] = gacc[mt][ns][1] = 0.0;
        }
    }
    auto backward_previous = [&](const uint32_t parity) {
        // (ptxas orders the 2 NT8 NS8 DMMAs by A fragment whatever the source order is)
        const double* dl_prev = dl_tiles + parity * (kFdRows * kFdLdd);
        double        a[2][NT8];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
        {
#pragma unroll
            for (int mt = 0; mt < NT8; ++mt)
            {
                a[kk][mt] = dl_prev[(kk * 4 + tig) * kFdLdd + mt * 8 + g];
            }
        }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
        {
#pragma unroll
            for (int ns = 0; ns < NS8; ++ns)
            {
#pragma unroll
                for (int mt = 0; mt < NT8; ++mt)
                {
                    dmma_m8n8k4(gacc[mt][ns][0], gacc[mt][ns][1], a[kk][mt], xb[kk][ns]);
                }
            }
        }
    };

    int      stage_index = 0;
    uint32_t phase       = 0;
    uint32_t par         = 0; // parity of this CTA's iteration count
    bool     have_prev   = false;
    for (int64_t it = blockIdx.x; it < iterations; it += gridDim.x)
    {
        const int s = stage_index;
        mbar_wait(full_bar + s, phase);
        const double*  xs      = reinterpret_cast<const double*>(smem + p.stage_offset + static_cast<int64_t>(s) * p.stage_stride);
        const uint32_t wrapped = (stage_index + 1 == p.stages) ? 1u : 0u;
        stage_index            = wrapped ? 0 : stage_index + 1;
        phase ^= wrapped;

        // ----- forward, split-K: partial outputs of this warp's column slice -----
        {
            constexpr int CH = 4 / NT8; // four independent accumulator chains per warp
            double        part[CH][NT8][2];
#pragma unroll
            for (int c = 0; c < CH; ++c)
            {
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt)
                {
                    part[c][nt][0] = part[c][nt][1] = 0.0;
                }
            }
            const double* xa = xs + g * ld + c0 + tig;
            const double* wb = w_s + g * ld + c0 + tig;
            int           kk = 0;
#pragma unroll 4
            for (; kk + CH <= kend; kk += CH)
            {
#pragma unroll
                for (int c = 0; c < CH; ++c)
                {
                    const double a = xa[(kk + c) * 4];
#pragma unroll
                    for (int nt = 0; nt < NT8; ++nt)
                    {
                        dmma_m8n8k4(part[c][nt][0], part[c][nt][1], a, wb[nt * 8 * ld + (kk + c) * 4]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CH - 1; ++c) // (ld / 4 is odd: the last warp's range does not split evenly)
            {
                if (kk + c < kend)
                {
                    const double a = xa[(kk + c) * 4];
#pragma unroll
                    for (int nt = 0; nt < NT8; ++nt)
                    {
                        dmma_m8n8k4(part[c][nt][0], part[c][nt][1], a, wb[nt * 8 * ld + (kk + c) * 4]);
                    }
                }
            }
            mbar_wait(part_free, par ^ 1u); // the loss warps have read the partials of the previous iteration
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
                double s0 = part[0][nt][0], s1 = part[0][nt][1];
#pragma unroll
                for (int c = 1; c < CH; ++c) // chain order: fixed
                {
                    s0 += part[c][nt][0];
                    s1 += part[c][nt][1];
                }
                *reinterpret_cast<double2*>(exchange + (warp * kFdRows + g) * TPS + nt * 8 + tig * 2) = make_double2(s0, s1);
            }
        }
        // the dL of the previous iteration is waited for BEFORE the loss warps are let into this one: they cannot
        // complete a second phase of dl_ready while a consumer still waits for the first
        if (GRAD && have_prev)
        {
            mbar_wait(dl_ready, par ^ 1u);
        }
        __syncwarp();
        if (lane == 0)
        {
            mbar_arrive(part_ready);
        }
        if constexpr (GRAD)
        {
            // ----- backward of the previous tile, then this tile's rows -> registers -----
            if (have_prev)
            {
                backward_previous(par ^ 1u);
            }
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
            {
                const double* src = xs + (kk * 4 + tig) * ld + c0 + g;
#pragma unroll
                for (int ns = 0; ns < NS8; ++ns)
                {
                    xb[kk][ns] = src[ns * 8];
                }
            }
        }
        mbar_release_stage(empty_bar + s, lane);
        have_prev = true;
        par ^= 1u;
    }
    if constexpr (GRAD)
    {
        if (have_prev)
        {
            mbar_wait(dl_ready, par ^ 1u);
            backward_previous(par ^ 1u);
        }
#pragma unroll
        for (int mt = 0; mt < NT8; ++mt)
        {
            const int t = mt * 8 + g;
            if (t < T)
            {
#pragma unroll
                for (int ns = 0; ns < NS8; ++ns)
                {
                    const int col = c0 + ns * 8 + tig * 2;
                    if (col < D)
                    {
                        double* dst = out_gw + static_cast<int64_t>(t) * D + col; // 8-byte aligned only
                        dst[0]      = gacc[mt][ns][0];
                        dst[1]      = gacc[mt][ns][1];
                    }
                }
            }
        }
    }
}

#endif // __CUDACC__
} // namespace nb200
