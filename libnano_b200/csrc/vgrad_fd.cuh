// vgrad_fd.cuh — a HANDFUL of targets (5 <= T <= 16): ONE pass over X with both contractions on the FP64 tensor pipe.
//
// vgrad_ft.cuh reads X once but pays 4 T CUDA-core FMAs per element and a warp-wide reduction per 3-4 rows: beyond
// T ~ 6 it is bound by instruction issue. vgrad_mt.cuh uses DMMA but reads X twice in 128-byte column strips and
// round-trips dL through HBM. This kernel keeps the one-pass structure of vgrad_ft.cuh and does the two contractions
// of linear::function_t::do_eval (outputs = X W^T + b, src/linear/util.cpp:19-20; gW += dL^T X,
// src/linear/function.cpp:68-70) with mma.sync m8n8k4 f64 (SASS DMMA.8x8x4) on a tile of 8 whole rows that sits in
// shared memory:
//
//   * producer warp: per tile 8 bulk copies (one per row, into rows padded to a stride = 4 mod 16 doubles so that every
//     64-bit fragment load below is bank-conflict free) + the 8 x T targets, behind full / empty mbarriers;
//   * W (T x D) lives in shared memory for the whole launch with the same padding;
//   * forward: the eight consumer warps split the FEATURES (split-K): warp w multiplies its column slice,
//     O_w[8, TP] = X[:, slice] W[:, slice]^T, the partial tiles are exchanged through double-buffered shared-memory
//     slots behind ONE CTA-wide named barrier and summed in warp order — every warp then holds the outputs of the tile
//     in accumulator-fragment layout (lane (g, tig): row g, targets 8 nt + 2 tig + {0, 1});
//   * loss + dL straight from the fragments (softmax: the four lanes of a quad hold a whole row), every warp evaluates
//     the same tile (bit-identical), dL goes through a per-warp scratch tile to the A-fragment layout;
//   * backward: warp w accumulates gW[:, slice] += dL^T X[:, slice] in registers (NT8 x slice / 8 accumulator tiles)
//     for the whole launch — split-N, no cross-warp traffic; the stage is released when the tile is done;
//   * per-CTA partial rows [loss | sum dL | sum dL x^T] are folded by reduce_partials_kernel / finish_kernel.
#pragma once

#include "common.cuh"
#include "loss.cuh"
#include "vgrad_mt.cuh" // dmma_m8n8k4, quad_sum, quad_max

namespace nb200
{
constexpr int kFdWarps         = 8;
constexpr int kFdProducerWarps = 4;
constexpr int kFdThreads       = (kFdWarps + kFdProducerWarps) * 32;
constexpr int kFdRows          = 8;   // rows per tile: one m8 tile
constexpr int kFdMaxStages     = 4;
constexpr int kFdLdd           = 20;  // row stride of the per-warp dL tile [8][kFdLdd] (= 4 mod 16)

struct fd_params
{
    const double* X;        // [N, D] row-major
    const double* Y;        // [N, T], padded by >= 16 bytes at the end
    const double* W;        // [T, D] device
    const double* b;        // [T] device
    double*       partials; // [grid, 1 + T + T * D]: per-CTA rows [loss | sum dL (T) | sum dL x^T (T * D)]
    // batched single-target models (nb200_vgrad_batch): the K weight vectors play the targets, every
    // output column is compared with the same y_i, losses are kept per column, and the partial rows are split in two
    // buffers: partials = [grid, 1 + 2 T] rows [total | sum dL (T) | loss sums (T)], partials_gw = [grid, T * D]
    double*       partials_gw; // nullptr: the single-buffer layout above
    int           y_broadcast; // 1: Y is [N, 1]
    int64_t       N;
    int64_t       num_tiles;
    int           D;      // even
    int           T;      // <= 8 * NT8; 8 * T even is always true (the Y tile starts on a 16-byte boundary)
    int           slice;  // columns per warp: a multiple of 8, kFdWarps * slice >= D
    int           ld;     // shared row stride of X and W in doubles: >= kFdWarps * slice, = 4 mod 16
    int           stages;
    int           stage_stride; // bytes
    int           y_offset;     // byte offset of the Y tile inside a stage
    int           w_offset;     // byte offset of the W tile in shared memory
    int           stage_offset; // byte offset of stage 0
    int           loss;
    double        loss_param;
};

// shared memory header: barriers [0, 128) | exchange [2][8 warps][8][TPS] doubles | per-warp dL tiles [8][8][kFdLdd]
template <int NT8>
struct fd_layout
{
    static constexpr int TP        = 8 * NT8;
    // row stride of the exchange tiles: 128-bit accesses go a quarter warp (rows g, g + 1) at a time, so the stride
    // must be 4 (mod 8) in 16-byte units
    static constexpr int TPS       = (NT8 == 1) ? 8 : TP + 8;
    static constexpr int kExchange = 128;
    static constexpr int kDl       = kExchange + 2 * kFdWarps * kFdRows * TPS * 8;
    static constexpr int kHeader   = (kDl + kFdWarps * kFdRows * kFdLdd * 8 + 127) / 128 * 128;
};

#ifdef __CUDACC__

// NS8: n8 column tiles per warp slice (slice = 8 * NS8); NSETS: 1, or 2 sets of four warps that take the tiles in
// turn (one set multiplies while the other is between its barrier and the end of its loss chain)
template <int NT8, int NS8, int NSETS, bool GRAD>
__global__ void __launch_bounds__(kFdThreads, 1) vgrad_fd_kernel(const fd_params p)
{
    constexpr int WPS = kFdWarps / NSETS; // warps per set: together they cover all columns
    static_assert(NSETS == 1 || NSETS == 2, "one or two alternating warp sets");
    using L          = fd_layout<NT8>;
    constexpr int TP  = L::TP;
    constexpr int TPS = L::TPS;

    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar = full_bar + kFdMaxStages;
    double*   exchange  = reinterpret_cast<double*>(smem + L::kExchange);
    double*   dl_tiles  = reinterpret_cast<double*>(smem + L::kDl);
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
    // W -> shared memory (zero beyond T rows / D columns), and zero the pad columns of every stage row once: the bulk
    // copies write exactly D doubles per row, so whatever the contractions read beyond D stays zero
    for (int i = threadIdx.x; i < T * ld; i += kFdThreads) // only the T rows that exist are stored
    {
        const int t = i / ld, c = i - t * ld;
        w_s[i]      = (c < D) ? p.W[static_cast<int64_t>(t) * D + c] : 0.0;
    }
    for (int s = 0; s < p.stages; ++s)
    {
        double* xs = reinterpret_cast<double*>(smem + p.stage_offset + static_cast<int64_t>(s) * p.stage_stride);
        for (int i = threadIdx.x; i < kFdRows * (ld - D); i += kFdThreads)
        {
            const int r = i / (ld - D), c = D + (i - r * (ld - D));
            xs[r * ld + c] = 0.0;
        }
    }
    // make the generic-proxy stores above visible to the async proxy's later writes into the same stage rows
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp >= kFdWarps)
    {
        // ===== producer warpgroup: one warp feeds the ring, one bulk copy per row =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(40));
        if (warp != kFdWarps)
        {
            return;
        }
        const uint64_t policy = l2_policy_evict_first();
        int            s      = 0;
        uint32_t       phase  = 1;
        for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
        {
            mbar_wait(empty_bar + s, phase);
            const int64_t  row0    = tile * kFdRows;
            const int      rows    = static_cast<int>(min(static_cast<int64_t>(kFdRows), p.N - row0));
            const int      ty      = p.y_broadcast ? 1 : T; // targets stored per row of Y
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

    // ===== consumer warps =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(232));
    const int set = warp / WPS;
    const int wis = warp % WPS;    // warp in set
    const int c0  = wis * p.slice; // first column of this warp's slice (forward: K range, backward: N range)

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
    double bias[NT8][2];
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt)
    {
        const int t = nt * 8 + tig * 2;
        bias[nt][0] = (t < T) ? p.b[t] : 0.0;
        bias[nt][1] = (t + 1 < T) ? p.b[t + 1] : 0.0;
    }
    double  fsum = 0.0;          // loss of the rows g this lane has seen (identical on the 4 lanes of a quad)
    log_product logs;            // softmax: the rows' sums of exponentials, multiplied up (loss.cuh)
    double  gbsum[NT8][2];       // sum over rows of dL[g][8 nt + 2 tig + j]
    double  lsum[NT8][2];        // per-target loss sums (batched models)
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt)
    {
        gbsum[nt][0] = gbsum[nt][1] = 0.0;
        lsum[nt][0] = lsum[nt][1] = 0.0;
    }
    double* my_dl = dl_tiles + warp * (kFdRows * kFdLdd);

    int      stage_index = 0;
    uint32_t phase       = 0;
    int      parity      = 0;
    int      turn        = 0;
    double*  set_exchange = exchange + static_cast<int64_t>(set) * (2 * WPS * kFdRows * TPS);
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
    {
        // every warp observes and releases every stage, also the other set's tiles (parity waits tell only adjacent
        // phases apart)
        const int s = stage_index;
        mbar_wait(full_bar + s, phase);
        const uint32_t wrapped = (stage_index + 1 == p.stages) ? 1u : 0u;
        stage_index            = wrapped ? 0 : stage_index + 1;
        phase ^= wrapped;
        if constexpr (NSETS > 1)
        {
            const bool mine = turn == set;
            turn            = (turn + 1 == NSETS) ? 0 : turn + 1;
            if (!mine)
            {
                if (lane == 0)
                {
                    mbar_arrive(empty_bar + s);
                }
                continue;
            }
        }

        const unsigned char* stage = smem + p.stage_offset + static_cast<int64_t>(s) * p.stage_stride;
        const double*        xs    = reinterpret_cast<const double*>(stage);
        const double*        ys    = reinterpret_cast<const double*>(stage + p.y_offset);
        const int            rows  = static_cast<int>(min(static_cast<int64_t>(kFdRows), p.N - tile * kFdRows));

        // ----- forward, split-K: partial outputs of this warp's column slice -----
        // (CH independent accumulator chains per target tile: a DMMA depends on the previous one of its chain, and
        // with one or two chains per warp the tensor pipe would wait for its own latency)
        constexpr int CH = 4 / NT8;
        double        acc[NT8][2];
        {
            double part[CH][NT8][2];
#pragma unroll
            for (int c = 0; c < CH; ++c)
            {
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt)
                {
                    part[c][nt][0] = part[c][nt][1] = 0.0;
                }
            }
            const bool    row_in = g < rows; // rows beyond the tile hold stale data: keep them out of the sums
            const double* xa     = xs + g * ld + c0 + tig;
            const double* wb     = w_s + g * ld + c0 + tig;
            bool          t_in[NT8]; // target rows of W that exist (the others multiply as zero)
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
                t_in[nt] = nt * 8 + g < T;
            }
            static_assert((2 * NS8) % CH == 0, "the k-steps of a slice must split evenly over the chains");
#pragma unroll 4
            for (int kk = 0; kk < 2 * NS8; kk += CH) // slice / 4 k-steps, CH at a time
            {
#pragma unroll
                for (int c = 0; c < CH; ++c)
                {
                    const double a = row_in ? xa[(kk + c) * 4] : 0.0;
#pragma unroll
                    for (int nt = 0; nt < NT8; ++nt)
                    {
                        const double bv = t_in[nt] ? wb[nt * 8 * ld + (kk + c) * 4] : 0.0;
                        dmma_m8n8k4(part[c][nt][0], part[c][nt][1], a, bv);
                    }
                }
            }
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
                acc[nt][0] = s0;
                acc[nt][1] = s1;
            }
        }
        // the rows' second use — the B fragments of the backward product — is fetched into registers now, together
        // with the targets, so that the stage goes back to the producer before the exchange / loss / backward phase
        // (with W resident only two or three stages fit: holding one for the whole tile exposed the HBM latency)
        // Measured: + 5-13 % with two target tiles (T > 8), - 8 % with one, where it stays at the end of the tile.
        constexpr bool EARLY = NT8 > 1;
        double         xb[kFdRows / 4][NS8];
        if constexpr (GRAD && EARLY)
        {
#pragma unroll
            for (int kk = 0; kk < kFdRows / 4; ++kk)
            {
                const int     r   = kk * 4 + tig;
                const double* src = xs + r * ld + c0 + g;
#pragma unroll
                for (int ns = 0; ns < NS8; ++ns)
                {
                    xb[kk][ns] = (r < rows) ? src[ns * 8] : 0.0; // stale rows: 0 x NaN would poison the sums
                }
            }
        }
        const bool row_ok = g < rows;
        double     y[NT8][2];
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt)
        {
#pragma unroll
            for (int j = 0; j < 2; ++j)
            {
                const int t = nt * 8 + tig * 2 + j;
                y[nt][j]    = (row_ok && t < T) ? (p.y_broadcast ? ys[g] : ys[g * T + t]) : 0.0;
            }
        }
        if constexpr (EARLY || !GRAD)
        {
            mbar_release_stage(empty_bar + s, lane);
        }

        // exchange: every warp ends up with the full outputs in fragment layout, summed in warp order
        double* slot = set_exchange + static_cast<int64_t>(parity) * (WPS * kFdRows * TPS);
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt)
        {
            *reinterpret_cast<double2*>(slot + (wis * kFdRows + g) * TPS + nt * 8 + tig * 2) =
                make_double2(acc[nt][0], acc[nt][1]);
        }
        named_bar_sync(1 + set, WPS * 32);
        double o[NT8][2];
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt)
        {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int q = 0; q < WPS; ++q)
            {
                const double2 v = *reinterpret_cast<const double2*>(slot + (q * kFdRows + g) * TPS + nt * 8 + tig * 2);
                s0 += v.x;
                s1 += v.y;
            }
            o[nt][0] = s0 + bias[nt][0]; // product first, then + bias (src/linear/util.cpp:19-20)
            o[nt][1] = s1 + bias[nt][1];
        }
        parity ^= 1;

        // ----- loss + dL from the fragments: lane (g, tig) holds row g, targets 8 nt + 2 tig + {0, 1} -----
        double dl[NT8][2];
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt)
        {
            dl[nt][0] = dl[nt][1] = 0.0;
        }
        if (p.loss == LOSS_CLASSNLL && p.y_broadcast == 0) // (batched single-output models: elementwise)
        {
            // classnll.h:19-37: the four lanes of a quad hold the whole row
            double omax = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                {
                    if (nt * 8 + tig * 2 + j < T)
                    {
                        omax = max_of(omax, o[nt][j]);
                    }
                }
            }
            omax = quad_max(omax);
            double ex[NT8][2];
            double sumexp = 0.0, dot = 0.0;
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                {
                    // (no branch: the 2 NT8 exponentials of a lane are independent chains; targets past T enter as
                    // -inf -> 0)
                    const bool in = nt * 8 + tig * 2 + j < T;
                    ex[nt][j]     = exp_nonpositive(in ? __dsub_rn(o[nt][j], omax) : -INFINITY);
                    sumexp += ex[nt][j];
                    dot += in ? __dmul_rn(__dadd_rn(1.0, y[nt][j]), o[nt][j]) : 0.0;
                }
            }
            sumexp = quad_sum(sumexp);
            dot    = quad_sum(dot);
            // one reciprocal per row instead of 2 NT8 divisions per lane (exp x (1 / sum): within one rounding of the
            // reference's quotient); selects instead of branches around the logarithm and the gradient
            // (the logarithm of the value is taken once per kLogEvery rows: loss.cuh, log_product)
            const double rcp = __ddiv_rn(1.0, sumexp);
            fsum += row_ok ? __dsub_rn(omax, __dmul_rn(0.5, dot)) : 0.0;
            logs.product *= row_ok ? sumexp : 1.0;
            log_product_flush(logs, 1, fsum);
            if (GRAD)
            {
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt)
                {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                    {
                        const double gv =
                            __dsub_rn(__dmul_rn(ex[nt][j], rcp), __dmul_rn(0.5, __dadd_rn(1.0, y[nt][j])));
                        dl[nt][j] = (row_ok && nt * 8 + tig * 2 + j < T) ? gv : 0.0;
                    }
                }
            }
        }
        else
        {
            double sum = 0.0;
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                {
                    if (row_ok && nt * 8 + tig * 2 + j < T)
                    {
                        double term;
                        loss_term<GRAD>(p.loss, p.loss_param, y[nt][j], o[nt][j], term, dl[nt][j]);
                        sum += term;
                        lsum[nt][j] += loss_post(p.loss, term); // one output per model: the factor applies per term
                    }
                }
            }
            fsum += loss_post(p.loss, quad_sum(sum));
        }

        if constexpr (GRAD)
        {
            // dL tile [8 rows][kFdLdd] of this warp -> A fragments of the backward product
            __syncwarp();
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
                gbsum[nt][0] += dl[nt][0];
                gbsum[nt][1] += dl[nt][1];
                *reinterpret_cast<double2*>(my_dl + g * kFdLdd + nt * 8 + tig * 2) = make_double2(dl[nt][0], dl[nt][1]);
            }
            __syncwarp();
            // ----- backward, split-N: gW[:, slice] += dL^T X[:, slice]; k runs over the 8 rows of the tile -----
#pragma unroll
            for (int kk = 0; kk < kFdRows / 4; ++kk)
            {
                const int r = kk * 4 + tig;
                double    a[NT8];
#pragma unroll
                for (int mt = 0; mt < NT8; ++mt)
                {
                    a[mt] = my_dl[r * kFdLdd + mt * 8 + g]; // rows beyond the tile carry dL = 0
                }
#pragma unroll
                for (int ns = 0; ns < NS8; ++ns)
                {
                    double bv;
                    if constexpr (EARLY)
                    {
                        bv = xb[kk][ns];
                    }
                    else
                    {
                        bv = (r < rows) ? xs[r * ld + c0 + g + ns * 8] : 0.0; // stale rows: 0 x NaN would poison
                    }
#pragma unroll
                    for (int mt = 0; mt < NT8; ++mt)
                    {
                        dmma_m8n8k4(gacc[mt][ns][0], gacc[mt][ns][1], a[mt], bv);
                    }
                }
            }
        }
        if constexpr (GRAD && !EARLY)
        {
            mbar_release_stage(empty_bar + s, lane);
        }
    }

    // ----- this CTA's partial row: [loss | sum dL (T) | sum dL x^T (T * D)] -----
    const bool batched = p.partials_gw != nullptr;
    double*    out     = p.partials + static_cast<int64_t>(blockIdx.x) *
                                   (batched ? (1 + 2 * T) : (1 + T + static_cast<int64_t>(T) * D));
    double*    out_gw  = batched ? p.partials_gw + static_cast<int64_t>(blockIdx.x) * T * D : out + 1 + T;

    // loss and dL sums of this warp's set: fold the 8 row slots (g) in order; every warp of a set carries the same
    fsum += log(logs.product); // (1 when nothing is pending)
    double f = (tig == 0) ? fsum : 0.0;
    f += __shfl_xor_sync(0xffffffffu, f, 4);
    f += __shfl_xor_sync(0xffffffffu, f, 8);
    f += __shfl_xor_sync(0xffffffffu, f, 16);
    if constexpr (GRAD)
    {
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt)
        {
#pragma unroll
            for (int j = 0; j < 2; ++j)
            {
                double b = gbsum[nt][j];
                b += __shfl_xor_sync(0xffffffffu, b, 4);
                b += __shfl_xor_sync(0xffffffffu, b, 8);
                b += __shfl_xor_sync(0xffffffffu, b, 16);
                gbsum[nt][j] = b;
            }
        }
    }
    if (p.partials_gw != nullptr)
    {
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt)
        {
#pragma unroll
            for (int j = 0; j < 2; ++j)
            {
                double l = lsum[nt][j];
                l += __shfl_xor_sync(0xffffffffu, l, 4);
                l += __shfl_xor_sync(0xffffffffu, l, 8);
                l += __shfl_xor_sync(0xffffffffu, l, 16);
                lsum[nt][j] = l;
            }
        }
    }

    if constexpr (NSETS > 1)
    {
        // both sets cover the same columns: set 1 parks its sums in the (now idle) exchange / stage areas, set 0 adds
        named_bar_sync(15, kFdWarps * 32); // every issued copy has been consumed, every exchange slot is dead
        double* park  = reinterpret_cast<double*>(smem + p.stage_offset);
        double* parkf = exchange; // [1 + 2 TP]
        if (set == 1)
        {
            if constexpr (GRAD)
            {
#pragma unroll
                for (int mt = 0; mt < NT8; ++mt)
                {
#pragma unroll
                    for (int ns = 0; ns < NS8; ++ns)
                    {
                        park[((static_cast<int64_t>(wis) * NT8 + mt) * NS8 + ns) * 64 + lane * 2]     = gacc[mt][ns][0];
                        park[((static_cast<int64_t>(wis) * NT8 + mt) * NS8 + ns) * 64 + lane * 2 + 1] = gacc[mt][ns][1];
                    }
                }
            }
            if (wis == 0)
            {
                if (lane == 0)
                {
                    parkf[0] = f;
                }
                if (g == 0)
                {
#pragma unroll
                    for (int nt = 0; nt < NT8; ++nt)
                    {
                        parkf[1 + nt * 8 + tig * 2]          = gbsum[nt][0];
                        parkf[1 + nt * 8 + tig * 2 + 1]      = gbsum[nt][1];
                        parkf[1 + TP + nt * 8 + tig * 2]     = lsum[nt][0];
                        parkf[1 + TP + nt * 8 + tig * 2 + 1] = lsum[nt][1];
                    }
                }
            }
        }
        named_bar_sync(15, kFdWarps * 32);
        if (set == 1)
        {
            return;
        }
        if constexpr (GRAD)
        {
#pragma unroll
            for (int mt = 0; mt < NT8; ++mt)
            {
#pragma unroll
                for (int ns = 0; ns < NS8; ++ns)
                {
                    gacc[mt][ns][0] += park[((static_cast<int64_t>(wis) * NT8 + mt) * NS8 + ns) * 64 + lane * 2];
                    gacc[mt][ns][1] += park[((static_cast<int64_t>(wis) * NT8 + mt) * NS8 + ns) * 64 + lane * 2 + 1];
                }
            }
        }
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt)
        {
            gbsum[nt][0] += parkf[1 + nt * 8 + tig * 2];
            gbsum[nt][1] += parkf[1 + nt * 8 + tig * 2 + 1];
            lsum[nt][0] += parkf[1 + TP + nt * 8 + tig * 2];
            lsum[nt][1] += parkf[1 + TP + nt * 8 + tig * 2 + 1];
        }
        f += parkf[0];
    }

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
    if (warp == 0)
    {
        if (lane == 0)
        {
            out[0] = f;
        }
        if (g == 0)
        {
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                {
                    const int t = nt * 8 + tig * 2 + j;
                    if (t < T)
                    {
                        if (GRAD || batched)
                        {
                            out[1 + t] = GRAD ? gbsum[nt][j] : 0.0;
                        }
                        if (batched)
                        {
                            out[1 + T + t] = lsum[nt][j];
                        }
                    }
                }
            }
        }
    }
}

#endif // __CUDACC__
} // namespace nb200
