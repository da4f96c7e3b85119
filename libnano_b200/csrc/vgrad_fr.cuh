// vgrad_fr.cuh — a handful of targets (T <= 16) over SHORT rows (D <= 256): ONE pass over X, both contractions on the
// FP64 tensor pipe, the forward split over ROWS instead of features.
//
// vgrad_fd.cuh splits the forward over the features of one 8-row tile: every tile costs a CTA-wide exchange of partial
// outputs behind a barrier, and all eight warps then evaluate the same tile's loss. That fixed cost (~4000 cycles per
// tile, measured) does not shrink with the row: at D = 256 the kernel read X at 1.06 TB/s. Here a GROUP is 64 rows =
// eight m8 tiles in eight consecutive ring stages, and per group
//
//   * forward, split over rows: warp w multiplies ITS tile over the full width, O_w[8, TP] = X_w W^T (four
//     independent accumulator chains per target tile), src/linear/util.cpp:19-20 — no partial sums to exchange;
//   * the loss and dL of tile w come straight from warp w's accumulator fragments, once (loss.cuh; softmax: the four
//     lanes of a quad hold a row) — not eight times;
//   * every warp fetches the B fragments of the backward product — its column slice of all 64 rows — into registers
//     and hands the eight stages back to the producer, which refills them while the group is finished;
//   * the eight dL tiles meet in shared memory behind ONE named barrier per group (double-buffered);
//   * backward, split over columns: warp w accumulates gW[:, slice_w] += dL^T X[:, slice_w] over the 64 rows in registers
//     for the whole launch (src/linear/function.cpp:68-70);
//   * per-CTA partial rows [loss | sum dL | sum dL x^T] as in vgrad_fd.cuh (same reduce / finish kernels).
//
// Shared memory: barriers | dL tiles [2][64][kFdLdd] | fold scratch | W [T][ld] | stages of 8 rows [8][ld] + targets.
#pragma once

#include "vgrad_fd.cuh"

namespace nb200
{
constexpr int kFrMaxStages = 16;
constexpr int kFrGroup     = kFdWarps;            // tiles per group: one per consumer warp
constexpr int kFrGroupRows = kFrGroup * kFdRows;  // 64
constexpr int kFrBars      = 0;                   // full [16] | empty [16]
constexpr int kFrDl        = 256;                 // dL tiles [2][64][kFdLdd]
constexpr int kFrFold      = kFrDl + 2 * kFrGroupRows * kFdLdd * 8; // per-warp fold rows [8][1 + 16]
constexpr int kFrHeader    = (kFrFold + kFdWarps * 17 * 8 + 127) / 128 * 128;

#ifdef __CUDACC__

// fd_params: num_tiles = number of 64-row groups, slice = 8 * NS8 columns per warp, ld = 64 * NS8 + 4
template <int NT8, int NS8, bool GRAD>
__global__ void __launch_bounds__(kFdThreads, 1) vgrad_fr_kernel(const fd_params p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem + kFrBars);
    uint64_t* empty_bar = full_bar + kFrMaxStages;
    double*   dl_tiles  = reinterpret_cast<double*>(smem + kFrDl);
    double*   fold      = reinterpret_cast<double*>(smem + kFrFold);
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
    // W -> shared memory (zero beyond D columns), and zero the pad columns of every stage row once: the bulk copies
    // write exactly D doubles per row, so whatever the contractions read beyond D stays zero
    for (int i = threadIdx.x; i < T * ld; i += kFdThreads)
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

    const int64_t groups = p.num_tiles;
    const int     ty     = T; // targets stored per row of Y

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
        for (int64_t group = blockIdx.x; group < groups; group += gridDim.x)
        {
            for (int i = 0; i < kFrGroup; ++i)
            {
                mbar_wait(empty_bar + s, phase);
                const int64_t  row0    = group * kFrGroupRows + i * kFdRows;
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
    const int c0 = warp * p.slice; // first column of this warp's slice of the backward product

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
    double fsum = 0.0;    // loss of the rows g of this warp's tiles (identical on the 4 lanes of a quad)
    double gbsum[NT8][2]; // sum over this warp's rows of dL[g][8 nt + 2 tig + j]
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt)
    {
        gbsum[nt][0] = gbsum[nt][1] = 0.0;
    }
    bool t_in[NT8]; // target rows of W that exist (the others multiply as zero)
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt)
    {
        t_in[nt] = nt * 8 + g < T;
    }
    const int ksteps = (D + 3) / 4; // k-steps of the forward (the columns past D are zero in both operands)

    int it     = 0; // stage uses so far: use `it` sits in stage it % stages, phase (it / stages) & 1
    int parity = 0;
    for (int64_t group = blockIdx.x; group < groups; group += gridDim.x, it += kFrGroup)
    {
        const int64_t group_row0 = group * kFrGroupRows;
        // ----- forward of this warp's tile over the full width -----
        const int own_it = it + warp;
        const int own_s  = own_it % p.stages;
        mbar_wait(full_bar + own_s, static_cast<uint32_t>((own_it / p.stages) & 1));
        const double* xo   = reinterpret_cast<const double*>(smem + p.stage_offset + static_cast<int64_t>(own_s) * p.stage_stride);
        const double* yo   = reinterpret_cast<const double*>(smem + p.stage_offset + static_cast<int64_t>(own_s) * p.stage_stride + p.y_offset);
        const int     rows = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(kFdRows), p.N - group_row0 - warp * kFdRows)));
        const bool    row_ok = g < rows; // rows beyond the data hold stale bytes: keep them out of every sum
        double        o[NT8][2];
        {
            constexpr int CH = 4; // independent accumulator chains per target tile (a DMMA depends on the previous one)
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
            const double* xa = xo + g * ld + tig;
            const double* wb = w_s + g * ld + tig;
            int           kk = 0;
#pragma unroll 2
            for (; kk + CH <= ksteps; kk += CH)
            {
#pragma unroll
                for (int c = 0; c < CH; ++c)
                {
                    const double a = row_ok ? xa[(kk + c) * 4] : 0.0;
#pragma unroll
                    for (int nt = 0; nt < NT8; ++nt)
                    {
                        const double bv = t_in[nt] ? wb[nt * 8 * ld + (kk + c) * 4] : 0.0;
                        dmma_m8n8k4(part[c][nt][0], part[c][nt][1], a, bv);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CH - 1; ++c) // the last ksteps % CH k-steps
            {
                if (kk + c < ksteps)
                {
                    const double a = row_ok ? xa[(kk + c) * 4] : 0.0;
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
                o[nt][0] = s0 + bias[nt][0]; // product first, then + bias (src/linear/util.cpp:19-20)
                o[nt][1] = s1 + bias[nt][1];
            }
        }
        double y[NT8][2];
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt)
        {
#pragma unroll
            for (int j = 0; j < 2; ++j)
            {
                const int t = nt * 8 + tig * 2 + j;
                y[nt][j]    = (row_ok && t < T) ? yo[g * T + t] : 0.0;
            }
        }

        // ----- the group's rows, this warp's columns -> B fragments of the backward product; stages back to the producer
        double xb[2 * kFrGroup][NS8]; // k-step kk = rows 4 kk .. 4 kk + 3 of the group
#pragma unroll
        for (int i = 0; i < kFrGroup; ++i)
        {
            const int use = it + i;
            const int s   = use % p.stages;
            mbar_wait(full_bar + s, static_cast<uint32_t>((use / p.stages) & 1));
            if constexpr (GRAD)
            {
                const double* xs     = reinterpret_cast<const double*>(smem + p.stage_offset + static_cast<int64_t>(s) * p.stage_stride);
                const int     rows_i = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(kFdRows), p.N - group_row0 - i * kFdRows)));
#pragma unroll
                for (int h = 0; h < 2; ++h)
                {
                    const int     r   = h * 4 + tig;
                    const double* src = xs + r * ld + c0 + g;
#pragma unroll
                    for (int ns = 0; ns < NS8; ++ns)
                    {
                        xb[2 * i + h][ns] = (r < rows_i) ? src[ns * 8] : 0.0; // stale rows: 0 x NaN would poison the sums
                    }
                }
            }
            mbar_release_stage(empty_bar + s, lane);
        }

        // ----- loss + dL of this warp's tile from its fragments: lane (g, tig) holds row g, targets 8 nt + 2 tig + {0, 1}
        double dl[NT8][2];
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt)
        {
            dl[nt][0] = dl[nt][1] = 0.0;
        }
        if (p.loss == LOSS_CLASSNLL)
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
                    ex[nt][j] = 0.0;
                    if (nt * 8 + tig * 2 + j < T)
                    {
                        ex[nt][j] = exp(__dsub_rn(o[nt][j], omax));
                        sumexp += ex[nt][j];
                        dot += __dmul_rn(__dadd_rn(1.0, y[nt][j]), o[nt][j]);
                    }
                }
            }
            sumexp = quad_sum(sumexp);
            dot    = quad_sum(dot);
            if (row_ok)
            {
                fsum += __dadd_rn(__dsub_rn(log(sumexp), __dmul_rn(0.5, dot)), omax);
                if (GRAD)
                {
#pragma unroll
                    for (int nt = 0; nt < NT8; ++nt)
                    {
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                        {
                            if (nt * 8 + tig * 2 + j < T)
                            {
                                dl[nt][j] = __dsub_rn(__ddiv_rn(ex[nt][j], sumexp),
                                                      __dmul_rn(0.5, __dadd_rn(1.0, y[nt][j])));
                            }
                        }
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
                    }
                }
            }
            fsum += loss_post(p.loss, quad_sum(sum));
        }

        if constexpr (GRAD)
        {
            // ----- the eight dL tiles meet in shared memory: [64 rows][kFdLdd], rows beyond the data carry dL = 0 -----
            double* tiles = dl_tiles + parity * (kFrGroupRows * kFdLdd);
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
                gbsum[nt][0] += dl[nt][0];
                gbsum[nt][1] += dl[nt][1];
                *reinterpret_cast<double2*>(tiles + (warp * kFdRows + g) * kFdLdd + nt * 8 + tig * 2) =
                    make_double2(dl[nt][0], dl[nt][1]);
            }
            named_bar_sync(1, kFdWarps * 32);
            parity ^= 1;
            // ----- backward, split over columns: gW[:, slice] += dL^T X[:, slice]; k runs over the 64 rows -----
#pragma unroll
            for (int kk = 0; kk < 2 * kFrGroup; ++kk)
            {
                const int r = kk * 4 + tig;
                double    a[NT8];
#pragma unroll
                for (int mt = 0; mt < NT8; ++mt)
                {
                    a[mt] = tiles[r * kFdLdd + mt * 8 + g];
                }
#pragma unroll
                for (int ns = 0; ns < NS8; ++ns)
                {
#pragma unroll
                    for (int mt = 0; mt < NT8; ++mt)
                    {
                        dmma_m8n8k4(gacc[mt][ns][0], gacc[mt][ns][1], a[mt], xb[kk][ns]);
                    }
                }
            }
        }
    }

    // ----- this CTA's partial row: [loss | sum dL (T) | sum dL x^T (T * D)] -----
    double* out    = p.partials + static_cast<int64_t>(blockIdx.x) * (1 + T + static_cast<int64_t>(T) * D);
    double* out_gw = out + 1 + T;

    // loss and dL sums: fold the 8 row slots (g) of every warp in order, then the warps in warp order
    double f = (tig == 0) ? fsum : 0.0;
    f += __shfl_xor_sync(0xffffffffu, f, 4);
    f += __shfl_xor_sync(0xffffffffu, f, 8);
    f += __shfl_xor_sync(0xffffffffu, f, 16);
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
    double* mine = fold + warp * 17;
    if (lane == 0)
    {
        mine[0] = f;
    }
    if (g == 0)
    {
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt)
        {
            mine[1 + nt * 8 + tig * 2]     = gbsum[nt][0];
            mine[1 + nt * 8 + tig * 2 + 1] = gbsum[nt][1];
        }
    }
    named_bar_sync(1, kFdWarps * 32);
    if (warp == 0 && lane <= T)
    {
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < kFdWarps; ++q)
        {
            sum += fold[q * 17 + lane];
        }
        if (lane == 0 || GRAD)
        {
            out[lane] = sum;
        }
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
}

#endif // __CUDACC__
} // namespace nb200
