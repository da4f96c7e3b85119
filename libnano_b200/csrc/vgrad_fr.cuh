// vgrad_fr.cuh — a handful of targets (T <= 16) over SHORT rows (D <= 256): ONE pass over X, both contractions on the
// FP64 tensor pipe, the forward split over ROWS instead of features.
//
// vgrad_fd.cuh splits the forward over the features of one 8-row tile: every tile costs a CTA-wide exchange of partial
// outputs behind a barrier, and all eight warps then evaluate the same tile's loss. That fixed cost (~4000 cycles per
// tile, measured) does not shrink with the row: at D = 256 the kernel read X at 1.06 TB/s. Here a GROUP is 64 TPW rows
// = 8 TPW m8 tiles in as many consecutive ring stages (TPW = 2 tiles per warp for D <= 128), and per group
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
constexpr int kFrMaxStages = 32;
constexpr int kFrBars      = 0;    // full [32] | empty [32]
constexpr int kFrDl        = 512;  // dL tiles [2][64 TPW][kFdLdd], then the per-warp fold rows [8][1 + 16]

// TPW: tiles per warp and group (a group is 8 TPW tiles = 64 TPW rows in as many consecutive ring stages)
__host__ __device__ constexpr int fr_group_tiles(const int tpw)
{
    return kFdWarps * tpw;
}
__host__ __device__ constexpr int fr_fold_offset(const int tpw)
{
    return kFrDl + 2 * fr_group_tiles(tpw) * kFdRows * kFdLdd * 8;
}
__host__ __device__ constexpr int fr_header(const int tpw)
{
    return (fr_fold_offset(tpw) + kFdWarps * 17 * 8 + 127) / 128 * 128;
}

#ifdef __CUDACC__

// fd_params: num_tiles = number of groups, slice = 8 * NS8 columns per warp, ld = 64 * NS8 + 4
template <int NT8, int NS8, int TPW, bool GRAD>
__global__ void __launch_bounds__(kFdThreads, 1) vgrad_fr_kernel(const fd_params p)
{
    constexpr int kTiles = fr_group_tiles(TPW); // tiles (= ring stages) per group; warp w owns tiles w, w + 8, ...
    constexpr int kRows  = kTiles * kFdRows;    // rows per group
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem + kFrBars);
    uint64_t* empty_bar = full_bar + kFrMaxStages;
    double*   dl_tiles  = reinterpret_cast<double*>(smem + kFrDl);
    double*   fold      = reinterpret_cast<double*>(smem + fr_fold_offset(TPW));
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
            for (int i = 0; i < kTiles; ++i)
            {
                mbar_wait(empty_bar + s, phase);
                const int64_t  row0    = group * kRows + i * kFdRows;
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
    log_product logs;     // softmax: the rows' sums of exponentials, multiplied up (loss.cuh)
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
    for (int64_t group = blockIdx.x; group < groups; group += gridDim.x, it += kTiles)
    {
        const int64_t group_row0 = group * kRows;
        // ----- forward of this warp's TPW tiles (tile u * 8 + warp of the group) over the full width -----
        const double* xo[TPW];
        bool          row_ok[TPW]; // rows beyond the data hold stale bytes: keep them out of every sum
        double        y[TPW][NT8][2];
        double        o[TPW][NT8][2];
#pragma unroll
        for (int u = 0; u < TPW; ++u)
        {
            const int use = it + u * kFdWarps + warp;
            const int s   = use % p.stages;
            mbar_wait(full_bar + s, static_cast<uint32_t>((use / p.stages) & 1));
            xo[u] = reinterpret_cast<const double*>(smem + p.stage_offset + static_cast<int64_t>(s) * p.stage_stride);
            const double* yo = reinterpret_cast<const double*>(smem + p.stage_offset + static_cast<int64_t>(s) * p.stage_stride + p.y_offset);
            row_ok[u] = group_row0 + (u * kFdWarps + warp) * kFdRows + g < p.N;
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                {
                    const int t = nt * 8 + tig * 2 + j;
                    y[u][nt][j] = (row_ok[u] && t < T) ? yo[g * T + t] : 0.0;
                }
            }
        }
        {
            // CH independent accumulator chains per tile and target tile (a DMMA depends on the previous one of its chain)
            constexpr int CH = 4 / TPW;
            double        part[TPW][CH][NT8][2];
#pragma unroll
            for (int u = 0; u < TPW; ++u)
            {
#pragma unroll
                for (int c = 0; c < CH; ++c)
                {
#pragma unroll
                    for (int nt = 0; nt < NT8; ++nt)
                    {
                        part[u][c][nt][0] = part[u][c][nt][1] = 0.0;
                    }
                }
            }
            const double* wb = w_s + g * ld + tig;
            int           kk = 0;
#pragma unroll 2
            for (; kk + CH <= ksteps; kk += CH)
            {
#pragma unroll
                for (int c = 0; c < CH; ++c)
                {
                    double bv[NT8];
#pragma unroll
                    for (int nt = 0; nt < NT8; ++nt)
                    {
                        bv[nt] = t_in[nt] ? wb[nt * 8 * ld + (kk + c) * 4] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < TPW; ++u)
                    {
                        const double a = row_ok[u] ? xo[u][g * ld + tig + (kk + c) * 4] : 0.0;
#pragma unroll
                        for (int nt = 0; nt < NT8; ++nt)
                        {
                            dmma_m8n8k4(part[u][c][nt][0], part[u][c][nt][1], a, bv[nt]);
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CH - 1; ++c) // the last ksteps % CH k-steps
            {
                if (kk + c < ksteps)
                {
                    double bv[NT8];
#pragma unroll
                    for (int nt = 0; nt < NT8; ++nt)
                    {
                        bv[nt] = t_in[nt] ? wb[nt * 8 * ld + (kk + c) * 4] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < TPW; ++u)
                    {
                        const double a = row_ok[u] ? xo[u][g * ld + tig + (kk + c) * 4] : 0.0;
#pragma unroll
                        for (int nt = 0; nt < NT8; ++nt)
                        {
                            dmma_m8n8k4(part[u][c][nt][0], part[u][c][nt][1], a, bv[nt]);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < TPW; ++u)
            {
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt)
                {
                    double s0 = part[u][0][nt][0], s1 = part[u][0][nt][1];
#pragma unroll
                    for (int c = 1; c < CH; ++c) // chain order: fixed
                    {
                        s0 += part[u][c][nt][0];
                        s1 += part[u][c][nt][1];
                    }
                    o[u][nt][0] = s0 + bias[nt][0]; // product first, then + bias (src/linear/util.cpp:19-20)
                    o[u][nt][1] = s1 + bias[nt][1];
                }
            }
        }

        // ----- the group's rows, this warp's columns -> B fragments of the backward product; stages back to the producer
        double xb[2 * kTiles][NS8]; // k-step kk = rows 4 kk .. 4 kk + 3 of the group
#pragma unroll
        for (int i = 0; i < kTiles; ++i)
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
        }
        // one cross-proxy fence for the whole group (see mbar_release_stage: the loads above must be ordered before the
        // refills), then lane i hands stage i back — a fence per stage made every tile wait for its own loads
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane < kTiles)
        {
            mbar_arrive(empty_bar + (it + lane) % p.stages);
        }

        // ----- loss + dL of this warp's tiles from its fragments: lane (g, tig) holds row g, targets 8 nt + 2 tig + {0, 1}
        double dl[TPW][NT8][2];
#pragma unroll
        for (int u = 0; u < TPW; ++u)
        {
#pragma unroll
            for (int nt = 0; nt < NT8; ++nt)
            {
                dl[u][nt][0] = dl[u][nt][1] = 0.0;
            }
        }
        if (p.loss == LOSS_CLASSNLL)
        {
            // classnll.h:19-37: the four lanes of a quad hold the whole row; the TPW tiles' chains are independent
            double omax[TPW], sumexp[TPW], dot[TPW];
            double ex[TPW][NT8][2];
#pragma unroll
            for (int u = 0; u < TPW; ++u)
            {
                omax[u] = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt)
                {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                    {
                        if (nt * 8 + tig * 2 + j < T)
                        {
                            omax[u] = max_of(omax[u], o[u][nt][j]);
                        }
                    }
                }
                omax[u] = quad_max(omax[u]);
            }
#pragma unroll
            for (int u = 0; u < TPW; ++u)
            {
                sumexp[u] = dot[u] = 0.0;
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt)
                {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                    {
                        // (no branch: the 2 NT8 TPW exponentials of a lane are independent chains; targets past T
                        // enter as -inf -> 0)
                        const bool in = nt * 8 + tig * 2 + j < T;
                        ex[u][nt][j]  = exp_nonpositive(in ? __dsub_rn(o[u][nt][j], omax[u]) : -INFINITY);
                        sumexp[u] += ex[u][nt][j];
                        dot[u] += in ? __dmul_rn(__dadd_rn(1.0, y[u][nt][j]), o[u][nt][j]) : 0.0;
                    }
                }
                sumexp[u] = quad_sum(sumexp[u]);
                dot[u]    = quad_sum(dot[u]);
            }
#pragma unroll
            for (int u = 0; u < TPW; ++u)
            {
                // one reciprocal per row instead of 2 NT8 divisions per lane (exp x (1 / sum): within one rounding of
                // the reference's quotient); selects instead of branches around the logarithm and the gradient
                // (the logarithm of the value is taken once per kLogEvery rows: loss.cuh, log_product)
                const double rcp = __ddiv_rn(1.0, sumexp[u]);
                fsum += row_ok[u] ? __dsub_rn(omax[u], __dmul_rn(0.5, dot[u])) : 0.0;
                logs.product *= row_ok[u] ? sumexp[u] : 1.0;
                if (GRAD)
                {
#pragma unroll
                    for (int nt = 0; nt < NT8; ++nt)
                    {
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                        {
                            const double gv = __dsub_rn(__dmul_rn(ex[u][nt][j], rcp),
                                                        __dmul_rn(0.5, __dadd_rn(1.0, y[u][nt][j])));
                            dl[u][nt][j] = (row_ok[u] && nt * 8 + tig * 2 + j < T) ? gv : dl[u][nt][j];
                        }
                    }
                }
            }
            log_product_flush(logs, TPW, fsum);
        }
        else
        {
#pragma unroll
            for (int u = 0; u < TPW; ++u)
            {
                double sum = 0.0;
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt)
                {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                    {
                        if (row_ok[u] && nt * 8 + tig * 2 + j < T)
                        {
                            double term;
                            loss_term<GRAD>(p.loss, p.loss_param, y[u][nt][j], o[u][nt][j], term, dl[u][nt][j]);
                            sum += term;
                        }
                    }
                }
                fsum += loss_post(p.loss, quad_sum(sum)); // tile order: fixed
            }
        }

        if constexpr (GRAD)
        {
            // ----- the dL tiles meet in shared memory: [rows of the group][kFdLdd], rows beyond the data carry dL = 0 -----
            double* tiles = dl_tiles + parity * (kRows * kFdLdd);
#pragma unroll
            for (int u = 0; u < TPW; ++u)
            {
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt)
                {
                    gbsum[nt][0] += dl[u][nt][0];
                    gbsum[nt][1] += dl[u][nt][1];
                    *reinterpret_cast<double2*>(tiles + ((u * kFdWarps + warp) * kFdRows + g) * kFdLdd + nt * 8 + tig * 2) =
                        make_double2(dl[u][nt][0], dl[u][nt][1]);
                }
            }
            named_bar_sync(1, kFdWarps * 32);
            parity ^= 1;
            // ----- backward, split over columns: gW[:, slice] += dL^T X[:, slice]; k runs over the group's rows -----
#pragma unroll
            for (int kk = 0; kk < 2 * kTiles; ++kk)
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
    fsum += log(logs.product); // (1 when nothing is pending)
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
