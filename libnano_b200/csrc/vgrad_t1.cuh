// vgrad_t1.cuh — the hot kernel: ONE pass over X for a single-target linear model, ONE launch per evaluation.
//
// Replaces, for T == 1, the per-batch body of linear::function_t::do_eval (src/linear/function.cpp:53-72):
//   predict (src/linear/util.cpp:19-20)  o_i = x_i . w + b
//   loss value + vgrad (flatten.h:69-86) l_i, dL_i
//   accumulation (function.cpp:62,69-70) sum l_i ; sum dL_i ; sum dL_i * x_i
// the reduction over workers (sum_reduce, include/nano/core/reduce.h:23-31) and the regularisation block
// (function.cpp:127-143,158-167); likewise linear_model_t::eval (src/function/mlearn/linear.h:21-38, two passes
// over X in the reference) with the regularisers of ridge.cpp:63-80 / lasso.cpp:63-76 / elasticnet.cpp:67-86.
//
// Design (HBM-bound: 8 bytes of X per 4 flops):
//   * persistent CTAs (one per SM), static round-robin tiles of RPS consecutive rows => fixed summation order;
//   * a producer warp streams tiles into a ring of shared-memory stages with 1-D bulk async copies
//     (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP) — rows are contiguous, so a tile is one flat copy;
//   * consumer warps own rows: 128-bit conflict-free LDS of the row into REGISTERS, dot with w (shared memory),
//     butterfly all-reduce, loss + dL in registers, then dL * x is accumulated into per-lane column slices — X is
//     read from shared memory exactly once and from HBM exactly once;
//   * the stage is released as soon as its rows sit in registers, before the loss math;
//   * per-CTA partials [loss | gb | gW] are folded IN THE SAME LAUNCH: a grid-wide barrier (cooperative launch), then
//     every CTA reduces a slice of columns over all CTAs in a fixed order and — single-rank case — applies 1/N and the
//     regularisation terms and writes (g, f) straight to the output (which may be mapped host memory).
//     Sample-sharded ranks stop at the un-normalised sums and all-reduce them (sharded.py).
//
// Template knobs: V = doubles per lane per 32-lane chunk (2 => 128-bit loads, needs even D; 1 => any D),
// KC = chunks per warp (a warp covers 32*V*KC columns), WS = warps cooperating on one row (row split in WS
// column slices), NCW = consumer warps, RW = rows a warp handles per iteration (several short rows amortise the
// shuffle / loss chain: lane r evaluates the loss of row r), GRAD = accumulate the gradient (false: value-only call).
#pragma once

#include <type_traits>

#include "common.cuh"
#include "loss.cuh"

namespace nb200
{
constexpr int kT1MaxStages  = 8;
constexpr int kT1HeaderSize    = 1024; // barriers [0, 128) + dot-exchange slots [128, 640) + tail scratch
constexpr int kT1ExchangeBytes = 512; // tail scratch [640, 1024): 2 x (NCW + 4) doubles
constexpr int kT1MaxPeers   = 8;    // ranks of one NVSwitch domain

struct t1_params
{
    const double* X;        // [N, D] row-major, padded by >= 16 bytes at the end
    const double* Y;        // [N], padded likewise
    const double* w;        // [D] device
    const double* bias;     // 1 double, device
    double*       partials; // [grid, D + 2] : [loss sum | dL sum | dL*x sums]
    int64_t       N;
    int64_t       num_tiles;
    int           D;
    int           rows_per_stage; // RPS: even, = (NCW / WS) * rows_per_group
    int           rows_per_group; // rows each warp-group handles per stage
    int           stages;
    int           stage_stride; // bytes between stages (multiple of 128)
    int           y_offset;     // byte offset of the Y tile inside a stage (multiple of 128)
    int           loss;
    double        loss_param;
    int           l2_resident; // 1: X + Y fit in L2 and are re-read by the next evaluation

    // fused tail: cross-CTA reduction after a grid-wide barrier, optionally followed by the final scaling
    unsigned long long* grid_counter; // monotonically increasing arrival counter (never reset between launches)
    unsigned long long  grid_target;  // counter value at which every CTA of THIS launch has arrived
    double*             reduced;      // [D + 2] un-normalised sums out (always written)
    int                 fuse_finish;  // 1: also write gx_out / fx_out (single-rank evaluation)
    int                 flavour;      // 0 LINEAR, 1 SYNTH
    double*             gx_out;       // [D + 1] (LINEAR) or [D] (SYNTH); device or mapped host memory
    double*             fx_out;       // 1 double
    int64_t             n_global;
    double              l1;
    double              l2;

    // fused cross-rank all-reduce over NVLink peer memory (sample-sharded evaluation, world > 1): every rank owns a
    // mailbox [2 (parity)][world][col_stride] doubles and flags [2][world][grid]; rank r writes its locally reduced
    // column slices into slot r of EVERY rank's mailbox and raises the matching per-CTA flags.
    int                 world;       // 1: no cross-rank stage
    int                 rank;
    int                 col_stride;  // doubles per mailbox slot (>= D + 2)
    unsigned long long  epoch;       // evaluation counter, identical on all ranks
    double*             peer_mail[kT1MaxPeers];
    unsigned long long* peer_flag[kT1MaxPeers];
    int*                peer_error;  // mapped host int: set when a peer did not show up in time
    // peer_wait == 1: the launch itself waits for the peers' slices and folds them (compute + collective in ONE kernel:
    // every CTA spins on per-CTA flags, so the launch holds the whole GPU until every peer has published).
    // peer_wait == 0: the launch publishes its slices and leaves; the LAST CTA to publish raises ONE ready word per peer
    // (behind the per-CTA flags: [2][world] words), the stream then waits on those words (cuStreamWaitValue64: no SM is
    // held while waiting, so evaluations of different objectives cannot cross-wait) and peer_finish_kernel folds the
    // slots in rank order.
    int                 peer_wait;
    unsigned long long* publish_counter; // CTAs of this device that have published (monotone, like grid_counter)
    unsigned long long  publish_target;  // value at which every CTA of THIS launch has published
};

#ifdef __CUDACC__

// grid-wide barrier for a cooperative launch (all CTAs co-resident). The counter only grows: launch k of a plan
// with G CTAs waits for the value k * G, so no reset is needed between launches.
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, const unsigned long long target)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence(); // publish this CTA's partials device-wide before signalling
        atomicAdd(counter, 1ULL);
        unsigned long long seen;
        do
        {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(counter) : "memory");
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// consumer role: rows -> registers -> dot -> loss -> gradient slices; leaves this CTA's partial row in HBM
// ---------------------------------------------------------------------------------------------------------------
template <int V, int KC, int WS, int NCW, int RW, int ALT, bool GRAD>
__device__ __forceinline__ void consume_rows(const t1_params& p, uint64_t* full_bar, uint64_t* empty_bar,
                                             double* exchange, const double* w_s, unsigned char* stage_base,
                                             const int warp, const int lane)
{
    static_assert(RW >= 1 && RW <= 32, "rows per iteration");
    static_assert(WS == 1 || (NCW / WS) * 2 * WS * RW * 8 <= kT1ExchangeBytes, "the dot-exchange slots must fit the header");
    static_assert((RW & (RW - 1)) == 0, "rows per iteration must be a power of two");
    static_assert(ALT >= 1 && NCW % (ALT * WS) == 0, "the alternating warp sets must hold whole row groups");
    constexpr int NG  = NCW / WS;    // row groups of the CTA (all sets): one partial gradient row each in the fold
    constexpr int SW  = NCW / ALT;   // consumer warps per set; set a serves the CTA's tiles a, a + ALT, a + 2 ALT, ...
    constexpr int CPW = 32 * V * KC; // columns per warp
    // after warp_reduce_rows the total of row r sits in lanes [r << OWNER_SHIFT, (r + 1) << OWNER_SHIFT); the first
    // lane of each group evaluates the loss of its row
    constexpr int OWNER_SHIFT = 5 - log2_of(RW);
    const int     my_row      = lane >> OWNER_SHIFT;
    const bool    owner       = (lane & ((1 << OWNER_SHIFT) - 1)) == 0;
    const int     D   = p.D;
    const int64_t rps = p.rows_per_stage;

    const int set   = warp / SW;        // which of the CTA's tiles
    const int group = warp / WS;        // row group within the CTA (fold slot, named barrier)
    const int sgrp  = (warp % SW) / WS; // row group within the set: which rows of a stage
    const int slice = warp % WS;        // which columns of a row
    const int cbase = slice * CPW; // first column of this warp's slice

    double g[V * KC]; // per-lane slice of sum dL_i * x_i
#pragma unroll
    for (int k = 0; k < V * KC; ++k)
    {
        g[k] = 0.0;
    }
    // lane r carries the loss / dL sums of the r-th row of each iteration (the other lanes stay at zero)
    double fsum = 0.0, gbsum = 0.0;

    const double bias = *p.bias;
    const int    rpg  = p.rows_per_group;

    int      stage_index = 0; // the CTA's it-th tile lands in stage it % stages
    uint32_t phase       = 0;
    int      turn        = 0; // it % ALT: the set whose tile this is
    int      exparity    = 0;
    // EXACT: the warps' column slices tile the row exactly (D == WS * CPW), so no column predicate is needed
    const bool    exact      = D == WS * CPW;
    const int64_t full_tiles = p.N / rps; // tiles [0, full_tiles) hold rps rows; at most one partial tile follows
    const int     first      = sgrp * rpg;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
    {
        // every warp observes EVERY fill of the ring in order and releases it, also the tiles of the other sets: a
        // parity wait can tell only adjacent phases apart, so no warp may fall a whole lap behind the producer
        const int s = stage_index;
        mbar_wait(full_bar + s, phase);
        const uint32_t wrapped = (stage_index + 1 == p.stages) ? 1u : 0u;
        stage_index            = wrapped ? 0 : stage_index + 1;
        phase ^= wrapped;
        if constexpr (ALT > 1)
        {
            const bool mine = turn == set;
            turn            = (turn + 1 == ALT) ? 0 : turn + 1;
            if (!mine)
            {
                if (lane == 0)
                {
                    mbar_arrive(empty_bar + s);
                }
                continue;
            }
        }

        const unsigned char* stage = stage_base + static_cast<int64_t>(s) * p.stage_stride;
        const double*        xs    = reinterpret_cast<const double*>(stage);
        const double*        ys    = reinterpret_cast<const double*>(stage + p.y_offset);

        int count = rpg; // rows of this stage owned by this warp group
        if (tile >= full_tiles)
        {
            const int rows = static_cast<int>(p.N - tile * rps);
            count          = max(0, min(rpg, rows - first));
        }
        if (count == 0)
        {
            if (lane == 0)
            {
                mbar_arrive(empty_bar + s);
            }
            continue;
        }

        // one iteration = RW rows of this warp; FULL: all RW rows exist (no per-row predicates in the hot path)
        const auto iteration = [&](const int i, const int nrows, auto full_tag, auto exact_tag)
        {
            constexpr bool FULL  = decltype(full_tag)::value;
            constexpr bool EXACT = decltype(exact_tag)::value;

            // RW row slices -> registers; partial dots with w (w is read once per iteration)
            double x[RW][V * KC];
            double acc[RW][2];
#pragma unroll
            for (int r = 0; r < RW; ++r)
            {
                acc[r][0] = acc[r][1] = 0.0;
            }
            const double* rows0 = xs + static_cast<int64_t>(first + i) * D;
#pragma unroll
            for (int k = 0; k < KC; ++k)
            {
                const int  col = cbase + (k * 32 + lane) * V;
                const bool in  = EXACT || col < D;
                if constexpr (V == 2)
                {
                    const double2 wv = in ? *reinterpret_cast<const double2*>(w_s + col) : make_double2(0.0, 0.0);
#pragma unroll
                    for (int r = 0; r < RW; ++r)
                    {
                        double2 xv = make_double2(0.0, 0.0);
                        if (in && (FULL || r < nrows))
                        {
                            xv = *reinterpret_cast<const double2*>(rows0 + static_cast<int64_t>(r) * D + col);
                        }
                        x[r][2 * k]     = xv.x;
                        x[r][2 * k + 1] = xv.y;
                        acc[r][0]       = fma(xv.x, wv.x, acc[r][0]);
                        acc[r][1]       = fma(xv.y, wv.y, acc[r][1]);
                    }
                }
                else
                {
                    const double wv = in ? w_s[col] : 0.0;
#pragma unroll
                    for (int r = 0; r < RW; ++r)
                    {
                        const double xv = (in && (FULL || r < nrows)) ? rows0[static_cast<int64_t>(r) * D + col] : 0.0;
                        x[r][k]         = xv;
                        acc[r][k & 1]   = fma(xv, wv, acc[r][k & 1]);
                    }
                }
            }
            // target of "my" row: the lanes of group (lane >> OWNER_SHIFT) serve row (lane >> OWNER_SHIFT)
            const double y = ys[first + i + min(my_row, nrows - 1)];

            // the stage can be refilled as soon as the last rows this warp owns are in registers
            if (i + RW >= count)
            {
                mbar_release_stage(empty_bar + s, lane);
            }

            // the RW partial dots are reduced together (recursive halving: RW - 1 exchanges instead of 5 * RW); the
            // lanes of group (lane >> OWNER_SHIFT) end up with the total of that row
            double part[RW];
#pragma unroll
            for (int r = 0; r < RW; ++r)
            {
                part[r] = acc[r][0] + acc[r][1];
            }
            double mine = warp_reduce_rows<RW>(part, lane);
            if constexpr (WS > 1)
            {
                // combine the column slices of the WS warps sharing this row, in slice order
                double* slot = exchange + (group * 2 + exparity) * (WS * RW); // [WS][RW]
                if (owner)
                {
                    slot[slice * RW + my_row] = mine;
                }
                named_bar_sync(1 + group, WS * 32);
                mine = 0.0;
#pragma unroll
                for (int q = 0; q < WS; ++q)
                {
                    mine += slot[q * RW + my_row];
                }
                exparity ^= 1;
            }

            // predict + loss + dL (src/linear/util.cpp:19-20: product first, then + bias). ONE lane per row evaluates
            // the loss and the other lanes sit the transcendental chain out (32 redundant fp64 evaluations per row
            // were most of the kernel's arithmetic energy, and the board runs at its power cap); dL is broadcast.
            double dl = 0.0;
            if (owner && (FULL || my_row < nrows))
            {
                const double o = mine + bias;
                double       term;
                loss_term<GRAD>(p.loss, p.loss_param, y, o, term, dl);
                fsum += loss_post(p.loss, term);
                gbsum += dl;
            }

            if constexpr (GRAD)
            {
#pragma unroll
                for (int r = 0; r < RW; ++r)
                {
                    const double dlr = __shfl_sync(0xffffffffu, dl, r << OWNER_SHIFT);
#pragma unroll
                    for (int k = 0; k < V * KC; ++k)
                    {
                        g[k] = fma(dlr, x[r][k], g[k]);
                    }
                }
            }
        };

        int i = 0;
        if (exact)
        {
            for (; i + RW <= count; i += RW)
            {
                iteration(i, RW, std::true_type{}, std::true_type{});
            }
        }
        else
        {
            for (; i + RW <= count; i += RW)
            {
                iteration(i, RW, std::true_type{}, std::false_type{});
            }
        }
        if (i < count)
        {
            iteration(i, count - i, std::false_type{}, std::false_type{});
        }
    }
    if constexpr (RW > 1)
    {
        // the owner lanes carried the per-row sums: fold them (fixed order) so that every lane holds the totals
        fsum  = warp_allreduce_sum(fsum);
        gbsum = warp_allreduce_sum(gbsum);
    }

    // fixed-order reduction of the row groups inside the CTA, through the (now idle) stages
    named_bar_sync(15, NCW * 32); // every consumer left the main loop; all issued copies have been consumed

    const int dpad = WS * CPW;
    double*   red  = reinterpret_cast<double*>(stage_base);  // [NG][dpad]
    double*   redf = red + static_cast<int64_t>(NG) * dpad;  // [NG] loss sums, then [NG] dL sums

    if constexpr (GRAD)
    {
#pragma unroll
        for (int k = 0; k < KC; ++k)
        {
            const int col = cbase + (k * 32 + lane) * V;
#pragma unroll
            for (int v = 0; v < V; ++v)
            {
                red[static_cast<int64_t>(group) * dpad + col + v] = g[k * V + v];
            }
        }
    }
    if (slice == 0 && lane == 0)
    {
        redf[group]      = fsum;
        redf[NG + group] = gbsum;
    }
    named_bar_sync(15, NCW * 32);

    double* out = p.partials + static_cast<int64_t>(blockIdx.x) * (D + 2);
    if constexpr (GRAD)
    {
        for (int col = threadIdx.x; col < D; col += NCW * 32)
        {
            double sum = 0.0;
#pragma unroll
            for (int q = 0; q < NG; ++q)
            {
                sum += red[static_cast<int64_t>(q) * dpad + col];
            }
            out[2 + col] = sum;
        }
    }
    if (threadIdx.x == 0)
    {
        double f = 0.0, b = 0.0;
#pragma unroll
        for (int q = 0; q < NG; ++q)
        {
            f += redf[q];
            b += redf[NG + q];
        }
        out[0] = f;
        out[1] = b;
    }
}

// the final scaling + regularisation of one column of the reduced vector [loss | gb | gW] (one thread)
template <int NWARPS>
__device__ __forceinline__ void finish_column(const t1_params& p, const int c, const double sum, const double* w_s,
                                              const double* scratch)
{
    const int    D       = p.D;
    const double samples = static_cast<double>(p.n_global);
    const double wsize   = static_cast<double>(D); // T == 1
    if (c == 0)
    {
        double sabs = 0.0, ssqr = 0.0;
        for (int q = 0; q < NWARPS; ++q)
        {
            sabs += scratch[q];
            ssqr += scratch[NWARPS + q];
        }
        double fx = __ddiv_rn(sum, samples);
        if (p.flavour == 0)
        {
            if (p.l1 > 0.0)
            {
                fx = __dadd_rn(fx, __dmul_rn(p.l1, __ddiv_rn(sabs, wsize)));
            }
            if (p.l2 > 0.0)
            {
                fx = __dadd_rn(fx, __dmul_rn(0.5, __ddiv_rn(ssqr, wsize)));
            }
        }
        else
        {
            fx = __dadd_rn(fx, __dadd_rn(__dmul_rn(p.l1, sabs), __dmul_rn(0.5, ssqr)));
        }
        *p.fx_out = fx;
    }
    else if (c == 1)
    {
        if (p.flavour == 0 && p.gx_out != nullptr)
        {
            p.gx_out[D] = __ddiv_rn(sum, samples); // gb = acc.m_gb / N
        }
    }
    else if (p.gx_out != nullptr)
    {
        const int    j  = c - 2;
        const double wj = w_s[j];
        double       gj = __ddiv_rn(sum, samples);
        if (p.flavour == 0)
        {
            if (p.l1 > 0.0)
            {
                gj = __dadd_rn(gj, __ddiv_rn(__dmul_rn(p.l1, sign_of(wj)), wsize));
            }
            if (p.l2 > 0.0)
            {
                gj = __dadd_rn(gj, __ddiv_rn(__dmul_rn(p.l2, wj), wsize));
            }
        }
        else
        {
            gj = __dadd_rn(gj, __dadd_rn(__dmul_rn(p.l1, sign_of(wj)), __dmul_rn(p.l2, wj)));
        }
        p.gx_out[j] = gj;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// tail (all warps): grid barrier, then this CTA folds its slice of columns over every CTA's partial row
// (sum_reduce, include/nano/core/reduce.h:23-31) and, when fuse_finish, applies the final scaling and the
// regularisation terms (src/linear/function.cpp:127-143,158-167 / src/function/mlearn/{ridge,lasso,elasticnet}.cpp).
// ---------------------------------------------------------------------------------------------------------------
template <int NWARPS, bool GRAD>
__device__ __forceinline__ void fold_partials(const t1_params& p, const double* w_s, double* scratch, const int warp,
                                              const int lane)
{
    const int D    = p.D;
    const int G    = static_cast<int>(gridDim.x);
    const int cols = GRAD ? (D + 2) : 2;

    grid_barrier(p.grid_counter, p.grid_target);

    // regularisation sums over w, needed by the CTA that owns column 0 (the loss sum)
    const bool owns_value = blockIdx.x == 0;
    if (owns_value && p.fuse_finish != 0)
    {
        double a = 0.0, q = 0.0;
        if (p.l1 > 0.0 || p.l2 > 0.0)
        {
            const double root = sqrt(p.l2);
            for (int j = threadIdx.x; j < D; j += NWARPS * 32)
            {
                const double w = w_s[j];
                a += fabs(w);
                const double sw = __dmul_rn(root, w);
                q               = __dadd_rn(q, __dmul_rn(sw, sw));
            }
        }
        a = warp_allreduce_sum(a);
        q = warp_allreduce_sum(q);
        if (lane == 0)
        {
            scratch[warp]          = a;
            scratch[NWARPS + warp] = q;
        }
        __syncthreads();
    }

    const int     cpc      = (cols + G - 1) / G; // columns per CTA
    const int     c_begin  = blockIdx.x * cpc;
    const int     c_end    = min(cols, c_begin + cpc);
    const int64_t stride   = D + 2;

    const int                parity = static_cast<int>(p.epoch & 1ULL);
    const unsigned long long slot   = static_cast<unsigned long long>(parity) * p.world + p.rank;

    for (int c = c_begin + warp; c < c_end; c += NWARPS)
    {
        // lanes stride over the CTAs, then a butterfly: a fixed order that depends on the grid size only
        double sum = 0.0;
        for (int r = lane; r < G; r += 32)
        {
            sum += __ldcg(p.partials + static_cast<int64_t>(r) * stride + c);
        }
        sum = warp_allreduce_sum(sum);
        if (p.world > 1)
        {
            // all-gather over peer memory: lane q stores this rank's column sum into rank q's mailbox
            if (lane < p.world)
            {
                p.peer_mail[lane][slot * p.col_stride + c] = sum;
            }
        }
        else if (lane == 0)
        {
            p.reduced[c] = sum;
            if (p.fuse_finish != 0)
            {
                finish_column<NWARPS>(p, c, sum, w_s, scratch);
            }
        }
    }
    if (p.world <= 1)
    {
        return;
    }

    // ---- cross-rank stage: publish this CTA's slice; either leave (the stream waits) or wait here and fold ----------
    // Both signals are always raised, so ranks in different wait modes still understand each other: a per-CTA flag
    // (for peers that wait inside their launch, slice by slice) and, by the LAST CTA of this launch to publish, ONE
    // ready word per peer (for peers whose stream waits: cuStreamWaitValue64 on [2][world] words behind the flags).
    __threadfence_system();
    __syncthreads();
    if (warp == 0)
    {
        if (lane < p.world)
        {
            // release: the slice written above is visible to rank `lane` before it sees the flag
            unsigned long long* flag = p.peer_flag[lane] + slot * G + blockIdx.x;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(p.epoch) : "memory");
        }
        __syncwarp();
        if (lane == 0)
        {
            // every CTA fenced its peer stores (system scope) before this update, and the release below is cumulative
            // over what the acquiring CTA observed: a peer that sees the ready word sees every slice of this rank
            unsigned long long arrived;
            asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], 1;" : "=l"(arrived) : "l"(p.publish_counter) : "memory");
            if (arrived + 1ULL == p.publish_target)
            {
                __threadfence_system();
                for (int q = 0; q < p.world; ++q)
                {
                    unsigned long long* ready = p.peer_flag[q] + 2ULL * p.world * G + slot;
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(ready), "l"(p.epoch) : "memory");
                }
            }
        }
        if (p.peer_wait != 0 && lane < p.world)
        {
            // acquire: wait for rank `lane`'s slice in MY mailbox (flags only grow; parity double-buffers the mail)
            const unsigned long long* mine =
                p.peer_flag[p.rank] + (static_cast<unsigned long long>(parity) * p.world + lane) * G + blockIdx.x;
            unsigned long long seen  = 0;
            const long long    start = clock64();
            do
            {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
                if (seen < p.epoch && clock64() - start > 4000000000LL) // ~2 s: a peer never launched
                {
                    *p.peer_error = 1;
                    break;
                }
            } while (seen < p.epoch);
        }
        __syncwarp();
    }
    if (p.peer_wait == 0)
    {
        return; // publish-only launch: peer_finish (finish_kernel with the mailbox) folds the slots
    }
    __syncthreads();

    const double* mail = p.peer_mail[p.rank] + static_cast<unsigned long long>(parity) * p.world * p.col_stride;
    for (int c = c_begin + threadIdx.x; c < c_end; c += NWARPS * 32)
    {
        double sum = 0.0;
        for (int q = 0; q < p.world; ++q) // rank order: identical bits on every rank
        {
            sum += __ldcg(mail + static_cast<int64_t>(q) * p.col_stride + c);
        }
        p.reduced[c] = sum;
        if (p.fuse_finish != 0)
        {
            finish_column<NWARPS>(p, c, sum, w_s, scratch);
        }
    }
}

// Warp roles: NCW consumer warps (whole warpgroups) + one producer warpgroup (its first warp feeds the ring, the
// other three idle until the tail). The producer warpgroup hands its registers to the consumers (setmaxnreg): the
// 128-register row + gradient slices of the widest variant then fit without spilling.
constexpr int kT1ProducerWarps = 4;
// 40 / 232 is the split that works: 24 / 240 would need the producer warps to grow back for the tail fold, and a
// setmaxnreg.inc of the producer warpgroup after a consumer .dec was observed to hang on B200 (round-1 notes)
constexpr int kT1ProducerRegs  = 40;

template <int NCW>
struct t1_consumer_regs
{
    // ptxas allocates the launch-bound maximum per thread when a kernel uses setmaxnreg (checked with -Xptxas -v:
    // 168 / 128 / 96 registers for 8 / 12 / 16 consumer warps); the CTA's pool is that times the thread count, and
    // asking for more than the pool holds would block setmaxnreg.inc forever.
    static constexpr int kThreads    = (NCW + kT1ProducerWarps) * 32;
    static constexpr int kRawRegs    = (65536 / kThreads) / 8 * 8;
    static constexpr int kLaunchRegs = kRawRegs > 248 ? 248 : kRawRegs;
    static constexpr int kPool       = kThreads * kLaunchRegs;
    static constexpr int kBudget     = ((kPool - kT1ProducerWarps * 32 * kT1ProducerRegs) / (NCW * 32)) / 8 * 8;
    static constexpr int value       = kBudget > 232 ? 232 : kBudget;
    // with few consumer warps every thread already owns more registers than setmaxnreg could hand over
    static constexpr bool kHandOver  = kLaunchRegs < value;

    static_assert(kLaunchRegs <= 255, "register budget out of range");
};

template <int V, int KC, int WS, int NCW, int RW, int ALT, bool GRAD>
__global__ void __launch_bounds__((NCW + kT1ProducerWarps) * 32, 1) vgrad_t1_kernel(const t1_params p)
{
    static_assert(V == 1 || V == 2, "V is 1 (64-bit loads) or 2 (128-bit loads)");
    static_assert(NCW % WS == 0, "consumer warps must split evenly into row groups");
    static_assert(NCW % 4 == 0, "consumer warps must form whole warpgroups (setmaxnreg is per warpgroup)");

    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar = full_bar + kT1MaxStages;
    double*   exchange  = reinterpret_cast<double*>(smem + 2 * kT1MaxStages * sizeof(uint64_t)); // [NG][2][WS][RW]
    double*   scratch   = reinterpret_cast<double*>(smem + 128 + kT1ExchangeBytes);              // tail: [2][warps]
    const int w_bytes   = static_cast<int>((static_cast<int64_t>(p.D) * 8 + 127) / 128 * 128);
    double*   w_s       = reinterpret_cast<double*>(smem + kT1HeaderSize);
    unsigned char* stage_base = smem + kT1HeaderSize + w_bytes;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int D    = p.D;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < p.stages; ++s)
        {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, NCW); // every consumer warp releases every stage (its own tiles after loading)
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp >= NCW)
    {
        // ===== producer warpgroup: one lane of its first warp feeds the ring =====
        if constexpr (t1_consumer_regs<NCW>::kHandOver)
        {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kT1ProducerRegs));
        }
        if (warp == NCW && lane == 0)
        {
            const int64_t  rps    = p.rows_per_stage;
            // X larger than L2 is streamed once per evaluation (evict first); a small dataset is kept for the next call
            const uint64_t policy = p.l2_resident != 0 ? l2_policy_evict_last() : l2_policy_evict_first();
            int            s      = 0;
            uint32_t       phase  = 1; // the first lap passes at once: no consumer has arrived yet
            for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
            {
                mbar_wait(empty_bar + s, phase);

                const int64_t  row0    = tile * rps;
                const int64_t  rows    = min(rps, p.N - row0);
                const uint32_t bytes_x = static_cast<uint32_t>((rows * D * 8 + 15) / 16 * 16);
                const uint32_t bytes_y = static_cast<uint32_t>((rows * 8 + 15) / 16 * 16);
                unsigned char* stage   = stage_base + static_cast<int64_t>(s) * p.stage_stride;

                mbar_arrive_expect_tx(full_bar + s, bytes_x + bytes_y);
                bulk_copy_g2s(stage, p.X + row0 * D, bytes_x, full_bar + s, policy);
                bulk_copy_g2s(stage + p.y_offset, p.Y + row0, bytes_y, full_bar + s, policy);
                if (++s == p.stages)
                {
                    s = 0;
                    phase ^= 1u;
                }
            }
        }
        __syncwarp();
    }
    else
    {
        // ===== consumer warpgroups =====
        // the producer is already streaming: the parameters reach shared memory while the first tiles are in flight
        for (int j = threadIdx.x; j < D; j += NCW * 32)
        {
            w_s[j] = p.w[j];
        }
        named_bar_sync(15, NCW * 32);
        if constexpr (t1_consumer_regs<NCW>::kHandOver)
        {
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(t1_consumer_regs<NCW>::value));
        }
        consume_rows<V, KC, WS, NCW, RW, ALT, GRAD>(p, full_bar, empty_bar, exchange, w_s, stage_base, warp, lane);
    }

    fold_partials<NCW + kT1ProducerWarps, GRAD>(p, w_s, scratch, warp, lane);
}

#endif // __CUDACC__
} // namespace nb200
