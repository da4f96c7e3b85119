// nb200_api.cu — the C ABI of libnb200.so (include/nb200.h): host-side orchestration of the sm_100a kernels.
//
// Boundary replaced: nano::function_t::operator() (src/function.cpp:246-272) dispatching into
// nano::linear::function_t::do_eval (src/linear/function.cpp:44-168) or function_{ridge,lasso,elasticnet}_t::do_eval
// (src/function/mlearn/). There is no CPU fallback in this file: every evaluation is a CUDA launch.
#include "../../include/nb200.h"

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <numeric>
#include <random>
#include <vector>

#include "common.cuh"
#include "loss.cuh"
#include "reduce.cuh"
#include "scale.cuh"
#include "synth.cuh"
#include "vgrad_gen.cuh"
#include "kernel_tables.cuh"
#include "batch.cuh"
#include "probe.cuh"
#include "gboost.cuh"

using namespace nb200;

// ---------------------------------------------------------------------------------------------------------------
// opaque handles
// ---------------------------------------------------------------------------------------------------------------
struct nb200_ctx
{
    int          device{0};
    cudaStream_t stream{nullptr};
    int          sm_count{0};
    int          smem_optin{0};
    int          l2_bytes{0};
    // grow-only scratch of nb200_evaluate (called once per trial and fold by linear_t::fit): a device arena for
    // (W, b, outputs, errors, values) and a pinned host buffer the two result vectors land in
    std::mutex   scratch_lock;
    char*        arena{nullptr};
    size_t       arena_bytes{0};
    char*        staging{nullptr};
    size_t       staging_bytes{0};
    // a device group (nb200_init_multi): one context per device, owned by this one; datasets pinned on it are split
    // in contiguous row blocks over the parts and objectives on them evaluate on every device
    std::vector<nb200_ctx*> parts;
};

struct nb200_data
{
    std::atomic<int> refs{1};
    int              device{0};
    int              sm_count{0};
    int              smem_optin{0};
    int              l2_bytes{0};
    double*          X{nullptr}; // [N, D] + 256 bytes of zeroed slack (bulk copies round tails up to 16 bytes)
    double*          Y{nullptr}; // [N, T] + slack
    int64_t          N{0}, D{0}, T{0};
    nb200::column_stats_t inputs_stats;  // filled by nb200_dataset_scale
    nb200::column_stats_t targets_stats;
    // a dataset on a device group: rows [row0[r], row0[r + 1]) live on parts[r] (X, Y above stay null)
    std::vector<nb200_data*> parts;
    std::vector<int64_t>     row0;
    // tensor map of X for the many-target forward (box = 16 features x 128 rows, 128-byte swizzle); x_map_ok == false
    // (odd D, or the driver refused) keeps the cp.async-fed kernel
    CUtensorMap x_map{};
    bool        x_map_ok{false};
};

namespace
{
// the kernel variants are instantiated in kernels_t1_{a,b,c,d}.cu and kernels_mt.cu (kernel_tables.cuh)
struct t1_table_t
{
    std::vector<t1_variant> entries;
    t1_table_t()
    {
        for (const auto group : {t1_group_a, t1_group_b, t1_group_c, t1_group_d})
        {
            int               count = 0;
            const t1_variant* first = group(&count);
            entries.insert(entries.end(), first, first + count);
        }
    }
};
const t1_table_t kT1Table;
const t1_variant* const kT1Variants    = kT1Table.entries.data();
const int               kNumT1Variants = static_cast<int>(kT1Table.entries.size());
constexpr int           kNumT1Heuristic = 17; // entries scanned by the fall-back heuristic

// (kernels_fd.cu: the row-split and the feature-split kernels; kernels_fl.cu: the feature-split kernel with the loss
// spread over the lanes, NSETS >= 10)
struct fd_table_t
{
    std::vector<fd_variant> entries;
    fd_table_t()
    {
        for (const auto group : {fd_variants, fl_variants})
        {
            int               count = 0;
            const fd_variant* first = group(&count);
            entries.insert(entries.end(), first, first + count);
        }
    }
};
const fd_table_t        kFdTable;
const fd_variant* const kFdVariants    = kFdTable.entries.data();
const int               kNumFdVariants = static_cast<int>(kFdTable.entries.size());

int               g_num_ft_variants = 0;
const ft_variant* const kFtVariants    = ft_variants(&g_num_ft_variants);
const int               kNumFtVariants = g_num_ft_variants;

int               g_num_mt_variants = 0;
const mt_variant* const kMtVariants    = mt_variants(&g_num_mt_variants);
const int               kNumMtVariants = g_num_mt_variants;

int env_int(const char* name, const int fallback)
{
    const char* value = std::getenv(name);
    return (value != nullptr && *value != '\0') ? std::atoi(value) : fallback;
}

// launch geometry of the few-target one-pass kernel (vgrad_ft.cuh)
struct ft_plan
{
    int     variant{0};
    int     grid{0};
    int     rows_per_stage{0};
    int     stages{0};
    int     stage_stride{0};
    int     y_offset{0};
    int     smem_bytes{0};
    int64_t num_tiles{0};
};

// an exact-T variant whose 512 * KC columns cover D (even D: rows start on 16-byte boundaries); false otherwise
bool plan_ft(const int64_t N, const int64_t D, const int64_t T, const int sm_count, const int smem_optin, ft_plan& plan)
{
    static const bool skip_two_sets = env_int("NB200_FT_ONE_SET", 0) != 0; // A/B knob (tools/gpu_round40.sh)
    if (!(N > 0 && D > 0 && D % 2 == 0))
    {
        return false;
    }
    for (int v = 0; v < kNumFtVariants; ++v)
    {
        const ft_variant& var = kFtVariants[v];
        if (var.T != T || D > 512LL * var.KC / var.NS || (skip_two_sets && var.NS > 1))
        {
            continue;
        }
        const int64_t row_b = D * 8;
        const int64_t avail = static_cast<int64_t>(smem_optin) - kFtHeaderSize;
        // ~64 KiB stages of whole iterations; rows * T must be even (the Y tile starts on a 16-byte boundary)
        int64_t rps = std::max<int64_t>(1, 65536 / (row_b + T * 8)) / var.RW * var.RW;
        rps         = std::max<int64_t>(rps, var.RW);
        if ((rps * T) % 2 != 0)
        {
            rps += var.RW;
        }
        const int64_t x_bytes = round_up(rps * row_b, 128);
        const int64_t y_bytes = round_up(rps * T * 8, 128);
        const int64_t stride  = x_bytes + y_bytes;
        const int64_t stages  = std::min<int64_t>(kFtMaxStages, avail / stride);
        if (stages < 2)
        {
            return false;
        }
        plan.variant        = v;
        plan.rows_per_stage = static_cast<int>(rps);
        plan.stages         = static_cast<int>(std::min<int64_t>(stages, env_int("NB200_FT_STAGES", 3)));
        plan.stage_stride   = static_cast<int>(stride);
        plan.y_offset       = static_cast<int>(x_bytes);
        plan.smem_bytes     = static_cast<int>(kFtHeaderSize + plan.stages * stride);
        plan.num_tiles      = (N + rps - 1) / rps;
        plan.grid           = static_cast<int>(std::min<int64_t>(plan.num_tiles, sm_count));
        return plan.grid > 0;
    }
    return false;
}

// launch geometry of the one-pass tensor-pipe kernel for a handful of targets (vgrad_fd.cuh)
struct fd_plan
{
    int     variant{0};
    int     grid{0};
    int     slice{0}, ld{0};
    int     stages{0};
    int     stage_stride{0}, y_offset{0}, w_offset{0}, stage_offset{0};
    int     smem_bytes{0};
    int64_t num_tiles{0};
};

// W (T x ld) and at least two stages of 8 padded rows must fit next to the header; false otherwise
// lane_loss: 0 never the kernels of vgrad_fl.cuh, 1 where they win, 2 wherever they fit
bool plan_fd(const int64_t N, const int64_t D, const int64_t T, const int sm_count, const int smem_optin, fd_plan& plan,
             const bool row_split_ok = false, const int lane_loss = 0)
{
    static const bool fd_one_set = env_int("NB200_FD_ONE_SET", 0) != 0; // A/B knob (tools/gpu_round44.sh)
    const int         fl_tpi     = env_int("NB200_FL_TPI", 0);          // A/B knob: tiles per iteration (0: by room)
    if (!(N > 0 && D > 0 && D % 2 == 0 && T >= 2))
    {
        return false;
    }
    // the feature-split kernels with the loss spread over the lanes (vgrad_fl.cuh): rows longer than the row-split
    // kernel takes. Shared rows are as short as the fragment loads allow (ld = 4 mod 8), and 512 bytes of slack follow
    // the last stage (the last warps' slices may reach past ld). Three forms, tried in the order that measured
    // fastest (profiles/r02_lane_loss_sweep.jsonl): 'L2' / 'L1' = two barriers per iteration of two / one 8-row tiles
    // (they hold the iteration's stages: room for as many again in flight is asked for, at least two), 'P' = the
    // pipeline with the loss on its own warps (one tile per iteration, two stages in flight, one when nothing else fits).
    //   rows up to 512 features: L2 (the per-iteration latency is shared by 16 rows), then P, then L1;
    //   longer rows: P; one target tile (T <= 8) and 512 < D <= 800: vgrad_fd.cuh's two alternating warp sets stay.
    // NB200_FL=2: wherever a form fits; NB200_FL_PIPE=0 / 2: never / always P first; NB200_FL_TPI: only that L form.
    const int fl_pipe = env_int("NB200_FL_PIPE", 1);
    if (lane_loss > 0 && T <= 16 && D > 256)
    {
        const int   nt8 = T > 8 ? 2 : 1;
        const char* order;
        if (fl_pipe >= 2)
        {
            order = "P21";
        }
        else if (D <= 512)
        {
            order = "2P1";
        }
        else if (nt8 == 2 || D > 800 || lane_loss >= 2)
        {
            order = "P21";
        }
        else
        {
            order = "";
        }
        for (const char* kind = order; *kind != '\0'; ++kind)
        {
            for (int tight = 0; tight < 2; ++tight)
            {
                for (int v = 0; v < kNumFdVariants; ++v)
                {
                    const fd_variant& var  = kFdVariants[v];
                    const bool        pipe = var.NSETS >= 20;
                    const int         tpi  = var.NSETS - (pipe ? 20 : 10);
                    if (var.NSETS < 10 || var.NT8 != nt8 || 64LL * var.NS8 < D || pipe != (*kind == 'P') ||
                        (!pipe && tpi != *kind - '0') || (!pipe && fl_tpi > 0 && tpi != fl_tpi) ||
                        (pipe && fl_pipe == 0) || (tight == 1 && !pipe))
                    {
                        continue;
                    }
                    const int64_t slice   = 8LL * var.NS8;
                    const int64_t ld      = (D + 3) / 8 * 8 + 4; // the smallest value = 4 (mod 8) that is >= D
                    const int64_t w_bytes = round_up(T * ld * 8, 128);
                    const int64_t x_bytes = kFdRows * ld * 8;
                    const int64_t stride  = round_up(x_bytes + round_up(kFdRows * T * 8, 128), 128);
                    const int64_t offset  = var.header + w_bytes;
                    const int64_t stages =
                        std::min<int64_t>(std::min<int64_t>(kFlMaxStages, env_int("NB200_FD_STAGES", kFlMaxStages)),
                                          (smem_optin - 512 - offset) / stride);
                    if (stages < (tight == 1 ? 2 : std::max(2 * tpi, tpi + 2)))
                    {
                        continue;
                    }
                    plan.variant      = v;
                    plan.slice        = static_cast<int>(slice);
                    plan.ld           = static_cast<int>(ld);
                    plan.stages       = static_cast<int>(stages);
                    plan.stage_stride = static_cast<int>(stride);
                    plan.y_offset     = static_cast<int>(x_bytes);
                    plan.w_offset     = var.header;
                    plan.stage_offset = static_cast<int>(offset);
                    plan.smem_bytes   = static_cast<int>(offset + stages * stride + 512);
                    plan.num_tiles    = (N + tpi * kFdRows - 1) / (tpi * kFdRows); // iterations
                    plan.grid         = static_cast<int>(std::min<int64_t>(plan.num_tiles, sm_count));
                    return plan.grid > 0;
                }
            }
        }
    }
    for (int v = 0; v < kNumFdVariants; ++v)
    {
        const fd_variant& var = kFdVariants[v];
        if (var.NSETS >= 10)
        {
            continue;
        }
        if (var.NSETS < 0)
        {
            // the row-split kernel (vgrad_fr.cuh): groups of 8 TPW tiles, one 8-row tile per ring stage, at least a
            // group and two more tiles in flight
            const int64_t tiles = fr_group_tiles(-var.NSETS);
            if (!row_split_ok || 8 * var.NT8 < T || 64LL * var.NS8 < D)
            {
                continue;
            }
            const int64_t slice   = 8LL * var.NS8;
            const int64_t ld      = kFdWarps * slice + 4;
            const int64_t w_bytes = round_up(T * ld * 8, 128);
            const int64_t x_bytes = kFdRows * ld * 8;
            const int64_t stride  = round_up(x_bytes + round_up(kFdRows * T * 8, 128), 128);
            const int64_t offset  = var.header + w_bytes;
            const int64_t stages  = std::min<int64_t>(kFrMaxStages, (smem_optin - offset) / stride);
            if (stages < tiles + 2)
            {
                continue;
            }
            plan.variant      = v;
            plan.slice        = static_cast<int>(slice);
            plan.ld           = static_cast<int>(ld);
            plan.stages       = static_cast<int>(stages);
            plan.stage_stride = static_cast<int>(stride);
            plan.y_offset     = static_cast<int>(x_bytes);
            plan.w_offset     = var.header;
            plan.stage_offset = static_cast<int>(offset);
            plan.smem_bytes   = static_cast<int>(offset + stages * stride);
            plan.num_tiles    = (N + tiles * kFdRows - 1) / (tiles * kFdRows); // groups
            plan.grid         = static_cast<int>(std::min<int64_t>(plan.num_tiles, sm_count));
            return plan.grid > 0;
        }
        if (8 * var.NT8 < T || (64LL / var.NSETS) * var.NS8 < D || (fd_one_set && var.NSETS > 1))
        {
            continue;
        }
        const int64_t slice    = 8LL * var.NS8;
        const int64_t ld       = (kFdWarps / var.NSETS) * slice + 4; // = 4 mod 16: conflict-free fragment loads
        const int64_t w_bytes  = round_up(T * ld * 8, 128);
        const int64_t x_bytes  = kFdRows * ld * 8;
        const int64_t y_bytes  = round_up(kFdRows * T * 8, 128);
        const int64_t stride   = round_up(x_bytes + y_bytes, 128);
        const int64_t offset   = var.header + w_bytes;
        const int64_t stages   = std::min<int64_t>(std::min<int64_t>(kFdMaxStages, env_int("NB200_FD_STAGES", kFdMaxStages)),
                                                 (smem_optin - offset) / stride);
        if (stages < 2)
        {
            return false;
        }
        plan.variant      = v;
        plan.slice        = static_cast<int>(slice);
        plan.ld           = static_cast<int>(ld);
        plan.stages       = static_cast<int>(stages);
        plan.stage_stride = static_cast<int>(stride);
        plan.y_offset     = static_cast<int>(x_bytes);
        plan.w_offset     = var.header;
        plan.stage_offset = static_cast<int>(offset);
        plan.smem_bytes   = static_cast<int>(offset + stages * stride);
        plan.num_tiles    = (N + kFdRows - 1) / kFdRows;
        plan.grid         = static_cast<int>(std::min<int64_t>(plan.num_tiles, sm_count));
        return plan.grid > 0;
    }
    return false;
}

// launch geometry of the tensor-pipe kernels for an (N, D, T) problem
struct mt_geometry
{
    int variant{0};
    int fwd_grid{0}, fwd_stages{0}, fwd_smem{0};
    int tma_stages{0}, tma_smem{0}; // the tensor-map-fed forward (0 stages: does not fit)
    int splits{0}, bwd_grid{0}, bwd_stages{0}, bwd_smem{0}, t_ld{0};
};

// T in [2, 128], D even (rows of X start on 16-byte boundaries; dL rows are padded to an even stride); false when the
// shape is not covered
bool plan_mt(const int64_t N, const int64_t D, const int64_t T, const int sm_count, const int smem_optin,
             mt_geometry& geo)
{
    if (!(T >= 2 && T <= 128 && D % 2 == 0 && D > 0 && N > 0))
    {
        return false;
    }
    int v = 0;
    while (v < kNumMtVariants && kMtVariants[v].NT8 * 8 < T)
    {
        ++v;
    }
    const int64_t TP = kMtVariants[v].NT8 * 8;
    // forward: header (barriers + the [64][TP + 2] output tile, which also hosts the final fold + alignment slack)
    const int64_t fwd_header = 1024 + 8 * (kMtBM / 2) * (TP + 2) + 512;
    const int64_t fwd_stage  = (kMtBM + TP) * kMtLd * 8;
    int64_t       fwd_stages = std::min<int64_t>(kMtMaxStages, (smem_optin - fwd_header) / fwd_stage);
    if (const int forced = env_int("NB200_MT_STAGES", 0); forced >= 2)
    {
        fwd_stages = std::min<int64_t>(fwd_stages, forced); // tuning knob
    }
    // backward: dL tile rows padded to t_ld = 4 (mod 16) doubles, X tile rows to 132 doubles
    int64_t t_ld = T + (T & 1);
    while (t_ld % 16 != 4)
    {
        t_ld += 2;
    }
    const int64_t bwd_stage  = (round_up(kMtBwdRows * t_ld, 16) + kMtBwdRows * kMtBwdLd) * 8;
    const int64_t bwd_stages = std::min<int64_t>(kMtMaxStages, (smem_optin - 256) / bwd_stage);
    const int64_t ctiles     = (D + kMtBwdCols - 1) / kMtBwdCols;
    const int64_t chunks     = (N + kMtBwdRows - 1) / kMtBwdRows;
    int64_t       splits     = std::max<int64_t>(1, std::min<int64_t>(chunks, sm_count / std::max<int64_t>(ctiles, 1)));
    // one CTA per SM leaves SMs idle when ctiles does not divide their number (D = 2048: 16 x 9 = 144 of 148). With
    // enough rows the row ranges are cut finer instead, into the smallest count that makes the grid a whole number of
    // waves (16 x 37 = 4 x 148): CTAs are handed out as SMs free up, so every SM works until the end
    if (ctiles * splits < sm_count)
    {
        const int64_t unit = sm_count / std::gcd<int64_t>(sm_count, ctiles);
        if (chunks >= unit * 64 && unit * T * D * 8 <= (int64_t{256} << 20) && env_int("NB200_MT_WAVES", 1) != 0)
        {
            splits = unit;
        }
    }
    if (fwd_stages < 2 || bwd_stages < 2 || ctiles > sm_count)
    {
        return false;
    }
    geo.variant    = v;
    geo.fwd_grid   = static_cast<int>(std::min<int64_t>((N + kMtBM - 1) / kMtBM, sm_count));
    geo.fwd_stages = static_cast<int>(fwd_stages);
    geo.fwd_smem   = static_cast<int>(fwd_header + fwd_stages * fwd_stage);
    // tensor-map-fed forward: barriers, the [128][TP + 2] output tile (1 KiB granules: the stages behind it are 128-byte
    // swizzle atoms), then stages of a dense X tile + the packed W tile
    const int64_t tma_otile = round_up(kMtBM * (TP + 2) * 8, 1024);
    const int64_t tma_stage = kMtXTileBytes + round_up(TP * kMtLd * 8, 1024);
    int64_t       tma_stages = std::min<int64_t>(kMtMaxStages, (smem_optin - kMtTmaHeader - tma_otile) / tma_stage);
    if (const int forced = env_int("NB200_MT_STAGES", 0); forced >= 2)
    {
        tma_stages = std::min<int64_t>(tma_stages, forced);
    }
    geo.tma_stages = tma_stages >= 2 ? static_cast<int>(tma_stages) : 0;
    geo.tma_smem   = static_cast<int>(kMtTmaHeader + tma_otile + std::max<int64_t>(tma_stages, 0) * tma_stage);
    geo.splits     = static_cast<int>(splits);
    geo.bwd_grid   = static_cast<int>(ctiles * splits);
    geo.bwd_stages = static_cast<int>(bwd_stages);
    geo.bwd_smem   = static_cast<int>(256 + bwd_stages * bwd_stage);
    geo.t_ld       = static_cast<int>(t_ld);
    return true;
}

// scratch of nb200_vgrad_batch (K single-target models in one pass), sized for the largest K seen so far
struct batch_state
{
    int64_t     Kp{0}; // padded (even) model count the buffers are sized for
    mt_geometry geo{};
    bool        one_pass{false}; // the one-pass tensor-pipe kernel (vgrad_fd.cuh) covers (D, Kp)
    fd_plan     fd{};
    double *    d_x{nullptr}, *d_W{nullptr}, *d_Wp{nullptr}, *d_b{nullptr}, *d_l{nullptr}, *d_dL{nullptr};
    double *    d_partials_fb{nullptr}, *d_partials_gw{nullptr}, *d_red_fb{nullptr}, *d_red_gw{nullptr};
    double *    h_in{nullptr}, *h_out{nullptr}, *h_out_dev{nullptr};

    void release()
    {
        cudaFree(d_x);
        cudaFree(d_W);
        cudaFree(d_Wp);
        cudaFree(d_b);
        cudaFree(d_l);
        cudaFree(d_dL);
        cudaFree(d_partials_fb);
        cudaFree(d_partials_gw);
        cudaFree(d_red_fb);
        cudaFree(d_red_gw);
        cudaFreeHost(h_in);
        cudaFreeHost(h_out);
        *this = batch_state{};
    }
};

constexpr int kTimingRing = 64; // event pairs kept in flight when pass timing is enabled

struct t1_plan
{
    int     variant{-1};
    int     grid{0};
    int     rows_per_group{0};
    int     rows_per_stage{0};
    int     stages{0};
    int     stage_stride{0};
    int     y_offset{0};
    int     smem_bytes{0};
    int64_t num_tiles{0};
};

// Shared-memory geometry of variant `v` for a (N, D) problem; false if it cannot cover D or does not fit.
bool plan_t1(const int v, const int64_t N, const int64_t D, const int sm_count, const int smem_optin, t1_plan& plan)
{
    const t1_variant& var = kT1Variants[v];
    if (var.V == 2 && (D % 2) != 0)
    {
        return false;
    }
    const int64_t capacity = 32LL * var.V * var.KC * var.WS;
    if (D > capacity)
    {
        return false;
    }
    const int     NG       = var.NCW / var.WS / var.ALT; // row groups that share one stage
    const int     NG_fold  = var.NCW / var.WS;           // partial gradient rows folded through the stage area
    const int64_t w_bytes  = round_up(D * 8, 128);
    const int64_t avail    = static_cast<int64_t>(smem_optin) - kT1HeaderSize - w_bytes;
    const int64_t row_b    = D * 8;
    const int64_t target   = env_int("NB200_STAGE_BYTES", 65536);
    // TWO stages of ~64 KiB in flight per SM: measured on B200 (profiles/r02_stages_sweep.jsonl), a third stage costs
    // 5-8 % of the bandwidth at full clocks for every width (7.28 -> 6.8 TB/s: 28 MB of evict-first lines in flight over
    // the 148 SMs instead of 19 MB) and buys nothing at the power cap; 128 KiB covers latency x bandwidth per SM
    // (~2 us x 49 GB/s = 98 KB). NB200_STAGES overrides (tuning).
    const int     stages_x = env_int("NB200_STAGES", 2);

    int64_t rpg = std::max<int64_t>(1, target / (NG * row_b));
    if (rpg > var.RW)
    {
        rpg = rpg / var.RW * var.RW; // whole iterations of RW rows
    }
    if ((NG * rpg) % 2 != 0)
    {
        rpg += 1; // tiles must start on 16-byte boundaries for any D
    }
    // the epilogue folds [NG][WS*CPW] partial gradients through the stage area
    const int64_t epilogue = (static_cast<int64_t>(NG_fold) * capacity + 2 * NG_fold) * 8;
    for (;;)
    {
        const int64_t rps      = NG * rpg;
        const int64_t x_bytes  = round_up(rps * row_b, 128);
        const int64_t y_bytes  = round_up(rps * 8, 128);
        const int64_t stride   = x_bytes + y_bytes;
        int64_t       stages   = std::min<int64_t>(std::min(kT1MaxStages, stages_x), avail / stride);
        if (stages >= 2 && stages * stride >= epilogue)
        {
            plan.variant        = v;
            plan.rows_per_group = static_cast<int>(rpg);
            plan.rows_per_stage = static_cast<int>(rps);
            plan.stages         = static_cast<int>(stages);
            plan.stage_stride   = static_cast<int>(stride);
            plan.y_offset       = static_cast<int>(x_bytes);
            plan.smem_bytes     = static_cast<int>(kT1HeaderSize + w_bytes + stages * stride);
            plan.num_tiles      = (N + rps - 1) / rps;
            plan.grid           = static_cast<int>(std::min<int64_t>(plan.num_tiles, sm_count));
            return plan.grid > 0;
        }
        // shrink the stage (keeping NG * rpg even) until it fits
        const int64_t step = (NG % 2 == 0) ? 1 : 2;
        if (rpg <= step)
        {
            return false;
        }
        rpg -= step;
        if ((NG * rpg) % 2 != 0)
        {
            return false;
        }
    }
}
} // namespace

struct nb200_obj
{
    nb200_data*  data{nullptr};
    int          device{0};
    cudaStream_t stream{nullptr};     // the stream every launch and copy of this objective is ordered on
    cudaStream_t own_stream{nullptr}; // created with the objective; `stream` may be re-pointed at a caller's stream
    int          loss{0};
    double       loss_param{0.0};
    int          flavour{0};
    double       l1{0.0}, l2{0.0};
    int64_t      n{0};

    // device scratch
    double* d_x{nullptr};          // [n] parameters
    double* d_fixed_bias{nullptr}; // [T] (SYNTH)
    double* d_out{nullptr};        // [n + 1] gradient, then the function value
    double* d_reduced{nullptr};    // [1 + T + T*D]
    double* d_partials{nullptr};   // T1: [grid, D + 2]; generic: forward [blocks, 1 + T]
    double* d_partials_gw{nullptr}; // generic: [splits, T*D]
    double* d_dL{nullptr};         // generic: [N, T]
    double* d_Wp{nullptr};         // tensor-pipe forward: W packed per k-chunk [D / 16][T][20]
    unsigned long long* d_grid_counter{nullptr}; // arrival counter of the fused kernel's grid barrier
    unsigned long long  grid_epoch{0};           // launches issued on that counter under the current plan
    // pinned host staging; h_out is MAPPED: kernels write (g, f) straight into it through h_out_dev
    double* h_x{nullptr};
    double* h_out{nullptr};
    double* h_out_dev{nullptr};
    std::vector<double> fixed_bias;

    // plan
    bool    use_t1{false};
    t1_plan t1{};
    int     gen_blocks{0};
    int     gen_splits{0};
    int     gen_smem{0};
    // few-target one-pass plans (vgrad_ft.cuh: CUDA cores; vgrad_fd.cuh: tensor pipe)
    bool    use_ft{false};
    ft_plan ft{};
    bool    use_fd{false};
    fd_plan fd{};
    // many-target tensor-pipe plan (vgrad_mt.cuh)
    bool    use_mt{false};
    int     mt_variant{0}; // index into kMtVariants
    int     mt_fwd_grid{0}, mt_fwd_stages{0}, mt_fwd_smem{0};
    bool    mt_tma{false}; // the forward runs mt_forward_tma_kernel (mt_fwd_stages / smem are then that kernel's)
    int     mt_splits{0}, mt_bwd_grid{0}, mt_bwd_stages{0}, mt_bwd_smem{0}, mt_t_ld{0};

    // fused cross-rank all-reduce over peer memory (single-target path): mailbox + flags, see vgrad_t1.cuh
    int                 peer_world{1};
    int                 peer_rank{0};
    int                 peer_col_stride{0};
    unsigned long long  peer_epoch{0};
    unsigned char*      peer_local{nullptr};               // this rank's mailbox allocation (cudaMalloc, IPC-exported)
    unsigned char*      peer_base[kT1MaxPeers] = {nullptr}; // every rank's mailbox as seen from this device
    int*                h_peer_error{nullptr};             // mapped host flag raised by the kernel on a peer timeout
    int*                d_peer_error{nullptr};
    // how an evaluation meets its peers (vgrad_t1.cuh): 0 = the launch publishes and leaves, the STREAM waits for the
    // peers' ready words and peer-finish folds the slots (no SM is held while waiting: evaluations of different
    // objectives can never cross-wait); 1 = the launch itself spins on the peers' flags (one kernel, holds the GPU)
    int                 peer_wait{0};
    bool                peer_direct{false};                // peers live in this process (plain pointers, nothing to close)
    int                 peer_grid{0};                      // grid the mailbox was sized for
    unsigned long long  publish_epoch{0};                  // publish-only launches issued on d_grid_counter[1]
    bool                peer_pending{false};               // a stream wait on the peers may be outstanding
    unsigned long long* h_release{nullptr};                // pinned: the value a timed-out wait is released with
    cudaStream_t        release_stream{nullptr};

    batch_state batch; // nb200_vgrad_batch

    // an objective on a device group: one objective per device (owned), evaluated together; parts[0] talks to the host
    std::vector<nb200_obj*>  parts;
    bool                     group_fused{false};   // the parts publish into each other's mailboxes (distinct devices)
    bool                     peer_passive{false};  // a part other than the first: publishes, never folds
    double*                  group_gather{nullptr}; // two-phase groups: [parts][1 + T + T*D] on the first device
    std::vector<cudaEvent_t> group_events;          // [parts]: x is ready (slot 0) / part r's sums are ready

    // state of a two-phase evaluation
    bool    partial_has_grad{false};
    const double* partial_x{nullptr};

    // host-side wall clock of the public call, microseconds: [stage x + H2D | enqueue | wait + hand-over]
    double  host_us[3] = {0.0, 0.0, 0.0};
    int64_t host_calls{0};

    // statistics
    int64_t     launches{0};
    bool        timing{false};
    int         pass_pending{0}; // event pairs recorded and not yet read back
    std::vector<cudaEvent_t> ev0, ev1;
    double      last_pass_ms{-1.0};
    double      pass_ms_total{0.0};
    int64_t     pass_count{0};
    std::vector<float> pass_samples; // the passes since the last reset, one by one (bounded)
};

namespace
{
// cuStreamWaitValue64 through the runtime's driver entry-point query (libnb200.so does not link libcuda): the stream
// waits until the 64-bit word at `address` is >= `value` — no SM is held while it waits
using stream_wait_value_fn = int (*)(cudaStream_t stream, unsigned long long address, unsigned long long value,
                                     unsigned int flags);
stream_wait_value_fn stream_wait_value()
{
    static const stream_wait_value_fn fn = []() -> stream_wait_value_fn
    {
        void*                           entry = nullptr;
        cudaDriverEntryPointQueryResult found = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue64", &entry, cudaEnableDefault, &found) != cudaSuccess ||
            found != cudaDriverEntryPointSuccess)
        {
            cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<stream_wait_value_fn>(entry);
    }();
    return fn;
}

// forget the peers: unmap their mailboxes, free this rank's (a new plan may change the grid the mailbox is sized for,
// and flags of an earlier connection must never be mistaken for fresh mail)
void drop_peers(nb200_obj* obj)
{
    for (int r = 0; r < kT1MaxPeers; ++r)
    {
        if (obj->peer_base[r] != nullptr && obj->peer_base[r] != obj->peer_local && !obj->peer_direct)
        {
            cudaIpcCloseMemHandle(obj->peer_base[r]);
        }
        obj->peer_base[r] = nullptr;
    }
    cudaFree(obj->peer_local);
    obj->peer_local    = nullptr;
    obj->peer_world    = 1;
    obj->peer_rank     = 0;
    obj->peer_epoch    = 0;
    obj->publish_epoch = 0;
    obj->peer_direct   = false;
    obj->peer_grid     = 0;
}

void free_obj(nb200_obj* obj)
{
    if (obj == nullptr)
    {
        return;
    }
    cudaSetDevice(obj->device);
    cudaFree(obj->d_x);
    cudaFree(obj->d_fixed_bias);
    cudaFree(obj->d_out);
    cudaFree(obj->d_reduced);
    cudaFree(obj->d_partials);
    cudaFree(obj->d_partials_gw);
    cudaFree(obj->d_dL);
    cudaFree(obj->d_Wp);
    cudaFree(obj->d_grid_counter);
    for (nb200_obj* part : obj->parts)
    {
        free_obj(part);
    }
    for (cudaEvent_t event : obj->group_events)
    {
        cudaEventDestroy(event);
    }
    cudaFree(obj->group_gather);
    drop_peers(obj);
    cudaFreeHost(obj->h_peer_error);
    cudaFreeHost(obj->h_release);
    if (obj->release_stream != nullptr)
    {
        cudaStreamDestroy(obj->release_stream);
    }
    obj->batch.release();
    cudaFreeHost(obj->h_x);
    cudaFreeHost(obj->h_out);
    for (cudaEvent_t event : obj->ev0)
    {
        cudaEventDestroy(event);
    }
    for (cudaEvent_t event : obj->ev1)
    {
        cudaEventDestroy(event);
    }
    if (obj->own_stream != nullptr)
    {
        cudaStreamDestroy(obj->own_stream);
    }
    if (obj->data != nullptr)
    {
        nb200_dataset_release(obj->data);
    }
    delete obj;
}

int setup_plan(nb200_obj* obj, const int forced_variant)
{
    const nb200_data* data = obj->data;
    const int64_t     N = data->N, D = data->D, T = data->T;

    // release the previous plan's scratch
    cudaFree(obj->d_partials);
    cudaFree(obj->d_partials_gw);
    cudaFree(obj->d_dL);
    cudaFree(obj->d_Wp);
    obj->d_partials = obj->d_partials_gw = obj->d_dL = obj->d_Wp = nullptr;
    obj->use_t1 = false;
    NB200_CUDA_TRY(cudaMemset(obj->d_grid_counter, 0, 2 * sizeof(unsigned long long)));
    obj->grid_epoch = 0;
    drop_peers(obj); // a new plan may change the grid: peers must export and connect again

    const bool force_generic = forced_variant == -2 || env_int("NB200_FORCE_GENERIC", 0) != 0;
    if (T == 1 && !force_generic && D <= (1 << 20))
    {
        t1_plan plan;
        bool    ok = false;
        if (forced_variant >= 0)
        {
            NB200_REQUIRE(forced_variant < kNumT1Variants, "objective: unknown kernel variant");
            ok = plan_t1(forced_variant, N, D, data->sm_count, data->smem_optin, plan);
            NB200_REQUIRE(ok, "objective: the requested kernel variant cannot cover this (N, D)");
        }
        else
        {
            // measured on B200 (profiles/r01_sustained_*.jsonl, r02_variants_d1024.jsonl): the board sits at its 1000 W
            // power cap while it streams fp64 from HBM, so the variant with the fewest instructions per row wins the
            // steady state — several rows per warp iteration (one reduction and one loss chain per RW rows), two
            // alternating sets of four consumer warps for rows up to 4 KiB, rows split over warp pairs (37) for rows up to
            // 8 KiB, 16 KiB rows: variant 39. With two stages in flight these are also the fastest variants at full
            // clocks (D = 1024, logistic: 37 runs 1.181 ms cool / 1.219 ms at the cap, whole-row variant 27 1.182 / 1.346),
            // so one choice per width serves a single solve and a long tuner run alike.
            const int preferred[] = {36, 35, 34, 33, 37, 27, 39};
            for (const int v : preferred)
            {
                if (!ok)
                {
                    ok = plan_t1(v, N, D, data->sm_count, data->smem_optin, plan);
                }
            }
            for (int v = 0; v < kNumT1Heuristic && !ok; ++v)
            {
                ok = plan_t1(v, N, D, data->sm_count, data->smem_optin, plan);
            }
        }
        if (ok)
        {
            const t1_variant& var = kT1Variants[plan.variant];
            NB200_CUDA_TRY(cudaFuncSetAttribute(var.grad, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes));
            NB200_CUDA_TRY(cudaFuncSetAttribute(var.value, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes));
            NB200_CUDA_TRY(cudaMalloc(&obj->d_partials, static_cast<size_t>(plan.grid) * (D + 2) * sizeof(double)));
            obj->t1     = plan;
            obj->use_t1 = true;
            NB200_CUDA_TRY(cudaDeviceSynchronize()); // the memsets above ran on the legacy stream
            return NB200_OK;
        }
    }

    // few targets: one pass over X. forced_variant: -3 the generic kernels, -4 the two-pass tensor-pipe pair, -5 the
    // one-pass tensor-pipe kernel, -6 the one-pass CUDA-core kernel; heuristic: CUDA cores for T = 2 (HBM-bound
    // there), the tensor pipe from T = 5, in between by width (profiles/r01_few_targets_*.jsonl, r01_handful_*.jsonl)
    obj->use_ft = false;
    obj->use_fd = false;
    obj->use_mt = false;
    const bool one_pass = !force_generic && forced_variant != -3 && forced_variant != -4 && T >= 2;
    fd_plan    handful;
    // (T = 3, 4: the CUDA-core kernel wins where D fills its 512- / 1024-column layouts, the tensor-pipe one elsewhere)
    const bool ft_sweet  = (D > 384 && D <= 512) || D > 896;
    const bool prefer_fd = D >= 200 && (T >= 5 || (T >= 3 && !ft_sweet));
    // short rows: the row-split tensor-pipe kernel (vgrad_fr.cuh; NB200_FR=0: A/B knob); also for T = 2, where it beats
    // the CUDA-core kernel at these widths (256 x 2: 0.53 against 0.73 ms, 64 x 2: 1.34 against 2.60 at 2 GiB of X)
    const bool short_rows = D <= 256 && T >= 2 && T <= 16 && env_int("NB200_FR", 1) != 0;
    if (one_pass && forced_variant != -6 && (prefer_fd || short_rows || forced_variant == -5) &&
        plan_fd(N, D, T, data->sm_count, data->smem_optin, handful, short_rows, env_int("NB200_FL", 1)))
    {
        const fd_variant& var = kFdVariants[handful.variant];
        NB200_CUDA_TRY(cudaFuncSetAttribute(var.grad, cudaFuncAttributeMaxDynamicSharedMemorySize, handful.smem_bytes));
        NB200_CUDA_TRY(cudaFuncSetAttribute(var.value, cudaFuncAttributeMaxDynamicSharedMemorySize, handful.smem_bytes));
        NB200_CUDA_TRY(
            cudaMalloc(&obj->d_partials, static_cast<size_t>(handful.grid) * (1 + T + T * D) * sizeof(double)));
        NB200_CUDA_TRY(cudaDeviceSynchronize());
        obj->fd     = handful;
        obj->use_fd = true;
        return NB200_OK;
    }
    ft_plan few;
    // (rows shorter than ~1.6 KiB with five or more targets: the two-pass pair is faster than either one-pass kernel)
    if (one_pass && forced_variant != -5 && (T <= 4 || D >= 200 || forced_variant == -6) &&
        plan_ft(N, D, T, data->sm_count, data->smem_optin, few))
    {
        const ft_variant& var = kFtVariants[few.variant];
        NB200_CUDA_TRY(cudaFuncSetAttribute(var.grad, cudaFuncAttributeMaxDynamicSharedMemorySize, few.smem_bytes));
        NB200_CUDA_TRY(cudaFuncSetAttribute(var.value, cudaFuncAttributeMaxDynamicSharedMemorySize, few.smem_bytes));
        NB200_CUDA_TRY(cudaMalloc(&obj->d_partials, static_cast<size_t>(few.grid) * (1 + T + T * D) * sizeof(double)));
        NB200_CUDA_TRY(cudaDeviceSynchronize());
        obj->ft     = few;
        obj->use_ft = true;
        return NB200_OK;
    }

    // many-target tensor-pipe path
    mt_geometry geo;
    if (!force_generic && forced_variant != -3 && plan_mt(N, D, T, data->sm_count, data->smem_optin, geo))
    {
        const mt_variant& var = kMtVariants[geo.variant];
        obj->mt_variant    = geo.variant;
        obj->mt_fwd_grid   = geo.fwd_grid;
        obj->mt_tma        = data->x_map_ok && geo.tma_stages >= 2 && env_int("NB200_MT_LEGACY_FORWARD", 0) == 0;
        obj->mt_fwd_stages = obj->mt_tma ? geo.tma_stages : geo.fwd_stages;
        obj->mt_fwd_smem   = obj->mt_tma ? geo.tma_smem : geo.fwd_smem;
        obj->mt_splits     = geo.splits;
        obj->mt_bwd_grid   = geo.bwd_grid;
        obj->mt_bwd_stages = geo.bwd_stages;
        obj->mt_bwd_smem   = geo.bwd_smem;
        obj->mt_t_ld       = geo.t_ld;
        NB200_CUDA_TRY(cudaFuncSetAttribute(var.grad, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.fwd_smem));
        NB200_CUDA_TRY(cudaFuncSetAttribute(var.value, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.fwd_smem));
        if (obj->mt_tma)
        {
            NB200_CUDA_TRY(cudaFuncSetAttribute(var.grad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.tma_smem));
            NB200_CUDA_TRY(cudaFuncSetAttribute(var.value_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.tma_smem));
        }
        NB200_CUDA_TRY(cudaFuncSetAttribute(var.backward, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.bwd_smem));
        NB200_CUDA_TRY(cudaMalloc(&obj->d_partials, static_cast<size_t>(geo.fwd_grid) * (1 + T) * sizeof(double)));
        NB200_CUDA_TRY(cudaMalloc(&obj->d_partials_gw, static_cast<size_t>(geo.splits) * T * D * sizeof(double)));
        NB200_CUDA_TRY(cudaMalloc(&obj->d_dL, static_cast<size_t>(N) * (T + (T & 1)) * sizeof(double)));
        NB200_CUDA_TRY(cudaMalloc(&obj->d_Wp, static_cast<size_t>((D + kMtBK - 1) / kMtBK) * T * kMtLd * sizeof(double)));
        NB200_CUDA_TRY(cudaDeviceSynchronize());
        obj->use_mt = true;
        return NB200_OK;
    }

    // generic path
    const int64_t smem = (2 * kGenWarps * T + kGenWarps) * 8;
    if (smem > data->smem_optin)
    {
        return fail(NB200_ERR_UNSUPPORTED, "objective: too many targets for the generic forward kernel");
    }
    obj->gen_smem   = static_cast<int>(smem);
    obj->gen_blocks = static_cast<int>(
        std::max<int64_t>(1, std::min<int64_t>((N + kGenWarps - 1) / kGenWarps, 4LL * data->sm_count)));
    const int64_t col_tiles = (D + kBwdCols - 1) / kBwdCols;
    const int64_t t_tiles   = (T + kBwdTB - 1) / kBwdTB;
    int64_t       splits    = (4LL * data->sm_count + col_tiles * t_tiles - 1) / (col_tiles * t_tiles);
    splits                  = std::max<int64_t>(1, std::min<int64_t>(splits, std::min<int64_t>(256, (N + 255) / 256)));
    obj->gen_splits         = static_cast<int>(splits);
    if (smem > 48 * 1024)
    {
        NB200_CUDA_TRY(cudaFuncSetAttribute(gen_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            obj->gen_smem));
        NB200_CUDA_TRY(cudaFuncSetAttribute(gen_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            obj->gen_smem));
    }
    NB200_CUDA_TRY(cudaMalloc(&obj->d_partials, static_cast<size_t>(obj->gen_blocks) * (1 + T) * sizeof(double)));
    NB200_CUDA_TRY(cudaMalloc(&obj->d_partials_gw, static_cast<size_t>(splits) * T * D * sizeof(double)));
    NB200_CUDA_TRY(cudaMalloc(&obj->d_dL, static_cast<size_t>(N) * T * sizeof(double)));
    NB200_CUDA_TRY(cudaDeviceSynchronize()); // the memsets above ran on the legacy stream; launches use obj->stream
    return NB200_OK;
}

int make_objective(nb200_data* data, const int loss, const double loss_param, const int flavour, const double l1,
                   const double l2, const double* fixed_bias, const int variant, nb200_obj** out_obj)
{
    NB200_REQUIRE(data != nullptr && out_obj != nullptr, "objective: null argument");
    NB200_REQUIRE(loss >= 0 && loss < NB200_LOSS_COUNT, "objective: unknown loss");
    NB200_REQUIRE(flavour == NB200_FLAVOUR_LINEAR || flavour == NB200_FLAVOUR_SYNTH, "objective: unknown flavour");
    NB200_REQUIRE(flavour != NB200_FLAVOUR_SYNTH || (data->T == 1 && fixed_bias != nullptr),
                  "objective: the synthetic flavour needs one target and a fixed bias");
    NB200_REQUIRE(data->N > 0 && data->D > 0 && data->T > 0, "objective: empty dataset");

    NB200_CUDA_TRY(cudaSetDevice(data->device));
    auto* obj = new (std::nothrow) nb200_obj{};
    NB200_REQUIRE(obj != nullptr, "objective: out of host memory");
    nb200_dataset_retain(data);
    obj->data       = data;
    obj->device     = data->device;
    obj->loss       = loss;
    obj->loss_param = loss_param;
    obj->flavour    = flavour;
    obj->l1         = l1;
    obj->l2         = l2;
    obj->n          = (flavour == NB200_FLAVOUR_LINEAR) ? (data->D + 1) * data->T : data->D;

    const auto guard = [&](const int status)
    {
        if (status != NB200_OK)
        {
            const std::string message = g_last_error;
            free_obj(obj);
            g_last_error = message;
        }
        return status;
    };
    const auto alloc = [&]() -> int
    {
        const size_t n = static_cast<size_t>(obj->n);
        NB200_CUDA_TRY(cudaStreamCreateWithFlags(&obj->own_stream, cudaStreamNonBlocking));
        obj->stream = obj->own_stream;
        obj->ev0.assign(kTimingRing, nullptr);
        obj->ev1.assign(kTimingRing, nullptr);
        for (int i = 0; i < kTimingRing; ++i)
        {
            NB200_CUDA_TRY(cudaEventCreate(&obj->ev0[i]));
            NB200_CUDA_TRY(cudaEventCreate(&obj->ev1[i]));
        }
        NB200_CUDA_TRY(cudaMalloc(&obj->d_grid_counter, 2 * sizeof(unsigned long long))); // [grid barrier | published]
        NB200_CUDA_TRY(cudaMalloc(&obj->d_x, (n + 2) * sizeof(double)));
        NB200_CUDA_TRY(cudaMalloc(&obj->d_out, (n + 2) * sizeof(double)));
        NB200_CUDA_TRY(cudaMalloc(&obj->d_reduced, static_cast<size_t>(1 + data->T + data->T * data->D) * sizeof(double)));
        NB200_CUDA_TRY(cudaMemset(obj->d_reduced, 0, static_cast<size_t>(1 + data->T + data->T * data->D) * sizeof(double)));
        NB200_CUDA_TRY(cudaMallocHost(&obj->h_x, (n + 2) * sizeof(double)));
        NB200_CUDA_TRY(cudaHostAlloc(&obj->h_out, (n + 2) * sizeof(double), cudaHostAllocMapped));
        NB200_CUDA_TRY(cudaHostGetDevicePointer(&obj->h_out_dev, obj->h_out, 0));
        if (flavour == NB200_FLAVOUR_SYNTH)
        {
            obj->fixed_bias.assign(fixed_bias, fixed_bias + data->T);
            NB200_CUDA_TRY(cudaMalloc(&obj->d_fixed_bias, static_cast<size_t>(data->T) * sizeof(double)));
            NB200_CUDA_TRY(cudaMemcpy(obj->d_fixed_bias, fixed_bias, static_cast<size_t>(data->T) * sizeof(double),
                                      cudaMemcpyHostToDevice));
        }
        return setup_plan(obj, variant);
    };
    const int status = guard(alloc());
    if (status == NB200_OK)
    {
        *out_obj = obj;
    }
    return status;
}

void collect_timing(nb200_obj* obj)
{
    for (int i = 0; i < obj->pass_pending; ++i)
    {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, obj->ev0[i], obj->ev1[i]) == cudaSuccess)
        {
            obj->last_pass_ms = ms;
            obj->pass_ms_total += ms;
            obj->pass_count += 1;
            if (obj->pass_samples.size() < 65536)
            {
                obj->pass_samples.push_back(ms);
            }
        }
    }
    obj->pass_pending = 0;
}

int timing_begin(nb200_obj* obj)
{
    if (obj->timing)
    {
        if (obj->pass_pending == kTimingRing)
        {
            NB200_CUDA_TRY(cudaStreamSynchronize(obj->stream));
            collect_timing(obj);
        }
        NB200_CUDA_TRY(cudaEventRecord(obj->ev0[obj->pass_pending], obj->stream));
    }
    return NB200_OK;
}

int timing_end(nb200_obj* obj)
{
    if (obj->timing)
    {
        NB200_CUDA_TRY(cudaEventRecord(obj->ev1[obj->pass_pending], obj->stream));
        obj->pass_pending += 1;
    }
    return NB200_OK;
}

// peers: fold the `peer_world` reduced vectors the ranks published into this rank's mailbox (rank order) on the way
int enqueue_finish(nb200_obj* obj, const double* x_dev, const int64_t n_global, double* gx_dev, double* fx_dev,
                   const double* mail = nullptr)
{
    finish_params p{};
    p.reduced  = obj->d_reduced;
    if (mail != nullptr)
    {
        p.mail        = mail;
        p.world       = obj->peer_world;
        p.col_stride  = obj->peer_col_stride;
        p.reduced_out = obj->d_reduced;
    }
    p.x        = x_dev;
    p.gx       = gx_dev;
    p.fx       = fx_dev;
    p.D        = obj->data->D;
    p.T        = obj->data->T;
    p.n_global = n_global;
    p.flavour  = obj->flavour;
    p.l1       = obj->l1;
    p.l2       = obj->l2;
    const int64_t blocks = (gx_dev != nullptr) ? (obj->n + 255) / 256 : 1;
    finish_kernel<<<static_cast<unsigned>(blocks), 256, 0, obj->stream>>>(p);
    NB200_CUDA_TRY(cudaGetLastError());
    obj->launches += 1;
    return NB200_OK;
}

// Enqueue one evaluation on obj->stream.
//   finish == false: stop at the un-normalised [sum loss | sum dL | sum dL x^T] in obj->d_reduced (phase 1 of a
//                    sample-sharded evaluation);
//   finish == true:  also divide by n_global, add the regularisation terms and write gx_out (nullable) / fx_out.
// The single-target path does all of it in ONE cooperative launch; the generic path takes 3-6 small launches.
int enqueue_eval(nb200_obj* obj, const double* x_dev, const bool want_grad, const bool finish, const int64_t n_global,
                 double* gx_out, double* fx_out, const bool peers = false)
{
    const nb200_data* data = obj->data;
    const int64_t     N = data->N, D = data->D, T = data->T;
    const double*     W    = x_dev;
    const double*     bias = (obj->flavour == NB200_FLAVOUR_LINEAR) ? (x_dev + T * D) : obj->d_fixed_bias;

    int status = timing_begin(obj);
    if (status != NB200_OK)
    {
        return status;
    }

    if (obj->use_t1)
    {
        const t1_variant& var = kT1Variants[obj->t1.variant];
        // the counters only advance when the launch succeeded (a failed launch must not leave the next one waiting
        // for arrivals that never happen)
        const unsigned long long grid_epoch = obj->grid_epoch + 1;

        t1_params p{};
        p.X              = data->X;
        p.Y              = data->Y;
        p.w              = W;
        p.bias           = bias;
        p.partials       = obj->d_partials;
        p.N              = N;
        p.num_tiles      = obj->t1.num_tiles;
        p.D              = static_cast<int>(D);
        p.rows_per_stage = obj->t1.rows_per_stage;
        p.rows_per_group = obj->t1.rows_per_group;
        p.stages         = obj->t1.stages;
        p.stage_stride   = obj->t1.stage_stride;
        p.y_offset       = obj->t1.y_offset;
        p.loss           = obj->loss;
        p.loss_param     = obj->loss_param;
        p.l2_resident    = (static_cast<int64_t>(N) * (D + 1) * 8 <= (static_cast<int64_t>(data->l2_bytes) * 3) / 4) ? 1 : 0;
        p.grid_counter   = obj->d_grid_counter;
        p.grid_target    = grid_epoch * static_cast<unsigned long long>(obj->t1.grid);
        p.reduced        = obj->d_reduced;
        p.fuse_finish    = finish ? 1 : 0;
        p.flavour        = obj->flavour;
        p.gx_out         = finish ? gx_out : nullptr;
        p.fx_out         = fx_out;
        p.n_global       = n_global;
        p.l1             = obj->l1;
        p.l2             = obj->l2;
        p.world          = 1;
        p.peer_wait      = 1;
        const bool   publish_only = peers && obj->peer_wait == 0;
        const size_t mail_bytes   = static_cast<size_t>(2) * obj->peer_world * obj->peer_col_stride * sizeof(double);
        if (peers)
        {
            NB200_REQUIRE(finish, "peers: a fused evaluation always runs to (f, g)");
            p.world      = obj->peer_world;
            p.rank       = obj->peer_rank;
            p.col_stride = obj->peer_col_stride;
            p.epoch      = obj->peer_epoch + 1;
            p.peer_error = obj->d_peer_error;
            p.peer_wait  = obj->peer_wait;
            for (int r = 0; r < obj->peer_world; ++r)
            {
                p.peer_mail[r] = reinterpret_cast<double*>(obj->peer_base[r]);
                p.peer_flag[r] = reinterpret_cast<unsigned long long*>(obj->peer_base[r] + mail_bytes);
            }
            p.publish_counter = obj->d_grid_counter + 1;
            p.publish_target  = (obj->publish_epoch + 1) * static_cast<unsigned long long>(obj->t1.grid);
            if (publish_only)
            {
                p.fuse_finish = 0;
                p.gx_out      = nullptr;
            }
        }

        // cooperative launch: the grid barrier of the fused tail needs every CTA resident (grid <= #SMs, 1 CTA/SM)
        void*       args[] = {&p};
        const void* kernel = reinterpret_cast<const void*>(want_grad ? var.grad : var.value);
        NB200_CUDA_TRY(cudaLaunchCooperativeKernel(kernel, dim3(obj->t1.grid), dim3((var.NCW + kT1ProducerWarps) * 32),
                                                   args, static_cast<size_t>(obj->t1.smem_bytes), obj->stream));
        obj->grid_epoch = grid_epoch;
        obj->launches += 1;
        if (peers)
        {
            obj->peer_epoch += 1;
            obj->publish_epoch += 1;
        }
        status = timing_end(obj);
        if (status != NB200_OK || !publish_only || obj->peer_passive)
        {
            return status; // (a passive part of a device group only publishes: the first part folds and finishes)
        }

        // the stream (not an SM) waits until every peer's ready word for this evaluation is in my mailbox, then the
        // finish kernel folds the world's slots in rank order
        const unsigned long long parity = obj->peer_epoch & 1ULL;
        const auto* flags = reinterpret_cast<const unsigned long long*>(obj->peer_local + mail_bytes);
        const unsigned long long* ready = flags + 2ULL * obj->peer_world * obj->t1.grid + parity * obj->peer_world;
        for (int q = 0; q < obj->peer_world; ++q)
        {
            if (q == obj->peer_rank)
            {
                continue; // my own slices are ordered by the stream
            }
            const int rc = stream_wait_value()(obj->stream, reinterpret_cast<unsigned long long>(ready + q),
                                               obj->peer_epoch, 0u /*CU_STREAM_WAIT_VALUE_GEQ*/);
            if (rc != 0)
            {
                return fail(NB200_ERR_CUDA, "peers: cuStreamWaitValue64 failed with driver error " + std::to_string(rc));
            }
        }
        obj->peer_pending = true;
        const double* mail = reinterpret_cast<const double*>(obj->peer_local) + parity * obj->peer_world * obj->peer_col_stride;
        return enqueue_finish(obj, x_dev, n_global, want_grad ? gx_out : nullptr, fx_out, mail);
    }

    if (obj->use_fd)
    {
        const fd_variant& var = kFdVariants[obj->fd.variant];
        fd_params         f{};
        f.X            = data->X;
        f.Y            = data->Y;
        f.W            = W;
        f.b            = bias;
        f.partials     = obj->d_partials;
        f.N            = N;
        f.num_tiles    = obj->fd.num_tiles;
        f.D            = static_cast<int>(D);
        f.T            = static_cast<int>(T);
        f.slice        = obj->fd.slice;
        f.ld           = obj->fd.ld;
        f.stages       = obj->fd.stages;
        f.stage_stride = obj->fd.stage_stride;
        f.y_offset     = obj->fd.y_offset;
        f.w_offset     = obj->fd.w_offset;
        f.stage_offset = obj->fd.stage_offset;
        f.loss         = obj->loss;
        f.loss_param   = obj->loss_param;
        (want_grad ? var.grad : var.value)<<<obj->fd.grid, var.threads, obj->fd.smem_bytes, obj->stream>>>(f);
        NB200_CUDA_TRY(cudaGetLastError());
        obj->launches += 1;
        status = timing_end(obj);
        if (status != NB200_OK)
        {
            return status;
        }
        const int64_t stride = 1 + T + T * D;
        const int64_t cols   = want_grad ? stride : 1;
        reduce_partials_kernel<<<static_cast<unsigned>((cols + 127) / 128), 128, 0, obj->stream>>>(
            obj->d_partials, obj->fd.grid, stride, cols, obj->d_reduced);
        NB200_CUDA_TRY(cudaGetLastError());
        obj->launches += 1;
        return finish ? enqueue_finish(obj, x_dev, n_global, gx_out, fx_out) : NB200_OK;
    }

    if (obj->use_ft)
    {
        const ft_variant& var = kFtVariants[obj->ft.variant];
        ft_params         f{};
        f.X              = data->X;
        f.Y              = data->Y;
        f.W              = W;
        f.b              = bias;
        f.partials       = obj->d_partials;
        f.N              = N;
        f.num_tiles      = obj->ft.num_tiles;
        f.D              = static_cast<int>(D);
        f.rows_per_stage = obj->ft.rows_per_stage;
        f.stages         = obj->ft.stages;
        f.stage_stride   = obj->ft.stage_stride;
        f.y_offset       = obj->ft.y_offset;
        f.loss           = obj->loss;
        f.loss_param     = obj->loss_param;
        (want_grad ? var.grad : var.value)<<<obj->ft.grid, kFtThreads, obj->ft.smem_bytes, obj->stream>>>(f);
        NB200_CUDA_TRY(cudaGetLastError());
        obj->launches += 1;
        status = timing_end(obj);
        if (status != NB200_OK)
        {
            return status;
        }
        const int64_t stride = 1 + T + T * D;
        const int64_t cols   = want_grad ? stride : 1;
        reduce_partials_kernel<<<static_cast<unsigned>((cols + 127) / 128), 128, 0, obj->stream>>>(
            obj->d_partials, obj->ft.grid, stride, cols, obj->d_reduced);
        NB200_CUDA_TRY(cudaGetLastError());
        obj->launches += 1;
        return finish ? enqueue_finish(obj, x_dev, n_global, gx_out, fx_out) : NB200_OK;
    }

    if (obj->use_mt)
    {
        const mt_variant& var = kMtVariants[obj->mt_variant];
        mt_pack_weights_kernel<<<static_cast<unsigned>(std::min<int64_t>((T * D + 255) / 256, 2048)), 256, 0,
                                 obj->stream>>>(W, static_cast<int>(T), static_cast<int>(D), obj->d_Wp);
        NB200_CUDA_TRY(cudaGetLastError());
        obj->launches += 1;
        mt_forward_params f{};
        f.X           = data->X;
        f.Y           = data->Y;
        f.W           = obj->d_Wp;
        f.b           = bias;
        f.dL          = want_grad ? obj->d_dL : nullptr;
        f.partials_fb = obj->d_partials;
        f.N           = N;
        f.D           = static_cast<int>(D);
        f.T           = static_cast<int>(T);
        f.stages      = obj->mt_fwd_stages;
        f.loss        = obj->loss;
        f.loss_param  = obj->loss_param;
        f.y_broadcast     = 0;
        f.per_target_loss = 0;
        f.fb_stride       = static_cast<int>(1 + T);
        if (obj->mt_tma)
        {
            (want_grad ? var.grad_tma : var.value_tma)<<<obj->mt_fwd_grid, kMtThreads, obj->mt_fwd_smem, obj->stream>>>(
                f, data->x_map);
        }
        else
        {
            (want_grad ? var.grad : var.value)<<<obj->mt_fwd_grid, kMtThreads, obj->mt_fwd_smem, obj->stream>>>(f);
        }
        NB200_CUDA_TRY(cudaGetLastError());
        obj->launches += 1;
        if (want_grad)
        {
            mt_backward_params b{};
            b.X           = data->X;
            b.dL          = obj->d_dL;
            b.partials_gw = obj->d_partials_gw;
            b.N           = N;
            b.D           = static_cast<int>(D);
            b.T           = static_cast<int>(T);
            b.t_ld        = obj->mt_t_ld;
            b.splits      = obj->mt_splits;
            b.stages      = obj->mt_bwd_stages;
            var.backward<<<obj->mt_bwd_grid, kMtThreads, obj->mt_bwd_smem, obj->stream>>>(b);
            NB200_CUDA_TRY(cudaGetLastError());
            obj->launches += 1;
        }
        status = timing_end(obj);
        if (status != NB200_OK)
        {
            return status;
        }
        const int64_t cols_fb = want_grad ? (1 + T) : 1;
        reduce_partials_kernel<<<static_cast<unsigned>((cols_fb + 127) / 128), 128, 0, obj->stream>>>(
            obj->d_partials, obj->mt_fwd_grid, 1 + T, cols_fb, obj->d_reduced);
        NB200_CUDA_TRY(cudaGetLastError());
        obj->launches += 1;
        if (want_grad)
        {
            reduce_partials_kernel<<<static_cast<unsigned>((T * D + 127) / 128), 128, 0, obj->stream>>>(
                obj->d_partials_gw, obj->mt_splits, T * D, T * D, obj->d_reduced + 1 + T);
            NB200_CUDA_TRY(cudaGetLastError());
            obj->launches += 1;
        }
        return finish ? enqueue_finish(obj, x_dev, n_global, gx_out, fx_out) : NB200_OK;
    }

    gen_params p{};
    p.X           = data->X;
    p.Y           = data->Y;
    p.W           = W;
    p.b           = bias;
    p.dL          = want_grad ? obj->d_dL : nullptr;
    p.partials_fb = obj->d_partials;
    p.N           = N;
    p.D           = D;
    p.T           = T;
    p.loss        = obj->loss;
    p.loss_param  = obj->loss_param;
    if (want_grad)
    {
        gen_forward_kernel<true><<<obj->gen_blocks, kGenWarps * 32, obj->gen_smem, obj->stream>>>(p);
    }
    else
    {
        gen_forward_kernel<false><<<obj->gen_blocks, kGenWarps * 32, obj->gen_smem, obj->stream>>>(p);
    }
    NB200_CUDA_TRY(cudaGetLastError());
    obj->launches += 1;
    if (want_grad)
    {
        const dim3 grid(static_cast<unsigned>((D + kBwdCols - 1) / kBwdCols),
                        static_cast<unsigned>((T + kBwdTB - 1) / kBwdTB), static_cast<unsigned>(obj->gen_splits));
        gen_backward_kernel<<<grid, kBwdCols, 0, obj->stream>>>(data->X, obj->d_dL, N, D, T, obj->d_partials_gw);
        NB200_CUDA_TRY(cudaGetLastError());
        obj->launches += 1;
    }
    status = timing_end(obj);
    if (status != NB200_OK)
    {
        return status;
    }

    const int64_t cols_fb = want_grad ? (1 + T) : 1;
    reduce_partials_kernel<<<static_cast<unsigned>((cols_fb + 127) / 128), 128, 0, obj->stream>>>(
        obj->d_partials, obj->gen_blocks, 1 + T, cols_fb, obj->d_reduced);
    NB200_CUDA_TRY(cudaGetLastError());
    obj->launches += 1;
    if (want_grad)
    {
        reduce_partials_kernel<<<static_cast<unsigned>((T * D + 127) / 128), 128, 0, obj->stream>>>(
            obj->d_partials_gw, obj->gen_splits, T * D, T * D, obj->d_reduced + 1 + T);
        NB200_CUDA_TRY(cudaGetLastError());
        obj->launches += 1;
    }
    return finish ? enqueue_finish(obj, x_dev, n_global, gx_out, fx_out) : NB200_OK;
}

// Wait for the objective's stream. With a stream wait on the peers outstanding the host polls under a time limit
// (NB200_PEER_TIMEOUT_MS, default 30 s): a peer that never publishes would otherwise park the stream for ever. On
// expiry the wait is released by writing the awaited epoch into this rank's own ready words, and the evaluation is
// reported as failed — like the in-kernel wait's device-side time limit.
int obj_sync(nb200_obj* obj)
{
    if (obj->peer_pending && obj->peer_world > 1 && obj->peer_wait == 0)
    {
        static const int limit_ms = std::max(1, env_int("NB200_PEER_TIMEOUT_MS", 30000));
        const auto       start    = std::chrono::steady_clock::now();
        for (int spins = 0;; ++spins)
        {
            const cudaError_t state = cudaStreamQuery(obj->stream);
            if (state == cudaSuccess)
            {
                break;
            }
            if (state != cudaErrorNotReady)
            {
                NB200_CUDA_TRY(state);
            }
            if ((spins & 1023) == 1023 &&
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start).count() > limit_ms)
            {
                const size_t mail_bytes = static_cast<size_t>(2) * obj->peer_world * obj->peer_col_stride * sizeof(double);
                auto* flags = reinterpret_cast<unsigned long long*>(obj->peer_local + mail_bytes);
                unsigned long long* ready = flags + 2ULL * obj->peer_world * obj->t1.grid;
                for (int i = 0; i < 2 * obj->peer_world; ++i)
                {
                    obj->h_release[i] = obj->peer_epoch;
                }
                NB200_CUDA_TRY(cudaMemcpyAsync(ready, obj->h_release, 2 * obj->peer_world * sizeof(unsigned long long),
                                               cudaMemcpyHostToDevice, obj->release_stream));
                NB200_CUDA_TRY(cudaStreamSynchronize(obj->release_stream));
                NB200_CUDA_TRY(cudaStreamSynchronize(obj->stream));
                *obj->h_peer_error = 1;
                break;
            }
        }
        obj->peer_pending = false;
        return NB200_OK;
    }
    NB200_CUDA_TRY(cudaStreamSynchronize(obj->stream));
    obj->peer_pending = false;
    return NB200_OK;
}

// after a synchronisation: did an evaluation give up on a peer? (its sums are then stale: never hand them out)
int peer_check(nb200_obj* obj)
{
    if (obj->h_peer_error != nullptr && *obj->h_peer_error != 0)
    {
        *obj->h_peer_error = 0;
        return fail(NB200_ERR_CUDA, "peers: a rank did not reach the all-reduce within the time limit");
    }
    return NB200_OK;
}

int upload_x(nb200_obj* obj, const double* x, const int64_t n)
{
    NB200_REQUIRE(n == obj->n, "function: invalid input size, expecting (" + std::to_string(obj->n) + ",), got (" +
                                   std::to_string(n) + ",) instead!");
    std::memcpy(obj->h_x, x, static_cast<size_t>(n) * sizeof(double));
    NB200_CUDA_TRY(cudaMemcpyAsync(obj->d_x, obj->h_x, static_cast<size_t>(n) * sizeof(double),
                                   cudaMemcpyHostToDevice, obj->stream));
    return NB200_OK;
}

// where the kernels of a host-buffer call write (g, f): the mapped pinned buffer (a device buffer + one D2H copy was
// measured to be no faster, profiles/r01_e2e_breakdown.jsonl)
double* result_base(nb200_obj* obj)
{
    return obj->h_out_dev;
}

// the kernels wrote (g, f): wait for the stream, hand the values to the caller
int collect_result(nb200_obj* obj, double* gx, double* fx)
{
    const size_t n = static_cast<size_t>(obj->n);
    int          status = obj_sync(obj);
    if (status == NB200_OK)
    {
        status = peer_check(obj);
    }
    if (status != NB200_OK)
    {
        return status;
    }
    collect_timing(obj);
    if (gx != nullptr)
    {
        std::memcpy(gx, obj->h_out, n * sizeof(double));
    }
    *fx = obj->h_out[n];
    return NB200_OK;
}
} // namespace

#include "group.cuh"

namespace
{
int peer_allocate(nb200_obj* obj, int world);
int peer_choose_mode(nb200_obj* obj, int world);

// Wire the parts of a group objective: fused over peer memory when every part runs the single-target kernel on its own
// device with the same grid and every pair of devices has a peer path (NB200_GROUP_FUSED=0 keeps the two-phase path);
// otherwise the sums travel by peer copies. Called again after a change of plan.
int group_connect(nb200_obj* group)
{
    const int world = static_cast<int>(group->parts.size());
    group->group_fused = false;
    for (nb200_obj* part : group->parts)
    {
        drop_peers(part);
        part->peer_passive = false;
    }
    bool fused = world >= 2 && env_int("NB200_GROUP_FUSED", 1) != 0;
    for (int i = 0; i < world && fused; ++i)
    {
        const nb200_obj* a = group->parts[static_cast<size_t>(i)];
        fused = a->use_t1 && a->t1.grid == group->parts[0]->t1.grid;
        for (int j = 0; j < world && fused; ++j)
        {
            const nb200_obj* b = group->parts[static_cast<size_t>(j)];
            int can = 0;
            fused = i == j || (a->device != b->device &&
                               cudaDeviceCanAccessPeer(&can, a->device, b->device) == cudaSuccess && can != 0);
        }
    }
    cudaGetLastError();
    if (!fused)
    {
        return NB200_OK;
    }
    for (nb200_obj* part : group->parts)
    {
        NB200_CUDA_TRY(cudaSetDevice(part->device));
        const int status = peer_allocate(part, world);
        if (status != NB200_OK)
        {
            return status;
        }
    }
    NB200_CUDA_TRY(cudaSetDevice(group->parts[0]->device));
    int status = peer_choose_mode(group->parts[0], world);
    if (status != NB200_OK)
    {
        return status;
    }
    for (int r = 0; r < world; ++r)
    {
        nb200_obj* part = group->parts[static_cast<size_t>(r)];
        for (int q = 0; q < world; ++q)
        {
            part->peer_base[q] = group->parts[static_cast<size_t>(q)]->peer_local; // same process: plain pointers
        }
        part->peer_direct  = true;
        part->peer_world   = world;
        part->peer_rank    = r;
        part->peer_epoch   = 0;
        part->peer_wait    = group->parts[0]->peer_wait;
        part->peer_passive = r > 0;
    }
    group->group_fused = true;
    return NB200_OK;
}

int make_group_objective(nb200_data* data, const int loss, const double loss_param, const int flavour, const double l1,
                         const double l2, const double* fixed_bias, const int variant, nb200_obj** out_obj)
{
    NB200_REQUIRE(out_obj != nullptr, "objective: null argument");
    auto* group = new (std::nothrow) nb200_obj{};
    NB200_REQUIRE(group != nullptr, "objective: out of host memory");
    nb200_dataset_retain(data);
    group->data       = data;
    group->device     = data->device;
    group->loss       = loss;
    group->loss_param = loss_param;
    group->flavour    = flavour;
    group->l1         = l1;
    group->l2         = l2;
    const auto build = [&]() -> int
    {
        for (nb200_data* part_data : data->parts)
        {
            nb200_obj* part   = nullptr;
            const int  status = make_objective(part_data, loss, loss_param, flavour, l1, l2, fixed_bias, variant, &part);
            if (status != NB200_OK)
            {
                return status;
            }
            group->parts.push_back(part);
            cudaEvent_t event = nullptr;
            NB200_CUDA_TRY(cudaSetDevice(part->device));
            NB200_CUDA_TRY(cudaEventCreateWithFlags(&event, cudaEventDisableTiming));
            group->group_events.push_back(event);
        }
        group->n      = group->parts[0]->n;
        group->stream = group->parts[0]->stream;
        if (fixed_bias != nullptr)
        {
            group->fixed_bias.assign(fixed_bias, fixed_bias + data->T);
        }
        NB200_CUDA_TRY(cudaSetDevice(group->device));
        NB200_CUDA_TRY(cudaMalloc(&group->group_gather, data->parts.size() *
                                                            static_cast<size_t>(1 + data->T + data->T * data->D) *
                                                            sizeof(double)));
        return group_connect(group);
    };
    const int status = build();
    if (status != NB200_OK)
    {
        const std::string message = g_last_error;
        group->stream = nullptr;
        free_obj(group);
        g_last_error = message;
        return status;
    }
    *out_obj = group;
    return NB200_OK;
}

// entry points that only make sense for one device (the cross-process protocol, tuning knobs)
#define NB200_NOT_ON_A_GROUP(obj, what)                                                                                \
    do                                                                                                                 \
    {                                                                                                                  \
        if ((obj) != nullptr && !(obj)->parts.empty())                                                                 \
        {                                                                                                              \
            return ::nb200::fail(3 /*NB200_ERR_UNSUPPORTED*/,                                                           \
                                 std::string(what) + ": not available on a device group (it evaluates on all of its "  \
                                                     "devices by itself)");                                            \
        }                                                                                                              \
    } while (0)
} // namespace
#include "lbfgs.cuh"
#include "osga.cuh"

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

const char* nb200_last_error(void)
{
    return g_last_error.c_str();
}

const char* nb200_version(void)
{
    return "0.1.0";
}

int nb200_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return count;
}

int nb200_init(const int device, nb200_ctx** out_ctx)
{
    NB200_REQUIRE(out_ctx != nullptr, "init: null argument");
    int count = 0;
    NB200_CUDA_TRY(cudaGetDeviceCount(&count));
    if (count <= 0)
    {
        return fail(NB200_ERR_CUDA, "init: no CUDA device is visible; libnb200 has no CPU fallback");
    }
    NB200_REQUIRE(device >= 0 && device < count, "init: invalid device index");
    NB200_CUDA_TRY(cudaSetDevice(device));

    cudaDeviceProp prop{};
    NB200_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
    {
        return fail(NB200_ERR_CUDA, std::string("init: libnb200 is built for sm_100a (B200) only, found ") + prop.name);
    }

    auto* ctx = new (std::nothrow) nb200_ctx{};
    NB200_REQUIRE(ctx != nullptr, "init: out of host memory");
    ctx->device     = device;
    ctx->sm_count   = prop.multiProcessorCount;
    ctx->smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
    ctx->l2_bytes   = prop.l2CacheSize;
    const cudaError_t err = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (err != cudaSuccess)
    {
        delete ctx;
        return fail(NB200_ERR_CUDA, std::string("init: cannot create a stream: ") + cudaGetErrorString(err));
    }
    *out_ctx = ctx;
    return NB200_OK;
}

int nb200_init_multi(const int ndev, const int* devices, nb200_ctx** out_ctx)
{
    NB200_REQUIRE(out_ctx != nullptr && devices != nullptr, "init: null argument");
    NB200_REQUIRE(ndev >= 1 && ndev <= kT1MaxPeers, "init: a device group holds 1 to 8 devices");
    auto* group = new (std::nothrow) nb200_ctx{};
    NB200_REQUIRE(group != nullptr, "init: out of host memory");
    for (int i = 0; i < ndev; ++i)
    {
        nb200_ctx* part   = nullptr;
        const int  status = nb200_init(devices[i], &part);
        if (status != NB200_OK)
        {
            const std::string message = g_last_error;
            nb200_ctx_release(group);
            g_last_error = message;
            return status;
        }
        group->parts.push_back(part);
    }
    group->device     = group->parts[0]->device;
    group->sm_count   = group->parts[0]->sm_count;
    group->smem_optin = group->parts[0]->smem_optin;
    group->l2_bytes   = group->parts[0]->l2_bytes;
    // peer access both ways between distinct devices (the fused evaluation stores into the peers' mailboxes and the
    // two-phase one copies sums between devices); devices without a peer path still work through staged copies
    for (int i = 0; i < ndev; ++i)
    {
        for (int j = 0; j < ndev; ++j)
        {
            int can = 0;
            if (devices[i] != devices[j] && cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can != 0)
            {
                cudaSetDevice(devices[i]);
                cudaDeviceEnablePeerAccess(devices[j], 0); // cudaErrorPeerAccessAlreadyEnabled is fine
            }
            cudaGetLastError();
        }
    }
    *out_ctx = group;
    return NB200_OK;
}

int nb200_ctx_devices(const nb200_ctx* ctx)
{
    return ctx == nullptr ? 0 : (ctx->parts.empty() ? 1 : static_cast<int>(ctx->parts.size()));
}

void nb200_ctx_release(nb200_ctx* ctx)
{
    if (ctx != nullptr && !ctx->parts.empty())
    {
        for (nb200_ctx* part : ctx->parts)
        {
            nb200_ctx_release(part);
        }
        delete ctx;
        return;
    }
    if (ctx != nullptr)
    {
        cudaSetDevice(ctx->device);
        cudaStreamDestroy(ctx->stream);
        cudaFree(ctx->arena);
        cudaFreeHost(ctx->staging);
        delete ctx;
    }
}

// Tensor map of X[N, D] for the many-target forward: a box of 16 features x 128 rows lands in shared memory as dense
// 128-byte rows under the 128-byte swizzle; elements outside the matrix are filled with zeros. The encoder comes from
// the driver through the runtime (libnb200.so does not link libcuda).
static bool encode_x_map(const double* X, const int64_t N, const int64_t D, CUtensorMap* map)
{
    using encode_fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const encode_fn encode = []() -> encode_fn
    {
        void*                           entry = nullptr;
        cudaDriverEntryPointQueryResult found = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &entry, cudaEnableDefault, &found) != cudaSuccess ||
            found != cudaDriverEntryPointSuccess)
        {
            cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<encode_fn>(entry);
    }();
    if (encode == nullptr || D % 2 != 0 || N >= (1LL << 31) || D >= (1LL << 31))
    {
        return false; // (rows must start on 16-byte boundaries: the global stride is a multiple of 16 bytes)
    }
    const cuuint64_t dims[2]    = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(D) * sizeof(double)};
    const cuuint32_t box[2]     = {static_cast<cuuint32_t>(kMtBK), static_cast<cuuint32_t>(kMtBM)};
    const cuuint32_t estrides[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(X), dims, strides, box, estrides,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int alloc_dataset(nb200_ctx* ctx, const int64_t N, const int64_t D, const int64_t T, nb200_data** out_data)
{
    NB200_REQUIRE(ctx != nullptr && out_data != nullptr, "dataset: null argument");
    NB200_REQUIRE(N > 0 && D > 0 && T > 0, "dataset: sizes must be positive");
    NB200_REQUIRE(D < (1LL << 31) && T < (1LL << 31), "dataset: too many columns");
    NB200_CUDA_TRY(cudaSetDevice(ctx->device));

    auto* data = new (std::nothrow) nb200_data{};
    NB200_REQUIRE(data != nullptr, "dataset: out of host memory");
    data->device     = ctx->device;
    data->sm_count   = ctx->sm_count;
    data->smem_optin = ctx->smem_optin;
    data->l2_bytes   = ctx->l2_bytes;
    data->N          = N;
    data->D          = D;
    data->T          = T;
    const size_t      x_bytes = static_cast<size_t>(N) * D * sizeof(double);
    const size_t      y_bytes = static_cast<size_t>(N) * T * sizeof(double);
    constexpr size_t  slack   = 256;
    cudaError_t       err     = cudaMalloc(&data->X, x_bytes + slack);
    if (err == cudaSuccess)
    {
        err = cudaMalloc(&data->Y, y_bytes + slack);
    }
    if (err == cudaSuccess)
    {
        err = cudaMemset(reinterpret_cast<unsigned char*>(data->X) + x_bytes, 0, slack);
    }
    if (err == cudaSuccess)
    {
        err = cudaMemset(reinterpret_cast<unsigned char*>(data->Y) + y_bytes, 0, slack);
    }
    if (err != cudaSuccess)
    {
        cudaFree(data->X);
        cudaFree(data->Y);
        delete data;
        return fail(NB200_ERR_CUDA, std::string("dataset: cannot allocate HBM: ") + cudaGetErrorString(err));
    }
    data->x_map_ok = encode_x_map(data->X, N, D, &data->x_map);
    *out_data = data;
    return NB200_OK;
}

// a dataset on a device group: `make(part context, part index, first row, rows, out)` builds each part
extern "C++" template <typename tmake>
int group_dataset(nb200_ctx* ctx, const int64_t N, const int64_t D, const int64_t T, nb200_data** out_data,
                         const tmake& make)
{
    NB200_REQUIRE(out_data != nullptr, "dataset: null argument");
    const int world = static_cast<int>(ctx->parts.size());
    NB200_REQUIRE(N >= world && D > 0 && T > 0, "dataset: every device of the group needs at least one sample");
    auto* data = new (std::nothrow) nb200_data{};
    NB200_REQUIRE(data != nullptr, "dataset: out of host memory");
    data->device     = ctx->device;
    data->sm_count   = ctx->sm_count;
    data->smem_optin = ctx->smem_optin;
    data->l2_bytes   = ctx->l2_bytes;
    data->N          = N;
    data->D          = D;
    data->T          = T;
    group_split_rows(N, world, data->row0);
    for (int r = 0; r < world; ++r)
    {
        nb200_data* part   = nullptr;
        const int   status = make(ctx->parts[static_cast<size_t>(r)], r, data->row0[static_cast<size_t>(r)],
                                  data->row0[static_cast<size_t>(r) + 1] - data->row0[static_cast<size_t>(r)], &part);
        if (status != NB200_OK)
        {
            const std::string message = g_last_error;
            nb200_dataset_release(data);
            g_last_error = message;
            return status;
        }
        data->parts.push_back(part);
    }
    *out_data = data;
    return NB200_OK;
}

int nb200_dataset_pin(nb200_ctx* ctx, const double* X, const double* Y, const int64_t N, const int64_t D,
                      const int64_t T, nb200_data** out_data)
{
    NB200_REQUIRE(X != nullptr && Y != nullptr, "dataset: null inputs or targets");
    if (ctx != nullptr && !ctx->parts.empty())
    {
        return group_dataset(ctx, N, D, T, out_data,
                             [&](nb200_ctx* part, int, const int64_t row0, const int64_t rows, nb200_data** out)
                             { return nb200_dataset_pin(part, X + row0 * D, Y + row0 * T, rows, D, T, out); });
    }
    nb200_data* data   = nullptr;
    const int   status = alloc_dataset(ctx, N, D, T, &data);
    if (status != NB200_OK)
    {
        return status;
    }
    cudaError_t err = cudaMemcpy(data->X, X, static_cast<size_t>(N) * D * sizeof(double), cudaMemcpyHostToDevice);
    if (err == cudaSuccess)
    {
        err = cudaMemcpy(data->Y, Y, static_cast<size_t>(N) * T * sizeof(double), cudaMemcpyHostToDevice);
    }
    if (err != cudaSuccess)
    {
        nb200_dataset_release(data);
        return fail(NB200_ERR_CUDA, std::string("dataset: host to device copy failed: ") + cudaGetErrorString(err));
    }
    *out_data = data;
    return NB200_OK;
}

int nb200_dataset_synth(nb200_ctx* ctx, const uint64_t seed, const int synth_kind, const int64_t N, const int64_t D,
                        const int64_t T, const int64_t row_offset, const double noise, nb200_data** out_data)
{
    NB200_REQUIRE(synth_kind >= 0 && synth_kind <= 2, "dataset: unknown synthetic kind");
    NB200_REQUIRE(synth_kind == NB200_SYNTH_MULTICLASS ? T >= 2 : T == 1,
                  "dataset: classification/regression are single-target, multiclass needs T >= 2");
    NB200_REQUIRE(row_offset >= 0, "dataset: negative row offset");
    if (ctx != nullptr && !ctx->parts.empty())
    {
        return group_dataset(ctx, N, D, T, out_data,
                             [&](nb200_ctx* part, int, const int64_t row0, const int64_t rows, nb200_data** out) {
                                 return nb200_dataset_synth(part, seed, synth_kind, rows, D, T, row_offset + row0, noise, out);
                             });
    }
    nb200_data* data   = nullptr;
    int         status = alloc_dataset(ctx, N, D, T, &data);
    if (status != NB200_OK)
    {
        return status;
    }
    double*    W   = nullptr;
    double*    b   = nullptr;
    const auto run = [&]() -> int
    {
        NB200_CUDA_TRY(cudaMalloc(&W, static_cast<size_t>(T) * D * sizeof(double)));
        NB200_CUDA_TRY(cudaMalloc(&b, static_cast<size_t>(T) * sizeof(double)));
        const int blocks = ctx->sm_count * 8;
        synth_inputs_kernel<<<blocks, 256, 0, ctx->stream>>>(seed, N, D, row_offset, data->X);
        NB200_CUDA_TRY(cudaGetLastError());
        synth_weights_kernel<<<static_cast<unsigned>((T * D + 255) / 256), 256, 0, ctx->stream>>>(seed, D, T, W, b);
        NB200_CUDA_TRY(cudaGetLastError());
        predict_kernel<<<blocks, 256, 0, ctx->stream>>>(data->X, W, b, N, D, T, data->Y);
        NB200_CUDA_TRY(cudaGetLastError());
        synth_targets_kernel<<<static_cast<unsigned>((N + 255) / 256), 256, 0, ctx->stream>>>(seed, synth_kind, N, T,
                                                                                               row_offset, noise, data->Y);
        NB200_CUDA_TRY(cudaGetLastError());
        NB200_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return NB200_OK;
    };
    status = run();
    cudaFree(W);
    cudaFree(b);
    if (status != NB200_OK)
    {
        const std::string message = g_last_error;
        nb200_dataset_release(data);
        g_last_error = message;
        return status;
    }
    *out_data = data;
    return NB200_OK;
}

int nb200_dataset_read(const nb200_data* data, const int64_t row0, const int64_t rows, double* X, double* Y)
{
    NB200_REQUIRE(data != nullptr, "dataset: null argument");
    NB200_REQUIRE(row0 >= 0 && rows >= 0 && row0 + rows <= data->N, "dataset: row range out of bounds");
    if (!data->parts.empty())
    {
        for (size_t r = 0; r < data->parts.size(); ++r)
        {
            const int64_t lo = std::max(row0, data->row0[r]), hi = std::min(row0 + rows, data->row0[r + 1]);
            if (lo < hi)
            {
                const int status = nb200_dataset_read(data->parts[r], lo - data->row0[r], hi - lo,
                                                      X != nullptr ? X + (lo - row0) * data->D : nullptr,
                                                      Y != nullptr ? Y + (lo - row0) * data->T : nullptr);
                if (status != NB200_OK)
                {
                    return status;
                }
            }
        }
        return NB200_OK;
    }
    NB200_CUDA_TRY(cudaSetDevice(data->device));
    if (X != nullptr && rows > 0)
    {
        NB200_CUDA_TRY(cudaMemcpy(X, data->X + row0 * data->D, static_cast<size_t>(rows) * data->D * sizeof(double),
                                  cudaMemcpyDeviceToHost));
    }
    if (Y != nullptr && rows > 0)
    {
        NB200_CUDA_TRY(cudaMemcpy(Y, data->Y + row0 * data->T, static_cast<size_t>(rows) * data->T * sizeof(double),
                                  cudaMemcpyDeviceToHost));
    }
    return NB200_OK;
}

int64_t nb200_dataset_samples(const nb200_data* data)
{
    return data != nullptr ? data->N : 0;
}

int64_t nb200_dataset_inputs(const nb200_data* data)
{
    return data != nullptr ? data->D : 0;
}

int64_t nb200_dataset_targets(const nb200_data* data)
{
    return data != nullptr ? data->T : 0;
}

void nb200_dataset_retain(nb200_data* data)
{
    if (data != nullptr)
    {
        data->refs.fetch_add(1, std::memory_order_relaxed);
    }
}

void nb200_dataset_release(nb200_data* data)
{
    if (data != nullptr && data->refs.fetch_sub(1, std::memory_order_acq_rel) == 1)
    {
        for (nb200_data* part : data->parts)
        {
            nb200_dataset_release(part);
        }
        cudaSetDevice(data->device);
        cudaFree(data->X);
        cudaFree(data->Y);
        delete data;
    }
}

// raw [5][C]: count, sum, sum of squares, min, max over the finite values of every column
static int device_column_raw(nb200_data* data, const double* values, const int64_t C, std::vector<double>& raw)
{
    const int64_t N      = data->N;
    const int     chunks = static_cast<int>((N + kStatsRowsPerBlock - 1) / kStatsRowsPerBlock);
    double *      partial = nullptr, *folded = nullptr;
    const auto    run     = [&]() -> int
    {
        NB200_CUDA_TRY(cudaMalloc(&partial, static_cast<size_t>(chunks) * 5 * C * sizeof(double)));
        NB200_CUDA_TRY(cudaMalloc(&folded, static_cast<size_t>(5) * C * sizeof(double)));
        const dim3 grid(static_cast<unsigned>((C + 255) / 256), static_cast<unsigned>(chunks));
        column_stats_kernel<<<grid, 256>>>(values, N, C, partial);
        NB200_CUDA_TRY(cudaGetLastError());
        column_stats_fold_kernel<<<static_cast<unsigned>((C + 255) / 256), 256>>>(partial, chunks, C, folded);
        NB200_CUDA_TRY(cudaGetLastError());
        raw.assign(static_cast<size_t>(5 * C), 0.0);
        NB200_CUDA_TRY(cudaMemcpy(raw.data(), folded, raw.size() * sizeof(double), cudaMemcpyDeviceToHost));
        return NB200_OK;
    };
    const int status = run();
    cudaFree(partial);
    cudaFree(folded);
    return status;
}

static int device_column_stats(nb200_data* data, const double* values, const int64_t C, const unsigned char* enable,
                               column_stats_t& stats)
{
    std::vector<double> raw;
    const int           status = device_column_raw(data, values, C, raw);
    if (status == NB200_OK)
    {
        finish_column_stats(raw, C, enable, stats);
    }
    return status;
}

// statistics over ALL rows of a dataset on a device group: per-device sums folded in device order
static int group_column_stats(nb200_data* data, const bool inputs, const unsigned char* enable, column_stats_t& stats)
{
    const int64_t       C = inputs ? data->D : data->T;
    std::vector<double> total(static_cast<size_t>(5 * C), 0.0), raw;
    for (int64_t c = 0; c < C; ++c)
    {
        total[static_cast<size_t>(3 * C + c)] = INFINITY;
        total[static_cast<size_t>(4 * C + c)] = -INFINITY;
    }
    for (nb200_data* part : data->parts)
    {
        NB200_CUDA_TRY(cudaSetDevice(part->device));
        const int status = device_column_raw(part, inputs ? part->X : part->Y, C, raw);
        if (status != NB200_OK)
        {
            return status;
        }
        for (int64_t c = 0; c < C; ++c)
        {
            for (int k = 0; k < 3; ++k)
            {
                total[static_cast<size_t>(k * C + c)] += raw[static_cast<size_t>(k * C + c)];
            }
            total[static_cast<size_t>(3 * C + c)] = std::fmin(total[static_cast<size_t>(3 * C + c)], raw[static_cast<size_t>(3 * C + c)]);
            total[static_cast<size_t>(4 * C + c)] = std::fmax(total[static_cast<size_t>(4 * C + c)], raw[static_cast<size_t>(4 * C + c)]);
        }
    }
    finish_column_stats(total, C, enable, stats);
    return NB200_OK;
}

static int device_column_scale(double* values, const int64_t N, const int64_t C, const column_stats_t& stats,
                               const int scaling, const int sm_count)
{
    std::vector<double> shift(static_cast<size_t>(C), 0.0), mul(static_cast<size_t>(C), 1.0);
    for (int64_t i = 0; i < C; ++i)
    {
        switch (scaling) // scalar_stats_t::scale (stats.cpp:357-401)
        {
        case 1: shift[i] = stats.mean[i]; mul[i] = stats.div_range[i]; break;
        case 2: shift[i] = stats.min[i]; mul[i] = stats.div_range[i]; break;
        case 3: shift[i] = stats.mean[i]; mul[i] = stats.div_stdev[i]; break;
        default: break; // none: only non-finite -> 0
        }
    }
    double*    d   = nullptr;
    const auto run = [&]() -> int
    {
        NB200_CUDA_TRY(cudaMalloc(&d, static_cast<size_t>(2 * C) * sizeof(double)));
        NB200_CUDA_TRY(cudaMemcpy(d, shift.data(), static_cast<size_t>(C) * sizeof(double), cudaMemcpyHostToDevice));
        NB200_CUDA_TRY(cudaMemcpy(d + C, mul.data(), static_cast<size_t>(C) * sizeof(double), cudaMemcpyHostToDevice));
        column_scale_kernel<<<sm_count * 8, 256>>>(values, N, C, d, d + C);
        NB200_CUDA_TRY(cudaGetLastError());
        NB200_CUDA_TRY(cudaDeviceSynchronize());
        return NB200_OK;
    };
    const int status = run();
    cudaFree(d);
    return status;
}

int nb200_dataset_scale(nb200_data* data, const int inputs_scaling, const int targets_scaling,
                        const unsigned char* enable_inputs, const unsigned char* enable_targets)
{
    NB200_REQUIRE(data != nullptr, "dataset: null argument");
    NB200_REQUIRE(inputs_scaling >= 0 && inputs_scaling <= 3 && targets_scaling >= 0 && targets_scaling <= 3,
                  "dataset: unhandled scaling type");
    NB200_REQUIRE(data->inputs_stats.empty(), "dataset: already scaled");
    if (!data->parts.empty())
    {
        int status = group_column_stats(data, true, enable_inputs, data->inputs_stats);
        if (status == NB200_OK)
        {
            status = group_column_stats(data, false, enable_targets, data->targets_stats);
        }
        for (size_t r = 0; r < data->parts.size() && status == NB200_OK; ++r)
        {
            nb200_data* part = data->parts[r];
            NB200_CUDA_TRY(cudaSetDevice(part->device));
            status = device_column_scale(part->X, part->N, part->D, data->inputs_stats, inputs_scaling, part->sm_count);
            if (status == NB200_OK)
            {
                status = device_column_scale(part->Y, part->N, part->T, data->targets_stats, targets_scaling, part->sm_count);
            }
            part->inputs_stats          = data->inputs_stats;
            part->targets_stats         = data->targets_stats;
            part->inputs_stats.scaling  = inputs_scaling;
            part->targets_stats.scaling = targets_scaling;
        }
        if (status == NB200_OK)
        {
            data->inputs_stats.scaling  = inputs_scaling;
            data->targets_stats.scaling = targets_scaling;
        }
        return status;
    }
    NB200_CUDA_TRY(cudaSetDevice(data->device));
    int status = device_column_stats(data, data->X, data->D, enable_inputs, data->inputs_stats);
    if (status == NB200_OK)
    {
        status = device_column_stats(data, data->Y, data->T, enable_targets, data->targets_stats);
    }
    if (status == NB200_OK)
    {
        status = device_column_scale(data->X, data->N, data->D, data->inputs_stats, inputs_scaling, data->sm_count);
    }
    if (status == NB200_OK)
    {
        status = device_column_scale(data->Y, data->N, data->T, data->targets_stats, targets_scaling, data->sm_count);
    }
    if (status == NB200_OK)
    {
        data->inputs_stats.scaling  = inputs_scaling;
        data->targets_stats.scaling = targets_scaling;
    }
    return status;
}

int nb200_dataset_stats(const nb200_data* data, const int which, double* samples, double* min, double* max,
                        double* mean, double* stdev, double* div_range, double* div_stdev)
{
    NB200_REQUIRE(data != nullptr && (which == 0 || which == 1), "dataset: invalid argument");
    const column_stats_t& stats = (which == 0) ? data->inputs_stats : data->targets_stats;
    NB200_REQUIRE(!stats.empty(), "dataset: statistics are available after nb200_dataset_scale");
    const auto copy = [](const std::vector<double>& from, double* to)
    {
        if (to != nullptr)
        {
            std::memcpy(to, from.data(), from.size() * sizeof(double));
        }
    };
    copy(stats.samples, samples);
    copy(stats.min, min);
    copy(stats.max, max);
    copy(stats.mean, mean);
    copy(stats.stdev, stdev);
    copy(stats.div_range, div_range);
    copy(stats.div_stdev, div_stdev);
    return NB200_OK;
}

int nb200_upscale(const nb200_data* data, double* W, double* b)
{
    // nano::upscale (src/dataset/stats.cpp:211-230)
    NB200_REQUIRE(data != nullptr && W != nullptr && b != nullptr, "upscale: null argument");
    NB200_REQUIRE(!data->inputs_stats.empty(), "upscale: the dataset has not been scaled");
    std::vector<double> fw, fb, tw, tb;
    scaling_affine(data->inputs_stats, data->inputs_stats.scaling, fw, fb);
    scaling_affine(data->targets_stats, data->targets_stats.scaling, tw, tb);
    const int64_t D = data->D, T = data->T;
    for (int64_t t = 0; t < T; ++t)
    {
        double dot = 0.0;
        for (int64_t j = 0; j < D; ++j)
        {
            dot += W[t * D + j] * fb[static_cast<size_t>(j)];
        }
        b[t] = (dot + b[t] - tb[static_cast<size_t>(t)]) / tw[static_cast<size_t>(t)];
        for (int64_t j = 0; j < D; ++j)
        {
            W[t * D + j] = W[t * D + j] / tw[static_cast<size_t>(t)] * fw[static_cast<size_t>(j)];
        }
    }
    return NB200_OK;
}

int nb200_objective_create(nb200_data* data, const int loss, const double loss_param, const int flavour,
                           const double l1, const double l2, const double* fixed_bias, nb200_obj** out_obj)
{
    if (data != nullptr && !data->parts.empty())
    {
        return make_group_objective(data, loss, loss_param, flavour, l1, l2, fixed_bias, -1, out_obj);
    }
    return make_objective(data, loss, loss_param, flavour, l1, l2, fixed_bias, -1, out_obj);
}

int nb200_objective_clone(const nb200_obj* obj, nb200_obj** out_obj)
{
    NB200_REQUIRE(obj != nullptr && out_obj != nullptr, "objective: null argument");
    if (!obj->parts.empty())
    {
        const nb200_obj* first = obj->parts[0];
        const int status = make_group_objective(
            obj->data, obj->loss, obj->loss_param, obj->flavour, obj->l1, obj->l2,
            obj->fixed_bias.empty() ? nullptr : obj->fixed_bias.data(),
            first->use_t1 ? first->t1.variant : (first->use_fd ? -5 : (first->use_ft ? -6 : (first->use_mt ? -4 : -2))), out_obj);
        for (size_t r = 0; status == NB200_OK && r < obj->parts.size(); ++r)
        {
            (*out_obj)->parts[r]->timing = obj->parts[r]->timing;
        }
        return status;
    }
    const int status = make_objective(obj->data, obj->loss, obj->loss_param, obj->flavour, obj->l1, obj->l2,
                                      obj->fixed_bias.empty() ? nullptr : obj->fixed_bias.data(),
                                      obj->use_t1 ? obj->t1.variant : (obj->use_fd ? -5 : (obj->use_ft ? -6 : (obj->use_mt ? -4 : -2))), out_obj);
    if (status == NB200_OK)
    {
        (*out_obj)->timing = obj->timing;
    }
    return status;
}

void nb200_objective_release(nb200_obj* obj)
{
    free_obj(obj);
}

int64_t nb200_objective_size(const nb200_obj* obj)
{
    return obj != nullptr ? obj->n : 0;
}

int nb200_vgrad(nb200_obj* obj, const double* x, const int64_t n, double* gx, double* fx)
{
    NB200_REQUIRE(obj != nullptr && x != nullptr && fx != nullptr, "function: null argument");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    // (a device group stages x and collects (g, f) through its first part; every part evaluates)
    nb200_obj* const host = obj->parts.empty() ? obj : obj->parts[0];
    const auto t0     = std::chrono::steady_clock::now();
    int        status = upload_x(host, x, n);
    const auto t1     = std::chrono::steady_clock::now();
    if (status == NB200_OK)
    {
        double* out = result_base(host);
        status      = obj->parts.empty()
                          ? enqueue_eval(obj, obj->d_x, gx != nullptr, true, obj->data->N, gx != nullptr ? out : nullptr,
                                         out + obj->n)
                          : group_enqueue_eval(obj, host->d_x, gx != nullptr, obj->data->N, gx != nullptr ? out : nullptr,
                                               out + obj->n);
    }
    const auto t2 = std::chrono::steady_clock::now();
    if (status == NB200_OK)
    {
        status = collect_result(host, gx, fx);
    }
    const auto t3 = std::chrono::steady_clock::now();
    obj->host_us[0] += std::chrono::duration<double, std::micro>(t1 - t0).count();
    obj->host_us[1] += std::chrono::duration<double, std::micro>(t2 - t1).count();
    obj->host_us[2] += std::chrono::duration<double, std::micro>(t3 - t2).count();
    obj->host_calls += 1;
    return status;
}

int nb200_objective_host_profile(nb200_obj* obj, double* stage_us, double* enqueue_us, double* wait_us,
                                 int64_t* calls, const int reset)
{
    NB200_REQUIRE(obj != nullptr && stage_us != nullptr && enqueue_us != nullptr && wait_us != nullptr &&
                      calls != nullptr,
                  "objective: null argument");
    *stage_us   = obj->host_us[0];
    *enqueue_us = obj->host_us[1];
    *wait_us    = obj->host_us[2];
    *calls      = obj->host_calls;
    if (reset != 0)
    {
        obj->host_us[0] = obj->host_us[1] = obj->host_us[2] = 0.0;
        obj->host_calls = 0;
    }
    return NB200_OK;
}

int nb200_vgrad_device(nb200_obj* obj, const double* x_dev, const int64_t n, double* gx_dev, double* fx_dev)
{
    NB200_REQUIRE(obj != nullptr && x_dev != nullptr && fx_dev != nullptr, "function: null argument");
    NB200_REQUIRE(n == obj->n, "function: invalid input size");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    if (!obj->parts.empty())
    {
        return group_enqueue_eval(obj, x_dev, gx_dev != nullptr, obj->data->N, gx_dev, fx_dev); // x on the first device
    }
    return enqueue_eval(obj, x_dev, gx_dev != nullptr, true, obj->data->N, gx_dev, fx_dev);
}

int nb200_objective_sync(nb200_obj* obj)
{
    NB200_REQUIRE(obj != nullptr, "objective: null argument");
    if (!obj->parts.empty())
    {
        for (size_t r = obj->parts.size(); r-- > 0;)
        {
            const int status = nb200_objective_sync(obj->parts[r]);
            if (status != NB200_OK)
            {
                return status;
            }
        }
        return NB200_OK;
    }
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    const int status = obj_sync(obj);
    if (status != NB200_OK)
    {
        return status;
    }
    collect_timing(obj);
    return peer_check(obj); // nb200_vgrad_sharded_device is checked here
}

int nb200_vgrad_partial(nb200_obj* obj, const double* x, const int64_t n, const int want_grad,
                        double** out_partial_dev, int64_t* out_partial_len)
{
    NB200_NOT_ON_A_GROUP(obj, "nb200_vgrad_partial");
    NB200_REQUIRE(obj != nullptr && x != nullptr && out_partial_dev != nullptr && out_partial_len != nullptr,
                  "function: null argument");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    int status = upload_x(obj, x, n);
    if (status == NB200_OK)
    {
        status = enqueue_eval(obj, obj->d_x, want_grad != 0, false, 0, nullptr, nullptr);
    }
    if (status == NB200_OK)
    {
        obj->partial_has_grad = want_grad != 0;
        *out_partial_dev      = obj->d_reduced;
        *out_partial_len      = 1 + obj->data->T + obj->data->T * obj->data->D;
    }
    return status;
}

int nb200_vgrad_finish(nb200_obj* obj, const int64_t n_global, double* gx, double* fx)
{
    NB200_NOT_ON_A_GROUP(obj, "nb200_vgrad_finish");
    NB200_REQUIRE(obj != nullptr && fx != nullptr, "function: null argument");
    NB200_REQUIRE(n_global > 0, "function: the global sample count must be positive");
    NB200_REQUIRE(gx == nullptr || obj->partial_has_grad, "function: the gradient was not requested in phase 1");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    double* out    = result_base(obj);
    int     status = enqueue_finish(obj, obj->d_x, n_global, gx != nullptr ? out : nullptr, out + obj->n);
    if (status == NB200_OK)
    {
        status = collect_result(obj, gx, fx);
    }
    return status;
}

int nb200_vgrad_partial_device(nb200_obj* obj, const double* x_dev, const int64_t n, const int want_grad,
                               double** out_partial_dev, int64_t* out_partial_len)
{
    NB200_NOT_ON_A_GROUP(obj, "nb200_vgrad_partial_device");
    NB200_REQUIRE(obj != nullptr && x_dev != nullptr && out_partial_dev != nullptr && out_partial_len != nullptr,
                  "function: null argument");
    NB200_REQUIRE(n == obj->n, "function: invalid input size");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    const int status = enqueue_eval(obj, x_dev, want_grad != 0, false, 0, nullptr, nullptr);
    if (status == NB200_OK)
    {
        obj->partial_has_grad = want_grad != 0;
        *out_partial_dev      = obj->d_reduced;
        *out_partial_len      = 1 + obj->data->T + obj->data->T * obj->data->D;
    }
    return status;
}

int nb200_vgrad_finish_device(nb200_obj* obj, const double* x_dev, const int64_t n_global, double* gx_dev,
                              double* fx_dev)
{
    NB200_NOT_ON_A_GROUP(obj, "nb200_vgrad_finish_device");
    NB200_REQUIRE(obj != nullptr && x_dev != nullptr && fx_dev != nullptr, "function: null argument");
    NB200_REQUIRE(n_global > 0, "function: the global sample count must be positive");
    NB200_REQUIRE(gx_dev == nullptr || obj->partial_has_grad, "function: the gradient was not requested in phase 1");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    return enqueue_finish(obj, x_dev, n_global, gx_dev, fx_dev);
}

// this rank's mailbox: mail [2 (parity)][world][col_stride] doubles, per-CTA flags [2][world][grid] and the ready words
// [2][world] of the publish-only mode (64-bit each), zeroed
} // extern "C"
namespace
{
int peer_allocate(nb200_obj* obj, const int world)
{
    NB200_REQUIRE(world >= 2 && world <= kT1MaxPeers, "peers: the world size must be in [2, 8]");
    NB200_REQUIRE(obj->use_t1, "peers: the fused all-reduce exists for the single-target kernel only");
    NB200_REQUIRE(obj->peer_local == nullptr, "peers: already exported");
    obj->peer_col_stride = static_cast<int>(round_up(obj->data->D + 2, 2));
    obj->peer_grid       = obj->t1.grid;
    const size_t bytes   = (static_cast<size_t>(2) * world * obj->peer_col_stride +
                          static_cast<size_t>(2) * world * obj->t1.grid + static_cast<size_t>(2) * world) * sizeof(double);
    NB200_CUDA_TRY(cudaMalloc(&obj->peer_local, bytes));
    NB200_CUDA_TRY(cudaMemset(obj->peer_local, 0, bytes));
    if (obj->h_peer_error == nullptr)
    {
        NB200_CUDA_TRY(cudaHostAlloc(&obj->h_peer_error, sizeof(int), cudaHostAllocMapped));
        NB200_CUDA_TRY(cudaHostGetDevicePointer(&obj->d_peer_error, obj->h_peer_error, 0));
        NB200_CUDA_TRY(cudaMallocHost(&obj->h_release, 2 * kT1MaxPeers * sizeof(unsigned long long)));
        NB200_CUDA_TRY(cudaStreamCreateWithFlags(&obj->release_stream, cudaStreamNonBlocking));
    }
    *obj->h_peer_error = 0;
    // the publish counter restarts with the mailbox
    NB200_CUDA_TRY(cudaMemset(obj->d_grid_counter + 1, 0, sizeof(unsigned long long)));
    obj->publish_epoch = 0;
    NB200_CUDA_TRY(cudaDeviceSynchronize());
    return NB200_OK;
}

// Publish-only evaluations need the stream to wait on a word in HBM (cuStreamWaitValue64). Probe it once on this
// objective's own mailbox: a wait that is already satisfied must fall through. NB200_PEER_WAIT=1 forces the
// in-kernel wait, =0 insists on the stream wait.
int peer_choose_mode(nb200_obj* obj, const int world)
{
    const int forced = env_int("NB200_PEER_WAIT", -1);
    obj->peer_wait   = 1;
    if (forced == 1)
    {
        return NB200_OK;
    }
    bool usable = stream_wait_value() != nullptr;
    if (usable)
    {
        const size_t mail_bytes = static_cast<size_t>(2) * world * obj->peer_col_stride * sizeof(double);
        auto* flags = reinterpret_cast<unsigned long long*>(obj->peer_local + mail_bytes);
        // (the word probed is a ready word of parity 0 that no peer has written yet: it holds 0)
        usable = stream_wait_value()(obj->stream, reinterpret_cast<unsigned long long>(flags + 2ULL * world * obj->t1.grid),
                                     0ULL, 0u) == 0 &&
                 cudaStreamSynchronize(obj->stream) == cudaSuccess;
        cudaGetLastError();
    }
    NB200_REQUIRE(usable || forced != 0, "peers: NB200_PEER_WAIT=0 but cuStreamWaitValue64 is not usable on this device");
    obj->peer_wait = usable ? 0 : 1;
    return NB200_OK;
}
} // namespace
extern "C" {

int nb200_objective_peer_export(nb200_obj* obj, const int world, void* handle_out, int64_t* grid_out)
{
    NB200_NOT_ON_A_GROUP(obj, "nb200_objective_peer_export");
    NB200_REQUIRE(obj != nullptr && handle_out != nullptr && grid_out != nullptr, "peers: null argument");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are exchanged as 64 bytes");
    const int status = peer_allocate(obj, world);
    if (status != NB200_OK)
    {
        return status;
    }
    cudaIpcMemHandle_t handle;
    NB200_CUDA_TRY(cudaIpcGetMemHandle(&handle, obj->peer_local));
    std::memcpy(handle_out, &handle, sizeof(handle));
    *grid_out = obj->t1.grid;
    return NB200_OK;
}

int nb200_objective_peer_connect(nb200_obj* obj, const int rank, const int world, const void* handles)
{
    NB200_NOT_ON_A_GROUP(obj, "nb200_objective_peer_connect");
    NB200_REQUIRE(obj != nullptr && handles != nullptr, "peers: null argument");
    NB200_REQUIRE(obj->peer_local != nullptr, "peers: export first");
    NB200_REQUIRE(obj->peer_world == 1, "peers: already connected (a new plan drops the connection)");
    NB200_REQUIRE(obj->use_t1 && obj->peer_grid == obj->t1.grid, "peers: the launch plan changed since the export");
    NB200_REQUIRE(world >= 2 && world <= kT1MaxPeers && rank >= 0 && rank < world, "peers: invalid rank / world");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    for (int r = 0; r < world; ++r)
    {
        if (r == rank)
        {
            obj->peer_base[r] = obj->peer_local;
            continue;
        }
        cudaIpcMemHandle_t handle;
        std::memcpy(&handle, static_cast<const unsigned char*>(handles) + static_cast<size_t>(r) * sizeof(handle),
                    sizeof(handle));
        void* mapped = nullptr;
        NB200_CUDA_TRY(cudaIpcOpenMemHandle(&mapped, handle, cudaIpcMemLazyEnablePeerAccess));
        obj->peer_base[r] = static_cast<unsigned char*>(mapped);
    }
    const int status = peer_choose_mode(obj, world);
    if (status != NB200_OK)
    {
        return status;
    }
    obj->peer_world = world;
    obj->peer_rank  = rank;
    obj->peer_epoch = 0;
    return NB200_OK;
}

int nb200_objective_peer_mode(const nb200_obj* obj)
{
    return (obj == nullptr || obj->peer_world <= 1) ? -1 : obj->peer_wait;
}

static int check_peers(nb200_obj* obj)
{
    NB200_REQUIRE(obj->peer_world > 1 && obj->use_t1, "peers: nb200_objective_peer_connect has not been called");
    return NB200_OK;
}

int nb200_vgrad_sharded(nb200_obj* obj, const double* x, const int64_t n, const int64_t n_global, double* gx, double* fx)
{
    NB200_NOT_ON_A_GROUP(obj, "nb200_vgrad_sharded");
    NB200_REQUIRE(obj != nullptr && x != nullptr && fx != nullptr, "function: null argument");
    NB200_REQUIRE(n_global > 0, "function: the global sample count must be positive");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    int status = check_peers(obj);
    if (status == NB200_OK)
    {
        status = upload_x(obj, x, n);
    }
    if (status == NB200_OK)
    {
        double* out = result_base(obj);
        status      = enqueue_eval(obj, obj->d_x, gx != nullptr, true, n_global, gx != nullptr ? out : nullptr,
                                   out + obj->n, true);
    }
    if (status == NB200_OK)
    {
        status = collect_result(obj, gx, fx); // also reports a peer that never showed up
    }
    return status;
}

int nb200_vgrad_sharded_device(nb200_obj* obj, const double* x_dev, const int64_t n, const int64_t n_global,
                               double* gx_dev, double* fx_dev)
{
    NB200_NOT_ON_A_GROUP(obj, "nb200_vgrad_sharded_device");
    NB200_REQUIRE(obj != nullptr && x_dev != nullptr && fx_dev != nullptr, "function: null argument");
    NB200_REQUIRE(n == obj->n && n_global > 0, "function: invalid sizes");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    const int status = check_peers(obj);
    return status != NB200_OK ? status : enqueue_eval(obj, x_dev, gx_dev != nullptr, true, n_global, gx_dev, fx_dev, true);
}

void* nb200_objective_stream(const nb200_obj* obj)
{
    return obj != nullptr ? static_cast<void*>(obj->stream) : nullptr;
}

int nb200_objective_set_stream(nb200_obj* obj, void* stream, const int use_own)
{
    NB200_NOT_ON_A_GROUP(obj, "nb200_objective_set_stream");
    NB200_REQUIRE(obj != nullptr, "objective: null argument");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    NB200_CUDA_TRY(cudaStreamSynchronize(obj->stream));
    collect_timing(obj);
    // a null handle with use_own == 0 is the legacy default stream (what torch.cuda.current_stream() is by default)
    obj->stream = (use_own != 0) ? obj->own_stream : static_cast<cudaStream_t>(stream);
    return NB200_OK;
}

int64_t nb200_objective_launches(const nb200_obj* obj)
{
    int64_t total = obj != nullptr ? obj->launches : 0;
    for (size_t r = 0; obj != nullptr && r < obj->parts.size(); ++r)
    {
        total += obj->parts[r]->launches;
    }
    return total;
}

int nb200_objective_parts(const nb200_obj* obj)
{
    return obj == nullptr ? 0 : (obj->parts.empty() ? 1 : static_cast<int>(obj->parts.size()));
}

nb200_obj* nb200_objective_part(nb200_obj* obj, const int index)
{
    if (obj == nullptr || index < 0 || index >= nb200_objective_parts(obj))
    {
        return nullptr;
    }
    return obj->parts.empty() ? obj : obj->parts[static_cast<size_t>(index)];
}

int nb200_objective_enable_timing(nb200_obj* obj, const int enabled)
{
    NB200_REQUIRE(obj != nullptr, "objective: null argument");
    obj->timing = enabled != 0;
    for (nb200_obj* part : obj->parts)
    {
        part->timing = enabled != 0;
    }
    return NB200_OK;
}

double nb200_objective_last_pass_ms(nb200_obj* obj)
{
    if (obj == nullptr)
    {
        return -1.0;
    }
    if (!obj->parts.empty())
    {
        double slowest = -1.0; // the evaluation is as long as its slowest device
        for (nb200_obj* part : obj->parts)
        {
            slowest = std::max(slowest, nb200_objective_last_pass_ms(part));
        }
        return slowest;
    }
    if (obj->pass_pending > 0)
    {
        cudaSetDevice(obj->device);
        if (cudaStreamSynchronize(obj->stream) != cudaSuccess)
        {
            return -1.0;
        }
        collect_timing(obj);
    }
    return obj->last_pass_ms;
}

int nb200_objective_pass_stats(nb200_obj* obj, double* total_ms, int64_t* count, const int reset)
{
    NB200_REQUIRE(obj != nullptr, "objective: null argument");
    if (!obj->parts.empty())
    {
        // the device with the largest total: the critical path of the group's evaluations (per-device figures through
        // nb200_objective_part)
        double  worst = -1.0, part_ms = 0.0;
        int64_t worst_count = 0, part_count = 0;
        for (nb200_obj* part : obj->parts)
        {
            const int status = nb200_objective_pass_stats(part, &part_ms, &part_count, reset);
            if (status != NB200_OK)
            {
                return status;
            }
            if (part_ms > worst)
            {
                worst       = part_ms;
                worst_count = part_count;
            }
        }
        if (total_ms != nullptr)
        {
            *total_ms = worst;
        }
        if (count != nullptr)
        {
            *count = worst_count;
        }
        return NB200_OK;
    }
    if (total_ms != nullptr)
    {
        *total_ms = obj->pass_ms_total;
    }
    if (count != nullptr)
    {
        *count = obj->pass_count;
    }
    if (reset != 0)
    {
        obj->pass_ms_total = 0.0;
        obj->pass_count    = 0;
        obj->pass_samples.clear();
    }
    return NB200_OK;
}

int nb200_objective_pass_samples(nb200_obj* obj, double* pass_ms, const int64_t capacity, int64_t* count)
{
    NB200_REQUIRE(obj != nullptr && count != nullptr && (pass_ms != nullptr || capacity == 0), "objective: null argument");
    const nb200_obj* source = obj->parts.empty() ? obj : obj->parts[0];
    *count = static_cast<int64_t>(source->pass_samples.size());
    for (int64_t i = 0; i < std::min<int64_t>(capacity, *count); ++i)
    {
        pass_ms[i] = source->pass_samples[static_cast<size_t>(i)];
    }
    return NB200_OK;
}

// ---- gradient-boosting cousins (gboost.cuh) ------------------------------------------------------------------------
} // extern "C"

struct nb200_gboost
{
    int          device{0};
    cudaStream_t stream{nullptr};
    int          kind{0}, loss{0};
    double       loss_param{0.0};
    int64_t      N{0}, T{0}, K{0}, n{0}, n_out{0};
    int          blocks{0}, smem{0};
    double *     d_targets{nullptr}, *d_soutputs{nullptr}, *d_woutputs{nullptr};
    int*         d_cluster{nullptr};
    double *     d_x{nullptr}, *d_out{nullptr}, *d_partials{nullptr}, *d_reduced{nullptr};
    double *     h_x{nullptr}, *h_out{nullptr};
    int64_t      launches{0};
};

extern "C" {

void nb200_gboost_release(nb200_gboost* obj)
{
    if (obj == nullptr)
    {
        return;
    }
    cudaSetDevice(obj->device);
    cudaFree(obj->d_targets);
    cudaFree(obj->d_soutputs);
    cudaFree(obj->d_woutputs);
    cudaFree(obj->d_cluster);
    cudaFree(obj->d_x);
    cudaFree(obj->d_out);
    cudaFree(obj->d_partials);
    cudaFree(obj->d_reduced);
    cudaFreeHost(obj->h_x);
    cudaFreeHost(obj->h_out);
    if (obj->stream != nullptr)
    {
        cudaStreamDestroy(obj->stream);
    }
    delete obj;
}

int nb200_gboost_create(nb200_ctx* ctx, const int kind, const int loss, const double loss_param, const double* targets,
                        const int64_t N, const int64_t T, const int32_t* cluster, const int64_t groups,
                        const double* soutputs, const double* woutputs, nb200_gboost** out_obj)
{
    NB200_REQUIRE(ctx != nullptr && targets != nullptr && out_obj != nullptr, "gboost: null argument");
    NB200_REQUIRE(kind >= GBOOST_BIAS && kind <= GBOOST_GRADS, "gboost: unknown function kind");
    NB200_REQUIRE(loss >= 0 && loss < NB200_LOSS_SYNTH_CAUCHY, "gboost: unknown loss");
    NB200_REQUIRE(N > 0 && T > 0 && T < (1 << 20), "gboost: invalid dimensions");
    NB200_REQUIRE(kind != GBOOST_SCALE ||
                      (cluster != nullptr && soutputs != nullptr && woutputs != nullptr && groups > 0 && groups <= 2048),
                  "gboost: the scale function needs cluster ids, strong and weak outputs and 1 to 2048 groups");
    if (kind == GBOOST_SCALE)
    {
        for (int64_t i = 0; i < N; ++i)
        {
            NB200_REQUIRE(cluster[i] < groups, "gboost: cluster id out of range");
        }
    }
    if (!ctx->parts.empty())
    {
        ctx = ctx->parts[0];
    }
    NB200_CUDA_TRY(cudaSetDevice(ctx->device));
    auto* obj = new (std::nothrow) nb200_gboost{};
    NB200_REQUIRE(obj != nullptr, "gboost: out of host memory");
    obj->device     = ctx->device;
    obj->kind       = kind;
    obj->loss       = loss;
    obj->loss_param = loss_param;
    obj->N          = N;
    obj->T          = T;
    obj->K          = (kind == GBOOST_SCALE) ? groups : 0;
    obj->n          = (kind == GBOOST_BIAS) ? T : (kind == GBOOST_SCALE ? groups : N * T);
    obj->n_out      = (kind == GBOOST_BIAS) ? T : (kind == GBOOST_SCALE ? groups : 0);
    obj->smem       = static_cast<int>((kGbWarps * obj->n_out + kGbWarps) * sizeof(double));
    const auto build = [&]() -> int
    {
        NB200_REQUIRE(obj->smem <= ctx->smem_optin, "gboost: too many targets for the per-warp accumulators");
        const size_t  rows  = static_cast<size_t>(N) * T * sizeof(double);
        const int64_t iters = (N + 31) / 32; // at least this many warp iterations
        obj->blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((iters + kGbWarps - 1) / kGbWarps,
                                                                               4LL * ctx->sm_count)));
        NB200_CUDA_TRY(cudaStreamCreateWithFlags(&obj->stream, cudaStreamNonBlocking));
        NB200_CUDA_TRY(cudaMalloc(&obj->d_targets, rows));
        NB200_CUDA_TRY(cudaMemcpy(obj->d_targets, targets, rows, cudaMemcpyHostToDevice));
        if (kind == GBOOST_SCALE)
        {
            NB200_CUDA_TRY(cudaMalloc(&obj->d_soutputs, rows));
            NB200_CUDA_TRY(cudaMalloc(&obj->d_woutputs, rows));
            NB200_CUDA_TRY(cudaMalloc(&obj->d_cluster, static_cast<size_t>(N) * sizeof(int)));
            NB200_CUDA_TRY(cudaMemcpy(obj->d_soutputs, soutputs, rows, cudaMemcpyHostToDevice));
            NB200_CUDA_TRY(cudaMemcpy(obj->d_woutputs, woutputs, rows, cudaMemcpyHostToDevice));
            NB200_CUDA_TRY(cudaMemcpy(obj->d_cluster, cluster, static_cast<size_t>(N) * sizeof(int), cudaMemcpyHostToDevice));
        }
        const size_t n = static_cast<size_t>(obj->n);
        NB200_CUDA_TRY(cudaMalloc(&obj->d_x, (n + 2) * sizeof(double)));
        NB200_CUDA_TRY(cudaMalloc(&obj->d_out, (n + 2) * sizeof(double)));
        NB200_CUDA_TRY(cudaMalloc(&obj->d_partials, static_cast<size_t>(obj->blocks) * (1 + obj->n_out) * sizeof(double)));
        NB200_CUDA_TRY(cudaMalloc(&obj->d_reduced, static_cast<size_t>(1 + obj->n_out) * sizeof(double)));
        NB200_CUDA_TRY(cudaMallocHost(&obj->h_x, (n + 2) * sizeof(double)));
        NB200_CUDA_TRY(cudaMallocHost(&obj->h_out, (n + 2) * sizeof(double)));
        if (obj->smem > 48 * 1024)
        {
            NB200_CUDA_TRY(cudaFuncSetAttribute(gboost_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, obj->smem));
            NB200_CUDA_TRY(cudaFuncSetAttribute(gboost_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, obj->smem));
        }
        return NB200_OK;
    };
    const int status = build();
    if (status != NB200_OK)
    {
        const std::string message = g_last_error;
        nb200_gboost_release(obj);
        g_last_error = message;
        return status;
    }
    *out_obj = obj;
    return NB200_OK;
}

int64_t nb200_gboost_size(const nb200_gboost* obj)
{
    return obj != nullptr ? obj->n : 0;
}

int64_t nb200_gboost_launches(const nb200_gboost* obj)
{
    return obj != nullptr ? obj->launches : 0;
}

int nb200_gboost_vgrad(nb200_gboost* obj, const double* x, const int64_t n, double* gx, double* fx)
{
    NB200_REQUIRE(obj != nullptr && x != nullptr && fx != nullptr, "function: null argument");
    NB200_REQUIRE(n == obj->n, "function: invalid input size, expecting (" + std::to_string(obj->n) + ",), got (" +
                                   std::to_string(n) + ",) instead!");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    const size_t bytes = static_cast<size_t>(n) * sizeof(double);
    std::memcpy(obj->h_x, x, bytes);
    NB200_CUDA_TRY(cudaMemcpyAsync(obj->d_x, obj->h_x, bytes, cudaMemcpyHostToDevice, obj->stream));
    gboost_params p{};
    p.targets    = obj->d_targets;
    p.soutputs   = obj->d_soutputs;
    p.woutputs   = obj->d_woutputs;
    p.cluster    = obj->d_cluster;
    p.x          = obj->d_x;
    p.partials   = obj->d_partials;
    p.gx         = obj->d_out;
    p.N          = obj->N;
    p.T          = static_cast<int>(obj->T);
    p.K          = static_cast<int>(obj->K);
    p.kind       = obj->kind;
    p.loss       = obj->loss;
    p.loss_param = obj->loss_param;
    if (gx != nullptr)
    {
        gboost_kernel<true><<<obj->blocks, kGbWarps * 32, obj->smem, obj->stream>>>(p);
    }
    else
    {
        gboost_kernel<false><<<obj->blocks, kGbWarps * 32, obj->smem, obj->stream>>>(p);
    }
    NB200_CUDA_TRY(cudaGetLastError());
    const int64_t cols = (gx != nullptr) ? 1 + obj->n_out : 1;
    reduce_partials_kernel<<<static_cast<unsigned>((cols + 127) / 128), 128, 0, obj->stream>>>(
        obj->d_partials, obj->blocks, 1 + obj->n_out, cols, obj->d_reduced);
    NB200_CUDA_TRY(cudaGetLastError());
    // the per-sample gradients of the grads function are already in d_out; bias / scale gradients come from the sums
    double* g_dev = (gx != nullptr && obj->kind != GBOOST_GRADS) ? obj->d_out : nullptr;
    gboost_finish_kernel<<<static_cast<unsigned>((std::max<int64_t>(obj->n_out, 1) + 255) / 256), 256, 0, obj->stream>>>(
        obj->d_reduced, obj->n_out, obj->N, g_dev, obj->d_out + n);
    NB200_CUDA_TRY(cudaGetLastError());
    obj->launches += 3;
    if (gx != nullptr)
    {
        NB200_CUDA_TRY(cudaMemcpyAsync(obj->h_out, obj->d_out, bytes, cudaMemcpyDeviceToHost, obj->stream));
    }
    NB200_CUDA_TRY(cudaMemcpyAsync(obj->h_out + n, obj->d_out + n, sizeof(double), cudaMemcpyDeviceToHost, obj->stream));
    NB200_CUDA_TRY(cudaStreamSynchronize(obj->stream));
    if (gx != nullptr)
    {
        std::memcpy(gx, obj->h_out, bytes);
    }
    *fx = obj->h_out[n];
    return NB200_OK;
}

// the roofline denominators of this board, measured in the run that quotes them (probe.cuh)
int nb200_probe_peak(nb200_ctx* ctx, const int kind, double* value)
{
    NB200_REQUIRE(ctx != nullptr && value != nullptr, "probe: null argument");
    NB200_REQUIRE(kind >= NB200_PEAK_DMMA_TFLOPS && kind <= NB200_PEAK_HBM_READ_GBS, "probe: unknown kind");
    if (!ctx->parts.empty())
    {
        ctx = ctx->parts[0];
    }
    NB200_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaEvent_t    e0 = nullptr, e1 = nullptr;
    double*        out = nullptr;
    unsigned char* src = nullptr;
    const auto     run = [&]() -> int
    {
        NB200_CUDA_TRY(cudaEventCreate(&e0));
        NB200_CUDA_TRY(cudaEventCreate(&e1));
        const int blocks = ctx->sm_count * 8, threads = 256;
        NB200_CUDA_TRY(cudaMalloc(&out, static_cast<size_t>(blocks) * threads * sizeof(double)));
        constexpr int     kTile = 65536, kStages = 2;
        const int64_t     tiles = (4LL << 30) / kTile; // 4 GiB: far beyond L2
        const int         smem  = 128 + kStages * kTile;
        if (kind == NB200_PEAK_HBM_READ_GBS)
        {
            NB200_CUDA_TRY(cudaMalloc(&src, static_cast<size_t>(tiles) * kTile));
            NB200_CUDA_TRY(cudaMemsetAsync(src, 1, static_cast<size_t>(tiles) * kTile, ctx->stream));
            NB200_CUDA_TRY(cudaFuncSetAttribute(probe_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        }
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) // the first one warms up
        {
            NB200_CUDA_TRY(cudaEventRecord(e0, ctx->stream));
            switch (kind)
            {
            case NB200_PEAK_DMMA_TFLOPS: probe_dmma_kernel<<<blocks, threads, 0, ctx->stream>>>(out, 1.0000001, 1e-9); break;
            case NB200_PEAK_DFMA_TFLOPS: probe_dfma_kernel<<<blocks, threads, 0, ctx->stream>>>(out, 1.0000001, 1e-9); break;
            default: probe_read_kernel<<<ctx->sm_count, 128, smem, ctx->stream>>>(src, tiles, out); break;
            }
            NB200_CUDA_TRY(cudaGetLastError());
            NB200_CUDA_TRY(cudaEventRecord(e1, ctx->stream));
            NB200_CUDA_TRY(cudaEventSynchronize(e1));
            float ms = 0.0f;
            NB200_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0)
            {
                best = std::min(best, ms);
            }
        }
        const double seconds = best * 1e-3;
        const double warps   = static_cast<double>(blocks) * threads / 32.0;
        switch (kind)
        {
        case NB200_PEAK_DMMA_TFLOPS: *value = warps * kProbeIters * 8.0 * (2.0 * 8 * 8 * 4) / seconds * 1e-12; break;
        case NB200_PEAK_DFMA_TFLOPS: *value = warps * 32.0 * kProbeIters * 16.0 * 2.0 / seconds * 1e-12; break;
        default: *value = static_cast<double>(tiles) * kTile / seconds * 1e-9; break;
        }
        return NB200_OK;
    };
    const int status = run();
    cudaFree(out);
    cudaFree(src);
    if (e0 != nullptr)
    {
        cudaEventDestroy(e0);
    }
    if (e1 != nullptr)
    {
        cudaEventDestroy(e1);
    }
    return status;
}

int nb200_objective_set_variant(nb200_obj* obj, const int variant)
{
    NB200_REQUIRE(obj != nullptr, "objective: null argument");
    if (!obj->parts.empty())
    {
        for (nb200_obj* part : obj->parts)
        {
            const int status = nb200_objective_set_variant(part, variant);
            if (status != NB200_OK)
            {
                return status;
            }
        }
        return group_connect(obj);
    }
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    NB200_CUDA_TRY(cudaStreamSynchronize(obj->stream));
    return setup_plan(obj, variant);
}

int nb200_objective_describe(const nb200_obj* obj, char* buffer, const int64_t size)
{
    NB200_REQUIRE(obj != nullptr && buffer != nullptr && size > 0, "objective: null argument");
    std::string text;
    if (!obj->parts.empty())
    {
        const int status = nb200_objective_describe(obj->parts[0], buffer, size);
        if (status != NB200_OK)
        {
            return status;
        }
        text = std::string(buffer) + " | group of " + std::to_string(obj->parts.size()) + " devices, sums " +
               (obj->group_fused ? (obj->parts[0]->peer_wait == 0
                                        ? "published into the first device's mailbox over peer memory (stream wait)"
                                        : "exchanged inside the launch over peer memory")
                                 : "gathered by peer copies (two-phase)");
        std::strncpy(buffer, text.c_str(), static_cast<size_t>(size) - 1);
        buffer[size - 1] = '\0';
        return NB200_OK;
    }
    if (obj->use_t1)
    {
        const t1_variant& var = kT1Variants[obj->t1.variant];
        text = "t1 variant=" + std::to_string(obj->t1.variant) + " V=" + std::to_string(var.V) +
               " KC=" + std::to_string(var.KC) + " WS=" + std::to_string(var.WS) + " NCW=" + std::to_string(var.NCW) +
               " RW=" + std::to_string(var.RW) + " ALT=" + std::to_string(var.ALT) +
               " grid=" + std::to_string(obj->t1.grid) + " rows_per_stage=" + std::to_string(obj->t1.rows_per_stage) +
               " stages=" + std::to_string(obj->t1.stages) + " stage_bytes=" + std::to_string(obj->t1.stage_stride) +
               " smem=" + std::to_string(obj->t1.smem_bytes);
    }
    else if (obj->use_fd)
    {
        const fd_variant& var = kFdVariants[obj->fd.variant];
        text = std::string(var.NSETS < 0 ? "fr" : var.NSETS >= 20 ? "fl-pipe" : var.NSETS >= 10 ? "fl" : "fd") + " NT8=" + std::to_string(var.NT8) + " NS8=" + std::to_string(var.NS8) + " NSETS=" +
               std::to_string(var.NSETS) + " grid=" +
               std::to_string(obj->fd.grid) + " ld=" + std::to_string(obj->fd.ld) + " stages=" +
               std::to_string(obj->fd.stages) + " stage_bytes=" + std::to_string(obj->fd.stage_stride) +
               " smem=" + std::to_string(obj->fd.smem_bytes);
    }
    else if (obj->use_ft)
    {
        const ft_variant& var = kFtVariants[obj->ft.variant];
        text = "ft T=" + std::to_string(var.T) + " KC=" + std::to_string(var.KC) + " RW=" + std::to_string(var.RW) +
               " NS=" + std::to_string(var.NS) +
               " grid=" + std::to_string(obj->ft.grid) + " rows_per_stage=" + std::to_string(obj->ft.rows_per_stage) +
               " stages=" + std::to_string(obj->ft.stages) + " stage_bytes=" + std::to_string(obj->ft.stage_stride) +
               " smem=" + std::to_string(obj->ft.smem_bytes);
    }
    else if (obj->use_mt)
    {
        text = std::string(obj->mt_tma ? "dmma-tma" : "dmma") + " NT8=" + std::to_string(kMtVariants[obj->mt_variant].NT8) + " fwd_grid=" +
               std::to_string(obj->mt_fwd_grid) + " fwd_stages=" + std::to_string(obj->mt_fwd_stages) +
               " fwd_smem=" + std::to_string(obj->mt_fwd_smem) + " bwd_grid=" + std::to_string(obj->mt_bwd_grid) +
               " splits=" + std::to_string(obj->mt_splits) + " bwd_stages=" + std::to_string(obj->mt_bwd_stages) +
               " bwd_smem=" + std::to_string(obj->mt_bwd_smem) + " t_ld=" + std::to_string(obj->mt_t_ld);
    }
    else
    {
        text = "generic blocks=" + std::to_string(obj->gen_blocks) + " splits=" + std::to_string(obj->gen_splits) +
               " smem=" + std::to_string(obj->gen_smem);
    }
    std::strncpy(buffer, text.c_str(), static_cast<size_t>(size) - 1);
    buffer[size - 1] = '\0';
    return NB200_OK;
}

int nb200_lbfgs_minimize(nb200_obj* obj, double* x, const int64_t n, const double epsilon, const int64_t max_evals,
                         const int64_t history, const int64_t n_global, nb200_allreduce_fn allreduce,
                         void* allreduce_user, double* fx, double* gradient_test, int64_t* stats, int* status)
{
    NB200_REQUIRE(obj != nullptr && x != nullptr && fx != nullptr && gradient_test != nullptr && stats != nullptr &&
                      status != nullptr,
                  "lbfgs: null argument");
    NB200_REQUIRE(n == obj->n, "solver: incompatible initial point (" + std::to_string(n) + " dimensions), expecting " +
                                   std::to_string(obj->n) + " dimensions!");
    NB200_REQUIRE(epsilon > 0.0 && max_evals > 0 && history >= 1 && history <= 1000, "lbfgs: invalid parameters");
    NB200_REQUIRE(allreduce == nullptr || n_global > 0, "lbfgs: a sharded solve needs the global sample count");
    NB200_REQUIRE(allreduce != nullptr || n_global == 0 || n_global == obj->data->N || obj->peer_world > 1,
                  "lbfgs: a sharded solve needs the all-reduce callback or connected peers");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    if (!obj->parts.empty())
    {
        NB200_REQUIRE(allreduce == nullptr && (n_global == 0 || n_global == obj->data->N),
                      "lbfgs: a device group holds all samples; no all-reduce callback applies");
        device_lbfgs_t solver(obj->parts[0], obj->data->N, nullptr, nullptr, solver_method_t{}, obj);
        return solver.run(x, epsilon, max_evals, history, fx, gradient_test, stats, status);
    }
    device_lbfgs_t solver(obj, n_global > 0 ? n_global : obj->data->N, allreduce, allreduce_user);
    return solver.run(x, epsilon, max_evals, history, fx, gradient_test, stats, status);
}

int nb200_osga_minimize(nb200_obj* obj, double* x, const int64_t n, const double epsilon, const int64_t max_evals,
                        const int64_t patience, const double lambda, const double alpha_max, const double kappa_prime,
                        const double kappa, const double strong_convexity, const int64_t n_global, double* fx,
                        double* eta, int64_t* stats, int* status)
{
    NB200_REQUIRE(obj != nullptr && x != nullptr && fx != nullptr && eta != nullptr && stats != nullptr && status != nullptr,
                  "osga: null argument");
    NB200_REQUIRE(n == obj->n, "solver: incompatible initial point (" + std::to_string(n) + " dimensions), expecting " +
                                   std::to_string(obj->n) + " dimensions!");
    NB200_REQUIRE(epsilon > 0.0 && max_evals > 0 && patience > 0 && lambda > 0.0 && lambda < 1.0 && alpha_max > 0.0 &&
                      alpha_max < 1.0 && kappa_prime > 0.0 && kappa_prime < kappa && strong_convexity >= 0.0,
                  "osga: invalid parameters");
    NB200_REQUIRE(n_global == 0 || n_global == obj->data->N || (obj->parts.empty() && obj->peer_world > 1),
                  "osga: a sharded solve needs connected peers (or a device group)");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    osga_params_t q;
    q.epsilon          = epsilon;
    q.max_evals        = max_evals;
    q.patience         = patience;
    q.lambda           = lambda;
    q.alpha_max        = alpha_max;
    q.kappa_prime      = kappa_prime;
    q.kappa            = kappa;
    q.strong_convexity = strong_convexity;
    nb200_obj* const first = obj->parts.empty() ? obj : obj->parts[0];
    device_osga_t    solver(first, obj->parts.empty() ? nullptr : obj, n_global > 0 ? n_global : obj->data->N);
    return solver.run(x, q, fx, eta, stats, status);
}

int nb200_cgd_minimize(nb200_obj* obj, double* x, const int64_t n, const int beta, const double epsilon,
                       const int64_t max_evals, const double orthotest, const double eta, const int64_t n_global,
                       nb200_allreduce_fn allreduce, void* allreduce_user, double* fx, double* gradient_test,
                       int64_t* stats, int* status)
{
    NB200_REQUIRE(obj != nullptr && x != nullptr && fx != nullptr && gradient_test != nullptr && stats != nullptr &&
                      status != nullptr,
                  "cgd: null argument");
    NB200_REQUIRE(n == obj->n, "solver: incompatible initial point (" + std::to_string(n) + " dimensions), expecting " +
                                   std::to_string(obj->n) + " dimensions!");
    NB200_REQUIRE(beta >= 0 && beta < CGD_COUNT, "cgd: unknown beta formula");
    NB200_REQUIRE(epsilon > 0.0 && max_evals > 0 && orthotest > 0.0 && orthotest < 1.0 && eta > 0.0 && eta < 1e+6,
                  "cgd: invalid parameters");
    NB200_REQUIRE(allreduce == nullptr || n_global > 0, "cgd: a sharded solve needs the global sample count");
    NB200_REQUIRE(allreduce != nullptr || n_global == 0 || n_global == obj->data->N || obj->peer_world > 1,
                  "cgd: a sharded solve needs the all-reduce callback or connected peers");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    solver_method_t method;
    method.cgd       = true;
    method.beta      = beta;
    method.c2        = 1e-1; // solver::tolerance of the cgd family (cgd.cpp:77)
    method.orthotest = orthotest;
    method.eta       = eta;
    if (!obj->parts.empty())
    {
        NB200_REQUIRE(allreduce == nullptr && (n_global == 0 || n_global == obj->data->N),
                      "cgd: a device group holds all samples; no all-reduce callback applies");
        device_lbfgs_t solver(obj->parts[0], obj->data->N, nullptr, nullptr, method, obj);
        return solver.run(x, epsilon, max_evals, 1, fx, gradient_test, stats, status);
    }
    device_lbfgs_t solver(obj, n_global > 0 ? n_global : obj->data->N, allreduce, allreduce_user, method);
    return solver.run(x, epsilon, max_evals, 1, fx, gradient_test, stats, status);
}

// the one-pass tensor-pipe kernel for a batch of Kp single-target models (Y has one column)
static bool plan_fd_batch(const int64_t N, const int64_t D, const int64_t Kp, const int sm_count, const int smem_optin,
                          fd_plan& plan)
{
    return Kp <= 16 && D >= 200 && plan_fd(N, D, Kp, sm_count, smem_optin, plan, false, env_int("NB200_FL", 1));
}

static int ensure_batch(nb200_obj* obj, const int64_t Kp)
{
    batch_state&      bs   = obj->batch;
    const nb200_data* data = obj->data;
    const int64_t     N = data->N, D = data->D, n = obj->n;
    if (bs.Kp == Kp)
    {
        return NB200_OK;
    }
    bs.release();
    mt_geometry geo;
    fd_plan     one;
    const bool  one_pass = plan_fd_batch(N, D, Kp, data->sm_count, data->smem_optin, one);
    if (one_pass)
    {
        const fd_variant& var = kFdVariants[one.variant];
        NB200_CUDA_TRY(cudaFuncSetAttribute(var.grad, cudaFuncAttributeMaxDynamicSharedMemorySize, one.smem_bytes));
        NB200_CUDA_TRY(cudaFuncSetAttribute(var.value, cudaFuncAttributeMaxDynamicSharedMemorySize, one.smem_bytes));
        geo.fwd_grid = one.grid; // rows of the two partial buffers
        geo.splits   = one.grid;
    }
    else
    {
        if (!plan_mt(N, D, Kp, data->sm_count, data->smem_optin, geo))
        {
            return fail(NB200_ERR_UNSUPPORTED, "batch: this (D, K) is not covered by the tensor-pipe kernels");
        }
        const mt_variant& var = kMtVariants[geo.variant];
        NB200_CUDA_TRY(cudaFuncSetAttribute(var.grad, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.fwd_smem));
        NB200_CUDA_TRY(cudaFuncSetAttribute(var.value, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.fwd_smem));
        if (geo.tma_stages >= 2)
        {
            NB200_CUDA_TRY(cudaFuncSetAttribute(var.grad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.tma_smem));
            NB200_CUDA_TRY(cudaFuncSetAttribute(var.value_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.tma_smem));
        }
        NB200_CUDA_TRY(cudaFuncSetAttribute(var.backward, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.bwd_smem));
    }
    const size_t K = static_cast<size_t>(Kp);
    NB200_CUDA_TRY(cudaMalloc(&bs.d_x, K * (n + 2) * sizeof(double)));
    NB200_CUDA_TRY(cudaMalloc(&bs.d_W, K * D * sizeof(double)));
    NB200_CUDA_TRY(cudaMalloc(&bs.d_Wp, static_cast<size_t>((D + kMtBK - 1) / kMtBK) * K * kMtLd * sizeof(double)));
    NB200_CUDA_TRY(cudaMalloc(&bs.d_b, K * sizeof(double)));
    NB200_CUDA_TRY(cudaMalloc(&bs.d_l, 2 * K * sizeof(double)));
    if (!one_pass)
    {
        NB200_CUDA_TRY(cudaMalloc(&bs.d_dL, static_cast<size_t>(N) * K * sizeof(double)));
    }
    NB200_CUDA_TRY(cudaMalloc(&bs.d_partials_fb, static_cast<size_t>(geo.fwd_grid) * (1 + 2 * K) * sizeof(double)));
    NB200_CUDA_TRY(cudaMalloc(&bs.d_partials_gw, static_cast<size_t>(geo.splits) * K * D * sizeof(double)));
    NB200_CUDA_TRY(cudaMalloc(&bs.d_red_fb, (1 + 2 * K) * sizeof(double)));
    NB200_CUDA_TRY(cudaMalloc(&bs.d_red_gw, K * D * sizeof(double)));
    NB200_CUDA_TRY(cudaMallocHost(&bs.h_in, (K * (n + 2) + 2 * K) * sizeof(double)));
    NB200_CUDA_TRY(cudaHostAlloc(&bs.h_out, K * (n + 1) * sizeof(double), cudaHostAllocMapped));
    NB200_CUDA_TRY(cudaHostGetDevicePointer(&bs.h_out_dev, bs.h_out, 0));
    NB200_CUDA_TRY(cudaDeviceSynchronize());
    bs.Kp       = Kp;
    bs.geo      = geo;
    bs.one_pass = one_pass;
    bs.fd       = one;
    return NB200_OK;
}

static int vgrad_batch_impl(nb200_obj* obj, const int64_t K, const double* x, const int64_t n, const double* l1,
                            const double* l2, double* gx, double* fx, const int64_t Kp_min)
{
    NB200_REQUIRE(obj != nullptr && x != nullptr && l1 != nullptr && l2 != nullptr && fx != nullptr,
                  "batch: null argument");
    NB200_REQUIRE(K >= 1 && K <= 128, "batch: the number of models must be in [1, 128]");
    NB200_REQUIRE(n == obj->n, "function: invalid input size, expecting (" + std::to_string(obj->n) + ",), got (" +
                                   std::to_string(n) + ",) instead!");
    NB200_REQUIRE(obj->data->T == 1, "batch: the models must be single-target");
    NB200_CUDA_TRY(cudaSetDevice(obj->device));
    const nb200_data* data = obj->data;
    const int64_t     N = data->N, D = data->D;
    // padded model count (even; the chunks of a split batch all use 8 so that the scratch is allocated once)
    const int64_t     Kp = std::max<int64_t>((K % 2 == 0) ? K : K + 1, Kp_min);

    // fewer than 6 models: K single passes at ~6.4 TB/s beat the two tensor-pipe passes (measured cross-over K ~ 5)
    // 9 .. 24 models: chunks of 8 on the one-pass kernel with ONE target tile (2.5 ms per chunk at 2^20 x 1024) beat
    // both the two-target-tile one-pass kernel and the two-pass pair (7.2-7.6 ms for 12-16 models, ~8 ms for 32)
    {
        fd_plan chunk;
        if (Kp_min == 0 && K > 8 && K <= 24 && plan_fd_batch(N, D, 8, data->sm_count, data->smem_optin, chunk))
        {
            int status = NB200_OK;
            for (int64_t k0 = 0; k0 < K && status == NB200_OK; k0 += 8)
            {
                const int64_t kc = std::min<int64_t>(8, K - k0);
                status = vgrad_batch_impl(obj, kc, x + k0 * n, n, l1 + k0, l2 + k0, gx != nullptr ? gx + k0 * n : nullptr,
                                          fx + k0, 8);
            }
            return status;
        }
    }
    // ... unless the one-pass kernel covers the shape: it reads X once for any K >= 2
    mt_geometry probe;
    fd_plan     probe_one;
    const bool  one_pass_fits = K >= 2 && plan_fd_batch(N, D, Kp, data->sm_count, data->smem_optin, probe_one);
    if (!one_pass_fits && (Kp < 6 || !plan_mt(N, D, Kp, data->sm_count, data->smem_optin, probe)))
    {
        // shapes outside the tensor-pipe kernels: K passes of the single-model path (still the GPU, just not fused)
        const double keep1 = obj->l1, keep2 = obj->l2;
        int          status = NB200_OK;
        for (int64_t k = 0; k < K && status == NB200_OK; ++k)
        {
            obj->l1 = l1[k];
            obj->l2 = l2[k];
            status  = nb200_vgrad(obj, x + k * n, n, gx != nullptr ? gx + k * n : nullptr, fx + k);
        }
        obj->l1 = keep1;
        obj->l2 = keep2;
        return status;
    }

    int status = ensure_batch(obj, Kp);
    if (status != NB200_OK)
    {
        return status;
    }
    batch_state&      bs   = obj->batch;
    const mt_geometry geo  = bs.geo;
    const mt_variant& var  = kMtVariants[geo.variant];
    const bool        grad = gx != nullptr;
    cudaStream_t      st   = obj->stream;

    // upload the K parameter vectors and their regularisation factors
    std::memcpy(bs.h_in, x, static_cast<size_t>(K * n) * sizeof(double));
    double* h_l = bs.h_in + K * n;
    for (int64_t k = 0; k < Kp; ++k)
    {
        h_l[k]      = l1[std::min(k, K - 1)];
        h_l[Kp + k] = l2[std::min(k, K - 1)];
    }
    NB200_CUDA_TRY(cudaMemcpyAsync(bs.d_x, bs.h_in, static_cast<size_t>(K * n) * sizeof(double), cudaMemcpyHostToDevice, st));
    NB200_CUDA_TRY(cudaMemcpyAsync(bs.d_l, h_l, static_cast<size_t>(2 * Kp) * sizeof(double), cudaMemcpyHostToDevice, st));
    batch_pack_kernel<<<dim3(static_cast<unsigned>(std::min<int64_t>((D + 255) / 256, 64)), static_cast<unsigned>(Kp)),
                        256, 0, st>>>(static_cast<int>(K), static_cast<int>(Kp), D, n, obj->flavour, bs.d_x,
                                      obj->d_fixed_bias, bs.d_W, bs.d_b);
    NB200_CUDA_TRY(cudaGetLastError());
    obj->launches += 1;

    if (bs.one_pass)
    {
        status = timing_begin(obj);
        if (status != NB200_OK)
        {
            return status;
        }
        const fd_variant& one = kFdVariants[bs.fd.variant];
        fd_params         q{};
        q.X            = data->X;
        q.Y            = data->Y;
        q.W            = bs.d_W;
        q.b            = bs.d_b;
        q.partials     = bs.d_partials_fb;
        q.partials_gw  = bs.d_partials_gw;
        q.y_broadcast  = 1;
        q.N            = N;
        q.num_tiles    = bs.fd.num_tiles;
        q.D            = static_cast<int>(D);
        q.T            = static_cast<int>(Kp);
        q.slice        = bs.fd.slice;
        q.ld           = bs.fd.ld;
        q.stages       = bs.fd.stages;
        q.stage_stride = bs.fd.stage_stride;
        q.y_offset     = bs.fd.y_offset;
        q.w_offset     = bs.fd.w_offset;
        q.stage_offset = bs.fd.stage_offset;
        q.loss         = obj->loss;
        q.loss_param   = obj->loss_param;
        (grad ? one.grad : one.value)<<<bs.fd.grid, one.threads, bs.fd.smem_bytes, st>>>(q);
        NB200_CUDA_TRY(cudaGetLastError());
        obj->launches += 1;
        status = timing_end(obj);
        if (status != NB200_OK)
        {
            return status;
        }
    }
    else
    {
    mt_pack_weights_kernel<<<static_cast<unsigned>(std::min<int64_t>((Kp * D + 255) / 256, 2048)), 256, 0, st>>>(
        bs.d_W, static_cast<int>(Kp), static_cast<int>(D), bs.d_Wp);
    NB200_CUDA_TRY(cudaGetLastError());
    obj->launches += 1;

    status = timing_begin(obj);
    if (status != NB200_OK)
    {
        return status;
    }
    mt_forward_params f{};
    f.X               = data->X;
    f.Y               = data->Y;
    f.W               = bs.d_Wp;
    f.b               = bs.d_b;
    f.dL              = grad ? bs.d_dL : nullptr;
    f.partials_fb     = bs.d_partials_fb;
    f.N               = N;
    f.D               = static_cast<int>(D);
    f.T               = static_cast<int>(Kp);
    f.loss            = obj->loss;
    f.loss_param      = obj->loss_param;
    f.y_broadcast     = 1;
    f.per_target_loss = 1;
    f.fb_stride       = static_cast<int>(1 + 2 * Kp);
    if (data->x_map_ok && geo.tma_stages >= 2 && env_int("NB200_MT_LEGACY_FORWARD", 0) == 0)
    {
        f.stages = geo.tma_stages;
        (grad ? var.grad_tma : var.value_tma)<<<geo.fwd_grid, kMtThreads, geo.tma_smem, st>>>(f, data->x_map);
    }
    else
    {
        f.stages = geo.fwd_stages;
        (grad ? var.grad : var.value)<<<geo.fwd_grid, kMtThreads, geo.fwd_smem, st>>>(f);
    }
    NB200_CUDA_TRY(cudaGetLastError());
    obj->launches += 1;
    if (grad)
    {
        mt_backward_params b{};
        b.X           = data->X;
        b.dL          = bs.d_dL;
        b.partials_gw = bs.d_partials_gw;
        b.N           = N;
        b.D           = static_cast<int>(D);
        b.T           = static_cast<int>(Kp);
        b.t_ld        = geo.t_ld;
        b.splits      = geo.splits;
        b.stages      = geo.bwd_stages;
        var.backward<<<geo.bwd_grid, kMtThreads, geo.bwd_smem, st>>>(b);
        NB200_CUDA_TRY(cudaGetLastError());
        obj->launches += 1;
    }
    status = timing_end(obj);
    if (status != NB200_OK)
    {
        return status;
    }
    } // two-pass pair
    const int64_t cols_fb = 1 + 2 * Kp;
    reduce_partials_kernel<<<static_cast<unsigned>((cols_fb + 127) / 128), 128, 0, st>>>(
        bs.d_partials_fb, geo.fwd_grid, cols_fb, cols_fb, bs.d_red_fb);
    NB200_CUDA_TRY(cudaGetLastError());
    obj->launches += 1;
    if (grad)
    {
        reduce_partials_kernel<<<static_cast<unsigned>((Kp * D + 127) / 128), 128, 0, st>>>(
            bs.d_partials_gw, geo.splits, Kp * D, Kp * D, bs.d_red_gw);
        NB200_CUDA_TRY(cudaGetLastError());
        obj->launches += 1;
    }
    batch_finish_kernel<<<static_cast<unsigned>(K), 256, 0, st>>>(static_cast<int>(Kp), D, n, N, obj->flavour, bs.d_red_fb,
                                                                   bs.d_red_gw, bs.d_x, bs.d_l, bs.d_l + Kp, grad ? 1 : 0,
                                                                   bs.h_out_dev);
    NB200_CUDA_TRY(cudaGetLastError());
    obj->launches += 1;
    NB200_CUDA_TRY(cudaStreamSynchronize(st));
    collect_timing(obj);
    for (int64_t k = 0; k < K; ++k)
    {
        if (grad)
        {
            std::memcpy(gx + k * n, bs.h_out + k * (n + 1), static_cast<size_t>(n) * sizeof(double));
        }
        fx[k] = bs.h_out[k * (n + 1) + n];
    }
    return NB200_OK;
}

int nb200_vgrad_batch(nb200_obj* obj, const int64_t K, const double* x, const int64_t n, const double* l1,
                      const double* l2, double* gx, double* fx)
{
    NB200_NOT_ON_A_GROUP(obj, "nb200_vgrad_batch");
    return vgrad_batch_impl(obj, K, x, n, l1, l2, gx, fx, 0);
}

int nb200_predict(nb200_ctx* ctx, const nb200_data* data, const double* W, const double* b, double* outputs)
{
    NB200_REQUIRE(ctx != nullptr && data != nullptr && W != nullptr && b != nullptr && outputs != nullptr,
                  "predict: null argument");
    NB200_REQUIRE(ctx->parts.size() == data->parts.size(), "predict: context and dataset are not on the same device group");
    for (size_t r = 0; r < data->parts.size(); ++r) // a device group: every device predicts its own rows
    {
        const int status = nb200_predict(ctx->parts[r], data->parts[r], W, b, outputs + data->row0[r] * data->T);
        if (status != NB200_OK || r + 1 == data->parts.size())
        {
            return status;
        }
    }
    NB200_REQUIRE(ctx->device == data->device, "predict: context and dataset live on different devices");
    NB200_CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t N = data->N, D = data->D, T = data->T;
    double *      dW = nullptr, *db = nullptr, *dO = nullptr;
    const auto    run = [&]() -> int
    {
        NB200_CUDA_TRY(cudaMalloc(&dW, static_cast<size_t>(T) * D * sizeof(double)));
        NB200_CUDA_TRY(cudaMalloc(&db, static_cast<size_t>(T) * sizeof(double)));
        NB200_CUDA_TRY(cudaMalloc(&dO, static_cast<size_t>(N) * T * sizeof(double)));
        NB200_CUDA_TRY(cudaMemcpyAsync(dW, W, static_cast<size_t>(T) * D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        NB200_CUDA_TRY(cudaMemcpyAsync(db, b, static_cast<size_t>(T) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        predict_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(data->X, dW, db, N, D, T, dO);
        NB200_CUDA_TRY(cudaGetLastError());
        NB200_CUDA_TRY(cudaMemcpyAsync(outputs, dO, static_cast<size_t>(N) * T * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        NB200_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return NB200_OK;
    };
    const int status = run();
    cudaFree(dW);
    cudaFree(db);
    cudaFree(dO);
    return status;
}

static int loss_rows(nb200_ctx* ctx, const int loss, const double loss_param, const double* targets,
                     const double* outputs, const int64_t N, const int64_t T, double* values, double* vgrads)
{
    NB200_REQUIRE(ctx != nullptr && targets != nullptr && outputs != nullptr, "loss: null argument");
    if (!ctx->parts.empty())
    {
        ctx = ctx->parts[0]; // host buffers in, host buffers out: one device of the group does it
    }
    NB200_REQUIRE(loss >= 0 && loss < NB200_LOSS_COUNT, "loss: unknown loss");
    NB200_REQUIRE(N >= 0 && T > 0, "loss: invalid dimensions");
    if (N == 0)
    {
        return NB200_OK;
    }
    NB200_CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t bytes = static_cast<size_t>(N) * T * sizeof(double);
    double *     dT = nullptr, *dO = nullptr, *dV = nullptr, *dG = nullptr;
    const auto   run = [&]() -> int
    {
        NB200_CUDA_TRY(cudaMalloc(&dT, bytes));
        NB200_CUDA_TRY(cudaMalloc(&dO, bytes));
        if (values != nullptr)
        {
            NB200_CUDA_TRY(cudaMalloc(&dV, static_cast<size_t>(N) * sizeof(double)));
        }
        if (vgrads != nullptr)
        {
            NB200_CUDA_TRY(cudaMalloc(&dG, bytes));
        }
        NB200_CUDA_TRY(cudaMemcpyAsync(dT, targets, bytes, cudaMemcpyHostToDevice, ctx->stream));
        NB200_CUDA_TRY(cudaMemcpyAsync(dO, outputs, bytes, cudaMemcpyHostToDevice, ctx->stream));
        loss_rows_kernel<<<static_cast<unsigned>((N + 255) / 256), 256, 0, ctx->stream>>>(loss, loss_param, dT, dO, N,
                                                                                           T, dV, dG);
        NB200_CUDA_TRY(cudaGetLastError());
        if (values != nullptr)
        {
            NB200_CUDA_TRY(cudaMemcpyAsync(values, dV, static_cast<size_t>(N) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (vgrads != nullptr)
        {
            NB200_CUDA_TRY(cudaMemcpyAsync(vgrads, dG, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        }
        NB200_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return NB200_OK;
    };
    const int status = run();
    cudaFree(dT);
    cudaFree(dO);
    cudaFree(dV);
    cudaFree(dG);
    return status;
}

int nb200_loss_value(nb200_ctx* ctx, const int loss, const double loss_param, const double* targets,
                     const double* outputs, const int64_t N, const int64_t T, double* values)
{
    NB200_REQUIRE(values != nullptr, "loss: null value buffer");
    return loss_rows(ctx, loss, loss_param, targets, outputs, N, T, values, nullptr);
}

int nb200_loss_vgrad(nb200_ctx* ctx, const int loss, const double loss_param, const double* targets,
                     const double* outputs, const int64_t N, const int64_t T, double* vgrads)
{
    NB200_REQUIRE(vgrads != nullptr, "loss: null gradient buffer");
    return loss_rows(ctx, loss, loss_param, targets, outputs, N, T, nullptr, vgrads);
}

int nb200_loss_error(nb200_ctx* ctx, const int error_kind, const double* targets, const double* outputs,
                     const int64_t N, const int64_t T, double* errors)
{
    NB200_REQUIRE(ctx != nullptr && targets != nullptr && outputs != nullptr && errors != nullptr,
                  "loss error: null argument");
    NB200_REQUIRE(error_kind >= 0 && error_kind < NB200_ERROR_COUNT, "loss error: unknown error kind");
    NB200_REQUIRE(N >= 0 && T > 0, "loss error: invalid dimensions");
    if (!ctx->parts.empty())
    {
        ctx = ctx->parts[0];
    }
    if (N == 0)
    {
        return NB200_OK;
    }
    NB200_CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t bytes = static_cast<size_t>(N) * T * sizeof(double);
    double *     dT = nullptr, *dO = nullptr, *dE = nullptr;
    const auto   run = [&]() -> int
    {
        NB200_CUDA_TRY(cudaMalloc(&dT, bytes));
        NB200_CUDA_TRY(cudaMalloc(&dO, bytes));
        NB200_CUDA_TRY(cudaMalloc(&dE, static_cast<size_t>(N) * sizeof(double)));
        NB200_CUDA_TRY(cudaMemcpyAsync(dT, targets, bytes, cudaMemcpyHostToDevice, ctx->stream));
        NB200_CUDA_TRY(cudaMemcpyAsync(dO, outputs, bytes, cudaMemcpyHostToDevice, ctx->stream));
        loss_error_kernel<<<static_cast<unsigned>((N + 255) / 256), 256, 0, ctx->stream>>>(error_kind, dT, dO, N, T, dE);
        NB200_CUDA_TRY(cudaGetLastError());
        NB200_CUDA_TRY(cudaMemcpyAsync(errors, dE, static_cast<size_t>(N) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        NB200_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return NB200_OK;
    };
    const int status = run();
    cudaFree(dT);
    cudaFree(dO);
    cudaFree(dE);
    return status;
}

static int ctx_reserve(nb200_ctx* ctx, const size_t device_bytes, const size_t host_bytes)
{
    if (device_bytes > ctx->arena_bytes)
    {
        cudaFree(ctx->arena);
        ctx->arena       = nullptr;
        ctx->arena_bytes = 0;
        NB200_CUDA_TRY(cudaMalloc(&ctx->arena, device_bytes));
        ctx->arena_bytes = device_bytes;
    }
    if (host_bytes > ctx->staging_bytes)
    {
        cudaFreeHost(ctx->staging);
        ctx->staging       = nullptr;
        ctx->staging_bytes = 0;
        NB200_CUDA_TRY(cudaMallocHost(&ctx->staging, host_bytes));
        ctx->staging_bytes = host_bytes;
    }
    return NB200_OK;
}

int nb200_evaluate(nb200_ctx* ctx, const nb200_data* data, const int loss, const double loss_param,
                   const int error_kind, const double* W, const double* b, double* errors, double* values)
{
    NB200_REQUIRE(ctx != nullptr && data != nullptr && W != nullptr && b != nullptr, "evaluate: null argument");
    NB200_REQUIRE(errors != nullptr || values != nullptr, "evaluate: nothing to compute");
    NB200_REQUIRE(ctx->parts.size() == data->parts.size(), "evaluate: context and dataset are not on the same device group");
    for (size_t r = 0; r < data->parts.size(); ++r) // a device group: every device evaluates its own rows
    {
        const int status = nb200_evaluate(ctx->parts[r], data->parts[r], loss, loss_param, error_kind, W, b,
                                          errors != nullptr ? errors + data->row0[r] : nullptr,
                                          values != nullptr ? values + data->row0[r] : nullptr);
        if (status != NB200_OK || r + 1 == data->parts.size())
        {
            return status;
        }
    }
    NB200_REQUIRE(ctx->device == data->device, "evaluate: context and dataset live on different devices");
    NB200_REQUIRE(loss >= 0 && loss < NB200_LOSS_COUNT, "evaluate: unknown loss");
    NB200_REQUIRE(error_kind >= 0 && error_kind < NB200_ERROR_COUNT, "evaluate: unknown error kind");
    const int64_t N = data->N, D = data->D, T = data->T;
    if (N == 0)
    {
        return NB200_OK;
    }
    NB200_CUDA_TRY(cudaSetDevice(ctx->device));
    const std::lock_guard<std::mutex> guard(ctx->scratch_lock);

    // arena layout (256-byte aligned slices): W [T, D] | b [T] | outputs [N, T] | errors [N] | values [N]
    const auto   aligned = [](const size_t bytes) { return (bytes + 255) / 256 * 256; };
    const size_t rows = static_cast<size_t>(N) * sizeof(double);
    const size_t w_bytes = aligned(static_cast<size_t>(T) * D * sizeof(double));
    const size_t b_bytes = aligned(static_cast<size_t>(T) * sizeof(double));
    const size_t o_bytes = aligned(rows * T);
    const size_t r_bytes = aligned(rows);
    const int    status = ctx_reserve(ctx, w_bytes + b_bytes + o_bytes + 2 * r_bytes, 2 * r_bytes);
    if (status != NB200_OK)
    {
        return status;
    }
    auto* dW = reinterpret_cast<double*>(ctx->arena);
    auto* db = reinterpret_cast<double*>(ctx->arena + w_bytes);
    auto* dO = reinterpret_cast<double*>(ctx->arena + w_bytes + b_bytes);
    auto* dE = reinterpret_cast<double*>(ctx->arena + w_bytes + b_bytes + o_bytes);
    auto* dV = reinterpret_cast<double*>(ctx->arena + w_bytes + b_bytes + o_bytes + r_bytes);
    auto* hE = reinterpret_cast<double*>(ctx->staging);
    auto* hV = reinterpret_cast<double*>(ctx->staging + r_bytes);

    NB200_CUDA_TRY(cudaMemcpyAsync(dW, W, static_cast<size_t>(T) * D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NB200_CUDA_TRY(cudaMemcpyAsync(db, b, static_cast<size_t>(T) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    predict_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(data->X, dW, db, N, D, T, dO);
    NB200_CUDA_TRY(cudaGetLastError());
    const unsigned blocks = static_cast<unsigned>((N + 255) / 256);
    // pinball's error is its value (src/loss/pinball.cpp:20-23)
    const bool error_is_value = loss == NB200_LOSS_PINBALL;
    if (values != nullptr || error_is_value)
    {
        loss_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(loss, loss_param, data->Y, dO, N, T, dV, nullptr);
        NB200_CUDA_TRY(cudaGetLastError());
        NB200_CUDA_TRY(cudaMemcpyAsync(hV, dV, rows, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (errors != nullptr && !error_is_value)
    {
        loss_error_kernel<<<blocks, 256, 0, ctx->stream>>>(error_kind, data->Y, dO, N, T, dE);
        NB200_CUDA_TRY(cudaGetLastError());
        NB200_CUDA_TRY(cudaMemcpyAsync(hE, dE, rows, cudaMemcpyDeviceToHost, ctx->stream));
    }
    NB200_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (errors != nullptr)
    {
        std::memcpy(errors, error_is_value ? hV : hE, rows);
    }
    if (values != nullptr)
    {
        std::memcpy(values, hV, rows);
    }
    return NB200_OK;
}

int nb200_synth_linear_inputs(const int64_t samples, const int64_t inputs, const uint64_t seed, const int64_t modulo,
                              double* X, double* wopt, double* bopt)
{
    NB200_REQUIRE(samples > 0 && inputs > 0 && modulo > 0, "synthetic model: sizes must be positive");
    NB200_REQUIRE(X != nullptr && wopt != nullptr && bopt != nullptr, "synthetic model: null argument");
    std::mt19937_64                        rng(seed);
    std::uniform_real_distribution<double> udist(0.0, 1.0);
    for (int64_t k = 0; k < samples * inputs; ++k)
    {
        X[k] = udist(rng);
    }
    for (int64_t j = 0; j < inputs; ++j)
    {
        wopt[j] = udist(rng);
    }
    *bopt = udist(rng) - 0.5;

    double sum = 0.0;
    for (int64_t j = 0; j < inputs; ++j)
    {
        sum += wopt[j];
    }
    for (int64_t j = 0; j < inputs; ++j)
    {
        wopt[j] = (j % modulo != 0) ? 0.0 : wopt[j] / sum;
    }
    return NB200_OK;
}

} // extern "C"
