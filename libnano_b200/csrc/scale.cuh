// scale.cuh — "next row" N1 of SURVEY.md §8f: what flatten_iterator_t does to the data before the objective sees it,
// done once in HBM at pin time instead of per batch on the host:
//   column statistics over the finite values   scalar_stats_t update/done   src/dataset/stats.cpp:81-137
//   scaling none / mean / minmax / standard     scalar_stats_t::scale        src/dataset/stats.cpp:357-401
//   non-finite -> 0 (missing values)            nan2zero                     src/dataset/stats.cpp:8-22
// and nano::upscale (src/dataset/stats.cpp:211-230), which maps the fitted (W, b) back to the original units.
#pragma once

#include "common.cuh"

namespace nb200
{
constexpr int kStatsRowsPerBlock = 4096;

#ifdef __CUDACC__

// partial[chunk][5][C]: count, sum, sum of squares, min, max over the finite values of the chunk's rows.
// One thread per column (coalesced across a row), rows in order => fixed summation order.
__global__ void __launch_bounds__(256) column_stats_kernel(const double* __restrict__ values, const int64_t N,
                                                           const int64_t C, double* __restrict__ partial)
{
    const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= C)
    {
        return;
    }
    const int64_t begin = static_cast<int64_t>(blockIdx.y) * kStatsRowsPerBlock;
    const int64_t end   = min(N, begin + kStatsRowsPerBlock);
    double        count = 0.0, sum = 0.0, sumsq = 0.0, lo = INFINITY, hi = -INFINITY;
    for (int64_t i = begin; i < end; ++i)
    {
        const double v = values[i * C + c];
        if (isfinite(v))
        {
            count += 1.0;
            sum += v;
            sumsq = fma(v, v, sumsq);
            lo    = fmin(lo, v);
            hi    = fmax(hi, v);
        }
    }
    double* out = partial + static_cast<int64_t>(blockIdx.y) * 5 * C;
    out[0 * C + c] = count;
    out[1 * C + c] = sum;
    out[2 * C + c] = sumsq;
    out[3 * C + c] = lo;
    out[4 * C + c] = hi;
}

// fold the chunks in order: out[5][C]
__global__ void __launch_bounds__(256) column_stats_fold_kernel(const double* __restrict__ partial, const int chunks,
                                                                const int64_t C, double* __restrict__ out)
{
    const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= C)
    {
        return;
    }
    double count = 0.0, sum = 0.0, sumsq = 0.0, lo = INFINITY, hi = -INFINITY;
    for (int k = 0; k < chunks; ++k)
    {
        const double* p = partial + static_cast<int64_t>(k) * 5 * C;
        count += p[0 * C + c];
        sum += p[1 * C + c];
        sumsq += p[2 * C + c];
        lo = fmin(lo, p[3 * C + c]);
        hi = fmax(hi, p[4 * C + c]);
    }
    out[0 * C + c] = count;
    out[1 * C + c] = sum;
    out[2 * C + c] = sumsq;
    out[3 * C + c] = lo;
    out[4 * C + c] = hi;
}

// values[i, c] = (values[i, c] - shift[c]) * mul[c], non-finite -> 0  (stats.cpp:357-401)
__global__ void __launch_bounds__(256) column_scale_kernel(double* __restrict__ values, const int64_t N,
                                                           const int64_t C, const double* __restrict__ shift,
                                                           const double* __restrict__ mul)
{
    const int64_t total = N * C;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < total;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x)
    {
        const int64_t c = k % C;
        const double  v = __dmul_rn(__dsub_rn(values[k], shift[c]), mul[c]);
        values[k]       = isfinite(v) ? v : 0.0;
    }
}

#endif // __CUDACC__

// scalar_stats_t for one group of columns (host side, after `done`)
struct column_stats_t
{
    std::vector<double> samples, min, max, mean, stdev, div_range, div_stdev;
    int                 scaling{0}; // the scaling applied to the device data (0 = none)

    bool empty() const { return min.empty(); }
};

// done() of src/dataset/stats.cpp:100-137 from the raw sums
inline void finish_column_stats(const std::vector<double>& raw /* [5][C] */, const int64_t C,
                                const unsigned char* enable, column_stats_t& stats)
{
    const double epsilon = 1e-8; // epsilon2 (include/nano/core/numeric.h:107-111)
    stats.samples.assign(C, 0.0);
    stats.min.assign(C, 0.0);
    stats.max.assign(C, 0.0);
    stats.mean.assign(C, 0.0);
    stats.stdev.assign(C, 0.0);
    stats.div_range.assign(C, 1.0);
    stats.div_stdev.assign(C, 1.0);
    for (int64_t i = 0; i < C; ++i)
    {
        const double N = raw[0 * C + i], sum = raw[1 * C + i], sumsq = raw[2 * C + i];
        stats.samples[i] = N;
        if (N > 1.0)
        {
            stats.min[i]       = raw[3 * C + i];
            stats.max[i]       = raw[4 * C + i];
            stats.stdev[i]     = std::sqrt((sumsq - sum * sum / N) / (N - 1.0));
            stats.mean[i]      = sum / N;
            stats.div_range[i] = 1.0 / std::max(stats.max[i] - stats.min[i], epsilon);
            stats.div_stdev[i] = 1.0 / std::max(stats.stdev[i], epsilon);
        }
        else if (N == 1.0)
        {
            stats.min[i] = stats.max[i] = raw[3 * C + i];
            stats.mean[i]               = sum; // the reference leaves the un-divided sum here for N == 1
        }
        if (enable != nullptr && enable[i] == 0) // scaling disabled for this column (categorical)
        {
            stats.min[i] = stats.max[i] = stats.mean[i] = stats.stdev[i] = 0.0;
            stats.div_range[i] = stats.div_stdev[i] = 1.0;
        }
    }
}

// make_scaling (stats.cpp:46-78): value -> value * w + b
inline void scaling_affine(const column_stats_t& stats, const int scaling, std::vector<double>& w,
                           std::vector<double>& b)
{
    const size_t C = stats.min.size();
    w.assign(C, 1.0);
    b.assign(C, 0.0);
    for (size_t i = 0; i < C; ++i)
    {
        switch (scaling)
        {
        case 1: w[i] = stats.div_range[i]; b[i] = -stats.mean[i] * stats.div_range[i]; break;
        case 2: w[i] = stats.div_range[i]; b[i] = -stats.min[i] * stats.div_range[i]; break;
        case 3: w[i] = stats.div_stdev[i]; b[i] = -stats.mean[i] * stats.div_stdev[i]; break;
        default: break;
        }
    }
}
} // namespace nb200
