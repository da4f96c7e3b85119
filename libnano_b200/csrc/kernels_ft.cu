// kernels_ft.cu — the few-target one-pass kernel variants (vgrad_ft.cuh, kernel_tables.cuh), compiled on their own.
#include "kernel_tables.cuh"

namespace nb200
{
namespace
{
// RW: rows per iteration, as many as keep RW * T <= 32 (one lane per (row, target) pair) and the row slices within
// 32 doubles per lane. Register budget per lane: 2 * T * 2 KC doubles of W + gradient, RW * 2 KC doubles of X.
#define NB200_FT_ENTRY(T, KC, RW) ft_variant{T, KC, RW, vgrad_ft_kernel<T, KC, RW, true>, vgrad_ft_kernel<T, KC, RW, false>}
const ft_variant kVariants[] = {
    // D <= 512
    NB200_FT_ENTRY(2, 1, 16), NB200_FT_ENTRY(3, 1, 10), NB200_FT_ENTRY(4, 1, 8), NB200_FT_ENTRY(5, 1, 6),
    NB200_FT_ENTRY(6, 1, 5),  NB200_FT_ENTRY(8, 1, 4),  NB200_FT_ENTRY(10, 1, 3),
    // D <= 1024
    NB200_FT_ENTRY(2, 2, 8),  NB200_FT_ENTRY(3, 2, 8),  NB200_FT_ENTRY(4, 2, 8),  NB200_FT_ENTRY(5, 2, 6),
    NB200_FT_ENTRY(6, 2, 5),  NB200_FT_ENTRY(8, 2, 4),  NB200_FT_ENTRY(10, 2, 3),
    // D <= 2048
    NB200_FT_ENTRY(2, 4, 4),  NB200_FT_ENTRY(3, 4, 4),  NB200_FT_ENTRY(4, 4, 4),
};
} // namespace

const ft_variant* ft_variants(int* count)
{
    *count = static_cast<int>(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}
} // namespace nb200
