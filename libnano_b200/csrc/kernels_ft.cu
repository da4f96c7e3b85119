// kernels_ft.cu — the few-target one-pass kernel variants (vgrad_ft.cuh, kernel_tables.cuh), compiled on their own.
#include "kernel_tables.cuh"

namespace nb200
{
namespace
{
// RW: rows per iteration, as many as keep RW * T <= 32 (one lane per (row, target) pair) and the row slices within
// 32 doubles per lane. Register budget per lane: 2 * T * 2 KC doubles of W + gradient, RW * 2 KC doubles of X.
#define NB200_FT_ENTRY(T, KC, RW, NS)                                                                                  \
    ft_variant{T, KC, RW, NS, vgrad_ft_kernel<T, KC, RW, NS, true>, vgrad_ft_kernel<T, KC, RW, NS, false>}
// the first entry with the objective's T that covers D is taken. Measured (profiles/r01_few_targets_sets.jsonl):
// two alternating sets of four warps win for D <= 512 and T >= 3 (T = 4: 3995 vs 3082 GB/s, T = 6: 2756 vs 1942);
// T = 2 (16 rows per iteration already) and 512 < D <= 1024 (the sets would need 4 rows per iteration) are faster
// with one set
const ft_variant kVariants[] = {
    // two sets, D <= 512
    NB200_FT_ENTRY(3, 2, 8, 2), NB200_FT_ENTRY(4, 2, 8, 2), NB200_FT_ENTRY(5, 2, 6, 2), NB200_FT_ENTRY(6, 2, 5, 2),
    NB200_FT_ENTRY(8, 2, 4, 2), NB200_FT_ENTRY(10, 2, 3, 2),
    // one set, D <= 512
    NB200_FT_ENTRY(2, 1, 16, 1), NB200_FT_ENTRY(3, 1, 10, 1), NB200_FT_ENTRY(4, 1, 8, 1), NB200_FT_ENTRY(5, 1, 6, 1),
    NB200_FT_ENTRY(6, 1, 5, 1),  NB200_FT_ENTRY(8, 1, 4, 1),  NB200_FT_ENTRY(10, 1, 3, 1),
    // one set, D <= 1024
    NB200_FT_ENTRY(2, 2, 8, 1),  NB200_FT_ENTRY(3, 2, 8, 1),  NB200_FT_ENTRY(4, 2, 8, 1),  NB200_FT_ENTRY(5, 2, 6, 1),
    NB200_FT_ENTRY(6, 2, 5, 1),  NB200_FT_ENTRY(8, 2, 4, 1),  NB200_FT_ENTRY(10, 2, 3, 1),
    // one set, D <= 2048
    NB200_FT_ENTRY(2, 4, 4, 1),  NB200_FT_ENTRY(3, 4, 4, 1),  NB200_FT_ENTRY(4, 4, 4, 1),
};
} // namespace

const ft_variant* ft_variants(int* count)
{
    *count = static_cast<int>(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}
} // namespace nb200
