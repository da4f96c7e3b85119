// kernels_fd.cu — the one-pass tensor-pipe kernel variants for a handful of targets (vgrad_fd.cuh), compiled on their own.
#include "kernel_tables.cuh"

namespace nb200
{
namespace
{
#define NB200_FD_ENTRY(NT8, NS8, NSETS)                                                                                \
    fd_variant{NT8, NS8, NSETS, fd_layout<NT8>::kHeader, kFdThreads, vgrad_fd_kernel<NT8, NS8, NSETS, true>,                       \
               vgrad_fd_kernel<NT8, NS8, NSETS, false>}
#define NB200_FR_ENTRY(NT8, NS8, TPW)                                                                                  \
    fd_variant{NT8, NS8, -(TPW), fr_header(TPW), kFdThreads, vgrad_fr_kernel<NT8, NS8, TPW, true>,                                 \
               vgrad_fr_kernel<NT8, NS8, TPW, false>}
// the first entry that covers (T, D) is taken: short rows go to the row-split kernel (vgrad_fr.cuh) where the caller
// allows it; for T <= 8 two alternating sets of four warps (each warp twice the columns)
const fd_variant kVariants[] = {
    NB200_FR_ENTRY(1, 1, 2),  NB200_FR_ENTRY(1, 2, 2),  NB200_FR_ENTRY(1, 4, 1),
    NB200_FR_ENTRY(2, 1, 2),  NB200_FR_ENTRY(2, 2, 2),  NB200_FR_ENTRY(2, 4, 1),
    NB200_FD_ENTRY(1, 8, 2),  NB200_FD_ENTRY(1, 16, 2), NB200_FD_ENTRY(1, 24, 2), NB200_FD_ENTRY(1, 28, 2),
    NB200_FD_ENTRY(1, 32, 2),
    NB200_FD_ENTRY(1, 2, 1),  NB200_FD_ENTRY(1, 4, 1),  NB200_FD_ENTRY(1, 6, 1),  NB200_FD_ENTRY(1, 8, 1),
    NB200_FD_ENTRY(1, 10, 1), NB200_FD_ENTRY(1, 14, 1), NB200_FD_ENTRY(1, 16, 1),
    NB200_FD_ENTRY(2, 2, 1),  NB200_FD_ENTRY(2, 4, 1),  NB200_FD_ENTRY(2, 6, 1),  NB200_FD_ENTRY(2, 8, 1),
    NB200_FD_ENTRY(2, 10, 1), NB200_FD_ENTRY(2, 13, 1), NB200_FD_ENTRY(2, 16, 1),
};
} // namespace

const fd_variant* fd_variants(int* count)
{
    *count = static_cast<int>(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}
} // namespace nb200
