// kernels_fd.cu — the one-pass tensor-pipe kernel variants for a handful of targets (vgrad_fd.cuh), compiled on their own.
#include "kernel_tables.cuh"

namespace nb200
{
namespace
{
#define NB200_FD_ENTRY(NT8, NS8)                                                                                       \
    fd_variant{NT8, NS8, fd_layout<NT8>::kHeader, vgrad_fd_kernel<NT8, NS8, true>, vgrad_fd_kernel<NT8, NS8, false>}
// ordered by NT8, then NS8: the first entry that covers (T, D) is taken
const fd_variant kVariants[] = {
    NB200_FD_ENTRY(1, 2), NB200_FD_ENTRY(1, 4), NB200_FD_ENTRY(1, 6), NB200_FD_ENTRY(1, 8), NB200_FD_ENTRY(1, 10),
    NB200_FD_ENTRY(1, 14), NB200_FD_ENTRY(1, 16), NB200_FD_ENTRY(2, 2), NB200_FD_ENTRY(2, 4), NB200_FD_ENTRY(2, 6),
    NB200_FD_ENTRY(2, 8), NB200_FD_ENTRY(2, 10), NB200_FD_ENTRY(2, 13), NB200_FD_ENTRY(2, 16),
};
} // namespace

const fd_variant* fd_variants(int* count)
{
    *count = static_cast<int>(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}
} // namespace nb200
