// kernels_mt.cu — the FP64 tensor-pipe (DMMA) kernel variants (kernel_tables.cuh), compiled on their own.
#include "kernel_tables.cuh"

namespace nb200
{
namespace
{
#define NB200_MT_ENTRY(NT8)                                                                                            \
    mt_variant{NT8, mt_forward_kernel<NT8, true>, mt_forward_kernel<NT8, false>, mt_backward_kernel<NT8>,               \
               mt_forward_tma_kernel<NT8, true>, mt_forward_tma_kernel<NT8, false>}
const mt_variant kVariants[] = {NB200_MT_ENTRY(1), NB200_MT_ENTRY(2),  NB200_MT_ENTRY(3),
                                NB200_MT_ENTRY(4), NB200_MT_ENTRY(6),  NB200_MT_ENTRY(8),
                                NB200_MT_ENTRY(10), NB200_MT_ENTRY(13), NB200_MT_ENTRY(16)};
} // namespace

const mt_variant* mt_variants(int* count)
{
    *count = static_cast<int>(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}
} // namespace nb200
