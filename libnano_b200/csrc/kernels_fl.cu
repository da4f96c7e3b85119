// kernels_fl.cu — the one-pass tensor-pipe kernels with the loss spread over the lanes (vgrad_fl.cuh), compiled on their own.
#include "kernel_tables.cuh"

namespace nb200
{
namespace
{
// NSETS = 10 + TPI marks the two-barrier kernel, 20 + TPI the pipelined one (kernel_tables.cuh); the first entry that
// covers (T, D) and finds room for its ring stages is taken: one target tile before two, per width the pipelined kernel
// before the unpipelined one, two tiles per iteration before one
#define NB200_FL_ENTRY(NT8, NS8, TPI)                                                                                  \
    fd_variant{NT8, NS8, 10 + (TPI), fl_header(NT8, TPI), kFdThreads, vgrad_fl_kernel<NT8, NS8, TPI, true>,                        \
               vgrad_fl_kernel<NT8, NS8, TPI, false>}
#define NB200_FP_ENTRY(NT8, NS8)                                                                                       \
    fd_variant{NT8, NS8, 21, fp_header(NT8), fp_threads(NT8), vgrad_fp_kernel<NT8, NS8, true>, vgrad_fp_kernel<NT8, NS8, false>}
#define NB200_FL_GROUP(NT8, NS8) NB200_FP_ENTRY(NT8, NS8), NB200_FL_ENTRY(NT8, NS8, 2), NB200_FL_ENTRY(NT8, NS8, 1)
const fd_variant kVariants[] = {
    NB200_FL_GROUP(1, 5), NB200_FL_GROUP(1, 6), NB200_FL_GROUP(1, 8), NB200_FL_GROUP(1, 10), NB200_FL_GROUP(1, 13),
    NB200_FL_GROUP(1, 16),
    NB200_FL_GROUP(2, 5), NB200_FL_GROUP(2, 6), NB200_FL_GROUP(2, 8), NB200_FL_GROUP(2, 10), NB200_FL_GROUP(2, 13),
    NB200_FL_GROUP(2, 16),
};
} // namespace

const fd_variant* fl_variants(int* count)
{
    *count = static_cast<int>(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}
} // namespace nb200
