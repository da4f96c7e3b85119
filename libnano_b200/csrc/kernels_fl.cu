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
    fd_variant{NT8, NS8, 10 + (TPI), fl_header(NT8, TPI), vgrad_fl_kernel<NT8, NS8, TPI, true>,                        \
               vgrad_fl_kernel<NT8, NS8, TPI, false>}
#define NB200_FP_ENTRY(NT8, NS8, TPI)                                                                                  \
    fd_variant{NT8, NS8, 20 + (TPI), fp_header(NT8, TPI), vgrad_fp_kernel<NT8, NS8, TPI, true>,                        \
               vgrad_fp_kernel<NT8, NS8, TPI, false>}
#define NB200_FL_WIDE(NT8, NS8) NB200_FP_ENTRY(NT8, NS8, 1), NB200_FL_ENTRY(NT8, NS8, 2), NB200_FL_ENTRY(NT8, NS8, 1)
#define NB200_FL_NARROW(NT8, NS8) NB200_FP_ENTRY(NT8, NS8, 2), NB200_FL_WIDE(NT8, NS8)
const fd_variant kVariants[] = {
    NB200_FL_NARROW(1, 5), NB200_FL_NARROW(1, 6), NB200_FL_NARROW(1, 8), NB200_FL_WIDE(1, 10), NB200_FL_WIDE(1, 13),
    NB200_FL_WIDE(1, 16),
    NB200_FL_NARROW(2, 5), NB200_FL_NARROW(2, 6), NB200_FL_NARROW(2, 8), NB200_FL_WIDE(2, 10), NB200_FL_WIDE(2, 13),
    NB200_FL_WIDE(2, 16),
};
} // namespace

const fd_variant* fl_variants(int* count)
{
    *count = static_cast<int>(sizeof(kVariants) / sizeof(kVariants[0]));
    return kVariants;
}
} // namespace nb200
