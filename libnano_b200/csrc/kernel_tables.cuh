// kernel_tables.cuh — the instantiated kernel variants, split over several translation units.
//
// Every group below is compiled as its own .cu (kernels_t1_<group>.cu, kernels_mt.cu): the groups build in parallel,
// and — the reason for the split — the code ptxas generates for one kernel family no longer depends on which other
// kernels happen to share its module (with everything in one TU under -split-compile, adding three single-target
// variants changed the SASS of mt_backward_kernel and cost config 3 15 %; profiles/README.md).
#pragma once

#include "vgrad_fd.cuh"
#include "vgrad_fl.cuh"
#include "vgrad_fr.cuh"
#include "vgrad_ft.cuh"
#include "vgrad_mt.cuh"
#include "vgrad_t1.cuh"

namespace nb200
{
using t1_kernel_t = void (*)(const t1_params);

struct t1_variant
{
    int         V, KC, WS, NCW, RW, ALT;
    t1_kernel_t grad;
    t1_kernel_t value;
};

using mt_fwd_kernel_t = void (*)(const mt_forward_params);
using mt_bwd_kernel_t = void (*)(const mt_backward_params);
using mt_tma_kernel_t = void (*)(const mt_forward_params, const CUtensorMap);

struct mt_variant
{
    int             NT8; // target tiles of 8: covers T <= 8 * NT8
    mt_fwd_kernel_t grad;
    mt_fwd_kernel_t value;
    mt_bwd_kernel_t backward;
    mt_tma_kernel_t grad_tma;  // forward with the X tile fed through a tensor map and the epilogue on its own warps
    mt_tma_kernel_t value_tma;
};

using ft_kernel_t = void (*)(const ft_params);

struct ft_variant
{
    int         T, KC, RW, NS; // targets (exact), 64 * KC columns per warp, rows per iteration, alternating warp sets
                               // (the 8 / NS warps of a set cover D <= 512 * KC / NS columns)
    ft_kernel_t grad;
    ft_kernel_t value;
};

using fd_kernel_t = void (*)(const fd_params);

struct fd_variant
{
    int         NT8, NS8, NSETS; // covers T <= 8 * NT8 and D <= (64 / NSETS) * NS8; NSETS < 0: the row-split kernel of
                                 // vgrad_fr.cuh (short rows, D <= 64 * NS8) with -NSETS tiles per warp and group;
                                 // NSETS >= 10: the kernel of vgrad_fl.cuh (D <= 64 * NS8) with NSETS - 10 tiles per
                                 // iteration
    int         header;   // bytes of the shared-memory header (fd_layout<NT8>::kHeader / fr_header(TPW) / ...)
    int         threads;  // threads per CTA
    fd_kernel_t grad;
    fd_kernel_t value;
};

// X(V, KC, WS, NCW, RW, ALT). The variant id is the position in A, B, C, D concatenated.
// Ordered so that the first variant of 0..16 whose column capacity 32*V*KC*WS covers D is the fall-back heuristic
// choice; the later entries are alternative decompositions.
#define NB200_T1_GROUP_A(X)                                                                                            \
    /* 0*/ X(2, 1, 1, 8, 8, 1) /* 1*/ X(2, 2, 1, 8, 8, 1) /* 2*/ X(2, 4, 1, 8, 4, 1) /* 3*/ X(2, 8, 1, 8, 2, 1)        \
    /* 4*/ X(2, 16, 1, 8, 1, 1) /* 5*/ X(2, 16, 2, 8, 1, 1) /* 6*/ X(2, 16, 4, 8, 1, 1) /* 7*/ X(2, 16, 8, 8, 1, 1)    \
    /* 8*/ X(1, 1, 1, 8, 8, 1) /* 9*/ X(1, 2, 1, 8, 8, 1)
#define NB200_T1_GROUP_B(X)                                                                                            \
    /*10*/ X(1, 4, 1, 8, 8, 1) /*11*/ X(1, 8, 1, 8, 4, 1) /*12*/ X(1, 16, 1, 8, 2, 1) /*13*/ X(1, 32, 1, 8, 1, 1)      \
    /*14*/ X(1, 32, 2, 8, 1, 1) /*15*/ X(1, 32, 4, 8, 1, 1) /*16*/ X(1, 32, 8, 8, 1, 1)                                \
    /* alternatives (same coverage, different decomposition) */                                                       \
    /*17*/ X(2, 8, 2, 8, 1, 1) /*18*/ X(2, 8, 2, 16, 1, 1) /*19*/ X(2, 4, 4, 16, 1, 1)
#define NB200_T1_GROUP_C(X)                                                                                            \
    /*20*/ X(2, 16, 1, 12, 1, 1) /*21*/ X(2, 16, 1, 8, 2, 1) /*22*/ X(2, 8, 1, 8, 1, 1) /*23*/ X(2, 8, 1, 8, 4, 1)     \
    /*24*/ X(2, 4, 1, 8, 1, 1) /*25*/ X(2, 4, 1, 8, 8, 1) /*26*/ X(2, 2, 1, 8, 1, 1)                                   \
    /* four consumer warps, twice the rows per warp and stage: 64 doubles of X per lane for every width */            \
    /*27*/ X(2, 16, 1, 4, 2, 1) /*28*/ X(2, 8, 1, 4, 4, 1) /*29*/ X(2, 4, 1, 4, 8, 1)
#define NB200_T1_GROUP_D(X)                                                                                            \
    /*30*/ X(2, 2, 1, 4, 16, 1) /*31*/ X(2, 1, 1, 4, 32, 1)                                                            \
    /* eight consumer warps in two sets that alternate over the CTA's tiles */                                        \
    /*32*/ X(2, 16, 1, 8, 2, 2) /*33*/ X(2, 8, 1, 8, 4, 2) /*34*/ X(2, 4, 1, 8, 8, 2) /*35*/ X(2, 2, 1, 8, 16, 2)      \
    /*36*/ X(2, 1, 1, 8, 32, 2)                                                                                        \
    /* rows split over two warps with several rows per iteration */                                                   \
    /*37*/ X(2, 8, 2, 8, 4, 2) /*38*/ X(2, 8, 2, 8, 2, 1) /*39*/ X(2, 16, 2, 8, 2, 2)

#define NB200_T1_ENTRY(V, KC, WS, NCW, RW, ALT)                                                                        \
    t1_variant{V, KC, WS, NCW, RW, ALT, vgrad_t1_kernel<V, KC, WS, NCW, RW, ALT, true>,                                \
               vgrad_t1_kernel<V, KC, WS, NCW, RW, ALT, false>},

// defined in kernels_t1_{a,b,c,d}.cu / kernels_mt.cu / kernels_ft.cu; return the group's entries and their count
const t1_variant* t1_group_a(int* count);
const t1_variant* t1_group_b(int* count);
const t1_variant* t1_group_c(int* count);
const t1_variant* t1_group_d(int* count);
const mt_variant* mt_variants(int* count);
const ft_variant* ft_variants(int* count);
const fd_variant* fd_variants(int* count);
const fd_variant* fl_variants(int* count);
} // namespace nb200
