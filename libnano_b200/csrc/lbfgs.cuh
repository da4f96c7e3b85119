// lbfgs.cuh — device-resident L-BFGS: libnano's solver_lbfgs_t (src/solver/lbfgs.cpp:25-124) with its default line
// search (lsearch0 quadratic: src/lsearch0/quadratic.cpp:16-33; lsearchk CG-DESCENT: src/lsearchk/cgdescent.cpp,
// glue src/lsearchk.cpp:36-91, src/solver/lsearch.cpp:11-26, step formulas src/solver/lstep.cpp:25-88, predicates
// src/solver/state.cpp:165-210, stopping src/solver.cpp:114-173) driving the GPU objective WITHOUT moving vectors:
//
//   * x, g, the previous point, the descent direction, the (s, y) history ring and the line-search trial points
//     live in HBM; the two-loop recursion (lbfgs.cpp:49-84) is one single-CTA kernel with fixed-order block sums;
//   * per evaluation the host sees four scalars (f, g.d, |g|_inf, all-finite) written to mapped pinned memory by a
//     probe kernel — the control flow (bracketing, secant steps, Wolfe tests) is scalar and stays on the host.
//
// BASELINE.json north_star item (3): "the solver's small vector ops (dot, axpy, two-loop L-BFGS recursion) kept
// device-resident". The reference solvers still drive the objective unchanged through nb200_vgrad; this driver removes
// the 2 x 8 KB PCIe round trip and the host BLAS-1 work per evaluation for callers that can use it.
//
// Included by nb200_api.cu after enqueue_eval / enqueue_finish are defined.
#pragma once

namespace nb200
{
constexpr int kSolverThreads = 1024;

#ifdef __CUDACC__

// fixed-order sum of one value per thread over the whole CTA; every thread gets the result
__device__ __forceinline__ double cta_sum(double v, double* scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_allreduce_sum(v);
    if (lane == 0)
    {
        scratch[warp] = v;
    }
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < kSolverThreads / 32; ++w)
    {
        s += scratch[w];
    }
    __syncthreads();
    return s;
}

__device__ __forceinline__ double cta_max(double v, double* scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_allreduce_max(v);
    if (lane == 0)
    {
        scratch[warp] = v;
    }
    __syncthreads();
    double s = scratch[0];
    for (int w = 1; w < kSolverThreads / 32; ++w)
    {
        s = (s < scratch[w]) ? scratch[w] : s;
    }
    __syncthreads();
    return s;
}

// d = -H g with the L-BFGS two-loop recursion (Nocedal & Wright p.178 as written in lbfgs.cpp:49-84).
// History ring: logical pair j (0 = oldest) sits in slot (head + j) % cap of S / Y ([cap, n] each).
__global__ void __launch_bounds__(kSolverThreads) lbfgs_two_loop_kernel(const int64_t n, const double* __restrict__ g,
                                                                        const double* __restrict__ S,
                                                                        const double* __restrict__ Y, const int count,
                                                                        const int head, const int cap,
                                                                        double* __restrict__ q, double* __restrict__ alphas,
                                                                        double* __restrict__ d)
{
    __shared__ double scratch[kSolverThreads / 32];
    for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
    {
        q[i] = g[i];
    }
    for (int j = 0; j < count; ++j) // newest to oldest
    {
        const int     slot = (head + count - 1 - j) % cap;
        const double* s    = S + static_cast<int64_t>(slot) * n;
        const double* y    = Y + static_cast<int64_t>(slot) * n;
        double        sq = 0.0, sy = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
        {
            sq = fma(s[i], q[i], sq);
            sy = fma(s[i], y[i], sy);
        }
        sq = cta_sum(sq, scratch);
        sy = cta_sum(sy, scratch);
        const double alpha = sq / sy;
        for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
        {
            q[i] -= alpha * y[i];
        }
        if (threadIdx.x == 0)
        {
            alphas[j] = alpha;
        }
    }
    double scale = 1.0;
    if (count > 0)
    {
        const int     slot = (head + count - 1) % cap;
        const double* s    = S + static_cast<int64_t>(slot) * n;
        const double* y    = Y + static_cast<int64_t>(slot) * n;
        double        sy = 0.0, yy = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
        {
            sy = fma(s[i], y[i], sy);
            yy = fma(y[i], y[i], yy);
        }
        sy    = cta_sum(sy, scratch);
        yy    = cta_sum(yy, scratch);
        scale = sy / yy;
    }
    for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
    {
        q[i] = scale * q[i]; // r (kept in q)
    }
    __syncthreads(); // alphas written by thread 0 are read below
    for (int j = 0; j < count; ++j) // oldest to newest
    {
        const int     slot = (head + j) % cap;
        const double* s    = S + static_cast<int64_t>(slot) * n;
        const double* y    = Y + static_cast<int64_t>(slot) * n;
        double        yr = 0.0, sy = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
        {
            yr = fma(y[i], q[i], yr);
            sy = fma(s[i], y[i], sy);
        }
        yr = cta_sum(yr, scratch);
        sy = cta_sum(sy, scratch);
        const double coeff = alphas[count - 1 - j] - yr / sy;
        for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
        {
            q[i] += s[i] * coeff;
        }
    }
    for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
    {
        d[i] = -q[i];
    }
}

// out[0] = f, out[1] = g.d, out[2] = |g|_inf, out[3] = 1 if f, x and g are all finite (solver_state_t::valid)
__global__ void __launch_bounds__(kSolverThreads) solver_probe_kernel(const int64_t n, const double* __restrict__ x,
                                                                      const double* __restrict__ g,
                                                                      const double* __restrict__ d,
                                                                      const double* __restrict__ f, double* out)
{
    __shared__ double scratch[kSolverThreads / 32];
    double            dg = 0.0, gmax = 0.0, bad = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
    {
        const double gi = g[i];
        dg              = fma(gi, d[i], dg);
        gmax            = fmax(gmax, fabs(gi));
        bad += (isfinite(gi) && isfinite(x[i])) ? 0.0 : 1.0;
    }
    dg   = cta_sum(dg, scratch);
    gmax = cta_max(gmax, scratch);
    bad  = cta_sum(bad, scratch);
    if (threadIdx.x == 0)
    {
        out[0] = *f;
        out[1] = dg;
        out[2] = gmax;
        out[3] = (bad == 0.0 && isfinite(*f)) ? 1.0 : 0.0;
        __threadfence_system();
    }
}

__global__ void solver_axpy_kernel(const int64_t n, const double* __restrict__ x0, const double t,
                                   const double* __restrict__ d, double* __restrict__ out)
{
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    {
        out[i] = fma(t, d[i], x0[i]); // x0 + t * d
    }
}

__global__ void solver_negate_kernel(const int64_t n, const double* __restrict__ g, double* __restrict__ d)
{
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    {
        d[i] = -g[i];
    }
}

// CGD direction (src/solver/cgd.cpp:104-127): d = -cg + beta * pd with one of the ten beta formulas of cgd.cpp:7-71,
// restarted to -cg when it is not a descent direction or consecutive gradients are far from orthogonal. One CTA,
// fixed-order sums; `d` holds the previous direction on entry and the new one on exit.
enum cgd_beta : int
{
    CGD_HS = 0, CGD_FR = 1, CGD_PR = 2, CGD_CD = 3, CGD_LS = 4, CGD_DY = 5, CGD_N = 6, CGD_DYCD = 7, CGD_DYHS = 8,
    CGD_FRPR = 9, CGD_COUNT = 10
};

__global__ void __launch_bounds__(kSolverThreads) cgd_direction_kernel(const int64_t n, const double* __restrict__ cg,
                                                                       const double* __restrict__ pg, double* d,
                                                                       const int kind, const double eta0,
                                                                       const double orthotest)
{
    __shared__ double scratch[kSolverThreads / 32];
    double cg_y = 0.0, pd_y = 0.0, y_y = 0.0, cg_cg = 0.0, pg_pg = 0.0, pd_pg = 0.0, pd_pd = 0.0, cg_pg = 0.0, pd_cg = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
    {
        const double c = cg[i], p = pg[i], q = d[i], y = c - p;
        cg_y  = fma(c, y, cg_y);
        pd_y  = fma(q, y, pd_y);
        y_y   = fma(y, y, y_y);
        cg_cg = fma(c, c, cg_cg);
        pg_pg = fma(p, p, pg_pg);
        pd_pg = fma(q, p, pd_pg);
        pd_pd = fma(q, q, pd_pd);
        cg_pg = fma(c, p, cg_pg);
        pd_cg = fma(q, c, pd_cg);
    }
    cg_y  = cta_sum(cg_y, scratch);
    pd_y  = cta_sum(pd_y, scratch);
    y_y   = cta_sum(y_y, scratch);
    cg_cg = cta_sum(cg_cg, scratch);
    pg_pg = cta_sum(pg_pg, scratch);
    pd_pg = cta_sum(pd_pg, scratch);
    pd_pd = cta_sum(pd_pd, scratch);
    cg_pg = cta_sum(cg_pg, scratch);
    pd_cg = cta_sum(pd_cg, scratch);

    const double HS = cg_y / pd_y, FR = cg_cg / pg_pg, PR = cg_y / pg_pg, CD = -cg_cg / pd_pg, LS = -cg_y / pd_pg,
                 DY = cg_cg / pd_y;
    double beta;
    switch (kind)
    {
    case CGD_HS: beta = fmax(HS, 0.0); break;
    case CGD_FR: beta = FR; break;
    case CGD_PR: beta = fmax(PR, 0.0); break;
    case CGD_CD: beta = CD; break;
    case CGD_LS: beta = fmax(LS, 0.0); break;
    case CGD_DY: beta = DY; break;
    case CGD_N:
    {
        // max(eta, div * (y - 2 pd |y|^2 div) . cg), the dot product expanded by linearity
        const double div = 1.0 / pd_y;
        const double eta = -1.0 / (sqrt(pd_pd) * fmin(eta0, sqrt(pg_pg)));
        beta             = fmax(eta, div * (cg_y - 2.0 * y_y * div * pd_cg));
        break;
    }
    case CGD_DYCD: beta = cg_cg / fmax(pd_y, -pd_pg); break;
    case CGD_DYHS: beta = fmax(0.0, fmin(DY, HS)); break;
    default: beta = (PR < -FR) ? -FR : ((fabs(PR) <= FR) ? PR : FR); break;
    }

    double dg = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
    {
        const double c = cg[i];
        const double v = fma(beta, d[i], -c);
        d[i]           = v;
        dg             = fma(c, v, dg);
    }
    dg = cta_sum(dg, scratch);
    if (!(dg < 0.0) || fabs(cg_pg) >= orthotest * cg_cg)
    {
        for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
        {
            d[i] = -cg[i];
        }
    }
}

// s = x - x_prev, y = g - g_prev into one history slot (lbfgs.cpp:107-108)
__global__ void lbfgs_history_kernel(const int64_t n, const double* __restrict__ x, const double* __restrict__ xp,
                                     const double* __restrict__ g, const double* __restrict__ gp,
                                     double* __restrict__ s, double* __restrict__ y)
{
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    {
        s[i] = x[i] - xp[i];
        y[i] = g[i] - gp[i];
    }
}

#endif // __CUDACC__

// ---------------------------------------------------------------------------------------------------------------
// host side: scalar view of a solver state and the CG-DESCENT control flow on it
// ---------------------------------------------------------------------------------------------------------------
struct solver_point_t
{
    double f{0.0};     // function value
    double dg{0.0};    // gradient . descent direction
    double gmax{0.0};  // |gradient|_inf
    bool   valid{false};
};

struct lsearch_step_t // src/solver/lstep.cpp
{
    double t, f, g;
};

inline double lsearch_stpmin()
{
    return 10.0 * 2.220446049250313e-16;
}

inline double lsearch_secant(const lsearch_step_t& u, const lsearch_step_t& v)
{
    const double t = (v.t * u.g - u.t * v.g) / (u.g - v.g);
    const double lo = lsearch_stpmin(), hi = 1.0 / lsearch_stpmin();
    return (t < lo) ? lo : ((hi < t) ? hi : t); // std::clamp semantics (NaN falls through)
}

// The line-search solvers that share this skeleton: L-BFGS (lbfgs.cpp) and the CGD family (cgd.cpp) differ in how the
// descent direction is built and in the curvature tolerance c2 of the line search (0.9 resp. 0.1).
struct solver_method_t
{
    bool   cgd{false};
    int    beta{CGD_PR};
    double c2{0.9};
    double orthotest{0.1}; // solver::cgd::orthotest
    double eta{0.01};      // solver::cgdN::eta
};

class device_lbfgs_t
{
public:
    // obj: the objective whose stream carries the solver state (the first part of a device group); group: the group
    // objective when the evaluation spans several devices of this process (group.cuh), null otherwise
    device_lbfgs_t(nb200_obj* obj, const int64_t n_global, nb200_allreduce_fn allreduce, void* allreduce_user,
                   const solver_method_t method = solver_method_t{}, nb200_obj* group = nullptr)
        : m_obj(obj)
        , m_group(group)
        , m_n(obj->n)
        , m_n_global(n_global)
        , m_allreduce(allreduce)
        , m_allreduce_user(allreduce_user)
        , m_method(method)
    {
    }

    ~device_lbfgs_t()
    {
        cudaFree(m_block);
        cudaFreeHost(m_probe);
    }

    int run(double* x_host, const double epsilon, const int64_t max_evals, const int64_t history, double* fx,
            double* gradient_test, int64_t* stats, int* status_out)
    {
        const size_t n = static_cast<size_t>(m_n);
        m_cap          = static_cast<int>(history);
        // one allocation: x, g, x0, g0, d, q, xt(unused slot kept for alignment), f scalar, alphas, S, Y
        const size_t doubles = 7 * (n + 2) + 8 + static_cast<size_t>(m_cap) + 2 * static_cast<size_t>(m_cap) * n;
        NB200_CUDA_TRY(cudaMalloc(&m_block, doubles * sizeof(double)));
        m_x      = m_block;
        m_g      = m_x + (n + 2);
        m_x0     = m_g + (n + 2);
        m_g0     = m_x0 + (n + 2);
        m_d      = m_g0 + (n + 2);
        m_q      = m_d + (n + 2);
        m_f      = m_q + 2 * (n + 2);
        m_alphas = m_f + 8;
        m_S      = m_alphas + m_cap;
        m_Y      = m_S + static_cast<size_t>(m_cap) * n;
        NB200_CUDA_TRY(cudaHostAlloc(&m_probe, 8 * sizeof(double), cudaHostAllocMapped));
        NB200_CUDA_TRY(cudaHostGetDevicePointer(&m_probe_dev, m_probe, 0));

        cudaStream_t stream = m_obj->stream;
        NB200_CUDA_TRY(cudaMemcpyAsync(m_x, x_host, n * sizeof(double), cudaMemcpyHostToDevice, stream));
        NB200_CUDA_TRY(cudaMemsetAsync(m_d, 0, n * sizeof(double), stream));

        int        status     = 0; // 0 max_iters, 1 gradient_test, 4 failed (same codes as the restated CPU driver)
        int64_t    iterations = 0;
        solver_point_t c;          // current state
        int            rc = evaluate(c);
        if (rc != NB200_OK)
        {
            return rc;
        }
        solver_point_t p = c; // previous state
        bool           current_is_valid = c.valid;

        if (c.gmax < 1e-15) // lbfgs.cpp:34-38 (epsilon0)
        {
            status = c.valid ? 1 : 4;
        }
        else
        {
            double last_step_size = -1.0, prevf = 0.0, prevdg = 1.0; // lsearch.h:30, quadratic.h:33-34
            int    count = 0, head = 0;
            while (m_fcalls + m_gcalls < max_evals)
            {
                // descent direction
                if (!m_method.cgd)
                {
                    lbfgs_two_loop_kernel<<<1, kSolverThreads, 0, stream>>>(m_n, m_g, m_S, m_Y, count, head, m_cap,
                                                                            m_q, m_alphas, m_d);
                }
                else if (iterations == 0)
                {
                    solver_negate_kernel<<<blocks(), 256, 0, stream>>>(m_n, m_g, m_d); // cgd.cpp:106-109
                }
                else
                {
                    // m_g0 / m_d still hold the previous gradient and direction (cgd.cpp:112-126, restart included)
                    cgd_direction_kernel<<<1, kSolverThreads, 0, stream>>>(m_n, m_g, m_g0, m_d, m_method.beta,
                                                                           m_method.eta, m_method.orthotest);
                }
                NB200_CUDA_TRY(cudaGetLastError());
                m_obj->launches += 1;
                rc = probe(c); // refresh c.dg for the new direction
                if (rc != NB200_OK)
                {
                    return rc;
                }
                const bool has_descent = m_method.cgd || c.dg < 0.0; // lbfgs.cpp:89-94
                if (!has_descent)
                {
                    solver_negate_kernel<<<blocks(), 256, 0, stream>>>(m_n, m_g, m_d);
                    NB200_CUDA_TRY(cudaGetLastError());
                    m_obj->launches += 1;
                    rc = probe(c);
                    if (rc != NB200_OK)
                    {
                        return rc;
                    }
                }

                // pstate = cstate ; lsearch0 quadratic (quadratic.cpp:16-33)
                NB200_CUDA_TRY(cudaMemcpyAsync(m_x0, m_x, n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
                NB200_CUDA_TRY(cudaMemcpyAsync(m_g0, m_g, n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
                p = c;
                const double t0 = (last_step_size < 0.0) ? 1.0 : std::min(1.0, 1.01 * 2.0 * (c.f - prevf) / prevdg);
                prevf           = c.f;
                prevdg          = c.dg;

                bool   iter_ok   = false;
                double step_size = t0;
                rc               = line_search(p, c, t0, iter_ok, step_size);
                if (rc != NB200_OK)
                {
                    return rc;
                }
                last_step_size = step_size;
                ++iterations;

                // done_gradient_test for a smooth function (src/solver.cpp:149-173,114-135)
                current_is_valid = c.valid;
                if (!(iter_ok && c.valid))
                {
                    status = 4;
                    break;
                }
                if (c.gmax / std::max(1.0, std::fabs(c.f)) < epsilon)
                {
                    status = 1;
                    break;
                }

                if (m_method.cgd)
                {
                    continue; // no curvature history: the next direction needs the previous gradient and direction only
                }
                if (has_descent) // lbfgs.cpp:105-120
                {
                    int slot;
                    if (count < m_cap)
                    {
                        slot = (head + count) % m_cap;
                        ++count;
                    }
                    else
                    {
                        slot = head;
                        head = (head + 1) % m_cap;
                    }
                    lbfgs_history_kernel<<<blocks(), 256, 0, stream>>>(m_n, m_x, m_x0, m_g, m_g0,
                                                                       m_S + static_cast<size_t>(slot) * n,
                                                                       m_Y + static_cast<size_t>(slot) * n);
                    NB200_CUDA_TRY(cudaGetLastError());
                    m_obj->launches += 1;
                }
                else
                {
                    count = 0;
                    head  = 0;
                }
            }
        }

        // choose(cstate, pstate) (src/solver.cpp:219-225)
        const double*         x_best = current_is_valid ? m_x : m_x0;
        const solver_point_t& best   = current_is_valid ? c : p;
        NB200_CUDA_TRY(cudaMemcpyAsync(x_host, x_best, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
        rc = obj_sync(m_obj);
        if (rc != NB200_OK)
        {
            return rc;
        }
        collect_timing(m_obj);
        *fx            = best.f;
        *gradient_test = best.gmax / std::max(1.0, std::fabs(best.f));
        stats[0]       = m_fcalls;
        stats[1]       = m_gcalls;
        stats[2]       = iterations;
        *status_out    = status;
        return NB200_OK;
    }

private:
    unsigned blocks() const { return static_cast<unsigned>(std::min<int64_t>((m_n + 255) / 256, 1184)); }

    // f and g at m_x (one fused launch; partial + all-reduce + finish when sharded), then the probe
    int evaluate(solver_point_t& point)
    {
        int status;
        if (m_allreduce == nullptr)
        {
            status = solver_enqueue_eval(m_obj, m_group, m_x, m_n_global, m_g, m_f);
        }
        else
        {
            status = enqueue_eval(m_obj, m_x, true, false, 0, nullptr, nullptr);
            if (status == NB200_OK)
            {
                const int64_t len = 1 + m_obj->data->T + m_obj->data->T * m_obj->data->D;
                if (m_allreduce(m_allreduce_user, m_obj->d_reduced, len, static_cast<void*>(m_obj->stream)) != 0)
                {
                    return fail(NB200_ERR_INVALID, "lbfgs: the all-reduce callback failed");
                }
                status = enqueue_finish(m_obj, m_x, m_n_global, m_g, m_f);
            }
        }
        if (status != NB200_OK)
        {
            return status;
        }
        m_fcalls += 1;
        m_gcalls += 1;
        return probe(point);
    }

    int probe(solver_point_t& point)
    {
        solver_probe_kernel<<<1, kSolverThreads, 0, m_obj->stream>>>(m_n, m_x, m_g, m_d, m_f, m_probe_dev);
        NB200_CUDA_TRY(cudaGetLastError());
        m_obj->launches += 1;
        int status = obj_sync(m_obj);
        if (status == NB200_OK)
        {
            // a fused evaluation that gave up on a peer summed stale mail: stop instead of iterating on a wrong (f, g)
            status = peer_check(m_obj);
        }
        if (status != NB200_OK)
        {
            return status;
        }
        collect_timing(m_obj);
        point.f     = m_probe[0];
        point.dg    = m_probe[1];
        point.gmax  = m_probe[2];
        point.valid = m_probe[3] != 0.0;
        return NB200_OK;
    }

    // state.update(state0.x + t * descent) (src/lsearchk.cpp:77-91)
    int move(const double t, solver_point_t& c, bool& ok)
    {
        solver_axpy_kernel<<<blocks(), 256, 0, m_obj->stream>>>(m_n, m_x0, t, m_d, m_x);
        NB200_CUDA_TRY(cudaGetLastError());
        m_obj->launches += 1;
        const int status = evaluate(c);
        ok               = c.valid;
        return status;
    }

    // lsearchk_t::get (src/lsearchk.cpp:36-75) + lsearchk_cgdescent_t::do_get (src/lsearchk/cgdescent.cpp)
    int line_search(const solver_point_t& s0, solver_point_t& c, double t, bool& ok_out, double& t_out)
    {
        constexpr double c1 = 1e-4;                                 // lbfgs.cpp:10, cgd.cpp:77
        const double     c2 = m_method.c2;
        constexpr double theta = 0.5, gamma = 0.66, ro = 5.0;       // cgdescent.cpp:83-91
        constexpr int    max_iterations = 128;                      // lsearchk.cpp:15
        constexpr double epsilon0 = 1e-15, epsilon1 = 1e-10;        // numeric.h:92-105

        ok_out = false;
        t_out  = t;
        if (!(s0.dg < 0.0))
        {
            return NB200_OK; // not a descent direction
        }

        t        = std::isfinite(t) ? std::min(std::max(t, lsearch_stpmin()), 1.0) : 1.0;
        bool ok  = false;
        int  rc  = move(t, c, ok);
        for (int i = 0; rc == NB200_OK && i < max_iterations && !ok; ++i)
        {
            t *= 0.3;
            rc = move(t, c, ok);
        }
        for (int i = 0; rc == NB200_OK && i < max_iterations && std::fabs(c.f - s0.f) < epsilon1; ++i)
        {
            t *= 3.0;
            rc = move(t, c, ok);
            if (rc == NB200_OK && !ok)
            {
                t_out = t;
                return NB200_OK;
            }
        }
        if (rc != NB200_OK)
        {
            return rc;
        }

        const double epsilonk = 1e-6 * std::fabs(s0.f);
        int          budget   = max_iterations;
        double       ct       = t; // step size of the tentative point c
        lsearch_step_t a{0.0, s0.f, s0.dg}, b{t, c.f, c.dg};

        const auto has_armijo        = [&]() { return c.f <= s0.f + ct * c1 * s0.dg; };
        const auto has_approx_armijo = [&]() { return c.f <= s0.f + epsilonk; };
        const auto has_wolfe         = [&]() { return c.dg >= c2 * s0.dg; };
        const auto has_approx_wolfe  = [&]() { return (2.0 * c1 - 1.0) * s0.dg >= c.dg && c.dg >= c2 * s0.dg; };
        const auto done              = [&](const bool bracketed)
        {
            if ((bracketed && (a.f > s0.f + epsilonk || b.g < 0.0)) || !c.valid)
            {
                return true;
            }
            if (ct < a.t || ct > b.t)
            {
                return false;
            }
            return (has_armijo() && has_wolfe()) || (has_approx_armijo() && has_approx_wolfe());
        };
        const auto go = [&](const double tt)
        {
            ct      = tt;
            bool ok2 = false;
            rc      = move(tt, c, ok2);
            return ok2 && rc == NB200_OK;
        };
        const auto updateA = [&]() { a = lsearch_step_t{ct, c.f, c.dg}; };
        const auto updateB = [&]() { b = lsearch_step_t{ct, c.f, c.dg}; };
        const auto updateU = [&]()
        {
            for (; budget > 0 && (b.t - a.t) > lsearch_stpmin() && rc == NB200_OK; --budget)
            {
                go((1.0 - theta) * a.t + theta * b.t);
                if (!c.valid)
                {
                    return;
                }
                if (!(c.dg < 0.0))
                {
                    updateB();
                    return;
                }
                if (has_approx_armijo())
                {
                    updateA();
                }
                else
                {
                    updateB();
                }
            }
        };
        const auto update = [&]()
        {
            if (ct <= a.t || ct >= b.t)
            {
                return;
            }
            if (!(c.dg < 0.0))
            {
                updateB();
            }
            else if (has_approx_armijo())
            {
                updateA();
            }
            else
            {
                updateB();
                updateU();
            }
        };
        const auto finish = [&](const bool ok3)
        {
            ok_out = ok3;
            t_out  = ct;
            return rc;
        };

        if (done(false))
        {
            return finish(c.valid);
        }

        // bracket (cgdescent.cpp:151-175)
        {
            lsearch_step_t last_a = a;
            for (; budget > 0 && c.valid && rc == NB200_OK; --budget)
            {
                if (!(c.dg < 0.0))
                {
                    a = last_a;
                    updateB();
                    break;
                }
                if (!has_approx_armijo())
                {
                    a = lsearch_step_t{0.0, s0.f, s0.dg};
                    updateB();
                    updateU();
                    break;
                }
                last_a = lsearch_step_t{ct, c.f, c.dg};
                go(ro * ct);
            }
        }
        if (rc != NB200_OK)
        {
            return rc;
        }
        if (done(true))
        {
            return finish(c.valid);
        }

        const auto move_update_and_check_done = [&](const double tt)
        {
            if (!std::isfinite(tt) || !go(tt))
            {
                return false;
            }
            if (done(true))
            {
                return true;
            }
            update();
            return done(true);
        };

        for (int i = 0; i < budget && (b.t - a.t) > lsearch_stpmin() && rc == NB200_OK; ++i)
        {
            const lsearch_step_t a0 = a, b0 = b;
            const double         prev_width = b.t - a.t;
            const double         tc         = lsearch_secant(a0, b0);
            if (move_update_and_check_done(tc))
            {
                return finish(c.valid);
            }
            if (std::fabs(tc - a.t) < epsilon0)
            {
                if (move_update_and_check_done(lsearch_secant(a0, a)))
                {
                    return finish(c.valid);
                }
            }
            else if (std::fabs(tc - b.t) < epsilon0)
            {
                if (move_update_and_check_done(lsearch_secant(b0, b)))
                {
                    return finish(c.valid);
                }
            }
            if (b.t - a.t > gamma * prev_width)
            {
                if (move_update_and_check_done((a.t + b.t) / 2))
                {
                    return finish(c.valid);
                }
            }
        }
        return finish(false);
    }

    nb200_obj*         m_obj;
    nb200_obj*         m_group;
    int64_t            m_n;
    int64_t            m_n_global;
    nb200_allreduce_fn m_allreduce;
    void*              m_allreduce_user;
    solver_method_t    m_method;
    int                m_cap{50};
    int64_t            m_fcalls{0}, m_gcalls{0};
    double*            m_block{nullptr};
    double *           m_x{nullptr}, *m_g{nullptr}, *m_x0{nullptr}, *m_g0{nullptr}, *m_d{nullptr}, *m_q{nullptr};
    double *           m_f{nullptr}, *m_alphas{nullptr}, *m_S{nullptr}, *m_Y{nullptr};
    double *           m_probe{nullptr}, *m_probe_dev{nullptr};
};
} // namespace nb200
