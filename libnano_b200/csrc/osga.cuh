// osga.cuh — device-resident OSGA: libnano's solver_osga_t::do_minimize (src/solver/osga.cpp:59-147, proxy_t :8-39),
// the non-smooth driver of BASELINE config 5 (hinge + lasso / elastic net), with its vectors in HBM.
//
// Per iteration the reference evaluates f(x, g) at x = xb + alpha (u - xb) and f(x') at x' = xb + alpha (u' - xb), and
// does O(n) algebra on h, u and the prox-function Q(z) = Q0 + |z - z0|^2 / 2. Here
//   * z0, xb, x, x', g, h, h_hat stay on the device; u = U(gamma, h) = z0 - h / E(gamma, h) (osga.cpp:28-31) is never
//     materialised: a point kernel forms xb + alpha ((z0 - h / e) - xb) from h and the scalar e;
//   * one model kernel per iteration turns g into g - mu (x - z0), forms h_hat = h + alpha (g - h) and reduces the
//     five numbers the scalar logic needs (|x - z0|^2, g.x, h_hat.z0, h_hat.h_hat, all-finite) in a fixed order;
//   * the host sees those scalars and the two function values through mapped memory and runs the parameter update
//     (alpha, eta, gamma: osga.cpp:130-143) and the stopping tests (solver.cpp:137-182, track.cpp:11-59).
// Two synchronisations per iteration (one per evaluation), no vector crosses PCIe until the result is copied out.
//
// Included by nb200_api.cu after group.cuh.
#pragma once

namespace nb200
{
#ifdef __CUDACC__

// out = xb + alpha * ((z0 - h * inv_e) - xb); stats[0] = max|out - xb|, stats[1] = max|xb|, stats[2] = all finite
__global__ void __launch_bounds__(kSolverThreads) osga_point_kernel(const int64_t n, const double* __restrict__ xb,
                                                                    const double* __restrict__ z0,
                                                                    const double* __restrict__ h, const double inv_e,
                                                                    const double alpha, double* __restrict__ out,
                                                                    double* stats)
{
    __shared__ double scratch[kSolverThreads / 32];
    double dmax = 0.0, bmax = 0.0, bad = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
    {
        const double b = xb[i];
        const double u = z0[i] - h[i] * inv_e;
        const double v = b + alpha * (u - b);
        out[i]         = v;
        dmax           = fmax(dmax, fabs(v - b));
        bmax           = fmax(bmax, fabs(b));
        bad += isfinite(v) ? 0.0 : 1.0;
    }
    dmax = cta_max(dmax, scratch);
    bmax = cta_max(bmax, scratch);
    bad  = cta_sum(bad, scratch);
    if (threadIdx.x == 0)
    {
        stats[0] = dmax;
        stats[1] = bmax;
        stats[2] = (bad == 0.0) ? 1.0 : 0.0;
    }
}

// g <- g - mu (x - z0)   (g - miu * proxy.gQ(x), osga.cpp:110);  h_hat = h + alpha (g - h)   (:112)
// sums[0] = |x - z0|^2, [1] = g.x, [2] = h_hat.z0, [3] = h_hat.h_hat, [4] = max|g| before the shift, [5] = all finite
__global__ void __launch_bounds__(kSolverThreads) osga_model_kernel(const int64_t n, const double* __restrict__ x,
                                                                    double* __restrict__ g,
                                                                    const double* __restrict__ z0,
                                                                    const double* __restrict__ h, const double alpha,
                                                                    const double mu, double* __restrict__ h_hat,
                                                                    double* sums)
{
    __shared__ double scratch[kSolverThreads / 32];
    double dxz = 0.0, gx = 0.0, hz = 0.0, hh = 0.0, gmax = 0.0, bad = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += kSolverThreads)
    {
        const double xi = x[i], zi = z0[i], graw = g[i];
        const double d  = xi - zi;
        const double gq = graw - mu * d;
        const double hi = h[i];
        const double hn = hi + alpha * (gq - hi);
        g[i]            = gq;
        h_hat[i]        = hn;
        dxz             = fma(d, d, dxz);
        gx              = fma(gq, xi, gx);
        hz              = fma(hn, zi, hz);
        hh              = fma(hn, hn, hh);
        gmax            = fmax(gmax, fabs(graw));
        bad += (isfinite(graw) && isfinite(xi)) ? 0.0 : 1.0;
    }
    dxz  = cta_sum(dxz, scratch);
    gx   = cta_sum(gx, scratch);
    hz   = cta_sum(hz, scratch);
    hh   = cta_sum(hh, scratch);
    gmax = cta_max(gmax, scratch);
    bad  = cta_sum(bad, scratch);
    if (threadIdx.x == 0)
    {
        sums[0] = dxz;
        sums[1] = gx;
        sums[2] = hz;
        sums[3] = hh;
        sums[4] = gmax;
        sums[5] = (bad == 0.0) ? 1.0 : 0.0;
        __threadfence_system();
    }
}

#endif // __CUDACC__

struct osga_params_t
{
    double  epsilon{1e-8};
    int64_t max_evals{1000};
    int64_t patience{100};
    double  lambda{0.9};       // solver::osga::lambda (osga.cpp:49)
    double  alpha_max{0.7};    // solver::osga::alpha_max (:50)
    double  kappa_prime{0.1};  // solver::osga::kappas (:51)
    double  kappa{1.1};
    double  strong_convexity{0.0};
};

class device_osga_t
{
public:
    device_osga_t(nb200_obj* obj, nb200_obj* group, const int64_t n_global) : m_obj(obj), m_group(group), m_n(obj->n), m_n_global(n_global) {}

    ~device_osga_t()
    {
        cudaFree(m_block);
        cudaFreeHost(m_host);
    }

    // status: 0 max_evals, 1 gradient test (|g0|_inf < 1e-15), 2 specific test (eta < epsilon), 3 value test, 4 failed
    int run(double* x_host, const osga_params_t& q, double* fx, double* eta_out, int64_t* stats, int* status_out)
    {
        const size_t n = static_cast<size_t>(m_n);
        NB200_CUDA_TRY(cudaMalloc(&m_block, 7 * (n + 2) * sizeof(double)));
        double* z0   = m_block;
        double* xb   = z0 + (n + 2);
        double* x    = xb + (n + 2);
        double* xp   = x + (n + 2);
        double* g    = xp + (n + 2);
        double* h    = g + (n + 2);
        double* hhat = h + (n + 2);
        // mapped scalars: [0..5] model sums, [8] f, [9] f', [16..18] point stats of x, [24..26] point stats of x'
        NB200_CUDA_TRY(cudaHostAlloc(&m_host, 32 * sizeof(double), cudaHostAllocMapped));
        NB200_CUDA_TRY(cudaHostGetDevicePointer(&m_dev, m_host, 0));
        cudaStream_t stream = m_obj->stream;

        double norm2 = 0.0;
        for (size_t i = 0; i < n; ++i)
        {
            norm2 += x_host[i] * x_host[i];
        }
        const double mu = q.strong_convexity / 2.0;                                         // osga.cpp:70
        const double Q0 = 0.5 * std::sqrt(std::sqrt(norm2) + 2.220446049250313e-16);        // proxy_t (:10-14)
        const auto   E  = [&](const double gamma, const double hz, const double hh)          // proxy_t::E (:20-26)
        {
            const double beta = gamma + hz;
            const double root = std::sqrt(beta * beta + 2.0 * Q0 * hh);
            return (beta <= 0.0) ? (-beta + root) / (2.0 * Q0) : hh / (beta + root);
        };

        NB200_CUDA_TRY(cudaMemcpyAsync(z0, x_host, n * sizeof(double), cudaMemcpyHostToDevice, stream));
        NB200_CUDA_TRY(cudaMemcpyAsync(xb, z0, n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        NB200_CUDA_TRY(cudaMemsetAsync(h, 0, n * sizeof(double), stream));

        // the initial state: f0, g0; h = g0 - mu gQ(x0) (= the model kernel with alpha = 1 on h = 0)
        int rc = evaluate(xb, g, m_dev + 8, true);
        if (rc == NB200_OK)
        {
            rc = model(xb, g, z0, h, 1.0, mu, hhat);
        }
        if (rc != NB200_OK)
        {
            return rc;
        }
        std::swap(h, hhat);
        double       fb     = m_host[8];
        const double gmax0  = m_host[4];
        bool         valid  = m_host[5] != 0.0 && std::isfinite(fb);
        double       hz = m_host[2], hh = m_host[3];
        double       gamma  = fb - mu * (Q0 + 0.5 * m_host[0]) - m_host[1];                    // osga.cpp:77
        double       e_u    = E(gamma - fb, hz, hh);                                        // u = z0 - h / e_u (:78)
        double       eta    = e_u - mu;                                                     // :79
        double       alpha  = q.alpha_max;
        int          status = 0;
        int64_t      iterations = 0;
        std::vector<double> history_df, history_dx;
        const auto value_test = [&]() // solver_track_t::value_test_unconstrained (track.cpp:24-59)
        {
            const size_t size = history_df.size();
            size_t       last = size;
            double       dd   = std::numeric_limits<double>::max();
            for (size_t it = size; it > 0; --it)
            {
                if (history_df[it - 1] > 0.0)
                {
                    dd   = std::max(history_dx[it - 1], history_df[it - 1]);
                    last = it - 1;
                    break;
                }
            }
            if (last == size)
            {
                return size >= static_cast<size_t>(q.patience) ? 0.0 : dd;
            }
            return (last + static_cast<size_t>(q.patience) >= size) ? dd : 0.0;
        };

        if (!valid)
        {
            status = 4;
        }
        while (status == 0 && m_fcalls + m_gcalls < q.max_evals)
        {
            if (gmax0 < 1e-15) // the state's gradient is the initial one throughout (state.cpp:73-98, osga.cpp:98-103)
            {
                status = 1;
                break;
            }
            // x = xb + alpha (u - xb); f(x, g)
            rc = point(xb, z0, h, 1.0 / e_u, alpha, x, m_dev + 16);
            if (rc == NB200_OK)
            {
                rc = evaluate(x, g, m_dev + 8, true);
            }
            if (rc == NB200_OK)
            {
                rc = model(x, g, z0, h, alpha, mu, hhat);
            }
            if (rc != NB200_OK)
            {
                return rc;
            }
            const double f         = m_host[8];
            const double Qx        = Q0 + 0.5 * m_host[0];
            const double hz_hat    = m_host[2], hh_hat = m_host[3];
            const double gamma_hat = gamma + alpha * (f - mu * Qx - m_host[1] - gamma);       // :113
            const bool   x_better  = f < fb;                                                // :115-116
            const double fb_prime  = x_better ? f : fb;
            const double dx_x      = m_host[16] / std::max(1.0, m_host[17]);

            // x' = xb + alpha (u' - xb), u' = U(gamma_hat - fb', h_hat); f(x')
            const double e_prime = E(gamma_hat - fb_prime, hz_hat, hh_hat);
            rc                   = point(xb, z0, hhat, 1.0 / e_prime, alpha, xp, m_dev + 24);
            if (rc == NB200_OK)
            {
                rc = evaluate(xp, nullptr, m_dev + 9, false);
            }
            if (rc == NB200_OK)
            {
                rc = sync();
            }
            if (rc != NB200_OK)
            {
                return rc;
            }
            const double f_prime   = m_host[9];
            const bool   xp_better = f_prime < fb_prime;                                    // :122-123
            const double fb_hat    = xp_better ? f_prime : fb_prime;
            const double dx_hat    = xp_better ? m_host[24] / std::max(1.0, m_host[25]) : (x_better ? dx_x : 0.0);
            const double e_hat     = E(gamma_hat - fb_hat, hz_hat, hh_hat);                 // u_hat, eta_hat (:127-128)
            const double eta_hat   = e_hat - mu;
            ++iterations;

            // state.update_if_better(xb_hat, fb_hat): the best point is xb throughout (fb never increases)
            const bool improved = std::isfinite(fb_hat) && fb > fb_hat;
            valid = std::isfinite(improved ? fb_hat : fb) && (xp_better ? m_host[26] != 0.0 : (x_better ? m_host[18] != 0.0 : true));
            history_df.push_back(improved ? fb - fb_hat : 0.0); // done_specific_test -> update_history (solver.cpp:175-182)
            history_dx.push_back(improved ? dx_hat : 0.0);
            if (improved)
            {
                std::swap(xb, xp_better ? xp : x); // xb = xb_hat (:145)
                fb = fb_hat;
            }
            if (!valid)
            {
                status = 4;
                break;
            }
            if (eta_hat < q.epsilon)
            {
                status = 2;
                eta    = eta_hat;
                break;
            }
            history_df.push_back(0.0); // done_value_test records the (unchanged) state once more (solver.cpp:137-147)
            history_dx.push_back(0.0);
            if (value_test() < q.epsilon)
            {
                status = 3;
                eta    = eta_hat;
                break;
            }

            // the parameters (alpha, h, gamma, eta, u) (osga.cpp:130-143)
            const double R = (eta - eta_hat) / (q.lambda * alpha * eta);
            alpha = (R < 1.0) ? alpha * std::exp(-q.kappa) : std::min(alpha * std::exp(q.kappa_prime * (R - 1.0)), q.alpha_max);
            if (eta_hat < eta)
            {
                std::swap(h, hhat);
                e_u   = e_hat;
                eta   = eta_hat;
                gamma = gamma_hat;
            }
        }

        NB200_CUDA_TRY(cudaMemcpyAsync(x_host, xb, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
        rc = sync();
        if (rc != NB200_OK)
        {
            return rc;
        }
        *fx         = fb;
        *eta_out    = eta;
        stats[0]    = m_fcalls;
        stats[1]    = m_gcalls;
        stats[2]    = iterations;
        *status_out = status;
        return NB200_OK;
    }

private:
    int sync()
    {
        int status = obj_sync(m_obj);
        if (status == NB200_OK)
        {
            status = peer_check(m_obj);
        }
        if (status == NB200_OK)
        {
            collect_timing(m_obj);
        }
        return status;
    }

    int evaluate(const double* x, double* g, double* f_out, const bool want_grad)
    {
        const int status = solver_enqueue_eval(m_obj, m_group, x, m_n_global, g, f_out, want_grad);
        m_fcalls += 1;
        m_gcalls += want_grad ? 1 : 0;
        return status;
    }

    int point(const double* xb, const double* z0, const double* h, const double inv_e, const double alpha, double* out,
              double* stats)
    {
        osga_point_kernel<<<1, kSolverThreads, 0, m_obj->stream>>>(m_n, xb, z0, h, inv_e, alpha, out, stats);
        NB200_CUDA_TRY(cudaGetLastError());
        m_obj->launches += 1;
        return NB200_OK;
    }

    int model(const double* x, double* g, const double* z0, const double* h, const double alpha, const double mu,
              double* hhat)
    {
        osga_model_kernel<<<1, kSolverThreads, 0, m_obj->stream>>>(m_n, x, g, z0, h, alpha, mu, hhat, m_dev);
        NB200_CUDA_TRY(cudaGetLastError());
        m_obj->launches += 1;
        return sync();
    }

    nb200_obj* m_obj;
    nb200_obj* m_group;
    int64_t    m_n;
    int64_t    m_n_global;
    int64_t    m_fcalls{0}, m_gcalls{0};
    double*    m_block{nullptr};
    double *   m_host{nullptr}, *m_dev{nullptr};
};
} // namespace nb200
