// host_qp.cu — host-side algebra of the proximal bundle driver (no device code, no CUDA call).
//
// The bundle solvers of libnano (src/solver/bundle/bundle.cpp:109-168) hand the proximal-point problem
//     min_(y, r)  r + y.M.y / (2 t)   s.t.  h_j + g_j.y <= r  for every cutting plane j
// to an interior-point solver in n + 1 variables. Its dual has one variable per plane,
//     min_a  a.K.a / 2 - h.a   over the unit simplex,  K = t G M^-1 G',
// which is what a linear model with n = (D + 1) T parameters and a bundle of <= ~100 planes wants. This file solves
// that dual: Mehrotra predictor-corrector on  K a - h + nu 1 - s = 0,  1.a = 1,  a o s = 0,  a, s >= 0.
// The host drivers (libnano_b200/solver.py::rqb) call it once per trial point of the curve search.
#include "../../include/nb200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace
{
// in-place Cholesky factor (lower) of the symmetric positive definite H[m, m]; false if a pivot is not positive
bool cholesky(std::vector<double>& H, const int64_t m)
{
    for (int64_t j = 0; j < m; ++j)
    {
        double diag = H[j * m + j];
        for (int64_t k = 0; k < j; ++k)
        {
            diag -= H[j * m + k] * H[j * m + k];
        }
        if (!(diag > 0.0) || !std::isfinite(diag))
        {
            return false;
        }
        const double root = std::sqrt(diag);
        H[j * m + j]      = root;
        for (int64_t i = j + 1; i < m; ++i)
        {
            double value = H[i * m + j];
            for (int64_t k = 0; k < j; ++k)
            {
                value -= H[i * m + k] * H[j * m + k];
            }
            H[i * m + j] = value / root;
        }
    }
    return true;
}

// x <- (L L')^-1 x
void cholesky_solve(const std::vector<double>& L, const int64_t m, std::vector<double>& x)
{
    for (int64_t i = 0; i < m; ++i)
    {
        double value = x[i];
        for (int64_t k = 0; k < i; ++k)
        {
            value -= L[i * m + k] * x[k];
        }
        x[i] = value / L[i * m + i];
    }
    for (int64_t i = m - 1; i >= 0; --i)
    {
        double value = x[i];
        for (int64_t k = i + 1; k < m; ++k)
        {
            value -= L[k * m + i] * x[k];
        }
        x[i] = value / L[i * m + i];
    }
}

double step_to_boundary(const std::vector<double>& v, const std::vector<double>& dv)
{
    double step = 1.0;
    for (size_t i = 0; i < v.size(); ++i)
    {
        if (dv[i] < 0.0)
        {
            step = std::min(step, -v[i] / dv[i]);
        }
    }
    return step;
}
} // namespace

extern "C" int nb200_simplex_qp(const double* K_in, const double* h_in, const int64_t m, double* alpha)
{
    if (K_in == nullptr || h_in == nullptr || alpha == nullptr || m <= 0)
    {
        return NB200_ERR_INVALID;
    }
    if (m == 1)
    {
        alpha[0] = 1.0;
        return NB200_OK;
    }
    // the program is invariant to a common positive factor on (K, h): normalise it
    double scale = 1e-300;
    for (int64_t i = 0; i < m * m; ++i)
    {
        scale = std::max(scale, std::fabs(K_in[i]));
    }
    for (int64_t i = 0; i < m; ++i)
    {
        scale = std::max(scale, std::fabs(h_in[i]));
    }
    if (!std::isfinite(scale))
    {
        return NB200_ERR_INVALID;
    }
    std::vector<double> K(static_cast<size_t>(m * m)), h(static_cast<size_t>(m));
    for (int64_t i = 0; i < m; ++i)
    {
        for (int64_t j = 0; j < m; ++j)
        {
            K[i * m + j] = 0.5 * (K_in[i * m + j] + K_in[j * m + i]) / scale;
        }
        h[i] = h_in[i] / scale;
    }

    const size_t        um = static_cast<size_t>(m);
    std::vector<double> a(um, 1.0 / static_cast<double>(m)), s(um, 1.0), rd(um), H(um * um), z1(um), z2(um);
    std::vector<double> da(um), ds(um), da_aff(um), ds_aff(um), rc(um);
    double              nu = 0.0;

    for (int iteration = 0; iteration < 200; ++iteration)
    {
        double rd_max = 0.0, rp = -1.0, mu = 0.0;
        for (int64_t i = 0; i < m; ++i)
        {
            double value = -h[i] + nu - s[i];
            for (int64_t j = 0; j < m; ++j)
            {
                value += K[i * m + j] * a[j];
            }
            rd[i]  = value;
            rd_max = std::max(rd_max, std::fabs(value));
            rp += a[i];
            mu += a[i] * s[i];
        }
        mu /= static_cast<double>(m);
        if (std::max(rd_max, std::fabs(rp)) <= 1e-14 && mu <= 1e-16)
        {
            break;
        }

        // H = K + S / A, positive definite in exact arithmetic; a growing ridge covers the rounding of a
        // semi-definite K. Close to the solution its condition number exceeds what fp64 factorises: a nearly
        // complementary iterate is accepted as it is, anything else means K is not positive semi-definite.
        bool factorised = false;
        for (const double ridge : {0.0, 1e-14, 1e-11, 1e-8})
        {
            double diag_max = 0.0;
            for (int64_t i = 0; i < m; ++i)
            {
                diag_max = std::max(diag_max, std::fabs(K[i * m + i] + s[i] / a[i]));
            }
            H = K;
            for (int64_t i = 0; i < m; ++i)
            {
                H[i * m + i] += s[i] / a[i] + ridge * (1.0 + diag_max);
            }
            if (cholesky(H, m))
            {
                factorised = true;
                break;
            }
        }
        std::fill(z2.begin(), z2.end(), 1.0);
        double z2_sum = 0.0;
        if (factorised)
        {
            cholesky_solve(H, m, z2);
            for (const double v : z2)
            {
                z2_sum += v;
            }
        }
        if (!factorised || !(z2_sum > 0.0) || !std::isfinite(z2_sum))
        {
            if (mu <= 1e-10)
            {
                break;
            }
            return NB200_ERR_INVALID; // indefinite metric
        }

        // (K + S/A) da + 1 dnu = -rd - rc / a,  1.da = -rp,  ds = -(rc + s da) / a   (rc: complementarity residual)
        const auto solve = [&](const std::vector<double>& residual, std::vector<double>& out_da,
                               std::vector<double>& out_ds) -> double
        {
            for (int64_t i = 0; i < m; ++i)
            {
                z1[i] = -rd[i] - residual[i] / a[i];
            }
            cholesky_solve(H, m, z1);
            double z1_sum = 0.0;
            for (const double v : z1)
            {
                z1_sum += v;
            }
            const double dnu = (z1_sum + rp) / z2_sum;
            for (int64_t i = 0; i < m; ++i)
            {
                out_da[i] = z1[i] - dnu * z2[i];
                out_ds[i] = -(residual[i] + s[i] * out_da[i]) / a[i];
            }
            return dnu;
        };

        for (int64_t i = 0; i < m; ++i)
        {
            rc[i] = a[i] * s[i];
        }
        solve(rc, da_aff, ds_aff); // predictor (affine scaling)
        const double step_aff = std::min(step_to_boundary(a, da_aff), step_to_boundary(s, ds_aff));
        double       mu_aff = 0.0;
        for (int64_t i = 0; i < m; ++i)
        {
            mu_aff += (a[i] + step_aff * da_aff[i]) * (s[i] + step_aff * ds_aff[i]);
        }
        mu_aff /= static_cast<double>(m);
        const double ratio = mu > 0.0 ? mu_aff / mu : 0.0;
        const double sigma = ratio * ratio * ratio;
        for (int64_t i = 0; i < m; ++i)
        {
            rc[i] = a[i] * s[i] + da_aff[i] * ds_aff[i] - sigma * mu;
        }
        const double dnu = solve(rc, da, ds); // corrector
        // 99.5 % of the way to the boundary; the full step when no component decreases
        const auto   decreases = [](const std::vector<double>& d) { return std::any_of(d.begin(), d.end(), [](const double v) { return v < 0.0; }); };
        const double step_p = decreases(da) ? 0.995 * step_to_boundary(a, da) : 1.0;
        const double step_d = decreases(ds) ? 0.995 * step_to_boundary(s, ds) : 1.0;
        for (int64_t i = 0; i < m; ++i)
        {
            a[i] += step_p * da[i];
            s[i] += step_d * ds[i];
        }
        nu += step_d * dnu;
    }

    // planes on the inactive side of their complementarity pair get an exact zero (the bundle drops those)
    double total = 0.0, largest = 0.0;
    for (int64_t i = 0; i < m; ++i)
    {
        largest = std::max(largest, a[i]);
    }
    bool any_support = false;
    for (int64_t i = 0; i < m; ++i)
    {
        any_support = any_support || a[i] > s[i];
    }
    for (int64_t i = 0; i < m; ++i)
    {
        const bool active = any_support ? a[i] > s[i] : a[i] >= largest;
        alpha[i]          = active ? a[i] : 0.0;
        total += alpha[i];
    }
    if (!(total > 0.0) || !std::isfinite(total))
    {
        return NB200_ERR_INVALID;
    }
    for (int64_t i = 0; i < m; ++i)
    {
        alpha[i] /= total;
    }

    // one equality-constrained solve on the support removes the interior-point method's O(mu) bias:
    // K_SS a_S + nu 1 = h_S, 1.a_S = 1  =>  a_S = u - nu w  with  K_SS u = h_S,  K_SS w = 1  (a whiff of ridge: K_SS
    // may be singular). Kept only if it stays inside the simplex and does not increase the objective.
    std::vector<int64_t> support;
    for (int64_t i = 0; i < m; ++i)
    {
        if (alpha[i] > 0.0)
        {
            support.push_back(i);
        }
    }
    const int64_t k = static_cast<int64_t>(support.size());
    if (k >= 2)
    {
        std::vector<double> KS(static_cast<size_t>(k * k)), u(static_cast<size_t>(k)), w(static_cast<size_t>(k), 1.0);
        double              diag_max = 0.0;
        for (int64_t i = 0; i < k; ++i)
        {
            for (int64_t j = 0; j < k; ++j)
            {
                KS[i * k + j] = K[support[i] * m + support[j]];
            }
            u[i]     = h[support[i]];
            diag_max = std::max(diag_max, KS[i * k + i]);
        }
        for (int64_t i = 0; i < k; ++i)
        {
            KS[i * k + i] += 1e-13 * (1e-3 + diag_max);
        }
        if (cholesky(KS, k))
        {
            cholesky_solve(KS, k, u);
            cholesky_solve(KS, k, w);
            double u_sum = 0.0, w_sum = 0.0;
            for (int64_t i = 0; i < k; ++i)
            {
                u_sum += u[i];
                w_sum += w[i];
            }
            const double        shift = (u_sum - 1.0) / w_sum;
            std::vector<double> candidate(um, 0.0);
            bool                inside = std::isfinite(shift);
            double              sum = 0.0;
            for (int64_t i = 0; inside && i < k; ++i)
            {
                const double value    = u[i] - shift * w[i];
                inside                = value > 0.0 && std::isfinite(value);
                candidate[support[i]] = value;
                sum += value;
            }
            if (inside && std::fabs(sum - 1.0) < 1e-9)
            {
                const auto objective = [&](const double* v)
                {
                    double value = 0.0;
                    for (int64_t i = 0; i < m; ++i)
                    {
                        double row = 0.0;
                        for (int64_t j = 0; j < m; ++j)
                        {
                            row += K[i * m + j] * v[j];
                        }
                        value += v[i] * (0.5 * row - h[i]);
                    }
                    return value;
                };
                for (int64_t i = 0; i < m; ++i)
                {
                    candidate[i] /= sum;
                }
                if (objective(candidate.data()) <= objective(alpha))
                {
                    std::copy(candidate.begin(), candidate.end(), alpha);
                }
            }
        }
    }
    return NB200_OK;
}
