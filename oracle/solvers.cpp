// solvers.cpp — restated DRIVERS of the hot path (test infrastructure, like the rest of oracle/).
//
// libnano's solvers are callers of function_t::operator() and are out of scope to re-implement as product code
// (SURVEY.md §2); they cannot be compiled here either (every solver header includes Eigen3). To check the north-star
// requirement "solver iterates reach the same optimum within the solver's epsilon" end to end, the two drivers the
// BASELINE configs name are restated over a plain (x, gx) -> fx callback, so the SAME driver can minimise the CPU
// oracle objective and the GPU objective and the optima can be compared:
//   L-BFGS            src/solver/lbfgs.cpp:25-124      (history 50, tolerance (1e-4, 0.9))
//   line-search glue  src/solver/lsearch.cpp:11-26, src/lsearchk.cpp:36-91
//   lsearch0          src/lsearch0/quadratic.cpp:16-33 (alpha = 1.01)
//   lsearchk          src/lsearchk/cgdescent.cpp       (Hager-Zhang CG-DESCENT: epsilon 1e-6, theta 0.5, gamma 0.66, ro 5)
//   step formulas     src/solver/lstep.cpp:25-88
//   state predicates  src/solver/state.cpp:165-210, include/nano/solver/state.h:139-146
//   stopping          src/solver.cpp:114-173, src/solver/track.cpp:11-59
//   OSGA              src/solver/osga.cpp:8-147        (lambda 0.9, alpha_max 0.7, kappas (0.1, 1.1))
//   CGD               src/solver/cgd.cpp:7-143,240-299 (ten beta formulas, tolerance (1e-4, 1e-1), orthotest 0.1)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <deque>
#include <limits>
#include <vector>

namespace
{
using vec = std::vector<double>;

constexpr double kEpsilon0 = 1e-15; // include/nano/core/numeric.h:92-97
constexpr double kEpsilon1 = 1e-10; // include/nano/core/numeric.h:99-105

typedef double (*callback_t)(void* ctx, const double* x, double* gx);

double dot(const vec& a, const vec& b)
{
    double s = 0.0;
    for (size_t i = 0; i < a.size(); ++i)
    {
        s += a[i] * b[i];
    }
    return s;
}

double norm_inf(const vec& a)
{
    double m = 0.0;
    for (const double v : a)
    {
        m = std::max(m, std::fabs(v));
    }
    return m;
}

bool all_finite(const vec& a)
{
    return std::all_of(a.begin(), a.end(), [](const double v) { return std::isfinite(v); });
}

// function_t's call bookkeeping (src/function.cpp:258-260): one f per call, one g per call with a gradient buffer
struct counted_function_t
{
    callback_t m_callback;
    void*      m_ctx;
    int64_t    m_fcalls{0};
    int64_t    m_gcalls{0};

    double operator()(const vec& x, vec* gx = nullptr)
    {
        ++m_fcalls;
        m_gcalls += (gx != nullptr) ? 1 : 0;
        return m_callback(m_ctx, x.data(), gx != nullptr ? gx->data() : nullptr);
    }

    int64_t calls() const { return m_fcalls + m_gcalls; }
};

// solver_track_t (src/solver/track.cpp:11-59)
struct track_t
{
    vec                 m_prev_x;
    double              m_prev_fx{0.0};
    std::vector<double> m_df, m_dx;

    void update(const vec& x, const double fx)
    {
        double dmax = 0.0;
        for (size_t i = 0; i < x.size(); ++i)
        {
            dmax = std::max(dmax, std::fabs(m_prev_x[i] - x[i]));
        }
        m_df.push_back(m_prev_fx - fx);
        m_dx.push_back(dmax / std::max(1.0, norm_inf(m_prev_x)));
        m_prev_x  = x;
        m_prev_fx = fx;
    }

    double value_test(const int64_t patience) const
    {
        const size_t size = m_df.size();
        size_t       last = size;
        double       dd   = std::numeric_limits<double>::max();
        for (size_t it = size; it > 0; --it)
        {
            if (m_df[it - 1] > 0.0)
            {
                dd   = std::max(m_dx[it - 1], m_df[it - 1]);
                last = it - 1;
                break;
            }
        }
        if (last == size)
        {
            return size >= static_cast<size_t>(patience) ? 0.0 : dd;
        }
        return (last + static_cast<size_t>(patience) >= size) ? dd : 0.0;
    }
};

// solver_state_t restricted to unconstrained problems (src/solver/state.cpp)
struct state_t
{
    counted_function_t* m_function{nullptr};
    vec                 m_x, m_gx;
    double              m_fx{0.0};
    track_t             m_track;

    state_t(counted_function_t& function, const vec& x0)
        : m_function(&function)
        , m_x(x0)
        , m_gx(x0.size(), 0.0)
    {
        m_fx              = function(m_x, &m_gx);
        m_track.m_prev_x  = m_x;
        m_track.m_prev_fx = m_fx;
    }

    bool update(const vec& x) // include/nano/solver/state.h:45-52
    {
        m_x  = x;
        m_fx = (*m_function)(m_x, &m_gx);
        return valid();
    }

    bool update_if_better(const vec& x, const double fx) // src/solver/state.cpp:73-98 (gradient kept)
    {
        if (std::isfinite(fx) && m_fx > fx)
        {
            m_x  = x;
            m_fx = fx;
            return true;
        }
        return false;
    }

    bool   valid() const { return std::isfinite(m_fx) && all_finite(m_x) && all_finite(m_gx); } // state.cpp:170-174
    double dg(const vec& descent) const { return dot(m_gx, descent); }                          // state.h:139
    bool   has_descent(const vec& descent) const { return dg(descent) < 0.0; }                  // state.h:146
    double gradient_test() const { return norm_inf(m_gx) / std::max(1.0, std::fabs(m_fx)); }    // state.cpp:165-168

    // src/solver/state.cpp:176-210
    bool has_armijo(const state_t& origin, const vec& descent, const double t, const double c1) const
    {
        return m_fx <= origin.m_fx + t * c1 * origin.dg(descent);
    }
    bool has_approx_armijo(const state_t& origin, const double epsilon) const { return m_fx <= origin.m_fx + epsilon; }
    bool has_wolfe(const state_t& origin, const vec& descent, const double c2) const
    {
        return dg(descent) >= c2 * origin.dg(descent);
    }
    bool has_approx_wolfe(const state_t& origin, const vec& descent, const double c1, const double c2) const
    {
        return (2.0 * c1 - 1.0) * origin.dg(descent) >= dg(descent) && dg(descent) >= c2 * origin.dg(descent);
    }
};

// lsearch_step_t (src/solver/lstep.cpp)
struct step_t
{
    double t, f, g;
};

double stpmin()
{
    return 10.0 * std::numeric_limits<double>::epsilon();
}

double stpmax()
{
    return 1.0 / stpmin();
}

double secant(const step_t& u, const step_t& v)
{
    return std::clamp((v.t * u.g - u.t * v.g) / (u.g - v.g), stpmin(), stpmax());
}

step_t make_step(const state_t& state, const vec& descent, const double t)
{
    return {t, state.m_fx, state.dg(descent)};
}

// CG-DESCENT line search (src/lsearchk/cgdescent.cpp) behind the generic glue of src/lsearchk.cpp:36-91
class cgdescent_t
{
public:
    cgdescent_t(const double c1, const double c2)
        : m_c1(c1)
        , m_c2(c2)
    {
    }

    // returns (ok, step size); `state` moves to the accepted point
    std::pair<bool, double> get(state_t& state, const vec& descent, double t) const
    {
        if (!state.has_descent(descent))
        {
            return {false, t};
        }
        const state_t state0 = state;

        t = std::isfinite(t) ? std::clamp(t, stpmin(), 1.0) : 1.0;
        for (int i = 0; i < m_max_iterations && !move_to(state, state0, descent, t); ++i)
        {
            t *= 0.3; // initial step length is too large
        }
        for (int i = 0; i < m_max_iterations && std::fabs(state.m_fx - state0.m_fx) < kEpsilon1; ++i)
        {
            t *= 3.0; // initial step length is too small
            if (!move_to(state, state0, descent, t))
            {
                return {false, t};
            }
        }
        return search(state0, descent, t, state);
    }

private:
    struct interval_t
    {
        const state_t& state0;
        const vec&     descent;
        double         t;
        state_t&       c;
        step_t         a, b;

        void updateA() { a = make_step(c, descent, t); }
        void updateB() { b = make_step(c, descent, t); }
    };

    static bool move_to(state_t& state, const state_t& state0, const vec& descent, const double t)
    {
        vec x(state0.m_x.size());
        for (size_t i = 0; i < x.size(); ++i)
        {
            x[i] = state0.m_x[i] + t * descent[i];
        }
        return state.update(x);
    }

    bool move(interval_t& iv, const double t) const
    {
        iv.t = t;
        return move_to(iv.c, iv.state0, iv.descent, t);
    }

    bool done(const interval_t& iv, const double epsilonk, const bool bracketed = true) const
    {
        if ((bracketed && (iv.a.f > iv.state0.m_fx + epsilonk || iv.b.g < 0.0)) || !iv.c.valid())
        {
            return true; // bracketing failed or diverged
        }
        if (iv.t < iv.a.t || iv.t > iv.b.t)
        {
            return false; // tentative point outside the bracketing interval
        }
        return (iv.c.has_armijo(iv.state0, iv.descent, iv.t, m_c1) && iv.c.has_wolfe(iv.state0, iv.descent, m_c2)) ||
               (iv.c.has_approx_armijo(iv.state0, epsilonk) &&
                iv.c.has_approx_wolfe(iv.state0, iv.descent, m_c1, m_c2));
    }

    void updateU(interval_t& iv, const double epsilonk, int& budget) const
    {
        for (; budget > 0 && (iv.b.t - iv.a.t) > stpmin(); --budget)
        {
            move(iv, (1.0 - m_theta) * iv.a.t + m_theta * iv.b.t);
            if (!iv.c.valid())
            {
                return;
            }
            if (!iv.c.has_descent(iv.descent))
            {
                iv.updateB();
                return;
            }
            if (iv.c.has_approx_armijo(iv.state0, epsilonk))
            {
                iv.updateA();
            }
            else
            {
                iv.updateB();
            }
        }
    }

    void update(interval_t& iv, const double epsilonk, int& budget) const
    {
        if (iv.t <= iv.a.t || iv.t >= iv.b.t)
        {
            return;
        }
        if (!iv.c.has_descent(iv.descent))
        {
            iv.updateB();
        }
        else if (iv.c.has_approx_armijo(iv.state0, epsilonk))
        {
            iv.updateA();
        }
        else
        {
            iv.updateB();
            updateU(iv, epsilonk, budget);
        }
    }

    void bracket(interval_t& iv, const double epsilonk, int& budget) const
    {
        step_t last_a = iv.a;
        for (; budget > 0 && iv.c.valid(); --budget)
        {
            if (!iv.c.has_descent(iv.descent))
            {
                iv.a = last_a;
                iv.updateB();
                return;
            }
            if (!iv.c.has_approx_armijo(iv.state0, epsilonk))
            {
                iv.a = make_step(iv.state0, iv.descent, 0.0);
                iv.updateB();
                updateU(iv, epsilonk, budget);
                return;
            }
            last_a = make_step(iv.c, iv.descent, iv.t);
            move(iv, m_ro * iv.t);
        }
    }

    std::pair<bool, double> search(const state_t& state0, const vec& descent, const double t0, state_t& state) const
    {
        const double epsilonk = m_epsilon * std::fabs(state0.m_fx);
        int          budget   = m_max_iterations;

        interval_t iv{state0, descent, t0, state, make_step(state0, descent, 0.0), make_step(state, descent, t0)};
        if (done(iv, epsilonk, false))
        {
            return {state.valid(), iv.t};
        }
        bracket(iv, epsilonk, budget);
        if (done(iv, epsilonk))
        {
            return {state.valid(), iv.t};
        }

        const auto move_update_and_check_done = [&](const double t)
        {
            if (!std::isfinite(t) || !move(iv, t))
            {
                return false; // interpolation failed, go on with bisection
            }
            if (done(iv, epsilonk))
            {
                return true;
            }
            update(iv, epsilonk, budget);
            return done(iv, epsilonk);
        };

        for (int i = 0; i < budget && (iv.b.t - iv.a.t) > stpmin(); ++i)
        {
            const step_t a0 = iv.a, b0 = iv.b;
            const double prev_width = iv.b.t - iv.a.t;

            const double tc = secant(a0, b0);
            if (move_update_and_check_done(tc))
            {
                return {state.valid(), iv.t};
            }
            if (std::fabs(tc - iv.a.t) < kEpsilon0)
            {
                if (move_update_and_check_done(secant(a0, iv.a)))
                {
                    return {state.valid(), iv.t};
                }
            }
            else if (std::fabs(tc - iv.b.t) < kEpsilon0)
            {
                if (move_update_and_check_done(secant(b0, iv.b)))
                {
                    return {state.valid(), iv.t};
                }
            }
            if (iv.b.t - iv.a.t > m_gamma * prev_width)
            {
                if (move_update_and_check_done((iv.a.t + iv.b.t) / 2))
                {
                    return {state.valid(), iv.t};
                }
            }
        }
        return {false, iv.t};
    }

    double m_c1, m_c2;
    int    m_max_iterations{128};
    double m_epsilon{1e-6}, m_theta{0.5}, m_gamma{0.66}, m_ro{5.0};
};

enum status_t : int
{
    STATUS_MAX_ITERS     = 0,
    STATUS_GRADIENT_TEST = 1,
    STATUS_VALUE_TEST    = 2,
    STATUS_SPECIFIC_TEST = 3,
    STATUS_FAILED        = 4
};
} // namespace

extern "C" {

// solver_lbfgs_t::do_minimize (src/solver/lbfgs.cpp:25-124) with lsearch0 = quadratic, lsearchk = cgdescent.
// x: in = x0, out = the returned state's x. stats = {fcalls, gcalls, iterations}. Returns a status_t.
int oracle_lbfgs(callback_t callback, void* ctx, const int64_t n, double* x, const double epsilon,
                 const int64_t max_evals, const int64_t history, double* fx, double* gradient_test, int64_t* stats)
{
    counted_function_t function{callback, ctx};
    const vec          x0(x, x + n);

    state_t cstate(function, x0);
    int     status     = STATUS_MAX_ITERS;
    int64_t iterations = 0;

    const auto finish = [&](const state_t& state)
    {
        std::copy(state.m_x.begin(), state.m_x.end(), x);
        *fx            = state.m_fx;
        *gradient_test = state.gradient_test();
        stats[0]       = function.m_fcalls;
        stats[1]       = function.m_gcalls;
        stats[2]       = iterations;
        return status;
    };

    if (norm_inf(cstate.m_gx) < kEpsilon0) // lbfgs.cpp:34-38
    {
        status = cstate.valid() ? STATUS_GRADIENT_TEST : STATUS_FAILED;
        return finish(cstate);
    }

    state_t           pstate = cstate;
    const cgdescent_t lsearchk(1e-4, 0.9); // solver::tolerance of lbfgs (lbfgs.cpp:10)
    double            last_step_size = -1.0, prevf = 0.0, prevdg = 1.0; // lsearch.h:30, quadratic.h:33-34

    std::deque<vec> ss, ys;
    vec             q, r(static_cast<size_t>(n));
    while (function.calls() < max_evals)
    {
        // two-loop recursion (Nocedal & Wright, p.178; lbfgs.cpp:49-84)
        q                  = cstate.m_gx;
        const size_t hsize = ss.size();
        vec          alphas(hsize);
        for (size_t j = 0; j < hsize; ++j)
        {
            const vec&   s     = ss[hsize - 1 - j];
            const vec&   y     = ys[hsize - 1 - j];
            const double alpha = dot(s, q) / dot(s, y);
            for (int64_t i = 0; i < n; ++i)
            {
                q[i] -= alpha * y[i];
            }
            alphas[j] = alpha;
        }
        if (ss.empty())
        {
            r = q;
        }
        else
        {
            const double scale = dot(ss[hsize - 1], ys[hsize - 1]) / dot(ys[hsize - 1], ys[hsize - 1]);
            for (int64_t i = 0; i < n; ++i)
            {
                r[i] = scale * q[i];
            }
        }
        for (size_t j = 0; j < hsize; ++j)
        {
            const double alpha = alphas[hsize - 1 - j];
            const double beta  = dot(ys[j], r) / dot(ss[j], ys[j]);
            for (int64_t i = 0; i < n; ++i)
            {
                r[i] += ss[j][i] * (alpha - beta);
            }
        }
        vec descent(static_cast<size_t>(n));
        for (int64_t i = 0; i < n; ++i)
        {
            descent[i] = -r[i];
        }

        const bool has_descent = cstate.has_descent(descent); // lbfgs.cpp:89-94
        if (!has_descent)
        {
            for (int64_t i = 0; i < n; ++i)
            {
                descent[i] = -cstate.m_gx[i];
            }
        }

        // lsearch_t::get (src/solver/lsearch.cpp:11-26) with lsearch0_quadratic_t::get (quadratic.cpp:16-33)
        pstate = cstate;
        double t0;
        if (last_step_size < 0.0)
        {
            t0 = 1.0;
        }
        else
        {
            t0 = std::min(1.0, 1.01 * 2.0 * (cstate.m_fx - prevf) / prevdg);
        }
        prevf  = cstate.m_fx;
        prevdg = cstate.dg(descent);

        const auto [iter_ok, step_size] = lsearchk.get(cstate, descent, t0);
        last_step_size                  = step_size;
        ++iterations;

        // solver_t::done_gradient_test for a smooth function (src/solver.cpp:149-173, 114-135)
        cstate.m_track.update(cstate.m_x, cstate.m_fx);
        const bool step_ok   = iter_ok && cstate.valid();
        const bool converged = cstate.gradient_test() < epsilon;
        if (!step_ok)
        {
            status = STATUS_FAILED;
            break;
        }
        if (converged)
        {
            status = STATUS_GRADIENT_TEST;
            break;
        }

        if (has_descent) // lbfgs.cpp:105-120
        {
            vec s(static_cast<size_t>(n)), y(static_cast<size_t>(n));
            for (int64_t i = 0; i < n; ++i)
            {
                s[i] = cstate.m_x[i] - pstate.m_x[i];
                y[i] = cstate.m_gx[i] - pstate.m_gx[i];
            }
            ss.push_back(std::move(s));
            ys.push_back(std::move(y));
            if (ss.size() > static_cast<size_t>(history))
            {
                ss.pop_front();
                ys.pop_front();
            }
        }
        else
        {
            ss.clear();
            ys.clear();
        }
    }

    return finish(cstate.valid() ? cstate : pstate); // solver_t::choose (src/solver.cpp:219-225)
}

// solver_cgd_t::do_minimize (src/solver/cgd.cpp:86-143) with lsearch0 = quadratic, lsearchk = cgdescent and the
// tolerance (1e-4, 1e-1) of cgd.cpp:77. beta: 0 HS+ 1 FR 2 PR+ 3 CD 4 LS+ 5 DY 6 N 7 DYCD 8 DYHS 9 FRPR
// (cgd.cpp:7-71,240-299). x: in = x0, out = the returned state's x. stats = {fcalls, gcalls, iterations}.
int oracle_cgd(callback_t callback, void* ctx, const int64_t n, double* x, const int beta_kind, const double epsilon,
               const int64_t max_evals, const double orthotest, const double eta0, double* fx, double* gradient_test,
               int64_t* stats)
{
    counted_function_t function{callback, ctx};
    const vec          x0(x, x + n);

    state_t cstate(function, x0);
    int     status     = STATUS_MAX_ITERS;
    int64_t iterations = 0;

    const auto finish = [&](const state_t& state)
    {
        std::copy(state.m_x.begin(), state.m_x.end(), x);
        *fx            = state.m_fx;
        *gradient_test = state.gradient_test();
        stats[0]       = function.m_fcalls;
        stats[1]       = function.m_gcalls;
        stats[2]       = iterations;
        return status;
    };

    if (norm_inf(cstate.m_gx) < kEpsilon0) // cgd.cpp:95-99
    {
        status = cstate.valid() ? STATUS_GRADIENT_TEST : STATUS_FAILED;
        return finish(cstate);
    }

    // the ten formulas, written over the dot products the reference takes (cgd.cpp:7-71)
    const auto beta = [&](const vec& pg, const vec& pd, const vec& cg)
    {
        double cg_y = 0.0, pd_y = 0.0, y_y = 0.0; // y = cg - pg
        for (int64_t i = 0; i < n; ++i)
        {
            const double y = cg[i] - pg[i];
            cg_y += cg[i] * y;
            pd_y += pd[i] * y;
            y_y += y * y;
        }
        const double cg_cg = dot(cg, cg), pg_pg = dot(pg, pg), pd_pg = dot(pd, pg);
        const double HS = cg_y / pd_y, FR = cg_cg / pg_pg, PR = cg_y / pg_pg, CD = -cg_cg / pd_pg, LS = -cg_y / pd_pg,
                     DY = cg_cg / pd_y;
        switch (beta_kind)
        {
        case 0: return std::max(HS, 0.0);
        case 1: return FR;
        case 2: return std::max(PR, 0.0);
        case 3: return CD;
        case 4: return std::max(LS, 0.0);
        case 5: return DY;
        case 6:
        {
            const double div = 1.0 / pd_y;
            const double eta = -1.0 / (std::sqrt(dot(pd, pd)) * std::min(eta0, std::sqrt(pg_pg)));
            double       sum = 0.0;
            for (int64_t i = 0; i < n; ++i)
            {
                sum += ((cg[i] - pg[i]) - 2.0 * pd[i] * y_y * div) * cg[i];
            }
            return std::max(eta, div * sum);
        }
        case 7: return cg_cg / std::max(pd_y, -pd_pg);
        case 8: return std::max(0.0, std::min(DY, HS));
        default: return (PR < -FR) ? -FR : ((std::fabs(PR) <= FR) ? PR : FR);
        }
    };

    state_t           pstate = cstate;
    const cgdescent_t lsearchk(1e-4, 1e-1);
    double            last_step_size = -1.0, prevf = 0.0, prevdg = 1.0;
    vec               cdescent, pdescent;
    while (function.calls() < max_evals)
    {
        if (cdescent.empty())
        {
            cdescent.resize(static_cast<size_t>(n));
            for (int64_t i = 0; i < n; ++i)
            {
                cdescent[i] = -cstate.m_gx[i];
            }
        }
        else
        {
            const double b = beta(pstate.m_gx, pdescent, cstate.m_gx);
            for (int64_t i = 0; i < n; ++i)
            {
                cdescent[i] = -cstate.m_gx[i] + b * pdescent[i];
            }
            // restart: not a descent direction, or consecutive gradients far from orthogonal (cgd.cpp:121-126)
            if (!cstate.has_descent(cdescent) ||
                std::fabs(dot(cstate.m_gx, pstate.m_gx)) >= orthotest * dot(cstate.m_gx, cstate.m_gx))
            {
                for (int64_t i = 0; i < n; ++i)
                {
                    cdescent[i] = -cstate.m_gx[i];
                }
            }
        }
        pstate   = cstate;
        pdescent = cdescent;

        double t0;
        if (last_step_size < 0.0)
        {
            t0 = 1.0;
        }
        else
        {
            t0 = std::min(1.0, 1.01 * 2.0 * (cstate.m_fx - prevf) / prevdg);
        }
        prevf  = cstate.m_fx;
        prevdg = cstate.dg(cdescent);

        const auto [iter_ok, step_size] = lsearchk.get(cstate, cdescent, t0);
        last_step_size                  = step_size;
        ++iterations;

        cstate.m_track.update(cstate.m_x, cstate.m_fx);
        if (!(iter_ok && cstate.valid()))
        {
            status = STATUS_FAILED;
            break;
        }
        if (cstate.gradient_test() < epsilon)
        {
            status = STATUS_GRADIENT_TEST;
            break;
        }
    }
    return finish(cstate.valid() ? cstate : pstate);
}

// solver_osga_t::do_minimize (src/solver/osga.cpp:59-147). strong_convexity = function.strong_convexity().
int oracle_osga(callback_t callback, void* ctx, const int64_t n, double* x, const double epsilon,
                const double strong_convexity, const int64_t max_evals, const int64_t patience, double* fx,
                double* eta_out, int64_t* stats)
{
    constexpr double lambda = 0.9, alpha_max = 0.7, kappa_prime = 0.1, kappa = 1.1; // osga.cpp:49-51

    counted_function_t function{callback, ctx};
    const vec          z0(x, x + n);
    const double       miu = strong_convexity / 2.0;

    // prox-function Q(z) = Q0 + 0.5 |z - z0|^2 (osga.cpp:8-39)
    double norm2 = 0.0;
    for (const double v : z0)
    {
        norm2 += v * v;
    }
    const double Q0 = 0.5 * std::sqrt(std::sqrt(norm2) + std::numeric_limits<double>::epsilon());
    const auto   Q  = [&](const vec& z)
    {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i)
        {
            s += (z[i] - z0[i]) * (z[i] - z0[i]);
        }
        return Q0 + 0.5 * s;
    };
    const auto E = [&](const double gamma, const vec& h)
    {
        const double beta = gamma + dot(h, z0);
        const double hh   = dot(h, h);
        const double root = std::sqrt(beta * beta + 2.0 * Q0 * hh);
        return (beta <= 0.0) ? (-beta + root) / (2.0 * Q0) : hh / (beta + root);
    };
    const auto U = [&](const double gamma, const vec& h, vec& u)
    {
        const double e = E(gamma, h);
        for (int64_t i = 0; i < n; ++i)
        {
            u[i] = z0[i] - h[i] / e;
        }
    };

    state_t state(function, z0); // keeps track of the best state
    int     status     = STATUS_MAX_ITERS;
    int64_t iterations = 0;

    const size_t sn = static_cast<size_t>(n);
    vec          h(sn), u(sn), g = state.m_gx, xb = state.m_x, xv(sn), x_prime(sn), h_hat(sn), u_hat(sn), u_prime(sn);
    for (int64_t i = 0; i < n; ++i)
    {
        h[i] = state.m_gx[i] - miu * (state.m_x[i] - z0[i]);
    }
    double gamma = state.m_fx - miu * Q(state.m_x) - dot(h, state.m_x);
    U(gamma - state.m_fx, h, u);
    double eta   = E(gamma - state.m_fx, h) - miu;
    double alpha = alpha_max;
    double fb    = state.m_fx;

    while (function.calls() < max_evals)
    {
        if (norm_inf(state.m_gx) < kEpsilon0)
        {
            status = state.valid() ? STATUS_GRADIENT_TEST : STATUS_FAILED;
            break;
        }

        for (int64_t i = 0; i < n; ++i)
        {
            xv[i] = xb[i] + alpha * (u[i] - xb[i]);
        }
        const double f = function(xv, &g);
        for (int64_t i = 0; i < n; ++i)
        {
            g[i] -= miu * (xv[i] - z0[i]);
        }
        for (int64_t i = 0; i < n; ++i)
        {
            h_hat[i] = h[i] + alpha * (g[i] - h[i]);
        }
        const double gamma_hat = gamma + alpha * (f - miu * Q(xv) - dot(g, xv) - gamma);

        const vec&   xb_prime = (f < fb) ? xv : xb;
        const double fb_prime = (f < fb) ? f : fb;

        U(gamma_hat - fb_prime, h_hat, u_prime);
        for (int64_t i = 0; i < n; ++i)
        {
            x_prime[i] = xb[i] + alpha * (u_prime[i] - xb[i]);
        }
        const double f_prime = function(x_prime);

        const vec    xb_hat = (f_prime < fb_prime) ? x_prime : xb_prime;
        const double fb_hat = (f_prime < fb_prime) ? f_prime : fb_prime;

        state.update_if_better(xb_hat, fb_hat);

        U(gamma_hat - fb_hat, h_hat, u_hat);
        const double eta_hat = E(gamma_hat - fb_hat, h_hat) - miu;
        ++iterations;

        // done_specific_test || done_value_test (osga.cpp:123-128; src/solver.cpp:114-147,175-182)
        const bool iter_ok = state.valid();
        state.m_track.update(state.m_x, state.m_fx);
        if (!iter_ok)
        {
            status = STATUS_FAILED;
            break;
        }
        if (eta_hat < epsilon)
        {
            status = STATUS_SPECIFIC_TEST;
            eta    = eta_hat;
            break;
        }
        state.m_track.update(state.m_x, state.m_fx);
        if (state.m_track.value_test(patience) < epsilon)
        {
            status = STATUS_VALUE_TEST;
            eta    = eta_hat;
            break;
        }

        const double R = (eta - eta_hat) / (lambda * alpha * eta);
        alpha = (R < 1.0) ? (alpha * std::exp(-kappa)) : std::min(alpha * std::exp(kappa_prime * (R - 1.0)), alpha_max);
        if (eta_hat < eta)
        {
            h     = h_hat;
            u     = u_hat;
            eta   = eta_hat;
            gamma = gamma_hat;
        }
        xb = xb_hat;
        fb = fb_hat;
    }

    std::copy(state.m_x.begin(), state.m_x.end(), x);
    *fx      = state.m_fx;
    *eta_out = eta;
    stats[0] = function.m_fcalls;
    stats[1] = function.m_gcalls;
    stats[2] = iterations;
    return status;
}

} // extern "C"
