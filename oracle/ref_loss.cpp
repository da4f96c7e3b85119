// oracle/ref_loss.cpp — TEST INFRASTRUCTURE: C entry points over the reference's OWN loss headers, compiled in place from
// /root/reference (make -C oracle ref -> oracle/_ref/libref_loss.so; the stand-in for <Eigen/Core> is oracle/ref_shim/).
// kind follows oracle/__init__.py LOSS_KINDS: 0 mse 1 mae 2 cauchy 3 hinge 4 squared-hinge 5 classnll 6 savage 7 tangent
// 8 logistic 9 exponential (pinball lives in src/loss/pinball.cpp behind tensor maps: its golden vectors are the
// reference's test/test_loss.cpp:271-307).
#include <nano/loss/error.h> // over ref_shim/nano/tensor.h

#include <nano/loss/cauchy.h>
#include <nano/loss/classnll.h>
#include <nano/loss/exponential.h>
#include <nano/loss/hinge.h>
#include <nano/loss/logistic.h>
#include <nano/loss/mae.h>
#include <nano/loss/mse.h>
#include <nano/loss/savage.h>
#include <nano/loss/squared_hinge.h>
#include <nano/loss/tangent.h>

// objective (B): the reference's src/function/mlearn/loss.{h,cpp}, unmodified; <function/mlearn/linear.h> resolves to
// the stand-in in oracle/ref_shim/ (the real one needs Eigen3 proper)
#include <function/mlearn/loss.h>

#include "src/function/mlearn/loss.cpp" // -I /root/reference: compiled in place

// the random stream of the builtin test functions (linear_model_t, src/function/mlearn/linear.cpp:15-20): the
// reference's own make_rng / make_udist (std-only sources, compiled in place)
#include <nano/core/random.h>

#include "src/core/random.cpp"

#include <cstdint>

namespace
{
struct no_error_t // the terror base class only contributes error(), which is not on the vgrad path
{
};
using array_t     = nano::refshim::array_t;
using array_ref_t = nano::refshim::array_ref_t;

template <template <class> class tloss>
double value(const array_t& target, const array_t& output)
{
    return tloss<no_error_t>::value(target, output);
}

template <template <class> class tloss>
void vgrad(const array_t& target, const array_t& output, array_t& out)
{
    tloss<no_error_t>::vgrad(target, output, array_ref_t{out});
}
} // namespace

extern "C" {
// values[N]; targets / outputs [N, T] row-major (one sample = one array of T, as flatten_loss_t::do_value loops)
int ref_loss_value(const int kind, const double* targets, const double* outputs, const int64_t N, const int64_t T,
                   double* values)
{
    using namespace nano::detail;
    for (int64_t i = 0; i < N; ++i)
    {
        const array_t t(targets + i * T, T), o(outputs + i * T, T);
        switch (kind)
        {
        case 0: values[i] = value<mse_t>(t, o); break;
        case 1: values[i] = value<mae_t>(t, o); break;
        case 2: values[i] = value<cauchy_t>(t, o); break;
        case 3: values[i] = value<hinge_t>(t, o); break;
        case 4: values[i] = value<squared_hinge_t>(t, o); break;
        case 5: values[i] = value<classnll_t>(t, o); break;
        case 6: values[i] = value<savage_t>(t, o); break;
        case 7: values[i] = value<tangent_t>(t, o); break;
        case 8: values[i] = value<logistic_t>(t, o); break;
        case 9: values[i] = value<exponential_t>(t, o); break;
        default: return 1;
        }
    }
    return 0;
}

int ref_loss_vgrad(const int kind, const double* targets, const double* outputs, const int64_t N, const int64_t T,
                   double* vgrads)
{
    using namespace nano::detail;
    for (int64_t i = 0; i < N; ++i)
    {
        const array_t t(targets + i * T, T), o(outputs + i * T, T);
        array_t       g(T);
        switch (kind)
        {
        case 0: vgrad<mse_t>(t, o, g); break;
        case 1: vgrad<mae_t>(t, o, g); break;
        case 2: vgrad<cauchy_t>(t, o, g); break;
        case 3: vgrad<hinge_t>(t, o, g); break;
        case 4: vgrad<squared_hinge_t>(t, o, g); break;
        case 5: vgrad<classnll_t>(t, o, g); break;
        case 6: vgrad<savage_t>(t, o, g); break;
        case 7: vgrad<tangent_t>(t, o, g); break;
        case 8: vgrad<logistic_t>(t, o, g); break;
        case 9: vgrad<exponential_t>(t, o, g); break;
        default: return 1;
        }
        for (int64_t k = 0; k < T; ++k)
        {
            vgrads[i * T + k] = g(k);
        }
    }
    return 0;
}
}

// loss_t::error per sample through the reference's own error functors (include/nano/loss/error.h:11-70);
// error_kind: 0 absdiff (regression), 1 sclass (s- ids), 2 mclass (m- ids)
extern "C" int ref_loss_error(const int error_kind, const double* targets, const double* outputs, const int64_t N,
                              const int64_t T, double* errors)
{
    using namespace nano::loss::detail;
    for (int64_t i = 0; i < N; ++i)
    {
        const array_t t(targets + i * T, T), o(outputs + i * T, T);
        switch (error_kind)
        {
        case 0: errors[i] = absdiff_t::error(t, o); break;
        case 1: errors[i] = sclass_t::error(t, o); break;
        case 2: errors[i] = mclass_t::error(t, o); break;
        default: return 1;
        }
    }
    return 0;
}

// objective (B) losses, kind: 0 mse 1 mae 2 cauchy 3 hinge 4 logistic (oracle/__init__.py SYNTH_LOSSES).
// outputs / targets [N, T] row-major -> fx (already divided by N, src/function/mlearn/loss.cpp) and gx [N, T]
extern "C" int ref_synth_loss(const int kind, const double* outputs, const double* targets, const int64_t N, const int64_t T,
                              double* fx, double* gx)
{
    using nano::refshim::array_t;
    using nano::refshim::matrix_out_t;
    using nano::refshim::matrix_val_t;
    const matrix_val_t O(N, T, array_t(outputs, N * T)), Y(N, T, array_t(targets, N * T));
    matrix_val_t       G(N, T, array_t(N * T));
    switch (kind)
    {
    case 0: *fx = nano::loss_mse_t::fx(O, Y); nano::loss_mse_t::gx(O, Y, matrix_out_t{G}); break;
    case 1: *fx = nano::loss_mae_t::fx(O, Y); nano::loss_mae_t::gx(O, Y, matrix_out_t{G}); break;
    case 2: *fx = nano::loss_cauchy_t::fx(O, Y); nano::loss_cauchy_t::gx(O, Y, matrix_out_t{G}); break;
    case 3: *fx = nano::loss_hinge_t::fx(O, Y); nano::loss_hinge_t::gx(O, Y, matrix_out_t{G}); break;
    case 4: *fx = nano::loss_logistic_t::fx(O, Y); nano::loss_logistic_t::gx(O, Y, matrix_out_t{G}); break;
    default: return 1;
    }
    for (int64_t i = 0; i < N * T; ++i)
    {
        gx[i] = G.data()(i);
    }
    return 0;
}

// the first n draws of `auto rng = make_rng(seed); auto udist = make_udist<scalar_t>(0.0, 1.0);` (linear.cpp:15-16)
extern "C" int ref_rng_draws(const uint64_t seed, const int64_t n, double* out)
{
    auto rng   = nano::make_rng(seed);
    auto udist = nano::make_udist<nano::scalar_t>(0.0, 1.0);
    for (int64_t i = 0; i < n; ++i)
    {
        out[i] = udist(rng);
    }
    return 0;
}
