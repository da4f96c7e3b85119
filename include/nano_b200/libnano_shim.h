// nano_b200/libnano_shim.h — the drop-in for the REAL libnano: a nano::function_t whose do_eval is one nb200_vgrad.
//
// Only compiled where libnano's headers (and therefore Eigen3) are available; the rest of this repository does not
// need them. L-BFGS, CGD, OSGA, the bundle solvers and every line search drive it unchanged, because they only see
// nano::function_t::operator() (src/function.cpp:246-272): counters, size checks and the smooth/convex flags are
// kept by the base class. See INTEGRATION.md for how to wire it into linear_t::fit.
#pragma once

#if __has_include(<nano/function.h>) && __has_include(<nano/dataset/iterator.h>) && __has_include(<nano/loss.h>)

#include <nano/critical.h>
#include <nano/dataset/iterator.h>
#include <nano/function.h>
#include <nano/loss.h>

#include "../nb200.h"

#include <algorithm>
#include <vector>

namespace nano::cuda
{
inline int loss_kind(const loss_t& loss)
{
    // src/loss.cpp:105-135: the s-/m- prefix only changes error()
    auto id = loss.type_id();
    if (id.rfind("s-", 0) == 0 || id.rfind("m-", 0) == 0)
    {
        id = id.substr(2);
    }
    if (id == "mse") return NB200_LOSS_MSE;
    if (id == "mae") return NB200_LOSS_MAE;
    if (id == "cauchy") return NB200_LOSS_CAUCHY;
    if (id == "hinge") return NB200_LOSS_HINGE;
    if (id == "squared-hinge") return NB200_LOSS_SQUARED_HINGE;
    if (id == "classnll") return NB200_LOSS_CLASSNLL;
    if (id == "savage") return NB200_LOSS_SAVAGE;
    if (id == "tangent") return NB200_LOSS_TANGENT;
    if (id == "logistic") return NB200_LOSS_LOGISTIC;
    if (id == "exponential") return NB200_LOSS_EXPONENTIAL;
    if (id == "pinball") return NB200_LOSS_PINBALL;
    critical(false, "cuda: unsupported loss <", loss.type_id(), ">!");
    return -1;
}

///
/// \brief the ERM objective of a linear model (see nano::linear::function_t) evaluated on a B200.
///
/// The (already scaled) inputs and targets are gathered ONCE through flatten_iterator_t::loop — exactly like
/// check_vgrad does in test/test_linear_function.cpp:29-38 — and pinned in HBM.
///
class linear_function_t final : public ::nano::function_t
{
public:
    /// \param devices the GPUs to shard the samples over (empty: every visible device, i.e. the whole box, like
    ///        libnano's pool uses every core); the rows are split in contiguous blocks, one per device.
    linear_function_t(const flatten_iterator_t& iterator, const loss_t& loss, scalar_t l1reg, scalar_t l2reg,
                      std::vector<int> devices = {})
        : ::nano::function_t("linear", (iterator.dataset().columns() + 1) * ::nano::size(iterator.dataset().target_dims()))
        , m_isize(iterator.dataset().columns())
        , m_tsize(::nano::size(iterator.dataset().target_dims()))
    {
        const auto samples = iterator.samples().size();

        tensor2d_t inputs(samples, m_isize);
        tensor2d_t targets(samples, m_tsize);
        iterator.loop(
            [&](tensor_range_t range, size_t, tensor2d_cmap_t input, tensor4d_cmap_t target)
            {
                inputs.slice(range)                                      = input;
                targets.slice(range).reshape(range.size() * m_tsize) = target.reshape(range.size() * m_tsize);
            });

        if (devices.empty())
        {
            for (int device = 0; device < nb200_device_count(); ++device)
            {
                devices.push_back(device);
            }
        }
        // (never more devices than samples: every device needs at least one row)
        devices.resize(std::min(devices.size(), static_cast<size_t>(std::max<tensor_size_t>(samples, 1))));
        nb200_ctx* ctx = nullptr;
        check(nb200_init_multi(static_cast<int>(devices.size()), devices.data(), &ctx));
        m_ctx.reset(ctx, nb200_ctx_release);

        nb200_data* data = nullptr;
        check(nb200_dataset_pin(ctx, inputs.data(), targets.data(), samples, m_isize, m_tsize, &data));
        m_data.reset(data, nb200_dataset_release);

        const auto alpha = (loss.type_id() == "pinball") ? loss.parameter("loss::pinball::alpha").value<scalar_t>() : 0.5;
        m_loss_kind      = loss_kind(loss);
        m_loss_param     = alpha;
        m_l1reg          = l1reg;
        m_l2reg          = l2reg;
        make_objective();

        convex(loss.convex() ? convexity::yes : convexity::no);                          // src/linear/function.cpp:34
        smooth((loss.smooth() && l1reg <= 0.0) ? smoothness::yes : smoothness::no);       // :35
        strong_convexity(l2reg / static_cast<scalar_t>(m_isize * m_tsize));               // :36
    }

    linear_function_t(const linear_function_t& other)
        : ::nano::function_t(other)
        , m_isize(other.m_isize)
        , m_tsize(other.m_tsize)
        , m_ctx(other.m_ctx)
        , m_data(other.m_data) // clones share the pinned dataset (ref-counted) ...
        , m_loss_kind(other.m_loss_kind)
        , m_loss_param(other.m_loss_param)
        , m_l1reg(other.m_l1reg)
        , m_l2reg(other.m_l2reg)
    {
        make_objective(); // ... and own their stream + scratch (one thread per function object)
    }

    rfunction_t clone() const override { return std::make_unique<linear_function_t>(*this); }

    scalar_t do_eval(eval_t eval) const override
    {
        // the Hessian branch (src/linear/function.cpp:73-121) is outside the GPU hot path
        critical(!eval.has_hess(), "cuda: the Hessian of the linear objective is not available on the GPU path!");

        scalar_t fx = 0.0;
        check(nb200_vgrad(m_obj.get(), eval.m_x.data(), eval.m_x.size(), eval.has_grad() ? eval.m_gx.data() : nullptr,
                          &fx));
        return fx; // non-finite values flow back as they are: solver_state_t::valid() decides
    }

private:
    static void check(const int status) { critical(status == NB200_OK, "cuda: ", nb200_last_error()); }

    void make_objective()
    {
        nb200_obj* obj = nullptr;
        check(nb200_objective_create(m_data.get(), m_loss_kind, m_loss_param, NB200_FLAVOUR_LINEAR, m_l1reg, m_l2reg,
                                     nullptr, &obj));
        m_obj.reset(obj, nb200_objective_release);
    }

    tensor_size_t               m_isize{0};
    tensor_size_t               m_tsize{0};
    std::shared_ptr<nb200_ctx>  m_ctx;
    std::shared_ptr<nb200_data> m_data;
    std::shared_ptr<nb200_obj>  m_obj;
    int                         m_loss_kind{0};
    scalar_t                    m_loss_param{0.5};
    scalar_t                    m_l1reg{0.0};
    scalar_t                    m_l2reg{0.0};
};
} // namespace nano::cuda

#endif // libnano headers available
