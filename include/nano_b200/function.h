// nano_b200/function.h — header-only C++ host layer over the C ABI (nb200.h), Eigen-free.
//
// Mirrors the reference interface of the hot path with plain pointer + size views instead of Eigen maps, so that a
// C++ host that cannot (or does not want to) pull Eigen can drive the GPU objective exactly like libnano's solvers do:
//   nano::function_t                       include/nano/function.h:28-216, src/function.cpp:246-272
//   nano::linear::function_t               include/nano/linear/function.h:19-80, src/linear/function.cpp:21-168
//   function_{ridge,lasso,elasticnet}_t    src/function/mlearn/*.cpp
// Errors follow nano::critical (include/nano/critical.h:13-29): std::runtime_error("critical check failed: ...").
// The drop-in subclass of the REAL nano::function_t (needs Eigen + libnano headers) is in libnano_shim.h.
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../nb200.h"

namespace nano_b200
{
using scalar_t      = double;        // include/nano/scalar.h:5
using tensor_size_t = std::int64_t;  // include/nano/tensor/index.h:7

[[noreturn]] inline void raise(const std::string& message)
{
    throw std::runtime_error("critical check failed: " + message);
}

inline void critical(const bool condition, const std::string& message)
{
    if (!condition)
    {
        raise(message);
    }
}

inline void check(const int status)
{
    if (status != NB200_OK)
    {
        throw std::runtime_error(nb200_last_error());
    }
}

// contiguous fp64 views: what vector_cmap_t / vector_map_t are at the boundary (pointer + size)
struct vector_cmap_t
{
    const scalar_t* m_data{nullptr};
    tensor_size_t   m_size{0};
    const scalar_t* data() const { return m_data; }
    tensor_size_t   size() const { return m_size; }
};

struct vector_map_t
{
    scalar_t*     m_data{nullptr};
    tensor_size_t m_size{0};
    scalar_t*     data() const { return m_data; }
    tensor_size_t size() const { return m_size; }
};

class context_t
{
public:
    explicit context_t(const int device = 0) { check(nb200_init(device, &m_handle)); }
    // a device GROUP (nb200_init_multi): datasets pinned on it are split over the devices in contiguous row blocks and
    // objectives on them evaluate on every device from this one process (the thread pool's role in libnano)
    explicit context_t(const std::vector<int>& devices)
    {
        check(nb200_init_multi(static_cast<int>(devices.size()), devices.data(), &m_handle));
    }
    int devices() const { return nb200_ctx_devices(m_handle); }
    ~context_t() { nb200_ctx_release(m_handle); }
    context_t(const context_t&)            = delete;
    context_t& operator=(const context_t&) = delete;
    nb200_ctx* handle() const { return m_handle; }

private:
    nb200_ctx* m_handle{nullptr};
};

// what flatten_iterator_t::cache_flatten / cache_targets hold, pinned in HBM once per fit; shared by clones
class dataset_t
{
public:
    dataset_t(const context_t& ctx, const scalar_t* X, const scalar_t* Y, tensor_size_t samples, tensor_size_t inputs,
              tensor_size_t targets)
    {
        nb200_data* data = nullptr;
        check(nb200_dataset_pin(ctx.handle(), X, Y, samples, inputs, targets, &data));
        m_handle.reset(data, nb200_dataset_release);
    }

    nb200_data*   handle() const { return m_handle.get(); }
    tensor_size_t samples() const { return nb200_dataset_samples(handle()); }
    tensor_size_t inputs() const { return nb200_dataset_inputs(handle()); }
    tensor_size_t targets() const { return nb200_dataset_targets(handle()); }

private:
    std::shared_ptr<nb200_data> m_handle;
};

class function_t;
using rfunction_t = std::unique_ptr<function_t>;

// nano::function_t: size checks that throw, call counters, virtual do_eval (src/function.cpp:246-272)
class function_t
{
public:
    struct eval_t
    {
        bool          has_grad() const { return m_gx.size() == m_x.size(); }
        vector_cmap_t m_x{};
        vector_map_t  m_gx{};
    };

    function_t(std::string id, const tensor_size_t size)
        : m_id(std::move(id))
        , m_size(size)
    {
    }
    virtual ~function_t() = default;

    virtual rfunction_t clone() const = 0;

    const std::string& type_id() const { return m_id; }
    std::string        name(const bool with_size = true) const
    {
        return with_size ? (m_id + "[" + std::to_string(m_size) + "D]") : m_id;
    }
    tensor_size_t size() const { return m_size; }
    bool          convex() const { return m_convex; }
    bool          smooth() const { return m_smooth; }
    scalar_t      strong_convexity() const { return m_strong_convexity; }

    scalar_t operator()(const vector_cmap_t x, const vector_map_t gx = {}) const
    {
        critical(x.size() == size(), "function: invalid input size, expecting (" + std::to_string(size()) + ",), got (" +
                                         std::to_string(x.size()) + ",) instead!");
        critical(gx.size() == 0 || gx.size() == size(), "function: invalid gradient size, expecting (" +
                                                            std::to_string(size()) + ",) or empty, got (" +
                                                            std::to_string(gx.size()) + ",) instead!");
        m_fcalls += 1;
        m_gcalls += (gx.size() == size()) ? 1 : 0;
        return do_eval(eval_t{x, gx});
    }

    tensor_size_t fcalls() const { return m_fcalls; }
    tensor_size_t gcalls() const { return m_gcalls; }
    void          clear_statistics() const { m_fcalls = m_gcalls = 0; }
    // evaluations made on the device on this function's behalf (the device-resident solvers below)
    void          count_calls(const tensor_size_t fcalls, const tensor_size_t gcalls) const
    {
        m_fcalls += fcalls;
        m_gcalls += gcalls;
    }

protected:
    void convex(const bool value) { m_convex = value; }
    void smooth(const bool value) { m_smooth = value; }
    void strong_convexity(const scalar_t value) { m_strong_convexity = value; }

    virtual scalar_t do_eval(eval_t) const = 0;

private:
    std::string           m_id;
    tensor_size_t         m_size{0};
    bool                  m_convex{false};
    bool                  m_smooth{false};
    scalar_t              m_strong_convexity{0};
    mutable tensor_size_t m_fcalls{0};
    mutable tensor_size_t m_gcalls{0};
};

struct loss_traits_t
{
    int  m_kind;
    bool m_convex;
    bool m_smooth;
};

// loss ids of src/loss.cpp:105-135 -> device kind + the convex/smooth flags of include/nano/loss/*.h
inline loss_traits_t loss_traits(const std::string& id)
{
    const auto base = (id.rfind("s-", 0) == 0 || id.rfind("m-", 0) == 0) ? id.substr(2) : id;
    if (base == "mse") return {NB200_LOSS_MSE, true, true};
    if (base == "mae") return {NB200_LOSS_MAE, true, false};
    if (base == "cauchy") return {NB200_LOSS_CAUCHY, false, true};
    if (base == "hinge") return {NB200_LOSS_HINGE, true, false};
    if (base == "squared-hinge") return {NB200_LOSS_SQUARED_HINGE, true, false};
    if (base == "classnll") return {NB200_LOSS_CLASSNLL, true, true};
    if (base == "savage") return {NB200_LOSS_SAVAGE, false, true};
    if (base == "tangent") return {NB200_LOSS_TANGENT, false, true};
    if (base == "logistic") return {NB200_LOSS_LOGISTIC, true, true};
    if (base == "exponential") return {NB200_LOSS_EXPONENTIAL, true, true};
    if (base == "pinball") return {NB200_LOSS_PINBALL, true, false};
    raise("loss: unknown id <" + id + ">");
}

// the terror policy behind a loss id (include/nano/loss/flatten.h:103-122): "s-" / "m-" prefix, else regression
inline int error_kind(const std::string& id)
{
    return id.rfind("s-", 0) == 0 ? NB200_ERROR_SCLASS : id.rfind("m-", 0) == 0 ? NB200_ERROR_MCLASS : NB200_ERROR_ABSDIFF;
}

namespace linear
{
// linear::evaluate (src/linear/util.cpp:30-49) over a pinned dataset: per-sample errors and loss values of the
// predictions X W^T + b; the outputs stay in HBM. `errors` / `values` hold samples() doubles each.
inline void evaluate(const context_t& ctx, const dataset_t& dataset, const std::string& loss_id, const scalar_t* weights,
                     const scalar_t* bias, scalar_t* errors, scalar_t* values, const scalar_t loss_param = 0.5)
{
    check(nb200_evaluate(ctx.handle(), dataset.handle(), loss_traits(loss_id).m_kind, loss_param, error_kind(loss_id),
                         weights, bias, errors, values));
}

// nano::linear::function_t on the GPU: x = [W (T x D row-major) | b (T)]
class function_t final : public ::nano_b200::function_t
{
public:
    function_t(dataset_t dataset, const std::string& loss_id, const scalar_t l1reg, const scalar_t l2reg,
               const scalar_t loss_param = 0.5)
        : ::nano_b200::function_t("linear", (dataset.inputs() + 1) * dataset.targets())
        , m_dataset(std::move(dataset))
        , m_loss_id(loss_id)
        , m_l1reg(l1reg)
        , m_l2reg(l2reg)
        , m_loss_param(loss_param)
    {
        const auto traits = loss_traits(loss_id);
        nb200_obj* handle = nullptr;
        check(nb200_objective_create(m_dataset.handle(), traits.m_kind, loss_param, NB200_FLAVOUR_LINEAR, l1reg, l2reg,
                                     nullptr, &handle));
        m_handle.reset(handle);
        convex(traits.m_convex);                                   // src/linear/function.cpp:34
        smooth(traits.m_smooth && l1reg <= 0.0);                   // :35
        strong_convexity(l2reg / static_cast<scalar_t>(m_dataset.inputs() * m_dataset.targets())); // :36
    }

    rfunction_t clone() const override
    {
        // shares the pinned dataset, owns stream + scratch (src/linear/function.cpp:39-42; SURVEY.md §8b Ownership)
        return std::make_unique<function_t>(m_dataset, m_loss_id, m_l1reg, m_l2reg, m_loss_param);
    }

    nb200_obj* handle() const { return m_handle.get(); }

    // include/nano/linear/function.h:30-59
    const scalar_t* weights(const scalar_t* x) const { return x; }
    const scalar_t* bias(const scalar_t* x) const { return x + m_dataset.inputs() * m_dataset.targets(); }

protected:
    scalar_t do_eval(eval_t eval) const override
    {
        scalar_t fx = 0.0;
        check(nb200_vgrad(m_handle.get(), eval.m_x.data(), eval.m_x.size(), eval.has_grad() ? eval.m_gx.data() : nullptr,
                          &fx));
        return fx;
    }

private:
    struct release_t
    {
        void operator()(nb200_obj* obj) const { nb200_objective_release(obj); }
    };

    dataset_t                             m_dataset;
    std::string                           m_loss_id;
    scalar_t                              m_l1reg{0.0};
    scalar_t                              m_l2reg{0.0};
    scalar_t                              m_loss_param{0.5};
    std::unique_ptr<nb200_obj, release_t> m_handle;
};
} // namespace linear

// ---------------------------------------------------------------------------------------------------------------
// device-resident solvers: nano::solver_lbfgs_t / solver_cgd_*_t (src/solver/lbfgs.cpp, src/solver/cgd.cpp) with the
// vectors in HBM (nb200_lbfgs_minimize / nb200_cgd_minimize). The result mirrors solver_state_t's fields.
// ---------------------------------------------------------------------------------------------------------------
enum class solver_status
{
    max_iters     = 0, // include/nano/solver/status.h
    gradient_test = 1,
    specific_test = 2, // OSGA: eta < epsilon
    value_test    = 3, // OSGA: no improvement within `patience` iterations
    failed        = 4
};

struct solver_state_t
{
    std::vector<scalar_t> m_x;
    scalar_t              m_fx{0.0};
    scalar_t              m_gradient_test{0.0};
    solver_status         m_status{solver_status::max_iters};
    tensor_size_t         m_fcalls{0}, m_gcalls{0}, m_iterations{0};
};

namespace detail
{
template <typename tcall>
solver_state_t minimize(const linear::function_t& function, const std::vector<scalar_t>& x0, const tcall& call)
{
    critical(static_cast<tensor_size_t>(x0.size()) == function.size(),
             "solver: incompatible initial point (" + std::to_string(x0.size()) + " dimensions), expecting " +
                 std::to_string(function.size()) + " dimensions!");
    function.clear_statistics(); // solver_t::minimize, src/solver.cpp:107
    solver_state_t state;
    state.m_x       = x0;
    int64_t stats[3] = {0, 0, 0};
    int     status   = 0;
    check(call(function.handle(), state.m_x.data(), static_cast<int64_t>(state.m_x.size()), &state.m_fx,
               &state.m_gradient_test, stats, &status));
    state.m_status     = static_cast<solver_status>(status);
    state.m_fcalls     = stats[0];
    state.m_gcalls     = stats[1];
    state.m_iterations = stats[2];
    function.count_calls(stats[0], stats[1]);
    return state;
}
} // namespace detail

// defaults: src/solver.cpp:45-51 (epsilon 1e-8, max_evals 1000), src/solver/lbfgs.cpp:12 (history 50)
inline solver_state_t minimize_lbfgs(const linear::function_t& function, const std::vector<scalar_t>& x0,
                                     const scalar_t epsilon = 1e-8, const tensor_size_t max_evals = 1000,
                                     const tensor_size_t history = 50)
{
    return detail::minimize(function, x0,
                            [&](nb200_obj* obj, scalar_t* x, int64_t n, scalar_t* fx, scalar_t* gtest, int64_t* stats,
                                int* status) {
                                return nb200_lbfgs_minimize(obj, x, n, epsilon, max_evals, history, 0, nullptr, nullptr,
                                                            fx, gtest, stats, status);
                            });
}

// beta: NB200_CGD_HS ... NB200_CGD_FRPR (libnano's "cgd-hs" ... "cgd-frpr"; NB200_CGD_PR is its plain "cgd")
inline solver_state_t minimize_cgd(const linear::function_t& function, const std::vector<scalar_t>& x0,
                                   const int beta = NB200_CGD_PR, const scalar_t epsilon = 1e-8,
                                   const tensor_size_t max_evals = 1000, const scalar_t orthotest = 0.1,
                                   const scalar_t eta = 0.01)
{
    return detail::minimize(function, x0,
                            [&](nb200_obj* obj, scalar_t* x, int64_t n, scalar_t* fx, scalar_t* gtest, int64_t* stats,
                                int* status) {
                                return nb200_cgd_minimize(obj, x, n, beta, epsilon, max_evals, orthotest, eta, 0, nullptr,
                                                          nullptr, fx, gtest, stats, status);
                            });
}

// solver_osga_t::do_minimize (src/solver/osga.cpp:59-147), device-resident: the driver for non-smooth objectives
// (hinge / mae losses, lasso and elastic-net terms). m_gradient_test carries OSGA's eta on return.
// defaults: src/solver.cpp:45-51 (epsilon, max_evals, patience), osga.cpp:49-51 (lambda, alpha_max, kappas)
inline solver_state_t minimize_osga(const linear::function_t& function, const std::vector<scalar_t>& x0,
                                    const scalar_t epsilon = 1e-8, const tensor_size_t max_evals = 1000,
                                    const tensor_size_t patience = 100, const scalar_t lambda = 0.9,
                                    const scalar_t alpha_max = 0.7, const scalar_t kappa_prime = 0.1,
                                    const scalar_t kappa = 1.1)
{
    return detail::minimize(function, x0,
                            [&](nb200_obj* obj, scalar_t* x, int64_t n, scalar_t* fx, scalar_t* eta, int64_t* stats,
                                int* status) {
                                return nb200_osga_minimize(obj, x, n, epsilon, max_evals, patience, lambda, alpha_max,
                                                           kappa_prime, kappa, function.strong_convexity(), 0, fx, eta,
                                                           stats, status);
                            });
}
} // namespace nano_b200
