// oracle/ref_shim/nano/tensor/eigen.h — TEST INFRASTRUCTURE. A stand-in for the one header the reference's loss
// formulas include (/root/reference/include/nano/tensor/eigen.h -> <Eigen/Core>, which is not installed here), so that
// the reference's OWN loss headers (include/nano/loss/{mse,mae,cauchy,hinge,squared_hinge,savage,tangent,exponential,
// logistic,classnll}.h) compile unmodified, from where they lie, into oracle/_ref/libref_loss.so (oracle/Makefile,
// target `ref`). Only the slice of Eigen's array API those headers use is provided: eager, element-wise IEEE fp64
// operations in the written order; `.sum()` adds left to right (Eigen's packet reduction order is unspecified anyway).
// Nothing in the product or in the GPU tests reads /root/reference: the golden vectors this produces are committed
// under tests/golden/ together with the generating script.
#pragma once

#include <cassert>
#include <cmath>
#include <cstddef>
#include <limits>
#include <type_traits>
#include <vector>

namespace nano
{
using scalar_t      = double;         // include/nano/scalar.h:5
using tensor_size_t = std::ptrdiff_t; // include/nano/tensor/index.h:7 (Eigen::Index)

template <class tscalar>
tscalar square(const tscalar value) noexcept // include/nano/core/numeric.h:14
{
    return value * value;
}

// named by the (never instantiated) Hessian members of the loss headers: must merely be declared
struct matrix_t
{
    template <class... targs>
    static int identity(targs...)
    {
        return 0;
    }
};

namespace refshim
{
// value semantics; an output argument `tgarray vgrad` is passed as a view (array_ref_t) that writes through
class array_t
{
public:
    array_t() = default;
    explicit array_t(const tensor_size_t size, const scalar_t value = 0.0)
        : m_data(static_cast<size_t>(size), value)
    {
    }
    array_t(const scalar_t* data, const tensor_size_t size)
        : m_data(data, data + size)
    {
    }

    tensor_size_t size() const { return static_cast<tensor_size_t>(m_data.size()); }
    scalar_t      operator()(const tensor_size_t i) const { return m_data[static_cast<size_t>(i)]; }
    scalar_t&     operator()(const tensor_size_t i) { return m_data[static_cast<size_t>(i)]; }

    template <class top>
    array_t map(const top& op) const
    {
        array_t out(size());
        for (tensor_size_t i = 0; i < size(); ++i)
        {
            out(i) = op((*this)(i));
        }
        return out;
    }

    array_t square() const { return map([](const scalar_t v) { return v * v; }); }
    array_t abs() const { return map([](const scalar_t v) { return std::fabs(v); }); }
    array_t exp() const { return map([](const scalar_t v) { return std::exp(v); }); }
    array_t log() const { return map([](const scalar_t v) { return std::log(v); }); }
    array_t atan() const { return map([](const scalar_t v) { return std::atan(v); }); }
    // Eigen: sign(a) = (a > 0) - (a < 0); max(scalar): plain compare
    array_t sign() const { return map([](const scalar_t v) { return static_cast<scalar_t>((v > 0.0) - (v < 0.0)); }); }
    array_t max(const scalar_t other) const { return map([=](const scalar_t v) { return (v < other) ? other : v; }); }

    scalar_t sum() const
    {
        scalar_t total = 0.0;
        for (const auto v : m_data)
        {
            total += v;
        }
        return total;
    }
    scalar_t maxCoeff(tensor_size_t* index) const
    {
        tensor_size_t best = 0;
        for (tensor_size_t i = 1; i < size(); ++i)
        {
            if ((*this)(i) > (*this)(best))
            {
                best = i;
            }
        }
        *index = best;
        return (*this)(best);
    }

    // the slice include/nano/loss/error.h uses: `.array()` of an array, `(a < scalar).count()`
    const array_t& array() const { return *this; }
    struct mask_t
    {
        tensor_size_t m_count{0};
        tensor_size_t count() const { return m_count; }
    };
    mask_t operator<(const scalar_t other) const
    {
        mask_t mask;
        for (const auto v : m_data)
        {
            mask.m_count += (v < other) ? 1 : 0;
        }
        return mask;
    }

    // (Eigen: array -> vector -> diagonal matrix; only named by Hessian members that are never called)
    const array_t& matrix() const { return *this; }
    const array_t& asDiagonal() const { return *this; }

    array_t  operator-() const { return map([](const scalar_t v) { return -v; }); }
    array_t& operator/=(const scalar_t s)
    {
        for (auto& v : m_data) v /= s;
        return *this;
    }
    array_t& operator-=(const array_t& o)
    {
        for (tensor_size_t i = 0; i < size(); ++i) (*this)(i) -= o(i);
        return *this;
    }

private:
    std::vector<scalar_t> m_data;
};

template <class top>
array_t zip(const array_t& a, const array_t& b, const top& op)
{
    array_t out(a.size());
    for (tensor_size_t i = 0; i < a.size(); ++i)
    {
        out(i) = op(a(i), b(i));
    }
    return out;
}

inline array_t operator+(const array_t& a, const array_t& b) { return zip(a, b, [](scalar_t x, scalar_t y) { return x + y; }); }
inline array_t operator-(const array_t& a, const array_t& b) { return zip(a, b, [](scalar_t x, scalar_t y) { return x - y; }); }
inline array_t operator*(const array_t& a, const array_t& b) { return zip(a, b, [](scalar_t x, scalar_t y) { return x * y; }); }
inline array_t operator/(const array_t& a, const array_t& b) { return zip(a, b, [](scalar_t x, scalar_t y) { return x / y; }); }
inline array_t operator+(const array_t& a, const scalar_t s) { return a.map([=](scalar_t x) { return x + s; }); }
inline array_t operator+(const scalar_t s, const array_t& a) { return a.map([=](scalar_t x) { return s + x; }); }
inline array_t operator-(const array_t& a, const scalar_t s) { return a.map([=](scalar_t x) { return x - s; }); }
inline array_t operator-(const scalar_t s, const array_t& a) { return a.map([=](scalar_t x) { return s - x; }); }
inline array_t operator*(const array_t& a, const scalar_t s) { return a.map([=](scalar_t x) { return x * s; }); }
inline array_t operator*(const scalar_t s, const array_t& a) { return a.map([=](scalar_t x) { return s * x; }); }
inline array_t operator/(const array_t& a, const scalar_t s) { return a.map([=](scalar_t x) { return x / s; }); }
inline array_t operator/(const scalar_t s, const array_t& a) { return a.map([=](scalar_t x) { return s / x; }); }

// the by-value output argument of vgrad(target, output, tgarray vgrad): assignments write through to the caller
class array_ref_t
{
public:
    explicit array_ref_t(array_t& target)
        : m_target(&target)
    {
    }
    array_ref_t& operator=(const array_t& value)
    {
        *m_target = value;
        return *this;
    }
    array_ref_t& operator/=(const scalar_t s)
    {
        *m_target /= s;
        return *this;
    }
    array_ref_t& operator-=(const array_t& o)
    {
        *m_target -= o;
        return *this;
    }
    scalar_t& operator()(const tensor_size_t i) { return (*m_target)(i); }
    scalar_t  sum() const { return m_target->sum(); }
    operator const array_t&() const { return *m_target; }

private:
    array_t* m_target;
};
} // namespace refshim

template <class T>
inline constexpr bool is_eigen_v = std::is_same_v<std::remove_cv_t<T>, refshim::array_t> ||
                                   std::is_same_v<std::remove_cv_t<T>, refshim::array_ref_t>;
} // namespace nano
