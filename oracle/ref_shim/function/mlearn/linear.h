// oracle/ref_shim/function/mlearn/linear.h — TEST INFRASTRUCTURE. Shadows the reference header of the same name (which
// pulls nano/function.h and, through it, Eigen3 proper) so that the reference's OWN src/function/mlearn/loss.{h,cpp} —
// the five losses of the builtin linear test functions, objective (B) of SURVEY.md — compile unmodified, in place, into
// oracle/_ref/libref_loss.so. Provides the matrix / array / tensor-map vocabulary those two files use, eagerly
// evaluated, sums left to right. The Hessian members only have to compile: they are never called.
#pragma once

#include <nano/tensor/eigen.h> // the stand-in next to this file

#include <utility>

#ifndef NANO_PUBLIC
    #define NANO_PUBLIC
#endif

namespace nano
{
namespace refshim
{
class marray_t; // element-wise view of a matrix (Eigen's .array())

struct diagonal_t
{
    array_t m_values;
};

class matrix_val_t
{
public:
    matrix_val_t() = default;
    matrix_val_t(const tensor_size_t rows, const tensor_size_t cols, array_t data)
        : m_rows(rows)
        , m_cols(cols)
        , m_data(std::move(data))
    {
    }

    tensor_size_t rows() const { return m_rows; }
    tensor_size_t cols() const { return m_cols; }
    std::pair<tensor_size_t, tensor_size_t> dims() const { return {m_rows, m_cols}; }
    const array_t& data() const { return m_data; }

    scalar_t squaredNorm() const { return m_data.square().sum(); }
    marray_t array() const;
    array_t  array(const tensor_size_t row) const // one sample
    {
        array_t out(m_cols);
        for (tensor_size_t c = 0; c < m_cols; ++c)
        {
            out(c) = m_data(row * m_cols + c);
        }
        return out;
    }
    diagonal_t asDiagonal() const { return {m_data}; }

private:
    tensor_size_t m_rows{0}, m_cols{0};
    array_t       m_data;
};

inline matrix_val_t operator-(const matrix_val_t& a, const matrix_val_t& b)
{
    return {a.rows(), a.cols(), a.data() - b.data()};
}

// Eigen array expression over a whole matrix: element-wise operations, .matrix() goes back
class marray_t
{
public:
    marray_t(const tensor_size_t rows, const tensor_size_t cols, array_t data)
        : m_rows(rows)
        , m_cols(cols)
        , m_data(std::move(data))
    {
    }
    marray_t with(array_t data) const { return {m_rows, m_cols, std::move(data)}; }
    const array_t& data() const { return m_data; }

    marray_t abs() const { return with(m_data.abs()); }
    marray_t sign() const { return with(m_data.sign()); }
    marray_t square() const { return with(m_data.square()); }
    marray_t log() const { return with(m_data.log()); }
    marray_t exp() const { return with(m_data.exp()); }
    marray_t max(const scalar_t v) const { return with(m_data.max(v)); }
    scalar_t sum() const { return m_data.sum(); }
    marray_t operator-() const { return with(-m_data); }
    matrix_val_t matrix() const { return {m_rows, m_cols, m_data}; }

private:
    tensor_size_t m_rows, m_cols;
    array_t       m_data;
};

inline marray_t matrix_val_t::array() const
{
    return {m_rows, m_cols, m_data};
}

inline marray_t operator*(const marray_t& a, const marray_t& b) { return a.with(a.data() * b.data()); }
inline marray_t operator/(const marray_t& a, const marray_t& b) { return a.with(a.data() / b.data()); }
inline marray_t operator+(const scalar_t s, const marray_t& a) { return a.with(s + a.data()); }
inline marray_t operator+(const marray_t& a, const scalar_t s) { return a.with(a.data() + s); }
inline marray_t operator*(const scalar_t s, const marray_t& a) { return a.with(s * a.data()); }
inline marray_t operator*(const marray_t& a, const scalar_t s) { return a.with(a.data() * s); }

// matrix_map_t: the gradient output argument, passed by value and written through
class matrix_out_t
{
public:
    explicit matrix_out_t(matrix_val_t& target)
        : m_target(&target)
    {
    }
    matrix_out_t& operator=(const matrix_val_t& value)
    {
        *m_target = value;
        return *this;
    }

private:
    matrix_val_t* m_target;
};

// tensor3d_map_t of the Hessian members: compiles, never used
struct tensor_slot_t
{
    template <class tvalue>
    tensor_slot_t& operator=(const tvalue&)
    {
        return *this;
    }
};

struct tensor3d_out_t
{
    tensor_slot_t tensor(tensor_size_t) { return {}; }
    tensor_slot_t array(tensor_size_t) { return {}; }
    void          full(scalar_t) {}
    scalar_t&     operator()(tensor_size_t, tensor_size_t, tensor_size_t) { return m_dummy; }
    scalar_t      m_dummy{0.0};
};
} // namespace refshim

using matrix_cmap_t  = const refshim::matrix_val_t&;
using matrix_map_t   = refshim::matrix_out_t;
using tensor3d_map_t = refshim::tensor3d_out_t;
} // namespace nano
