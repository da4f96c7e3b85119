// oracle/ref_shim/nano/tensor.h — TEST INFRASTRUCTURE. include/nano/loss/error.h includes <nano/tensor.h> (the whole
// tensor library, i.e. Eigen3, which is not installed here); it only needs the array slice of the stand-in next door,
// plus is_pos_target from include/nano/loss/class.h:26-29 — that header declares tensor-typed functions as well and
// cannot be included, so its one-line predicate is restated here.
#pragma once

#include <nano/tensor/eigen.h>

namespace nano
{
inline bool is_pos_target(const scalar_t target)
{
    return target > 0;
}
} // namespace nano
