// kernels_t1_b.cu — group B of the single-target kernel variants (kernel_tables.cuh), compiled on its own.
#include "kernel_tables.cuh"

namespace nb200
{
namespace
{
const t1_variant kGroup[] = {NB200_T1_GROUP_B(NB200_T1_ENTRY)};
}

const t1_variant* t1_group_b(int* count)
{
    *count = static_cast<int>(sizeof(kGroup) / sizeof(kGroup[0]));
    return kGroup;
}
} // namespace nb200
