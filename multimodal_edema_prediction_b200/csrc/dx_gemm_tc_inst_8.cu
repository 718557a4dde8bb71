// Explicit instantiations of the tcgen05 GEMM launcher, group 8: fp32 operands on kind::tf32 (see DX_TC_GROUP_8).
#include "dx_gemm_tc_impl.cuh"

namespace dx_tc {
DX_TC_GROUP_8(DX_TC32_INSTANTIATE)
}
