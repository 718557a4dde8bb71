// Explicit instantiations of the tcgen05 GEMM launcher, group 7 (see DX_TC_GROUP_7 in dx_gemm_tc_impl.cuh).
#include "dx_gemm_tc_impl.cuh"

namespace dx_tc {
DX_TC_GROUP_7(DX_TC_INSTANTIATE)
}
