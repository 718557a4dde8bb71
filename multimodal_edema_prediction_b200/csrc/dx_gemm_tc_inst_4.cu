// Explicit instantiations of the tcgen05 GEMM launcher, group 4 (see DX_TC_GROUP_4 in dx_gemm_tc_impl.cuh).
#include "dx_gemm_tc_impl.cuh"

namespace dx_tc {
DX_TC_GROUP_4(DX_TC_INSTANTIATE)
}
