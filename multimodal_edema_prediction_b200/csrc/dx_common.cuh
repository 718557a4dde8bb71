// Shared device/host helpers for the duett_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define DX_OK 0
#define DX_ERR_ARG (-1)
#define DX_ERR_CUDA (-2)
#define DX_ERR_UNSUPPORTED (-3)

// dtype codes used across the C ABI (include/duett_b200.h)
#define DX_F32 0
#define DX_BF16 1

void dx_set_error(const char* fmt, ...);

#define DX_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      dx_set_error(__VA_ARGS__);           \
      return DX_ERR_ARG;                   \
    }                                      \
  } while (0)

#define DX_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      dx_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(e__), \
                   cudaGetErrorString(e__));                                            \
      return DX_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define DX_LAUNCH_CHECK() DX_CUDA(cudaGetLastError())

typedef __nv_bfloat16 bf16;

// ---- scalar/vector load-store as float, templated on storage type ----------------------
template <typename T> struct dx_type;
template <> struct dx_type<float> { static constexpr int code = DX_F32; };
template <> struct dx_type<bf16> { static constexpr int code = DX_BF16; };

__device__ __forceinline__ float dx_ld(const float* p) { return *p; }
__device__ __forceinline__ float dx_ld(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void dx_st(float* p, float v) { *p = v; }
__device__ __forceinline__ void dx_st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 8-element vector access (16 B for bf16, 32 B for f32). Pointers must be 16 B aligned.
__device__ __forceinline__ void dx_ld8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void dx_ld8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void dx_st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void dx_st8(bf16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ---- dropout ---------------------------------------------------------------------------
// Counter-based generator shared by every kernel that drops activations (attention probabilities, FFN hidden, heads):
// element `idx` of a dropout site is kept iff splitmix64(seed + idx * golden) >> 32 >= p * 2^32, and scaled by 1/(1-p).
// Nothing is stored: the backward pass regenerates the mask from (seed, idx).  tests/ops_emulator.py and the oracle restate
// the same function in numpy, so parity tests run both sides on identical masks.  (torch's Philox stream is not
// reproduced: the reference's dropout masks are not a portable contract.)
__host__ __device__ __forceinline__ uint32_t dx_rng32(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed + idx * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}
struct DxDrop {
  unsigned long long seed;
  uint32_t thresh;   // 0 = dropout disabled
  float scale;
};
static inline DxDrop dx_make_drop(float p, unsigned long long seed) {
  DxDrop d;
  d.seed = seed;
  double t = (double)p * 4294967296.0;
  d.thresh = p > 0.f ? (uint32_t)(t > 4294967295.0 ? 4294967295.0 : t) : 0u;
  d.scale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  return d;
}
// seed_dev: optional device-resident offset (advanced once per step so that a replayed CUDA graph draws fresh masks)
__device__ __forceinline__ DxDrop dx_drop_resolve(DxDrop d, const unsigned long long* seed_dev) {
  if (seed_dev) d.seed += seed_dev[0];
  return d;
}
__device__ __forceinline__ float dx_drop_factor(const DxDrop& d, unsigned long long idx) {
  return dx_rng32(d.seed, idx) >= d.thresh ? d.scale : 0.f;
}

// ---- reductions ----------------------------------------------------------------------
__device__ __forceinline__ float dx_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float dx_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum; every thread gets the result. `sh` must hold >= 33 floats.
__device__ __forceinline__ float dx_block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = dx_warp_sum(v);
  __syncthreads();  // protect sh reuse across consecutive calls
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : 0.f;
  r = dx_warp_sum(r);
  return r;
}

// ---- math ------------------------------------------------------------------------------
__device__ __forceinline__ float dx_gelu(float x) {  // exact erf GELU (torch nn.GELU default)
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float dx_gelu_grad(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float dx_sigmoid(float x) { return 1.f / (1.f + __expf(-x)); }
// softplus(x) = log(1+exp(x)), numerically stable
__device__ __forceinline__ float dx_softplus(float x) {
  return fmaxf(x, 0.f) + log1pf(__expf(-fabsf(x)));
}

static inline int dx_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
