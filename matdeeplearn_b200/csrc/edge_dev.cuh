// edge_dev.cuh -- device helpers shared by the fused edge kernels (cgconv_tc.cu, cgconv_tt.cu)
#pragma once
#include "umma.cuh"

namespace mdl {

struct TileInfo { int n_lo, n_hi, e_lo, e_hi; };

__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(umma::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(umma::smem_u32(dst)), "l"(src) : "memory");
}
// 16 bytes, L2 only (streamed node rows are re-read from shared memory, never from L1)
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(umma::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- gate math on the MUFU pipe (ex2 / lg2 / rcp approximations, abs. error ~2e-7) ----
__device__ __forceinline__ float ex2_(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float sigmoid_mufu(float x) { return rcp_(1.0f + ex2_(-kLog2e * x)); }
// softplus(x) = max(x,0) + log1p(exp(-|x|))   (== F.softplus incl. its x>20 branch to fp32 rounding)
__device__ __forceinline__ float softplus_mufu(float x) {
  return fmaf(kLn2, lg2_(1.0f + ex2_(-kLog2e * fabsf(x))), fmaxf(x, 0.0f));
}

// 1/u for u in [1, 2^126] on the FMA / integer pipes: bit-trick seed (|1 - u r0| < 0.15) and two
// third-order steps r <- r (1 + e + e^2), e = 1 - u r  (error 0.15 -> 3.4e-3 -> 4e-8).  Eight
// instructions next to one MUFU op (8 issue cycles per warp on its 4-lane pipe): used to move
// part of the gate math off the transcendental pipe, which otherwise bounds the epilogue.
__device__ __forceinline__ float rcp_fma(float u) {
  float r = __uint_as_float(0x7EF311C7u - __float_as_uint(u));
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const float e = fmaf(-u, r, 1.0f);
    r = fmaf(r, fmaf(e, e, e), r);
  }
  return r;
}
// sigmoid with the reciprocal on the FMA pipe; the exponent is clamped so that 1 + 2^z stays finite
__device__ __forceinline__ float sigmoid_mixed(float x) {
  return rcp_fma(1.0f + ex2_(fminf(-kLog2e * x, 126.0f)));
}
// softplus(x) and sigmoid(x) of the SAME argument from one exponential: with t = exp(-|x|),
// softplus = max(x,0) + log(1+t) and sigmoid = 1/(1+t) (x >= 0) or t/(1+t) (x < 0): 3 MUFU ops, not 4.
__device__ __forceinline__ void softplus_sigmoid_mufu(float x, float& sp, float& sg) {
  const float t = ex2_(-kLog2e * fabsf(x));
  const float u = 1.0f + t;
  sp = fmaf(kLn2, lg2_(u), fmaxf(x, 0.0f));
  const float r = rcp_(u);
  sg = x >= 0.0f ? r : t * r;
}

}  // namespace mdl
