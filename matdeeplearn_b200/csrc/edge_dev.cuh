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


}  // namespace mdl
