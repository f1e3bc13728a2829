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

// ---- GaussianSmearing inside the edge kernels (reference process.py:580-590: exp(coeff (d - mu_k)^2), mu = linspace)
// Eight consecutive basis values t_k .. t_{k+7} of one edge from two exponentials: with uniform spacing dmu,
//   t_{k+1} = t_k rho_k,  rho_k = 2^(c2 dmu (dmu - 2 (d - mu_k))),  rho_{k+1} = rho_k q2,  q2 = 2^(2 c2 dmu^2)
// (c2 = coeff log2 e).  The chunk restarts from the module's own mu_k0 (table `mu`), so position errors do not
// accumulate beyond seven steps; against torch.exp(coeff * (d - mu)^2) in fp64, in a float32 restatement with exact
// exponentials (tests/test_data_lazy_cpu.py): 2.2e-6 of the basis' scale (1) at the reference's parameters (G = 50),
// up to 3.8e-6 for a coarse basis (G = 8).
struct SmearConst { float c2, dmu, q2; };
__device__ __forceinline__ SmearConst smear_const(const float* mu, int G, float coeff) {
  SmearConst s;
  s.c2 = coeff * kLog2e;
  s.dmu = (G > 1) ? (mu[G - 1] - mu[0]) / (float)(G - 1) : 0.0f;
  s.q2 = ex2_(2.0f * s.c2 * s.dmu * s.dmu);
  return s;
}
__device__ __forceinline__ void smear_chunk8(float d, const float* mu, int k0, int G, const SmearConst& s, float (&v)[8]) {
  const float diff = d - mu[k0];
  float t = ex2_(s.c2 * (diff * diff));
  float rho = ex2_(s.c2 * s.dmu * (s.dmu - 2.0f * diff));
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = (k0 + j < G) ? t : 0.0f;
    t *= rho;
    rho *= s.q2;
  }
}

// ---- packed fp32 pairs (one FMA-pipe instruction per two values on sm_100)
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk2(float a, float b) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) {
  f2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) {
  f2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) {
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// sigmoid(a_f) * softplus(a_s) / ln 2 for two channels, from yf = -log2(e) a_f and ys = +log2(e) a_s:
//   1 / (1 + 2^yf)  *  (max(ys, 0) + log2(1 + 2^-|ys|))
// three MUFU ops per channel (ex2, ex2, lg2); the reciprocal runs on the FMA pipe as in rcp_fma (bit-trick seed,
// two third-order steps), on packed pairs.
__device__ __forceinline__ f2_t gate_pair(float yf0, float yf1, float ys0, float ys1) {
  const f2_t one = pk2(1.0f, 1.0f);
  const f2_t u = add2(pk2(ex2_(fminf(yf0, 126.0f)), ex2_(fminf(yf1, 126.0f))), one);
  float u0, u1;
  upk2(u, u0, u1);
  f2_t r = pk2(__uint_as_float(0x7EF311C7u - __float_as_uint(u0)), __uint_as_float(0x7EF311C7u - __float_as_uint(u1)));
  const f2_t nu = pk2(-u0, -u1);
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const f2_t e = fma2(nu, r, one);
    r = fma2(r, fma2(e, e, e), r);
  }
  const f2_t w = add2(pk2(ex2_(-fabsf(ys0)), ex2_(-fabsf(ys1))), one);
  float w0, w1;
  upk2(w, w0, w1);
  const f2_t sp = add2(pk2(lg2_(w0), lg2_(w1)), pk2(fmaxf(ys0, 0.0f), fmaxf(ys1, 0.0f)));
  return mul2(r, sp);
}


// Backward counterpart: d m / d a_f and d m / d a_s times g (= grad_out[dst], already divided by the degree) for two
// channels, from the same base-2 pre-activations:
//   sg = sigmoid(a_f) = 1 / (1 + 2^yf);  t = 2^-|ys|;  sp = ln2 (max(ys, 0) + log2(1 + t));  sgs = sigmoid(a_s) = (ys >= 0 ? 1 : t) / (1 + t)
//   r0 = g sp sg (1 - sg),  r1 = g sg sgs
// four MUFU ops per channel (ex2, ex2, lg2, rcp); sigmoid(a_f)'s reciprocal on the FMA pipe, packed.
__device__ __forceinline__ void gate_pair_bwd(float yf0, float yf1, float ys0, float ys1, float g0, float g1, f2_t& r0, f2_t& r1) {
  const f2_t one = pk2(1.0f, 1.0f);
  const f2_t u = add2(pk2(ex2_(fminf(yf0, 126.0f)), ex2_(fminf(yf1, 126.0f))), one);
  float u0, u1;
  upk2(u, u0, u1);
  f2_t sg = pk2(__uint_as_float(0x7EF311C7u - __float_as_uint(u0)), __uint_as_float(0x7EF311C7u - __float_as_uint(u1)));
  const f2_t nu = pk2(-u0, -u1);
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const f2_t e = fma2(nu, sg, one);
    sg = fma2(sg, fma2(e, e, e), sg);
  }
  const float t0 = ex2_(-fabsf(ys0)), t1 = ex2_(-fabsf(ys1));
  const f2_t w = add2(pk2(t0, t1), one);
  float w0, w1;
  upk2(w, w0, w1);
  const f2_t sp = mul2(pk2(kLn2, kLn2), add2(pk2(lg2_(w0), lg2_(w1)), pk2(fmaxf(ys0, 0.0f), fmaxf(ys1, 0.0f))));
  const f2_t sgs = mul2(pk2(ys0 >= 0.0f ? 1.0f : t0, ys1 >= 0.0f ? 1.0f : t1), pk2(rcp_(w0), rcp_(w1)));
  const f2_t gsg = mul2(pk2(g0, g1), sg);
  r1 = mul2(gsg, sgs);
  r0 = mul2(mul2(gsg, fma2(sg, pk2(-1.0f, -1.0f), one)), sp);
}

}  // namespace mdl
