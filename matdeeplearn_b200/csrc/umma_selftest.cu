// umma_selftest.cu -- D[128,N] = A[128,K] . B[N,K]^T on the tcgen05 tensor core,
// single CTA.  Exists so tests/test_gpu_umma.py can pin the descriptor / layout /
// TMEM-readback conventions of umma.cuh against a CPU matmul independently of the
// fused edge kernels that build on them.
#include "common.cuh"
#include "umma.cuh"
#include "../../include/mdl_b200_selftest.h"

namespace mdl {

__global__ void __launch_bounds__(128, 1)
k_umma_selftest(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                int N, int K, int split, int lbo_a, int sbo_a, int lbo_b, int sbo_b) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int a_bytes = 128 * K * 4, b_bytes = N * K * 4;
  uint8_t* sAhi = sm;
  uint8_t* sAlo = sAhi + a_bytes;
  uint8_t* sBhi = sAlo + a_bytes;
  uint8_t* sBlo = sBhi + b_bytes;
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, ncols);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < 128 * K; i += blockDim.x) {
    const int r = i / K, k = i - r * K;
    const float x = A[i];
    const float hi = split ? umma::tf32_hi(x) : x;
    const int off = umma::tile_offset_bytes(r, k, 128);
    *reinterpret_cast<float*>(sAhi + off) = hi;
    *reinterpret_cast<float*>(sAlo + off) = x - hi;
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int r = i / K, k = i - r * K;
    const float x = B[i];
    const float hi = split ? umma::tf32_hi(x) : x;
    const int off = umma::tile_offset_bytes(r, k, N);
    *reinterpret_cast<float*>(sBhi + off) = hi;
    *reinterpret_cast<float*>(sBlo + off) = x - hi;
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = umma::make_idesc_tf32(128, N);
    const uint32_t step_a = 2 * 128 * 16, step_b = 2 * N * 16;  // two k-chunks per MMA
    uint32_t acc = 0;
    for (int pass = 0; pass < (split ? 3 : 1); ++pass) {
      const uint8_t* a = (pass == 2) ? sAlo : sAhi;
      const uint8_t* b = (pass == 1) ? sBlo : sBhi;
      for (int kk = 0; kk < K / 8; ++kk) {
        const uint64_t ad = umma::make_desc(umma::smem_u32(a) + kk * step_a, lbo_a, sbo_a);
        const uint64_t bd = umma::make_desc(umma::smem_u32(b) + kk * step_b, lbo_b, sbo_b);
        umma::mma_tf32(tmem, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  const int row = tid;  // TMEM lane == D row for M=128, cta_group::1
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(umma::tmem_addr(tmem, warp, c0), v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < N) D[(size_t)row * N + c0 + j] = v[j];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}

// Same product with the A operand staged in TENSOR MEMORY (tcgen05.st by the thread that owns the
// row's lane) and B in shared memory: pins the A-from-TMEM convention the transposed edge kernel uses
// (lane = A row, one 32-bit column per tf32 element, 8 columns per MMA).
__global__ void __launch_bounds__(128, 1)
k_umma_selftest_ts(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
                   int split) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b_bytes = N * K * 4;
  uint8_t* sBhi = sm;
  uint8_t* sBlo = sBhi + b_bytes;
  uint32_t ncols = 32;
  while ((int)ncols < N + 2 * K) ncols <<= 1;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, ncols);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int r = i / K, k = i - r * K;
    const float x = B[i];
    const float hi = split ? umma::tf32_hi(x) : x;
    const int off = umma::tile_offset_bytes(r, k, N);
    *reinterpret_cast<float*>(sBhi + off) = hi;
    *reinterpret_cast<float*>(sBlo + off) = x - hi;
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t a_hi = tmem + N, a_lo = tmem + N + K;
  for (int k0 = 0; k0 < K; k0 += 8) {   // thread = lane = row of A
    float hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = A[(size_t)tid * K + k0 + j];
      hi[j] = split ? umma::tf32_hi(x) : x;
      lo[j] = x - hi[j];
    }
    umma::tmem_st8(umma::tmem_addr(a_hi, warp, k0), hi);
    umma::tmem_st8(umma::tmem_addr(a_lo, warp, k0), lo);
  }
  umma::tmem_st_wait();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = umma::make_idesc_tf32(128, N);
    const uint32_t step_b = 2 * N * 16;
    uint32_t acc = 0;
    for (int pass = 0; pass < (split ? 3 : 1); ++pass) {
      const uint32_t a = (pass == 2) ? a_lo : a_hi;
      const uint8_t* b = (pass == 1) ? sBlo : sBhi;
      for (int kk = 0; kk < K / 8; ++kk) {
        const uint64_t bd = umma::make_desc(umma::smem_u32(b) + kk * step_b, N * 16, 128);
        umma::mma_tf32_ts(tmem, a + kk * 8, bd, idesc, acc);
        acc = 1;
      }
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(umma::tmem_addr(tmem, warp, c0), v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < N) D[(size_t)tid * N + c0 + j] = v[j];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}


// Layout probe (development aid): the caller supplies the RAW shared-memory images of both operand tiles and
// the descriptor fields (LBO / SBO / layout type / major bits); nmma K = 8 MMAs are issued on them.  With
// index-coded tile contents and one-hot rows on the other side, D spells out which shared-memory word the
// tensor core reads for every (row, k).  Finding of round 2 (profiles/r2_umma_mn_probe.txt): kind::tf32 with
// an MN-major operand returns zeros for the no-swizzle, 32B, 64B and 128B layout types -- MN-major tf32
// needs SWIZZLE_128B_BASE32B -- so the K = E weight-gradient contractions stage their operands K-major.
__global__ void __launch_bounds__(128, 1)
k_umma_probe(const float* __restrict__ rawA, int a_floats, const float* __restrict__ rawB, int b_floats,
             float* __restrict__ D, int N, uint32_t lbo_a, uint32_t sbo_a, uint32_t lbo_b, uint32_t sbo_b,
             uint32_t idesc_extra, int nmma, uint32_t step_a, uint32_t step_b, uint32_t layout_a, uint32_t layout_b) {
  extern __shared__ __align__(128) uint8_t sm[];  // dynamic smem starts 1024-aligned on sm_100 (only static 16 B here)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  float* sA = reinterpret_cast<float*>(sm);
  float* sB = sA + ((a_floats + 255) & ~255);   // both tiles 1024-byte aligned (swizzled layouts)
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, ncols);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < a_floats; i += blockDim.x) sA[i] = rawA[i];
  for (int i = tid; i < b_floats; i += blockDim.x) sB[i] = rawB[i];
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = umma::make_idesc_tf32(128, N) | idesc_extra;
    for (int i = 0; i < nmma; ++i)
      umma::mma_tf32(tmem, umma::make_desc(umma::smem_u32(sA) + i * step_a, lbo_a, sbo_a) | ((uint64_t)layout_a << 61),
                     umma::make_desc(umma::smem_u32(sB) + i * step_b, lbo_b, sbo_b) | ((uint64_t)layout_b << 61), idesc,
                     i > 0);
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(umma::tmem_addr(tmem, warp, c0), v);
    umma::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < N) D[(size_t)tid * N + c0 + j] = v[j];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, ncols);
}


// Microbenchmark (development aid): cost of tcgen05.st for a warp, alone and while another warp keeps the tensor
// pipe busy with 128x128x8 tf32 MMAs (A from tensor memory).  out[warp] = cycles for `nstores` stores of `width`
// columns + the final tcgen05.wait::st; out[16] = cycles the MMA issuer spent.
__global__ void __launch_bounds__(576, 1)
k_tmem_st_bench(long long* __restrict__ out, int nwarps, int nstores, int width, int mma_count, int wait_each) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < 128 * 64; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.001f * (i & 15);
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 17) {  // MMA issuer: D = columns [0,128), A = columns [128,192) (whatever they hold), B = sm
    if (lane == 0 && mma_count > 0) {
      const long long t0 = clock64();
      const uint32_t idesc = umma::make_idesc_tf32(128, 128);
      for (int i = 0; i < mma_count; ++i)
        umma::mma_tf32_ts(tmem, tmem + 128 + 8 * (i & 7), umma::make_desc(umma::smem_u32(sm) + (i & 7) * 4096, 128 * 16, 128),
                          idesc, 1u);
      umma::mma_commit(&bar);
      out[16] = clock64() - t0;
      umma::mbar_wait(&bar, 0);
      out[17] = clock64() - t0;
    }
  } else if (warp < nwarps) {
    float v[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) v[t] = (float)(tid + t);
    const uint32_t base = umma::tmem_addr(tmem + 256, warp, 0);   // columns [256, 512)
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < nstores; ++i) {
      const uint32_t a = base + (uint32_t)((i * width) & 255);
      if (width == 8) { float w[8]; for (int t = 0; t < 8; ++t) w[t] = v[t] + i; umma::tmem_st8(a, w); }
      else if (width == 16) { float w[16]; for (int t = 0; t < 16; ++t) w[t] = v[t] + i; umma::tmem_st16(a, w); }
      else { float w[32]; for (int t = 0; t < 32; ++t) w[t] = v[t] + i; umma::tmem_st32(a & ~31u | (base & 31u), w); }
      if (wait_each) umma::tmem_st_wait();
    }
    umma::tmem_st_wait();
    const long long t1 = clock64();
    if (lane == 0) out[warp] = t1 - t0;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace mdl

using namespace mdl;

extern "C" int mdl_selftest_tmem_st_bench(long long* out, int32_t nwarps, int32_t nstores, int32_t width,
                                          int32_t mma_count, int32_t wait_each, void* stream) {
  MDL_REQUIRE(out && nwarps >= 1 && nwarps <= 16 && (width == 8 || width == 16 || width == 32), "tmem_st_bench: bad arguments");
  MDL_CUDA(cudaFuncSetAttribute(k_tmem_st_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  k_tmem_st_bench<<<1, 576, 32 * 1024, as_stream(stream)>>>(out, nwarps, nstores, width, mma_count, wait_each);
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_selftest_umma_probe(const float* rawA, int32_t a_floats, const float* rawB, int32_t b_floats,
                                       float* D, int32_t N, int32_t lbo_a, int32_t sbo_a, int32_t lbo_b,
                                       int32_t sbo_b, int32_t a_mn, int32_t b_mn, int32_t nmma, int32_t step_a,
                                       int32_t step_b, int32_t layout_a, int32_t layout_b, void* stream) {
  MDL_REQUIRE(rawA && rawB && D, "umma_probe: null pointer");
  MDL_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "umma_probe: N must be a multiple of 16 in [16,256]");
  const size_t smem = ((size_t)((a_floats + 255) & ~255) + b_floats) * 4;
  MDL_REQUIRE(smem <= 200 * 1024, "umma_probe: tiles too large");
  MDL_CUDA(cudaFuncSetAttribute(k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  k_umma_probe<<<1, 128, smem, as_stream(stream)>>>(rawA, a_floats, rawB, b_floats, D, N, lbo_a, sbo_a, lbo_b, sbo_b,
                                                     (a_mn ? 1u << 15 : 0u) | (b_mn ? 1u << 16 : 0u), nmma,
                                                     (uint32_t)step_a, (uint32_t)step_b, (uint32_t)layout_a,
                                                     (uint32_t)layout_b);
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_selftest_umma_ts(const float* A, const float* B, float* D, int32_t N, int32_t K,
                                    int32_t split, void* stream) {
  MDL_REQUIRE(A && B && D, "selftest_umma_ts: null pointer");
  MDL_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "selftest_umma_ts: N must be a multiple of 16 in [16,256]");
  MDL_REQUIRE(K >= 8 && K % 8 == 0 && N + 2 * K <= 512, "selftest_umma_ts: K must be a multiple of 8, N+2K <= 512");
  size_t smem = (size_t)2 * N * K * 4;
  MDL_REQUIRE(smem <= 200 * 1024, "selftest_umma_ts: tile too large");
  MDL_CUDA(cudaFuncSetAttribute(k_umma_selftest_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_umma_selftest_ts<<<1, 128, smem, as_stream(stream)>>>(A, B, D, N, K, split);
  MDL_LAUNCHED();
  return MDL_OK;
}

static int selftest_launch(const float* A, const float* B, float* D, int32_t N, int32_t K,
                           int32_t split, int lbo_a, int sbo_a, int lbo_b, int sbo_b, void* stream) {
  MDL_REQUIRE(A && B && D, "selftest_umma: null pointer");
  MDL_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "selftest_umma: N must be a multiple of 16 in [16,256]");
  MDL_REQUIRE(K >= 8 && K % 8 == 0, "selftest_umma: K must be a multiple of 8");
  size_t smem = (size_t)2 * (128 + N) * K * 4;
  MDL_REQUIRE(smem <= 200 * 1024, "selftest_umma: tile too large");
  MDL_CUDA(cudaFuncSetAttribute(k_umma_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_umma_selftest<<<1, 128, smem, as_stream(stream)>>>(A, B, D, N, K, split, lbo_a, sbo_a, lbo_b, sbo_b);
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_selftest_umma(const float* A, const float* B, float* D, int32_t N, int32_t K,
                                 int32_t split, void* stream) {
  return selftest_launch(A, B, D, N, K, split, 128 * 16, 128, N * 16, 128, stream);
}

// descriptor-field probe (development aid; same staging layout, caller-chosen LBO/SBO)
extern "C" int mdl_selftest_umma_ex(
    const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t split, int32_t lbo_a,
    int32_t sbo_a, int32_t lbo_b, int32_t sbo_b, void* stream) {
  return selftest_launch(A, B, D, N, K, split, lbo_a, sbo_a, lbo_b, sbo_b, stream);
}
