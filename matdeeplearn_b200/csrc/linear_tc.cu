// linear_tc.cu -- dense layer over a LONG batch on tcgen05:   Y[r, :] = act( X[r, :] . B^T + bias )
//     X [R, K] row-major (K <= 128), B given through strides as B[n][k] = W[n * ldn + k * ldk] (N <= 128):
//       forward  y = x W^T, W [O, I]:   n = o, k = i  ->  ldn = I, ldk = 1
//       backward dx = g W,  W [O, I]:   n = i, k = o  ->  ldn = 1, ldk = I
//
// Reference: the edge-level Linear layers of the reference's models -- MEGNet's edge / embedding MLPs
// (matdeeplearn/models/megnet.py:28-56, 222-247), the NNConv edge network's first layer (mpnn.py:83-85) -- and their
// input gradients; on the reference's path cuBLAS SIMT SGEMMs over E rows.  A round = 128 rows: the rows are read
// once (thread = row = TMEM lane), split hi / lo into tensor memory (A operand), contracted with the resident weight
// tile (3xTF32, fp32-faithful, fp32 accumulation in TMEM), and written once with bias and activation applied.
#include "common.cuh"
#include "umma.cuh"
#include "edge_dev.cuh"

namespace mdl {
namespace {

constexpr int kLW = 512, kLWarps = 16, kLLaunch = kLW + 32, kLRows = 128;

unsigned long long* g_lin_phase_buf = nullptr;

struct LinParams {
  unsigned long long* prof;
  const float* X; const float* W; const float* bias; float* Y;
  int64_t R, ldn, ldk;
  int K, N, KP, NP, act;
};

__global__ void __launch_bounds__(kLLaunch, 1) k_linear_tc(const LinParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ int sMail[2];
  __shared__ float sBias[128];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = p.K, N = p.N, KP = p.KP, NP = p.NP;
  uint8_t* sWhi = smem;                                  // B tile [NP rows][KP] canonical K-major, hi / lo
  uint8_t* sWlo = smem + (size_t)NP * KP * 4;
  auto sync_issuer = [] { asm volatile("bar.sync 1, %0;" ::"n"(kLLaunch) : "memory"); };
  const int64_t n_tiles = (p.R + kLRows - 1) / kLRows;

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) {
    umma::mbar_init(&bar_mma, 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < NP * KP; i += kLLaunch) {
    // consecutive threads along the weight's contiguous index
    const int n = (p.ldk == 1) ? i / KP : i % NP, k = (p.ldk == 1) ? i % KP : i / NP;
    const float w = (n < N && k < K) ? __ldg(p.W + n * p.ldn + k * p.ldk) : 0.0f;
    const float hi = umma::tf32_hi(w);
    const int off = umma::tile_offset_bytes(n, k, NP);
    *reinterpret_cast<float*>(sWhi + off) = hi;
    *reinterpret_cast<float*>(sWlo + off) = w - hi;
  }
  if (tid < 128) sBias[tid] = (p.bias && tid < N) ? __ldg(p.bias + tid) : 0.0f;
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma::make_idesc_tf32(kLRows, NP);
  const uint32_t colA = 128;                             // D [0, 128) | A hi [128, 256) | A lo [256, 384)

  if (warp == kLWarps) {
    for (uint32_t mb = 0;; mb ^= 1) {
      sync_issuer();
      if (sMail[mb] == 3) break;
      if (lane == 0) {
        umma::fence_after_sync();
        const uint32_t step = 2 * (uint32_t)NP * 16;
        uint32_t acc = 0;
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a = tmem + colA + ((pass == 2) ? 128u : 0u);
          const uint32_t b = umma::smem_u32((pass == 1) ? sWlo : sWhi);
          for (int kk = 0; kk < (KP >> 3); ++kk) {
            umma::mma_tf32_ts(tmem, a + kk * 8, umma::make_desc(b + kk * step, (uint32_t)NP * 16, 128), idesc, acc);
            acc = 1;
          }
        }
        umma::mma_commit(&bar_mma);
      }
      __syncwarp();
    }
    __syncthreads();
    return;
  }

  long long t_prev = clock64();
  auto mark = [&](int slot) {   // development aid: per-phase cycles of thread 0 (mdl_debug_set_phase_buffer)
    if (p.prof && tid == 0) {
      const long long now = clock64();
      atomicAdd(p.prof + slot, (unsigned long long)(now - t_prev));
      t_prev = now;
    }
  };
  const int q = warp & 3, part = warp >> 2;
  const int e = 32 * q + lane;
  const bool vec4 = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(p.X) & 15) == 0;
  const bool vec4o = (N & 3) == 0 && (reinterpret_cast<uintptr_t>(p.Y) & 15) == 0;
  const bool vec8o = (N & 7) == 0 && (reinterpret_cast<uintptr_t>(p.Y) & 31) == 0;
  uint32_t mb = 0, ph = 0;
  const bool vec8 = (K & 7) == 0 && (reinterpret_cast<uintptr_t>(p.X) & 31) == 0;
  // this thread's 32-column piece of row e of tile t (zero beyond the matrix)
  auto load_piece = [&](int64_t t, float (&x)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = 0.0f;
    const int64_t r = t * kLRows + e;
    if (t < n_tiles && r < p.R && 32 * part < KP) {
      const float* src = p.X + r * K + 32 * part;
      if (vec8 && 32 * part + 32 <= K) {
#pragma unroll
        for (int j = 0; j < 4; ++j) umma::ldg256(src + 8 * j, x + 8 * j);
      } else if (vec4 && 32 * part + 32 <= K) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src) + j);
          x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (32 * part + j < K) x[j] = __ldg(src + j);
      }
    }
  };
  float x[32];
  load_piece(blockIdx.x, x);
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t r_lo = t * kLRows;
    const int cnt = (int)min((int64_t)kLRows, p.R - r_lo);
    const bool live = e < cnt;
    mark(0);
    // ---- X row piece (32 columns per thread, requested one tile ahead) -> hi / lo -> A operand
    if (32 * part < KP) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (32 * part + 16 * half < KP) {
          float hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { hi[j] = umma::tf32_hi(x[16 * half + j]); lo[j] = x[16 * half + j] - hi[j]; }
          umma::tmem_st16(umma::tmem_addr(tmem + colA, warp, 32 * part + 16 * half), hi);
          umma::tmem_st16(umma::tmem_addr(tmem + colA + 128, warp, 32 * part + 16 * half), lo);
        }
      }
    }
    mark(1);
    umma::tmem_st_wait();
    mark(2);
    if (tid == 0) sMail[mb] = 1;
    umma::fence_before_sync();
    sync_issuer();   // also: every thread has read D of the previous tile
    mb ^= 1;
    mark(3);
    load_piece(t + gridDim.x, x);   // the next tile's rows: in flight under this tile's MMAs and epilogue
    umma::mbar_wait(&bar_mma, ph);
    ph ^= 1;
    umma::fence_after_sync();
    mark(4);
    if (32 * part < NP) {
      float d[32];
      umma::tmem_ld16(umma::tmem_addr(tmem, q, 32 * part), *reinterpret_cast<float(*)[16]>(d));
      if (32 * part + 16 < NP) umma::tmem_ld16(umma::tmem_addr(tmem, q, 32 * part + 16), *reinterpret_cast<float(*)[16]>(d + 16));
      umma::tmem_ld_wait();
      mark(5);
      if (live) {
        float* dst = p.Y + (r_lo + e) * N + 32 * part;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v = d[j] + sBias[(32 * part + j) & 127];
          if (p.act == 1) v = fmaxf(v, 0.0f);
          else if (p.act == 2) v = fmaf(kLn2, lg2_(1.0f + ex2_(-kLog2e * fabsf(v))), fmaxf(v, 0.0f)) - kLn2;
          d[j] = v;
        }
        if (vec8o && 32 * part + 32 <= N) {
#pragma unroll
          for (int j = 0; j < 4; ++j) umma::stg256(dst + 8 * j, d + 8 * j);
        } else if (vec4o && 32 * part + 32 <= N) {
#pragma unroll
          for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(dst)[j] = make_float4(d[4 * j], d[4 * j + 1], d[4 * j + 2], d[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (32 * part + j < N) dst[j] = d[j];
        }
      }
    }
    mark(6);
    if (p.prof && tid == 0) atomicAdd(p.prof + 31, 1ull);
  }
  if (tid == 0) sMail[mb] = 3;
  umma::fence_before_sync();
  sync_issuer();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace
void lin_set_phase_buffer(unsigned long long* dev_ptr) { g_lin_phase_buf = dev_ptr; }
}  // namespace mdl

using namespace mdl;

extern "C" int mdl_linear_tc_supported(int64_t R, int32_t K, int32_t N) {
  return (R >= 16384 && K >= 8 && K <= 128 && N >= 8 && N <= 128) ? 1 : 0;
}

extern "C" int mdl_linear_tc(const float* X, const float* W, int64_t ldn, int64_t ldk, const float* bias, float* Y,
                             int64_t R, int32_t K, int32_t N, int32_t act, void* stream) {
  MDL_REQUIRE(X && W && Y && R >= 0, "linear_tc: null pointer");
  MDL_REQUIRE(K >= 1 && K <= 128 && N >= 1 && N <= 128, "linear_tc: K and N must be in [1, 128] (got %d, %d)", K, N);
  MDL_REQUIRE(act >= 0 && act <= 2, "linear_tc: act 0 (none), 1 (relu), 2 (shifted softplus)");
  if (R == 0) return MDL_OK;
  LinParams p{};
  p.prof = g_lin_phase_buf;
  p.X = X; p.W = W; p.bias = bias; p.Y = Y; p.R = R; p.ldn = ldn; p.ldk = ldk;
  p.K = K; p.N = N; p.KP = (K + 7) & ~7; p.NP = (N + 15) & ~15; p.act = act;
  const int smem = 2 * p.NP * p.KP * 4;
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MDL_CUDA(cudaFuncSetAttribute(k_linear_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured.store(1, std::memory_order_release);
  }
  const int64_t n_tiles = (R + kLRows - 1) / kLRows;
  // one CTA per SM (each allocates the whole tensor memory): ask for at least half an SM's shared memory
  const int smem_req = smem < 120 * 1024 ? 120 * 1024 : smem;
  k_linear_tc<<<(int)std::min<int64_t>(n_tiles, kNumSMs), kLLaunch, smem_req, as_stream(stream)>>>(p);
  MDL_LAUNCHED();
  return MDL_OK;
}
