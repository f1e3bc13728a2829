// wgrad.cu -- weight/bias gradient of a dense layer y = x W^T + b over a tall batch:
//     dW[o][i] = sum_r G[r][o] * X[r][i],    db[o] = sum_r G[r][o]          (r over N rows)
//
// The reference gets these from autograd's Linear backward (torch.nn.Linear in
// matdeeplearn/models/cgcnn.py:64-77,97-111 and inside PyG's CGConv).  Shapes on this path are tall
// and skinny (N = all nodes of a batch, O x I <= 256 x 128): a library GEMM tiles only the tiny
// O x I output and leaves most SMs idle.  Here the ROWS are split over the grid, each CTA keeps an
// 8x8 register tile per thread over its rows, and the per-CTA partials are summed in CTA order
// (sum_partials): deterministic, fp32 FMA.
#include "common.cuh"
#include "cgconv.cuh"   // sum_partials

namespace mdl {

constexpr int kWgRows = 32;       // rows staged per chunk
constexpr int kWgThreads = 256;

__host__ __device__ inline int wg_pad4(int n) { return (n + 3) & ~3; }
// row stride of the staged chunks: multiple of 4 floats and covering whole 8-wide tiles
__host__ __device__ inline int wg_stride(int n) { return ((n + 7) & ~7) + 4; }

static int wg_grid(int64_t N) {
  int64_t g = ceil_div<int64_t>(N, kWgRows);
  if (g > kNumSMs) g = kNumSMs;
  return (int)(g > 0 ? g : 1);
}

__global__ void __launch_bounds__(kWgThreads)
k_linear_wgrad(const float* __restrict__ X, const float* __restrict__ G, int64_t N, int I, int O,
               float* __restrict__ part) {
  extern __shared__ __align__(16) float wsm[];
  const int sx = wg_stride(I), sg = wg_stride(O);
  float* sX = wsm;                  // [kWgRows][sx]
  float* sG = wsm + kWgRows * sx;   // [kWgRows][sg]
  const int tiles_i = (I + 7) >> 3, tiles_o = (O + 7) >> 3, ntiles = tiles_i * tiles_o;
  const int64_t nchunks = ceil_div<int64_t>(N, kWgRows);
  float* mine = part + (size_t)blockIdx.x * ((size_t)O * I + O);
  for (int i = threadIdx.x; i < kWgRows * sx; i += kWgThreads) sX[i] = 0.f;   // pads stay zero
  for (int i = threadIdx.x; i < kWgRows * sg; i += kWgThreads) sG[i] = 0.f;
  for (int t0 = 0; t0 < ntiles; t0 += kWgThreads) {
    const int t = t0 + threadIdx.x;
    const bool live = t < ntiles;
    const int to = live ? t / tiles_i : 0, ti = live ? t - to * tiles_i : 0;
    float acc[8][8];
    float accb[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      accb[a] = 0.f;
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
    }
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
      const int64_t r0 = ch * kWgRows;
      const int rows = (int)min((int64_t)kWgRows, N - r0);
      __syncthreads();
      for (int i = threadIdx.x; i < kWgRows * I; i += kWgThreads) {
        const int r = i / I, c = i - r * I;
        sX[r * sx + c] = (r < rows) ? __ldg(X + (r0 + r) * I + c) : 0.f;
      }
      for (int i = threadIdx.x; i < kWgRows * O; i += kWgThreads) {
        const int r = i / O, c = i - r * O;
        sG[r * sg + c] = (r < rows) ? __ldg(G + (r0 + r) * O + c) : 0.f;
      }
      __syncthreads();
      if (live) {
#pragma unroll 4
        for (int r = 0; r < kWgRows; ++r) {
          const float4 g0 = *reinterpret_cast<const float4*>(sG + r * sg + to * 8);
          const float4 g1 = *reinterpret_cast<const float4*>(sG + r * sg + to * 8 + 4);
          const float4 x0 = *reinterpret_cast<const float4*>(sX + r * sx + ti * 8);
          const float4 x1 = *reinterpret_cast<const float4*>(sX + r * sx + ti * 8 + 4);
          const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int a = 0; a < 8; ++a) {
            accb[a] += gv[a];
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(gv[a], xv[b], acc[a][b]);
          }
        }
      }
    }
    if (live) {
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int o = to * 8 + a;
        if (o >= O) continue;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int i = ti * 8 + b;
          if (i < I) mine[(size_t)o * I + i] = acc[a][b];
        }
        if (ti == 0) mine[(size_t)O * I + o] = accb[a];
      }
    }
  }
}

static size_t wg_ws_bytes(int64_t N, int I, int O) {
  return (size_t)wg_grid(N) * ((size_t)O * I + O) * sizeof(float);
}

}  // namespace mdl

using namespace mdl;

extern "C" size_t mdl_linear_wgrad_workspace_bytes(int64_t N, int32_t I, int32_t O) {
  if (N < 0 || I <= 0 || O <= 0) return 0;
  return wg_ws_bytes(N, I, O);
}

extern "C" int mdl_linear_wgrad(const float* X, const float* G, int64_t N, int32_t I, int32_t O,
                                float* dW, float* db, void* workspace, size_t workspace_bytes,
                                void* stream) {
  MDL_REQUIRE(N >= 0 && I > 0 && O > 0, "linear_wgrad: bad shape");
  MDL_REQUIRE(dW, "linear_wgrad: null pointer");
  cudaStream_t st = as_stream(stream);
  if (N == 0) {
    MDL_CUDA(cudaMemsetAsync(dW, 0, (size_t)O * I * 4, st));
    if (db) MDL_CUDA(cudaMemsetAsync(db, 0, (size_t)O * 4, st));
    return MDL_OK;
  }
  MDL_REQUIRE(X && G && workspace, "linear_wgrad: null pointer");
  MDL_REQUIRE(workspace_bytes >= wg_ws_bytes(N, I, O), "linear_wgrad: workspace too small");
  const size_t smem = (size_t)kWgRows * (wg_stride(I) + wg_stride(O)) * sizeof(float);
  MDL_REQUIRE(smem <= 200 * 1024, "linear_wgrad: layer too wide (I + O <= ~1500)");
  static bool attr_set = false;
  if (!attr_set) {
    MDL_CUDA(cudaFuncSetAttribute(k_linear_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int grid = wg_grid(N);
  float* part = reinterpret_cast<float*>(workspace);
  k_linear_wgrad<<<grid, kWgThreads, smem, st>>>(X, G, N, I, O, part);
  MDL_LAUNCHED();
  const int64_t len = (int64_t)O * I + O;
  // db may be NULL: the bias sums then land in a scratch tail of the first partial (already consumed)
  return sum_partials(part, grid, len, db ? len : (int64_t)O * I, dW, (int64_t)O * I, db, st);
}
