// nodegrad.cu -- weight/bias gradient of a node-level Linear whose K dimension is the node count:
//     dW[r, c] = sum_n dY[n, r] * x[n, c]        db[r] = sum_n dY[n, r]   (r < RB)
// i.e. the dense tail of the CGConv backward (dW_n = dPQ^T x, db = column sums of dP; reference:
// autograd through lin_f / lin_s of PyG CGConv as called at matdeeplearn/models/cgcnn.py:142).
// The library GEMM for this shape ([4C x N] . [N x C], N ~ 8k) runs on 4 thread blocks; here the
// node range is split over the whole chip (one slice per CTA, register-tiled, no atomics) and the
// per-CTA partials are summed in a fixed order -> deterministic.
#include "common.cuh"

namespace mdl {

constexpr int kNgCols = 64;   // x columns per pass (register tile)
constexpr int kNgTile = 16;   // nodes staged in smem per step

__global__ void __launch_bounds__(512)
k_node_grad(const float* __restrict__ dY, const float* __restrict__ x, float* __restrict__ part,
            int64_t N, int R, int C, int RB, int nodes_per_cta) {
  extern __shared__ __align__(16) float sx[];  // [kNgTile][C]
  const int r = threadIdx.x;                   // output row (column of dY)
  const int64_t n0 = (int64_t)blockIdx.x * nodes_per_cta;
  const int64_t n1 = (n0 + nodes_per_cta < N) ? (n0 + nodes_per_cta) : N;
  float* my = part + (size_t)blockIdx.x * ((size_t)R * C + RB);
  for (int c0 = 0; c0 < C; c0 += kNgCols) {
    const int cw = min(kNgCols, C - c0);
    float acc[kNgCols];
#pragma unroll
    for (int c = 0; c < kNgCols; ++c) acc[c] = 0.0f;
    float accb = 0.0f;
    for (int64_t t0 = n0; t0 < n1; t0 += kNgTile) {
      const int tn = (int)((n1 - t0 < kNgTile) ? (n1 - t0) : kNgTile);
      __syncthreads();
      for (int i = threadIdx.x; i < tn * C; i += blockDim.x) sx[i] = __ldg(x + t0 * C + i);
      __syncthreads();
      if (r < R) {
        float a[kNgTile];
#pragma unroll
        for (int j = 0; j < kNgTile; ++j)  // all loads of the tile in flight before the first use
          a[j] = (j < tn) ? __ldg(dY + (t0 + j) * R + r) : 0.0f;
#pragma unroll 4
        for (int j = 0; j < kNgTile; ++j) {
          if (j < tn) {
            accb += a[j];
            const float* row = sx + j * C + c0;
#pragma unroll
            for (int c = 0; c < kNgCols; c += 4) {
              if (c < cw) {
                const float4 v = *reinterpret_cast<const float4*>(row + c);
                acc[c] = fmaf(a[j], v.x, acc[c]); acc[c + 1] = fmaf(a[j], v.y, acc[c + 1]);
                acc[c + 2] = fmaf(a[j], v.z, acc[c + 2]); acc[c + 3] = fmaf(a[j], v.w, acc[c + 3]);
              }
            }
          }
        }
      }
    }
    if (r < R) {
#pragma unroll
      for (int c = 0; c < kNgCols; ++c)
        if (c < cw) my[(size_t)r * C + c0 + c] = acc[c];
      if (c0 == 0 && r < RB) my[(size_t)R * C + r] = accb;
    }
  }
}

// out[i] = sum_b part[b][i], i < len.  Block = 32 outputs x 8 partial-groups: warp w sums partials
// w, w+8, ... (coalesced 128-byte reads), the 8 group sums are combined in a fixed order.
__global__ void __launch_bounds__(256)
k_sum_partials(const float* __restrict__ part, int nparts, int64_t stride, int64_t len,
               float* __restrict__ out0, int64_t len0, float* __restrict__ out1) {
  __shared__ float sm[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  float acc = 0.0f;
  if (i < len) {
#pragma unroll 4
    for (int b = warp; b < nparts; b += 8) acc += __ldg(part + (size_t)b * stride + i);
  }
  sm[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && i < len) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][lane];
    if (i < len0) out0[i] = t;
    else if (out1) out1[i - len0] = t;
  }
}

int sum_partials(const float* part, int nparts, int64_t stride, int64_t len, float* out0, int64_t len0,
                 float* out1, cudaStream_t st) {
  k_sum_partials<<<(int)ceil_div<int64_t>(len, 32), 256, 0, st>>>(part, nparts, stride, len, out0, len0, out1);
  MDL_LAUNCHED();
  return MDL_OK;
}

}  // namespace mdl

using namespace mdl;

static int ng_ctas(int64_t N) {
  int64_t c = ceil_div<int64_t>(N, 32);
  return (int)std::max<int64_t>(1, std::min<int64_t>(c, kNumSMs));
}

extern "C" size_t mdl_node_grad_workspace_bytes(int64_t N, int32_t R, int32_t C, int32_t RB) {
  return (size_t)ng_ctas(N) * ((size_t)R * C + RB) * 4 + 256;
}

extern "C" int mdl_node_grad(const float* dY, const float* x, float* dW, float* db, int64_t N, int32_t R,
                             int32_t C, int32_t RB, void* workspace, size_t workspace_bytes, void* stream) {
  MDL_REQUIRE(N > 0 && R > 0 && R <= 512 && C > 0 && C % 4 == 0 && RB >= 0 && RB <= R,
              "node_grad: unsupported shape R=%d C=%d RB=%d", R, C, RB);
  MDL_REQUIRE(dY && x && dW && (RB == 0 || db) && workspace, "node_grad: null pointer");
  if (workspace_bytes < mdl_node_grad_workspace_bytes(N, R, C, RB)) {
    set_error("node_grad: workspace too small");
    return MDL_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  const int ctas = ng_ctas(N);
  const int per = (int)ceil_div<int64_t>(N, ctas);
  const int threads = ((R + 31) / 32) * 32;
  const size_t smem = (size_t)kNgTile * C * 4;
  k_node_grad<<<ctas, threads, smem, st>>>(dY, x, (float*)workspace, N, R, C, RB, per);
  MDL_LAUNCHED();
  const int64_t len = (int64_t)R * C + RB;
  return sum_partials((const float*)workspace, ctas, len, len, dW, (int64_t)R * C, db, st);
}
