// wgrad.cu -- weight / bias gradient of a dense layer y = x W^T + b over a tall batch, written
// straight into the caller's parameter-gradient storage:
//     dW[o][i] = sum_r G[r][o] * X[r][i],    db[o] = sum_r G[r][o]          (r over N rows)
//
// The reference gets these from autograd's Linear backward (torch.nn.Linear in
// matdeeplearn/models/cgcnn.py:64-77,97-111 and the lin_f / lin_s of PyG's CGConv), followed by
// a copy of every .grad into DDP's flat bucket.  On this path N is the node count of a batch and
// O x I <= 256 x 128: a library GEMM tiles only the small O x I output (4 CTAs for 256 x 64) and
// the bias needs a second reduction kernel.  Here the ROWS are split over the grid, each thread
// keeps an 8x8 register tile over its rows, per-CTA partials are summed in CTA order
// (deterministic, fp32 FMA) and the result lands -- through a block map -- wherever the parameter's
// gradient lives (e.g. the column blocks of CGConv's lin_f.weight / lin_s.weight inside the flat
// gradient buffer): no concatenation, no separate bias reduction.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

namespace mdl {

constexpr int kWgRows = 32;       // rows staged per chunk
constexpr int kWgThreads = 256;

__host__ __device__ inline int wg_stride(int n) { return ((n + 7) & ~7) + 4; }  // floats, 16-byte multiple

static int wg_grid(int64_t N) {
  int64_t g = ceil_div<int64_t>(N, 2 * kWgRows);   // two chunks per CTA when possible: halves the partials
  if (g > 2 * kNumSMs) g = 2 * kNumSMs;
  return (int)(g > 0 ? g : 1);
}

__device__ __forceinline__ void wg_cp_async(void* dst, const void* src, int bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  if (bytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  else if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}

// stage rows [r0, r0+rows) of a row-major [*, W] matrix into smem rows of stride `ld`: a warp per row,
// every piece an asynchronous copy (nothing waits on a load inside the loop); rows beyond `rows` are zeroed
__device__ __forceinline__ void wg_stage(float* dst, int ld, const float* __restrict__ src, int W, int64_t r0,
                                         int rows, int vec, int warp, int lane) {
  const int pieces = W / vec;
  for (int r = warp; r < kWgRows; r += kWgThreads / 32) {
    float* drow = dst + r * ld;
    if (r < rows) {
      const float* srow = src + (r0 + r) * W;
      for (int c = lane; c < pieces; c += 32) wg_cp_async(drow + c * vec, srow + c * vec, vec * 4);
    } else {
      for (int c = lane; c < W; c += 32) drow[c] = 0.f;
    }
  }
}

__global__ void __launch_bounds__(kWgThreads, 2)
k_linear_wgrad(const float* __restrict__ X, const float* __restrict__ G, int64_t N, int I, int O,
               int vx, int vg, float* __restrict__ part) {
  extern __shared__ __align__(16) float wsm[];
  const int sx = wg_stride(I), sg = wg_stride(O);
  const int buf_floats = kWgRows * (sx + sg);   // one stage: X chunk [kWgRows][sx] then G chunk [kWgRows][sg]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_i = (I + 7) >> 3, tiles_o = (O + 7) >> 3, ntiles = tiles_i * tiles_o;
  const int64_t nchunks = ceil_div<int64_t>(N, kWgRows);
  float* mine = part + (size_t)blockIdx.x * ((size_t)O * I + O);
  for (int i = threadIdx.x; i < 2 * buf_floats; i += kWgThreads) wsm[i] = 0.f;   // pad columns stay zero
  __syncthreads();
  auto stage = [&](int64_t ch, int b) {
    const int64_t r0 = ch * kWgRows;
    const int rows = (int)min((int64_t)kWgRows, N - r0);
    float* sX = wsm + b * buf_floats;
    wg_stage(sX, sx, X, I, r0, rows, vx, warp, lane);
    wg_stage(sX + kWgRows * sx, sg, G, O, r0, rows, vg, warp, lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
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
    int b = 0;
    if ((int64_t)blockIdx.x < nchunks) stage(blockIdx.x, 0);
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x, b ^= 1) {
      const bool more = ch + gridDim.x < nchunks;
      if (more) stage(ch + gridDim.x, b ^ 1);   // next chunk lands under this chunk's FMAs
      if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      const float* sX = wsm + b * buf_floats;
      const float* sG = sX + kWgRows * sx;
      if (live) {
#pragma unroll 2
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
            for (int bb = 0; bb < 8; ++bb) acc[a][bb] = fmaf(gv[a], xv[bb], acc[a][bb]);
          }
        }
      }
      __syncthreads();   // this stage is free to be overwritten by the copy issued next iteration
    }
    if (live) {
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int o = to * 8 + a;
        if (o >= O) continue;
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) {
          const int i = ti * 8 + bb;
          if (i < I) mine[(size_t)o * I + i] = acc[a][bb];
        }
        if (ti == 0) mine[(size_t)O * I + o] = accb[a];
      }
    }
  }
}

// element (o, i) of the [O x I] result -> its home in the caller's gradient storage
__device__ __forceinline__ float* wg_dest(const mdl_wgrad_out& m, int o, int i) {
  const int b = o / m.block_rows;
  float* base = m.w[b];
  return base ? base + (size_t)(o - b * m.block_rows) * m.ldw + i : nullptr;
}

// out(o, i) = sum_b part[b][o*I + i]; bias sums sit after the O*I weights of each partial.
// Block = 32 outputs x 8 partial groups (coalesced 128-byte reads), groups combined in a fixed order.
__global__ void __launch_bounds__(256)
k_wgrad_reduce(const float* __restrict__ part, int nparts, int I, int O, const mdl_wgrad_out m) {
  __shared__ float sm[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t len = (int64_t)O * I + O, stride = len;
  const int64_t idx = (int64_t)blockIdx.x * 32 + lane;
  float acc = 0.0f;
  if (idx < len) {
#pragma unroll 4
    for (int b = warp; b < nparts; b += 8) acc += __ldg(part + (size_t)b * stride + idx);
  }
  sm[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && idx < len) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][lane];
    if (idx < (int64_t)O * I) {
      const int o = (int)(idx / I), i = (int)(idx - (int64_t)o * I);
      float* d = wg_dest(m, o, i);
      if (d) *d = t;
    } else {
      const int o = (int)(idx - (int64_t)O * I);
      const int b = o / m.block_rows;
      if (m.b[b]) m.b[b][o - b * m.block_rows] = t;
    }
  }
}

// out(o, i) = src[i*O + o] (transposed == 1) or src[o*I + i]: scatter of a small dense result
__global__ void k_copy_mapped(const float* __restrict__ src, int I, int O, int transposed, const mdl_wgrad_out m) {
  const int total = I * O;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    int o, i;
    if (transposed) { i = t / O; o = t - i * O; }
    else { o = t / I; i = t - o * I; }
    float* d = wg_dest(m, o, i);
    if (d) *d = __ldg(src + t);
  }
}

// shared with wgrad_tc.cu: sum `nparts` partials ([O*I | O] each) in order and deliver through the block map
int wgrad_reduce_launch(const float* part, int nparts, int I, int O, const mdl_wgrad_out& m, cudaStream_t st) {
  const int64_t len = (int64_t)O * I + O;
  k_wgrad_reduce<<<(int)ceil_div<int64_t>(len, 32), 256, 0, st>>>(part, nparts, I, O, m);
  MDL_LAUNCHED();
  return MDL_OK;
}
bool wgrad_tc_supported(int64_t R, int I, int O);
int wgrad_tc_launch(const float* X, const float* G, const float* rs, int64_t R, int I, int O, const mdl_wgrad_out& out,
                    float* part, cudaStream_t st);

static size_t wg_ws_bytes(int64_t N, int I, int O) {
  return (size_t)wg_grid(N) * ((size_t)O * I + O) * sizeof(float);
}

static int wg_check_map(const mdl_wgrad_out* out, int O, const char* who) {
  MDL_REQUIRE(out, "%s: null output map", who);
  MDL_REQUIRE(out->block_rows > 0 && out->num_blocks > 0 && out->num_blocks <= 8 &&
                  (int64_t)out->block_rows * out->num_blocks >= O, "%s: block map does not cover %d rows", who, O);
  return MDL_OK;
}

}  // namespace mdl

using namespace mdl;

extern "C" size_t mdl_linear_wgrad_workspace_bytes(int64_t N, int32_t I, int32_t O) {
  if (N < 0 || I <= 0 || O <= 0) return 0;
  return wg_ws_bytes(N, I, O);
}

extern "C" int mdl_linear_wgrad(const float* X, const float* G, int64_t N, int32_t I, int32_t O,
                                const mdl_wgrad_out* out, void* workspace, size_t workspace_bytes,
                                void* stream) {
  MDL_REQUIRE(N > 0 && I > 0 && O > 0, "linear_wgrad: bad shape");
  if (int rc = wg_check_map(out, O, "linear_wgrad")) return rc;
  MDL_REQUIRE(X && G && workspace, "linear_wgrad: null pointer");
  MDL_REQUIRE(workspace_bytes >= wg_ws_bytes(N, I, O), "linear_wgrad: workspace too small");
  const size_t smem = (size_t)2 * kWgRows * (wg_stride(I) + wg_stride(O)) * sizeof(float);
  {  // long batches (edge-level layers, >= 16k rows) and layers too wide for the SIMT kernel's row stages: the tcgen05
     // kernel (wgrad_tc.cu).  Node-level layers (a few thousand rows) stay on the SIMT kernel, which is faster there
     // (profiles/r2_launches_c1_summary.txt: 24.8 vs 17.4 us).  MDL_WGRAD=simt / =tc force one or the other (A/B).
    const char* env = getenv("MDL_WGRAD");
    const bool force_tc = env && strcmp(env, "tc") == 0, force_simt = env && strcmp(env, "simt") == 0;
    if (!force_simt && wgrad_tc_supported(N, I, O) && (force_tc || N >= 16384 || smem > 200 * 1024))
      return wgrad_tc_launch(X, G, nullptr, N, I, O, *out, reinterpret_cast<float*>(workspace), as_stream(stream));
  }
  // two stages of 32 rows of X and G: I + O <= ~780 floats per row pair (one CTA per SM above ~100 KB)
  MDL_REQUIRE(smem <= 200 * 1024, "linear_wgrad: layer too wide (needs %zu bytes of shared memory, I + O <= ~780)", smem);
  static std::atomic<int> attr_set{0};
  if (!attr_set.load(std::memory_order_acquire)) {
    MDL_CUDA(cudaFuncSetAttribute(k_linear_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set.store(1, std::memory_order_release);
  }
  cudaStream_t st = as_stream(stream);
  const int grid = wg_grid(N);
  float* part = reinterpret_cast<float*>(workspace);
  // widest asynchronous copy the row and base alignment allow (rows start at multiples of W floats)
  auto vec = [](const float* p, int W) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    return (W % 4 == 0 && a % 16 == 0) ? 4 : (W % 2 == 0 && a % 8 == 0) ? 2 : 1;
  };
  k_linear_wgrad<<<grid, kWgThreads, smem, st>>>(X, G, N, I, O, vec(X, I), vec(G, O), part);
  MDL_LAUNCHED();
  const int64_t len = (int64_t)O * I + O;
  k_wgrad_reduce<<<(int)ceil_div<int64_t>(len, 32), 256, 0, st>>>(part, grid, I, O, *out);
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_copy_mapped(const float* src, int32_t I, int32_t O, int32_t transposed,
                               const mdl_wgrad_out* out, void* stream) {
  MDL_REQUIRE(src && I > 0 && O > 0, "copy_mapped: bad arguments");
  if (int rc = wg_check_map(out, O, "copy_mapped")) return rc;
  k_copy_mapped<<<(int)std::min<int64_t>(ceil_div<int64_t>((int64_t)I * O, 256), 4 * kNumSMs), 256, 0,
                  as_stream(stream)>>>(src, I, O, transposed, *out);
  MDL_LAUNCHED();
  return MDL_OK;
}

// dW = (diag(rowscale) G)^T X, db = sum_r rowscale[r] G[r]: the weight gradient of a layer whose output is scaled per
// row afterwards (SchNet: filter * cosine cutoff, reference schnet.py:81 / PyG CFConv.forward).  Long batches only
// (the tcgen05 kernel, wgrad_tc.cu).
extern "C" int mdl_linear_wgrad_rs(const float* X, const float* G, const float* rowscale, int64_t N, int32_t I, int32_t O,
                                   const mdl_wgrad_out* out, void* workspace, size_t workspace_bytes, void* stream) {
  MDL_REQUIRE(N > 0 && I > 0 && O > 0, "linear_wgrad_rs: bad shape");
  if (int rc = wg_check_map(out, O, "linear_wgrad_rs")) return rc;
  MDL_REQUIRE(X && G && workspace, "linear_wgrad_rs: null pointer");
  MDL_REQUIRE(workspace_bytes >= wg_ws_bytes(N, I, O), "linear_wgrad_rs: workspace too small");
  MDL_REQUIRE(wgrad_tc_supported(N, I, O), "linear_wgrad_rs: needs N >= 2048 rows and I <= 256 (got N=%lld I=%d)", (long long)N, I);
  return wgrad_tc_launch(X, G, rowscale, N, I, O, *out, reinterpret_cast<float*>(workspace), as_stream(stream));
}
