// batchnorm.cu -- training-mode BatchNorm1d over the first n_valid rows of x[N,C].
//
// The reference normalises node features after every conv with torch.nn.BatchNorm1d
// (matdeeplearn/models/cgcnn.py:88-92,141-147).  The engine's capacity-padded batches
// (assemble.cu) carry inert rows beyond the batch's real node count, and that count lives in
// device memory so that one CUDA graph serves every batch: the statistics must therefore be
// masked by a device-side row count, which torch's batch_norm cannot do.
//
// Statistics: per-CTA Welford partials over a row chunk, combined in CTA order (Chan et al.) by
// the last CTA to finish -- one pass over x, deterministic, no E[x^2]-E[x]^2 cancellation.
#include "common.cuh"

namespace mdl {

constexpr int kBnThreads = 256;
constexpr int kBnRowsPerCta = 64;

struct Welford {
  float n, mean, m2;
};
__device__ __forceinline__ void wf_add(Welford& w, float x) {
  w.n += 1.0f;
  float d = x - w.mean;
  w.mean += d * __frcp_rn(w.n);
  w.m2 += d * (x - w.mean);
}
__device__ __forceinline__ void wf_merge(Welford& a, const Welford& b) {
  if (b.n == 0.0f) return;
  float n = a.n + b.n, d = b.mean - a.mean, f = b.n / n;
  a.mean += d * f;
  a.m2 += b.m2 + d * d * a.n * f;
  a.n = n;
}

static int bn_grid(int64_t N) {
  int64_t g = ceil_div<int64_t>(N, kBnRowsPerCta);
  if (g > 2 * kNumSMs) g = 2 * kNumSMs;
  return (int)(g > 0 ? g : 1);
}

// column tiling of a CTA: cw columns side by side (power of two <= 256), kBnThreads/cw row lanes
__host__ __device__ inline int bn_cw(int C) {
  int cw = 1;
  while (cw * 2 <= C && cw * 2 <= kBnThreads) cw *= 2;
  return cw;
}

// workspace: [0] ticket (u32, zero between launches), then partials [grid][3][C]
__global__ void __launch_bounds__(kBnThreads)
k_bn_stats(const float* __restrict__ x, const int32_t* __restrict__ n_valid, int64_t N, int C,
           float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps,
           float* __restrict__ save_mean, float* __restrict__ save_invstd, unsigned* ticket,
           float* __restrict__ part) {
  __shared__ Welford sh[kBnThreads];
  __shared__ bool last;
  const int64_t n = n_valid ? min((int64_t)__ldg(n_valid), N) : N;
  const int cw = bn_cw(C), th = kBnThreads / cw;
  const int tx = threadIdx.x % cw, ty = threadIdx.x / cw;
  const int64_t per = ceil_div<int64_t>(n > 0 ? n : 1, gridDim.x);
  const int64_t r0 = blockIdx.x * per, r1 = min(n, r0 + per);
  float* mine = part + (size_t)blockIdx.x * 3 * C;
  for (int c0 = 0; c0 < C; c0 += cw) {
    const int c = c0 + tx;
    Welford w{0.f, 0.f, 0.f};
    if (c < C) {
      for (int64_t rb = r0 + ty; rb < r1; rb += 8 * th) {   // 8 rows in flight per thread
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t r = rb + (int64_t)u * th;
          v[u] = (r < r1) ? __ldg(x + r * C + c) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (rb + (int64_t)u * th < r1) wf_add(w, v[u]);
      }
    }
    sh[threadIdx.x] = w;
    __syncthreads();
    if (ty == 0 && c < C) {
      for (int k = 1; k < th; ++k) wf_merge(w, sh[k * cw + tx]);
      mine[c] = w.n;
      mine[C + c] = w.mean;
      mine[2 * C + c] = w.m2;
    }
    __syncthreads();
  }
  __threadfence();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last CTA merges the per-CTA partials: row lane ty takes partials ty, ty+th, ... (loads batched
  // so their L2 latencies overlap), then lane 0 merges the th results -- a fixed order either way
  for (int c0 = 0; c0 < C; c0 += cw) {
    const int c = c0 + tx;
    Welford w{0.f, 0.f, 0.f};
    if (c < C) {
      for (unsigned b0 = ty; b0 < gridDim.x; b0 += 8 * th) {
        Welford q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const unsigned b = b0 + u * th;
          const float* p = part + (size_t)b * 3 * C;
          q[u] = (b < gridDim.x) ? Welford{__ldcg(p + c), __ldcg(p + C + c), __ldcg(p + 2 * C + c)}
                                 : Welford{0.f, 0.f, 0.f};
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) wf_merge(w, q[u]);
      }
    }
    sh[threadIdx.x] = w;
    __syncthreads();
    if (ty == 0 && c < C) {
      for (int k = 1; k < th; ++k) wf_merge(w, sh[k * cw + tx]);
      const float var = w.n > 0.f ? w.m2 / w.n : 0.f;
      save_mean[c] = w.mean;
      save_invstd[c] = rsqrtf(var + eps);
      if (running_mean) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * w.mean;
        const float unbiased = w.n > 1.f ? w.m2 / (w.n - 1.f) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

__global__ void __launch_bounds__(256)
k_bn_apply(const float* __restrict__ x, const int32_t* __restrict__ n_valid, int64_t N, int C,
           const float* __restrict__ weight, const float* __restrict__ bias,
           const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ out) {
  const int64_t n = n_valid ? min((int64_t)__ldg(n_valid), N) : N;
  const int64_t total = N * C, valid = n * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float y = 0.f;
    if (i < valid) {
      y = (__ldg(x + i) - __ldg(mean + c)) * __ldg(invstd + c);
      y = y * (weight ? __ldg(weight + c) : 1.f) + (bias ? __ldg(bias + c) : 0.f);
    }
    out[i] = y;
  }
}

// partials [grid][2][C]: sum g, sum g*xhat; the last CTA reduces them in CTA order
__global__ void __launch_bounds__(kBnThreads)
k_bn_bwd_stats(const float* __restrict__ g, const float* __restrict__ x, const int32_t* __restrict__ n_valid,
               int64_t N, int C, const float* __restrict__ mean, const float* __restrict__ invstd,
               float* __restrict__ gweight, float* __restrict__ gbias, float* __restrict__ sums,
               unsigned* ticket, float* __restrict__ part) {
  __shared__ float sh[2][kBnThreads];
  __shared__ bool last;
  const int64_t n = n_valid ? min((int64_t)__ldg(n_valid), N) : N;
  const int cw = bn_cw(C), th = kBnThreads / cw;
  const int tx = threadIdx.x % cw, ty = threadIdx.x / cw;
  const int64_t per = ceil_div<int64_t>(n > 0 ? n : 1, gridDim.x);
  const int64_t r0 = blockIdx.x * per, r1 = min(n, r0 + per);
  float* mine = part + (size_t)blockIdx.x * 2 * C;
  for (int c0 = 0; c0 < C; c0 += cw) {
    const int c = c0 + tx;
    float sg = 0.f, sgx = 0.f;
    if (c < C) {
      const float m = __ldg(mean + c), is = __ldg(invstd + c);
      for (int64_t rb = r0 + ty; rb < r1; rb += 8 * th) {   // 8 rows in flight per thread
        float gv[8], xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t r = rb + (int64_t)u * th;
          const bool ok = r < r1;
          gv[u] = ok ? __ldg(g + r * C + c) : 0.f;
          xv[u] = ok ? __ldg(x + r * C + c) : m;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          sg += gv[u];
          sgx += gv[u] * ((xv[u] - m) * is);
        }
      }
    }
    sh[0][threadIdx.x] = sg;
    sh[1][threadIdx.x] = sgx;
    __syncthreads();
    if (ty == 0 && c < C) {
      for (int k = 1; k < th; ++k) {
        sg += sh[0][k * cw + tx];
        sgx += sh[1][k * cw + tx];
      }
      mine[c] = sg;
      mine[C + c] = sgx;
    }
    __syncthreads();
  }
  __threadfence();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int c0 = 0; c0 < C; c0 += cw) {
    const int c = c0 + tx;
    float sg = 0.f, sgx = 0.f;
    if (c < C) {
      for (unsigned b0 = ty; b0 < gridDim.x; b0 += 8 * th) {
        float a[8], bx[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const unsigned b = b0 + u * th;
          const float* p = part + (size_t)b * 2 * C;
          a[u] = (b < gridDim.x) ? __ldcg(p + c) : 0.f;
          bx[u] = (b < gridDim.x) ? __ldcg(p + C + c) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          sg += a[u];
          sgx += bx[u];
        }
      }
    }
    sh[0][threadIdx.x] = sg;
    sh[1][threadIdx.x] = sgx;
    __syncthreads();
    if (ty == 0 && c < C) {
      for (int k = 1; k < th; ++k) {
        sg += sh[0][k * cw + tx];
        sgx += sh[1][k * cw + tx];
      }
      sums[c] = sg;
      sums[C + c] = sgx;
      if (gbias) gbias[c] = sg;
      if (gweight) gweight[c] = sgx;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

__global__ void __launch_bounds__(256)
k_bn_bwd_apply(const float* __restrict__ g, const float* __restrict__ x, const int32_t* __restrict__ n_valid,
               int64_t N, int C, const float* __restrict__ weight, const float* __restrict__ mean,
               const float* __restrict__ invstd, const float* __restrict__ sums, float* __restrict__ gx) {
  const int64_t n = n_valid ? min((int64_t)__ldg(n_valid), N) : N;
  const int64_t total = N * C, valid = n * C;
  const float inv_n = n > 0 ? 1.0f / (float)n : 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float v = 0.f;
    if (i < valid) {
      const float is = __ldg(invstd + c);
      const float xh = (__ldg(x + i) - __ldg(mean + c)) * is;
      v = (weight ? __ldg(weight + c) : 1.f) * is *
          (__ldg(g + i) - __ldg(sums + c) * inv_n - xh * __ldg(sums + C + c) * inv_n);
    }
    gx[i] = v;
  }
}

static size_t bn_ws_bytes(int64_t N, int C) {
  return 256 + (size_t)bn_grid(N) * 3 * C * sizeof(float) + 2 * (size_t)C * sizeof(float);
}

}  // namespace mdl

using namespace mdl;

extern "C" size_t mdl_batchnorm_workspace_bytes(int64_t N, int32_t C) {
  if (N < 0 || C <= 0) return 0;
  return bn_ws_bytes(N, C);
}

extern "C" int mdl_batchnorm_fwd(const float* x, const int32_t* n_valid, int64_t N, int32_t C,
                                 const float* weight, const float* bias, float* running_mean,
                                 float* running_var, float momentum, float eps, float* out,
                                 float* save_mean, float* save_invstd, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  MDL_REQUIRE(N >= 0 && C > 0, "batchnorm_fwd: bad shape");
  if (N == 0) return MDL_OK;
  MDL_REQUIRE(x && out && save_mean && save_invstd && workspace, "batchnorm_fwd: null pointer");
  MDL_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "batchnorm_fwd: running stats come in pairs");
  MDL_REQUIRE(workspace_bytes >= bn_ws_bytes(N, C), "batchnorm_fwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  unsigned* ticket = reinterpret_cast<unsigned*>(workspace);
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  k_bn_stats<<<bn_grid(N), kBnThreads, 0, st>>>(x, n_valid, N, C, running_mean, running_var, momentum, eps,
                                                 save_mean, save_invstd, ticket, part);
  MDL_LAUNCHED();
  int grid = (int)std::min<int64_t>(ceil_div<int64_t>(N * C, 256), (int64_t)kNumSMs * 8);
  k_bn_apply<<<grid, 256, 0, st>>>(x, n_valid, N, C, weight, bias, save_mean, save_invstd, out);
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_batchnorm_bwd(const float* gout, const float* x, const int32_t* n_valid, int64_t N,
                                 int32_t C, const float* weight, const float* save_mean,
                                 const float* save_invstd, float* gx, float* gweight, float* gbias,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  MDL_REQUIRE(N >= 0 && C > 0, "batchnorm_bwd: bad shape");
  if (N == 0) return MDL_OK;
  MDL_REQUIRE(gout && x && save_mean && save_invstd && gx && workspace, "batchnorm_bwd: null pointer");
  MDL_REQUIRE(workspace_bytes >= bn_ws_bytes(N, C), "batchnorm_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  unsigned* ticket = reinterpret_cast<unsigned*>(workspace);
  float* part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  float* sums = part + (size_t)bn_grid(N) * 3 * C;
  k_bn_bwd_stats<<<bn_grid(N), kBnThreads, 0, st>>>(gout, x, n_valid, N, C, save_mean, save_invstd, gweight,
                                                     gbias, sums, ticket, part);
  MDL_LAUNCHED();
  int grid = (int)std::min<int64_t>(ceil_div<int64_t>(N * C, 256), (int64_t)kNumSMs * 8);
  k_bn_bwd_apply<<<grid, 256, 0, st>>>(gout, x, n_valid, N, C, weight, save_mean, save_invstd, sums, gx);
  MDL_LAUNCHED();
  return MDL_OK;
}
