// segment.cu -- warp-per-segment reduction over contiguous row ranges.
// Replaces torch_scatter.scatter(reduce=sum|mean|max) and PyG global_*_pool at
// reference matdeeplearn/models/cgcnn.py:154,169 and megnet.py:86,130-132,346-348.
// `batch` / CSR order make every segment a contiguous range, so the reduction is
// a deterministic sequential sum per (segment, channel) with lanes across
// channels (coalesced 128 B row reads) -- no atomics.
#include "common.cuh"

namespace mdl {

template <int REDUCE>
__global__ void __launch_bounds__(256)
k_segment_fwd(const float* __restrict__ src, const int32_t* __restrict__ ptr,
              const int32_t* __restrict__ perm, float* __restrict__ out,
              int32_t* __restrict__ argmax, int64_t S, int width) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = warp; s < S; s += nwarps) {
    const int lo = __ldg(ptr + s), hi = __ldg(ptr + s + 1);
    for (int c0 = 0; c0 < width; c0 += 128) {
      // each lane owns up to 4 channels strided by 32
      float acc[4];
      int arg[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[j] = (REDUCE == MDL_REDUCE_MAX) ? -INFINITY : 0.0f; arg[j] = -1; }
      // rows in batches of four: all loads of a batch are issued before the first is consumed (the
      // sum stays in row order, so the result does not depend on the batching)
      for (int r0 = lo; r0 < hi; r0 += 4) {
        float v[4][4];
        int64_t rows[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + u;
          rows[u] = (r < hi) ? (perm ? (int64_t)__ldg(perm + r) : (int64_t)r) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float* p = src + (rows[u] < 0 ? 0 : rows[u]) * width + c0 + lane;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            v[u][j] = (rows[u] >= 0 && c0 + lane + 32 * j < width) ? __ldg(p + 32 * j) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (rows[u] < 0) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (c0 + lane + 32 * j < width) {
              if (REDUCE == MDL_REDUCE_MAX) {
                if (v[u][j] > acc[j]) { acc[j] = v[u][j]; arg[j] = (int)rows[u]; }
              } else {
                acc[j] += v[u][j];
              }
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + lane + 32 * j;
        if (c < width) {
          float v = acc[j];
          if (REDUCE == MDL_REDUCE_MEAN) v *= 1.0f / (float)(hi - lo > 1 ? hi - lo : 1);
          if (REDUCE == MDL_REDUCE_MAX && hi == lo) v = 0.0f;
          out[s * width + c] = v;
          if (REDUCE == MDL_REDUCE_MAX && argmax) argmax[s * width + c] = arg[j];
        }
      }
    }
  }
}

template <int REDUCE>
__global__ void __launch_bounds__(256)
k_segment_bwd(const float* __restrict__ gout, const int32_t* __restrict__ ptr,
              const int32_t* __restrict__ perm, float* __restrict__ gsrc, int64_t S, int width,
              int64_t num_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  if (!perm) {  // rows past the last segment (capacity padding) belong to no segment: zero gradient
    const int64_t first = (int64_t)__ldg(ptr + S) * width, total = num_rows * width;
    for (int64_t i = first + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x)
      gsrc[i] = 0.0f;
  }
  for (int64_t s = warp; s < S; s += nwarps) {
    const int lo = __ldg(ptr + s), hi = __ldg(ptr + s + 1);
    const float w = (REDUCE == MDL_REDUCE_MEAN) ? 1.0f / (float)(hi - lo > 1 ? hi - lo : 1) : 1.0f;
    for (int c = lane; c < width; c += 32) {
      const float g = __ldg(gout + s * width + c) * w;
      for (int r = lo; r < hi; ++r) {
        const int64_t row = perm ? __ldg(perm + r) : r;
        gsrc[row * width + c] = g;
      }
    }
  }
}

__global__ void k_segment_bwd_max(const float* __restrict__ gout, const int32_t* __restrict__ argmax,
                                  float* __restrict__ gsrc, int64_t total, int width) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int a = argmax[i];
    if (a >= 0) gsrc[(int64_t)a * width + (i % width)] = gout[i];
  }
}

static int seg_grid(int64_t S) {
  int64_t blocks = ceil_div<int64_t>(S, 8);
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  return (int)(blocks > 0 ? blocks : 1);
}

}  // namespace mdl

using namespace mdl;

extern "C" int mdl_segment_reduce_fwd(const float* src, const int32_t* ptr, const int32_t* perm,
                                      float* out, int32_t* argmax, int64_t S, int64_t width,
                                      int32_t reduce, void* stream) {
  MDL_REQUIRE(S >= 0 && width > 0 && width < (1 << 30), "segment_reduce_fwd: bad shape");
  if (S == 0) return MDL_OK;
  MDL_REQUIRE(ptr && out, "segment_reduce_fwd: null pointer");
  cudaStream_t st = as_stream(stream);
  int grid = seg_grid(S);
  switch (reduce) {
    case MDL_REDUCE_SUM:
      k_segment_fwd<MDL_REDUCE_SUM><<<grid, 256, 0, st>>>(src, ptr, perm, out, nullptr, S, (int)width);
      break;
    case MDL_REDUCE_MEAN:
      k_segment_fwd<MDL_REDUCE_MEAN><<<grid, 256, 0, st>>>(src, ptr, perm, out, nullptr, S, (int)width);
      break;
    case MDL_REDUCE_MAX:
      k_segment_fwd<MDL_REDUCE_MAX><<<grid, 256, 0, st>>>(src, ptr, perm, out, argmax, S, (int)width);
      break;
    default:
      MDL_REQUIRE(false, "segment_reduce_fwd: unknown reduce %d", reduce);
  }
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_segment_reduce_bwd(const float* gout, const int32_t* ptr, const int32_t* perm,
                                      const int32_t* argmax, float* gsrc, int64_t S,
                                      int64_t num_rows, int64_t width, int32_t reduce,
                                      void* stream) {
  MDL_REQUIRE(S >= 0 && width > 0 && num_rows >= 0, "segment_reduce_bwd: bad shape");
  if (num_rows == 0) return MDL_OK;
  MDL_REQUIRE(gsrc, "segment_reduce_bwd: null pointer");
  cudaStream_t st = as_stream(stream);
  if (S == 0) {
    MDL_CUDA(cudaMemsetAsync(gsrc, 0, (size_t)num_rows * width * 4, st));
    return MDL_OK;
  }
  MDL_REQUIRE(ptr && gout, "segment_reduce_bwd: null pointer");
  int grid = seg_grid(S);
  switch (reduce) {
    case MDL_REDUCE_SUM:
      k_segment_bwd<MDL_REDUCE_SUM><<<grid, 256, 0, st>>>(gout, ptr, perm, gsrc, S, (int)width, num_rows);
      break;
    case MDL_REDUCE_MEAN:
      k_segment_bwd<MDL_REDUCE_MEAN><<<grid, 256, 0, st>>>(gout, ptr, perm, gsrc, S, (int)width, num_rows);
      break;
    case MDL_REDUCE_MAX: {
      MDL_REQUIRE(argmax, "segment_reduce_bwd: max needs argmax");
      MDL_CUDA(cudaMemsetAsync(gsrc, 0, (size_t)num_rows * width * 4, st));
      int64_t total = S * width;
      int g2 = (int)std::min<int64_t>(ceil_div<int64_t>(total, 256), (int64_t)kNumSMs * 16);
      k_segment_bwd_max<<<g2, 256, 0, st>>>(gout, argmax, gsrc, total, (int)width);
      break;
    }
    default:
      MDL_REQUIRE(false, "segment_reduce_bwd: unknown reduce %d", reduce);
  }
  MDL_LAUNCHED();
  return MDL_OK;
}
