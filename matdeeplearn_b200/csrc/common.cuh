// common.cuh -- shared helpers for libmdl_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/mdl_b200.h"

namespace mdl {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define MDL_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      mdl::set_error(__VA_ARGS__);        \
      return MDL_ERR_ARG;                 \
    }                                     \
  } while (0)

#define MDL_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      mdl::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return MDL_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

// after every kernel launch: count it and surface launch-time errors
#define MDL_LAUNCHED()                   \
  do {                                   \
    mdl::g_launches.fetch_add(1, std::memory_order_relaxed); \
    MDL_CUDA(cudaPeekAtLastError());     \
  } while (0)

constexpr int kNumSMs = 148;  // B200

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// first n in [0, len] with ptr[n] >= key  (ptr non-decreasing)
__device__ inline int lower_bound_i32(const int32_t* __restrict__ ptr, int len, int key) {
  int lo = 0, hi = len;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(ptr + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// F.softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplusf_(float x) {
  return x > 20.0f ? x : log1pf(expf(x));
}

}  // namespace mdl
