// csr.cu -- once-per-batch graph layout for the engine (include/mdl_b200.h,
// mdl_csr_from_coo).  Replaces the per-layer gather/scatter index bookkeeping
// PyG's MessagePassing does for every conv call of the reference
// (matdeeplearn/models/cgcnn.py:142) with two stable radix sorts:
//   slots      = edges sorted by destination (col), ties in reference order
//   positions  = slots sorted by source (row),      ties in slot order
// Deterministic (no atomics), fully asynchronous, graph-capturable.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace mdl {

__global__ void k_extract_dst(const int64_t* __restrict__ edge_index, int64_t E,
                              int32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < E) {
    keys[e] = (int32_t)edge_index[E + e];  // col = destination
    vals[e] = (int32_t)e;
  }
}

__global__ void k_extract_src(const int64_t* __restrict__ edge_index, int64_t E,
                              const int32_t* __restrict__ dst_eid, int32_t* __restrict__ dst_src,
                              int32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s < E) {
    int32_t src = (int32_t)edge_index[dst_eid[s]];  // row = source
    dst_src[s] = src;
    keys[s] = src;
    vals[s] = (int32_t)s;
  }
}

// ptr[n] = first position whose key >= n, for sorted keys; ptr has num_seg+1 entries.
template <typename K>
__global__ void k_boundaries(const K* __restrict__ keys, int64_t count, int64_t num_seg,
                             int32_t* __restrict__ ptr) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s > count) return;
  int64_t lo = (s == 0) ? 0 : (int64_t)keys[s - 1] + 1;
  int64_t hi = (s == count) ? num_seg : (int64_t)keys[s];
  if (hi > num_seg) hi = num_seg;
  for (int64_t n = lo; n <= hi; ++n) ptr[n] = (int32_t)s;
}

__global__ void k_inv_deg(const int32_t* __restrict__ ptr, int64_t N, float* __restrict__ inv) {
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n < N) {
    int d = ptr[n + 1] - ptr[n];
    inv[n] = 1.0f / (float)(d > 1 ? d : 1);
  }
}

static int key_bits(int64_t n) {
  int b = 1;
  while (b < 31 && (1LL << b) < n) ++b;
  return b;
}

static size_t cub_temp_bytes(int64_t E, int bits) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)E, 0, bits);
  return bytes;
}

}  // namespace mdl

using namespace mdl;

extern "C" size_t mdl_csr_workspace_bytes(int64_t N, int64_t E) {
  if (E < 0 || N < 0 || E >= (1LL << 31) || N >= (1LL << 31)) return 0;
  size_t e4 = align_up((size_t)(E > 0 ? E : 1) * 4, 256);
  return 3 * e4 + align_up(cub_temp_bytes(E > 0 ? E : 1, key_bits(N)), 256) + 256;
}

extern "C" int mdl_csr_from_coo(const int64_t* edge_index, const int64_t* batch, int64_t N,
                                int64_t E, int64_t B, int32_t* dst_ptr, int32_t* dst_src,
                                int32_t* dst_dst, int32_t* dst_eid, int32_t* src_ptr,
                                int32_t* src_slot, float* inv_deg_dst, float* inv_deg_src,
                                int32_t* graph_ptr, void* workspace, size_t workspace_bytes,
                                void* stream_) {
  cudaStream_t st = as_stream(stream_);
  MDL_REQUIRE(N >= 0 && E >= 0 && N < (1LL << 31) && E < (1LL << 31), "csr: N/E out of int32 range");
  MDL_REQUIRE(dst_ptr && src_ptr && inv_deg_dst && inv_deg_src, "csr: null output");
  MDL_REQUIRE(E == 0 || (edge_index && dst_src && dst_dst && dst_eid && src_slot), "csr: null edge buffers");
  MDL_REQUIRE((batch && graph_ptr) || B == 0, "csr: batch/graph_ptr null with num_graphs > 0");
  if (workspace_bytes < mdl_csr_workspace_bytes(N, E)) {
    set_error("csr: workspace %zu < %zu", workspace_bytes, mdl_csr_workspace_bytes(N, E));
    return MDL_ERR_WORKSPACE;
  }
  const int T = 256;
  size_t e4 = align_up((size_t)(E > 0 ? E : 1) * 4, 256);
  char* w = (char*)workspace;
  int32_t* keys_a = (int32_t*)w;
  int32_t* vals_a = (int32_t*)(w + e4);
  int32_t* keys_b = (int32_t*)(w + 2 * e4);
  void* cub_tmp = w + 3 * e4;
  if (E > 0) {
    int bits = key_bits(N);
    size_t cub_bytes = cub_temp_bytes(E, bits);
    int grid = (int)ceil_div<int64_t>(E, T);
    // slots: stable sort of reference edge ids by destination
    k_extract_dst<<<grid, T, 0, st>>>(edge_index, E, keys_a, vals_a);
    MDL_LAUNCHED();
    MDL_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_a, dst_dst, vals_a, dst_eid,
                                             (int)E, 0, bits, st));
    // positions: stable sort of slots by source
    k_extract_src<<<grid, T, 0, st>>>(edge_index, E, dst_eid, dst_src, keys_a, vals_a);
    MDL_LAUNCHED();
    MDL_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_a, keys_b, vals_a, src_slot,
                                             (int)E, 0, bits, st));
  }
  {
    int grid = (int)ceil_div<int64_t>(E + 1, T);
    k_boundaries<int32_t><<<grid, T, 0, st>>>(dst_dst, E, N, dst_ptr);
    MDL_LAUNCHED();
    k_boundaries<int32_t><<<grid, T, 0, st>>>(keys_b, E, N, src_ptr);
    MDL_LAUNCHED();
  }
  if (N > 0) {
    int grid = (int)ceil_div<int64_t>(N, T);
    k_inv_deg<<<grid, T, 0, st>>>(dst_ptr, N, inv_deg_dst);
    MDL_LAUNCHED();
    k_inv_deg<<<grid, T, 0, st>>>(src_ptr, N, inv_deg_src);
    MDL_LAUNCHED();
  }
  if (graph_ptr) {
    int grid = (int)ceil_div<int64_t>(N + 1, T);
    k_boundaries<int64_t><<<grid, T, 0, st>>>(batch, N, B, graph_ptr);
    MDL_LAUNCHED();
  }
  return MDL_OK;
}

namespace mdl {
__global__ void k_gather_rows(const float* __restrict__ src, const int32_t* __restrict__ idx,
                              float* __restrict__ out, int64_t rows, int64_t width, bool scatter) {
  int64_t total = rows * width;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / width, c = i - r * width;
    int64_t o = idx[r];
    if (scatter) out[o * width + c] = src[i];
    else out[i] = src[o * width + c];
  }
}
}  // namespace mdl

static int launch_rows(const float* src, const int32_t* idx, float* out, int64_t rows,
                       int64_t width, bool scatter, void* stream) {
  MDL_REQUIRE(rows >= 0 && width > 0, "gather_rows: bad shape");
  if (rows == 0) return MDL_OK;
  MDL_REQUIRE(src && idx && out, "gather_rows: null pointer");
  int64_t total = rows * width;
  int grid = (int)std::min<int64_t>(ceil_div<int64_t>(total, 256), (int64_t)kNumSMs * 16);
  k_gather_rows<<<grid, 256, 0, as_stream(stream)>>>(src, idx, out, rows, width, scatter);
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_gather_rows(const float* src, const int32_t* idx, float* out, int64_t rows,
                               int64_t width, void* stream) {
  return launch_rows(src, idx, out, rows, width, false, stream);
}
extern "C" int mdl_scatter_rows(const float* src, const int32_t* idx, float* out, int64_t rows,
                                int64_t width, void* stream) {
  return launch_rows(src, idx, out, rows, width, true, stream);
}

namespace mdl {
__global__ void k_gaussian_smear(const float* __restrict__ d, const float* __restrict__ offset,
                                 float* __restrict__ out, int64_t E, int G, float coeff) {
  int64_t total = E * G;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t e = i / G;
    int k = (int)(i - e * G);
    float diff = d[e] - __ldg(offset + k);
    out[i] = expf(coeff * (diff * diff));
  }
}
}  // namespace mdl

extern "C" int mdl_gaussian_smear(const float* d, const float* offset, float* out, int64_t E,
                                  int32_t G, float coeff, void* stream) {
  MDL_REQUIRE(E >= 0 && G > 0, "gaussian_smear: bad shape");
  if (E == 0) return MDL_OK;
  MDL_REQUIRE(d && out && offset, "gaussian_smear: null pointer");
  int64_t total = E * (int64_t)G;
  int grid = (int)std::min<int64_t>(ceil_div<int64_t>(total, 256), (int64_t)kNumSMs * 16);
  k_gaussian_smear<<<grid, 256, 0, as_stream(stream)>>>(d, offset, out, E, G, coeff);
  MDL_LAUNCHED();
  return MDL_OK;
}
