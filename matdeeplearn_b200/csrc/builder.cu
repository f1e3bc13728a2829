// builder.cu -- graph builder on the GPU: the per-structure part of the
// reference's process_data (matdeeplearn/process/process.py:284-305, 540-560, 365-388, 594-605):
//
//   all-pairs distances (fp64; minimum image for an orthorhombic periodic cell, as the repo's host
//   builder does), threshold_sort(r, k): row i keeps the k+1 closest columns within r (ordinal rank,
//   ties to the lower column -- a stable sort), zero distances dropped, edges emitted row-major with
//   ascending column, one loop (i,i) of weight 0 per node appended after the structure's edges, node
//   features one-hot(Z-1, 100) ++ one-hot(out-degree incl. the loop).
//
// The reference does this in Python/numpy per structure (O(n^2) host work, the slowest stage of a
// run).  Here one CTA owns a structure: positions sit in shared memory, a warp owns a row, selects
// its k+1 minima by (distance, column) with shuffle reductions -- no sort of the whole row, no atomics
// -- and writes a fixed-width neighbour table; a second kernel scatters the table into the edge
// arrays at offsets the caller obtained from a prefix sum of the per-node counts.
//
// Arithmetic is written with explicit round-to-nearest fp64 intrinsics (no FMA contraction) so that
// distances equal numpy's bit for bit.
#include "common.cuh"

namespace mdl {

constexpr int kBuildThreads = 256;
constexpr int kBuildWarps = kBuildThreads / 32;
constexpr int kBuildMaxK = 32;   // neighbours + 1 must fit one warp

__device__ __forceinline__ double mic_delta(double a, double b, double L) {
  double d = __dsub_rn(a, b);
  if (L > 0.0) d = __dsub_rn(d, __dmul_rn(rint(__ddiv_rn(d, L)), L));
  return d;
}

// General (triclinic) cell: the shortest lattice translate of the difference vector d.  `lat` = 28 doubles per
// structure written by the host (process.lattice_record): cell rows [0..8], inverse [9..17], image-shift range per
// axis [18..20], periodicity flags [21..23], "general cell" flag [24].  The expressions -- fractional coordinates,
// wrap, Cartesian vector, image search in (i, j, k) order with a strict `<` -- are the host builder's
// (process._general_minimum_image), evaluated with explicit round-to-nearest operations in the same order.
__device__ __forceinline__ double mic_general_dist(double d0, double d1, double d2, const double* __restrict__ lat) {
  const double* cell = lat;
  const double* inv = lat + 9;
  double frac[3], w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    frac[k] = __dadd_rn(__dadd_rn(__dmul_rn(d0, inv[k]), __dmul_rn(d1, inv[3 + k])), __dmul_rn(d2, inv[6 + k]));
    if (lat[21 + k] != 0.0) frac[k] = __dsub_rn(frac[k], rint(frac[k]));
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
    w[c] = __dadd_rn(__dadd_rn(__dmul_rn(frac[0], cell[c]), __dmul_rn(frac[1], cell[3 + c])), __dmul_rn(frac[2], cell[6 + c]));
  double b0 = w[0], b1 = w[1], b2 = w[2];
  double best = __dadd_rn(__dadd_rn(__dmul_rn(b0, b0), __dmul_rn(b1, b1)), __dmul_rn(b2, b2));
  const int n0 = (int)lat[18], n1 = (int)lat[19], n2 = (int)lat[20];
  for (int i = -n0; i <= n0; ++i)
    for (int j = -n1; j <= n1; ++j)
      for (int k = -n2; k <= n2; ++k) {
        if (i == 0 && j == 0 && k == 0) continue;
        double c[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double shift = __dadd_rn(__dadd_rn(__dmul_rn((double)i, cell[a]), __dmul_rn((double)j, cell[3 + a])),
                                         __dmul_rn((double)k, cell[6 + a]));
          c[a] = __dadd_rn(w[a], shift);
        }
        const double d2v = __dadd_rn(__dadd_rn(__dmul_rn(c[0], c[0]), __dmul_rn(c[1], c[1])), __dmul_rn(c[2], c[2]));
        if (d2v < best) { best = d2v; b0 = c[0]; b1 = c[1]; b2 = c[2]; }
      }
  return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(b0, b0), __dmul_rn(b1, b1)), __dmul_rn(b2, b2)));
}

// (d, j) lexicographic minimum across the warp
__device__ __forceinline__ void warp_argmin(double& d, int& j) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, d, off);
    const int oj = __shfl_xor_sync(0xffffffffu, j, off);
    if (od < d || (od == d && oj < j)) { d = od; j = oj; }
  }
}

__global__ void __launch_bounds__(kBuildThreads)
k_build_neighbors(const double* __restrict__ pos, const double* __restrict__ cell, const double* __restrict__ lattice,
                  const int64_t* __restrict__ node_ptr, double radius, int K, int max_n,
                  int32_t* __restrict__ nbr_col, float* __restrict__ nbr_w, int32_t* __restrict__ cnt) {
  extern __shared__ __align__(16) double bsm[];
  __shared__ double sLat[28];
  double* sPos = bsm;                       // [max_n][3]
  double* sRow = bsm + 3 * (size_t)max_n;   // [warps][max_n]
  const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n0 = __ldg(node_ptr + g);
  const int n = (int)(__ldg(node_ptr + g + 1) - n0);
  const double Lx = cell ? __ldg(cell + 3 * (size_t)g) : 0.0;
  const double Ly = cell ? __ldg(cell + 3 * (size_t)g + 1) : 0.0;
  const double Lz = cell ? __ldg(cell + 3 * (size_t)g + 2) : 0.0;
  for (int i = threadIdx.x; i < 3 * n; i += kBuildThreads) sPos[i] = __ldg(pos + 3 * n0 + i);
  if (threadIdx.x < 28) sLat[threadIdx.x] = lattice ? __ldg(lattice + 28 * (size_t)g + threadIdx.x) : 0.0;
  __syncthreads();
  const bool general = sLat[24] != 0.0;
  double* row = sRow + (size_t)warp * max_n;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  for (int i = warp; i < n; i += kBuildWarps) {
    const double xi = sPos[3 * i], yi = sPos[3 * i + 1], zi = sPos[3 * i + 2];
    for (int j = lane; j < n; j += 32) {
      if (general) {
        row[j] = mic_general_dist(__dsub_rn(xi, sPos[3 * j]), __dsub_rn(yi, sPos[3 * j + 1]), __dsub_rn(zi, sPos[3 * j + 2]), sLat);
        continue;
      }
      const double dx = mic_delta(xi, sPos[3 * j], Lx);
      const double dy = mic_delta(yi, sPos[3 * j + 1], Ly);
      const double dz = mic_delta(zi, sPos[3 * j + 2], Lz);
      row[j] = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    }
    __syncwarp();
    // K rounds of extract-min; lane r keeps the r-th selected (column, distance)
    int my_col = -1;
    double my_d = 0.0;
    int kept = 0;   // selected entries with non-zero distance (warp-uniform)
    for (int r = 0; r < K; ++r) {
      double bd = inf;
      int bj = 0x7fffffff;
      for (int j = lane; j < n; j += 32) {
        const double d = row[j];
        if (d < bd) { bd = d; bj = j; }   // ascending j per lane: first minimum wins
      }
      warp_argmin(bd, bj);
      if (!(bd <= radius)) break;          // also ends when nothing is left (inf)
      if (lane == (bj & 31)) row[bj] = inf;
      __syncwarp();
      if (bd != 0.0) {
        if (lane == kept) { my_col = bj; my_d = bd; }
        ++kept;
      }
    }
    // emit in ascending column order
    int rank = 0;
    for (int t = 0; t < kept; ++t) {
      const int oc = __shfl_sync(0xffffffffu, my_col, t);
      if (lane < kept && oc < my_col) ++rank;
    }
    const int64_t node = n0 + i;
    if (lane < kept) {
      nbr_col[node * K + rank] = my_col;
      nbr_w[node * K + rank] = (float)my_d;
    }
    if (lane == 0) cnt[node] = kept;
    __syncwarp();
  }
}

// neighbour table -> edge arrays (store-global node ids) + loops + node features.
// first_edge[node] = position of the node's first edge; loop_pos[node] = position of its loop.
__global__ void k_build_emit(const int32_t* __restrict__ nbr_col, const float* __restrict__ nbr_w,
                             const int32_t* __restrict__ cnt, const int64_t* __restrict__ first_edge,
                             const int64_t* __restrict__ loop_pos, const int64_t* __restrict__ node_graph_start,
                             const int32_t* __restrict__ numbers, int64_t num_nodes, int K, int z_width, int F,
                             int32_t* __restrict__ src, int32_t* __restrict__ dst, float* __restrict__ w,
                             float* __restrict__ x) {
  for (int64_t node = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; node < num_nodes;
       node += (int64_t)gridDim.x * blockDim.x) {
    const int c = __ldg(cnt + node);
    const int64_t base = __ldg(first_edge + node), g0 = __ldg(node_graph_start + node);
    for (int t = 0; t < c; ++t) {
      src[base + t] = (int32_t)node;
      dst[base + t] = (int32_t)(g0 + __ldg(nbr_col + node * K + t));
      w[base + t] = __ldg(nbr_w + node * K + t);
    }
    const int64_t lp = __ldg(loop_pos + node);
    src[lp] = (int32_t)node;
    dst[lp] = (int32_t)node;
    w[lp] = 0.0f;
    float* xr = x + node * F;   // x is zero-filled by the caller
    const int z = __ldg(numbers + node);
    if (z >= 1 && z <= z_width) xr[z - 1] = 1.0f;
    const int deg = c + 1;      // out-degree including the loop
    if (z_width + deg < F) xr[z_width + deg] = 1.0f;
  }
}

}  // namespace mdl

using namespace mdl;

extern "C" int mdl_build_neighbors(const double* pos, const double* cell, const int64_t* node_ptr,
                                   int64_t num_graphs, int32_t max_nodes, double radius, int32_t neighbors,
                                   int32_t* nbr_col, float* nbr_w, int32_t* cnt, void* stream) {
  return mdl_build_neighbors_lattice(pos, cell, nullptr, node_ptr, num_graphs, max_nodes, radius, neighbors, nbr_col,
                                     nbr_w, cnt, stream);
}

extern "C" int mdl_build_neighbors_lattice(const double* pos, const double* cell, const double* lattice,
                                           const int64_t* node_ptr, int64_t num_graphs, int32_t max_nodes,
                                           double radius, int32_t neighbors, int32_t* nbr_col, float* nbr_w,
                                           int32_t* cnt, void* stream) {
  MDL_REQUIRE(num_graphs >= 0 && max_nodes >= 0 && neighbors >= 0, "build_neighbors: bad shape");
  if (num_graphs == 0) return MDL_OK;
  MDL_REQUIRE(pos && node_ptr && nbr_col && nbr_w && cnt, "build_neighbors: null pointer");
  const int K = neighbors + 1;
  MDL_REQUIRE(K <= kBuildMaxK, "build_neighbors: at most %d neighbours", kBuildMaxK - 1);
  const size_t smem = ((size_t)3 * max_nodes + (size_t)kBuildWarps * max_nodes) * sizeof(double);
  MDL_REQUIRE(smem <= 200 * 1024, "build_neighbors: structures of more than %d atoms need a tiled builder",
              (int)(200 * 1024 / (sizeof(double) * (3 + kBuildWarps))));
  MDL_REQUIRE(num_graphs <= 0x7fffffff, "build_neighbors: too many structures");
  static bool attr_set = false;
  if (!attr_set) {
    MDL_CUDA(cudaFuncSetAttribute(k_build_neighbors, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  k_build_neighbors<<<(unsigned)num_graphs, kBuildThreads, smem, as_stream(stream)>>>(
      pos, cell, lattice, node_ptr, radius, K, max_nodes, nbr_col, nbr_w, cnt);
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_build_emit(const int32_t* nbr_col, const float* nbr_w, const int32_t* cnt,
                              const int64_t* first_edge, const int64_t* loop_pos,
                              const int64_t* node_graph_start, const int32_t* numbers, int64_t num_nodes,
                              int32_t neighbors, int32_t z_width, int32_t F, int32_t* src, int32_t* dst,
                              float* w, float* x, void* stream) {
  MDL_REQUIRE(num_nodes >= 0 && neighbors >= 0 && z_width > 0 && F >= z_width + neighbors + 2,
              "build_emit: bad shape");
  if (num_nodes == 0) return MDL_OK;
  MDL_REQUIRE(nbr_col && nbr_w && cnt && first_edge && loop_pos && node_graph_start && numbers && src && dst &&
                  w && x, "build_emit: null pointer");
  const int grid = (int)std::min<int64_t>(ceil_div<int64_t>(num_nodes, 256), 8 * kNumSMs);
  k_build_emit<<<grid, 256, 0, as_stream(stream)>>>(nbr_col, nbr_w, cnt, first_edge, loop_pos, node_graph_start,
                                                     numbers, num_nodes, neighbors + 1, z_width, F, src, dst, w, x);
  MDL_LAUNCHED();
  return MDL_OK;
}
