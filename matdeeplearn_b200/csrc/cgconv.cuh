// cgconv.cuh -- declarations shared by the SIMT (cgconv.cu) and tensor-core
// (cgconv_tc.cu) implementations of the fused CGConv edge kernels.
#pragma once
#include "common.cuh"

namespace mdl {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxDynSmem = 226 * 1024;  // 227 KB opt-in limit minus static __shared__

enum CgMode { CG_FWD = 0, CG_BWD_DST = 1, CG_BWD_SRC = 2 };

struct CgParams {
  const float* x;         // FWD  [N,C]
  const float* gout;      // BWD  [N,C]
  const float* PQ;        // [N,4C]
  const float* ea;        // [E,G] slot order (nullptr in the smearing-fused form)
  const float* dhat;      // smearing-fused form: [E] slot order, normalised distances; the Gaussian basis
  const float* sm_offset; //   e[k] = exp(sm_coeff (dhat - sm_offset[k])^2), sm_offset = linspace (uniform spacing), [G]
  float sm_coeff;         //   (reference GaussianSmearing, process.py:580-590) is expanded inside the kernel
  const float* WeT;       // [G,2C] (k-major: f channels then s channels)
  const int32_t* seg_ptr; // dst_ptr (FWD, BWD_DST) or src_ptr (BWD_SRC)  [N+1]
  const int32_t* dst_src; // [E] slot -> source node
  const int32_t* dst_dst; // [E] slot -> destination node
  const int32_t* src_slot;// [E] by-source position -> slot (BWD_SRC)
  const float* inv_deg;   // [N] destination 1/deg (mean) or nullptr (sum)
  float* out;             // FWD: out [N,C]; BWD: dPQ [N,4C]
  float* dW_part;         // BWD_DST: [gridDim.x][G][2*CC] partials
  int N, E, C, G;
  int c_off, CC;          // channel chunk handled by this launch
  int c_skip;             // tensor-core backward: leading chunk channels whose dQ the previous chunk's launch added
  int cap, te, n_tiles;
};

// First segment whose start offset seg_ptr[n] is >= position s (what a lower_bound over
// seg_ptr returns), found through the position -> segment maps the CSR already holds:
// 2-3 dependent loads instead of a ~20-step binary search per tile.
template <int MODE>
__device__ __forceinline__ int first_segment_at_or_after(const CgParams& p, int s) {
  if (s >= p.E) return (p.E == 0 && s == 0) ? 0 : p.N;
  const int slot = (MODE == CG_BWD_SRC) ? __ldg(p.src_slot + s) : s;
  const int seg = (MODE == CG_BWD_SRC) ? __ldg(p.dst_src + slot) : __ldg(p.dst_dst + slot);
  int n = (__ldg(p.seg_ptr + seg) == s) ? seg : seg + 1;
  while (n > 0 && __ldg(p.seg_ptr + n - 1) >= s) --n;  // empty segments that start exactly at s
  return n;
}

// tile t owns segments [bounds0, bounds1): computed by two threads in parallel
template <int MODE>
__device__ __forceinline__ void tile_bounds(const CgParams& p, int tile, int te, int tid, int* sh_bounds) {
  if (tid == 0) sh_bounds[0] = first_segment_at_or_after<MODE>(p, tile * te);
  if (tid == 32)
    sh_bounds[1] = (tile == p.n_tiles - 1) ? p.N : first_segment_at_or_after<MODE>(p, (tile + 1) * te);
}

// 1/(1+exp(-x)) with an approximate (<= 1 ulp) reciprocal instead of an IEEE division
__device__ __forceinline__ float sigmoid_fast_(float x) { return __fdividef(1.0f, 1.0f + expf(-x)); }

// tensor-core path (cgconv_tc.cu).  Returns false if (C, G, mode) does not fit its
// shared-memory / TMEM plan, in which case the caller uses the SIMT kernel.
bool cgtc_supported(int mode, int C, int G);
int cgtc_launch(int mode, CgParams p, cudaStream_t st, int* grid_out, int dq_atomic = 0);
void cgtc_set_phase_buffer(unsigned long long* dev_ptr);
// software-pipelined forward kernel (cgconv_fwd.cu): contraction one round ahead of the epilogue
bool cgfwd_supported(const CgParams& p);
int cgfwd_launch(CgParams p, cudaStream_t st);
void cgfwd_set_phase_buffer(unsigned long long* dev_ptr);
// warp-specialised forward kernel (cgconv_fwd_ws.cu): staging, MMA issue and epilogue on separate warps
bool cgws_supported(const CgParams& p);
int cgws_launch(CgParams p, cudaStream_t st);
void cgws_set_phase_buffer(unsigned long long* dev_ptr);
void lin_set_phase_buffer(unsigned long long* dev_ptr);  // linear_tc.cu (development aid)
// single-pass backward with dW_e on tcgen05 (cgconv_bwd.cu)
bool cgbwd_supported(const CgParams& p);
int cgbwd_launch(CgParams p, cudaStream_t st, float* dWeT);
// dWeT[k][chunk columns] = sum over the nparts per-CTA partials ([nparts][G][2 CC], f then s), fixed order (cgconv.cu)
int reduce_dw_partials(const float* part, int nparts, int G, int C, int c_off, int CC, float* dWeT, cudaStream_t st);
void cgbwd_set_phase_buffer(unsigned long long* dev_ptr);
// out0[i] (i < len0) / out1[i - len0] = sum over nparts partial vectors, fixed order (cgconv.cu)
int sum_partials(const float* part, int nparts, int64_t stride, int64_t len, float* out0, int64_t len0,
                 float* out1, cudaStream_t st);

}  // namespace mdl
