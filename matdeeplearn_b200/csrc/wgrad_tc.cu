// wgrad_tc.cu -- weight / bias gradient of a dense layer over a LONG batch on the tcgen05 tensor core:
//     dW[o][i] = sum_r G[r][o] * X[r][i],    db[o] = sum_r G[r][o]          (r over R rows, R = edges or nodes)
//
// Reference: autograd's Linear backward for the edge-level MLPs of the reference's models -- the SchNet filter
// network (matdeeplearn/models/schnet.py:81 -> PyG InteractionBlock.mlp), the NNConv edge network (mpnn.py:83-85),
// the MEGNet edge / embedding MLPs (megnet.py:28-56, 222-247) -- where it is a K = E contraction (cuBLAS SIMT
// sgemm "nt" + a separate column-sum kernel: half of the SchNet / MEGNet step before this kernel).
//
// The contraction runs over the ROW index, so both tensor-core operands must be row-contiguous, i.e.
// transposed relative to how X and G sit in memory (kind::tf32 has no usable MN-major descriptor:
// profiles/r2_umma_mn_probe.txt).  Per chunk of 64 rows:
//   A = G^T  [M = 128 output channels][K = 64 rows]  in TENSOR MEMORY (lane = channel o, column = row): thread o
//            reads its column of the chunk (coalesced across lanes), splits hi / lo, tcgen05.st; the same values
//            give the bias sum for free
//   B = X^T  [N = I][K = 64 rows]  in shared memory, canonical K-major tiles hi / lo, written by scatter stores
//            (thread = row; the 16-byte chunk padding keeps them conflict-free)
//   D[o][i] += A . B^T  as 3 x 8 tcgen05.mma (3xTF32, fp32-faithful) accumulated in TMEM over the CTA's whole life
// Per-CTA partials ([O*I | O], the layout of wgrad.cu) are summed in CTA order and delivered through the same
// block map (k_wgrad_reduce): deterministic, and the result lands wherever the parameter's gradient lives.
// Two CTAs share an SM (256 TMEM columns each for I <= 128): one stages while the other's MMAs run.
#include "common.cuh"
#include "umma.cuh"

namespace mdl {

int wgrad_reduce_launch(const float* part, int nparts, int I, int O, const mdl_wgrad_out& m, cudaStream_t st);  // wgrad.cu

namespace {

constexpr int kWtThreads = 512;
constexpr int kWtRows = 64;       // rows per chunk = MMA K extent per accumulation step
constexpr int kWtM = 128;         // output channels per CTA (TMEM lanes)

__host__ __device__ inline uint32_t wt_lbo(int NI) { return (uint32_t)NI * 16 + 16; }  // k-chunk stride of the X^T tiles

__global__ void __launch_bounds__(kWtThreads, 2)
k_wgrad_tc(const float* __restrict__ X, const float* __restrict__ G, const float* __restrict__ rs, int64_t R, int I, int O,
           int NI, int tmem_cols, float* __restrict__ part) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sBias[4][kWtM];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lbo = wt_lbo(NI), tile_bytes = (kWtRows / 4) * lbo;
  uint8_t* sThi = smem;
  uint8_t* sTlo = smem + tile_bytes;
  const int o_base = blockIdx.y * kWtM;

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, (uint32_t)tmem_cols);
  if (tid == 32) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  for (uint32_t i = tid; i < 2 * tile_bytes / 4; i += kWtThreads) reinterpret_cast<float*>(smem)[i] = 0.0f;  // pad columns
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmA_hi = tmem + (uint32_t)NI, tmA_lo = tmA_hi + kWtRows;
  const uint32_t idesc = umma::make_idesc_tf32(kWtM, NI);

  // A staging: thread = (output channel o = TMEM lane, 16-row group rg)
  const int o = tid & (kWtM - 1), rg = tid >> 7;
  const bool o_ok = o_base + o < O;
  // B staging: thread = (row r of the chunk, 16-column group cg, cg + 8, ...)
  const int br = tid & (kWtRows - 1), cg0 = tid >> 6;
  const uint32_t b_off = (uint32_t)(br >> 2) * lbo + (uint32_t)(br & 3) * 4;
  const bool vec4 = (I & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0;

  const int64_t nchunks = (R + kWtRows - 1) / kWtRows;
  float accb = 0.0f;
  uint32_t phase = 0;
  bool any = false;
  for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int64_t r0 = ch * kWtRows;
    // ---- issue this chunk's global loads (nothing below waits on memory until the previous MMAs are awaited)
    float ga[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int64_t r = r0 + rg * 16 + j;
      ga[j] = (o_ok && r < R) ? __ldg(G + r * O + o_base + o) * (rs ? __ldg(rs + r) : 1.0f) : 0.0f;  // optional row scale
    }
    // the previous chunk's MMAs read the A columns and the X^T tiles: wait before overwriting them
    if (any) {
      umma::mbar_wait(&bar, phase);
      phase ^= 1;
      umma::fence_after_sync();
    }
    {
      float hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        accb += ga[j];
        hi[j] = umma::tf32_hi(ga[j]);
        lo[j] = ga[j] - hi[j];
      }
      umma::tmem_st16(umma::tmem_addr(tmA_hi, warp, 16 * rg), hi);
      umma::tmem_st16(umma::tmem_addr(tmA_lo, warp, 16 * rg), lo);
    }
    for (int cg = cg0; cg * 16 < I; cg += kWtThreads / kWtRows) {
      const int64_t r = r0 + br;
      float xv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) xv[j] = 0.0f;
      if (r < R) {
        const float* xr = X + r * I + cg * 16;
        if (vec4 && cg * 16 + 16 <= I) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xr) + q);
            xv[4 * q] = v.x; xv[4 * q + 1] = v.y; xv[4 * q + 2] = v.z; xv[4 * q + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (cg * 16 + j < I) xv[j] = __ldg(xr + j);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {  // element (i = 16 cg + j, row br): core-matrix row i % 8 of 8-row group i / 8
        const int i = cg * 16 + j;
        const float hi = umma::tf32_hi(xv[j]);
        const uint32_t off = b_off + (uint32_t)(i >> 3) * 128 + (uint32_t)(i & 7) * 16;
        *reinterpret_cast<float*>(sThi + off) = hi;
        *reinterpret_cast<float*>(sTlo + off) = xv[j] - hi;
      }
    }
    umma::tmem_st_wait();
    umma::fence_proxy_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
      const uint32_t t_hi = umma::smem_u32(sThi), t_lo = umma::smem_u32(sTlo);
#pragma unroll 1
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = (pass == 2) ? tmA_lo : tmA_hi;
        const uint32_t b = (pass == 1) ? t_lo : t_hi;
        for (int kk = 0; kk < kWtRows / 8; ++kk)
          umma::mma_tf32_ts(tmem, a + kk * 8, umma::make_desc(b + kk * 2 * lbo, lbo, 128), idesc, (any || pass || kk) ? 1u : 0u);
      }
      umma::mma_commit(&bar);
    }
    any = true;
  }
  if (any) {
    umma::mbar_wait(&bar, phase);
    umma::fence_after_sync();
  }
  // ---- the CTA's partial: weights [O][I] then the bias sums (the four row groups of a channel combined in order)
  sBias[rg][o] = accb;
  __syncthreads();
  float* mine = part + ((size_t)blockIdx.x) * ((size_t)O * I + O);
  if (tid < kWtM && o_base + tid < O)
    mine[(size_t)O * I + o_base + tid] = (sBias[0][tid] + sBias[1][tid]) + (sBias[2][tid] + sBias[3][tid]);
  {
    const int q = warp & 3, cgrp = warp >> 2;  // TMEM lane quadrant, column group (4 groups)
    const int oo = o_base + 32 * q + lane;
    for (int c0 = cgrp * 16; c0 < NI; c0 += 64) {
      float d[16];
      umma::tmem_ld16(umma::tmem_addr(tmem, q, c0), d);
      umma::tmem_ld_wait();
      if (oo < O) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < I) mine[(size_t)oo * I + c0 + j] = any ? d[j] : 0.0f;
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

}  // namespace

// shapes the tensor-core kernel takes: I <= 256, any O (blocks of 128 channels), enough rows to be worth it
bool wgrad_tc_supported(int64_t R, int I, int O) { return R >= 2048 && I >= 8 && I <= 256 && O >= 8; }

int wgrad_tc_grid(int64_t R) {
  const int64_t chunks = (R + kWtRows - 1) / kWtRows;
  return (int)std::min<int64_t>(chunks, 2 * kNumSMs);
}

int wgrad_tc_launch(const float* X, const float* G, const float* rs, int64_t R, int I, int O, const mdl_wgrad_out& out,
                    float* part, cudaStream_t st) {
  const int NI = (I + 15) & ~15;
  const int tmem_cols = (NI + 2 * kWtRows <= 256) ? 256 : 512;
  size_t smem = (size_t)2 * (kWtRows / 4) * wt_lbo(NI);
  if (tmem_cols == 512) smem = std::max<size_t>(smem, 120 * 1024);  // one CTA per SM: it holds the whole tensor memory
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MDL_CUDA(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured.store(1, std::memory_order_release);
  }
  const int grid = wgrad_tc_grid(R);
  dim3 g(grid, (O + kWtM - 1) / kWtM);
  k_wgrad_tc<<<g, kWtThreads, smem, st>>>(X, G, rs, R, I, O, NI, tmem_cols, part);
  MDL_LAUNCHED();
  return wgrad_reduce_launch(part, grid, I, O, out, st);
}

}  // namespace mdl
