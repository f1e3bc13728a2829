// cgconv_tc.cu -- tensor-core (tcgen05 / TMEM) implementation of the fused CGConv
// edge kernels declared in cgconv.cuh.  Same tile-ownership scheme, same
// epilogue and same deterministic segmented reduction as the SIMT kernel in
// cgconv.cu; the per-edge contraction  [128 slots x G] . [G x 2C]  moves to the
// 5th-gen tensor core:
//
//   stage   ea rows  --cp.async-->  smem row-major  --split-->  A_hi / A_lo tiles
//           (canonical K-major no-swizzle layout, umma.cuh)
//   MMA     one thread issues 3 x (KP/8) tcgen05.mma kind::tf32 (3xTF32: hi*hi +
//           hi*lo + lo*hi, fp32-faithful) into a [128 x NP] fp32 accumulator in TMEM,
//           tcgen05.commit -> mbarrier
//   epilog  8 warps read the accumulator with tcgen05.ld (lane = slot), add the
//           gathered node projections P[dst] + Q[src], evaluate the gates, park the
//           per-slot values in smem
//   reduce  warp-per-segment sums in slot order -> coalesced stores (no atomics)
//
// Weights (W_e, split hi/lo once per CTA) stay resident in smem for the CTA's life.
#include "cgconv.cuh"
#include "umma.cuh"

namespace mdl {

constexpr int kTcRows = 128;  // slots per round = MMA M
constexpr int kTcTE = 112;    // ownership granularity (leaves head-room for straddling segments)

struct TcPlan {
  int NP, KP, GS, VW, tmem_cols, nitem;
  uint32_t offBhi, offBlo, offAhi, offAlo, offEA, offV, offIdx, total;
};

static bool tc_plan(int mode, int C, int G, TcPlan* pl) {
  if (C < 8 || C % 4 != 0) return false;
  const int NP = (2 * C + 15) & ~15;
  if (NP > 256) return false;
  const int KP = (G + 7) & ~7;
  int GS = (G + 3) & ~3;
  if (((GS >> 2) & 1) == 0) GS += 4;  // odd number of 16-byte chunks per row: conflict-free float4 column reads
  const int VW = (mode == CG_FWD ? C : 2 * C) + 4;
  uint32_t b = (uint32_t)NP * KP * 4, a = (uint32_t)kTcRows * KP * 4;
  uint32_t ea = (uint32_t)kTcRows * GS * 4, v = (uint32_t)kTcRows * VW * 4, idx = 3 * kTcRows * 4;
  pl->NP = NP; pl->KP = KP; pl->GS = GS; pl->VW = VW;
  pl->tmem_cols = 32;
  while (pl->tmem_cols < NP) pl->tmem_cols <<= 1;
  const int n_dw = (2 * C / 4) * ((GS + 7) / 8);
  pl->nitem = (n_dw + kThreads - 1) / kThreads;
  if (mode == CG_BWD_DST && pl->nitem > 2) return false;
  pl->offBhi = 0; pl->offBlo = b; pl->offAhi = 2 * b; pl->offAlo = 2 * b + a; pl->offEA = 2 * b + 2 * a;
  // value tile: own region if it fits, else aliased over the A tiles (dead once the MMAs retired)
  // and, outside BWD_DST (whose dWe pass still reads the row-major ea copy), over the ea copy too.
  uint32_t end = pl->offEA + ea;
  if (end + v + idx <= (uint32_t)kMaxDynSmem) {
    pl->offV = end; pl->offIdx = end + v; pl->total = end + v + idx;
    return true;
  }
  const uint32_t alias_room = 2 * a + (mode == CG_BWD_DST ? 0 : ea);
  if (v <= alias_room && end + idx <= (uint32_t)kMaxDynSmem) {
    pl->offV = pl->offAhi; pl->offIdx = end; pl->total = end + idx;
    return true;
  }
  return false;
}

__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(umma::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(umma::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int MODE, int NITEM>
__global__ void __launch_bounds__(kThreads, 1) k_cgconv_tc(const CgParams p, const TcPlan pl) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ int sh_bounds[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, G = p.G, W2 = 2 * C;
  const int NP = pl.NP, KP = pl.KP, GS = pl.GS, VW = pl.VW;

  uint8_t* sBhi = smem + pl.offBhi;
  uint8_t* sBlo = smem + pl.offBlo;
  uint8_t* sAhi = smem + pl.offAhi;
  uint8_t* sAlo = smem + pl.offAlo;
  float* sEA = reinterpret_cast<float*>(smem + pl.offEA);  // [128][GS] row-major, pads zero
  float* sV = reinterpret_cast<float*>(smem + pl.offV);    // [128][VW]
  int* sSrc = reinterpret_cast<int*>(smem + pl.offIdx);
  int* sDst = sSrc + kTcRows;
  int* sSlot = sDst + kTcRows;

  // ---- one-time setup: TMEM, barrier, resident weight tiles (hi/lo split)
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, (uint32_t)pl.tmem_cols);
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < NP * KP; i += kThreads) {
    const int n = i % NP, k = i / NP;  // consecutive threads -> consecutive columns of WeT (coalesced)
    const float w = (k < G && n < W2) ? __ldg(p.WeT + (size_t)k * W2 + n) : 0.0f;
    const float hi = umma::tf32_hi(w);
    const int off = umma::tile_offset_bytes(n, k, NP);
    *reinterpret_cast<float*>(sBhi + off) = hi;
    *reinterpret_cast<float*>(sBlo + off) = w - hi;
  }
  for (int i = tid; i < kTcRows * GS; i += kThreads) sEA[i] = 0.0f;

  float dw[NITEM][4][8];
  if (MODE == CG_BWD_DST) {
#pragma unroll
    for (int j = 0; j < NITEM; ++j)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) dw[j][a][b] = 0.0f;
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma::make_idesc_tf32(kTcRows, NP);
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    tile_bounds<MODE>(p, tile, kTcTE, tid, sh_bounds);
    __syncthreads();
    const int n_lo = sh_bounds[0], n_hi = sh_bounds[1];
    if (n_hi <= n_lo) { __syncthreads(); continue; }
    const int e_lo = __ldg(p.seg_ptr + n_lo), e_hi = __ldg(p.seg_ptr + n_hi);
    const int rounds = max(1, (e_hi - e_lo + kTcRows - 1) / kTcRows);

    for (int rd = 0; rd < rounds; ++rd) {
      const int r_lo = e_lo + rd * kTcRows;
      const int r_hi = min(e_hi, r_lo + kTcRows);
      const int cnt = r_hi - r_lo;

      // ---- stage 1: indices, then raw ea rows (coalesced async copies, row-major)
      if (tid < kTcRows) {
        int slot = 0, s = 0, d = 0;
        if (tid < cnt) {
          slot = (MODE == CG_BWD_SRC) ? __ldg(p.src_slot + r_lo + tid) : (r_lo + tid);
          s = __ldg(p.dst_src + slot);
          d = __ldg(p.dst_dst + slot);
        }
        sSlot[tid] = slot; sSrc[tid] = s; sDst[tid] = d;
      }
      if (MODE == CG_BWD_SRC) __syncthreads();
      for (int e = warp; e < cnt; e += kWarps) {
        const int slot = (MODE == CG_BWD_SRC) ? sSlot[e] : (r_lo + e);
        const float* row = p.ea + (size_t)slot * G;
        float* dst = sEA + e * GS;
        if ((G & 1) == 0) {
          for (int k2 = lane; k2 < (G >> 1); k2 += 32) cp_async8(dst + 2 * k2, row + 2 * k2);
        } else {
          for (int k = lane; k < G; k += 32) cp_async4(dst + k, row + k);
        }
      }
      cp_async_wait_all();
      __syncthreads();

      // ---- stage 2: split hi/lo into the canonical MMA operand layout
      {
        const int e = tid & (kTcRows - 1);
        const uint32_t row_off = (uint32_t)(e >> 3) * 128 + (uint32_t)(e & 7) * 16;
        for (int j = (tid >> 7); j < (KP >> 2); j += 2) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e < cnt && 4 * j < GS) {
            v = *reinterpret_cast<const float4*>(sEA + e * GS + 4 * j);
            // columns >= G are padding: force exact zeros (the value tile may alias this buffer)
            if (4 * j + 0 >= G) v.x = 0.f;
            if (4 * j + 1 >= G) v.y = 0.f;
            if (4 * j + 2 >= G) v.z = 0.f;
            if (4 * j + 3 >= G) v.w = 0.f;
          }
          float4 hi;
          hi.x = umma::tf32_hi(v.x); hi.y = umma::tf32_hi(v.y);
          hi.z = umma::tf32_hi(v.z); hi.w = umma::tf32_hi(v.w);
          const float4 lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
          const uint32_t off = (uint32_t)j * (kTcRows * 16) + row_off;
          *reinterpret_cast<float4*>(sAhi + off) = hi;
          *reinterpret_cast<float4*>(sAlo + off) = lo;
        }
      }
      umma::fence_proxy_async_smem();
      umma::fence_before_sync();
      __syncthreads();

      // ---- contraction on the tensor core
      if (tid == 0) {
        umma::fence_after_sync();
        const uint32_t step_a = 2 * kTcRows * 16, step_b = 2 * (uint32_t)NP * 16;
        const uint32_t a_hi = umma::smem_u32(sAhi), a_lo = umma::smem_u32(sAlo);
        const uint32_t b_hi = umma::smem_u32(sBhi), b_lo = umma::smem_u32(sBlo);
        uint32_t acc = 0;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a = (pass == 2) ? a_lo : a_hi;
          const uint32_t b = (pass == 1) ? b_lo : b_hi;
          for (int kk = 0; kk < (KP >> 3); ++kk) {
            const uint64_t ad = umma::make_desc(a + kk * step_a, kTcRows * 16, 128);
            const uint64_t bd = umma::make_desc(b + kk * step_b, (uint32_t)NP * 16, 128);
            umma::mma_tf32(tmem, ad, bd, idesc, acc);
            acc = 1;
          }
        }
        umma::mma_commit(&bar);
      }
      umma::mbar_wait(&bar, phase);
      phase ^= 1;
      umma::fence_after_sync();

      // ---- epilogue: thread = slot (TMEM lane), warps split the channel range
      {
        const int q = warp & 3, half = warp >> 2;
        const int e = 32 * q + lane;
        const int chh = ((C / 2 + 15) / 16) * 16;
        const int c_begin = half * chh;
        const int c_end = min(C, c_begin + chh);
        const bool live = e < cnt;
        const int d_node = sDst[e];
        const float* Pd = p.PQ + (size_t)d_node * (4 * C);
        const float* Qs = p.PQ + (size_t)sSrc[e] * (4 * C) + 2 * C;
        float gsc = 1.0f;
        const float* grow = nullptr;
        if (MODE != CG_FWD) {
          grow = p.gout + (size_t)d_node * C;
          if (p.inv_deg && live) gsc = __ldg(p.inv_deg + d_node);
        }
        for (int c0 = c_begin; c0 < c_end; c0 += 16) {
          float f[16], s[16];
          umma::tmem_ld16(umma::tmem_addr(tmem, q, c0), f);
          umma::tmem_ld16(umma::tmem_addr(tmem, q, C + c0), s);
          umma::tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
              const int c = c0 + j4;
              if (c < c_end) {
                const float4 pf = __ldg(reinterpret_cast<const float4*>(Pd + c));
                const float4 ps = __ldg(reinterpret_cast<const float4*>(Pd + C + c));
                const float4 qf = __ldg(reinterpret_cast<const float4*>(Qs + c));
                const float4 qs = __ldg(reinterpret_cast<const float4*>(Qs + C + c));
                const float af[4] = {f[j4] + pf.x + qf.x, f[j4 + 1] + pf.y + qf.y,
                                     f[j4 + 2] + pf.z + qf.z, f[j4 + 3] + pf.w + qf.w};
                const float as[4] = {s[j4] + ps.x + qs.x, s[j4 + 1] + ps.y + qs.y,
                                     s[j4 + 2] + ps.z + qs.z, s[j4 + 3] + ps.w + qs.w};
                if (MODE == CG_FWD) {
                  float4 m;
                  m.x = sigmoid_fast_(af[0]) * softplusf_(as[0]);
                  m.y = sigmoid_fast_(af[1]) * softplusf_(as[1]);
                  m.z = sigmoid_fast_(af[2]) * softplusf_(as[2]);
                  m.w = sigmoid_fast_(af[3]) * softplusf_(as[3]);
                  *reinterpret_cast<float4*>(sV + e * VW + c) = m;
                } else {
                  const float4 g = __ldg(reinterpret_cast<const float4*>(grow + c));
                  const float gg[4] = {g.x * gsc, g.y * gsc, g.z * gsc, g.w * gsc};
                  float dfv[4], dsv[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float sg = sigmoid_fast_(af[j]);
                    const float sp = softplusf_(as[j]);
                    dfv[j] = gg[j] * sp * sg * (1.0f - sg);
                    dsv[j] = gg[j] * sg * sigmoid_fast_(as[j]);
                  }
                  *reinterpret_cast<float4*>(sV + e * VW + c) = make_float4(dfv[0], dfv[1], dfv[2], dfv[3]);
                  *reinterpret_cast<float4*>(sV + e * VW + C + c) = make_float4(dsv[0], dsv[1], dsv[2], dsv[3]);
                }
              }
            }
          }
        }
      }
      umma::fence_before_sync();  // accumulator reads done before the next round's MMAs overwrite it
      __syncthreads();

      // ---- segmented sum over the owned segments that have slots in this round
      for (int n = n_lo + warp; n < n_hi; n += kWarps) {
        const int a = __ldg(p.seg_ptr + n), b = __ldg(p.seg_ptr + n + 1);
        const int lo = max(a, r_lo), hi = min(b, r_hi);
        const bool empty_seg = (a == b);
        if (empty_seg ? (rd != 0) : (lo >= hi)) continue;
        const bool first = empty_seg || (a >= r_lo);
        const bool last = empty_seg || (b <= r_hi);
        if (MODE == CG_FWD) {
          float* o = p.out + (size_t)n * C;
          const float* xr = p.x + (size_t)n * C;
          const float sc = p.inv_deg ? __ldg(p.inv_deg + n) : 1.0f;
          for (int c = lane; c < C; c += 32) {
            float acc = first ? 0.0f : o[c];
            for (int s = lo; s < hi; ++s) acc += sV[(s - r_lo) * VW + c];
            o[c] = last ? fmaf(acc, sc, __ldg(xr + c)) : acc;
          }
        } else {
          float* o = p.out + (size_t)n * (4 * C) + (MODE == CG_BWD_SRC ? 2 * C : 0);
          for (int c = lane; c < W2; c += 32) {
            float acc = first ? 0.0f : o[c];
            for (int s = lo; s < hi; ++s) acc += sV[(s - r_lo) * VW + c];
            o[c] = acc;
          }
        }
      }

      // ---- dWe += da^T . ea (register tiles 4 channels x 8 k per work item)
      if (MODE == CG_BWD_DST) {
        const int n_c4 = W2 >> 2;
        const int n_dw = n_c4 * ((GS + 7) >> 3);
#pragma unroll
        for (int j = 0; j < NITEM; ++j) {
          const int it = tid + j * kThreads;
          if (it < n_dw) {
            const int c4 = it % n_c4, k8 = it / n_c4;
            const bool second = (8 * k8 + 4) < GS;
            for (int e = 0; e < cnt; ++e) {
              const float4 da = *reinterpret_cast<const float4*>(sV + e * VW + 4 * c4);
              const float4 e0 = *reinterpret_cast<const float4*>(sEA + e * GS + 8 * k8);
              const float4 e1 = second ? *reinterpret_cast<const float4*>(sEA + e * GS + 8 * k8 + 4)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
              const float dv[4] = {da.x, da.y, da.z, da.w};
              const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
              for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) dw[j][a][b] = fmaf(dv[a], ev[b], dw[j][a][b]);
            }
          }
        }
      }
      __syncthreads();
    }  // rounds
  }    // tiles

  if (MODE == CG_BWD_DST) {
    const int n_c4 = W2 >> 2;
    const int n_dw = n_c4 * ((GS + 7) >> 3);
    float* part = p.dW_part + (size_t)blockIdx.x * G * W2;
#pragma unroll
    for (int j = 0; j < NITEM; ++j) {
      const int it = tid + j * kThreads;
      if (it < n_dw) {
        const int c4 = it % n_c4, k8 = it / n_c4;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int k = 8 * k8 + b;
          if (k < G)
            *reinterpret_cast<float4*>(part + (size_t)k * W2 + 4 * c4) =
                make_float4(dw[j][0][b], dw[j][1][b], dw[j][2][b], dw[j][3][b]);
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, (uint32_t)pl.tmem_cols);
}

template <int MODE, int NITEM>
static int tc_launch_t(const CgParams& p, const TcPlan& pl, int grid, cudaStream_t st) {
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MDL_CUDA(cudaFuncSetAttribute(k_cgconv_tc<MODE, NITEM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kMaxDynSmem));
    configured.store(1, std::memory_order_release);
  }
  k_cgconv_tc<MODE, NITEM><<<grid, kThreads, pl.total, st>>>(p, pl);
  MDL_LAUNCHED();
  return MDL_OK;
}

bool cgtc_supported(int mode, int C, int G) {
  TcPlan pl;
  return tc_plan(mode, C, G, &pl);
}

int cgtc_launch(int mode, CgParams p, cudaStream_t st, int* grid_out) {
  TcPlan pl;
  MDL_REQUIRE(tc_plan(mode, p.C, p.G, &pl), "cgconv_tc: unsupported shape C=%d G=%d", p.C, p.G);
  p.c_off = 0; p.CC = p.C; p.cap = kTcRows; p.te = kTcTE;
  p.n_tiles = (int)std::max<int64_t>(1, ceil_div<int64_t>(p.E, kTcTE));
  const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
  if (grid_out) *grid_out = grid;
  switch (mode) {
    case CG_FWD: return tc_launch_t<CG_FWD, 1>(p, pl, grid, st);
    case CG_BWD_SRC: return tc_launch_t<CG_BWD_SRC, 1>(p, pl, grid, st);
    case CG_BWD_DST:
      return pl.nitem <= 1 ? tc_launch_t<CG_BWD_DST, 1>(p, pl, grid, st)
                           : tc_launch_t<CG_BWD_DST, 2>(p, pl, grid, st);
  }
  MDL_REQUIRE(false, "cgconv_tc: bad mode");
}

}  // namespace mdl
