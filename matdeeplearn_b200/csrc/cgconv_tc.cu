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

constexpr int kTcThreads = 512;  // 16 warps: 4 per scheduler, the epilogue/gather math is latency-bound otherwise
constexpr int kTcWarps = kTcThreads / 32;
constexpr int kTcRows = 128;  // slots per round = MMA M
constexpr int kTcTE = 112;    // ownership granularity (leaves head-room for straddling segments)

struct TcPlan {
  int NP, KP, GS, VW, tmem_cols, nitem;
  uint32_t offBhi, offBlo, offAhi, offAlo, offEA, offV, offIdx, total;
};

static bool tc_plan(int mode, int C, int G, TcPlan* pl) {
  if (C != 64) return false;  // epilogue mapping: 4 lane quadrants x 4 channel quarters of 16
  const int NP = (2 * C + 15) & ~15;
  if (NP > 256) return false;
  const int KP = (G + 7) & ~7;
  int GS = (G + 3) & ~3;
  if (((GS >> 2) & 1) == 0) GS += 4;  // odd number of 16-byte chunks per row: conflict-free float4 column reads
  const int VW = 2 * C + 4;  // [f | s] per slot (+4 floats: conflict-free float4 row access)
  uint32_t b = (uint32_t)NP * KP * 4, a = (uint32_t)kTcRows * KP * 4;
  uint32_t ea = (uint32_t)kTcRows * GS * 4, v = (uint32_t)kTcRows * VW * 4, idx = 4 * kTcRows * 4;
  pl->NP = NP; pl->KP = KP; pl->GS = GS; pl->VW = VW;
  pl->tmem_cols = 32;
  while (pl->tmem_cols < NP) pl->tmem_cols <<= 1;
  const int n_dw = (2 * C / 4) * (KP / 8);
  pl->nitem = (n_dw + kTcThreads / 2 - 1) / (kTcThreads / 2);  // two thread groups split the slots
  if (mode == CG_BWD_DST && pl->nitem > 1) return false;
  pl->offBhi = 0; pl->offBlo = b; pl->offAhi = 2 * b; pl->offAlo = 2 * b + a; pl->offEA = 2 * b + 2 * a;
  // The value tile has its own region: it is filled (gathered node projections) while the MMAs of
  // the same round are still reading the operand tiles, so it cannot alias them.
  uint32_t end = pl->offEA + ea;
  if (end + v + idx <= (uint32_t)kMaxDynSmem) {
    pl->offV = end; pl->offIdx = end + v; pl->total = end + v + idx;
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

// ---- gate math on the MUFU pipe (ex2 / lg2 / rcp approximations, abs. error ~2e-7) ----
__device__ __forceinline__ float ex2_(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float sigmoid_mufu(float x) { return rcp_(1.0f + ex2_(-kLog2e * x)); }
// softplus(x) = max(x,0) + log1p(exp(-|x|))   (== F.softplus incl. its x>20 branch to fp32 rounding)
__device__ __forceinline__ float softplus_mufu(float x) {
  return fmaf(kLn2, lg2_(1.0f + ex2_(-kLog2e * fabsf(x))), fmaxf(x, 0.0f));
}

struct TileInfo { int n_lo, n_hi, e_lo, e_hi; };

// Persistent CTA, software-pipelined over "rounds" of <=128 slots:
//   while round r's MMAs run and its epilogue executes, round r+1's indices and ea rows are
//   already in flight (cp.async) and the node projections for round r are being gathered.
template <int MODE, int NITEM>
__global__ void __launch_bounds__(kTcThreads, 1) k_cgconv_tc(const CgParams p, const TcPlan pl) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ TileInfo sh_tile[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, G = p.G, W2 = 2 * C;
  const int NP = pl.NP, KP = pl.KP, GS = pl.GS, VW = pl.VW;

  uint8_t* sBhi = smem + pl.offBhi;
  uint8_t* sBlo = smem + pl.offBlo;
  uint8_t* sAhi = smem + pl.offAhi;
  uint8_t* sAlo = smem + pl.offAlo;
  float* sEA = reinterpret_cast<float*>(smem + pl.offEA);  // [128][GS] row-major landing zone
  float* sV = reinterpret_cast<float*>(smem + pl.offV);    // [128][VW]
  int* sIdx = reinterpret_cast<int*>(smem + pl.offIdx);    // [2 buffers][src|dst][128]

  // tiles of this CTA: blockIdx.x, +gridDim.x, ...
  const int my_tiles = (p.n_tiles > (int)blockIdx.x) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto compute_info = [&](int k) {  // one thread
    TileInfo t;
    if (k < my_tiles) {
      const int tile = blockIdx.x + k * gridDim.x;
      t.n_lo = first_segment_at_or_after<MODE>(p, tile * kTcTE);
      t.n_hi = (tile == p.n_tiles - 1) ? p.N : first_segment_at_or_after<MODE>(p, (tile + 1) * kTcTE);
      if (t.n_hi < t.n_lo) t.n_hi = t.n_lo;
      t.e_lo = __ldg(p.seg_ptr + t.n_lo);
      t.e_hi = __ldg(p.seg_ptr + t.n_hi);
    } else {
      t.n_lo = t.n_hi = t.e_lo = t.e_hi = 0;
    }
    sh_tile[k % 3] = t;
  };

  // ---- one-time setup
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, (uint32_t)pl.tmem_cols);
  if (tid == 32) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  if (tid == 64) compute_info(0);
  if (tid == 96) compute_info(1);
  for (int i = tid; i < NP * KP; i += kTcThreads) {
    const int n = i % NP, k = i / NP;
    const float w = (k < G && n < W2) ? __ldg(p.WeT + (size_t)k * W2 + n) : 0.0f;
    const float hi = umma::tf32_hi(w);
    const int off = umma::tile_offset_bytes(n, k, NP);
    *reinterpret_cast<float*>(sBhi + off) = hi;
    *reinterpret_cast<float*>(sBlo + off) = w - hi;
  }
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

  // prefetch of one round: warp w owns rows [8w, 8w+8): indices -> smem, ea rows -> cp.async
  auto prefetch = [&](int r_lo, int cnt, int buf) {
    int* bSrc = sIdx + buf * 2 * kTcRows;
    int* bDst = bSrc + kTcRows;
    const int row0 = warp * (kTcRows / kTcWarps);
    int slot = 0;
    if (lane < kTcRows / kTcWarps) {
      const int e = row0 + lane;
      int s = 0, d = 0;
      if (e < cnt) {
        slot = (MODE == CG_BWD_SRC) ? __ldg(p.src_slot + r_lo + e) : (r_lo + e);
        s = __ldg(p.dst_src + slot);
        d = __ldg(p.dst_dst + slot);
      }
      bSrc[e] = s;
      bDst[e] = d;
    }
#pragma unroll
    for (int i = 0; i < kTcRows / kTcWarps; ++i) {
      const int e = row0 + i;
      const int sl = __shfl_sync(0xffffffffu, slot, i);
      if (e < cnt) {
        const float* row = p.ea + (size_t)sl * G;
        float* dst = sEA + e * GS;
        if ((G & 1) == 0) {
          for (int k2 = lane; k2 < (G >> 1); k2 += 32) cp_async8(dst + 2 * k2, row + 2 * k2);
        } else {
          for (int k = lane; k < G; k += 32) cp_async4(dst + k, row + k);
        }
      }
    }
  };

  int k = 0, rd = 0, buf = 0;
  if (my_tiles > 0) {
    const TileInfo t0 = sh_tile[0];
    prefetch(t0.e_lo, min(t0.e_hi - t0.e_lo, kTcRows), 0);
  }

  while (k < my_tiles) {
    const TileInfo T = sh_tile[k % 3];
    const int rounds = max(1, (T.e_hi - T.e_lo + kTcRows - 1) / kTcRows);
    const int r_lo = T.e_lo + rd * kTcRows;
    const int r_hi = min(T.e_hi, r_lo + kTcRows);
    const int cnt = r_hi - r_lo;
    const int n_lo = T.n_lo, n_hi = T.n_hi;
    // next work item
    const bool same_tile = (rd + 1 < rounds);
    const int nk = same_tile ? k : k + 1, nrd = same_tile ? rd + 1 : 0;
    const int* bSrc = sIdx + buf * 2 * kTcRows;
    const int* bDst = bSrc + kTcRows;

    cp_async_wait_all();
    __syncthreads();  // [S1] rows + indices of this round visible; sV / A tiles free

    // ---- split hi/lo into the canonical MMA operand layout
    {
      const int e = tid & (kTcRows - 1);
      const uint32_t row_off = (uint32_t)(e >> 3) * 128 + (uint32_t)(e & 7) * 16;
      for (int j = (tid >> 7); j < (KP >> 2); j += kTcThreads / kTcRows) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < cnt && 4 * j < GS) {
          v = *reinterpret_cast<const float4*>(sEA + e * GS + 4 * j);
          if (4 * j + 0 >= G) v.x = 0.f;  // padding columns: exact zeros
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
    __syncthreads();  // [S2] operands staged; sEA free for the next round's rows

    // ---- contraction on the tensor core (asynchronous)
    if (tid == 0 && cnt > 0) {
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

    // ---- overlap window: look two tiles ahead, put the next round's loads in flight
    if (rd == 0 && tid == 64) compute_info(k + 2);
    if (nk < my_tiles) {
      const TileInfo Tn = sh_tile[nk % 3];
      const int nr_lo = Tn.e_lo + nrd * kTcRows;
      prefetch(nr_lo, min(Tn.e_hi - nr_lo, kTcRows), buf ^ 1);
    }

    // ---- gather P[dst] + Q[src] for this round, coalesced, into the value tile.
    // Warp (q, half) owns slots 32q..32q+31 and channels [half*chh, half*chh+chh) for BOTH the
    // gather and the epilogue below, so only a __syncwarp separates the two.
    // One LDG.128 per lane: lanes [0,chh/2) = 16-byte chunks of P (f half | s half), lanes
    // [chh/2, chh) = the same chunks of Q; with chh = 32 a warp instruction covers one slot
    // (4 full 128-byte lines), with chh = 16 two slots.
    const int q = warp & 3, part = warp >> 2;        // lane quadrant, channel quarter
    constexpr int chh = 16;                         // channels per warp
    const int c_begin = part * chh;
    const bool has_ch = c_begin < C;
    if (cnt > 0 && has_ch) {
      constexpr int rows_per_inst = 32 / chh;         // 2
      const int sub = lane / chh;                    // slot within the instruction
      const int l = lane % chh;                      // lane within the slot's group
      const int is_q = l / (chh / 2);                // 0: P (destination side), 1: Q (source side)
      const int l2 = l % (chh / 2);
      const int is_s = l2 / (chh / 4);               // 0: f channels, 1: s channels
      const int c = c_begin + 4 * (l2 % (chh / 4));  // first of this lane's 4 channels
      const int col = is_q * 2 * C + is_s * C + c;   // column inside a PQ row
      float4 v[32 / rows_per_inst];
#pragma unroll
      for (int i = 0; i < 32; i += rows_per_inst) {  // all loads first: one latency exposure
        const int e = 32 * q + i + sub;
        v[i / rows_per_inst] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < cnt) {
          const int node = is_q ? bSrc[e] : bDst[e];
          v[i / rows_per_inst] = __ldg(reinterpret_cast<const float4*>(p.PQ + (size_t)node * (4 * C) + col));
        }
      }
#pragma unroll
      for (int i = 0; i < 32; i += rows_per_inst) {
        const int e = 32 * q + i + sub;
        float4 t = v[i / rows_per_inst];
        // P + Q: partner lane holds the other operand of the same (slot, channel chunk)
        t.x += __shfl_xor_sync(0xffffffffu, t.x, chh / 2);
        t.y += __shfl_xor_sync(0xffffffffu, t.y, chh / 2);
        t.z += __shfl_xor_sync(0xffffffffu, t.z, chh / 2);
        t.w += __shfl_xor_sync(0xffffffffu, t.w, chh / 2);
        if (!is_q && e < cnt) *reinterpret_cast<float4*>(sV + e * VW + is_s * C + c) = t;
      }
      __syncwarp();
    }

    // ---- epilogue: thread = slot (TMEM lane); a = accumulator + gathered projections
    if (cnt > 0) {
      const int e = 32 * q + lane;
      const bool live = e < cnt;
      umma::mbar_wait(&bar, phase);
      umma::fence_after_sync();
      phase ^= 1;
      if (has_ch) {
#pragma unroll
        for (int c0 = c_begin; c0 < c_begin + chh; c0 += 16) {  // one pass: chh == 16
          float f[16], sacc[16];
          umma::tmem_ld16(umma::tmem_addr(tmem, q, c0), f);
          umma::tmem_ld16(umma::tmem_addr(tmem, q, C + c0), sacc);
          umma::tmem_ld_wait();
          if (live) {
            float* rowv = sV + e * VW;
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
              const int c = c0 + j4;
              const float4 bf = *reinterpret_cast<const float4*>(rowv + c);
              const float4 bs = *reinterpret_cast<const float4*>(rowv + C + c);
              const float af[4] = {f[j4] + bf.x, f[j4 + 1] + bf.y, f[j4 + 2] + bf.z, f[j4 + 3] + bf.w};
              const float as[4] = {sacc[j4] + bs.x, sacc[j4 + 1] + bs.y, sacc[j4 + 2] + bs.z, sacc[j4 + 3] + bs.w};
              float r0[4], r1[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float sg = sigmoid_mufu(af[j]);
                const float sp = softplus_mufu(as[j]);
                if (MODE == CG_FWD) {
                  r0[j] = sg * sp;
                } else {  // d(message)/d(a_f), d(message)/d(a_s); the grad_out factor is applied in the reduce stage
                  r0[j] = sp * sg * (1.0f - sg);
                  r1[j] = sg * sigmoid_mufu(as[j]);
                }
              }
              *reinterpret_cast<float4*>(rowv + c) = make_float4(r0[0], r0[1], r0[2], r0[3]);
              if (MODE != CG_FWD)
                *reinterpret_cast<float4*>(rowv + C + c) = make_float4(r1[0], r1[1], r1[2], r1[3]);
            }
          }
        }
      }
    }
    umma::fence_before_sync();  // accumulator reads done before the next round's MMAs overwrite it
    __syncthreads();            // [S3] value tile complete

    // ---- segmented sum over the owned segments that have slots in this round
    for (int n = n_lo + warp; n < n_hi; n += kTcWarps) {
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
          const float xv = last ? __ldg(xr + c) : 0.0f;  // issued before the sum: latency overlaps it
          float acc = first ? 0.0f : o[c];
          for (int s = lo; s < hi; ++s) acc += sV[(s - r_lo) * VW + c];
          o[c] = last ? fmaf(acc, sc, xv) : acc;
        }
      } else if (MODE == CG_BWD_DST) {
        // segment = destination: one grad_out row scales every slot of the segment.  The scaled
        // per-slot values are written back: the dWe pass below needs da = g * bracket per slot.
        float* o = p.out + (size_t)n * (4 * C);
        const float sc = p.inv_deg ? __ldg(p.inv_deg + n) : 1.0f;
        for (int c = lane; c < W2; c += 32) {
          const float g = __ldg(p.gout + (size_t)n * C + (c < C ? c : c - C)) * sc;
          float acc = first ? 0.0f : o[c];
          for (int s = lo; s < hi; ++s) {
            const float v = sV[(s - r_lo) * VW + c] * g;
            sV[(s - r_lo) * VW + c] = v;
            acc += v;
          }
          o[c] = acc;
        }
      } else {
        // segment = source: every slot has its own destination, hence its own grad_out row
        float* o = p.out + (size_t)n * (4 * C) + 2 * C;
        for (int c = lane; c < W2; c += 32) {
          const int cc = (c < C ? c : c - C);
          float acc = first ? 0.0f : o[c];
#pragma unroll 4
          for (int s = lo; s < hi; ++s) {
            const int d = bDst[s - r_lo];
            const float g = __ldg(p.gout + (size_t)d * C + cc) * (p.inv_deg ? __ldg(p.inv_deg + d) : 1.0f);
            acc = fmaf(sV[(s - r_lo) * VW + c], g, acc);
          }
          o[c] = acc;
        }
      }
    }
    if (MODE == CG_BWD_DST) __syncthreads();  // scaled value tile visible to the dWe pass

    // ---- dWe += da^T . ea   (ea re-assembled exactly as hi + lo from the operand tiles).
    // Two thread groups take alternate slots; each keeps its own register tile and partial.
    if (MODE == CG_BWD_DST) {
      const int n_c4 = W2 >> 2;
      const int n_dw = n_c4 * (KP >> 3);
      const int grp = tid / (kTcThreads / 2), it = tid % (kTcThreads / 2);
      if (it < n_dw) {
        const int c4 = it % n_c4, k8 = it / n_c4;
        const uint8_t* h0 = sAhi + (uint32_t)(2 * k8) * (kTcRows * 16);
        const uint8_t* l0 = sAlo + (uint32_t)(2 * k8) * (kTcRows * 16);
#pragma unroll 2
        for (int e = grp; e < cnt; e += 2) {
          const uint32_t ro = (uint32_t)(e >> 3) * 128 + (uint32_t)(e & 7) * 16;
          const float4 da = *reinterpret_cast<const float4*>(sV + e * VW + 4 * c4);
          const float4 ah = *reinterpret_cast<const float4*>(h0 + ro);
          const float4 al = *reinterpret_cast<const float4*>(l0 + ro);
          const float4 bh = *reinterpret_cast<const float4*>(h0 + kTcRows * 16 + ro);
          const float4 bl = *reinterpret_cast<const float4*>(l0 + kTcRows * 16 + ro);
          const float dv[4] = {da.x, da.y, da.z, da.w};
          const float ev[8] = {ah.x + al.x, ah.y + al.y, ah.z + al.z, ah.w + al.w,
                               bh.x + bl.x, bh.y + bl.y, bh.z + bl.z, bh.w + bl.w};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) dw[0][a][b] = fmaf(dv[a], ev[b], dw[0][a][b]);
        }
      }
    }
    k = nk; rd = nrd; buf ^= 1;
  }  // work items

  if (MODE == CG_BWD_DST) {
    const int n_c4 = W2 >> 2;
    const int n_dw = n_c4 * (KP >> 3);
    const int grp = tid / (kTcThreads / 2), it = tid % (kTcThreads / 2);
    float* part = p.dW_part + ((size_t)blockIdx.x * 2 + grp) * G * W2;
    if (it < n_dw) {
      const int c4 = it % n_c4, k8 = it / n_c4;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int kcol = 8 * k8 + b;
        if (kcol < G)
          *reinterpret_cast<float4*>(part + (size_t)kcol * W2 + 4 * c4) =
              make_float4(dw[0][0][b], dw[0][1][b], dw[0][2][b], dw[0][3][b]);
      }
    }
  }
  cp_async_wait_all();
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
  k_cgconv_tc<MODE, NITEM><<<grid, kTcThreads, pl.total, st>>>(p, pl);
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
    case CG_BWD_DST: return tc_launch_t<CG_BWD_DST, 1>(p, pl, grid, st);
  }
  MDL_REQUIRE(false, "cgconv_tc: bad mode");
}

}  // namespace mdl
