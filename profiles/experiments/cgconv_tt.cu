// cgconv_tt.cu -- "transposed tile" tensor-core CGConv forward (sm_100a, tcgen05 / TMEM).
//
// Same operator and the same tile-ownership scheme as cgconv_tc.cu, with the contraction
// transposed so that the CHANNEL, not the edge slot, is the TMEM lane:
//
//     D[m, e] = sum_k We[m, k] * ea[e, k]        m = 2*c + gate  (128 lanes),  e = slot (128 columns)
//
//   * A operand = the edge weights We, split hi/lo ONCE per CTA and parked in tensor memory
//     (tcgen05.st by the thread that owns lane m; umma::mma_tf32_ts) -- no shared memory.
//   * B operand = the round's edge_attr rows [slot][k], K-major: exactly the tile the rows land in.
//     cp.async writes the raw rows straight into the operand layout, the hi/lo split happens in place.
//   * epilogue thread = (lane m, half of the slots).  Its loads are coalesced by construction (a warp
//     reads 16 consecutive channels of both gates of one node row: two full 64-byte pieces), so the
//     node projections P[dst], Q[src] need no staging tile; the sum over a destination's slots runs
//     along the thread's own registers, so there is no value tile and no separate reduce pass.
//     The two gates of a channel sit in ADJACENT lanes and meet through one shuffle per slot pair.
//     MUFU budget: exp2 for every lane, log2 on the softplus lanes; the sigmoid lanes take their
//     reciprocal on the FMA pipe (quadratic seed + 2 Newton steps, ~1 ulp), so the transcendental
//     pipe sees 2 warp instructions per slot exactly as in a gate-homogeneous warp.
//   * warp 8 is the producer: it issues the MMAs, then the next round's index loads and cp.async
//     rows as soon as the tensor core has released the operand tiles, and writes the rows of
//     slot-less owned segments.  The 8 epilogue warps never wait for each other inside a round.
//
// Shared memory drops from 219 KB to ~65 KB and TMEM to 256 columns, so TWO CTAs share an SM and
// the hardware overlaps one CTA's load/split/MMA phases with the other's epilogue.
//
// Ownership / ordering of the output rows (deterministic, no atomics): a segment's slots are summed
// in slot order by 64-slot half-rounds; a half that STARTS a segment writes, a half that continues
// one adds to the row after a CTA barrier ("carry"), the half that ends it applies 1/deg and the
// residual x.
#include "cgconv.cuh"
#include "edge_dev.cuh"
#include "umma.cuh"

namespace mdl {

constexpr int kTtEpiThreads = 256;                   // 8 epilogue warps
constexpr int kTtThreads = kTtEpiThreads + 32;       // + the producer warp
constexpr int kTtRows = 128;                         // slots per round = MMA N
constexpr int kTtTE = 112;                           // ownership granularity (as cgconv_tc.cu)
constexpr uint32_t kTtChunk = kTtRows * 16 + 16;     // k-chunk stride of the operand tiles (bank padding)
constexpr int kTtInfoCap = 256;
constexpr int kTtBatch = 16;                         // slots per epilogue batch
constexpr int kTtC = 64;                             // channels (lane map: 64 channels x 2 gates = 128 lanes)

struct TtPlan {
  unsigned long long* prof;
  int KP;
  uint32_t offHi, offLo, offIdx, offInfo, total;
};

static bool tt_plan(int C, int G, TtPlan* pl) {
  if (C != kTtC) return false;
  const int KP = (G + 7) & ~7;
  if (128 + 2 * KP > 256) return false;  // TMEM: D (128 slot columns) + We hi/lo
  const uint32_t tile = (uint32_t)(KP / 4) * kTtChunk;
  pl->prof = nullptr;
  pl->KP = KP;
  pl->offHi = 0;
  pl->offLo = tile;
  pl->offIdx = 2 * tile;                                 // [2 buffers][src|dst][128]
  pl->offInfo = pl->offIdx + 2 * 2 * kTtRows * 4;
  pl->total = pl->offInfo + kTtInfoCap * 16;
  return pl->total <= 112 * 1024;
}

// 1/u for u in [1, 2] on the FMA pipe: minimax quadratic seed (1% error) + two Newton steps
__device__ __forceinline__ float rcp_newton_1_2(float u) {
  float r = fmaf(fmaf(0.3232323232f, u, -1.4545454545f), u, 2.1212121212f);
  r = fmaf(r, fmaf(-u, r, 1.0f), r);
  r = fmaf(r, fmaf(-u, r, 1.0f), r);
  return r;
}

__global__ void __launch_bounds__(kTtThreads, 2) k_cgconv_tt_fwd(const CgParams p, const TtPlan pl) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int C = kTtC, W2 = 2 * kTtC;
  const int G = p.G, KP = pl.KP;
  const bool producer = warp == 8;
  const int q = warp & 3, h = (warp >> 2) & 1;   // TMEM lane quadrant, slot half (epilogue warps)
  const int m = 32 * q + lane;                   // TMEM lane: 2*c + gate
  const int gate = m & 1, c = m >> 1;
  const int col = gate * C + c;                  // column of [f | s] in WeT / P / Q rows

  uint8_t* sHi = smem + pl.offHi;
  uint8_t* sLo = smem + pl.offLo;
  int* sIdx = reinterpret_cast<int*>(smem + pl.offIdx);
  TileInfo* sInfo = reinterpret_cast<TileInfo*>(smem + pl.offInfo);

  const int my_tiles = (p.n_tiles > (int)blockIdx.x) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  int info_base = 0;
  auto fill_infos = [&](int base) {
    for (int k = base + tid; k < min(my_tiles, base + kTtInfoCap); k += kTtThreads) {
      TileInfo t;
      const int tile = blockIdx.x + k * gridDim.x;
      t.n_lo = first_segment_at_or_after<CG_FWD>(p, tile * kTtTE);
      t.n_hi = (tile == p.n_tiles - 1) ? p.N : first_segment_at_or_after<CG_FWD>(p, (tile + 1) * kTtTE);
      if (t.n_hi < t.n_lo) t.n_hi = t.n_lo;
      t.e_lo = __ldg(p.seg_ptr + t.n_lo);
      t.e_hi = __ldg(p.seg_ptr + t.n_hi);
      sInfo[k - base] = t;
    }
  };

  // ---- one-time setup: TMEM, barrier, tile bounds, weights -> tensor memory
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256u);
  if (tid == 32) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  fill_infos(0);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tm_d = tmem, tm_whi = tmem + 128, tm_wlo = tmem + 128 + (uint32_t)KP;
  if (warp < 4) {  // thread = lane m = column `col` of WeT [G, 2C]
    for (int k0 = 0; k0 < KP; k0 += 8) {
      float hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float w = (k0 + j < G) ? __ldg(p.WeT + (size_t)(k0 + j) * W2 + col) : 0.0f;
        hi[j] = umma::tf32_hi(w);
        lo[j] = w - hi[j];
      }
      umma::tmem_st8(umma::tmem_addr(tm_whi, q, k0), hi);
      umma::tmem_st8(umma::tmem_addr(tm_wlo, q, k0), lo);
    }
    umma::tmem_st_wait();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t idesc = umma::make_idesc_tf32(128, kTtRows);
  uint32_t phase = 0;

  // ---- producer: indices of a round (4 rows per lane, registers), then its ea rows via cp.async
  // straight into the operand layout
  struct Idx4 { int s[4], d[4]; };
  auto issue_idx = [&](int r_lo, int cnt) {
    Idx4 v;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = lane + 32 * u;
      v.s[u] = v.d[u] = 0;
      if (e < cnt) {
        v.s[u] = __ldg(p.dst_src + r_lo + e);
        v.d[u] = __ldg(p.dst_dst + r_lo + e);
      }
    }
    return v;
  };
  auto land_idx_and_rows = [&](const Idx4& v, int r_lo, int cnt, int buf) {
    int* bS = sIdx + buf * 2 * kTtRows;
    int* bD = bS + kTtRows;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      bS[lane + 32 * u] = v.s[u];
      bD[lane + 32 * u] = v.d[u];
    }
    // the round's rows are one contiguous span of cnt*G floats (slot order)
    const float* src = p.ea + (size_t)r_lo * G;
    if ((G & 1) == 0) {  // 8-byte pieces (rows are 8-byte aligned when G is even)
      const int cpr = G >> 1, total = cnt * cpr;
      for (int i = lane; i < total; i += 32) {
        const int e = i / cpr, j = i - e * cpr;
        const uint32_t off = (uint32_t)(j >> 1) * kTtChunk + (uint32_t)(e >> 3) * 128 + (uint32_t)(e & 7) * 16 +
                             (uint32_t)(j & 1) * 8;
        cp_async8(sHi + off, src + 2 * i);
      }
    } else {
      const int total = cnt * G;
      for (int i = lane; i < total; i += 32) {
        const int e = i / G, j = i - e * G;
        const uint32_t off = (uint32_t)(j >> 2) * kTtChunk + (uint32_t)(e >> 3) * 128 + (uint32_t)(e & 7) * 16 +
                             (uint32_t)(j & 3) * 4;
        cp_async4(sHi + off, src + i);
      }
    }
  };

  long long t_prev = clock64();
  auto mark = [&](int slot) {
    if (pl.prof && (tid == 0 || tid == kTtEpiThreads)) {
      const long long now = clock64();
      atomicAdd(pl.prof + slot + (producer ? 8 : 0), (unsigned long long)(now - t_prev));
      t_prev = now;
    }
  };

  int k = 0, rd = 0, buf = 0;
  if (producer && my_tiles > 0) {
    const TileInfo t0 = sInfo[0];
    const int c0 = min(t0.e_hi - t0.e_lo, kTtRows);
    const Idx4 v = issue_idx(t0.e_lo, c0);
    land_idx_and_rows(v, t0.e_lo, c0, 0);
  }

  // carry of the (at most one) segment of this thread's previous half-round that had started before
  // its slot range; applied after a CTA barrier (see header).  Lives in the even (sigmoid) lanes.
  bool carry_on = false, carry_final = false;
  int carry_n = 0;
  float carry_v = 0.f, carry_x = 0.f, carry_sc = 1.f;
  auto apply_carry = [&]() {
    if (carry_on) {
      float* o = p.out + (size_t)carry_n * C + c;
      const float tot = *o + carry_v;
      *o = carry_final ? fmaf(tot, carry_sc, carry_x) : tot;
      carry_on = false;
    }
  };

  while (k < my_tiles) {
    if (k + 1 >= info_base + kTtInfoCap && info_base + kTtInfoCap < my_tiles) {
      __syncthreads();
      info_base = k;
      fill_infos(info_base);
      __syncthreads();
    }
    const TileInfo T = sInfo[k - info_base];
    const int rounds = max(1, (T.e_hi - T.e_lo + kTtRows - 1) / kTtRows);
    const int r_lo = T.e_lo + rd * kTtRows;
    const int r_hi = min(T.e_hi, r_lo + kTtRows);
    const int cnt = r_hi - r_lo;
    const int n_lo = T.n_lo, n_hi = T.n_hi;
    const bool same_tile = (rd + 1 < rounds);
    const int nk = same_tile ? k : k + 1, nrd = same_tile ? rd + 1 : 0;
    const int* bSrc = sIdx + buf * 2 * kTtRows;
    const int* bDst = bSrc + kTtRows;

    mark(0);
    if (producer) cp_async_wait_all();
    __syncthreads();  // [S1] rows + indices of this round landed; previous round fully retired
    mark(1);
    if (!producer && h == 1) apply_carry();  // second-half carries: after the first halves' (applied at S3)

    // ---- split hi/lo in place (raw rows sit in the hi tile)
    if (!producer) {
      const int e = tid & (kTtRows - 1);
      const uint32_t row_off = (uint32_t)(e >> 3) * 128 + (uint32_t)(e & 7) * 16;
      for (int j = (tid >> 7); j < (KP >> 2); j += kTtEpiThreads / kTtRows) {
        const uint32_t off = (uint32_t)j * kTtChunk + row_off;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < cnt) {
          v = *reinterpret_cast<const float4*>(sHi + off);
          if (4 * j + 0 >= G) v.x = 0.f;  // padding columns hold stale bytes: exact zeros
          if (4 * j + 1 >= G) v.y = 0.f;
          if (4 * j + 2 >= G) v.z = 0.f;
          if (4 * j + 3 >= G) v.w = 0.f;
        }
        float4 hi;
        hi.x = umma::tf32_hi(v.x); hi.y = umma::tf32_hi(v.y);
        hi.z = umma::tf32_hi(v.z); hi.w = umma::tf32_hi(v.w);
        *reinterpret_cast<float4*>(sHi + off) = hi;
        *reinterpret_cast<float4*>(sLo + off) = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
      }
    }
    umma::fence_proxy_async_smem();
    umma::fence_before_sync();
    __syncthreads();  // [S2] operands staged
    mark(2);

    if (producer) {
      // ================= producer warp =================
      if (lane == 0 && cnt > 0) {
        umma::fence_after_sync();
        const uint32_t step_b = 2 * kTtChunk;
        const uint32_t b_hi = umma::smem_u32(sHi), b_lo = umma::smem_u32(sLo);
        uint32_t acc = 0;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a = (pass == 2) ? tm_wlo : tm_whi;
          const uint32_t b = (pass == 1) ? b_lo : b_hi;
          for (int kk = 0; kk < (KP >> 3); ++kk) {
            const uint64_t bd = umma::make_desc(b + kk * step_b, kTtChunk, 128);
            umma::mma_tf32_ts(tm_d, a + kk * 8, bd, idesc, acc);
            acc = 1;
          }
        }
        umma::mma_commit(&bar);
      }
      __syncwarp();
      mark(3);
      // owned segments without any slot: out = x  (once per tile)
      if (rd == 0) {
        for (int i = lane; i < (n_hi - n_lo) * C; i += 32) {
          const int n = n_lo + i / C, cc = i % C;
          if (__ldg(p.seg_ptr + n) == __ldg(p.seg_ptr + n + 1))
            p.out[(size_t)n * C + cc] = __ldg(p.x + (size_t)n * C + cc);
        }
      }
      int ncnt = 0, nr_lo = 0;
      Idx4 nv{};
      if (nk < my_tiles) {
        const TileInfo Tn = sInfo[nk - info_base];
        nr_lo = Tn.e_lo + nrd * kTtRows;
        ncnt = min(Tn.e_hi - nr_lo, kTtRows);
        nv = issue_idx(nr_lo, ncnt);
      }
      mark(4);
      if (cnt > 0) {
        umma::mbar_wait(&bar, phase);
        umma::fence_after_sync();
      }
      mark(5);
      // operand tiles are free again: the next round's rows land under this round's epilogue
      if (nk < my_tiles) land_idx_and_rows(nv, nr_lo, ncnt, buf ^ 1);
      mark(6);
    } else {
      // ================= epilogue warps: thread = (TMEM lane m, slot half h) =================
      const int my_lo = 64 * h, my_hi = min(cnt, 64 * h + 64);
      float qn[kTtBatch], pn[kTtBatch];
      auto prefetch = [&](int e0) {
#pragma unroll
        for (int j = 0; j < kTtBatch; ++j) {
          qn[j] = 0.f;
          pn[j] = 0.f;
          if (e0 + j < my_hi) {
            qn[j] = __ldg(p.PQ + (size_t)bSrc[e0 + j] * (4 * C) + 2 * C + col);
            pn[j] = __ldg(p.PQ + (size_t)bDst[e0 + j] * (4 * C) + col);
          }
        }
      };
      if (my_lo < my_hi) prefetch(my_lo);
      mark(3);
      if (cnt > 0) {
        umma::mbar_wait(&bar, phase);
        umma::fence_after_sync();
      }
      mark(4);
      if (my_lo < my_hi) {
        int cur = -1, cur_end = 0;   // destination node of the running segment, its end slot
        bool cur_cont = false;       // it started before this half's slot range
        float acc = 0.f, cur_x = 0.f, cur_sc = 1.f;
        auto flush = [&]() {
          const float tot = acc + __shfl_xor_sync(0xffffffffu, acc, 1);
          if (gate == 0) {
            const bool ended = cur_end <= r_lo + my_hi;
            if (cur_cont) {
              carry_on = true; carry_final = ended; carry_n = cur; carry_v = tot; carry_x = cur_x; carry_sc = cur_sc;
            } else {
              p.out[(size_t)cur * C + c] = ended ? fmaf(tot, cur_sc, cur_x) : tot;
            }
          }
        };
        for (int e0 = my_lo; e0 < my_hi; e0 += kTtBatch) {
          float z[kTtBatch];
          umma::tmem_ld16(umma::tmem_addr(tm_d, q, e0), z);
          umma::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < kTtBatch; ++j) z[j] += qn[j] + pn[j];
          if (e0 + kTtBatch < my_hi) prefetch(e0 + kTtBatch);   // next batch's loads fly under this batch's math
          // gate values: sigmoid on even lanes, softplus on odd lanes, one exp2 shared by both forms
#pragma unroll
          for (int j = 0; j < kTtBatch; ++j) {
            const float t = ex2_(-kLog2e * fabsf(z[j]));
            const float u = 1.0f + t;
            if (gate == 0) {
              const float r = rcp_newton_1_2(u);
              z[j] = z[j] >= 0.f ? r : t * r;
            } else {
              z[j] = fmaf(kLn2, lg2_(u), fmaxf(z[j], 0.f));
            }
          }
          // products: the even lane takes slots j < 8 of the batch, the odd lane slots j >= 8
          float prod[kTtBatch / 2];
#pragma unroll
          for (int j = 0; j < kTtBatch / 2; ++j) {
            const float send = gate ? z[j] : z[j + 8];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
            prod[j] = (gate ? z[j + 8] : z[j]) * recv;
          }
#pragma unroll
          for (int j = 0; j < kTtBatch; ++j) {
            if (e0 + j < my_hi) {
              const int d = bDst[e0 + j];
              if (d != cur) {
                if (cur >= 0) flush();
                cur = d;
                acc = 0.f;
                cur_cont = (e0 + j == my_lo) && (__ldg(p.seg_ptr + d) < r_lo + my_lo);
                cur_end = __ldg(p.seg_ptr + d + 1);
                if (gate == 0) {
                  cur_x = __ldg(p.x + (size_t)d * C + c);
                  cur_sc = p.inv_deg ? __ldg(p.inv_deg + d) : 1.0f;
                }
              }
              if ((j >> 3) == gate) acc += prod[j & 7];
            }
          }
        }
        flush();
      }
      mark(5);
    }
    if (cnt > 0) phase ^= 1;
    umma::fence_before_sync();   // accumulator reads retire before the next round's MMAs overwrite D
    __syncthreads();             // [S3] every direct row write of this round is visible
    if (!producer && h == 0) apply_carry();
    if (!producer) mark(6);
    else t_prev = clock64();
    if (pl.prof && tid == 0) atomicAdd(pl.prof + 15, 1ull);
    k = nk; rd = nrd; buf ^= 1;
  }

  if (producer) cp_async_wait_all();
  __syncthreads();
  if (!producer && h == 1) apply_carry();
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256u);
}

static unsigned long long* g_tt_phase_buf = nullptr;
void cgtt_set_phase_buffer(unsigned long long* dev_ptr) { g_tt_phase_buf = dev_ptr; }

bool cgtt_supported(int mode, int C, int G) {
  TtPlan pl;
  return mode == CG_FWD && tt_plan(C, G, &pl);
}

int cgtt_launch(int mode, CgParams p, cudaStream_t st) {
  TtPlan pl;
  MDL_REQUIRE(mode == CG_FWD && tt_plan(p.C, p.G, &pl), "cgconv_tt: unsupported mode/shape C=%d G=%d", p.C, p.G);
  pl.prof = g_tt_phase_buf;
  p.c_off = 0; p.CC = p.C; p.cap = kTtRows; p.te = kTtTE;
  p.n_tiles = (int)std::max<int64_t>(1, ceil_div<int64_t>(p.E, kTtTE));
  const int grid = std::min(p.n_tiles, 2 * kNumSMs);
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MDL_CUDA(cudaFuncSetAttribute(k_cgconv_tt_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    configured.store(1, std::memory_order_release);
  }
  k_cgconv_tt_fwd<<<grid, kTtThreads, pl.total, st>>>(p, pl);
  MDL_LAUNCHED();
  return MDL_OK;
}

}  // namespace mdl
