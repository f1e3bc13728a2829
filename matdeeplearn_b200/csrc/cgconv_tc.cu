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
#include "edge_dev.cuh"

namespace mdl {

constexpr int kTcThreads = 512;  // 16 warps: 4 per scheduler, the epilogue/gather math is latency-bound otherwise
constexpr int kTcWarps = kTcThreads / 32;
constexpr int kTcRows = 128;  // slots per round = MMA M
// k-chunk stride of the A operand tiles: 128 rows x 16 B plus 16 B of padding, so that lanes that
// read the same row of DIFFERENT chunks (the dWe B-fragments) fall on different banks.  The UMMA
// descriptor simply carries this value as its leading-dimension byte offset.
constexpr uint32_t kAChunk = kTcRows * 16 + 16;
constexpr int kTcTE = 112;
constexpr int kInfoCap = 512;  // tile-bound records kept in smem per CTA (refilled if a CTA owns more tiles)    // ownership granularity (leaves head-room for straddling segments)

// optional per-phase cycle accounting (development aid): 16 counters, thread 0 of each CTA adds
// the cycles it spent between consecutive phase marks.  Enabled by mdl_debug_set_phase_buffer.
static unsigned long long* g_phase_buf = nullptr;

struct TcPlan {
  unsigned long long* prof;
  int dq_atomic;  // BWD_DST also scatters da into dQ[src] with vector float atomics (no BWD_SRC pass)
  int window;     // stage the round's contiguous P / Q node-row ranges in shared memory when they fit
  int ea_bulk;    // a round's edge rows are one contiguous block of ea: fetch it with ONE bulk (TMA) copy
  int NP, KP, GS, VW, tmem_cols, nitem;
  uint32_t offBhi, offBlo, offAhi, offAlo, offEA, offV, offIdx, offRange, offInfo, total;
};

static bool tc_plan(int mode, int C, int G, TcPlan* pl) {
  if (C != 64) return false;  // epilogue mapping: 4 lane quadrants x 4 channel quarters of 16
  const int NP = (2 * C + 15) & ~15;
  if (NP > 256) return false;
  const int KP = (G + 7) & ~7;
  int GS = (G + 3) & ~3;
  if (((GS >> 2) & 1) == 0) GS += 4;  // odd number of 16-byte chunks per row: conflict-free float4 column reads
  const int VW = 2 * C + 4;  // [f | s] per slot (+4 floats: conflict-free float4 row access)
  uint32_t b = (uint32_t)NP * KP * 4, a = (uint32_t)(KP / 4) * kAChunk;
  uint32_t ea = (uint32_t)kTcRows * GS * 4 + 16, v = (uint32_t)kTcRows * VW * 4;  // +16: bulk-copy alignment slack
  const uint32_t idx = 4 * kTcRows * 4 + 512;  // indices [2][src|dst][128] + per-warp node ranges [2][16][4]
  pl->prof = g_phase_buf;
  pl->dq_atomic = 0;
  pl->window = 1;
  pl->ea_bulk = 0;
  pl->NP = NP; pl->KP = KP; pl->GS = GS; pl->VW = VW;
  pl->tmem_cols = 32;
  while (pl->tmem_cols < NP) pl->tmem_cols <<= 1;
  const int n_dw = (2 * C / 4) * (KP / 8);
  pl->nitem = 1;
  (void)n_dw;
  if (mode == CG_BWD_DST && KP > 64) return false;  // dWe warp tiling: 4 column groups of 16
  pl->offBhi = 0; pl->offBlo = b; pl->offAhi = 2 * b; pl->offAlo = 2 * b + a; pl->offEA = 2 * b + 2 * a;
  // The value tile has its own region: it is filled (gathered node projections) while the MMAs of
  // the same round are still reading the operand tiles, so it cannot alias them.
  uint32_t end = pl->offEA + ea;
  const uint32_t info = kInfoCap * 16;  // per-CTA table of tile bounds (TileInfo)
  if (end + v + idx + info <= (uint32_t)kMaxDynSmem) {
    pl->offV = end; pl->offIdx = end + v; pl->offRange = end + v + idx - 512; pl->offInfo = end + v + idx;
    pl->total = end + v + idx + info;
    return true;
  }
  return false;
}


// 16-byte vector float reduction into global memory (sm_90+): no return value, resolved in L2
__device__ __forceinline__ void red_add_v4(float* addr, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// D[16x8] += A[16x8] * B[8x8]  (tf32 inputs, fp32 accumulate; legacy warp-level tensor path)
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const float (&a)[4], float b0, float b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
        "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// Persistent CTA, software-pipelined over "rounds" of <=128 slots:
//   while round r's MMAs run and its epilogue executes, round r+1's indices and ea rows are
//   already in flight (cp.async) and the node projections for round r are being gathered.
template <int MODE, int PROFILE, int GATE, int EA_BULK>
__global__ void __launch_bounds__(kTcThreads, 1) k_cgconv_tc(const CgParams p, const TcPlan pl) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;       // MMA completion (tcgen05.commit)
  __shared__ uint64_t bar_rows;  // this round's node rows (bulk copies into the value tile)
  __shared__ uint64_t bar_ea;    // next round's edge rows (one bulk copy into the landing zone)
  __shared__ uint32_t tmem_base_s;
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
  TileInfo* sInfo = reinterpret_cast<TileInfo*>(smem + pl.offInfo);  // bounds of this CTA's tiles
  // node ranges touched by a round, [2 buffers][16 warps][src min, src max, dst min, dst max] (one
  // record per warp for its 8 slots, combined by whoever needs the range): edges of a crystal graph
  // stay inside the graph, so the P / Q rows a round gathers lie in two short CONTIGUOUS node ranges.
  // When both fit, they are streamed into the (still idle) value tile by bulk copies and the epilogue
  // reads them from shared memory: a few KB from L2 per round instead of one 512-byte row per slot
  // and operand, and no register staging.
  int* sRange = reinterpret_cast<int*>(smem + pl.offRange);

  // tiles of this CTA: blockIdx.x, +gridDim.x, ...
  const int my_tiles = (p.n_tiles > (int)blockIdx.x) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // Bounds of tile k of this CTA (two short dependent-load chains).  ALL tiles of the CTA are
  // resolved up front, one per thread, so the chains run in parallel once instead of serially
  // on one thread inside every round.
  int info_base = 0;
  auto fill_infos = [&](int base) {
    for (int k = base + tid; k < min(my_tiles, base + kInfoCap); k += kTcThreads) {
      TileInfo t;
      const int tile = blockIdx.x + k * gridDim.x;
      t.n_lo = first_segment_at_or_after<MODE>(p, tile * kTcTE);
      t.n_hi = (tile == p.n_tiles - 1) ? p.N : first_segment_at_or_after<MODE>(p, (tile + 1) * kTcTE);
      if (t.n_hi < t.n_lo) t.n_hi = t.n_lo;
      t.e_lo = __ldg(p.seg_ptr + t.n_lo);
      t.e_hi = __ldg(p.seg_ptr + t.n_hi);
      sInfo[k - base] = t;
    }
  };

  // ---- one-time setup
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, (uint32_t)pl.tmem_cols);
  if (tid == 32) {
    umma::mbar_init(&bar, 1);
    umma::mbar_init(&bar_rows, 1);
    umma::mbar_init(&bar_ea, 1);
    umma::fence_mbar_init();
  }

  fill_infos(0);
  for (int i = tid; i < NP * KP; i += kTcThreads) {
    const int n = i % NP, k = i / NP;
    const float w = (k < G && n < W2) ? __ldg(p.WeT + (size_t)k * W2 + n) : 0.0f;
    const float hi = umma::tf32_hi(w);
    const int off = umma::tile_offset_bytes(n, k, NP);
    *reinterpret_cast<float*>(sBhi + off) = hi;
    *reinterpret_cast<float*>(sBlo + off) = w - hi;
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma::make_idesc_tf32(kTcRows, NP);
  uint32_t phase = 0;

  // ---- per-round loads are split into an "issue" half (registers) and a "land" half so that
  // every global load of the overlap window is in flight before the first one is consumed.
  constexpr int kRowsPerWarp = kTcRows / kTcWarps;  // 8
  const int row0 = warp * kRowsPerWarp;
  struct NextIdx { int slot, s, d; };
  auto issue_idx = [&](int r_lo, int cnt) {  // lanes < 8: indices of this warp's rows of a round
    NextIdx ni{0, 0, 0};
    if (lane < kRowsPerWarp) {
      const int e = row0 + lane;
      if (e < cnt) {
        ni.slot = (MODE == CG_BWD_SRC) ? __ldg(p.src_slot + r_lo + e) : (r_lo + e);
        ni.s = __ldg(p.dst_src + ni.slot);
        ni.d = __ldg(p.dst_dst + ni.slot);
      }
    }
    return ni;
  };
  auto land_idx_and_rows = [&](const NextIdx& ni, int cnt, int buf) {
    int* bS = sIdx + buf * 2 * kTcRows;
    int* bD = bS + kTcRows;
    if (lane < kRowsPerWarp) {
      bS[row0 + lane] = ni.s;
      bD[row0 + lane] = ni.d;
    }
    if (pl.window) {  // this warp's record (neutral when it holds no real slot)
      const bool real = lane < kRowsPerWarp && row0 + lane < cnt;
      const int s_lo = __reduce_min_sync(0xffffffffu, real ? ni.s : 0x7fffffff);
      const int s_hi = __reduce_max_sync(0xffffffffu, real ? ni.s : -0x40000000);
      const int d_lo = __reduce_min_sync(0xffffffffu, real ? ni.d : 0x7fffffff);
      const int d_hi = __reduce_max_sync(0xffffffffu, real ? ni.d : -0x40000000);
      if (lane == 0) *reinterpret_cast<int4*>(sRange + (buf * kTcWarps + warp) * 4) = make_int4(s_lo, s_hi, d_lo, d_hi);
    }
    if (EA_BULK) return;  // the edge rows arrive by one bulk copy (issue_ea_bulk)
    // the warp's 8 rows, lanes across the 8-byte (even G) or 4-byte chunks of a row
    const int wcnt = min(kRowsPerWarp, cnt - row0);  // rows of this warp that exist (may be <= 0)
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const int sl = __shfl_sync(0xffffffffu, ni.slot, r);
      if (r < wcnt) {
        float* d = sEA + (row0 + r) * GS;
        const float* g = p.ea + (size_t)sl * G;
        if ((G & 1) == 0) {
#pragma unroll 1
          for (int c = 2 * lane; c < G; c += 64) cp_async8(d + c, g + c);
        } else {
#pragma unroll 1
          for (int c = lane; c < G; c += 32) cp_async4(d + c, g + c);
        }
      }
    }
  };

  // Bulk variant of the edge-row fetch (FWD / BWD_DST: slots r_lo.. are consecutive rows of ea, i.e.
  // one contiguous block).  The block starts at any 4-byte offset; it is copied from the 16-byte
  // boundary below it, so element (e, k) lands at sEA[ea_off + e*G + k], ea_off = (r_lo*G) & 3.  The
  // copy length is rounded up to 16 bytes when ea continues for >= 12 bytes behind the block,
  // otherwise down, and thread 0 moves the last few floats by hand.
  auto ea_bulk_bytes = [&](int r_lo, int cnt) -> uint32_t {
    if (cnt <= 0) return 0u;
    const long long first = (long long)r_lo * G;
    const uint32_t bytes = (uint32_t)(((int)(first & 3) + cnt * G) * 4);
    const bool more = ((long long)p.E * G - (first + (long long)cnt * G)) >= 3;
    return more ? ((bytes + 15u) & ~15u) : (bytes & ~15u);
  };
  auto issue_ea_bulk = [&](int r_lo, int cnt) {  // thread 0 only
    const long long first = (long long)r_lo * G;
    const int off = (int)(first & 3);
    const float* src = p.ea + (first - off);
    const uint32_t bytes = (uint32_t)((off + cnt * G) * 4), nb = ea_bulk_bytes(r_lo, cnt);
    if (nb) {
      umma::mbar_arrive_expect_tx(&bar_ea, nb);
      umma::bulk_g2s(sEA, src, nb, &bar_ea);
    }
    for (uint32_t t = nb / 4; t < bytes / 4; ++t) sEA[t] = __ldg(src + t);  // tail at the very end of ea
  };
  uint32_t ph_rows = 0, ph_ea = 0;
  int dbg_round = 0;
  // debugging aid (instrumented build only): a wait that does not complete within 2^20 polls records
  // (barrier id, CTA, round, thread) in the phase buffer and carries on instead of trapping
  auto wait_dbg = [&](uint64_t* b, uint32_t parity, int id) {
    if (!PROFILE || !pl.prof) { umma::mbar_wait(b, parity); return; }
    const uint32_t addr = umma::smem_u32(b);
    for (uint32_t spin = 0; spin < (1u << 20); ++spin) {
      uint32_t done;
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(addr), "r"(parity) : "memory");
      if (done) return;
    }
    if (atomicCAS(pl.prof + 25, 0ull, 1ull) == 0ull) {
      pl.prof[26] = (unsigned long long)id;
      pl.prof[27] = blockIdx.x;
      pl.prof[28] = (unsigned long long)dbg_round;
      pl.prof[29] = (unsigned long long)tid;
      pl.prof[30] = parity;
    }
  };

  long long t_prev = PROFILE ? clock64() : 0;
  auto mark = [&](int slot) {
    if (PROFILE && pl.prof && tid == 0) {
      const long long now = clock64();
      atomicAdd(pl.prof + slot, (unsigned long long)(now - t_prev));
      t_prev = now;
    }
  };

  // dWe accumulators of the mma.sync path (BWD_DST): 2 x 2 tiles of 16 channels x 8 k-columns
  float dacc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) dacc[a][b] = 0.0f;

  int k = 0, rd = 0, buf = 0;
  if (my_tiles > 0) {
    const TileInfo t0 = sInfo[0];
    const int c0 = min(t0.e_hi - t0.e_lo, kTcRows);
    const NextIdx ni = issue_idx(t0.e_lo, c0);
    land_idx_and_rows(ni, c0, 0);
    if (EA_BULK && tid == 0) issue_ea_bulk(t0.e_lo, c0);
  }

  const int q = warp & 3, part = warp >> 2;  // TMEM lane quadrant, channel quarter (16 channels)
  constexpr int chh = 16;
  const int c_begin = part * chh;

  while (k < my_tiles) {
    if (k + 1 >= info_base + kInfoCap && info_base + kInfoCap < my_tiles) {  // table exhausted: refill
      __syncthreads();
      info_base = k;
      fill_infos(info_base);
      __syncthreads();
    }
    const TileInfo T = sInfo[k - info_base];
    const int rounds = max(1, (T.e_hi - T.e_lo + kTcRows - 1) / kTcRows);
    const int r_lo = T.e_lo + rd * kTcRows;
    const int r_hi = min(T.e_hi, r_lo + kTcRows);
    const int cnt = r_hi - r_lo;
    const int n_lo = T.n_lo, n_hi = T.n_hi;
    const bool same_tile = (rd + 1 < rounds);
    const int nk = same_tile ? k : k + 1, nrd = same_tile ? rd + 1 : 0;
    const int* bSrc = sIdx + buf * 2 * kTcRows;
    const int* bDst = bSrc + kTcRows;

    mark(0);
    if (!EA_BULK) cp_async_wait_all();
    umma::fence_proxy_async_smem();  // generic accesses of the value tile (last round) before bulk writes into it
    __syncthreads();  // [S1] indices of this round visible; value / operand tiles free
    mark(1);

    // ---- indices of the next round: issued first, landed after the MMA issue (in BWD_SRC they are a
    // dependent chain); everything up to there overlaps their latency
    int ncnt = 0;
    NextIdx ni{0, 0, 0};
    if (nk < my_tiles) {
      const TileInfo Tn = sInfo[nk - info_base];
      const int nr_lo = Tn.e_lo + nrd * kTcRows;
      ncnt = min(Tn.e_hi - nr_lo, kTcRows);
      ni = issue_idx(nr_lo, ncnt);
    }
    mark(16);

    // ---- node terms of this round, staged into the (idle) value tile by cp.async (CTA-uniform choice):
    //   window   rows [0,nq) = Q[smin..smax], rows [nq,nq+np) = P[dmin..dmax]
    //   per slot row e = the operand of the UNSORTED side (Q[src[e]]; P[dst[e]] in BWD_SRC); the sorted
    //            side repeats a few rows per warp and is read straight from global in the epilogue
    bool win = false;
    int w_smin = 0, w_dmin = 0, w_nq = 0;
    if (cnt > 0) {
      int4 rg = make_int4(0, 0, 0, 0);
      int nq = kTcRows, np_ = kTcRows;
      if (pl.window) {  // ranges are only maintained in window mode
        int4 w = make_int4(0x7fffffff, -0x40000000, 0x7fffffff, -0x40000000);
        if (lane < kTcWarps) w = *reinterpret_cast<const int4*>(sRange + (buf * kTcWarps + lane) * 4);
        rg.x = __reduce_min_sync(0xffffffffu, w.x); rg.y = __reduce_max_sync(0xffffffffu, w.y);
        rg.z = __reduce_min_sync(0xffffffffu, w.z); rg.w = __reduce_max_sync(0xffffffffu, w.w);
        nq = rg.y - rg.x + 1; np_ = rg.w - rg.z + 1;
      }
      win = nq + np_ <= kTcRows;
      w_smin = rg.x; w_dmin = rg.z; w_nq = nq;
      const int nrows = win ? nq + np_ : cnt;  // one 512-byte bulk copy per row, one thread each
      // The expectation is registered here, next to the copies (copies of other warps may complete first:
      // the transaction count of the phase simply goes negative until this arrives).  Registering it
      // ahead of [S1] instead failed intermittently on the GPU (profiles/r1_bulk_copy_race.txt).
      if (tid == 0) umma::mbar_arrive_expect_tx(&bar_rows, (uint32_t)nrows * (uint32_t)(2 * C * 4));
      // one bulk copy per 512-byte row; a bulk copy is a warp-serial instruction (~70 cycles each when
      // one warp issues a batch), so the rows are dealt out to lane 0 of all 16 warps
      if (lane == 0) {
        for (int r = warp; r < nrows; r += kTcWarps) {
          const float* g;
          if (win) {
            g = (r < nq) ? p.PQ + (size_t)(rg.x + r) * (4 * C) + 2 * C : p.PQ + (size_t)(rg.z + r - nq) * (4 * C);
          } else {
            g = (MODE == CG_BWD_SRC) ? p.PQ + (size_t)bDst[r] * (4 * C) : p.PQ + (size_t)bSrc[r] * (4 * C) + 2 * C;
          }
          umma::bulk_g2s(sV + r * VW, g, (uint32_t)(2 * C * 4), &bar_rows);
        }
      }
    }
    mark(17);
    int ea_off = 0;
    if (EA_BULK) {
      ea_off = (int)(((long long)r_lo * G) & 3);
      if (ea_bulk_bytes(r_lo, cnt)) {  // CTA-uniform
        wait_dbg(&bar_ea, ph_ea, 1);
        ph_ea ^= 1;
      }
    }
    mark(24);
    // ---- split hi/lo into the canonical MMA operand layout
    {
      const int e = tid & (kTcRows - 1);
      const uint32_t row_off = (uint32_t)(e >> 3) * 128 + (uint32_t)(e & 7) * 16;
      for (int j = (tid >> 7); j < (KP >> 2); j += kTcThreads / kTcRows) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EA_BULK) {
          if (e < cnt && 4 * j < G) {  // dense rows (stride G) at a 4- or 8-byte offset: 8-byte loads for even G
            const float* r = sEA + ea_off + e * G + 4 * j;
            if ((G & 1) == 0) {
              const float2 a = *reinterpret_cast<const float2*>(r);
              v.x = a.x; v.y = a.y;
              if (4 * j + 2 < G) { const float2 b2 = *reinterpret_cast<const float2*>(r + 2); v.z = b2.x; v.w = b2.y; }
            } else {
              v.x = r[0];
              if (4 * j + 1 < G) v.y = r[1];
              if (4 * j + 2 < G) v.z = r[2];
              if (4 * j + 3 < G) v.w = r[3];
            }
          }
        } else if (e < cnt && 4 * j < GS) {
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
        const uint32_t off = (uint32_t)j * kAChunk + row_off;
        *reinterpret_cast<float4*>(sAhi + off) = hi;
        *reinterpret_cast<float4*>(sAlo + off) = lo;
      }
    }
    mark(18);
    umma::fence_proxy_async_smem();
    umma::fence_before_sync();
    mark(19);
    __syncthreads();  // [S2] operands staged; the ea landing zone is free for the next round
    mark(2);
    if (EA_BULK && tid == 0 && nk < my_tiles) {
      const TileInfo Tn = sInfo[nk - info_base];
      const int nr_lo = Tn.e_lo + nrd * kTcRows;
      issue_ea_bulk(nr_lo, min(Tn.e_hi - nr_lo, kTcRows));
    }

    // ---- contraction on the tensor core (asynchronous; the issuing lane belongs to the LAST warp,
    // whose other duties in this window are the lightest)
    if (tid == kTcThreads - 32 && cnt > 0) {
      umma::fence_after_sync();
      const uint32_t step_a = 2 * kAChunk, step_b = 2 * (uint32_t)NP * 16;
      const uint32_t a_hi = umma::smem_u32(sAhi), a_lo = umma::smem_u32(sAlo);
      const uint32_t b_hi = umma::smem_u32(sBhi), b_lo = umma::smem_u32(sBlo);
      uint32_t acc = 0;
#pragma unroll 1
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a = (pass == 2) ? a_lo : a_hi;
        const uint32_t b = (pass == 1) ? b_lo : b_hi;
        for (int kk = 0; kk < (KP >> 3); ++kk) {
          const uint64_t ad = umma::make_desc(a + kk * step_a, kAChunk, 128);
          const uint64_t bd = umma::make_desc(b + kk * step_b, (uint32_t)NP * 16, 128);
          umma::mma_tf32(tmem, ad, bd, idesc, acc);
          acc = 1;
        }
      }
      umma::mma_commit(&bar);
    }
    mark(3);

    // ---- overlap window, part 1: ISSUE every global load (nothing below waits on memory yet)
    // (b) what the reduce stage will need for this warp's first segment
    const int n0 = n_lo + warp;
    int seg_a = 0, seg_b = 0;
    float seg_sc = 1.0f, seg_x0 = 0.0f, seg_x1 = 0.0f;
    if (n0 < n_hi) {
      seg_a = __ldg(p.seg_ptr + n0);
      seg_b = __ldg(p.seg_ptr + n0 + 1);
      if (MODE == CG_FWD) {
        if (p.inv_deg) seg_sc = __ldg(p.inv_deg + n0);
        seg_x0 = __ldg(p.x + (size_t)n0 * C + lane);
        seg_x1 = __ldg(p.x + (size_t)n0 * C + 32 + lane);
      }
    }
    // (c) backward: grad_out rows of this warp's 8 slots, lanes = channels.  For mean aggregation
    // the rows arrive pre-divided by the destination degree (p.gout = grad_out * inv_deg, a node-level
    // elementwise op done by the host wrapper), so nothing here depends on a second load.
    float g0[(MODE == CG_BWD_SRC) ? kRowsPerWarp : 1], g1[(MODE == CG_BWD_SRC) ? kRowsPerWarp : 1];
    if (MODE == CG_BWD_SRC) {  // destinations are scattered here: coalesced row loads, applied in the scale pass
#pragma unroll
      for (int i = 0; i < kRowsPerWarp; ++i) {
        const int e = row0 + i;
        g0[i] = g1[i] = 0.0f;
        if (e < cnt) {
          const int d = bDst[e];
          g0[i] = __ldg(p.gout + (size_t)d * C + lane);
          g1[i] = __ldg(p.gout + (size_t)d * C + 32 + lane);
        }
      }
    }
    // BWD_DST: slots are destination-sorted, so the 32 slots of a warp touch only a few distinct
    // grad_out rows -- each epilogue thread loads its own 16 channels directly (L1 hits, few
    // wavefronts) and the factor goes straight into the gate-derivative math: no scale pass.
    float4 gq[(MODE == CG_BWD_DST) ? 4 : 1];
    if (MODE == CG_BWD_DST) {
      const int e = 32 * q + lane;
#pragma unroll
      for (int j = 0; j < 4; ++j) gq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < cnt) {
        const float4* gp = reinterpret_cast<const float4*>(p.gout + (size_t)bDst[e] * C + c_begin);
#pragma unroll
        for (int j = 0; j < 4; ++j) gq[j] = __ldg(gp + j);
      }
    }
    mark(4);

    // ---- overlap window, part 2: LAND: this round's node rows, then the next round's indices and,
    // from them, its ea rows.
    if (nk < my_tiles) land_idx_and_rows(ni, ncnt, buf ^ 1);
    mark(21);

    // ---- epilogue, part 1: thread = slot (TMEM lane); a = accumulator + (P[dst] + Q[src]), the node
    // terms read from the window rows (or, without a window, from the slot's own value-tile row)
    float f[16], sacc[16];
    const int e_ep = 32 * q + lane;
    const bool live = e_ep < cnt;
    if (cnt > 0) {
      wait_dbg(&bar_rows, ph_rows, 2);  // node rows landed (each thread observes the barrier itself)
      ph_rows ^= 1;
      mark(20);
      wait_dbg(&bar, phase, 3);
      umma::fence_after_sync();
      phase ^= 1;
      mark(6);
      umma::tmem_ld16(umma::tmem_addr(tmem, q, c_begin), f);
      umma::tmem_ld16(umma::tmem_addr(tmem, q, C + c_begin), sacc);
      umma::tmem_ld_wait();
      mark(22);
      if (live) {
        const int sd = bDst[e_ep], ss = bSrc[e_ep];
        auto add_rows = [&](const float* r0, const float* r1) {  // a += r0[.] + r1[.]   ([f | s] rows)
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {
            const float4 pf = *reinterpret_cast<const float4*>(r0 + j4);
            const float4 ps = *reinterpret_cast<const float4*>(r0 + C + j4);
            const float4 qf = *reinterpret_cast<const float4*>(r1 + j4);
            const float4 qs = *reinterpret_cast<const float4*>(r1 + C + j4);
            f[j4] += pf.x + qf.x; f[j4 + 1] += pf.y + qf.y; f[j4 + 2] += pf.z + qf.z; f[j4 + 3] += pf.w + qf.w;
            sacc[j4] += ps.x + qs.x; sacc[j4 + 1] += ps.y + qs.y;
            sacc[j4 + 2] += ps.z + qs.z; sacc[j4 + 3] += ps.w + qs.w;
          }
        };
        if (win) {
          add_rows(sV + (w_nq + sd - w_dmin) * VW + c_begin, sV + (ss - w_smin) * VW + c_begin);
        } else {
          const float* direct = (MODE == CG_BWD_SRC) ? p.PQ + (size_t)ss * (4 * C) + 2 * C + c_begin
                                                     : p.PQ + (size_t)sd * (4 * C) + c_begin;
          add_rows(direct, sV + e_ep * VW + c_begin);
        }
      }
    }
    mark(23);
    __syncthreads();  // [S2d] every read of the staged node rows done: the value tile may be overwritten
    mark(13);
    // ---- epilogue, part 2: gate math, per-slot values parked in the value tile
    if (live) {
      float* rowv = sV + e_ep * VW;
#pragma unroll
      for (int j4 = 0; j4 < 16; j4 += 4) {
        const int c = c_begin + j4;
        float r0[4], r1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float sg = GATE ? sigmoid_mixed(f[j4 + j]) : sigmoid_mufu(f[j4 + j]);
          if (MODE == CG_FWD) {
            r0[j] = sg * softplus_mufu(sacc[j4 + j]);
          } else {  // d m / d a_f and d m / d a_s  (x grad_out here for BWD_DST, in the scale pass for BWD_SRC)
            float sp, sgs;
            softplus_sigmoid_mufu(sacc[j4 + j], sp, sgs);
            float g = 1.0f;
            if (MODE == CG_BWD_DST) {
              const float4 gv = gq[j4 >> 2];
              g = j == 0 ? gv.x : j == 1 ? gv.y : j == 2 ? gv.z : gv.w;
            }
            r0[j] = g * sp * sg * (1.0f - sg);
            r1[j] = g * sg * sgs;
          }
        }
        *reinterpret_cast<float4*>(rowv + c) = make_float4(r0[0], r0[1], r0[2], r0[3]);
        if (MODE != CG_FWD)
          *reinterpret_cast<float4*>(rowv + C + c) = make_float4(r1[0], r1[1], r1[2], r1[3]);
      }
    }
    mark(7);
    umma::fence_before_sync();  // accumulator reads done before the next round's MMAs overwrite it
    __syncthreads();            // [S3] value tile complete
    mark(8);

    // ---- BWD_SRC: da = grad_out[dst] * bracket, applied row-wise (lanes = channels)
    if (MODE == CG_BWD_SRC) {
#pragma unroll
      for (int i = 0; i < kRowsPerWarp; ++i) {
        const int e = row0 + i;
        if (e < cnt) {
          float* rowv = sV + e * VW;
          rowv[lane] *= g0[i];
          rowv[32 + lane] *= g1[i];
          rowv[C + lane] *= g0[i];
          rowv[C + 32 + lane] *= g1[i];
        }
      }
      __syncthreads();  // [S3b]
      mark(11);  // scale pass
    }
    // ---- BWD_DST single-pass mode: dQ[src] += da, a whole 512-byte row per warp instruction
    // (lane l owns floats [4l, 4l+4) of [d a_f | d a_s]); the order of the float adds is not fixed
    if (MODE == CG_BWD_DST && pl.dq_atomic) {
#pragma unroll
      for (int i = 0; i < kRowsPerWarp; ++i) {
        const int e = row0 + i;
        if (e < cnt) {
          const float4 v = *(reinterpret_cast<const float4*>(sV + e * VW) + lane);
          red_add_v4(p.out + (size_t)bSrc[e] * (4 * C) + 2 * C + 4 * lane, v);
        }
      }
      mark(12);  // dQ atomics
    }

    // ---- segmented sum over the owned segments that have slots in this round
    for (int n = n0; n < n_hi; n += kTcWarps) {
      int a, b;
      if (n == n0) { a = seg_a; b = seg_b; }
      else { a = __ldg(p.seg_ptr + n); b = __ldg(p.seg_ptr + n + 1); }
      const int lo = max(a, r_lo), hi = min(b, r_hi);
      const bool empty_seg = (a == b);
      if (empty_seg ? (rd != 0) : (lo >= hi)) continue;
      const bool first = empty_seg || (a >= r_lo);
      const bool last = empty_seg || (b <= r_hi);
      if (MODE == CG_FWD) {
        float* o = p.out + (size_t)n * C;
        float sc = seg_sc, x0 = seg_x0, x1 = seg_x1;
        if (n != n0) {
          sc = p.inv_deg ? __ldg(p.inv_deg + n) : 1.0f;
          x0 = __ldg(p.x + (size_t)n * C + lane);
          x1 = __ldg(p.x + (size_t)n * C + 32 + lane);
        }
        float acc0 = first ? 0.0f : o[lane], acc1 = first ? 0.0f : o[32 + lane];
        for (int s = lo; s < hi; ++s) {
          acc0 += sV[(s - r_lo) * VW + lane];
          acc1 += sV[(s - r_lo) * VW + 32 + lane];
        }
        o[lane] = last ? fmaf(acc0, sc, x0) : acc0;
        o[32 + lane] = last ? fmaf(acc1, sc, x1) : acc1;
      } else {
        float* o = p.out + (size_t)n * (4 * C) + (MODE == CG_BWD_SRC ? 2 * C : 0);
        float acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = first ? 0.0f : o[32 * u + lane];
        for (int s = lo; s < hi; ++s) {
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] += sV[(s - r_lo) * VW + 32 * u + lane];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) o[32 * u + lane] = acc[u];
      }
    }
    mark(9);

    // ---- dWe^T[ch, k] += sum_slots da[slot, ch] * ea[slot, k] on the warp-level tensor path
    // (mma.sync m16n8k8 tf32, 3xTF32).  Warp (mgrp, ngrp) owns 32 channels x 16 k-columns = 2x2
    // MMA tiles whose rows/columns are PERMUTED so that every fragment is a vector load:
    //   channel(tile mi, fragment row gid | gid+8) = mgrp*32 + 4*gid + 2*mi + {0 | 1}
    //   column (tile ni, fragment col gid)        = ngrp*16 + 2*gid + ni
    // A = da^T from the value tile (float4 per slot, split hi/lo in registers); B = ea straight
    // from the already split operand tiles (float2 per slot, element (slot e, column g) sits at
    // (g/4)*2048 + (e/8)*128 + (e%8)*16 + (g%4)*4 in sAhi / sAlo).
    if (MODE == CG_BWD_DST && cnt > 0) {
      const int gid = lane >> 2, tig = lane & 3;
      const int mgrp = warp & 3, ngrp = warp >> 2;
      const int chunk = ngrp * 4 + (gid >> 1);           // 16-byte chunk of the operand tiles
      const bool col_ok = chunk < (KP >> 2);
      const uint32_t cb = (uint32_t)chunk * kAChunk + (uint32_t)(gid & 1) * 8;
      for (int e0 = 0; e0 < cnt; e0 += 8) {
        const int ea_ = e0 + tig, eb_ = e0 + tig + 4;
        float4 a_lo_row = make_float4(0.f, 0.f, 0.f, 0.f), a_hi_row = a_lo_row;
        if (ea_ < cnt) a_lo_row = *reinterpret_cast<const float4*>(sV + ea_ * VW + mgrp * 32 + 4 * gid);
        if (eb_ < cnt) a_hi_row = *reinterpret_cast<const float4*>(sV + eb_ * VW + mgrp * 32 + 4 * gid);
        float2 bh0 = make_float2(0.f, 0.f), bh1 = bh0, bl0 = bh0, bl1 = bh0;
        if (col_ok) {
          const uint32_t rb0 = (uint32_t)(ea_ >> 3) * 128 + (uint32_t)(ea_ & 7) * 16;
          const uint32_t rb1 = (uint32_t)(eb_ >> 3) * 128 + (uint32_t)(eb_ & 7) * 16;
          bh0 = *reinterpret_cast<const float2*>(sAhi + cb + rb0);  // rows >= cnt are zero in the tiles
          bh1 = *reinterpret_cast<const float2*>(sAhi + cb + rb1);
          bl0 = *reinterpret_cast<const float2*>(sAlo + cb + rb0);
          bl1 = *reinterpret_cast<const float2*>(sAlo + cb + rb1);
        }
        const float ar0[4] = {a_lo_row.x, a_lo_row.y, a_lo_row.z, a_lo_row.w};  // slot e0+tig
        const float ar1[4] = {a_hi_row.x, a_hi_row.y, a_hi_row.z, a_hi_row.w};  // slot e0+tig+4
        float ah[2][4], al[2][4];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          // fragment order: (row gid, k tig), (row gid+8, k tig), (row gid, k tig+4), (row gid+8, k tig+4)
          const float av[4] = {ar0[2 * mi], ar0[2 * mi + 1], ar1[2 * mi], ar1[2 * mi + 1]};
#pragma unroll
          for (int u = 0; u < 4; ++u) { ah[mi][u] = umma::tf32_hi(av[u]); al[mi][u] = av[u] - ah[mi][u]; }
        }
        // pass-major order: four independent accumulators back to back (no dependent MMA chain)
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
          for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
              const float b0 = (pass == 1) ? (ni ? bl0.y : bl0.x) : (ni ? bh0.y : bh0.x);
              const float b1 = (pass == 1) ? (ni ? bl1.y : bl1.x) : (ni ? bh1.y : bh1.x);
              mma_tf32_16x8x8(dacc[2 * mi + ni], (pass == 2) ? al[mi] : ah[mi], b0, b1);
            }
      }
    }
    mark(10);
    if (PROFILE && pl.prof && tid == 0) atomicAdd(pl.prof + 31, 1ull);
    k = nk; rd = nrd; buf ^= 1; ++dbg_round;
  }  // work items

  if (MODE == CG_BWD_DST) {
    // C fragment of tile (mi, ni): rows gid | gid+8 -> channels 2*mi | 2*mi+1 of this lane's quad,
    // cols 2*tig | 2*tig+1 -> fragment columns, i.e. k = ngrp*16 + 2*(2*tig | 2*tig+1) + ni
    const int gid = lane >> 2, tig = lane & 3;
    const int mgrp = warp & 3, ngrp = warp >> 2;
    float* part = p.dW_part + (size_t)blockIdx.x * G * W2;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) {
        const int ch = mgrp * 32 + 4 * gid + 2 * mi;
        const int k0 = ngrp * 16 + 2 * (2 * tig) + ni, k1 = ngrp * 16 + 2 * (2 * tig + 1) + ni;
        const float* d = dacc[2 * mi + ni];
        if (k0 < G) { part[(size_t)k0 * W2 + ch] = d[0]; part[(size_t)k0 * W2 + ch + 1] = d[2]; }
        if (k1 < G) { part[(size_t)k1 * W2 + ch] = d[1]; part[(size_t)k1 * W2 + ch + 1] = d[3]; }
      }
  }
  cp_async_wait_all();
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, (uint32_t)pl.tmem_cols);
}

template <int MODE, int PROFILE, int GATE, int EA_BULK>
static int tc_launch_t(const CgParams& p, const TcPlan& pl, int grid, cudaStream_t st) {
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MDL_CUDA(cudaFuncSetAttribute(k_cgconv_tc<MODE, PROFILE, GATE, EA_BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kMaxDynSmem));
    configured.store(1, std::memory_order_release);
  }
  k_cgconv_tc<MODE, PROFILE, GATE, EA_BULK><<<grid, kTcThreads, pl.total, st>>>(p, pl);
  MDL_LAUNCHED();
  return MDL_OK;
}

void cgtc_set_phase_buffer(unsigned long long* dev_ptr) { g_phase_buf = dev_ptr; }

bool cgtc_supported(int mode, int C, int G) {
  TcPlan pl;
  return tc_plan(mode, C, G, &pl);
}

int cgtc_launch(int mode, CgParams p, cudaStream_t st, int* grid_out, int dq_atomic) {
  TcPlan pl;
  MDL_REQUIRE(tc_plan(mode, p.C, p.G, &pl), "cgconv_tc: unsupported shape C=%d G=%d", p.C, p.G);
  pl.dq_atomic = (mode == CG_BWD_DST) ? dq_atomic : 0;
  const char* wenv = getenv("MDL_CGCONV_WINDOW");  // "0": always gather per slot (A/B and test switch)
  pl.window = !(wenv && wenv[0] == '0');
  p.c_off = 0; p.CC = p.C; p.cap = kTcRows; p.te = kTcTE;
  p.n_tiles = (int)std::max<int64_t>(1, ceil_div<int64_t>(p.E, kTcTE));
  const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
  if (grid_out) *grid_out = grid;
  const bool prof = pl.prof != nullptr;  // instrumented instantiation only while a phase buffer is set
  const char* genv = getenv("MDL_CGCONV_GATE");  // "mufu": every reciprocal on the transcendental pipe
  const bool mixed = !(genv && genv[0] == 'm' && genv[1] == 'u');
  const char* benv = getenv("MDL_CGCONV_EA");  // "rows": per-row cp.async instead of the bulk copy
  const bool bulk = mode != CG_BWD_SRC && (reinterpret_cast<uintptr_t>(p.ea) & 15) == 0 && !(benv && benv[0] == 'r');
  MDL_REQUIRE((reinterpret_cast<uintptr_t>(p.PQ) & 15) == 0, "cgconv_tc: PQ must be 16-byte aligned");
#define MDL_TC_G(M, PR, GT) (bulk ? tc_launch_t<M, PR, GT, 1>(p, pl, grid, st) : tc_launch_t<M, PR, GT, 0>(p, pl, grid, st))
#define MDL_TC_CASE(M) \
  case M:              \
    return prof ? (mixed ? MDL_TC_G(M, 1, 1) : MDL_TC_G(M, 1, 0)) : (mixed ? MDL_TC_G(M, 0, 1) : MDL_TC_G(M, 0, 0));
  switch (mode) {
    MDL_TC_CASE(CG_FWD)
    MDL_TC_CASE(CG_BWD_DST)
    case CG_BWD_SRC:
      return prof ? (mixed ? tc_launch_t<CG_BWD_SRC, 1, 1, 0>(p, pl, grid, st) : tc_launch_t<CG_BWD_SRC, 1, 0, 0>(p, pl, grid, st))
                  : (mixed ? tc_launch_t<CG_BWD_SRC, 0, 1, 0>(p, pl, grid, st) : tc_launch_t<CG_BWD_SRC, 0, 0, 0>(p, pl, grid, st));
  }
#undef MDL_TC_G
#undef MDL_TC_CASE
  MDL_REQUIRE(false, "cgconv_tc: bad mode");
}

}  // namespace mdl
