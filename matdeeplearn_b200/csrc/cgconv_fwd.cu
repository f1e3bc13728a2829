// cgconv_fwd.cu -- software-pipelined forward kernel of the fused CGConv edge op (C = 64).
//
// Same operator, tile ownership, 3xTF32 contraction, gate math and deterministic per-segment sum as
// k_cgconv_tc<FWD> (cgconv_tc.cu); what changes is the schedule.  There, a round of <= 128 slots runs
// its phases back to back (split -> MMA -> epilogue -> reduce) and the tensor core's ~1.7k cycles sit
// on the critical path.  Here the contraction runs ONE ROUND AHEAD into a second TMEM accumulator:
//
//   iteration r:   [S1]  node rows of round r requested (LDG.128 -> registers), indices of round r+1 requested
//                        wait MMA(r)                         (issued an iteration ago: long finished)
//                        split ea(r+1): landing zone -> hi / lo -> tcgen05.st (the A operand lives in TENSOR
//                        MEMORY: no operand tiles in shared memory; its bulk copy was issued an iteration ago)
//                        node rows / next indices -> smem
//                  [S2]  issuer warp (its own warp, it does nothing else): bulk copy of ea(r+2), then the 21 MMAs
//                        of round r+1 into the other accumulator -- issuing them blocks a thread for ~2.4k cycles
//                        epilogue(r): tcgen05.ld -> + P[dst] + Q[src] (node-row tile) -> gates -> message tile
//                  [S3]  per-segment sums of round r -> out
//
// so the tensor core, the bulk copy engine and the L2 round trips of the node rows all run under the
// epilogue of the previous round.  640 threads: 16 epilogue warps + one warpgroup for the issuer warp, with
// setmaxnreg moving registers from the latter to the former (112 / 32).  The backward kernels cannot use this
// schedule with the same shared memory: their dW_e stage reads round r's operand tiles at the END of the round.
//
// Node terms: the P / Q rows a round needs lie in two short contiguous node ranges (edges never leave
// their crystal graph); when both fit in 128 rows they are staged in the node-row tile (separate from the
// message tile, so row reads and gate math share one phase) and the epilogue reads them from shared memory
// ("window"), else Q[src[e]] is staged per slot and P[dst[e]] (few distinct rows per warp, slots are
// destination-sorted) is read straight from global memory.
#include "cgconv.cuh"
#include "umma.cuh"
#include "edge_dev.cuh"

namespace mdl {

namespace {

constexpr int kThreadsF = 512;                    // epilogue ("consumer") threads: 16 warps
constexpr int kWarpsF = kThreadsF / 32;
constexpr int kS2Threads = kThreadsF + 32;        // + one warp that only issues MMAs and bulk copies
constexpr int kLaunchF = kThreadsF + 128;         // registers are re-balanced per warpgroup: launch a whole one for it
constexpr int kRowsF = 128;                       // slots per round = MMA M
constexpr int kTileSlots = 112;                   // ownership granularity (as cgconv_tc.cu)
constexpr int kInfoCapF = 512;
constexpr int kC = 64, kNP = 2 * kC;
constexpr int kVW = 2 * kC + 4;                   // row stride of the node-row tile ([f | s] rows + 16 B: bank spread)
constexpr int kVP = kC + 4;                       // row stride of the per-slot message tile
constexpr int kTmemCols = 512;                    // 2 accumulators (2 x 128) + A operand hi / lo (2 x KP <= 256)
constexpr int kRowRegs = 4;                       // 16-byte node-row chunks a thread keeps in flight across the split (64 rows)

unsigned long long* g_fwd_phase_buf = nullptr;

struct FwdPlan {
  unsigned long long* prof;
  int window, KP;
  uint32_t offBhi, offBlo, offEA, offW, offV, offIdx, offInfo, total;
};

bool fwd_plan(int C, int G, FwdPlan* pl) {
  if (C != kC || G < 1) return false;
  const int KP = (G + 7) & ~7;
  if (2 * kNP + 2 * KP > kTmemCols) return false;
  const uint32_t b = (uint32_t)kNP * KP * 4;
  const uint32_t ea = (((uint32_t)kRowsF * G * 4 + 32) + 15u) & ~15u;  // dense rows + alignment slack of the bulk copy
  const uint32_t w = (uint32_t)kRowsF * kVW * 4, v = (uint32_t)kRowsF * kVP * 4;
  const uint32_t idx = 4 * kRowsF * 4, info = kInfoCapF * 16;
  pl->prof = g_fwd_phase_buf;
  pl->window = 1;
  pl->KP = KP;
  pl->offBhi = 0; pl->offBlo = b; pl->offEA = 2 * b;
  pl->offW = pl->offEA + ea; pl->offV = pl->offW + w; pl->offIdx = pl->offV + v; pl->offInfo = pl->offIdx + idx;
  pl->total = pl->offInfo + info;
  return pl->total <= (uint32_t)kMaxDynSmem;
}

struct Round {
  int k, rd;      // tile of this CTA (k >= my_tiles: no such round), round inside the tile
  int r_lo, cnt;  // slots [r_lo, r_lo + cnt)
  bool last;      // last round of its tile
};

template <int PROFILE>
__global__ void __launch_bounds__(kLaunchF, 1) k_cgconv_fwd_pipe(const CgParams p, const FwdPlan pl) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_mma;  // tcgen05.commit of a round's MMAs
  __shared__ uint64_t bar_ea;   // bulk copy of a round's edge rows
  __shared__ uint32_t tmem_base_s;
  __shared__ int sMail[4];  // consumers -> issuer warp at [S2]: {round staged?, its cnt, next round's r_lo, cnt (or -1)}
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = p.G, KP = pl.KP;
  // barriers: 0 = whole CTA (setup / teardown), 1 = whole CTA at [S2] (hands the staged tiles to the
  // issuer warp), 2 = the 512 consumer threads only
  auto sync_consumers = [] { asm volatile("bar.sync 2, %0;" ::"n"(kThreadsF) : "memory"); };
  auto sync_s2 = [] { asm volatile("bar.sync 1, %0;" ::"n"(kS2Threads) : "memory"); };

  uint8_t* sBhi = smem + pl.offBhi;
  uint8_t* sBlo = smem + pl.offBlo;
  float* sEA = reinterpret_cast<float*>(smem + pl.offEA);  // dense [cnt][G] block at a 0..12 byte offset
  float* sW = reinterpret_cast<float*>(smem + pl.offW);    // [128][VW]: staged node rows (window or per slot)
  float* sV = reinterpret_cast<float*>(smem + pl.offV);    // [128][VP]: per-slot messages
  int* sIdx = reinterpret_cast<int*>(smem + pl.offIdx);    // [2 buffers][src | dst][128]
  TileInfo* sInfo = reinterpret_cast<TileInfo*>(smem + pl.offInfo);

  const int my_tiles = (p.n_tiles > (int)blockIdx.x) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  int info_base = 0;
  auto fill_infos = [&](int base) {
    if (tid >= kThreadsF) return;
    for (int k = base + tid; k < min(my_tiles, base + kInfoCapF); k += kThreadsF) {
      TileInfo t;
      const int tile = blockIdx.x + k * gridDim.x;
      t.n_lo = first_segment_at_or_after<CG_FWD>(p, tile * kTileSlots);
      t.n_hi = (tile == p.n_tiles - 1) ? p.N : first_segment_at_or_after<CG_FWD>(p, (tile + 1) * kTileSlots);
      if (t.n_hi < t.n_lo) t.n_hi = t.n_lo;
      t.e_lo = __ldg(p.seg_ptr + t.n_lo);
      t.e_hi = __ldg(p.seg_ptr + t.n_hi);
      sInfo[k - base] = t;
    }
  };
  auto make_round = [&](int k, int rd) -> Round {
    Round R{k, rd, 0, 0, true};
    if (k < my_tiles) {
      const TileInfo T = sInfo[k - info_base];
      R.r_lo = T.e_lo + rd * kRowsF;
      R.cnt = max(0, min(T.e_hi - R.r_lo, kRowsF));
      R.last = R.r_lo + kRowsF >= T.e_hi;
    }
    return R;
  };
  auto valid = [&](const Round& R) { return R.k < my_tiles; };
  auto next_round = [&](const Round& R) -> Round { return R.last ? make_round(R.k + 1, 0) : make_round(R.k, R.rd + 1); };

  // ---- one-time setup: TMEM (two accumulators), barriers, tile table, resident W_e split hi/lo
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, kTmemCols);
  if (tid == 32) {
    umma::mbar_init(&bar_mma, 1);
    umma::mbar_init(&bar_ea, 1);
    umma::fence_mbar_init();
  }
  fill_infos(0);
  for (int i = tid; i < kNP * KP; i += kLaunchF) {
    const int n = i % kNP, k = i / kNP;
    const float w = (k < G) ? __ldg(p.WeT + (size_t)k * kNP + n) : 0.0f;
    const float hi = umma::tf32_hi(w);
    const int off = umma::tile_offset_bytes(n, k, kNP);
    *reinterpret_cast<float*>(sBhi + off) = hi;
    *reinterpret_cast<float*>(sBlo + off) = w - hi;
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma::make_idesc_tf32(kRowsF, kNP);
  const uint32_t tmA_hi = tmem + 2 * kNP, tmA_lo = tmA_hi + (uint32_t)KP;  // A operand (edge rows) in tensor memory
  uint32_t ph_mma = 0, ph_ea = 0;

  // ---- edge rows of a round: one bulk copy from the 16-byte boundary below the block (see cgconv_tc.cu)
  auto ea_bulk_bytes = [&](int r_lo, int cnt) -> uint32_t {
    if (cnt <= 0) return 0u;
    const long long first = (long long)r_lo * G;
    const uint32_t bytes = (uint32_t)(((int)(first & 3) + cnt * G) * 4);
    const bool more = ((long long)p.E * G - (first + (long long)cnt * G)) >= 3;
    return more ? ((bytes + 15u) & ~15u) : (bytes & ~15u);
  };
  auto issue_ea_bulk = [&](int r_lo, int cnt) {  // one thread
    if (cnt <= 0) return;
    const long long first = (long long)r_lo * G;
    const int off = (int)(first & 3);
    const float* src = p.ea + (first - off);
    const uint32_t bytes = (uint32_t)((off + cnt * G) * 4), nb = ea_bulk_bytes(r_lo, cnt);
    (void)bytes;
    if (nb) {
      umma::mbar_arrive_expect_tx(&bar_ea, nb);
      umma::bulk_g2s(sEA, src, nb, &bar_ea);
    }
    // The last block of ea is copied rounded DOWN to 16 bytes; its last (<= 3) floats are read from
    // global memory by the split itself (front()).  Storing them here would race with the split: the
    // mbarrier only orders the bulk copy's bytes, and no barrier joins this thread and the readers.
  };

  long long t_prev = PROFILE ? clock64() : 0;
  auto mark = [&](int slot) {
    if (PROFILE && pl.prof && tid == 0) {
      const long long now = clock64();
      atomicAdd(pl.prof + slot, (unsigned long long)(now - t_prev));
      t_prev = now;
    }
  };

  // ---- the front half of a round: edge rows -> split operand tiles -> (mid) -> barrier
  auto front = [&](const Round& X, auto&& mid) {
    if (valid(X)) {
      const int ea_off = (int)(((long long)X.r_lo * G) & 3);
      if (ea_bulk_bytes(X.r_lo, X.cnt)) {  // CTA-uniform
        umma::mbar_wait(&bar_ea, ph_ea);
        ph_ea ^= 1;
      }
      mark(2);
      // thread = slot = TMEM lane (quadrant warp & 3); the four warps of a quadrant share the row's
      // 8-column chunks.  The split halves go straight to tensor memory (A operand of the MMAs): no
      // operand tiles in shared memory, and the MMAs read half as much of it.
      const int e = tid & (kRowsF - 1);
      const float* row = sEA + ea_off + e * G;
      // first float (index into the landing zone) the bulk copy did NOT deliver: only in the last block of ea
      const int landed = (int)(ea_bulk_bytes(X.r_lo, X.cnt) >> 2);
      const bool patch = e < X.cnt && landed < ea_off + (e + 1) * G;  // this row reaches past the delivered part
      for (int ch = (tid >> 7); ch < (KP >> 3); ch += kThreadsF / kRowsF) {
        float v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = 0.0f;
        if (e < X.cnt) {
          if ((G & 1) == 0) {  // rows start at an 8-byte offset: 8-byte loads
#pragma unroll
            for (int t = 0; t < 8; t += 2)
              if (8 * ch + t < G) {
                const float2 a = *reinterpret_cast<const float2*>(row + 8 * ch + t);
                v[t] = a.x; v[t + 1] = a.y;
              }
          } else {
#pragma unroll
            for (int t = 0; t < 8; ++t)
              if (8 * ch + t < G) v[t] = row[8 * ch + t];
          }
        }
        if (patch) {
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int k = 8 * ch + t;
            if (k < G && ea_off + e * G + k >= landed) v[t] = __ldg(p.ea + ((long long)X.r_lo + e) * G + k);
          }
        }
        float hi[8], lo[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) { hi[t] = umma::tf32_hi(v[t]); lo[t] = v[t] - hi[t]; }
        umma::tmem_st8(umma::tmem_addr(tmA_hi, warp, 8 * ch), hi);
        umma::tmem_st8(umma::tmem_addr(tmA_lo, warp, 8 * ch), lo);
      }
      umma::tmem_st_wait();
      mark(3);
    }
    mid();
    if (tid == 0) {
      const Round Y = valid(X) ? next_round(X) : X;
      sMail[0] = valid(X); sMail[1] = X.cnt;
      sMail[2] = Y.r_lo; sMail[3] = (valid(X) && valid(Y)) ? Y.cnt : -1;
    }
    umma::fence_proxy_async_smem();
    umma::fence_before_sync();
    sync_s2();  // [S2] operand tiles, staged node rows, next indices visible; landing zone free; issuer warp released
    mark(4);
  };
  // the 21 MMAs of a round into accumulator column `acc_col` (ONE thread; the tiles were staged by front())
  auto issue_mma = [&](uint32_t acc_col) {
    umma::fence_after_sync();
    const uint32_t step_b = 2 * (uint32_t)kNP * 16;
    const uint32_t b_hi = umma::smem_u32(sBhi), b_lo = umma::smem_u32(sBlo);
    uint32_t acc = 0;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t a = (pass == 2) ? tmA_lo : tmA_hi;
      const uint32_t b = (pass == 1) ? b_lo : b_hi;
      for (int kk = 0; kk < (KP >> 3); ++kk) {
        const uint64_t bd = umma::make_desc(b + kk * step_b, (uint32_t)kNP * 16, 128);
        umma::mma_tf32_ts(tmem + acc_col, a + kk * 8, bd, idesc, acc);
        acc = 1;
      }
    }
    umma::mma_commit(&bar_mma);
  };

  // ---- the issuer warp: per staged round, the MMAs into the alternating accumulator and the bulk copy
  // of the FOLLOWING round's edge rows (the landing zone is free once the split is done).  tcgen05.mma
  // issue blocks for most of the MMAs' run time (~2.4k cycles per round measured), which is why this
  // is not an epilogue warp's side job.
  if (warp >= kWarpsF) {
    // 640 threads leave 96 registers each at launch; this warpgroup hands most of its share to the
    // epilogue warpgroups (512 x 120 + 128 x 32 = 64 K)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    for (uint32_t i = 0; warp == kWarpsF; ++i) {
      sync_s2();
      const int staged = sMail[0], xcnt = sMail[1], y_lo = sMail[2], ycnt = sMail[3];
      if (!staged) break;
      if (lane == 0) {
        if (ycnt >= 0) issue_ea_bulk(y_lo, ycnt);  // first: the MMA issue below blocks for ~2k cycles
        if (xcnt > 0) issue_mma((i & 1) * kNP);
      }
      __syncwarp();
    }
    __syncthreads();  // teardown barrier of the CTA
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");

  const int q = warp & 3, part = warp >> 2;  // TMEM lane quadrant, channel quarter (16 channels)
  const int c_begin = part * 16;

  // ---- prologue: indices and edge rows of the first round, its contraction started
  Round cur = make_round(0, 0);
  Round nxt = next_round(cur);
  int buf = 0;
  uint32_t it = 0;
  if (valid(cur)) {
    if (tid == 0) issue_ea_bulk(cur.r_lo, cur.cnt);
    if (tid < 2 * kRowsF) {
      const int e = tid & (kRowsF - 1);
      if (e < cur.cnt) sIdx[tid] = __ldg((tid < kRowsF ? p.dst_src : p.dst_dst) + cur.r_lo + e);
    }
    front(cur, [] {});
  } else {
    if (tid == 0) sMail[0] = 0;
    sync_s2();  // releases the issuer warp, which leaves
  }

  while (valid(cur)) {
    if (cur.k + 2 >= info_base + kInfoCapF && info_base + kInfoCapF < my_tiles) {  // table exhausted: refill
      sync_consumers();
      info_base = cur.k;
      fill_infos(info_base);
      sync_consumers();
    }
    const int cnt = cur.cnt, r_lo = cur.r_lo, r_hi = cur.r_lo + cur.cnt;
    const int n_lo = sInfo[cur.k - info_base].n_lo, n_hi = sInfo[cur.k - info_base].n_hi;
    const int* bSrc = sIdx + buf * 2 * kRowsF;
    const int* bDst = bSrc + kRowsF;
    mark(0);
    sync_consumers();  // [S1] value tile free (last round's sums done); this round's indices visible
    mark(1);

    // ---- indices of the next round: one coalesced load per thread (slots are consecutive)
    int nidx = 0;
    if (valid(nxt) && tid < 2 * kRowsF) {
      const int e = tid & (kRowsF - 1);
      if (e < nxt.cnt) nidx = __ldg((tid < kRowsF ? p.dst_src : p.dst_dst) + nxt.r_lo + e);
    }
    // ---- node rows of this round: window decision (every warp derives it from the index tile) and
    // the row chunks requested into registers; they land in the value tile inside front()
    bool win = false;
    int w_smin = 0, w_dmin = 0, w_nq = 0, nrows = 0;
    if (cnt > 0) {
      int s_lo = 0x7fffffff, s_hi = -1;
      {
        const int4 sv = *reinterpret_cast<const int4*>(bSrc + 4 * lane);
        const int e0 = 4 * lane;
        if (e0 + 0 < cnt) { s_lo = min(s_lo, sv.x); s_hi = max(s_hi, sv.x); }
        if (e0 + 1 < cnt) { s_lo = min(s_lo, sv.y); s_hi = max(s_hi, sv.y); }
        if (e0 + 2 < cnt) { s_lo = min(s_lo, sv.z); s_hi = max(s_hi, sv.z); }
        if (e0 + 3 < cnt) { s_lo = min(s_lo, sv.w); s_hi = max(s_hi, sv.w); }
      }
      s_lo = __reduce_min_sync(0xffffffffu, s_lo);
      s_hi = __reduce_max_sync(0xffffffffu, s_hi);
      const int d_lo = bDst[0], d_hi = bDst[cnt - 1];  // slots are sorted by destination
      const int nq = s_hi - s_lo + 1, np_ = d_hi - d_lo + 1;
      win = pl.window && nq + np_ <= kRowsF;
      w_smin = s_lo; w_dmin = d_lo; w_nq = nq;
      nrows = win ? nq + np_ : cnt;
    }
    auto row_src = [&](int r) -> const float4* {  // global 512-byte row that value-tile row r stages
      const float* g;
      if (win) g = (r < w_nq) ? p.PQ + (size_t)(w_smin + r) * (4 * kC) + 2 * kC : p.PQ + (size_t)(w_dmin + r - w_nq) * (4 * kC);
      else g = p.PQ + (size_t)bSrc[r] * (4 * kC) + 2 * kC;
      return reinterpret_cast<const float4*>(g);
    };
    mark(13);
    float4 rr[kRowRegs];  // rows 0..63 stay in flight across the split; rows 64.. (rare) are copied in mid
#pragma unroll
    for (int i = 0; i < kRowRegs; ++i) {
      const int c = tid + kThreadsF * i, r = c >> 5, col = c & 31;
      if (r < nrows) rr[i] = __ldg(row_src(r) + col);
    }
    mark(14);
    // ---- reduce-stage node data of this warp's first segment
    const int n0 = n_lo + warp;
    int seg_a = 0, seg_b = 0;
    float seg_sc = 1.0f;
    float2 seg_x = make_float2(0.0f, 0.0f);  // lane l owns channels 2l, 2l+1
    if (n0 < n_hi) {
      seg_a = __ldg(p.seg_ptr + n0);
      seg_b = __ldg(p.seg_ptr + n0 + 1);
      if (p.inv_deg) seg_sc = __ldg(p.inv_deg + n0);
      seg_x = __ldg(reinterpret_cast<const float2*>(p.x + (size_t)n0 * kC) + lane);
    }
    mark(5);

    // ---- this round's contraction (issued an iteration ago) must have retired before its operand
    // tiles are overwritten -- and its accumulator is what the epilogue below reads
    if (cnt > 0) {
      umma::mbar_wait(&bar_mma, ph_mma);
      ph_mma ^= 1;
      umma::fence_after_sync();
    }
    mark(6);
    auto mid = [&] {
#pragma unroll
      for (int i = 0; i < kRowRegs; ++i) {
        const int c = tid + kThreadsF * i, r = c >> 5, col = c & 31;
        if (r < nrows) *(reinterpret_cast<float4*>(sW + r * kVW) + col) = rr[i];
      }
      for (int c = tid + kThreadsF * kRowRegs; c < nrows * 32; c += kThreadsF)
        *(reinterpret_cast<float4*>(sW + (c >> 5) * kVW) + (c & 31)) = __ldg(row_src(c >> 5) + (c & 31));
      if (valid(nxt) && tid < 2 * kRowsF) sIdx[(buf ^ 1) * 2 * kRowsF + tid] = nidx;
    };
    front(nxt, mid);

    // ---- epilogue: thread = slot (TMEM lane), 16 channels; a = accumulator + P[dst] + Q[src] (node rows
    // from the staged tile), gates, message parked in the message tile.  Row reads (LDS pipe) and gate
    // math (MUFU pipe) of different warps overlap: one phase, no barrier in between.
    const int e_ep = 32 * q + lane;
    const bool live = e_ep < cnt;
    if (cnt > 0) {
      float f[16], sacc[16];
      const uint32_t acc_col = (it & 1) * kNP;
      umma::tmem_ld16(umma::tmem_addr(tmem, q, acc_col + c_begin), f);
      umma::tmem_ld16(umma::tmem_addr(tmem, q, acc_col + kC + c_begin), sacc);
      umma::tmem_ld_wait();
      mark(7);
      if (live) {
        const int sd = bDst[e_ep], ss = bSrc[e_ep];
        const float* r0 = win ? sW + (w_nq + sd - w_dmin) * kVW + c_begin : p.PQ + (size_t)sd * (4 * kC) + c_begin;
        const float* r1 = win ? sW + (ss - w_smin) * kVW + c_begin : sW + e_ep * kVW + c_begin;
        float* rowv = sV + e_ep * kVP + c_begin;
        auto run = [&](auto ld0) {  // ld0: how the P row is read (shared or global memory)
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {
            const float4 pf = ld0(r0 + j4), ps = ld0(r0 + kC + j4);
            const float4 qf = *reinterpret_cast<const float4*>(r1 + j4);
            const float4 qs = *reinterpret_cast<const float4*>(r1 + kC + j4);
            const float af[4] = {f[j4] + (pf.x + qf.x), f[j4 + 1] + (pf.y + qf.y), f[j4 + 2] + (pf.z + qf.z),
                                 f[j4 + 3] + (pf.w + qf.w)};
            const float as[4] = {sacc[j4] + (ps.x + qs.x), sacc[j4 + 1] + (ps.y + qs.y), sacc[j4 + 2] + (ps.z + qs.z),
                                 sacc[j4 + 3] + (ps.w + qs.w)};
            float m[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) m[j] = sigmoid_mixed(af[j]) * softplus_mufu(as[j]);
            *reinterpret_cast<float4*>(rowv + j4) = make_float4(m[0], m[1], m[2], m[3]);
          }
        };
        if (win) run([](const float* a) { return *reinterpret_cast<const float4*>(a); });
        else run([](const float* a) { return __ldg(reinterpret_cast<const float4*>(a)); });
      }
    }
    mark(10);
    umma::fence_before_sync();  // accumulator reads done before a later round's MMAs overwrite it
    sync_consumers();           // [S3] message tile complete
    mark(11);

    // ---- segmented sum over the owned segments that have slots in this round (slot order: deterministic)
    for (int n = n0; n < n_hi; n += kWarpsF) {
      int a, b;
      if (n == n0) { a = seg_a; b = seg_b; }
      else { a = __ldg(p.seg_ptr + n); b = __ldg(p.seg_ptr + n + 1); }
      const int lo = max(a, r_lo), hi = min(b, r_hi);
      const bool empty_seg = (a == b);
      if (empty_seg ? (cur.rd != 0) : (lo >= hi)) continue;
      const bool first = empty_seg || (a >= r_lo);
      const bool lastp = empty_seg || (b <= r_hi);
      float2* o = reinterpret_cast<float2*>(p.out + (size_t)n * kC) + lane;
      float sc = seg_sc;
      float2 x = seg_x;
      if (n != n0) {
        sc = p.inv_deg ? __ldg(p.inv_deg + n) : 1.0f;
        x = __ldg(reinterpret_cast<const float2*>(p.x + (size_t)n * kC) + lane);
      }
      float2 acc = first ? make_float2(0.0f, 0.0f) : *o;
      for (int s = lo; s < hi; ++s) {
        const float2 v = *(reinterpret_cast<const float2*>(sV + (s - r_lo) * kVP) + lane);
        acc.x += v.x; acc.y += v.y;
      }
      *o = lastp ? make_float2(fmaf(acc.x, sc, x.x), fmaf(acc.y, sc, x.y)) : acc;
    }
    mark(12);
    if (PROFILE && pl.prof && tid == 0) atomicAdd(pl.prof + 31, 1ull);
    cur = nxt; nxt = next_round(nxt); buf ^= 1; ++it;
  }

  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
}

template <int PROFILE>
int fwd_launch_t(const CgParams& p, const FwdPlan& pl, int grid, cudaStream_t st) {
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MDL_CUDA(cudaFuncSetAttribute(k_cgconv_fwd_pipe<PROFILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    configured.store(1, std::memory_order_release);
  }
  k_cgconv_fwd_pipe<PROFILE><<<grid, kLaunchF, pl.total, st>>>(p, pl);
  MDL_LAUNCHED();
  return MDL_OK;
}

}  // namespace

void cgfwd_set_phase_buffer(unsigned long long* dev_ptr) { g_fwd_phase_buf = dev_ptr; }

// The pipelined kernel needs C = 64, a shared-memory plan that fits, and 16-byte aligned ea / PQ bases
// (bulk copy, vector loads); anything else stays on k_cgconv_tc<FWD>.
bool cgfwd_supported(const CgParams& p) {
  FwdPlan pl;
  return fwd_plan(p.C, p.G, &pl) && (reinterpret_cast<uintptr_t>(p.ea) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.PQ) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 7) == 0 &&
         (reinterpret_cast<uintptr_t>(p.out) & 7) == 0 && (int64_t)p.N * 4 * p.C < (int64_t)1 << 31;
}

int cgfwd_launch(CgParams p, cudaStream_t st) {
  FwdPlan pl;
  MDL_REQUIRE(fwd_plan(p.C, p.G, &pl), "cgconv_fwd: unsupported shape C=%d G=%d", p.C, p.G);
  const char* wenv = getenv("MDL_CGCONV_WINDOW");  // "0": per-slot rows only (A/B and test switch)
  pl.window = !(wenv && wenv[0] == '0');
  p.c_off = 0; p.CC = p.C; p.cap = kRowsF; p.te = kTileSlots;
  p.n_tiles = (int)std::max<int64_t>(1, ceil_div<int64_t>(p.E, kTileSlots));
  const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
  return pl.prof ? fwd_launch_t<1>(p, pl, grid, st) : fwd_launch_t<0>(p, pl, grid, st);
}

}  // namespace mdl
