// cgconv_fwd_ws.cu -- warp-specialised forward kernel of the fused CGConv edge op: 64 channels [c_off, c_off + 64) of a
// C-wide layer per launch (any C >= 64, C % 4 == 0), G <= 64; edge data = materialised edge_attr rows or, in the
// smearing-fused form, the normalised distance d_hat expanded to the Gaussian basis by the splitter warps.
//
// Reference: PyG CGConv.forward as the reference calls it (matdeeplearn/models/cgcnn.py:80-82,136-145):
//   out_i = x_i + mean_{j->i} sigmoid(W_f z + b_f) * softplus(W_s z + b_s),  z = [x_i | x_j | e_ij].
// Same operator, tile ownership, 3xTF32 contraction, gate math and deterministic per-segment sums as
// k_cgconv_fwd_pipe (cgconv_fwd.cu).  There, the 16 epilogue warps also stage everything a round needs
// (indices, node rows, the hi/lo split of the edge rows into tensor memory) between CTA-wide barriers, and
// that staging is ~45 % of the round.  Here every stage has its own warps and its own mbarrier hand-offs, so
// the epilogue warps do nothing but epilogue + per-segment sums while the next rounds are staged under them:
//
//   warp 20 (one lane)   issuer: bulk (TMA) copy of a round's edge rows; the 21 tcgen05.mma of a round
//   warps 16-19          splitters: thread = slot = TMEM lane; landing zone -> hi / lo -> tcgen05.st (A operand)
//   warps 21-23          loaders: indices + node-row window decision one live round ahead, one bulk copy per P / Q row
//   warps 0-15           consumers: tcgen05.ld -> + P[dst] + Q[src] -> gates -> message tile -> per-segment sums
//
//   edge rows   issuer --bar_ea_full--> splitters --bar_a_full[b]--> issuer (MMA) --bar_mma[b]--> consumers
//               consumers --bar_acc_free[b]--> issuer     splitters wait bar_mma[b] before reusing A buffer b
//   node rows   loaders --bar_rows_full[b]--> consumers --bar_rows_free[b]--> loaders
//
// Measured dead ends (profiles/r2_phase_profile_ws_v3.txt, _v4, _v5; profiles/experiments/cgconv_fwd_ws_v5_reducers.cu):
// node terms preloaded into the accumulator by extra splitter warps, separate reducer warps with double-buffered
// message tiles, 8 splitter warps with 32-column stores -- tensor-memory stores queue behind the running MMAs and the
// extra warps starve each other (1.30 - 2.08 ms against 1.20 ms for this layout).
//
// Exponents are taken in base 2 with the scale folded in up front: the f-gate columns of W_e carry -log2(e), the
// s-gate columns +log2(e), the node terms enter as c * (P + Q) by one packed fma, and the final ln(2) of the softplus
// rides in the per-node scale.  Gate math runs on packed f32x2 (FFMA2 / FADD2 / FMUL2) and the MUFU pipe.
//
// Everything per round is double-buffered (two accumulators, two A-operand buffers, two index / node-row
// buffers), so the producers run up to two rounds ahead.  The node-row window (the P / Q rows a round needs
// lie in two short contiguous node ranges) holds WR rows per buffer; a round whose ranges do not fit reads its
// node terms from global memory (L2) in the epilogue.
#include "cgconv.cuh"
#include "umma.cuh"
#include "edge_dev.cuh"

namespace mdl {

namespace {

constexpr int kCons = 512, kConsWarps = 16;       // consumer (epilogue) threads
constexpr int kSplitWarp0 = 16;                   // warps 16..19: TMEM lane quadrants 0..3
constexpr int kIssuerWarp = 20;
constexpr int kLoadWarp0 = 21, kLoaders = 96;     // warps 21..23
constexpr int kLaunchW = 768;
constexpr int kRowsW = 128, kTileW = 112, kInfoCapW = 512;
constexpr int kC = 64, kNP = 2 * kC;
constexpr int kVW = 2 * kC + 4;                   // row stride of the node-row tiles (bank spread)
constexpr int kVP = kC + 4;                       // row stride of the message tile
constexpr int kTmemColsW = 512;
constexpr int kIdxRing = 4;                     // index / window-record buffers (staged one live round ahead of the node rows)

unsigned long long* g_ws_phase_buf = nullptr;

struct WsPlan {
  unsigned long long* prof;
  int window, KP, WR, sleep_ns, rows_cpasync;
  uint32_t offBhi, offBlo, offEA, offW, wbytes, offV, offIdx, offWin, offInfo, total;
};

// smear: the edge rows are expanded in the kernel from d_hat (no landing zone; its space goes to the node-row windows)
bool ws_plan(int C, int G, bool smear, WsPlan* pl) {
  if (C < kC || (C & 3) || G < 1) return false;  // a launch serves 64 channels [c_off, c_off + 64) of a C-wide layer
  const int KP = (G + 7) & ~7;
  if (2 * kNP + 4 * KP > kTmemColsW) return false;  // two accumulators + two hi/lo A-operand buffers
  const uint32_t b = (uint32_t)kNP * KP * 4;
  const uint32_t ea = smear ? 0u : ((((uint32_t)kRowsW * G * 4 + 32) + 15u) & ~15u);
  const uint32_t v = (uint32_t)kRowsW * kVP * 4, idx = kIdxRing * 2 * kRowsW * 4, win = kIdxRing * 16, info = kInfoCapW * 16;
  const uint32_t fixed = 2 * b + ea + v + idx + win + info;
  if (fixed + 2 * 32 * kVW * 4 > (uint32_t)kMaxDynSmem) return false;
  int WR = (int)(((uint32_t)kMaxDynSmem - fixed) / (2 * kVW * 4)) & ~7;
  if (WR > kRowsW) WR = kRowsW;
  pl->prof = g_ws_phase_buf;
  pl->window = 1; pl->KP = KP; pl->WR = WR; pl->rows_cpasync = 0;
  pl->wbytes = (uint32_t)WR * kVW * 4;
  pl->offBhi = 0; pl->offBlo = b; pl->offEA = 2 * b; pl->offW = pl->offEA + ea;
  pl->offV = pl->offW + 2 * pl->wbytes; pl->offIdx = pl->offV + v; pl->offWin = pl->offIdx + idx;
  pl->offInfo = pl->offWin + win; pl->total = pl->offInfo + info;
  return pl->total <= (uint32_t)kMaxDynSmem;
}

struct RoundW { int k, rd, r_lo, cnt; bool last; };

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}


template <int PROFILE>
__global__ void __launch_bounds__(kLaunchW, 1) k_cgconv_fwd_ws(const CgParams p, const WsPlan pl) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_ea_full, bar_a_full[2], bar_mma[2], bar_acc_free[2], bar_rows_full[2], bar_rows_free[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int sRed[3][2];  // loaders: per-warp (min, max) of the round's source nodes
  __shared__ float sMu[64];   // smearing-fused form: the basis centres (GaussianSmearing.offset), KP <= 64
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = p.G, KP = pl.KP, WR = pl.WR;
  const int C = p.C, c_off = p.c_off;  // layer width (row strides of x / out / PQ / WeT) and this launch's channel chunk
  const uint32_t sleep_ns = (uint32_t)pl.sleep_ns;
#define WAIT(bar, parity) umma::mbar_wait_sleep(bar, parity, sleep_ns)

  uint8_t* sBhi = smem + pl.offBhi;
  uint8_t* sBlo = smem + pl.offBlo;
  float* sEA = reinterpret_cast<float*>(smem + pl.offEA);   // landing zone of a round's edge rows
  float* sV = reinterpret_cast<float*>(smem + pl.offV);     // [128][VP] per-slot messages
  int* sIdx = reinterpret_cast<int*>(smem + pl.offIdx);     // [kIdxRing][src | dst][128]
  int4* sWin = reinterpret_cast<int4*>(smem + pl.offWin);   // [kIdxRing] {window?, src min, dst min, nq}
  TileInfo* sInfo = reinterpret_cast<TileInfo*>(smem + pl.offInfo);
  auto sWbuf = [&](int b) { return reinterpret_cast<float*>(smem + pl.offW + (uint32_t)b * pl.wbytes); };

  const int my_tiles = (p.n_tiles > (int)blockIdx.x) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto make_round = [&](int k, int rd) -> RoundW {
    RoundW R{k, rd, 0, 0, true};
    if (k < my_tiles) {
      const TileInfo T = sInfo[k];
      R.r_lo = T.e_lo + rd * kRowsW;
      R.cnt = max(0, min(T.e_hi - R.r_lo, kRowsW));
      R.last = R.r_lo + kRowsW >= T.e_hi;
    }
    return R;
  };
  auto valid = [&](const RoundW& R) { return R.k < my_tiles; };
  auto next_round = [&](const RoundW& R) -> RoundW { return R.last ? make_round(R.k + 1, 0) : make_round(R.k, R.rd + 1); };

  // ---- one-time setup (all 768 threads): TMEM, barriers, the CTA's whole tile table, resident W_e split hi/lo
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, kTmemColsW);
  if (tid == 32) {
    umma::mbar_init(&bar_ea_full, 1);
    for (int b = 0; b < 2; ++b) {
      umma::mbar_init(&bar_a_full[b], 4);            // one arrival per splitter warp
      umma::mbar_init(&bar_mma[b], 1);               // tcgen05.commit
      umma::mbar_init(&bar_acc_free[b], kConsWarps); // one arrival per consumer warp
      // node rows: one bulk copy per row, loader thread 0 arrives + the rows' bytes; or (MDL_CGCONV_ROWS=cpasync) 16-byte
      // cp.async (warp instruction = one 512-byte row), every loader thread arrives when its copies have landed
      umma::mbar_init(&bar_rows_full[b], pl.rows_cpasync ? kLoaders : 1);
      umma::mbar_init(&bar_rows_free[b], kConsWarps);
    }
    umma::fence_mbar_init();
  }
  for (int k = tid; k < my_tiles; k += kLaunchW) {   // my_tiles <= kInfoCapW (checked by the host)
    TileInfo t;
    const int tile = blockIdx.x + k * gridDim.x;
    t.n_lo = first_segment_at_or_after<CG_FWD>(p, tile * kTileW);
    t.n_hi = (tile == p.n_tiles - 1) ? p.N : first_segment_at_or_after<CG_FWD>(p, (tile + 1) * kTileW);
    if (t.n_hi < t.n_lo) t.n_hi = t.n_lo;
    t.e_lo = __ldg(p.seg_ptr + t.n_lo);
    t.e_hi = __ldg(p.seg_ptr + t.n_hi);
    sInfo[k] = t;
  }
  for (int i = tid; i < kNP * KP; i += kLaunchW) {
    const int n = i % kNP, k = i / kNP;
    const float w = (k < G) ? __ldg(p.WeT + (size_t)k * (2 * C) + (n < kC ? c_off + n : C + c_off + n - kC)) * (n < kC ? -kLog2e : kLog2e) : 0.0f;
    const float hi = umma::tf32_hi(w);
    const int off = umma::tile_offset_bytes(n, k, kNP);
    *reinterpret_cast<float*>(sBhi + off) = hi;
    *reinterpret_cast<float*>(sBlo + off) = w - hi;
  }
  if (p.dhat && tid < 64) sMu[tid] = __ldg(p.sm_offset + min(tid, G - 1));
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  auto tm_acc = [&](int b) { return tmem + (uint32_t)b * kNP; };
  auto tm_a_hi = [&](int b) { return tmem + 2 * kNP + (uint32_t)b * 2 * KP; };
  auto tm_a_lo = [&](int b) { return tmem + 2 * kNP + (uint32_t)b * 2 * KP + (uint32_t)KP; };

  // ---- edge rows of a round: one bulk copy from the 16-byte boundary below the block (see cgconv_tc.cu)
  auto ea_bulk_bytes = [&](int r_lo, int cnt) -> uint32_t {
    if (cnt <= 0 || p.dhat) return 0u;  // smearing-fused form: nothing to copy
    const long long first = (long long)r_lo * G;
    const uint32_t bytes = (uint32_t)(((int)(first & 3) + cnt * G) * 4);
    const bool more = ((long long)p.E * G - (first + (long long)cnt * G)) >= 3;
    return more ? ((bytes + 15u) & ~15u) : (bytes & ~15u);
  };

  // =====================================================================================================
  if (warp >= kIssuerWarp) {
    if (warp == kIssuerWarp) {
      // ---------------- issuer: bulk copies of the edge rows, MMAs
      if (lane == 0) {
        auto issue_ea_bulk = [&](const RoundW& R) {
          const uint32_t nb = ea_bulk_bytes(R.r_lo, R.cnt);
          if (!nb) return;
          const long long first = (long long)R.r_lo * G;
          umma::mbar_arrive_expect_tx(&bar_ea_full, nb);
          umma::bulk_g2s(sEA, p.ea + (first - (first & 3)), nb, &bar_ea_full);
        };
        const uint32_t idesc = umma::make_idesc_tf32(kRowsW, kNP);
        const uint32_t step_b = 2 * (uint32_t)kNP * 16;
        const uint32_t b_hi = umma::smem_u32(sBhi), b_lo = umma::smem_u32(sBlo);
        uint32_t ph_a = 0, ph_f = 0;  // phase parities, bit b = buffer b (plain registers: no dynamically indexed arrays)
        uint32_t busy = 0;            // bit b: accumulator b holds a round whose epilogue has not been waited for
        RoundW R = make_round(0, 0);
        if (valid(R)) issue_ea_bulk(R);
        for (uint32_t it = 0; valid(R); ++it) {
          const int b = it & 1;
          const RoundW Rn = next_round(R);
          if (R.cnt > 0) {  // split of this round done: its A operand is staged and the landing zone is free
            WAIT(&bar_a_full[b], (ph_a >> b) & 1);
            ph_a ^= 1u << b;
          }
          if (valid(Rn)) issue_ea_bulk(Rn);  // first: the MMA issue below blocks for ~2k cycles
          if (R.cnt > 0) {
            if ((busy >> b) & 1) {  // the epilogue of the round that used this accumulator two rounds ago has read it
              WAIT(&bar_acc_free[b], (ph_f >> b) & 1);
              ph_f ^= 1u << b;
            }
            umma::fence_after_sync();
            uint32_t acc = 0;
#pragma unroll 1
            for (int pass = 0; pass < 3; ++pass) {
              const uint32_t a = (pass == 2) ? tm_a_lo(b) : tm_a_hi(b);
              const uint32_t bb = (pass == 1) ? b_lo : b_hi;
              for (int kk = 0; kk < (KP >> 3); ++kk) {
                umma::mma_tf32_ts(tm_acc(b), a + kk * 8, umma::make_desc(bb + kk * step_b, (uint32_t)kNP * 16, 128), idesc, acc);
                acc = 1;
              }
            }
            umma::mma_commit(&bar_mma[b]);
            busy |= 1u << b;
          }
          R = Rn;
        }
      }
      __syncwarp();
    } else {
      // ---------------- loaders: indices, window decision, node rows of a round.  Live rounds (cnt > 0) are counted
      // by j on both sides: node rows in buffer j & 1, indices + window record in ring slot j & 3.  The indices and
      // the window decision of live round j + 1 are staged while round j's rows are in flight, so what follows the
      // consumers' release of a row buffer is only the bulk copies themselves.
      const int lt = tid - kLoadWarp0 * 32, lw = warp - kLoadWarp0;
      auto sync_loaders = [] { asm volatile("bar.sync 3, %0;" ::"n"(kLoaders) : "memory"); };
      struct Dec { int win, s_lo, d_lo, nq, nrows; };
      auto next_live = [&](RoundW R) { while (valid(R) && R.cnt == 0) R = next_round(R); return R; };
      auto load_idx = [&](const RoundW& R, int (&v)[3]) {  // this thread's (up to) three indices: global loads in flight
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int i = lt + kLoaders * k, e = i & (kRowsW - 1);
          v[k] = (i < 2 * kRowsW && e < R.cnt) ? __ldg((i < kRowsW ? p.dst_src : p.dst_dst) + R.r_lo + e) : 0;
        }
      };
      auto finish_idx = [&](const RoundW& R, int slot, const int (&v)[3]) -> Dec {
        int* bS = sIdx + slot * 2 * kRowsW;
        int s_lo = 0x7fffffff, s_hi = -1;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int i = lt + kLoaders * k, e = i & (kRowsW - 1);
          if (i < 2 * kRowsW) {
            if (i < kRowsW && e < R.cnt) { s_lo = min(s_lo, v[k]); s_hi = max(s_hi, v[k]); }
            bS[i] = v[k];
          }
        }
        s_lo = __reduce_min_sync(0xffffffffu, s_lo);
        s_hi = __reduce_max_sync(0xffffffffu, s_hi);
        if (lane == 0) { sRed[lw][0] = s_lo; sRed[lw][1] = s_hi; }
        sync_loaders();  // indices and per-warp ranges visible to all loaders
        s_lo = min(sRed[0][0], min(sRed[1][0], sRed[2][0]));
        s_hi = max(sRed[0][1], max(sRed[1][1], sRed[2][1]));
        const int d_lo = bS[kRowsW], d_hi = bS[kRowsW + R.cnt - 1];  // slots are sorted by destination
        const int nq = s_hi - s_lo + 1, np_ = d_hi - d_lo + 1;
        const bool win = pl.window && nq + np_ <= WR;
        if (lt == 0) sWin[slot] = make_int4(win ? 1 : 0, s_lo, d_lo, nq);
        sync_loaders();  // sRed is rewritten by the next call
        return Dec{win ? 1 : 0, s_lo, d_lo, nq, win ? nq + np_ : 0};
      };
      uint32_t ph_rf = 0;
      RoundW R = next_live(make_round(0, 0));
      Dec D{0, 0, 0, 0, 0};
      int v[3];
      if (valid(R)) {
        load_idx(R, v);
        D = finish_idx(R, 0, v);
      }
      for (uint32_t j = 0; valid(R); ++j) {
        const int b = j & 1;
        const RoundW Rn = next_live(next_round(R));
        if (valid(Rn)) load_idx(Rn, v);
        if (j >= 2) {  // consumers have finished with this row buffer (live round j - 2)
          WAIT(&bar_rows_free[b], (ph_rf >> b) & 1);
          ph_rf ^= 1u << b;
        }
        float* W = sWbuf(b);
        if (pl.rows_cpasync) {
          // warp per row, lane = 16-byte chunk: lanes 0-15 the chunk's f piece, 16-31 its s piece (C floats further on in
          // a PQ row [P_f | P_s | Q_f | Q_s]).  ~14 LDGSTS.128 per warp and round instead of ~14 serialised bulk-copy
          // issues (a UBLKCP occupies its thread for a few hundred cycles)
          for (int r = lw; r < D.nrows; r += kLoaders / 32) {
            const float* g = (r < D.nq) ? p.PQ + (size_t)(D.s_lo + r) * (4 * C) + 2 * C + c_off
                                        : p.PQ + (size_t)(D.d_lo + r - D.nq) * (4 * C) + c_off;
            cp_async16(W + r * kVW + 4 * lane, g + (lane < 16 ? 4 * lane : C + 4 * (lane - 16)));
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(umma::smem_u32(&bar_rows_full[b])) : "memory");
        } else if (lt == 0) {
          if (D.nrows) umma::mbar_arrive_expect_tx(&bar_rows_full[b], (uint32_t)D.nrows * (uint32_t)(2 * kC * 4));
          else mbar_arrive(&bar_rows_full[b]);
        }
        for (int r = lt; !pl.rows_cpasync && r < D.nrows; r += kLoaders) {  // rows [0,nq) = Q[smin..smax], rows [nq,nq+np) = P[dmin..dmax]
          // a PQ row is [P_f | P_s | Q_f | Q_s], C floats each: the chunk's f and s pieces are adjacent only when C = 64
          const float* g = (r < D.nq) ? p.PQ + (size_t)(D.s_lo + r) * (4 * C) + 2 * C + c_off
                                      : p.PQ + (size_t)(D.d_lo + r - D.nq) * (4 * C) + c_off;
          if (C == kC) {
            umma::bulk_g2s(W + r * kVW, g, (uint32_t)(2 * kC * 4), &bar_rows_full[b]);
          } else {
            umma::bulk_g2s(W + r * kVW, g, (uint32_t)(kC * 4), &bar_rows_full[b]);
            umma::bulk_g2s(W + r * kVW + kC, g + C, (uint32_t)(kC * 4), &bar_rows_full[b]);
          }
        }
        if (valid(Rn)) D = finish_idx(Rn, (int)((j + 1) & (kIdxRing - 1)), v);
        R = Rn;
      }
    }
    __syncthreads();  // teardown barrier of the CTA
    return;
  }
  if (warp >= kSplitWarp0) {
    // ---------------- splitters: thread = slot = TMEM lane; the whole edge row -> hi / lo -> tensor memory
    const int e = tid - kSplitWarp0 * 32;
    uint32_t ph_ea = 0, ph_m = 0, used = 0;
    const SmearConst sc = p.dhat ? smear_const(sMu, G, p.sm_coeff) : SmearConst{0.0f, 0.0f, 0.0f};
    RoundW R = make_round(0, 0);
    for (uint32_t it = 0; valid(R); ++it) {
      const int b = it & 1;
      if (R.cnt > 0) {
        // smearing-fused form: this slot's normalised distance (requested before the waits below)
        const float dh = (p.dhat && e < R.cnt) ? __ldg(p.dhat + R.r_lo + e) : 0.0f;
        const uint32_t nb = ea_bulk_bytes(R.r_lo, R.cnt);
        if (nb) {
          WAIT(&bar_ea_full, ph_ea);
          ph_ea ^= 1;
        }
        if ((used >> b) & 1) {  // the MMAs that read this A buffer two rounds ago have retired
          WAIT(&bar_mma[b], (ph_m >> b) & 1);
          ph_m ^= 1u << b;
          umma::fence_after_sync();
        }
        const int ea_off = (int)(((long long)R.r_lo * G) & 3);
        const float* row = sEA + ea_off + e * G;
        const int landed = (int)(nb >> 2);  // first float of the landing zone the bulk copy did NOT deliver
        const bool patch = e < R.cnt && landed < ea_off + (e + 1) * G;
        const uint32_t a_hi = umma::tmem_addr(tm_a_hi(b), warp, 0), a_lo = umma::tmem_addr(tm_a_lo(b), warp, 0);
        for (int ch = 0; ch < (KP >> 3); ++ch) {
          float v[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] = 0.0f;
          if (p.dhat) {  // Gaussian basis of this slot, columns 8 ch .. 8 ch + 7 (reference process.py:580-590)
            if (e < R.cnt) smear_chunk8(dh, sMu, 8 * ch, G, sc, v);
          } else if (e < R.cnt) {
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
            if (patch) {
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const int k = 8 * ch + t;
                if (k < G && ea_off + e * G + k >= landed) v[t] = __ldg(p.ea + ((long long)R.r_lo + e) * G + k);
              }
            }
          }
          float hi[8], lo[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) { hi[t] = umma::tf32_hi(v[t]); lo[t] = v[t] - hi[t]; }
          umma::tmem_st8(a_hi + 8 * ch, hi);
          umma::tmem_st8(a_lo + 8 * ch, lo);
        }
        umma::tmem_st_wait();
        umma::fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_a_full[b]);
        used |= 1u << b;
      }
      R = next_round(R);
    }
    __syncthreads();  // teardown barrier of the CTA
    return;
  }

  // ---------------- consumers: epilogue + per-segment sums
  auto sync_consumers = [] { asm volatile("bar.sync 2, %0;" ::"n"(kCons) : "memory"); };
  long long t_prev = PROFILE ? clock64() : 0;
  auto mark = [&](int slot) {
    if (PROFILE && pl.prof && tid == 0) {
      const long long now = clock64();
      atomicAdd(pl.prof + slot, (unsigned long long)(now - t_prev));
      t_prev = now;
    }
  };
  const int q = warp & 3, part = warp >> 2;  // TMEM lane quadrant, channel quarter (16 channels)
  const int c_begin = part * 16;
  uint32_t ph_r = 0, ph_m = 0;
  uint32_t live = 0;  // live rounds (cnt > 0) so far: node rows in buffer live & 1, indices / window record in ring slot live & 3
  RoundW cur = make_round(0, 0);
  for (uint32_t it = 0; valid(cur); ++it) {
    const int b = it & 1;                                        // accumulator / A-operand buffer (issuer, splitters)
    const int jb = live & 1, js = live & (kIdxRing - 1);         // node-row buffer, index ring slot (loaders)
    const int cnt = cur.cnt, r_lo = cur.r_lo, r_hi = cur.r_lo + cur.cnt;
    const int n_lo = sInfo[cur.k].n_lo, n_hi = sInfo[cur.k].n_hi;
    const int* bSrc = sIdx + js * 2 * kRowsW;
    const int* bDst = bSrc + kRowsW;
    mark(0);
    // ---- reduce-stage node data of this warp's first segment (global loads in flight across the epilogue)
    const int n0 = n_lo + warp;
    int seg_a = 0, seg_b = 0;
    float seg_inv = 1.0f;                    // 1/deg, multiplied by the softplus' ln 2 only where it is used (in the
                                             // reduce stage): a multiply here would wait for the load on the spot
    float2 seg_x = make_float2(0.0f, 0.0f);  // lane l owns channels 2l, 2l+1
    if (n0 < n_hi) {
      seg_a = __ldg(p.seg_ptr + n0);
      seg_b = __ldg(p.seg_ptr + n0 + 1);
      if (p.inv_deg) seg_inv = __ldg(p.inv_deg + n0);
      seg_x = __ldg(reinterpret_cast<const float2*>(p.x + (size_t)n0 * C + c_off) + lane);
    }
    if (cnt > 0) {
      WAIT(&bar_rows_full[jb], (ph_r >> jb) & 1);  // indices, window record, node rows of this round
      ph_r ^= 1u << jb;
      mark(1);
      const int4 wr = sWin[js];
      const bool win = wr.x != 0;
      const int w_smin = wr.y, w_dmin = wr.z, w_nq = wr.w;
      const float* sW = sWbuf(jb);
      WAIT(&bar_mma[b], (ph_m >> b) & 1);         // this round's contraction
      ph_m ^= 1u << b;
      umma::fence_after_sync();
      mark(2);
      float f[16], sacc[16];
      umma::tmem_ld16(umma::tmem_addr(tm_acc(b), q, c_begin), f);
      umma::tmem_ld16(umma::tmem_addr(tm_acc(b), q, kC + c_begin), sacc);
      umma::tmem_ld_wait();
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acc_free[b]);  // the accumulator may be overwritten (round it + 2)
      mark(3);
      // ---- epilogue: thread = slot (TMEM lane), 16 channels; a = accumulator + P[dst] + Q[src], gates,
      // message parked in the message tile
      const int e_ep = 32 * q + lane;
      if (e_ep < cnt) {
        const int sd = bDst[e_ep], ss = bSrc[e_ep];
        float* rowv = sV + e_ep * kVP + c_begin;
        const f2_t cf = pk2(-kLog2e, -kLog2e), cs = pk2(kLog2e, kLog2e);
        auto run = [&](auto ldP, auto ldQ) {  // how the P[dst] / Q[src] rows are read (shared or global memory), by float offset
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {
            const float4 pf = ldP(0, j4), ps = ldP(1, j4);   // (gate f / s, channel offset)
            const float4 qf = ldQ(0, j4), qs = ldQ(1, j4);
            // y = accumulator (already in base-2 units: W_e is pre-scaled) + c (P + Q)
            float yf0, yf1, yf2, yf3, ys0, ys1, ys2, ys3;
            upk2(fma2(cf, add2(pk2(pf.x, pf.y), pk2(qf.x, qf.y)), pk2(f[j4], f[j4 + 1])), yf0, yf1);
            upk2(fma2(cf, add2(pk2(pf.z, pf.w), pk2(qf.z, qf.w)), pk2(f[j4 + 2], f[j4 + 3])), yf2, yf3);
            upk2(fma2(cs, add2(pk2(ps.x, ps.y), pk2(qs.x, qs.y)), pk2(sacc[j4], sacc[j4 + 1])), ys0, ys1);
            upk2(fma2(cs, add2(pk2(ps.z, ps.w), pk2(qs.z, qs.w)), pk2(sacc[j4 + 2], sacc[j4 + 3])), ys2, ys3);
            float m0, m1, m2, m3;
            upk2(gate_pair(yf0, yf1, ys0, ys1), m0, m1);
            upk2(gate_pair(yf2, yf3, ys2, ys3), m2, m3);
            *reinterpret_cast<float4*>(rowv + j4) = make_float4(m0, m1, m2, m3);
          }
        };
        if (win) {  // explicit shared-space loads (a pointer that may be either space compiles to generic LD.E)
          const uint32_t a0 = umma::smem_u32(sW + (w_nq + sd - w_dmin) * kVW + c_begin);
          const uint32_t a1 = umma::smem_u32(sW + (ss - w_smin) * kVW + c_begin);
          run([&](int gate, int o) { return umma::lds128(a0 + 4u * (uint32_t)(gate * kC + o)); },
              [&](int gate, int o) { return umma::lds128(a1 + 4u * (uint32_t)(gate * kC + o)); });
        } else {
          const float* r0 = p.PQ + (size_t)sd * (4 * C) + c_off + c_begin;
          const float* r1 = p.PQ + (size_t)ss * (4 * C) + 2 * C + c_off + c_begin;
          run([&](int gate, int o) { return __ldg(reinterpret_cast<const float4*>(r0 + gate * C + o)); },
              [&](int gate, int o) { return __ldg(reinterpret_cast<const float4*>(r1 + gate * C + o)); });
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_rows_free[jb]);  // indices / node rows of this buffer are no longer needed
      ++live;
    }
    mark(4);
    sync_consumers();  // [S3] message tile complete
    mark(5);
    // ---- segmented sum over the owned segments that have slots in this round (slot order: deterministic)
    for (int n = n0; n < n_hi; n += kConsWarps) {
      int a, bq;
      if (n == n0) { a = seg_a; bq = seg_b; }
      else { a = __ldg(p.seg_ptr + n); bq = __ldg(p.seg_ptr + n + 1); }
      const int lo = max(a, r_lo), hi = min(bq, r_hi);
      const bool empty_seg = (a == bq);
      if (empty_seg ? (cur.rd != 0) : (lo >= hi)) continue;
      const bool first = empty_seg || (a >= r_lo);
      const bool lastp = empty_seg || (bq <= r_hi);
      float2* o = reinterpret_cast<float2*>(p.out + (size_t)n * C + c_off) + lane;
      float sc = kLn2 * seg_inv;
      float2 x = seg_x;
      if (n != n0) {
        sc = p.inv_deg ? kLn2 * __ldg(p.inv_deg + n) : kLn2;
        x = __ldg(reinterpret_cast<const float2*>(p.x + (size_t)n * C + c_off) + lane);
      }
      float2 acc = first ? make_float2(0.0f, 0.0f) : *o;
      const float2* vp = reinterpret_cast<const float2*>(sV + (lo - r_lo) * kVP) + lane;
      int s = lo;
      for (; s + 4 <= hi; s += 4, vp += 4 * (kVP / 2)) {  // four loads in flight, added in slot order
        const float2 v0 = vp[0], v1 = vp[kVP / 2], v2 = vp[2 * (kVP / 2)], v3 = vp[3 * (kVP / 2)];
        acc.x += v0.x; acc.y += v0.y;
        acc.x += v1.x; acc.y += v1.y;
        acc.x += v2.x; acc.y += v2.y;
        acc.x += v3.x; acc.y += v3.y;
      }
      for (; s < hi; ++s, vp += kVP / 2) {
        const float2 v = *vp;
        acc.x += v.x; acc.y += v.y;
      }
      *o = lastp ? make_float2(fmaf(acc.x, sc, x.x), fmaf(acc.y, sc, x.y)) : acc;
    }
    mark(6);
    sync_consumers();  // [S1] message tile free for the next round's epilogue
    mark(7);
    if (PROFILE && pl.prof && tid == 0) atomicAdd(pl.prof + 31, 1ull);
    cur = next_round(cur);
  }
  umma::fence_before_sync();
  __syncthreads();  // teardown barrier of the CTA
  if (warp == 0) umma::tmem_dealloc(tmem, kTmemColsW);
}

#undef WAIT

template <int PROFILE>
int ws_launch_t(const CgParams& p, const WsPlan& pl, int grid, cudaStream_t st) {
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MDL_CUDA(cudaFuncSetAttribute(k_cgconv_fwd_ws<PROFILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    configured.store(1, std::memory_order_release);
  }
  k_cgconv_fwd_ws<PROFILE><<<grid, kLaunchW, pl.total, st>>>(p, pl);
  MDL_LAUNCHED();
  return MDL_OK;
}

}  // namespace

void cgws_set_phase_buffer(unsigned long long* dev_ptr) { g_ws_phase_buf = dev_ptr; }

// C >= 64 (multiple of 4; 64 channels per launch), G <= 64, a tile table that fits (<= 512 tiles per CTA), 16-byte
// aligned ea / PQ (bulk copies)
bool cgws_supported(const CgParams& p) {
  WsPlan pl;
  const int64_t n_tiles = std::max<int64_t>(1, ceil_div<int64_t>(p.E, kTileW));
  const int64_t grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  return ws_plan(p.C, p.G, p.dhat != nullptr, &pl) && ceil_div<int64_t>(n_tiles, grid) <= kInfoCapW &&
         (reinterpret_cast<uintptr_t>(p.ea) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.PQ) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.x) & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 7) == 0 &&
         (int64_t)p.N * 4 * p.C < (int64_t)1 << 31;
}

int cgws_launch(CgParams p, cudaStream_t st) {
  WsPlan pl;
  MDL_REQUIRE(ws_plan(p.C, p.G, p.dhat != nullptr, &pl), "cgconv_fwd_ws: unsupported shape C=%d G=%d", p.C, p.G);
  const char* wenv = getenv("MDL_CGCONV_WINDOW");  // "0": node terms from global memory only (A/B and test switch)
  pl.window = !(wenv && wenv[0] == '0');
  const char* senv = getenv("MDL_WS_SLEEP");  // ns slept between mbarrier polls (A/B switch)
  pl.sleep_ns = senv ? atoi(senv) : 0;
  // node rows: one bulk (TMA) copy per row (default) or, MDL_CGCONV_ROWS=cpasync, 16-byte cp.async with one warp
  // instruction per row -- measured equal (1.144 vs 1.156 ms on the 16384-graph workload: the consumers' wait for the
  // rows is not set by the loaders' issue cost), kept as a switch
  const char* renv = getenv("MDL_CGCONV_ROWS");
  pl.rows_cpasync = (renv && strcmp(renv, "cpasync") == 0) ? 1 : 0;
  p.CC = kC; p.cap = kRowsW; p.te = kTileW;
  p.n_tiles = (int)std::max<int64_t>(1, ceil_div<int64_t>(p.E, kTileW));
  const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
  // one launch per 64-channel chunk; the last chunk of a width that is not a multiple of 64 starts at C - 64 and
  // recomputes the channels it shares with the one before (identical values, plain stores)
  for (int c0 = 0; c0 < p.C; c0 += kC) {
    p.c_off = std::min(c0, p.C - kC);
    if (int rc = pl.prof ? ws_launch_t<1>(p, pl, grid, st) : ws_launch_t<0>(p, pl, grid, st)) return rc;
  }
  return MDL_OK;
}

}  // namespace mdl
