// cgconv_fwd_ws.cu -- warp-specialised forward kernel of the fused CGConv edge op (C = 64, G <= 64).
//
// Reference: PyG CGConv.forward as the reference calls it (matdeeplearn/models/cgcnn.py:80-82,136-145):
//   out_i = x_i + mean_{j->i} sigmoid(W_f z + b_f) * softplus(W_s z + b_s),  z = [x_i | x_j | e_ij].
// Same operator, tile ownership, 3xTF32 contraction and deterministic per-segment sums as k_cgconv_fwd_pipe
// (cgconv_fwd.cu).  There, the 16 epilogue warps also stage everything a round needs and sum the segments,
// phase after phase between CTA-wide barriers.  Here every stage has its own warps and mbarrier hand-offs;
// there is no CTA-wide barrier inside the loop and the gate warps do nothing but gate math:
//
//   warps 25-27  loaders    the round's indices, its node-row window decision, one bulk (TMA) copy per P / Q row
//   warps 16-23  splitters  thread = slot = TMEM lane: edge row (landing zone) -> hi (warps 16-19) / lo (warps 20-23)
//                           -> tcgen05.st, 32 columns per instruction (a tensor-memory store costs its warp ~600
//                           cycles whatever its width, so few wide stores on many warps)
//   warp 24      issuer     bulk (TMA) copy of a round's edge rows; the 21 tcgen05.mma (3xTF32) of a round
//   warps 0-15   gates      tcgen05.ld -> + c (P[dst] + Q[src]) -> sigmoid * softplus on packed f32x2 / MUFU -> message tile
//   warps 28-31  reducers   per-destination sums of a round's message tile in slot order -> out (+ x, * 1/deg)
//
//   issuer --ea_full--> splitters --a_full[b]--> issuer (MMA) --mma[b]--> gates --v_full[b]--> reducers --v_free[b]--> gates
//   loaders --rows_full[b]--> gates --rows_free[b]--> loaders      gates --acc_free[b]--> issuer
//   mma[b] also frees A buffer b for the splitters
//
// (Measured dead end, profiles/r2_phase_profile_ws_v3.txt: preloading the node terms into the accumulator with
// tcgen05.st from extra splitter warps -- tensor-memory stores queue behind the running MMAs, the splitters became the
// bottleneck at ~5k cycles per round.  The node terms are added by the gate warps from shared memory.)
//
// Every per-round resource is double-buffered (accumulators, A-operand columns, index / node-row buffers, message
// tiles); only the edge-row landing zone is single (its copy for round r+1 is issued as soon as round r is split).
// Exponents are taken in base 2 with the scale folded in up front: the f-gate columns of W_e and of the node terms
// carry -log2(e), the s-gate columns +log2(e), and the final ln(2) of the softplus rides in the per-node scale.
// The node-row window (the P / Q rows a round needs lie in two short contiguous node ranges) holds WR rows per
// buffer; a round whose ranges do not fit reads its node rows from global memory (L2) in the splitters.
#include "cgconv.cuh"
#include "umma.cuh"
#include "edge_dev.cuh"

namespace mdl {

namespace {

constexpr int kGateWarps = 16;                    // warps 0..15
constexpr int kSplitWarp0 = 16, kSplitWarps = 8;  // warps 16..19: hi halves, 20..23: lo halves (TMEM lane quadrants 0..3)
constexpr int kIssuerWarp = 24;
constexpr int kLoadWarp0 = 25, kLoaders = 96;     // warps 25..27
constexpr int kRedWarp0 = 28, kRedWarps = 4;      // warps 28..31
constexpr int kLaunchW = 1024;
constexpr int kAW = 64;                           // columns of one A-operand half (hi or lo) in tensor memory
constexpr int kRowsW = 128, kTileW = 112, kInfoCapW = 512;
constexpr int kC = 64, kNP = 2 * kC;
constexpr int kVW = 2 * kC + 4;                   // row stride of the node-row tiles (bank spread)
constexpr int kVP = kC + 4;                       // row stride of the message tiles
constexpr int kTmemColsW = 512;

unsigned long long* g_ws_phase_buf = nullptr;

struct WsPlan {
  unsigned long long* prof;
  int window, KP, WR;
  uint32_t offBhi, offBlo, offEA, offW, wbytes, offV, vbytes, offIdx, offWin, offInfo, total;
};

bool ws_plan(int C, int G, WsPlan* pl) {
  if (C != kC || G < 1) return false;
  const int KP = (G + 7) & ~7;
  if (KP > kAW) return false;  // tensor memory: two accumulators (2 x 128) + two hi / lo A-operand buffers (4 x 64)
  const uint32_t b = (uint32_t)kNP * KP * 4;
  const uint32_t ea = (((uint32_t)kRowsW * G * 4 + 32) + 15u) & ~15u;
  const uint32_t v = (uint32_t)kRowsW * kVP * 4, idx = 2 * 2 * kRowsW * 4, win = 64, info = kInfoCapW * 16;
  const uint32_t fixed = 2 * b + ea + 2 * v + idx + win + info;
  if (fixed + 2 * 32 * kVW * 4 > (uint32_t)kMaxDynSmem) return false;
  int WR = (int)(((uint32_t)kMaxDynSmem - fixed) / (2 * kVW * 4)) & ~7;
  if (WR > kRowsW) WR = kRowsW;
  pl->prof = g_ws_phase_buf;
  pl->window = 1; pl->KP = KP; pl->WR = WR;
  pl->wbytes = (uint32_t)WR * kVW * 4; pl->vbytes = v;
  pl->offBhi = 0; pl->offBlo = b; pl->offEA = 2 * b; pl->offW = pl->offEA + ea;
  pl->offV = pl->offW + 2 * pl->wbytes; pl->offIdx = pl->offV + 2 * v; pl->offWin = pl->offIdx + idx;
  pl->offInfo = pl->offWin + win; pl->total = pl->offInfo + info;
  return pl->total <= (uint32_t)kMaxDynSmem;
}

struct RoundW { int k, rd, r_lo, cnt; bool last; };

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

// ---- packed fp32 pairs (one FMA-pipe instruction per two values on sm_100)
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t pk2(float a, float b) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) {
  f2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) {
  f2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) {
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// sigmoid(a_f) * softplus(a_s) / ln 2 for two channels, from yf = -log2(e) a_f and ys = +log2(e) a_s:
//   1 / (1 + 2^yf)  *  (max(ys, 0) + log2(1 + 2^-|ys|))
// three MUFU ops per channel (ex2, ex2, lg2); the reciprocal runs on the FMA pipe as in rcp_fma (bit-trick seed,
// two third-order steps), on packed pairs.
__device__ __forceinline__ f2_t gate_pair(float yf0, float yf1, float ys0, float ys1) {
  const f2_t one = pk2(1.0f, 1.0f);
  const f2_t u = add2(pk2(ex2_(fminf(yf0, 126.0f)), ex2_(fminf(yf1, 126.0f))), one);
  float u0, u1;
  upk2(u, u0, u1);
  f2_t r = pk2(__uint_as_float(0x7EF311C7u - __float_as_uint(u0)), __uint_as_float(0x7EF311C7u - __float_as_uint(u1)));
  const f2_t nu = pk2(-u0, -u1);
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const f2_t e = fma2(nu, r, one);
    r = fma2(r, fma2(e, e, e), r);
  }
  const f2_t w = add2(pk2(ex2_(-fabsf(ys0)), ex2_(-fabsf(ys1))), one);
  float w0, w1;
  upk2(w, w0, w1);
  const f2_t sp = add2(pk2(lg2_(w0), lg2_(w1)), pk2(fmaxf(ys0, 0.0f), fmaxf(ys1, 0.0f)));
  return mul2(r, sp);
}

template <int PROFILE>
__global__ void __launch_bounds__(kLaunchW, 1) k_cgconv_fwd_ws(const CgParams p, const WsPlan pl) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_ea_full, bar_a_full[2], bar_mma[2], bar_acc_free[2], bar_rows_full[2], bar_rows_free[2],
      bar_v_full[2], bar_v_free[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int sRed[3][2];  // loaders: per-warp (min, max) of the round's source nodes
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = p.G, KP = pl.KP, WR = pl.WR;

  uint8_t* sBhi = smem + pl.offBhi;
  uint8_t* sBlo = smem + pl.offBlo;
  float* sEA = reinterpret_cast<float*>(smem + pl.offEA);   // landing zone of a round's edge rows
  int* sIdx = reinterpret_cast<int*>(smem + pl.offIdx);     // [2 buffers][src | dst][128]
  int4* sWin = reinterpret_cast<int4*>(smem + pl.offWin);   // [2 buffers] {window?, src min, dst min, nq}
  TileInfo* sInfo = reinterpret_cast<TileInfo*>(smem + pl.offInfo);
  auto sWbuf = [&](int b) { return reinterpret_cast<float*>(smem + pl.offW + (uint32_t)b * pl.wbytes); };  // node rows
  auto sVbuf = [&](int b) { return reinterpret_cast<float*>(smem + pl.offV + (uint32_t)b * pl.vbytes); };  // [128][VP] messages

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

  // per-role cycle accounting (instrumented build): one thread of a role adds the cycles between its marks
  long long t_prev = PROFILE ? clock64() : 0;
  auto mark = [&](int slot) {
    if (PROFILE && pl.prof) {
      const long long now = clock64();
      atomicAdd(pl.prof + slot, (unsigned long long)(now - t_prev));
      t_prev = now;
    }
  };

  // ---- one-time setup (all threads): TMEM, barriers, the CTA's whole tile table, resident W_e split hi/lo.
  // The exponent scale rides in the weights: f-gate columns x -log2(e), s-gate columns x +log2(e).
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, kTmemColsW);
  if (tid == 32) {
    umma::mbar_init(&bar_ea_full, 1);
    for (int b = 0; b < 2; ++b) {
      umma::mbar_init(&bar_a_full[b], kSplitWarps);    // splitter warps: A operand staged (and the landing zone read)
      umma::mbar_init(&bar_mma[b], 1);                 // tcgen05.commit
      umma::mbar_init(&bar_acc_free[b], kGateWarps);   // gate warps: accumulator read
      umma::mbar_init(&bar_rows_full[b], 1);           // loader thread 0 (+ the rows' bytes)
      umma::mbar_init(&bar_rows_free[b], kGateWarps);  // gate warps: indices / node rows read
      umma::mbar_init(&bar_v_full[b], kGateWarps);     // gate warps: message tile written
      umma::mbar_init(&bar_v_free[b], kRedWarps);      // reducer warps: message tile summed
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
    const float w = (k < G) ? __ldg(p.WeT + (size_t)k * kNP + n) * (n < kC ? -kLog2e : kLog2e) : 0.0f;
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
  auto tm_acc = [&](int b) { return tmem + (uint32_t)b * kNP; };
  auto tm_a_hi = [&](int b) { return tmem + 2 * kNP + (uint32_t)b * 2 * kAW; };
  auto tm_a_lo = [&](int b) { return tmem + 2 * kNP + (uint32_t)b * 2 * kAW + (uint32_t)kAW; };

  // ---- edge rows of a round: one bulk copy from the 16-byte boundary below the block (see cgconv_tc.cu)
  auto ea_bulk_bytes = [&](int r_lo, int cnt) -> uint32_t {
    if (cnt <= 0) return 0u;
    const long long first = (long long)r_lo * G;
    const uint32_t bytes = (uint32_t)(((int)(first & 3) + cnt * G) * 4);
    const bool more = ((long long)p.E * G - (first + (long long)cnt * G)) >= 3;
    return more ? ((bytes + 15u) & ~15u) : (bytes & ~15u);
  };

  // =====================================================================================================
  if (warp >= kRedWarp0) {
    // ---------------- reducers: per-destination sums of a round's message tile (slot order: deterministic).
    // The node data (segment bounds, 1/deg, x row) of a warp's first kPre segments of a round are requested one
    // whole round ahead: they stream from HBM, and a segment's sum is far shorter than that latency.
    const int rw = warp - kRedWarp0;
    const bool prof_me = (tid == kRedWarp0 * 32);
    uint32_t ph_v = 0;
    constexpr int kPre = 4;
    struct Seg { int a, b; float sc; float2 x; };
    auto load_seg = [&](int n, int n_hi) -> Seg {  // node data of segment n (lane l owns channels 2l, 2l+1)
      Seg s{0, 0, kLn2, make_float2(0.0f, 0.0f)};
      if (n < n_hi) {
        s.a = __ldg(p.seg_ptr + n);
        s.b = __ldg(p.seg_ptr + n + 1);
        if (p.inv_deg) s.sc = kLn2 * __ldg(p.inv_deg + n);
        s.x = __ldg(reinterpret_cast<const float2*>(p.x + (size_t)n * kC) + lane);
      }
      return s;
    };
    auto load_round = [&](const RoundW& R, Seg (&sg)[kPre]) {
      if (!valid(R)) return;
      const int n_lo = sInfo[R.k].n_lo, n_hi = sInfo[R.k].n_hi;
#pragma unroll
      for (int j = 0; j < kPre; ++j) sg[j] = load_seg(n_lo + rw + j * kRedWarps, n_hi);
    };
    RoundW cur = make_round(0, 0);
    Seg pre[kPre], nxt[kPre];
    load_round(cur, pre);
    for (uint32_t it = 0; valid(cur); ++it) {
      const int b = it & 1;
      const int cnt = cur.cnt, r_lo = cur.r_lo, r_hi = cur.r_lo + cur.cnt;
      const int n_lo = sInfo[cur.k].n_lo, n_hi = sInfo[cur.k].n_hi;
      const float* sV = sVbuf(b);
      const RoundW nr = next_round(cur);
      load_round(nr, nxt);  // in flight under this whole round
      if (prof_me) mark(21);
      if (cnt > 0) {
        umma::mbar_wait(&bar_v_full[b], (ph_v >> b) & 1);
        ph_v ^= 1u << b;
      }
      if (prof_me) mark(22);
      auto sum_seg = [&](int n, const Seg& sg) {
        const int lo = max(sg.a, r_lo), hi = min(sg.b, r_hi);
        const bool empty_seg = (sg.a == sg.b);
        if (empty_seg ? (cur.rd != 0) : (lo >= hi)) return;
        const bool first = empty_seg || (sg.a >= r_lo);
        const bool lastp = empty_seg || (sg.b <= r_hi);
        float2* o = reinterpret_cast<float2*>(p.out + (size_t)n * kC) + lane;
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
        *o = lastp ? make_float2(fmaf(acc.x, sg.sc, sg.x.x), fmaf(acc.y, sg.sc, sg.x.y)) : acc;
      };
#pragma unroll
      for (int j = 0; j < kPre; ++j) {
        const int n = n_lo + rw + j * kRedWarps;
        if (n < n_hi) sum_seg(n, pre[j]);
      }
      for (int n = n_lo + rw + kPre * kRedWarps; n < n_hi; n += kRedWarps) sum_seg(n, load_seg(n, n_hi));  // many tiny segments
      if (cnt > 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_v_free[b]);
      }
      if (prof_me) mark(23);
#pragma unroll
      for (int j = 0; j < kPre; ++j) pre[j] = nxt[j];
      cur = nr;
    }
    __syncthreads();  // teardown barrier of the CTA
    return;
  }
  if (warp > kIssuerWarp) {
    // ---------------- loaders: indices, window decision, node rows of a round.  The indices of round r+1 are
    // requested (into registers) before round r is processed: their HBM latency runs under this round's work.
    const int lt = tid - kLoadWarp0 * 32, lw = warp - kLoadWarp0;
    auto sync_loaders = [] { asm volatile("bar.sync 3, %0;" ::"n"(kLoaders) : "memory"); };
    constexpr int kPer = (2 * kRowsW + kLoaders - 1) / kLoaders;  // index entries per loader thread (3)
    auto fetch = [&](const RoundW& R, int (&v)[kPer]) {
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int i = lt + j * kLoaders, e = i & (kRowsW - 1);
        v[j] = 0;
        if (valid(R) && i < 2 * kRowsW && e < R.cnt) v[j] = __ldg((i < kRowsW ? p.dst_src : p.dst_dst) + R.r_lo + e);
      }
    };
    uint32_t ph_rf = 0, used = 0;
    RoundW R = make_round(0, 0);
    int vcur[kPer], vnext[kPer];
    fetch(R, vcur);
    for (uint32_t it = 0; valid(R); ++it) {
      const int b = it & 1;
      const RoundW Rn = next_round(R);
      fetch(Rn, vnext);
      if (R.cnt > 0) {
        if (lt == 0) mark(18);
        if ((used >> b) & 1) {  // the gate warps have finished with this buffer (two rounds ago)
          umma::mbar_wait(&bar_rows_free[b], (ph_rf >> b) & 1);
          ph_rf ^= 1u << b;
        }
        if (lt == 0) mark(19);
        int* bS = sIdx + b * 2 * kRowsW;
        int s_lo = 0x7fffffff, s_hi = -1;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
          const int i = lt + j * kLoaders, e = i & (kRowsW - 1);
          if (i < 2 * kRowsW) bS[i] = vcur[j];
          if (i < kRowsW && e < R.cnt) { s_lo = min(s_lo, vcur[j]); s_hi = max(s_hi, vcur[j]); }
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
        const int nrows = win ? nq + np_ : 0;
        if (lt == 0) {
          sWin[b] = make_int4(win ? 1 : 0, s_lo, d_lo, nq);
          if (nrows) umma::mbar_arrive_expect_tx(&bar_rows_full[b], (uint32_t)nrows * (uint32_t)(2 * kC * 4));
          else mbar_arrive(&bar_rows_full[b]);
        }
        float* W = sWbuf(b);
        for (int r = lt; r < nrows; r += kLoaders) {  // rows [0,nq) = Q[smin..smax], rows [nq,nq+np) = P[dmin..dmax]
          const float* g = (r < nq) ? p.PQ + (size_t)(s_lo + r) * (4 * kC) + 2 * kC : p.PQ + (size_t)(d_lo + r - nq) * (4 * kC);
          umma::bulk_g2s(W + r * kVW, g, (uint32_t)(2 * kC * 4), &bar_rows_full[b]);
        }
        used |= 1u << b;
        sync_loaders();  // sRed is rewritten next round
        if (lt == 0) mark(20);
      }
#pragma unroll
      for (int j = 0; j < kPer; ++j) vcur[j] = vnext[j];
      R = Rn;
    }
    __syncthreads();  // teardown barrier of the CTA
    return;
  }
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
      uint32_t ph_a = 0, ph_f = 0;  // phase parities, bit b = buffer b
      uint32_t busy = 0;            // bit b: accumulator b holds a round whose epilogue has not been waited for
      RoundW R = make_round(0, 0);
      if (valid(R)) issue_ea_bulk(R);
      for (uint32_t it = 0; valid(R); ++it) {
        const int b = it & 1;
        const RoundW Rn = next_round(R);
        mark(13);
        if (R.cnt > 0) {  // split of this round done: its A operand is staged and the landing zone is free
          umma::mbar_wait(&bar_a_full[b], (ph_a >> b) & 1);
          ph_a ^= 1u << b;
        }
        if (valid(Rn)) issue_ea_bulk(Rn);  // first: the MMA issue below blocks for the MMAs' run time
        if (R.cnt > 0) {
          if ((busy >> b) & 1) {  // the gate warps have read the round that used this accumulator two rounds ago
            umma::mbar_wait(&bar_acc_free[b], (ph_f >> b) & 1);
            ph_f ^= 1u << b;
          }
          umma::fence_after_sync();
          mark(14);
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
          mark(15);
        }
        R = Rn;
      }
    }
    __syncwarp();
    __syncthreads();  // teardown barrier of the CTA
    return;
  }
  if (warp >= kSplitWarp0) {
    // ---------------- splitters: thread = slot = TMEM lane; the edge row -> its hi (warps 16..19) or lo (20..23)
    // tf32 half -> A operand columns, two 32-column tensor-memory stores per round and warp
    const int e = (tid - kSplitWarp0 * 32) & (kRowsW - 1);
    const bool lo_half = warp >= kSplitWarp0 + 4;
    const bool prof_me = (tid == kSplitWarp0 * 32);
    uint32_t ph_ea = 0, ph_m = 0, used = 0;
    RoundW R = make_round(0, 0);
    for (uint32_t it = 0; valid(R); ++it) {
      const int b = it & 1;
      if (R.cnt > 0) {
        if (prof_me) mark(5);
        if ((used >> b) & 1) {  // the MMAs that read this A buffer two rounds ago have retired
          umma::mbar_wait(&bar_mma[b], (ph_m >> b) & 1);
          ph_m ^= 1u << b;
          umma::fence_after_sync();
        }
        if (prof_me) mark(6);
        const uint32_t nb = ea_bulk_bytes(R.r_lo, R.cnt);
        if (nb) {
          umma::mbar_wait(&bar_ea_full, ph_ea);
          ph_ea ^= 1;
        }
        if (prof_me) mark(7);
        const int ea_off = (int)(((long long)R.r_lo * G) & 3);
        const float* row = sEA + ea_off + e * G;
        const int landed = (int)(nb >> 2);  // first float of the landing zone the bulk copy did NOT deliver
        const bool patch = e < R.cnt && landed < ea_off + (e + 1) * G;
        const uint32_t dst = umma::tmem_addr(lo_half ? tm_a_lo(b) : tm_a_hi(b), warp, 0);
#pragma unroll 1
        for (int k0 = 0; k0 < KP; k0 += 32) {
          float v[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] = 0.0f;
          if (e < R.cnt) {
            if ((G & 1) == 0) {  // rows start at an 8-byte offset: 8-byte loads
#pragma unroll
              for (int t = 0; t < 32; t += 2)
                if (k0 + t < G) {
                  const float2 a = *reinterpret_cast<const float2*>(row + k0 + t);
                  v[t] = a.x; v[t + 1] = a.y;
                }
            } else {
#pragma unroll
              for (int t = 0; t < 32; ++t)
                if (k0 + t < G) v[t] = row[k0 + t];
            }
            if (patch) {
#pragma unroll
              for (int t = 0; t < 32; ++t) {
                const int k = k0 + t;
                if (k < G && ea_off + e * G + k >= landed) v[t] = __ldg(p.ea + ((long long)R.r_lo + e) * G + k);
              }
            }
          }
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const float hi = umma::tf32_hi(v[t]);
            v[t] = lo_half ? v[t] - hi : hi;
          }
          umma::tmem_st32(dst + k0, v);
        }
        umma::tmem_st_wait();
        umma::fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_a_full[b]);  // A operand half staged; the landing zone has been read
        used |= 1u << b;
        if (prof_me) mark(8);
      }
      R = next_round(R);
    }
    __syncthreads();  // teardown barrier of the CTA
    return;
  }

  // ---------------- gate warps: thread = slot (TMEM lane), 16 channels; no CTA-wide barrier in the loop
  const bool prof_me = (tid == 0);
  const int q = warp & 3, part = warp >> 2;  // TMEM lane quadrant, channel quarter (16 channels)
  const int c_begin = part * 16;
  uint32_t ph_m = 0, ph_vf = 0, ph_r = 0, used_v = 0;
  RoundW cur = make_round(0, 0);
  for (uint32_t it = 0; valid(cur); ++it) {
    const int b = it & 1;
    const int cnt = cur.cnt;
    if (prof_me) mark(0);
    if (cnt > 0) {
      umma::mbar_wait(&bar_rows_full[b], (ph_r >> b) & 1);  // indices, window record, node rows of this round
      ph_r ^= 1u << b;
      if ((used_v >> b) & 1) {  // the reducers have summed the round that used this message tile two rounds ago
        umma::mbar_wait(&bar_v_free[b], (ph_vf >> b) & 1);
        ph_vf ^= 1u << b;
      }
      if (prof_me) mark(1);
      // node rows of this slot: requested from shared memory (window) or L2 before the wait on the contraction
      const int e_ep = 32 * q + lane;
      const bool live = e_ep < cnt;
      const int4 wr = sWin[b];
      const bool win = wr.x != 0;
      const int* bSrc = sIdx + b * 2 * kRowsW;
      const int ss = live ? bSrc[e_ep] : 0, sd = live ? bSrc[kRowsW + e_ep] : 0;
      const float* sW = sWbuf(b);
      const float* r0 = win ? sW + (wr.w + sd - wr.z) * kVW + c_begin : p.PQ + (size_t)sd * (4 * kC) + c_begin;
      const float* r1 = win ? sW + (ss - wr.y) * kVW + c_begin : p.PQ + (size_t)ss * (4 * kC) + 2 * kC + c_begin;
      umma::mbar_wait(&bar_mma[b], (ph_m >> b) & 1);  // this round's contraction
      ph_m ^= 1u << b;
      umma::fence_after_sync();
      if (prof_me) mark(2);
      float f[16], sacc[16];
      umma::tmem_ld16(umma::tmem_addr(tm_acc(b), q, c_begin), f);
      umma::tmem_ld16(umma::tmem_addr(tm_acc(b), q, kC + c_begin), sacc);
      umma::tmem_ld_wait();
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acc_free[b]);  // the accumulator may be rewritten (round it + 2)
      if (prof_me) mark(3);
      if (live) {
        float* rowv = sVbuf(b) + e_ep * kVP + c_begin;
        const f2_t cf = pk2(-kLog2e, -kLog2e), cs = pk2(kLog2e, kLog2e);
        auto run = [&](auto ld) {  // ld: how the node rows are read (shared or global memory)
#pragma unroll
          for (int j4 = 0; j4 < 16; j4 += 4) {
            const float4 pf = ld(r0 + j4), ps = ld(r0 + kC + j4);
            const float4 qf = ld(r1 + j4), qs = ld(r1 + kC + j4);
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
        if (win) run([](const float* a) { return *reinterpret_cast<const float4*>(a); });
        else run([](const float* a) { return __ldg(reinterpret_cast<const float4*>(a)); });
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bar_rows_free[b]);  // indices / node rows of this buffer are no longer needed
        mbar_arrive(&bar_v_full[b]);     // this warp's part of the message tile is written
      }
      used_v |= 1u << b;
      if (prof_me) mark(4);
    }
    if (PROFILE && pl.prof && tid == 0) atomicAdd(pl.prof + 31, 1ull);
    cur = next_round(cur);
  }
  umma::fence_before_sync();
  __syncthreads();  // teardown barrier of the CTA
  if (warp == 0) umma::tmem_dealloc(tmem, kTmemColsW);
}

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

// C = 64, G <= 64, a tile table that fits (<= 512 tiles per CTA), 16-byte aligned ea / PQ (bulk copies)
bool cgws_supported(const CgParams& p) {
  WsPlan pl;
  const int64_t n_tiles = std::max<int64_t>(1, ceil_div<int64_t>(p.E, kTileW));
  const int64_t grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  return ws_plan(p.C, p.G, &pl) && ceil_div<int64_t>(n_tiles, grid) <= kInfoCapW &&
         (reinterpret_cast<uintptr_t>(p.ea) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.PQ) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.x) & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 7) == 0 &&
         (int64_t)p.N * 4 * p.C < (int64_t)1 << 31;
}

int cgws_launch(CgParams p, cudaStream_t st) {
  WsPlan pl;
  MDL_REQUIRE(ws_plan(p.C, p.G, &pl), "cgconv_fwd_ws: unsupported shape C=%d G=%d", p.C, p.G);
  const char* wenv = getenv("MDL_CGCONV_WINDOW");  // "0": node terms from global memory only (A/B and test switch)
  pl.window = !(wenv && wenv[0] == '0');
  p.c_off = 0; p.CC = p.C; p.cap = kRowsW; p.te = kTileW;
  p.n_tiles = (int)std::max<int64_t>(1, ceil_div<int64_t>(p.E, kTileW));
  const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
  return pl.prof ? ws_launch_t<1>(p, pl, grid, st) : ws_launch_t<0>(p, pl, grid, st);
}

}  // namespace mdl
