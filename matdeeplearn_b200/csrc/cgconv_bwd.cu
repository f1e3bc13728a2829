// cgconv_bwd.cu -- single-pass CGConv backward (dP, dQ, dW_e) with BOTH contractions on tcgen05; 64 channels
// [c_off, c_off + 64) of a C-wide layer per launch (any C >= 64), edge data = edge_attr rows or d_hat (smearing-fused form).
//
// Reference: autograd through PyG's CGConv (matdeeplearn/models/cgcnn.py:80-82,136-145); SURVEY.md Appendix E.1.
// Same tile ownership, staging, gate recompute and deterministic per-destination sums as k_cgconv_tc<BWD_DST>
// (cgconv_tc.cu); what changes:
//   * 640 threads as in cgconv_fwd.cu: 16 epilogue warps + an issuer warp (every tcgen05.mma and bulk copy),
//     setmaxnreg 112 / 32; barrier 1 hands work to the issuer through a double-buffered mailbox
//     (kind 0 = gate recompute, 1 = dW_e, 2 = leave).
//   * gate recompute: A operand (edge rows, hi/lo) in tensor memory, B = W_e in shared memory (as the forward).
//   * dW_e^T[m, k] += sum_slot da[slot, m] * ea[slot, k]  as  D[M = 128 gate-channels][N = 64] accumulated in TMEM
//     over the CTA's whole life (read once at the end):  A = da^T from TENSOR MEMORY (lane = gate-channel m,
//     column = slot; written after [S3] by a column-wise re-read of the value tile), B = ea^T in shared memory,
//     K-major [k][slot] hi/lo, written by the split pass (conflict-free with the 16-byte chunk padding).
//     3 x 16 MMAs of 128x64x8 per round, asynchronous, instead of ~10.5k cycles of mma.sync in the epilogue warps.
//     (kind::tf32 cannot read an MN-major operand in the plain layouts -- profiles/r2_umma_mn_probe.txt -- so the
//     slot-contiguous operand is staged explicitly.)
//   * TMEM: [0,128) recompute accumulator | [128,192) dW_e accumulator | [192,448) operand columns: the edge-row
//     operand (2 x KP) of the recompute first, da^T (2 x 128) after it -- the former is dead when the latter is written.
//   * the operand region and the ea^T tiles are busy until the round's dW_e MMAs retire: the NEXT round waits on
//     bar_dwe before its split (they run under the dQ atomics, the per-destination sums and the next round's loads).
#include "cgconv.cuh"
#include "umma.cuh"
#include "edge_dev.cuh"

namespace mdl {
namespace {

constexpr int kThreads = 512, kWarps = 16, kS2 = kThreads + 32, kLaunch = kThreads + 128;
constexpr int kRows = 128, kTile = 112, kInfoCap = 128;
constexpr int kC = 64, kNP = 2 * kC, kVW = 2 * kC + 4;
constexpr int kKT = 64;                                  // k rows of the ea^T tiles (G <= 64)
constexpr uint32_t kTChunk = kKT * 16 + 16;              // bytes per 4-slot chunk of an ea^T tile (padded)
constexpr uint32_t kColAcc = 0, kColDwe = 128, kColOp = 192, kTmemCols = 512;

unsigned long long* g_bwd_phase_buf = nullptr;

struct Plan {
  unsigned long long* prof;
  int window, KP, dq_bulk;
  uint32_t offBhi, offBlo, offThi, offTlo, offEA, offV, offIdx, offInfo, total;
};

inline bool plan(int C, int G, bool smear, Plan* pl) {
  if (C < kC || (C & 3) || G < 1 || G > kKT) return false;  // a launch serves 64 channels [c_off, c_off + 64) of a C-wide layer
  const int KP = (G + 7) & ~7;
  const uint32_t b = (uint32_t)kNP * KP * 4, t = (uint32_t)(kRows / 4) * kTChunk;
  // smearing-fused form: the edge rows are expanded from d_hat inside the split, no landing zone
  const uint32_t ea = smear ? 0u : ((((uint32_t)kRows * G * 4 + 32) + 15u) & ~15u), v = (uint32_t)kRows * kVW * 4;
  const uint32_t idx = 4 * kRows * 4, info = kInfoCap * 16;
  pl->prof = g_bwd_phase_buf; pl->window = 1; pl->KP = KP; pl->dq_bulk = 0;
  pl->offBhi = 0; pl->offBlo = b; pl->offThi = 2 * b; pl->offTlo = 2 * b + t; pl->offEA = 2 * b + 2 * t;
  pl->offV = pl->offEA + ea; pl->offIdx = pl->offV + v; pl->offInfo = pl->offIdx + idx;
  pl->total = pl->offInfo + info;
  return pl->total <= (uint32_t)kMaxDynSmem;
}

struct Round { int k, rd, r_lo, cnt; bool last; };

template <int PROFILE>
__global__ void __launch_bounds__(kLaunch, 1) k_cgconv_bwd_pipe(const CgParams p, const Plan pl) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_mma, bar_dwe, bar_ea;
  __shared__ uint32_t tmem_base_s;
  __shared__ int sMail[2][4];  // double-buffered {kind, cnt, next r_lo, next cnt (or -1)}: consumers -> issuer warp
  __shared__ float sMu[64];    // smearing-fused form: the basis centres (GaussianSmearing.offset)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = p.G, KP = pl.KP;
  const int C = p.C, c_off = p.c_off;  // layer width (row strides) and this launch's 64-channel chunk
  uint8_t* sBhi = smem + pl.offBhi;
  uint8_t* sBlo = smem + pl.offBlo;
  uint8_t* sThi = smem + pl.offThi;  // ea^T hi: element (k, slot) at (slot/4)*kTChunk + (k/8)*128 + (k%8)*16 + (slot%4)*4
  uint8_t* sTlo = smem + pl.offTlo;
  float* sEA = reinterpret_cast<float*>(smem + pl.offEA);
  float* sV = reinterpret_cast<float*>(smem + pl.offV);    // node rows (window / per slot), then [d a_f | d a_s] per slot
  int* sIdx = reinterpret_cast<int*>(smem + pl.offIdx);    // [2][src | dst][128]
  TileInfo* sInfo = reinterpret_cast<TileInfo*>(smem + pl.offInfo);
  auto sync_consumers = [] { asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory"); };
  auto sync_issuer = [] { asm volatile("bar.sync 1, %0;" ::"n"(kS2) : "memory"); };

  const int my_tiles = (p.n_tiles > (int)blockIdx.x) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  int info_base = 0;
  auto fill_infos = [&](int base) {
    if (tid >= kThreads) return;
    for (int k = base + tid; k < min(my_tiles, base + kInfoCap); k += kThreads) {
      TileInfo t;
      const int tile = blockIdx.x + k * gridDim.x;
      t.n_lo = first_segment_at_or_after<CG_BWD_DST>(p, tile * kTile);
      t.n_hi = (tile == p.n_tiles - 1) ? p.N : first_segment_at_or_after<CG_BWD_DST>(p, (tile + 1) * kTile);
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
      R.r_lo = T.e_lo + rd * kRows;
      R.cnt = max(0, min(T.e_hi - R.r_lo, kRows));
      R.last = R.r_lo + kRows >= T.e_hi;
    }
    return R;
  };
  auto valid = [&](const Round& R) { return R.k < my_tiles; };
  auto next_round = [&](const Round& R) -> Round { return R.last ? make_round(R.k + 1, 0) : make_round(R.k, R.rd + 1); };

  // ---- setup
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, kTmemCols);
  if (tid == 32) {
    umma::mbar_init(&bar_mma, 1);
    umma::mbar_init(&bar_dwe, 1);
    umma::mbar_init(&bar_ea, 1);
    umma::fence_mbar_init();
  }
  fill_infos(0);
  for (int i = tid; i < kNP * KP; i += kLaunch) {
    const int n = i % kNP, k = i / kNP;
    // exponents in base 2: the f-gate columns carry -log2(e), the s-gate columns +log2(e) (as cgconv_fwd_ws.cu)
    const float w = (k < G) ? __ldg(p.WeT + (size_t)k * (2 * C) + (n < kC ? c_off + n : C + c_off + n - kC)) * (n < kC ? -kLog2e : kLog2e) : 0.0f;
    const float hi = umma::tf32_hi(w);
    const int off = umma::tile_offset_bytes(n, k, kNP);
    *reinterpret_cast<float*>(sBhi + off) = hi;
    *reinterpret_cast<float*>(sBlo + off) = w - hi;
  }
  for (uint32_t i = tid; i < (kRows / 4) * kTChunk / 4; i += kLaunch) {  // ea^T tiles: rows k >= G stay zero
    reinterpret_cast<float*>(sThi)[i] = 0.0f;
    reinterpret_cast<float*>(sTlo)[i] = 0.0f;
  }
  if (p.dhat && tid < 64) sMu[tid] = __ldg(p.sm_offset + min(tid, G - 1));
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc_g = umma::make_idesc_tf32(kRows, kNP);  // gate recompute: M = 128 slots, N = 128
  const uint32_t idesc_w = umma::make_idesc_tf32(kNP, kKT);    // dW_e: M = 128 gate-channels, N = 64
  const uint32_t tmA_hi = tmem + kColOp, tmA_lo = tmA_hi + 64;            // edge rows (recompute), 64 columns per half
  const uint32_t tmD_hi = tmem + kColOp, tmD_lo = tmD_hi + kRows;         // da^T (aliases the above)
  uint32_t ph_mma = 0, ph_dwe = 0, ph_ea = 0;

  auto ea_bulk_bytes = [&](int r_lo, int cnt) -> uint32_t {
    if (cnt <= 0 || p.dhat) return 0u;  // smearing-fused form: nothing to copy
    const long long first = (long long)r_lo * G;
    const uint32_t bytes = (uint32_t)(((int)(first & 3) + cnt * G) * 4);
    const bool more = ((long long)p.E * G - (first + (long long)cnt * G)) >= 3;
    return more ? ((bytes + 15u) & ~15u) : (bytes & ~15u);
  };
  auto issue_ea_bulk = [&](int r_lo, int cnt) {
    const uint32_t nb = ea_bulk_bytes(r_lo, cnt);
    if (!nb) return;
    const long long first = (long long)r_lo * G;
    umma::mbar_arrive_expect_tx(&bar_ea, nb);
    umma::bulk_g2s(sEA, p.ea + (first - (first & 3)), nb, &bar_ea);
  };

  // ---- issuer warp
  if (warp >= kWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    bool first_dwe = true;
    for (uint32_t mb = 0; warp == kWarps; mb ^= 1) {
      sync_issuer();
      const int kind = sMail[mb][0], xcnt = sMail[mb][1], y_lo = sMail[mb][2], ycnt = sMail[mb][3];
      if (kind == 2) break;
      if (lane == 0) {
        umma::fence_after_sync();
        if (kind == 0) {  // gate recompute of the staged round, then the next round's edge rows
          if (ycnt >= 0) issue_ea_bulk(y_lo, ycnt);
          if (xcnt > 0) {
            const uint32_t step_b = 2 * (uint32_t)kNP * 16;
            const uint32_t b_hi = umma::smem_u32(sBhi), b_lo = umma::smem_u32(sBlo);
            uint32_t acc = 0;
            for (int pass = 0; pass < 3; ++pass) {
              const uint32_t a = (pass == 2) ? tmA_lo : tmA_hi;
              const uint32_t b = (pass == 1) ? b_lo : b_hi;
              for (int kk = 0; kk < (KP >> 3); ++kk) {
                umma::mma_tf32_ts(tmem + kColAcc, a + kk * 8, umma::make_desc(b + kk * step_b, (uint32_t)kNP * 16, 128),
                                  idesc_g, acc);
                acc = 1;
              }
            }
            umma::mma_commit(&bar_mma);
          }
        } else if (xcnt > 0) {  // dW_e of the round: contraction over its 128 slots (columns of da^T, k-chunks of ea^T)
          const uint32_t t_hi = umma::smem_u32(sThi), t_lo = umma::smem_u32(sTlo);
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a = (pass == 2) ? tmD_lo : tmD_hi;
            const uint32_t b = (pass == 1) ? t_lo : t_hi;
            for (int kk = 0; kk < kRows / 8; ++kk) {
              umma::mma_tf32_ts(tmem + kColDwe, a + kk * 8, umma::make_desc(b + kk * 2 * kTChunk, kTChunk, 128), idesc_w,
                                first_dwe ? 0u : 1u);
              first_dwe = false;
            }
          }
          umma::mma_commit(&bar_dwe);
        }
      }
      __syncwarp();
    }
    __syncthreads();
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");

  const int q = warp & 3, part = warp >> 2;
  const int c_begin = part * 16;
  Round cur = make_round(0, 0);
  int buf = 0;
  uint32_t mb = 0;  // mailbox slot of the next hand-off
  auto post = [&](int kind, int cnt_, int y_lo, int y_cnt) {  // tid 0, before the issuer barrier
    sMail[mb][0] = kind; sMail[mb][1] = cnt_; sMail[mb][2] = y_lo; sMail[mb][3] = y_cnt;
  };
  long long t_prev = PROFILE ? clock64() : 0;
  auto mark = [&](int slot) {
    if (PROFILE && pl.prof && tid == 0) {
      const long long now = clock64();
      atomicAdd(pl.prof + slot, (unsigned long long)(now - t_prev));
      t_prev = now;
    }
  };
  bool dwe_pending = false;  // a dW_e batch has been handed to the issuer and not waited for yet
  bool dwe_ever = false;     // the dW_e accumulator has been written at all (else the partial is zero)
  const SmearConst sc = p.dhat ? smear_const(sMu, G, p.sm_coeff) : SmearConst{0.0f, 0.0f, 0.0f};
  if (valid(cur)) {
    if (tid == 0) issue_ea_bulk(cur.r_lo, cur.cnt);
    if (tid < 2 * kRows) {
      const int e = tid & (kRows - 1);
      if (e < cur.cnt) sIdx[tid] = __ldg((tid < kRows ? p.dst_src : p.dst_dst) + cur.r_lo + e);
    }
  }

  while (valid(cur)) {
    if (cur.k + 2 >= info_base + kInfoCap && info_base + kInfoCap < my_tiles) {
      sync_consumers();
      info_base = cur.k;
      fill_infos(info_base);
      sync_consumers();
    }
    const Round nxt = next_round(cur);
    const int cnt = cur.cnt, r_lo = cur.r_lo, r_hi = cur.r_lo + cur.cnt;
    const int n_lo = sInfo[cur.k - info_base].n_lo, n_hi = sInfo[cur.k - info_base].n_hi;
    const int* bSrc = sIdx + buf * 2 * kRows;
    const int* bDst = bSrc + kRows;
    mark(0);
    sync_consumers();  // [S1] value tile free; this round's indices visible
    mark(1);

    // ---- next indices, window decision, node rows -> registers (as cgconv_fwd.cu)
    int nidx = 0;
    if (valid(nxt) && tid < 2 * kRows) {
      const int e = tid & (kRows - 1);
      if (e < nxt.cnt) nidx = __ldg((tid < kRows ? p.dst_src : p.dst_dst) + nxt.r_lo + e);
    }
    bool win = false;
    int w_smin = 0, w_dmin = 0, w_nq = 0, nrows = 0;
    if (cnt > 0) {
      int s_lo = 0x7fffffff, s_hi = -1;
      const int4 sv = *reinterpret_cast<const int4*>(bSrc + 4 * lane);
      const int e0 = 4 * lane;
      if (e0 + 0 < cnt) { s_lo = min(s_lo, sv.x); s_hi = max(s_hi, sv.x); }
      if (e0 + 1 < cnt) { s_lo = min(s_lo, sv.y); s_hi = max(s_hi, sv.y); }
      if (e0 + 2 < cnt) { s_lo = min(s_lo, sv.z); s_hi = max(s_hi, sv.z); }
      if (e0 + 3 < cnt) { s_lo = min(s_lo, sv.w); s_hi = max(s_hi, sv.w); }
      s_lo = __reduce_min_sync(0xffffffffu, s_lo);
      s_hi = __reduce_max_sync(0xffffffffu, s_hi);
      const int d_lo = bDst[0], d_hi = bDst[cnt - 1];
      const int nq = s_hi - s_lo + 1, np_ = d_hi - d_lo + 1;
      win = pl.window && nq + np_ <= kRows;
      w_smin = s_lo; w_dmin = d_lo; w_nq = nq;
      nrows = win ? nq + np_ : cnt;
    }
    // float4 `col` (0..31) of staged row r: the chunk's f piece (16 float4) then its s piece, C floats further on
    // in a PQ row [P_f | P_s | Q_f | Q_s]
    auto row_src = [&](int r, int col) -> const float4* {
      const float* g;
      if (win) g = (r < w_nq) ? p.PQ + (size_t)(w_smin + r) * (4 * C) + 2 * C : p.PQ + (size_t)(w_dmin + r - w_nq) * (4 * C);
      else g = p.PQ + (size_t)bSrc[r] * (4 * C) + 2 * C;
      return reinterpret_cast<const float4*>(g + c_off + (col < 16 ? 0 : C - kC)) + col;
    };
    // the value tile is free after [S1]: the rows go straight into it with 16-byte cp.async (no register staging, no
    // store phase); they are waited for before the [S2] hand-off
    for (int c = tid; c < nrows * 32; c += kThreads)
      cp_async16(reinterpret_cast<float4*>(sV + (c >> 5) * kVW) + (c & 31), row_src(c >> 5, c & 31));
    const int n0 = n_lo + warp;
    int seg_a = 0, seg_b = 0;
    if (n0 < n_hi) { seg_a = __ldg(p.seg_ptr + n0); seg_b = __ldg(p.seg_ptr + n0 + 1); }
    // smearing-fused form: the normalised distance of this thread's slot (consumed by the split below)
    const float dh = (p.dhat && (tid & (kRows - 1)) < cnt) ? __ldg(p.dhat + r_lo + (tid & (kRows - 1))) : 0.0f;
    mark(2);

    // ---- the last round's dW_e MMAs read the operand columns and the ea^T tiles: both are rewritten below
    if (dwe_pending) {
      umma::mbar_wait(&bar_dwe, ph_dwe);
      ph_dwe ^= 1;
      umma::fence_after_sync();
      dwe_pending = false;
    }
    mark(3);
    // ---- split: edge rows -> hi / lo -> (a) tensor memory, A operand of the gate recompute
    //                                  -> (b) ea^T tiles, B operand of this round's dW_e
    {
      const int ea_off = (int)(((long long)r_lo * G) & 3);
      if (ea_bulk_bytes(r_lo, cnt)) {
        umma::mbar_wait(&bar_ea, ph_ea);
        ph_ea ^= 1;
      }
      const int e = tid & (kRows - 1);
      const float* row = sEA + ea_off + e * G;
      const int landed = (int)(ea_bulk_bytes(r_lo, cnt) >> 2);
      const bool patch = e < cnt && landed < ea_off + (e + 1) * G;
      const uint32_t t_off = (uint32_t)(e >> 2) * kTChunk + (uint32_t)(e & 3) * 4;
      // thread = (slot e, 16-column group): ONE 16-column tensor-memory store per half and warp (a tcgen05.st costs
      // its warp several hundred cycles whatever its width: few wide stores spread over all 16 warps)
      const int k0 = 16 * (tid >> 7);
      if (k0 < KP) {
        float v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t] = 0.0f;
        if (p.dhat) {  // Gaussian basis of this slot, columns k0 .. k0 + 15 (reference process.py:580-590)
          if (e < cnt) {
            smear_chunk8(dh, sMu, k0, G, sc, *reinterpret_cast<float(*)[8]>(v));
            smear_chunk8(dh, sMu, k0 + 8, G, sc, *reinterpret_cast<float(*)[8]>(v + 8));
          }
        } else if (e < cnt) {
          if ((G & 1) == 0) {  // rows start at an 8-byte offset: 8-byte loads
#pragma unroll
            for (int t = 0; t < 16; t += 2)
              if (k0 + t < G) {
                const float2 a = *reinterpret_cast<const float2*>(row + k0 + t);
                v[t] = a.x; v[t + 1] = a.y;
              }
          } else {
#pragma unroll
            for (int t = 0; t < 16; ++t)
              if (k0 + t < G) v[t] = row[k0 + t];
          }
          if (patch) {
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              const int k = k0 + t;
              if (k < G && ea_off + e * G + k >= landed) v[t] = __ldg(p.ea + ((long long)r_lo + e) * G + k);
            }
          }
        }
        float hi[16], lo[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) { hi[t] = umma::tf32_hi(v[t]); lo[t] = v[t] - hi[t]; }
        umma::tmem_st16(umma::tmem_addr(tmA_hi, warp, k0), hi);
        umma::tmem_st16(umma::tmem_addr(tmA_lo, warp, k0), lo);
#pragma unroll
        for (int t = 0; t < 16; ++t) {  // k = k0 + t: core-matrix row k % 8 of k-group k / 8; lanes = slots -> distinct banks
          if (k0 + t < kKT) {
            const uint32_t o = t_off + (uint32_t)((k0 + t) >> 3) * 128 + (uint32_t)((k0 + t) & 7) * 16;
            *reinterpret_cast<float*>(sThi + o) = hi[t];
            *reinterpret_cast<float*>(sTlo + o) = lo[t];
          }
        }
      }
      umma::tmem_st_wait();
    }
    mark(4);
    cp_async_wait_all();   // this thread's node-row chunks have landed in the value tile
    if (valid(nxt) && tid < 2 * kRows) sIdx[(buf ^ 1) * 2 * kRows + tid] = nidx;
    if (tid == 0) post(0, cnt, nxt.r_lo, valid(nxt) ? nxt.cnt : -1);
    umma::fence_proxy_async_smem();
    umma::fence_before_sync();
    mark(5);
    sync_issuer();  // [S2] -> issuer: next edge rows, gate recompute
    mb ^= 1;
    mark(6);

    // ---- epilogue, part 1: a = accumulator + P[dst] + Q[src]
    float f[16], sacc[16];
    const int e_ep = 32 * q + lane;
    const bool live = e_ep < cnt;
    int sd = 0;
    // grad_out row of this slot's destination (pre-divided by the degree): requested before the wait on the
    // contraction, consumed in part 2 (slots are destination-sorted: a warp touches a few distinct rows)
    float4 gq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) gq[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      sd = bDst[e_ep];
      const float4* gp = reinterpret_cast<const float4*>(p.gout + (size_t)sd * C + c_off + c_begin);
#pragma unroll
      for (int j = 0; j < 4; ++j) gq[j] = __ldg(gp + j);
    }
    if (cnt > 0) {
      umma::mbar_wait(&bar_mma, ph_mma);
      ph_mma ^= 1;
      umma::fence_after_sync();
      mark(7);
      umma::tmem_ld16(umma::tmem_addr(tmem, q, kColAcc + c_begin), f);
      umma::tmem_ld16(umma::tmem_addr(tmem, q, kColAcc + kC + c_begin), sacc);
      umma::tmem_ld_wait();
      if (live) {
        const int ss = bSrc[e_ep];
        const float* r0 = win ? sV + (w_nq + sd - w_dmin) * kVW + c_begin : p.PQ + (size_t)sd * (4 * C) + c_off + c_begin;
        const float* r1 = win ? sV + (ss - w_smin) * kVW + c_begin : sV + e_ep * kVW + c_begin;
        const f2_t cf = pk2(-kLog2e, -kLog2e), cs = pk2(kLog2e, kLog2e);
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
          const float4 pf = win ? *reinterpret_cast<const float4*>(r0 + j4) : __ldg(reinterpret_cast<const float4*>(r0 + j4));
          const float4 ps = win ? *reinterpret_cast<const float4*>(r0 + kC + j4)
                                : __ldg(reinterpret_cast<const float4*>(r0 + C + j4));
          const float4 qf = *reinterpret_cast<const float4*>(r1 + j4);
          const float4 qs = *reinterpret_cast<const float4*>(r1 + kC + j4);
          // y = accumulator (base-2 units: W_e is pre-scaled) + c (P + Q), one packed fma per pair
          upk2(fma2(cf, add2(pk2(pf.x, pf.y), pk2(qf.x, qf.y)), pk2(f[j4], f[j4 + 1])), f[j4], f[j4 + 1]);
          upk2(fma2(cf, add2(pk2(pf.z, pf.w), pk2(qf.z, qf.w)), pk2(f[j4 + 2], f[j4 + 3])), f[j4 + 2], f[j4 + 3]);
          upk2(fma2(cs, add2(pk2(ps.x, ps.y), pk2(qs.x, qs.y)), pk2(sacc[j4], sacc[j4 + 1])), sacc[j4], sacc[j4 + 1]);
          upk2(fma2(cs, add2(pk2(ps.z, ps.w), pk2(qs.z, qs.w)), pk2(sacc[j4 + 2], sacc[j4 + 3])), sacc[j4 + 2], sacc[j4 + 3]);
        }
      }
    }
    mark(8);
    umma::fence_before_sync();
    sync_consumers();  // [S2d] node rows read: the value tile may be overwritten
    mark(9);
    // ---- part 2: d m / d a_f, d m / d a_s times grad_out[dst] (pre-divided by the degree) -> value tile
    if (live) {
      float* rowv = sV + e_ep * kVW;
#pragma unroll
      for (int j4 = 0; j4 < 16; j4 += 4) {
        const float4 gv = gq[j4 >> 2];
        f2_t a0, a1, b0, b1;
        gate_pair_bwd(f[j4], f[j4 + 1], sacc[j4], sacc[j4 + 1], gv.x, gv.y, a0, a1);
        gate_pair_bwd(f[j4 + 2], f[j4 + 3], sacc[j4 + 2], sacc[j4 + 3], gv.z, gv.w, b0, b1);
        float4 o0, o1;
        upk2(a0, o0.x, o0.y); upk2(b0, o0.z, o0.w);
        upk2(a1, o1.x, o1.y); upk2(b1, o1.z, o1.w);
        *reinterpret_cast<float4*>(rowv + c_begin + j4) = o0;
        *reinterpret_cast<float4*>(rowv + kC + c_begin + j4) = o1;
      }
    }
    umma::fence_proxy_async_smem();  // the value tile is also read by the bulk (async-proxy) reductions of dQ below
    mark(10);
    sync_consumers();  // [S3] value tile = da of the round
    mark(11);

    // ---- da^T -> tensor memory: thread = gate-channel m (TMEM lane), its quarter of the slots; rows >= cnt are zero
    if (cnt > 0) {
      const int m = tid & (kNP - 1);
      for (int s16 = 2 * part; s16 < 2 * part + 2; ++s16) {  // 16-slot groups 2*part, 2*part+1
        float hi[16], lo[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const int s = 16 * s16 + t;
          const float v = s < cnt ? sV[s * kVW + m] : 0.0f;  // bank = 4 s + m: lanes = consecutive m
          hi[t] = umma::tf32_hi(v);
          lo[t] = v - hi[t];
        }
        umma::tmem_st16(umma::tmem_addr(tmD_hi, warp, 16 * s16), hi);
        umma::tmem_st16(umma::tmem_addr(tmD_lo, warp, 16 * s16), lo);
      }
      umma::tmem_st_wait();
    }
    if (tid == 0) post(1, cnt, 0, -1);
    umma::fence_before_sync();
    mark(12);
    sync_issuer();  // [S4] -> issuer: dW_e MMAs of this round (they run under the atomics and sums below)
    mb ^= 1;
    mark(13);
    dwe_pending = cnt > 0;
    dwe_ever |= cnt > 0;

    // ---- dQ[src] += da.  Default: 16-byte red.global.add.v4.f32, one 512-byte row per warp instruction.
    // MDL_CGCONV_DQ=bulk: one bulk async reduction (TMA, fp32 add performed by the copy engine / L2) per slot row,
    // issued by 8 lanes of every warp -- measured equal (3.86 vs 3.87 ms on the 16384-graph workload: issuing a bulk
    // operation costs its thread ~300 cycles, as much as the vector atomics it replaces), kept as a switch.
    if (pl.dq_bulk) {
      const int e = warp * (kRows / kWarps) + lane;
      if (lane < kRows / kWarps && e < cnt) {
        float* dq = p.out + (size_t)bSrc[e] * (4 * C) + 2 * C + c_off + p.c_skip;   // channels below c_skip: previous chunk
        const float* sv = sV + e * kVW + p.c_skip;
        const uint32_t nb = (uint32_t)(kC - p.c_skip) * 4;
        if (C == kC) {
          umma::bulk_reduce_add_f32(dq, sv, 2 * kC * 4);          // [dQ_f | dQ_s] are adjacent: one 512-byte row
        } else {
          umma::bulk_reduce_add_f32(dq, sv, nb);
          umma::bulk_reduce_add_f32(dq + C, sv + kC, nb);
        }
      }
      umma::bulk_commit();
    } else {
      const int row0 = warp * (kRows / kWarps);
#pragma unroll
      for (int i = 0; i < kRows / kWarps; ++i) {
        const int e = row0 + i;
        if (e < cnt) {
          const float4 v = *(reinterpret_cast<const float4*>(sV + e * kVW) + lane);
          // lanes 0-15: dQ_f piece, 16-31: dQ_s piece; channels below c_skip belong to the previous chunk's launch
          const int ch = 4 * (lane & 15);
          if (ch >= p.c_skip)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.out + (size_t)bSrc[e] * (4 * C) + 2 * C +
                                                                                 (lane < 16 ? 0 : C) + c_off + ch),
                         "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                         : "memory");
        }
      }
    }
    mark(14);
    // ---- dP[i] = sum over the segment's slots, slot order
    for (int n = n0; n < n_hi; n += kWarps) {
      int a, b;
      if (n == n0) { a = seg_a; b = seg_b; }
      else { a = __ldg(p.seg_ptr + n); b = __ldg(p.seg_ptr + n + 1); }
      const int lo = max(a, r_lo), hi = min(b, r_hi);
      const bool empty_seg = (a == b);
      if (empty_seg ? (cur.rd != 0) : (lo >= hi)) continue;
      const bool first = empty_seg || (a >= r_lo);
      // value-tile column 32 u + lane: u = 0, 1 -> dP_f channels, u = 2, 3 -> dP_s channels of the chunk
      float* o = p.out + (size_t)n * (4 * C) + c_off + lane;
      const int ooff[4] = {0, 32, C, C + 32};
      float acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = first ? 0.0f : o[ooff[u]];
      for (int s = lo; s < hi; ++s) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += sV[(s - r_lo) * kVW + 32 * u + lane];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) o[ooff[u]] = acc[u];
    }
    mark(15);
    if (pl.dq_bulk) umma::bulk_wait_read();  // the reductions have read the value tile: the next round may overwrite it
    if (PROFILE && pl.prof && tid == 0) atomicAdd(pl.prof + 31, 1ull);
    cur = nxt; buf ^= 1;
  }
  if (pl.dq_bulk) umma::bulk_wait_all();     // every reduction of this thread has been performed

  // ---- the CTA's dW_e^T partial: accumulator lane = gate-channel m, column = k
  if (dwe_pending) {
    umma::mbar_wait(&bar_dwe, ph_dwe);
    umma::fence_after_sync();
  }
  if (my_tiles > 0) {
    float* part_out = p.dW_part + (size_t)blockIdx.x * G * kNP;
    const int m = 32 * q + lane;
    float d[16];
    umma::tmem_ld16(umma::tmem_addr(tmem, q, kColDwe + c_begin), d);  // columns k = c_begin .. c_begin + 15
    umma::tmem_ld_wait();
    if (!dwe_ever) {
#pragma unroll
      for (int j = 0; j < 16; ++j) d[j] = 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c_begin + j < G) part_out[(size_t)(c_begin + j) * kNP + m] = d[j];
  }
  if (tid == 0) post(2, 0, 0, -1);
  umma::fence_before_sync();
  sync_issuer();  // the issuer warp leaves
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
}

template <int PROFILE>
int bwd_launch_t(const CgParams& p, const Plan& pl, int grid, cudaStream_t st) {
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MDL_CUDA(cudaFuncSetAttribute(k_cgconv_bwd_pipe<PROFILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    configured.store(1, std::memory_order_release);
  }
  k_cgconv_bwd_pipe<PROFILE><<<grid, kLaunch, pl.total, st>>>(p, pl);
  MDL_LAUNCHED();
  return MDL_OK;
}

}  // namespace

void cgbwd_set_phase_buffer(unsigned long long* dev_ptr) { g_bwd_phase_buf = dev_ptr; }

// C >= 64 (multiple of 4; 64 channels per launch), G <= 64, 16-byte aligned ea / PQ / gout / dPQ (bulk copy, vector
// loads, vector atomics)
bool cgbwd_supported(const CgParams& p) {
  Plan pl;
  return plan(p.C, p.G, p.dhat != nullptr, &pl) && (reinterpret_cast<uintptr_t>(p.ea) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.PQ) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.gout) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && (int64_t)p.N * 4 * p.C < (int64_t)1 << 31;
}

// dP (by destination, deterministic) + dQ (vector atomics: the caller zeroes the dQ half first) + dW_e^T: per-CTA
// partials in p.dW_part ([grid][G][128]) summed in CTA order into dWeT [G][2C].  One launch per 64-channel chunk;
// the last chunk of a width that is not a multiple of 64 starts at C - 64 and recomputes the channels it shares with
// the one before (identical values; plain stores for dP / dW_e, its dQ atomics skip them: c_skip).
int cgbwd_launch(CgParams p, cudaStream_t st, float* dWeT) {
  Plan pl;
  MDL_REQUIRE(plan(p.C, p.G, p.dhat != nullptr, &pl), "cgconv_bwd: unsupported shape C=%d G=%d", p.C, p.G);
  const char* wenv = getenv("MDL_CGCONV_WINDOW");
  pl.window = !(wenv && wenv[0] == '0');
  const char* qenv = getenv("MDL_CGCONV_DQ");   // "bulk": bulk async reductions instead of vector atomics (A/B)
  pl.dq_bulk = (qenv && strcmp(qenv, "bulk") == 0) ? 1 : 0;
  p.CC = kC; p.cap = kRows; p.te = kTile;
  p.n_tiles = (int)std::max<int64_t>(1, ceil_div<int64_t>(p.E, kTile));
  const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;
  for (int c0 = 0; c0 < p.C; c0 += kC) {
    p.c_off = std::min(c0, p.C - kC);
    p.c_skip = c0 - p.c_off;
    if (int rc = pl.prof ? bwd_launch_t<1>(p, pl, grid, st) : bwd_launch_t<0>(p, pl, grid, st)) return rc;
    if (int rc = reduce_dw_partials(p.dW_part, grid, p.G, p.C, p.c_off, kC, dWeT, st)) return rc;
  }
  return MDL_OK;
}

}  // namespace mdl
