// edge_mlp.cu -- the two-layer edge MLP of SchNet's continuous-filter convolution, fused on tcgen05:
//     Y[e, :] = ( act1(X[e, :] W1^T + b1) W2^T + b2 ) * rs[e]            X = edge_attr [E, G], H = O = 128
//
// Reference: PyG InteractionBlock.mlp = Sequential(Linear(G, F), ShiftedSoftplus, Linear(F, F)) and the cosine
// cutoff multiplied onto its output in CFConv.forward, as the reference builds and calls it
// (matdeeplearn/models/schnet.py:81, 134-143): W = mlp(edge_attr) * C(edge_weight).
// On the reference's path that is two SGEMMs over E rows, an [E, F] hidden tensor written and re-read, and
// three elementwise passes per layer.  Here a round of 128 edge rows goes
//     X rows --bulk (TMA) copy--> landing zone --split hi/lo--> tensor memory (A operand)
//     MMA 1 (3xTF32, K = G)   -> D            epilogue 1: + b1, act1 -> T1 (kept for the backward) -> split -> A operand
//     MMA 2 (3xTF32, K = 128) -> D (reused)   epilogue 2: + b2, * rs[e] -> Y
// with both weight matrices resident in shared memory (hi / lo tf32 halves, canonical K-major tiles) for the CTA's
// life and the hidden activations never leaving the SM in the forward (T1 is written once, for the backward).
//
// Backward (k_edge_mlp2_bwd): dPre1 = ((dY * rs) W2) * act1'(pre1), with act1' recovered from T1
// (shifted softplus: sigmoid(pre1) = 1 - exp(-T1) / 2).  The four K = E weight / bias gradient contractions
// (dW2 = (dY * rs)^T T1, dW1 = dPre1^T X) are mdl_linear_wgrad_rs / mdl_linear_wgrad (wgrad_tc.cu).
#include "common.cuh"
#include "umma.cuh"
#include "edge_dev.cuh"

namespace mdl {
namespace {

constexpr int kMW = 512, kMWarps = 16;        // worker threads
constexpr int kMLaunch = kMW + 32;            // + the issuer warp
constexpr int kMRows = 128;                   // edge rows per round = MMA M
constexpr int kH = 128;                       // hidden = output width
constexpr int kMaxSmem = 226 * 1024;
constexpr uint32_t kColD = 0, kColA1 = 128, kColA2 = 256;   // tensor memory: D | A1 hi, lo (64 each) | A2 hi, lo (128 each)

struct Mlp2Params {
  const float* X; const float* W1; const float* b1; const float* W2; const float* b2; const float* rs;
  float* Y; float* T1;
  int64_t E;
  int G, KP, act1, act2;
  uint32_t offW1hi, offW1lo, offW2hi, offW2lo, offEA, offB, total;
};

__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == 1) return fmaxf(x, 0.0f);
  // shifted softplus: softplus(x) - ln 2  (F.softplus' x > 20 branch equals this to fp32 rounding)
  return fmaf(kLn2, lg2_(1.0f + ex2_(-kLog2e * fabsf(x))), fmaxf(x, 0.0f)) - kLn2;
}

__global__ void __launch_bounds__(kMLaunch, 1) k_edge_mlp2_fwd(const Mlp2Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_mma, bar_ea;
  __shared__ uint32_t tmem_base_s;
  __shared__ int sMail[2][4];  // double-buffered {kind (1 MMA1, 2 MMA2, 3 leave), -, next tile row, next tile rows}
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = p.G, KP = p.KP;
  uint8_t* sW1hi = smem + p.offW1hi;
  uint8_t* sW1lo = smem + p.offW1lo;
  uint8_t* sW2hi = smem + p.offW2hi;
  uint8_t* sW2lo = smem + p.offW2lo;
  float* sEA = reinterpret_cast<float*>(smem + p.offEA);
  float* sB = reinterpret_cast<float*>(smem + p.offB);  // b1 | b2
  auto sync_issuer = [] { asm volatile("bar.sync 1, %0;" ::"n"(kMLaunch) : "memory"); };

  const int64_t n_tiles = (p.E + kMRows - 1) / kMRows;
  // ---- setup: TMEM, barriers, both weight matrices split hi / lo into canonical K-major tiles
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) {
    umma::mbar_init(&bar_mma, 1);
    umma::mbar_init(&bar_ea, 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < kH * KP; i += kMLaunch) {   // W1 [H, G] -> B tile [N = h][K = k]
    const int n = i / KP, k = i - n * KP;
    const float w = (k < G) ? __ldg(p.W1 + (size_t)n * G + k) : 0.0f;
    const float hi = umma::tf32_hi(w);
    const int off = umma::tile_offset_bytes(n, k, kH);
    *reinterpret_cast<float*>(sW1hi + off) = hi;
    *reinterpret_cast<float*>(sW1lo + off) = w - hi;
  }
  for (int i = tid; i < kH * kH; i += kMLaunch) {   // W2 [O, H] -> B tile [N = o][K = h]
    const int n = i >> 7, k = i & 127;
    const float w = __ldg(p.W2 + i);
    const float hi = umma::tf32_hi(w);
    const int off = umma::tile_offset_bytes(n, k, kH);
    *reinterpret_cast<float*>(sW2hi + off) = hi;
    *reinterpret_cast<float*>(sW2lo + off) = w - hi;
  }
  for (int i = tid; i < 2 * kH; i += kMLaunch) sB[i] = (i < kH) ? (p.b1 ? __ldg(p.b1 + i) : 0.0f) : (p.b2 ? __ldg(p.b2 + i - kH) : 0.0f);
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma::make_idesc_tf32(kMRows, kH);

  // a round's X rows are one contiguous, 16-byte aligned block (128 rows x G floats); the last, partial block is
  // copied only if its size is a multiple of 16 bytes, else its rows are read from global memory by the split
  auto bulk_ok = [&](int cnt) { return cnt > 0 && (((int64_t)cnt * G) & 3) == 0; };
  auto issue_bulk = [&](int64_t r_lo, int cnt) {  // one thread
    if (!bulk_ok(cnt)) return;
    const uint32_t nb = (uint32_t)cnt * (uint32_t)G * 4;
    umma::mbar_arrive_expect_tx(&bar_ea, nb);
    umma::bulk_g2s(sEA, p.X + r_lo * G, nb, &bar_ea);
  };

  if (warp == kMWarps) {
    // ---------------- issuer warp: MMAs (+ the next round's bulk copy)
    for (uint32_t mb = 0;; mb ^= 1) {
      sync_issuer();
      const int kind = sMail[mb][0];
      if (kind == 3) break;
      if (lane == 0) {
        umma::fence_after_sync();
        const uint32_t step = 2 * (uint32_t)kH * 16;
        if (kind == 1) {
          const int ncnt = sMail[mb][3];
          if (ncnt > 0) issue_bulk(((int64_t)sMail[mb][1] << 31) | (int64_t)sMail[mb][2], ncnt);  // landing zone is free: split done
          uint32_t acc = 0;
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a = tmem + kColA1 + ((pass == 2) ? 64u : 0u);
            const uint32_t b = umma::smem_u32((pass == 1) ? sW1lo : sW1hi);
            for (int kk = 0; kk < (KP >> 3); ++kk) {
              umma::mma_tf32_ts(tmem + kColD, a + kk * 8, umma::make_desc(b + kk * step, (uint32_t)kH * 16, 128), idesc, acc);
              acc = 1;
            }
          }
        } else {
          uint32_t acc = 0;
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a = tmem + kColA2 + ((pass == 2) ? 128u : 0u);
            const uint32_t b = umma::smem_u32((pass == 1) ? sW2lo : sW2hi);
            for (int kk = 0; kk < kH / 8; ++kk) {
              umma::mma_tf32_ts(tmem + kColD, a + kk * 8, umma::make_desc(b + kk * step, (uint32_t)kH * 16, 128), idesc, acc);
              acc = 1;
            }
          }
        }
        umma::mma_commit(&bar_mma);
      }
      __syncwarp();
    }
    __syncthreads();
    return;
  }

  // ---------------- workers: thread = (edge row = TMEM lane, column group)
  const int q = warp & 3, part = warp >> 2;
  const int e = 32 * q + lane;
  uint32_t mb = 0, ph_mma = 0, ph_ea = 0;
  auto post = [&](int kind, int64_t next_lo, int next_cnt) {  // tid 0, before the issuer barrier
    sMail[mb][0] = kind; sMail[mb][1] = (int)(next_lo >> 31); sMail[mb][2] = (int)(next_lo & 0x7fffffff); sMail[mb][3] = next_cnt;
  };
  auto tile_rows = [&](int64_t t) -> int { return (t < n_tiles) ? (int)min((int64_t)kMRows, p.E - t * kMRows) : 0; };
  // split of a tile's X rows into A1 (16 columns per thread)
  auto split = [&](int64_t t) {
    const int cnt = tile_rows(t);
    const int64_t r_lo = t * kMRows;
    const bool from_lz = bulk_ok(cnt);
    if (from_lz) {
      umma::mbar_wait(&bar_ea, ph_ea);
      ph_ea ^= 1;
    }
    const int k0 = 16 * part;
    if (k0 < KP) {
      float v[16];
#pragma unroll
      for (int t2 = 0; t2 < 16; ++t2) v[t2] = 0.0f;
      if (e < cnt) {
        const float* row = from_lz ? sEA + e * G : p.X + (r_lo + e) * G;
        if (from_lz && (G & 1) == 0) {
#pragma unroll
          for (int t2 = 0; t2 < 16; t2 += 2)
            if (k0 + t2 < G) {
              const float2 a = *reinterpret_cast<const float2*>(row + k0 + t2);
              v[t2] = a.x; v[t2 + 1] = a.y;
            }
        } else {
#pragma unroll
          for (int t2 = 0; t2 < 16; ++t2)
            if (k0 + t2 < G) v[t2] = from_lz ? row[k0 + t2] : __ldg(row + k0 + t2);
        }
      }
      float hi[16], lo[16];
#pragma unroll
      for (int t2 = 0; t2 < 16; ++t2) { hi[t2] = umma::tf32_hi(v[t2]); lo[t2] = v[t2] - hi[t2]; }
      umma::tmem_st16(umma::tmem_addr(tmem + kColA1, warp, k0), hi);
      umma::tmem_st16(umma::tmem_addr(tmem + kColA1 + 64, warp, k0), lo);
    }
    umma::tmem_st_wait();
  };

  int64_t t = blockIdx.x;
  if (t < n_tiles) {
    if (tid == 0) issue_bulk(t * kMRows, tile_rows(t));
    split(t);
  }
  for (; t < n_tiles; t += gridDim.x) {
    const int cnt = tile_rows(t);
    const int64_t r_lo = t * kMRows, tn = t + gridDim.x;
    // ---- MMA 1 (its A operand was staged by split(t)); the next tile's X rows are requested with it
    if (tid == 0) post(1, tn * kMRows, tile_rows(tn));
    umma::fence_before_sync();
    sync_issuer();
    mb ^= 1;
    umma::mbar_wait(&bar_mma, ph_mma);
    ph_mma ^= 1;
    umma::fence_after_sync();
    // ---- epilogue 1: hidden activations -> T1 (global, for the backward) and -> A2 (hi / lo)
    {
      float h[32];
      umma::tmem_ld16(umma::tmem_addr(tmem + kColD, q, 32 * part), *reinterpret_cast<float(*)[16]>(h));
      umma::tmem_ld16(umma::tmem_addr(tmem + kColD, q, 32 * part + 16), *reinterpret_cast<float(*)[16]>(h + 16));
      umma::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] = act_fwd(h[j] + sB[32 * part + j], p.act1);
      if (p.T1 && e < cnt) {
        float* dst = p.T1 + (r_lo + e) * kH + 32 * part;   // 128 contiguous bytes per thread: four full-sector stores
#pragma unroll
        for (int j = 0; j < 4; ++j) umma::stg256(dst + 8 * j, h + 8 * j);
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { hi[j] = umma::tf32_hi(h[16 * half + j]); lo[j] = h[16 * half + j] - hi[j]; }
        umma::tmem_st16(umma::tmem_addr(tmem + kColA2, warp, 32 * part + 16 * half), hi);
        umma::tmem_st16(umma::tmem_addr(tmem + kColA2 + 128, warp, 32 * part + 16 * half), lo);
      }
      umma::tmem_st_wait();
    }
    // ---- MMA 2 (D is reused: every thread has read it -- the barrier below orders that)
    if (tid == 0) post(2, 0, 0);
    umma::fence_before_sync();
    sync_issuer();
    mb ^= 1;
    if (tn < n_tiles) split(tn);   // the next tile's A1 operand, staged under MMA 2
    umma::mbar_wait(&bar_mma, ph_mma);
    ph_mma ^= 1;
    umma::fence_after_sync();
    // ---- epilogue 2: + b2, act2, * rs[e] -> Y
    {
      float y[32];
      umma::tmem_ld16(umma::tmem_addr(tmem + kColD, q, 32 * part), *reinterpret_cast<float(*)[16]>(y));
      umma::tmem_ld16(umma::tmem_addr(tmem + kColD, q, 32 * part + 16), *reinterpret_cast<float(*)[16]>(y + 16));
      umma::tmem_ld_wait();
      if (e < cnt) {
        const float s = p.rs ? __ldg(p.rs + r_lo + e) : 1.0f;
        float* dst = p.Y + (r_lo + e) * kH + 32 * part;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v = y[j] + sB[kH + 32 * part + j];
          if (p.act2 == 1) v = fmaxf(v, 0.0f);
          y[j] = v * s;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) umma::stg256(dst + 8 * j, y + 8 * j);
      }
    }
  }
  if (tid == 0) post(3, 0, 0);
  umma::fence_before_sync();
  sync_issuer();  // the issuer warp leaves
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

// ---- backward: dPre1[e, :] = ((dY[e, :] * rs[e]) W2) * act1'(pre1[e, :])
struct Mlp2BwdParams {
  const float* dY; const float* rs; const float* W2; const float* T1; float* dP1;
  int64_t E;
  int act1;
};

__global__ void __launch_bounds__(kMLaunch, 1) k_edge_mlp2_bwd(const Mlp2BwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ int sMail[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint8_t* sWhi = smem;                       // W2^T as B tile [N = h][K = o]: element (h, o) = W2[o][h]
  uint8_t* sWlo = smem + (size_t)kH * kH * 4;
  auto sync_issuer = [] { asm volatile("bar.sync 1, %0;" ::"n"(kMLaunch) : "memory"); };
  const int64_t n_tiles = (p.E + kMRows - 1) / kMRows;

  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 32) {
    umma::mbar_init(&bar_mma, 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < kH * kH; i += kMLaunch) {
    const int o = i >> 7, h = i & 127;        // coalesced read of W2[o][h]
    const float w = __ldg(p.W2 + i);
    const float hi = umma::tf32_hi(w);
    const int off = umma::tile_offset_bytes(h, o, kH);
    *reinterpret_cast<float*>(sWhi + off) = hi;
    *reinterpret_cast<float*>(sWlo + off) = w - hi;
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma::make_idesc_tf32(kMRows, kH);
  const uint32_t colA = 128;                  // A hi [128, 256), lo [256, 384)

  if (warp == kMWarps) {
    for (uint32_t mb = 0;; mb ^= 1) {
      sync_issuer();
      if (sMail[mb] == 3) break;
      if (lane == 0) {
        umma::fence_after_sync();
        const uint32_t step = 2 * (uint32_t)kH * 16;
        uint32_t acc = 0;
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a = tmem + colA + ((pass == 2) ? 128u : 0u);
          const uint32_t b = umma::smem_u32((pass == 1) ? sWlo : sWhi);
          for (int kk = 0; kk < kH / 8; ++kk) {
            umma::mma_tf32_ts(tmem, a + kk * 8, umma::make_desc(b + kk * step, (uint32_t)kH * 16, 128), idesc, acc);
            acc = 1;
          }
        }
        umma::mma_commit(&bar_mma);
      }
      __syncwarp();
    }
    __syncthreads();
    return;
  }

  const int q = warp & 3, part = warp >> 2;
  const int e = 32 * q + lane;
  uint32_t mb = 0, ph = 0;
  // this thread's 32-column piece of (dY * rs) for row e of tile t (zero beyond the matrix)
  auto load_piece = [&](int64_t t, float (&g)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) g[j] = 0.0f;
    const int64_t r = t * kMRows + e;
    if (t < n_tiles && r < p.E) {
      const float s = p.rs ? __ldg(p.rs + r) : 1.0f;
      const float* src = p.dY + r * kH + 32 * part;
#pragma unroll
      for (int j = 0; j < 4; ++j) umma::ldg256(src + 8 * j, g + 8 * j);
#pragma unroll
      for (int j = 0; j < 32; ++j) g[j] *= s;
    }
  };
  float g[32];
  load_piece(blockIdx.x, g);
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t r_lo = t * kMRows;
    const int cnt = (int)min((int64_t)kMRows, p.E - r_lo);
    const bool live = e < cnt;
    // ---- (dY * rs) row piece (requested one tile ahead) -> hi / lo -> A operand (32 columns per thread)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { hi[j] = umma::tf32_hi(g[16 * half + j]); lo[j] = g[16 * half + j] - hi[j]; }
      umma::tmem_st16(umma::tmem_addr(tmem + colA, warp, 32 * part + 16 * half), hi);
      umma::tmem_st16(umma::tmem_addr(tmem + colA + 128, warp, 32 * part + 16 * half), lo);
    }
    umma::tmem_st_wait();
    if (tid == 0) sMail[mb] = 1;
    umma::fence_before_sync();
    sync_issuer();   // also: every thread has read D of the previous tile
    mb ^= 1;
    // ---- T1 row piece requested under the MMAs
    float t1[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) t1[j] = 0.0f;
    if (live) {
      const float* src = p.T1 + (r_lo + e) * kH + 32 * part;
#pragma unroll
      for (int j = 0; j < 4; ++j) umma::ldg256(src + 8 * j, t1 + 8 * j);
    }
    load_piece(t + gridDim.x, g);   // the next tile's gradient rows: in flight under this tile's MMAs and epilogue
    umma::mbar_wait(&bar_mma, ph);
    ph ^= 1;
    umma::fence_after_sync();
    float d[32];
    umma::tmem_ld16(umma::tmem_addr(tmem, q, 32 * part), *reinterpret_cast<float(*)[16]>(d));
    umma::tmem_ld16(umma::tmem_addr(tmem, q, 32 * part + 16), *reinterpret_cast<float(*)[16]>(d + 16));
    umma::tmem_ld_wait();
    if (live) {
      float* dst = p.dP1 + (r_lo + e) * kH + 32 * part;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float tv = t1[j];
        // act1'(pre1) from the saved activation: relu -> [T1 > 0]; shifted softplus -> sigmoid = 1 - exp(-T1) / 2
        const float dact = (p.act1 == 1) ? (tv > 0.0f ? 1.0f : 0.0f) : fmaf(-0.5f, ex2_(-kLog2e * tv), 1.0f);
        d[j] *= dact;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) umma::stg256(dst + 8 * j, d + 8 * j);
    }
  }
  if (tid == 0) sMail[mb] = 3;
  umma::fence_before_sync();
  sync_issuer();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

bool mlp2_plan(int G, Mlp2Params* p) {
  if (G < 1 || G > 64) return false;
  const int KP = (G + 7) & ~7;
  const uint32_t w1 = (uint32_t)kH * KP * 4, w2 = (uint32_t)kH * kH * 4;
  const uint32_t ea = (((uint32_t)kMRows * G * 4) + 15u) & ~15u;
  p->G = G; p->KP = KP;
  p->offW1hi = 0; p->offW1lo = w1; p->offW2hi = 2 * w1; p->offW2lo = 2 * w1 + w2; p->offEA = 2 * w1 + 2 * w2;
  p->offB = p->offEA + ea; p->total = p->offB + 2 * kH * 4;
  return p->total <= (uint32_t)kMaxSmem;
}

}  // namespace
}  // namespace mdl

using namespace mdl;

extern "C" int mdl_edge_mlp2_supported(int32_t G, int32_t H, int32_t O) {
  Mlp2Params p{};
  return (H == kH && O == kH && mlp2_plan(G, &p)) ? 1 : 0;
}

extern "C" int mdl_edge_mlp2_fwd(const float* X, const float* W1, const float* b1, const float* W2, const float* b2,
                                 const float* rowscale, float* Y, float* T1, int64_t E, int32_t G, int32_t H, int32_t O,
                                 int32_t act1, int32_t act2, void* stream) {
  MDL_REQUIRE(X && W1 && W2 && Y && E >= 0, "edge_mlp2_fwd: null pointer");
  MDL_REQUIRE(H == kH && O == kH, "edge_mlp2_fwd: hidden and output width must be 128 (got %d, %d)", H, O);
  MDL_REQUIRE((act1 == 0 || act1 == 1) && (act2 == 0 || act2 == 1), "edge_mlp2_fwd: act 0 (shifted softplus / none) or 1 (relu)");
  MDL_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 31) == 0 &&
                  (!T1 || (reinterpret_cast<uintptr_t>(T1) & 31) == 0), "edge_mlp2_fwd: X 16-byte, Y and T1 32-byte aligned");
  Mlp2Params p{};
  MDL_REQUIRE(mlp2_plan(G, &p), "edge_mlp2_fwd: edge width %d not supported (1..64)", G);
  if (E == 0) return MDL_OK;
  p.X = X; p.W1 = W1; p.b1 = b1; p.W2 = W2; p.b2 = b2; p.rs = rowscale; p.Y = Y; p.T1 = T1; p.E = E;
  p.act1 = act1; p.act2 = act2;
  static std::atomic<int> configured{0};
  if (!configured.load(std::memory_order_acquire)) {
    MDL_CUDA(cudaFuncSetAttribute(k_edge_mlp2_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    configured.store(1, std::memory_order_release);
  }
  const int64_t n_tiles = (E + kMRows - 1) / kMRows;
  const int grid = (int)std::min<int64_t>(n_tiles, kNumSMs);
  k_edge_mlp2_fwd<<<grid, kMLaunch, p.total, as_stream(stream)>>>(p);
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_edge_mlp2_bwd(const float* dY, const float* rowscale, const float* W2, const float* T1, float* dPre1,
                                 int64_t E, int32_t H, int32_t O, int32_t act1, void* stream) {
  MDL_REQUIRE(dY && W2 && T1 && dPre1 && E >= 0, "edge_mlp2_bwd: null pointer");
  MDL_REQUIRE(H == kH && O == kH, "edge_mlp2_bwd: hidden and output width must be 128");
  MDL_REQUIRE(((reinterpret_cast<uintptr_t>(dY) | reinterpret_cast<uintptr_t>(T1) | reinterpret_cast<uintptr_t>(dPre1)) & 31) == 0,
              "edge_mlp2_bwd: dY, T1, dPre1 must be 32-byte aligned");
  if (E == 0) return MDL_OK;
  Mlp2BwdParams p{dY, rowscale, W2, T1, dPre1, E, act1};
  static std::atomic<int> configured{0};
  const int smem = 2 * kH * kH * 4;
  if (!configured.load(std::memory_order_acquire)) {
    MDL_CUDA(cudaFuncSetAttribute(k_edge_mlp2_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    configured.store(1, std::memory_order_release);
  }
  const int64_t n_tiles = (E + kMRows - 1) / kMRows;
  const int grid = (int)std::min<int64_t>(n_tiles, kNumSMs);
  k_edge_mlp2_bwd<<<grid, kMLaunch, smem, as_stream(stream)>>>(p);
  MDL_LAUNCHED();
  return MDL_OK;
}
