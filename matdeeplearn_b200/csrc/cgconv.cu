// cgconv.cu -- fused gather -> per-edge gated message -> segmented aggregate for
// PyG CGConv as the reference builds it (matdeeplearn/models/cgcnn.py:80-82,
// called cgcnn.py:136-145).  One pass over destination-sorted edge slots:
//
//   a[s]   = P[dst(s)] + Q[src(s)] + We . ea[s]          (P,Q: node-level GEMMs done by
//   m[s]   = sigmoid(a_f) * softplus(a_s)                  the caller; We . ea: here)
//   out[i] = aggr_{s: dst(s)=i} m[s] + x[i]
//
// Nothing of size [E, *] is ever written to HBM, forward or backward.
//
// Work decomposition ("tile ownership"): tile t owns the segments (destination
// nodes forward / dP pass, source nodes for the dQ pass) whose FIRST slot lies in
// [t*TE, (t+1)*TE).  A CTA walks its owned slot range in rounds of CAP slots, so
// a segment is always reduced by exactly one CTA, in slot order: deterministic,
// no atomics, any degree (hubs just take more rounds).
//
// Per round:  stage ea rows + indices to smem -> register-tiled contraction
// [CAP x G] x [G x 2C] (8 edges x 8 outputs per thread) -> epilogue (gather P/Q
// rows as float4, gate math) -> smem -> warp-per-segment sum -> coalesced store.
// Backward recomputes the gates (saves nothing per edge), produces
//   dP[i] = sum_{s: dst=i} da[s]   (BWD_DST pass, also accumulates dWe = da^T ea in
//   dQ[j] = sum_{s: src=j} da[s]    registers across all tiles of the CTA)
#include "cgconv.cuh"

#include <stdlib.h>
#include <string.h>

namespace mdl {

template <int MODE, int NITEM>
__global__ void __launch_bounds__(kThreads, 1) k_cgconv(const CgParams p) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, G = p.G, CC = p.CC, W2 = 2 * CC;
  const int G4 = (G + 3) & ~3, GS = (G + 7) & ~7;
  const int CAP = p.cap, NEG = CAP >> 3, NCG = CC >> 2;
  const int VW = (MODE == CG_FWD) ? CC : W2;  // width of the per-slot value tile

  float* sW = smem;                  // [G4][W2]
  float* sEA = sW + G4 * W2;         // [CAP][GS]   (pads zero)
  float* sV = sEA + CAP * GS;        // [CAP][VW]
  int* sSrc = reinterpret_cast<int*>(sV + CAP * VW);  // [CAP]
  int* sDst = sSrc + CAP;            // [CAP]
  int* sSlot = sDst + CAP;           // [CAP]
  __shared__ int sh_bounds[2];

  // ---- one-time: weights -> smem (zero k-padding), zero the ea tile (its pads stay zero)
  for (int i = tid; i < G4 * W2; i += kThreads) {
    const int k = i / W2, c = i - k * W2;
    float v = 0.0f;
    if (k < G) {
      const int col = (c < CC) ? (p.c_off + c) : (C + p.c_off + (c - CC));
      v = __ldg(p.WeT + (size_t)k * (2 * C) + col);
    }
    sW[i] = v;
  }
  for (int i = tid; i < CAP * GS; i += kThreads) sEA[i] = 0.0f;

  // dWe accumulators live in registers for the whole kernel (BWD_DST only)
  float dw[NITEM][4][8];
  if (MODE == CG_BWD_DST) {
#pragma unroll
    for (int j = 0; j < NITEM; ++j)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) dw[j][a][b] = 0.0f;
  }
  __syncthreads();

  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    tile_bounds<MODE>(p, tile, p.te, tid, sh_bounds);  // the last tile also owns trailing empty segments
    __syncthreads();
    const int n_lo = sh_bounds[0], n_hi = sh_bounds[1];
    if (n_hi <= n_lo) { __syncthreads(); continue; }
    const int e_lo = __ldg(p.seg_ptr + n_lo), e_hi = __ldg(p.seg_ptr + n_hi);
    const int rounds = max(1, (e_hi - e_lo + CAP - 1) / CAP);

    for (int rd = 0; rd < rounds; ++rd) {
      const int r_lo = e_lo + rd * CAP;
      const int r_hi = min(e_hi, r_lo + CAP);
      const int cnt = r_hi - r_lo;

      // ---- stage indices and ea rows
      for (int i = tid; i < CAP; i += kThreads) {
        int slot = 0, s = 0, d = 0;
        if (i < cnt) {
          slot = (MODE == CG_BWD_SRC) ? __ldg(p.src_slot + r_lo + i) : (r_lo + i);
          s = __ldg(p.dst_src + slot);
          d = __ldg(p.dst_dst + slot);
        }
        sSlot[i] = slot; sSrc[i] = s; sDst[i] = d;
      }
      if (MODE == CG_BWD_SRC) __syncthreads();
      for (int e = warp; e < CAP; e += kWarps) {
        if (e < cnt) {
          const int slot = (MODE == CG_BWD_SRC) ? sSlot[e] : (r_lo + e);
          const float* row = p.ea + (size_t)slot * G;
          for (int k = lane; k < G; k += 32) sEA[e * GS + k] = __ldg(row + k);
        } else {
          for (int k = lane; k < G; k += 32) sEA[e * GS + k] = 0.0f;
        }
      }
      __syncthreads();

      // ---- contraction + epilogue, 8 slots x (4 f + 4 s) channels per work item
      const int n_items = NCG * NEG;
      for (int it = tid; it < n_items; it += kThreads) {
        const int cg = it % NCG, eg = it / NCG;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
        const float* wbase = sW + 4 * cg;
        for (int k4 = 0; k4 < G4; k4 += 4) {
          float4 a4[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            a4[i] = *reinterpret_cast<const float4*>(sEA + (i * NEG + eg) * GS + k4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float4 wf = *reinterpret_cast<const float4*>(wbase + (k4 + kk) * W2);
            const float4 ws = *reinterpret_cast<const float4*>(wbase + (k4 + kk) * W2 + CC);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float a = kk == 0 ? a4[i].x : kk == 1 ? a4[i].y : kk == 2 ? a4[i].z : a4[i].w;
              acc[i][0] = fmaf(a, wf.x, acc[i][0]);
              acc[i][1] = fmaf(a, wf.y, acc[i][1]);
              acc[i][2] = fmaf(a, wf.z, acc[i][2]);
              acc[i][3] = fmaf(a, wf.w, acc[i][3]);
              acc[i][4] = fmaf(a, ws.x, acc[i][4]);
              acc[i][5] = fmaf(a, ws.y, acc[i][5]);
              acc[i][6] = fmaf(a, ws.z, acc[i][6]);
              acc[i][7] = fmaf(a, ws.w, acc[i][7]);
            }
          }
        }
        // epilogue
        const int c0 = p.c_off + 4 * cg;  // global channel of this item's first lane
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int e = i * NEG + eg;
          if (e < cnt) {
            const float* Pd = p.PQ + (size_t)sDst[e] * (4 * C);
            const float* Qs = p.PQ + (size_t)sSrc[e] * (4 * C) + 2 * C;
            const float4 pf = __ldg(reinterpret_cast<const float4*>(Pd + c0));
            const float4 ps = __ldg(reinterpret_cast<const float4*>(Pd + C + c0));
            const float4 qf = __ldg(reinterpret_cast<const float4*>(Qs + c0));
            const float4 qs = __ldg(reinterpret_cast<const float4*>(Qs + C + c0));
            float af[4] = {acc[i][0] + pf.x + qf.x, acc[i][1] + pf.y + qf.y,
                           acc[i][2] + pf.z + qf.z, acc[i][3] + pf.w + qf.w};
            float as[4] = {acc[i][4] + ps.x + qs.x, acc[i][5] + ps.y + qs.y,
                           acc[i][6] + ps.z + qs.z, acc[i][7] + ps.w + qs.w};
            if (MODE == CG_FWD) {
              float4 m;
              m.x = sigmoid_fast_(af[0]) * softplusf_(as[0]);
              m.y = sigmoid_fast_(af[1]) * softplusf_(as[1]);
              m.z = sigmoid_fast_(af[2]) * softplusf_(as[2]);
              m.w = sigmoid_fast_(af[3]) * softplusf_(as[3]);
              *reinterpret_cast<float4*>(sV + e * VW + 4 * cg) = m;
            } else {
              const int d = sDst[e];
              float4 g = __ldg(reinterpret_cast<const float4*>(p.gout + (size_t)d * C + c0));
              float gg[4] = {g.x, g.y, g.z, g.w};  // grad_out arrives pre-divided by the degree (mean)
              float dfv[4], dsv[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float sg = sigmoid_fast_(af[j]);
                const float sp = softplusf_(as[j]);
                dfv[j] = gg[j] * sp * sg * (1.0f - sg);
                dsv[j] = gg[j] * sg * sigmoid_fast_(as[j]);
              }
              *reinterpret_cast<float4*>(sV + e * VW + 4 * cg) = make_float4(dfv[0], dfv[1], dfv[2], dfv[3]);
              *reinterpret_cast<float4*>(sV + e * VW + CC + 4 * cg) = make_float4(dsv[0], dsv[1], dsv[2], dsv[3]);
            }
          }
        }
      }
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
          float* o = p.out + (size_t)n * C + p.c_off;
          const float* xr = p.x + (size_t)n * C + p.c_off;
          const float sc = p.inv_deg ? __ldg(p.inv_deg + n) : 1.0f;
          for (int c = lane; c < CC; c += 32) {
            float acc = first ? 0.0f : o[c];
            for (int s = lo; s < hi; ++s) acc += sV[(s - r_lo) * VW + c];
            o[c] = last ? fmaf(acc, sc, __ldg(xr + c)) : acc;
          }
        } else {
          // dPQ row: [dP_f | dP_s | dQ_f | dQ_s], each C wide
          float* o = p.out + (size_t)n * (4 * C) + (MODE == CG_BWD_SRC ? 2 * C : 0) + p.c_off;
          for (int c = lane; c < W2; c += 32) {
            const int oc = (c < CC) ? c : (C + (c - CC));
            float acc = first ? 0.0f : o[oc];
            for (int s = lo; s < hi; ++s) acc += sV[(s - r_lo) * VW + c];
            o[oc] = acc;
          }
        }
      }

      // ---- dWe += da^T . ea over this round's slots (register tiles, 4 ch x 8 k)
      if (MODE == CG_BWD_DST) {
        const int n_c4 = W2 >> 2;
        const int n_dw = n_c4 * (GS >> 3);
#pragma unroll
        for (int j = 0; j < NITEM; ++j) {
          const int it = tid + j * kThreads;
          if (it < n_dw) {
            const int c4 = it % n_c4, k8 = it / n_c4;
            for (int e = 0; e < cnt; ++e) {
              const float4 da = *reinterpret_cast<const float4*>(sV + e * VW + 4 * c4);
              const float4 e0 = *reinterpret_cast<const float4*>(sEA + e * GS + 8 * k8);
              const float4 e1 = *reinterpret_cast<const float4*>(sEA + e * GS + 8 * k8 + 4);
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
    const int n_dw = n_c4 * (GS >> 3);
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
}

// dWeT[k][col(c)] = sum_b part[b][k][c]   (fixed order -> deterministic)
__global__ void k_cg_reduce_dw(const float* __restrict__ part, int nparts, int G, int C, int c_off,
                               int CC, float* __restrict__ dWeT) {
  const int W2 = 2 * CC;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * W2) return;
  float acc = 0.0f;
#pragma unroll 8
  for (int b = 0; b < nparts; ++b) acc += __ldg(part + (size_t)b * G * W2 + i);  // fixed order; 8 loads in flight
  const int k = i / W2, c = i - k * W2;
  const int col = (c < CC) ? (c_off + c) : (C + c_off + (c - CC));
  dWeT[(size_t)k * (2 * C) + col] = acc;
}

int reduce_dw_partials(const float* part, int nparts, int G, int C, int c_off, int CC, float* dWeT, cudaStream_t st) {
  const int tot = G * 2 * CC;
  k_cg_reduce_dw<<<ceil_div(tot, 256), 256, 0, st>>>(part, nparts, G, C, c_off, CC, dWeT);
  MDL_LAUNCHED();
  return MDL_OK;
}

struct CgPlan {
  int CC, cap, te, nitem;
  size_t smem;
};

static size_t cg_smem_bytes(int mode, int CC, int G, int cap) {
  const int G4 = (G + 3) & ~3, GS = (G + 7) & ~7, W2 = 2 * CC;
  const int VW = (mode == CG_FWD) ? CC : W2;
  return (size_t)(G4 * W2 + cap * GS + cap * VW) * 4 + (size_t)3 * cap * 4;
}

// largest channel chunk / slot capacity that fits the 227 KB opt-in limit
static bool cg_plan(int mode, int C, int G, CgPlan* plan) {
  const size_t limit = kMaxDynSmem;
  for (int CC = C; CC >= 4; CC = ((CC / 2) + 3) & ~3) {
    const int W2 = 2 * CC, GS = (G + 7) & ~7;
    const int n_dw = (W2 / 4) * (GS / 8);
    const int nitem = (n_dw + kThreads - 1) / kThreads;
    if (mode == CG_BWD_DST && nitem > 4) { if (CC == 4) break; continue; }
    for (int cap = 128; cap >= 32; cap >>= 1) {
      size_t s = cg_smem_bytes(mode, CC, G, cap);
      if (s <= limit) {
        plan->CC = CC; plan->cap = cap; plan->te = cap - cap / 8;
        plan->nitem = nitem <= 1 ? 1 : nitem <= 2 ? 2 : 4;
        plan->smem = s;
        return true;
      }
    }
    if (CC == 4) break;
  }
  return false;
}

template <int MODE, int NITEM>
static int cg_launch(const CgParams& p, size_t smem, int grid, cudaStream_t st) {
  static std::atomic<int> configured{0};  // opt in to >48 KB dynamic smem once per instantiation
  if (!configured.load(std::memory_order_acquire)) {
    MDL_CUDA(cudaFuncSetAttribute(k_cgconv<MODE, NITEM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kMaxDynSmem));
    configured.store(1, std::memory_order_release);
  }
  k_cgconv<MODE, NITEM><<<grid, kThreads, smem, st>>>(p);
  MDL_LAUNCHED();
  return MDL_OK;
}

static int cg_grid(int n_tiles) { return n_tiles < kNumSMs ? n_tiles : kNumSMs; }

static int cg_check(int64_t N, int64_t E, int C, int G, int reduce) {
  MDL_REQUIRE(N > 0 && E >= 0 && N < (1LL << 31) && E < (1LL << 31), "cgconv: bad N/E");
  MDL_REQUIRE(C >= 4 && C % 4 == 0, "cgconv: channels must be a multiple of 4 (got %d)", C);
  MDL_REQUIRE(G >= 1 && G <= 4096, "cgconv: unsupported edge feature width %d", G);
  MDL_REQUIRE(reduce == MDL_REDUCE_SUM || reduce == MDL_REDUCE_MEAN,
              "cgconv: aggr must be add or mean (reference uses mean, cgcnn.py:81)");
  return MDL_OK;
}

// out[i] = sum_b part[b][i], i < len.  Block = 32 outputs x 8 partial-groups: warp w sums partials
// w, w+8, ... (coalesced 128-byte reads), the 8 group sums are combined in a fixed order.
__global__ void __launch_bounds__(256)
k_sum_partials(const float* __restrict__ part, int nparts, int64_t stride, int64_t len,
               float* __restrict__ out0, int64_t len0, float* __restrict__ out1) {
  __shared__ float sm[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  float acc = 0.0f;
  if (i < len) {
#pragma unroll 4
    for (int b = warp; b < nparts; b += 8) acc += __ldg(part + (size_t)b * stride + i);
  }
  sm[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && i < len) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][lane];
    if (i < len0) out0[i] = t;
    else if (out1) out1[i - len0] = t;
  }
}

int sum_partials(const float* part, int nparts, int64_t stride, int64_t len, float* out0, int64_t len0,
                 float* out1, cudaStream_t st) {
  k_sum_partials<<<(int)ceil_div<int64_t>(len, 32), 256, 0, st>>>(part, nparts, stride, len, out0, len0, out1);
  MDL_LAUNCHED();
  return MDL_OK;
}

// MDL_CGCONV_IMPL=simt forces the SIMT kernels (cross-checks, fallback timing)
static bool use_tc(int mode, int C, int G) {
  const char* env = getenv("MDL_CGCONV_IMPL");
  if (env && strcmp(env, "simt") == 0) return false;
  // odd G takes the 4-byte-granule staging paths of the tensor-core kernels (tests: test_cgconv_odd_edge_width_c64)
  return cgtc_supported(mode, C, G);
}

}  // namespace mdl

using namespace mdl;

extern "C" size_t mdl_cgconv_workspace_bytes(int64_t N, int64_t E, int32_t C, int32_t G) {
  (void)N; (void)E;
  // per-CTA dWe partials for the BWD_DST pass
  return (size_t)2 * kNumSMs * (size_t)G * 2 * (size_t)C * 4 + 256;  // TC path: two partials per CTA
}

extern "C" int mdl_cgconv_fwd(const float* x, const float* PQ, const float* ea, const float* WeT,
                              const int32_t* dst_ptr, const int32_t* dst_src,
                              const int32_t* dst_dst, const float* inv_deg_dst, float* out,
                              int64_t N, int64_t E, int32_t C, int32_t G, int32_t reduce,
                              void* stream) {
  if (int rc = cg_check(N, E, C, G, reduce)) return rc;
  MDL_REQUIRE(x && PQ && WeT && dst_ptr && out && (E == 0 || (ea && dst_src && dst_dst)),
              "cgconv_fwd: null pointer");
  MDL_REQUIRE(reduce != MDL_REDUCE_MEAN || inv_deg_dst, "cgconv_fwd: mean needs inv_deg");
  CgPlan plan;
  MDL_REQUIRE(cg_plan(CG_FWD, C, G, &plan), "cgconv_fwd: C=%d G=%d does not fit shared memory", C, G);
  CgParams p{};
  p.x = x; p.PQ = PQ; p.ea = ea; p.WeT = WeT; p.seg_ptr = dst_ptr; p.dst_src = dst_src;
  p.dst_dst = dst_dst; p.inv_deg = (reduce == MDL_REDUCE_MEAN) ? inv_deg_dst : nullptr;
  p.out = out; p.N = (int)N; p.E = (int)E; p.C = C; p.G = G;
  {
    // default: the warp-specialised forward kernel (cgconv_fwd_ws.cu; any C >= 64 in 64-channel chunks);
    // MDL_CGCONV_IMPL=pipe selects the software-pipelined one (cgconv_fwd.cu), =tc the round-serial tensor-core kernel
    // (both C = 64 only: A/B, and the alignments the first declines), =simt the SIMT kernel below
    const char* env = getenv("MDL_CGCONV_IMPL");
    const bool want_tc = env && strcmp(env, "tc") == 0, want_pipe = env && strcmp(env, "pipe") == 0;
    const bool want_simt = env && strcmp(env, "simt") == 0;
    if (!want_simt && !want_tc && !want_pipe && cgws_supported(p)) return cgws_launch(p, as_stream(stream));
    if (use_tc(CG_FWD, C, G)) {
      if (!want_tc && cgfwd_supported(p)) return cgfwd_launch(p, as_stream(stream));
      return cgtc_launch(CG_FWD, p, as_stream(stream), nullptr);
    }
  }
  p.cap = plan.cap; p.te = plan.te;
  p.n_tiles = (int)std::max<int64_t>(1, ceil_div<int64_t>(E, plan.te));
  for (int c_off = 0; c_off < C; c_off += plan.CC) {
    p.c_off = c_off; p.CC = std::min(plan.CC, C - c_off);
    size_t smem = cg_smem_bytes(CG_FWD, p.CC, G, plan.cap);
    if (int rc = cg_launch<CG_FWD, 1>(p, smem, cg_grid(p.n_tiles), as_stream(stream))) return rc;
  }
  return MDL_OK;
}

extern "C" int mdl_cgconv_bwd(const float* gout, const float* PQ, const float* ea,
                              const float* WeT, const int32_t* dst_ptr, const int32_t* dst_src,
                              const int32_t* dst_dst, const int32_t* src_ptr,
                              const int32_t* src_slot, const float* inv_deg_dst, float* dPQ,
                              float* dWeT, int64_t N, int64_t E, int32_t C, int32_t G,
                              int32_t reduce, void* workspace, size_t workspace_bytes,
                              void* stream) {
  if (int rc = cg_check(N, E, C, G, reduce)) return rc;
  MDL_REQUIRE(gout && PQ && WeT && dst_ptr && src_ptr && dPQ && dWeT &&
                  (E == 0 || (ea && dst_src && dst_dst && src_slot)),
              "cgconv_bwd: null pointer");
  MDL_REQUIRE(reduce != MDL_REDUCE_MEAN || inv_deg_dst, "cgconv_bwd: mean needs inv_deg");
  if (!workspace || workspace_bytes < mdl_cgconv_workspace_bytes(N, E, C, G)) {
    set_error("cgconv_bwd: workspace %zu < %zu", workspace_bytes,
              mdl_cgconv_workspace_bytes(N, E, C, G));
    return MDL_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  CgParams p{};
  p.gout = gout; p.PQ = PQ; p.ea = ea; p.WeT = WeT; p.dst_src = dst_src; p.dst_dst = dst_dst;
  p.src_slot = src_slot; p.inv_deg = (reduce == MDL_REDUCE_MEAN) ? inv_deg_dst : nullptr;
  p.out = dPQ; p.dW_part = (float*)workspace; p.N = (int)N; p.E = (int)E; p.C = C; p.G = G;

  // Tensor-core path, default: ONE pass in destination order; dQ[src] is accumulated with 16-byte
  // vector float atomics (like the reference's own scatter kernels, the summation order then varies
  // run to run).  MDL_CGCONV_DETERMINISTIC=1 selects the two-pass scheme (second pass over the
  // by-source view, fixed summation order, bitwise reproducible).
  const char* det_env = getenv("MDL_CGCONV_DETERMINISTIC");
  const char* ienv = getenv("MDL_CGCONV_IMPL");
  const char* benv = getenv("MDL_CGCONV_BWD");
  const bool det = det_env && det_env[0] == '1';
  p.seg_ptr = dst_ptr;
  // default single-pass kernel: both contractions on tcgen05 (cgconv_bwd.cu; any C >= 64 in 64-channel chunks);
  // MDL_CGCONV_BWD=tc keeps the mma.sync one (C = 64, A/B), MDL_CGCONV_IMPL=simt the SIMT kernels
  const bool pipe_ok = !det && !(ienv && strcmp(ienv, "simt") == 0) && !(benv && strcmp(benv, "tc") == 0) && cgbwd_supported(p);
  const bool tc64 = use_tc(CG_BWD_DST, C, G);
  const bool dq_atomic = pipe_ok || (tc64 && !det);
  if (dq_atomic)
    MDL_CUDA(cudaMemset2DAsync(dPQ + 2 * C, (size_t)4 * C * 4, 0, (size_t)2 * C * 4, (size_t)N, st));
  // pass A: destination order -> dP and dWe
  if (pipe_ok) {
    if (int rc = cgbwd_launch(p, st, dWeT)) return rc;   // per 64-channel chunk: kernel + partial sums into dWeT
  } else if (tc64) {
    int grid = 0;
    if (int rc = cgtc_launch(CG_BWD_DST, p, st, &grid, dq_atomic ? 1 : 0)) return rc;
    const int64_t tot = (int64_t)G * 2 * C;  // partial layout [G][2C] == dWeT layout (no channel chunking)
    if (int rc = sum_partials(p.dW_part, grid, tot, tot, dWeT, tot, nullptr, st)) return rc;
  } else {
    CgPlan plan;
    MDL_REQUIRE(cg_plan(CG_BWD_DST, C, G, &plan), "cgconv_bwd: C=%d G=%d does not fit", C, G);
    p.seg_ptr = dst_ptr; p.cap = plan.cap; p.te = plan.te;
    p.n_tiles = (int)std::max<int64_t>(1, ceil_div<int64_t>(E, plan.te));
    const int grid = cg_grid(p.n_tiles);
    for (int c_off = 0; c_off < C; c_off += plan.CC) {
      p.c_off = c_off; p.CC = std::min(plan.CC, C - c_off);
      size_t smem = cg_smem_bytes(CG_BWD_DST, p.CC, G, plan.cap);
      int rc;
      if (plan.nitem == 1) rc = cg_launch<CG_BWD_DST, 1>(p, smem, grid, st);
      else if (plan.nitem == 2) rc = cg_launch<CG_BWD_DST, 2>(p, smem, grid, st);
      else rc = cg_launch<CG_BWD_DST, 4>(p, smem, grid, st);
      if (rc) return rc;
      const int tot = G * 2 * p.CC;
      k_cg_reduce_dw<<<ceil_div(tot, 256), 256, 0, st>>>(p.dW_part, grid, G, C, c_off, p.CC, dWeT);
      MDL_LAUNCHED();
    }
  }
  // pass B: source order -> dQ
  if (dq_atomic) {
    // done inside pass A
  } else if (use_tc(CG_BWD_SRC, C, G)) {
    p.seg_ptr = src_ptr;
    if (int rc = cgtc_launch(CG_BWD_SRC, p, st, nullptr)) return rc;
  } else {
    CgPlan plan;
    MDL_REQUIRE(cg_plan(CG_BWD_SRC, C, G, &plan), "cgconv_bwd: C=%d G=%d does not fit", C, G);
    p.seg_ptr = src_ptr; p.cap = plan.cap; p.te = plan.te;
    p.n_tiles = (int)std::max<int64_t>(1, ceil_div<int64_t>(E, plan.te));
    const int grid = cg_grid(p.n_tiles);
    for (int c_off = 0; c_off < C; c_off += plan.CC) {
      p.c_off = c_off; p.CC = std::min(plan.CC, C - c_off);
      size_t smem = cg_smem_bytes(CG_BWD_SRC, p.CC, G, plan.cap);
      if (int rc = cg_launch<CG_BWD_SRC, 1>(p, smem, grid, st)) return rc;
    }
  }
  return MDL_OK;
}

// ---- smearing-fused form: the kernels take the normalised distance d_hat[E] (slot order) and expand the Gaussian
// basis themselves (reference GaussianSmearing, process.py:580-590, applied at process.py:500-502) -- 4 B per edge
// of HBM traffic per layer instead of 4 G.  Tensor-core kernels only (cgconv_fwd_ws.cu, cgconv_bwd.cu).
namespace mdl {
static bool smear_env_ok() {
  const char* a = getenv("MDL_CGCONV_IMPL");
  const char* b = getenv("MDL_CGCONV_BWD");
  const char* d = getenv("MDL_CGCONV_DETERMINISTIC");
  return !(a && a[0]) && !(b && b[0]) && !(d && d[0] == '1');
}
static CgParams smear_params(const float* PQ, const float* d_hat, const float* offset, float coeff, const float* WeT,
                             const int32_t* dst_ptr, const int32_t* dst_src, const int32_t* dst_dst,
                             const float* inv_deg, int reduce, int64_t N, int64_t E, int C, int G) {
  CgParams p{};
  p.PQ = PQ; p.dhat = d_hat; p.sm_offset = offset; p.sm_coeff = coeff; p.WeT = WeT; p.seg_ptr = dst_ptr;
  p.dst_src = dst_src; p.dst_dst = dst_dst; p.inv_deg = (reduce == MDL_REDUCE_MEAN) ? inv_deg : nullptr;
  p.N = (int)N; p.E = (int)E; p.C = C; p.G = G;
  return p;
}
}  // namespace mdl

extern "C" int mdl_cgconv_tc_supported(int32_t C, int32_t G) {
  static const float dummy = 0.0f;
  CgParams p{};
  p.ea = nullptr; p.C = C; p.G = G; p.N = 1; p.E = 1;
  (void)dummy;
  return (cgws_supported(p) ? 1 : 0) | (cgbwd_supported(p) ? 2 : 0);
}

extern "C" int mdl_cgconv_smear_supported(int32_t C, int32_t G) {
  if (!smear_env_ok() || G < 2) return 0;
  static const float dummy = 0.0f;
  CgParams p{};
  p.dhat = &dummy; p.C = C; p.G = G; p.N = 1; p.E = 1;
  return (cgws_supported(p) && cgbwd_supported(p)) ? 1 : 0;
}

extern "C" int mdl_cgconv_smear_fwd(const float* x, const float* PQ, const float* d_hat, const float* offset, float coeff,
                                    const float* WeT, const int32_t* dst_ptr, const int32_t* dst_src,
                                    const int32_t* dst_dst, const float* inv_deg_dst, float* out, int64_t N, int64_t E,
                                    int32_t C, int32_t G, int32_t reduce, void* stream) {
  if (int rc = cg_check(N, E, C, G, reduce)) return rc;
  MDL_REQUIRE(x && PQ && WeT && dst_ptr && out && offset && (E == 0 || (d_hat && dst_src && dst_dst)),
              "cgconv_smear_fwd: null pointer");
  MDL_REQUIRE(reduce != MDL_REDUCE_MEAN || inv_deg_dst, "cgconv_smear_fwd: mean needs inv_deg");
  static const float dummy = 0.0f;
  CgParams p = smear_params(PQ, d_hat ? d_hat : &dummy, offset, coeff, WeT, dst_ptr, dst_src, dst_dst, inv_deg_dst, reduce,
                            N, E, C, G);
  p.x = x; p.out = out;
  MDL_REQUIRE(smear_env_ok() && G >= 2 && cgws_supported(p),
              "cgconv_smear_fwd: C=%d G=%d (or the kernel switches in the environment) not supported by the fused form", C, G);
  return cgws_launch(p, as_stream(stream));
}

extern "C" int mdl_cgconv_smear_bwd(const float* gout, const float* PQ, const float* d_hat, const float* offset,
                                    float coeff, const float* WeT, const int32_t* dst_ptr, const int32_t* dst_src,
                                    const int32_t* dst_dst, const float* inv_deg_dst, float* dPQ, float* dWeT, int64_t N,
                                    int64_t E, int32_t C, int32_t G, int32_t reduce, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  if (int rc = cg_check(N, E, C, G, reduce)) return rc;
  MDL_REQUIRE(gout && PQ && WeT && dst_ptr && dPQ && dWeT && offset && (E == 0 || (d_hat && dst_src && dst_dst)),
              "cgconv_smear_bwd: null pointer");
  if (!workspace || workspace_bytes < mdl_cgconv_workspace_bytes(N, E, C, G)) {
    set_error("cgconv_smear_bwd: workspace %zu < %zu", workspace_bytes, mdl_cgconv_workspace_bytes(N, E, C, G));
    return MDL_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  static const float dummy = 0.0f;
  CgParams p = smear_params(PQ, d_hat ? d_hat : &dummy, offset, coeff, WeT, dst_ptr, dst_src, dst_dst, inv_deg_dst, reduce,
                            N, E, C, G);
  p.gout = gout; p.out = dPQ; p.dW_part = (float*)workspace;
  MDL_REQUIRE(smear_env_ok() && G >= 2 && cgbwd_supported(p),
              "cgconv_smear_bwd: C=%d G=%d (or the kernel switches in the environment) not supported by the fused form", C, G);
  // dQ[src] is accumulated with vector atomics inside the single pass: zero the Q half first (as mdl_cgconv_bwd)
  MDL_CUDA(cudaMemset2DAsync(dPQ + 2 * C, (size_t)4 * C * 4, 0, (size_t)2 * C * 4, (size_t)N, st));
  return cgbwd_launch(p, st, dWeT);
}

namespace mdl {
// lin_f / lin_s of the reference layer ([C, 2C+G] each, cat order x_i | x_j | e) -> the hoisted operands
__global__ void k_cgconv_pack(const float* __restrict__ w_f, const float* __restrict__ b_f,
                              const float* __restrict__ w_s, const float* __restrict__ b_s, int C, int G,
                              float* __restrict__ Wn, float* __restrict__ bias, float* __restrict__ WeT) {
  const int ld = 2 * C + G;
  const int n_wn = 4 * C * C, n_b = 4 * C, n_we = G * 2 * C;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_wn + n_b + n_we; t += gridDim.x * blockDim.x) {
    if (t < n_wn) {
      const int r = t / C, i = t - r * C, blk = r / C, c = r - blk * C;
      const float* w = (blk & 1) ? w_s : w_f;
      Wn[t] = __ldg(w + (size_t)c * ld + (blk >> 1) * C + i);
    } else if (t < n_wn + n_b) {
      const int r = t - n_wn, blk = r / C, c = r - blk * C;
      const float* b = (blk == 0) ? b_f : (blk == 1) ? b_s : nullptr;
      bias[r] = b ? __ldg(b + c) : 0.0f;
    } else {
      const int u = t - n_wn - n_b, g = u / (2 * C), col = u - g * 2 * C;
      const float* w = (col < C) ? w_f : w_s;
      WeT[u] = __ldg(w + (size_t)(col < C ? col : col - C) * ld + 2 * C + g);
    }
  }
}
}  // namespace mdl

extern "C" int mdl_cgconv_pack_weights(const float* w_f, const float* b_f, const float* w_s, const float* b_s,
                                       int32_t C, int32_t G, float* Wn, float* bias, float* WeT, void* stream) {
  MDL_REQUIRE(C > 0 && G > 0 && w_f && w_s && Wn && bias && WeT, "cgconv_pack_weights: bad arguments");
  const int total = 4 * C * C + 4 * C + G * 2 * C;
  k_cgconv_pack<<<std::min(ceil_div(total, 256), 2 * kNumSMs), 256, 0, as_stream(stream)>>>(w_f, b_f, w_s, b_s, C, G,
                                                                                             Wn, bias, WeT);
  MDL_LAUNCHED();
  return MDL_OK;
}

// development aid: 32 x uint64 device counters receiving per-phase cycle sums of the
// tensor-core kernels (NULL disables).  Not part of the reference-facing surface.
extern "C" MDL_API int mdl_debug_set_phase_buffer(void* dev_ptr) {
  cgtc_set_phase_buffer(reinterpret_cast<unsigned long long*>(dev_ptr));
  cgfwd_set_phase_buffer(reinterpret_cast<unsigned long long*>(dev_ptr));
  cgbwd_set_phase_buffer(reinterpret_cast<unsigned long long*>(dev_ptr));
  cgws_set_phase_buffer(reinterpret_cast<unsigned long long*>(dev_ptr));
  lin_set_phase_buffer(reinterpret_cast<unsigned long long*>(dev_ptr));
  return MDL_OK;
}
