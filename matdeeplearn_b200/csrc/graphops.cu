// graphops.cu -- CSR graph primitives behind the SchNet / MPNN / MEGNet operator surface
// (include/mdl_b200.h).  All are deterministic (segment-ordered sums, no atomics) and keep the
// reference's edge order at the interface: `eid` arrays translate a CSR position to the
// reference edge id, so edge-level tensors never need to be permuted in HBM.
//
//   spmm_edge      out[i]      = sum_{p in seg(i)} h[nbr(p)] (*) w[eid(p)]      CFConv aggregate
//                                                                               (schnet.py:140 -> PyG CFConv)
//   edge_mul       out[eid(p)] = a[ia(p)] (*) b[ib(p)]                           its filter gradient
//   edge_gather_add out[e]     = act(base[e] + A[ia[e]] + B[ib[e]] + U[ig[e]] + bias)
//                                                                               Megnet_EdgeModel's first
//                                                                               Linear on cat[...] (megnet.py:41-47)
//   nnconv_msg     m[eid(p)]   = sum_k hid[eid(p),k] * XT[j,k,:] + XB[j,:]       NNConv message, re-associated
//                                                                               (mpnn.py:83-88 -> PyG NNConv)
#include "common.cuh"

namespace mdl {

// ---- out[i,:] = sum_{p in [ptr[i],ptr[i+1])} h[nbr[p],:] * w[eid[p],:]   (warp per segment)
__global__ void __launch_bounds__(256)
k_spmm_edge(const float* __restrict__ h, const float* __restrict__ w, const int32_t* __restrict__ ptr,
            const int32_t* __restrict__ nbr, const int32_t* __restrict__ eid, float* __restrict__ out,
            int64_t S, int width) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nv = width >> 2;  // float4 per row
  for (int64_t s = warp; s < S; s += nwarps) {
    const int lo = __ldg(ptr + s), hi = __ldg(ptr + s + 1);
    for (int v0 = 0; v0 < nv; v0 += 32) {
      const int v = v0 + lane;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < nv) {
#pragma unroll 4
        for (int p = lo; p < hi; ++p) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(h + (size_t)(nbr ? __ldg(nbr + p) : p) * width) + v);
          const float4 b = __ldg(reinterpret_cast<const float4*>(w + (size_t)(eid ? __ldg(eid + p) : p) * width) + v);
          acc.x = fmaf(a.x, b.x, acc.x); acc.y = fmaf(a.y, b.y, acc.y);
          acc.z = fmaf(a.z, b.z, acc.z); acc.w = fmaf(a.w, b.w, acc.w);
        }
        *(reinterpret_cast<float4*>(out + (size_t)s * width) + v) = acc;
      }
    }
  }
}

// ---- out[i,:] = sum_{p in [ptr[i],ptr[i+1])} c[eid[p]] * h[nbr[p],:]   (warp per segment; one coefficient per edge:
// GCNConv's normalised adjacency, reference gcn.py:80-82,141)
__global__ void __launch_bounds__(256)
k_spmm_edge_scalar(const float* __restrict__ h, const float* __restrict__ c, const int32_t* __restrict__ ptr,
                   const int32_t* __restrict__ nbr, const int32_t* __restrict__ eid, float* __restrict__ out,
                   int64_t S, int width) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nv = width >> 2;
  for (int64_t s = warp; s < S; s += nwarps) {
    const int lo = __ldg(ptr + s), hi = __ldg(ptr + s + 1);
    for (int v0 = 0; v0 < nv; v0 += 32) {
      const int v = v0 + lane;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < nv) {
#pragma unroll 4
        for (int p = lo; p < hi; ++p) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(h + (size_t)(nbr ? __ldg(nbr + p) : p) * width) + v);
          const float w = __ldg(c + (eid ? __ldg(eid + p) : p));
          acc.x = fmaf(a.x, w, acc.x); acc.y = fmaf(a.y, w, acc.y);
          acc.z = fmaf(a.z, w, acc.z); acc.w = fmaf(a.w, w, acc.w);
        }
        *(reinterpret_cast<float4*>(out + (size_t)s * width) + v) = acc;
      }
    }
  }
}

// ---- out[eid[p]] = < a[ia[p],:], b[ib[p],:] >   (warp per position, fixed-order lane sums + shuffle tree)
__global__ void __launch_bounds__(256)
k_edge_dot(const float* __restrict__ a, const float* __restrict__ b, const int32_t* __restrict__ ia,
           const int32_t* __restrict__ ib, const int32_t* __restrict__ eid, float* __restrict__ out,
           int64_t E, int width) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nv = width >> 2;
  for (int64_t p = warp; p < E; p += nwarps) {
    const float4* ra = reinterpret_cast<const float4*>(a + (size_t)__ldg(ia + p) * width);
    const float4* rb = reinterpret_cast<const float4*>(b + (size_t)__ldg(ib + p) * width);
    float acc = 0.0f;
    for (int v = lane; v < nv; v += 32) {
      const float4 x = __ldg(ra + v), y = __ldg(rb + v);
      acc = fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, fmaf(x.w, y.w, acc))));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[eid ? __ldg(eid + p) : p] = acc;
  }
}

// ---- out[eid[p],:] = a[ia[p],:] * b[ib[p],:]   (warp per position)
__global__ void __launch_bounds__(256)
k_edge_mul(const float* __restrict__ a, const float* __restrict__ b, const int32_t* __restrict__ ia,
           const int32_t* __restrict__ ib, const int32_t* __restrict__ eid, float* __restrict__ out,
           int64_t E, int width) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nv = width >> 2;
  for (int64_t p = warp; p < E; p += nwarps) {
    const float4* ra = reinterpret_cast<const float4*>(a + (size_t)__ldg(ia + p) * width);
    const float4* rb = reinterpret_cast<const float4*>(b + (size_t)__ldg(ib + p) * width);
    float4* ro = reinterpret_cast<float4*>(out + (size_t)(eid ? __ldg(eid + p) : p) * width);
    for (int v = lane; v < nv; v += 32) {
      const float4 x = __ldg(ra + v), y = __ldg(rb + v);
      ro[v] = make_float4(x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w);
    }
  }
}

// ---- out[e,:] = act(base[e,:] + A[ia[e],:] + B[ib[e],:] + U[ig[e],:] + bias)   (warp per edge)
__global__ void __launch_bounds__(256)
k_edge_gather_add(const float* __restrict__ base, const float* __restrict__ A, const float* __restrict__ B,
                  const float* __restrict__ U, const int64_t* __restrict__ ia, const int64_t* __restrict__ ib,
                  const int64_t* __restrict__ ig, const float* __restrict__ bias, float* __restrict__ out,
                  int64_t E, int width, int relu) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nv = width >> 2;
  for (int64_t e = warp; e < E; e += nwarps) {
    const int64_t ja = ia[e], jb = ib[e];
    const int64_t jg = U ? (ig ? ig[ja] : 0) : 0;  // graph of the edge = graph of its source node
    for (int v = lane; v < nv; v += 32) {
      float4 s = __ldg(reinterpret_cast<const float4*>(base + (size_t)e * width) + v);
      const float4 x = __ldg(reinterpret_cast<const float4*>(A + (size_t)ja * width) + v);
      const float4 y = __ldg(reinterpret_cast<const float4*>(B + (size_t)jb * width) + v);
      s.x += x.x + y.x; s.y += x.y + y.y; s.z += x.z + y.z; s.w += x.w + y.w;
      if (U) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(U + (size_t)jg * width) + v);
        s.x += u.x; s.y += u.y; s.z += u.z; s.w += u.w;
      }
      if (bias) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + v);
        s.x += bb.x; s.y += bb.y; s.z += bb.z; s.w += bb.w;
      }
      if (relu) { s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f); }
      *(reinterpret_cast<float4*>(out + (size_t)e * width) + v) = s;
    }
  }
}

// ---- NNConv message, by-source order.  One CTA per source node j: XT[j] ([K,O], the node's
// features already contracted with the edge-network's output weights) is staged in smem once
// and reused by every out-edge of j.
//   fwd: m[eid(p), o] = sum_k hid[eid(p), k] * XT[j, k, o] + XB[j, o]
//   bwd: dhid[eid(p), k] = sum_o XT[j, k, o] * dm[eid(p), o]
//        dXT[j, k, o]    = sum_p hid[eid(p), k] * dm[eid(p), o]
//        dXB[j, o]       = sum_p dm[eid(p), o]
constexpr int kNnChunk = 16;   // edges staged per pass in the backward kernel
constexpr int kNnMaxAcc = 40;  // dXT entries per thread (K*O <= 40*256)

template <bool BWD>
__global__ void __launch_bounds__(256)
k_nnconv_msg(const float* __restrict__ hid, const float* __restrict__ XT, const float* __restrict__ XB,
             const float* __restrict__ dm, const int32_t* __restrict__ ptr, const int32_t* __restrict__ eid,
             float* __restrict__ m, float* __restrict__ dhid, float* __restrict__ dXT,
             float* __restrict__ dXB, int64_t N, int K, int O) {
  extern __shared__ __align__(16) float sm[];
  const int OS = O + 1;                     // padded row: conflict-free both along k and along o
  float* sXT = sm;                          // [K][OS]
  float* sH = sm + K * OS;                  // [kNnChunk][K]   hidden rows of the staged edges
  float* sD = sH + kNnChunk * K;            // [kNnChunk][O]   dm rows (bwd)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int KO = K * O;
  for (int64_t j = blockIdx.x; j < N; j += gridDim.x) {
    const int lo = __ldg(ptr + j), hi = __ldg(ptr + j + 1);
    __syncthreads();
    for (int i = tid; i < KO; i += blockDim.x) {
      const int k = i / O, o = i - k * O;
      sXT[k * OS + o] = __ldg(XT + (size_t)j * KO + i);
    }
    float acc[kNnMaxAcc];
    if (BWD) {
#pragma unroll
      for (int r = 0; r < kNnMaxAcc; ++r) acc[r] = 0.0f;
    }
    float accb = 0.0f;  // dXB entry of thread tid (tid < O)
    for (int c0 = lo; c0 < hi; c0 += kNnChunk) {
      const int nc = min(kNnChunk, hi - c0);
      __syncthreads();
      for (int i = tid; i < nc * K; i += blockDim.x) {
        const int r = i / K, k = i - r * K;
        sH[r * K + k] = __ldg(hid + (size_t)__ldg(eid + c0 + r) * K + k);
      }
      if (BWD) {
        for (int i = tid; i < nc * O; i += blockDim.x) {
          const int r = i / O, o = i - r * O;
          sD[r * O + o] = __ldg(dm + (size_t)__ldg(eid + c0 + r) * O + o);
        }
      }
      __syncthreads();
      if (!BWD) {
        for (int r = warp; r < nc; r += 8) {
          const size_t e = (size_t)__ldg(eid + c0 + r);
          for (int o = lane; o < O; o += 32) {
            float a = __ldg(XB + (size_t)j * O + o);
            for (int k = 0; k < K; ++k) a = fmaf(sH[r * K + k], sXT[k * OS + o], a);
            m[e * O + o] = a;
          }
        }
      } else {
        for (int r = warp; r < nc; r += 8) {  // dhid: lanes over k
          const size_t e = (size_t)__ldg(eid + c0 + r);
          for (int k = lane; k < K; k += 32) {
            float a = 0.0f;
            for (int o = 0; o < O; ++o) a = fmaf(sXT[k * OS + o], sD[r * O + o], a);
            dhid[e * K + k] = a;
          }
        }
#pragma unroll
        for (int rr = 0; rr < kNnMaxAcc; ++rr) {  // dXT entries owned by this thread
          const int i = tid + rr * 256;
          if (i < KO) {
            const int k = i / O, o = i - k * O;
            float a = acc[rr];
            for (int r = 0; r < nc; ++r) a = fmaf(sH[r * K + k], sD[r * O + o], a);
            acc[rr] = a;
          }
        }
        if (tid < O)
          for (int r = 0; r < nc; ++r) accb += sD[r * O + tid];
      }
    }
    if (BWD) {
#pragma unroll
      for (int rr = 0; rr < kNnMaxAcc; ++rr) {
        const int i = tid + rr * 256;
        if (i < KO) dXT[(size_t)j * KO + i] = acc[rr];
      }
      for (int o = tid; o < O; o += blockDim.x) {
        // O > 256 never happens for the supported sizes; accb covers tid < O <= 256
        dXB[(size_t)j * O + o] = accb;
      }
    }
  }
}

// ---- NNConv message, register-tiled: one WARP per source node j, lane = output pair (2l, 2l+1), the node's
// out-edges in chunks of 16 whose hidden rows sit transposed in shared memory (sHT[k][16]: one broadcast LDS.128 feeds
// four edges), XT[j] streamed once per chunk with coalesced 8-byte loads: 4 LDS + 1 LDG per 32 FMAs instead of the
// 2 LDS per FMA of k_nnconv_msg above.  K, O <= 64, O even (MPNN: K = dim3, O = gc_dim; reference mpnn.py:83-88).
//   fwd: m[e, o] = XB[j, o] + sum_k hid[e, k] XT[j, k, o]
//   bwd: dXT[j, k, o] = sum_e hid[e, k] dm[e, o];  dXB[j, o] = sum_e dm[e, o];  dhid[e, k] = sum_o XT[j, k, o] dm[e, o]
//        (dhid: lane = k and k + 32, XT[j] staged per warp with a padded stride, dm rows broadcast from shared memory)
constexpr int kNw = 8;          // warps (= nodes in flight) per CTA
constexpr int kNwChunk = 16;    // out-edges per pass
constexpr int kHS = 20;         // row stride of the transposed hidden tile (16 + 4: 16-byte aligned rows, 4-way instead of
                                // 16-way bank conflicts when lanes = k write it)

template <bool BWD>
__global__ void __launch_bounds__(kNw * 32)
k_nnconv_msg_w(const float* __restrict__ hid, const float* __restrict__ XT, const float* __restrict__ XB,
               const float* __restrict__ dm, const int32_t* __restrict__ ptr, const int32_t* __restrict__ eid,
               float* __restrict__ m, float* __restrict__ dhid, float* __restrict__ dXT, float* __restrict__ dXB,
               int64_t N, int K, int O) {
  extern __shared__ __align__(16) float smw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int OS = O + 1;
  const int per_warp = kHS * 64 + (BWD ? kNwChunk * 64 + 64 * 65 : 0);
  float* sHT = smw + (size_t)warp * per_warp;   // [K][kHS] hidden rows of the chunk, transposed
  float* sD = sHT + kHS * 64;                   // [16][O]  dm rows of the chunk (bwd)
  float* sXT = sD + kNwChunk * 64;              // [K][O+1] XT[j] (bwd, for dhid)
  const int KO = K * O;
  const bool own = 2 * lane < O;                // this lane owns outputs 2 lane, 2 lane + 1
  for (int64_t j = (int64_t)blockIdx.x * kNw + warp; j < N; j += (int64_t)gridDim.x * kNw) {
    const int lo = __ldg(ptr + j), hi = __ldg(ptr + j + 1);
    const float* xt = XT + (size_t)j * KO;
    if (BWD) {
      __syncwarp();
      for (int i = lane; i < KO; i += 32) {       // coalesced read, padded rows: column reads are conflict-free
        const int k = i / O, o = i - k * O;
        sXT[k * OS + o] = __ldg(xt + i);
      }
    }
    float2 bsum = make_float2(0.0f, 0.0f);
    for (int c0 = lo; c0 < hi || (BWD && c0 == lo); c0 += kNwChunk) {   // bwd: a node without out-edges still writes zeros
      const int nc = max(0, min(kNwChunk, hi - c0));
      __syncwarp();
      // ---- stage the chunk: hidden rows transposed (zero rows beyond nc), dm rows (bwd)
      for (int r = 0; r < kNwChunk; ++r) {
        const size_t e = r < nc ? (size_t)__ldg(eid + c0 + r) : 0;
        for (int k = lane; k < K; k += 32) sHT[k * kHS + r] = r < nc ? __ldg(hid + e * K + k) : 0.0f;
        if (BWD)
          for (int o = lane; o < O; o += 32) sD[r * O + o] = r < nc ? __ldg(dm + e * O + o) : 0.0f;
      }
      __syncwarp();
      if (!BWD) {
        float2 acc[kNwChunk];
        const float2 b = own ? __ldg(reinterpret_cast<const float2*>(XB + (size_t)j * O) + lane) : make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kNwChunk; ++r) acc[r] = b;
#pragma unroll 4   // four rows of XT[j] requested before the first is used
        for (int k = 0; k < K; ++k) {
          const float2 x = own ? __ldg(reinterpret_cast<const float2*>(xt + (size_t)k * O) + lane) : make_float2(0.f, 0.f);
          const float4* hp = reinterpret_cast<const float4*>(sHT + k * kHS);
#pragma unroll
          for (int q = 0; q < kNwChunk / 4; ++q) {
            const float4 h = hp[q];
            acc[4 * q + 0].x = fmaf(h.x, x.x, acc[4 * q + 0].x); acc[4 * q + 0].y = fmaf(h.x, x.y, acc[4 * q + 0].y);
            acc[4 * q + 1].x = fmaf(h.y, x.x, acc[4 * q + 1].x); acc[4 * q + 1].y = fmaf(h.y, x.y, acc[4 * q + 1].y);
            acc[4 * q + 2].x = fmaf(h.z, x.x, acc[4 * q + 2].x); acc[4 * q + 2].y = fmaf(h.z, x.y, acc[4 * q + 2].y);
            acc[4 * q + 3].x = fmaf(h.w, x.x, acc[4 * q + 3].x); acc[4 * q + 3].y = fmaf(h.w, x.y, acc[4 * q + 3].y);
          }
        }
        if (own) {
#pragma unroll
          for (int r = 0; r < kNwChunk; ++r)
            if (r < nc) reinterpret_cast<float2*>(m + (size_t)__ldg(eid + c0 + r) * O)[lane] = acc[r];
        }
      } else {
        // ---- dXT[j, k, 2l..] (+)= sum_r hid[r, k] dm[r, 2l..];  dXB
        float2 d[kNwChunk];
#pragma unroll
        for (int r = 0; r < kNwChunk; ++r) {
          d[r] = own ? *(reinterpret_cast<const float2*>(sD + r * O) + lane) : make_float2(0.f, 0.f);
          bsum.x += d[r].x; bsum.y += d[r].y;
        }
        const bool first = c0 == lo;
        for (int k = 0; k < K; ++k) {
          const float4* hp = reinterpret_cast<const float4*>(sHT + k * kHS);
          float2* dst = reinterpret_cast<float2*>(dXT + (size_t)j * KO + (size_t)k * O) + lane;
          float2 a = (first || !own) ? make_float2(0.f, 0.f) : *dst;
#pragma unroll
          for (int q = 0; q < kNwChunk / 4; ++q) {
            const float4 h = hp[q];
            a.x = fmaf(h.x, d[4 * q + 0].x, a.x); a.y = fmaf(h.x, d[4 * q + 0].y, a.y);
            a.x = fmaf(h.y, d[4 * q + 1].x, a.x); a.y = fmaf(h.y, d[4 * q + 1].y, a.y);
            a.x = fmaf(h.z, d[4 * q + 2].x, a.x); a.y = fmaf(h.z, d[4 * q + 2].y, a.y);
            a.x = fmaf(h.w, d[4 * q + 3].x, a.x); a.y = fmaf(h.w, d[4 * q + 3].y, a.y);
          }
          if (own) *dst = a;
        }
        // ---- dhid[e_r, k] = sum_o XT[j, k, o] dm[r, o], lane = k (and k + 32)
        float g0[kNwChunk], g1[kNwChunk];
#pragma unroll
        for (int r = 0; r < kNwChunk; ++r) { g0[r] = 0.0f; g1[r] = 0.0f; }
        const int k0 = lane, k1 = lane + 32;
        const float* x0p = sXT + (k0 < K ? k0 : 0) * OS;
        const float* x1p = sXT + (k1 < K ? k1 : 0) * OS;
        for (int o = 0; o < O; o += 4) {          // O % 4 == 0: one broadcast LDS.128 of dm per edge and four outputs
          const float a0 = x0p[o], a1 = x0p[o + 1], a2 = x0p[o + 2], a3 = x0p[o + 3];
          const float b0 = x1p[o], b1 = x1p[o + 1], b2 = x1p[o + 2], b3 = x1p[o + 3];
#pragma unroll
          for (int r = 0; r < kNwChunk; ++r) {
            const float4 dv = *reinterpret_cast<const float4*>(sD + r * O + o);
            g0[r] = fmaf(a3, dv.w, fmaf(a2, dv.z, fmaf(a1, dv.y, fmaf(a0, dv.x, g0[r]))));
            g1[r] = fmaf(b3, dv.w, fmaf(b2, dv.z, fmaf(b1, dv.y, fmaf(b0, dv.x, g1[r]))));
          }
        }
#pragma unroll
        for (int r = 0; r < kNwChunk; ++r)
          if (r < nc) {
            float* row = dhid + (size_t)__ldg(eid + c0 + r) * K;
            if (k0 < K) row[k0] = g0[r];
            if (k1 < K) row[k1] = g1[r];
          }
      }
    }
    if (BWD && own) reinterpret_cast<float2*>(dXB + (size_t)j * O)[lane] = bsum;
  }
}

static bool nnconv_w_ok(int K, int O) { return K >= 1 && K <= 64 && O >= 4 && O <= 64 && (O & 3) == 0; }

static int warp_grid(int64_t items) {
  int64_t blocks = ceil_div<int64_t>(items, 8);
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  return (int)(blocks > 0 ? blocks : 1);
}

}  // namespace mdl

using namespace mdl;

extern "C" int mdl_spmm_edge(const float* h, const float* w, const int32_t* ptr, const int32_t* nbr,
                             const int32_t* eid, float* out, int64_t S, int64_t width, void* stream) {
  MDL_REQUIRE(S >= 0 && width > 0 && width % 4 == 0, "spmm_edge: width must be a multiple of 4");
  if (S == 0) return MDL_OK;
  MDL_REQUIRE(h && w && ptr && out, "spmm_edge: null pointer");
  k_spmm_edge<<<warp_grid(S), 256, 0, as_stream(stream)>>>(h, w, ptr, nbr, eid, out, S, (int)width);
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_spmm_edge_scalar(const float* h, const float* coef, const int32_t* ptr, const int32_t* nbr,
                                    const int32_t* eid, float* out, int64_t S, int64_t width, void* stream) {
  MDL_REQUIRE(S >= 0 && width > 0 && width % 4 == 0, "spmm_edge_scalar: width must be a multiple of 4");
  if (S == 0) return MDL_OK;
  MDL_REQUIRE(h && coef && ptr && out, "spmm_edge_scalar: null pointer");
  k_spmm_edge_scalar<<<warp_grid(S), 256, 0, as_stream(stream)>>>(h, coef, ptr, nbr, eid, out, S, (int)width);
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_edge_dot(const float* a, const float* b, const int32_t* ia, const int32_t* ib, const int32_t* eid,
                            float* out, int64_t E, int64_t width, void* stream) {
  MDL_REQUIRE(E >= 0 && width > 0 && width % 4 == 0, "edge_dot: width must be a multiple of 4");
  if (E == 0) return MDL_OK;
  MDL_REQUIRE(a && b && ia && ib && out, "edge_dot: null pointer");
  k_edge_dot<<<warp_grid(E), 256, 0, as_stream(stream)>>>(a, b, ia, ib, eid, out, E, (int)width);
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_edge_mul(const float* a, const float* b, const int32_t* ia, const int32_t* ib,
                            const int32_t* eid, float* out, int64_t E, int64_t width, void* stream) {
  MDL_REQUIRE(E >= 0 && width > 0 && width % 4 == 0, "edge_mul: width must be a multiple of 4");
  if (E == 0) return MDL_OK;
  MDL_REQUIRE(a && b && ia && ib && out, "edge_mul: null pointer");
  k_edge_mul<<<warp_grid(E), 256, 0, as_stream(stream)>>>(a, b, ia, ib, eid, out, E, (int)width);
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_edge_gather_add(const float* base, const float* A, const float* B, const float* U,
                                   const int64_t* ia, const int64_t* ib, const int64_t* ig,
                                   const float* bias, float* out, int64_t E, int64_t width,
                                   int32_t relu, void* stream) {
  MDL_REQUIRE(E >= 0 && width > 0 && width % 4 == 0, "edge_gather_add: width must be a multiple of 4");
  if (E == 0) return MDL_OK;
  MDL_REQUIRE(base && A && B && ia && ib && out, "edge_gather_add: null pointer");
  k_edge_gather_add<<<warp_grid(E), 256, 0, as_stream(stream)>>>(base, A, B, U, ia, ib, ig, bias, out, E,
                                                                   (int)width, relu);
  MDL_LAUNCHED();
  return MDL_OK;
}

static int nnconv_launch(bool bwd, const float* hid, const float* XT, const float* XB, const float* dm,
                         const int32_t* ptr, const int32_t* eid, float* m, float* dhid, float* dXT,
                         float* dXB, int64_t N, int32_t K, int32_t O, void* stream) {
  MDL_REQUIRE(N >= 0 && K > 0 && O > 0, "nnconv_msg: bad shape");
  if (N == 0) return MDL_OK;
  {  // register-tiled warp-per-node kernels (K, O <= 64); MDL_NNCONV=cta keeps the CTA-per-node ones (A/B)
    const char* env = getenv("MDL_NNCONV");
    if (nnconv_w_ok(K, O) && !(env && strcmp(env, "cta") == 0)) {
      const size_t smw_bytes = (size_t)kNw * (kHS * 64 + (bwd ? kNwChunk * 64 + 64 * 65 : 0)) * sizeof(float);
      const int grid = (int)std::min<int64_t>(ceil_div<int64_t>(N, kNw), (int64_t)kNumSMs * 4);
      if (bwd) {
        MDL_CUDA(cudaFuncSetAttribute(k_nnconv_msg_w<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smw_bytes));
        k_nnconv_msg_w<true><<<grid, kNw * 32, smw_bytes, as_stream(stream)>>>(hid, XT, XB, dm, ptr, eid, m, dhid, dXT, dXB, N, K, O);
      } else {
        k_nnconv_msg_w<false><<<grid, kNw * 32, smw_bytes, as_stream(stream)>>>(hid, XT, XB, dm, ptr, eid, m, dhid, dXT, dXB, N, K, O);
      }
      MDL_LAUNCHED();
      return MDL_OK;
    }
  }
  size_t smem = ((size_t)K * (O + 1) + (size_t)kNnChunk * (K + O)) * 4;
  MDL_REQUIRE(smem <= 200 * 1024 && (int64_t)K * O <= (int64_t)kNnMaxAcc * 256 && O <= 256,
              "nnconv_msg: hidden %d x out %d outside the supported range", K, O);
  int grid = (int)std::min<int64_t>(N, (int64_t)kNumSMs * 4);
  if (bwd) {
    MDL_CUDA(cudaFuncSetAttribute(k_nnconv_msg<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_nnconv_msg<true><<<grid, 256, smem, as_stream(stream)>>>(hid, XT, XB, dm, ptr, eid, m, dhid, dXT, dXB, N, K, O);
  } else {
    MDL_CUDA(cudaFuncSetAttribute(k_nnconv_msg<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_nnconv_msg<false><<<grid, 256, smem, as_stream(stream)>>>(hid, XT, XB, dm, ptr, eid, m, dhid, dXT, dXB, N, K, O);
  }
  MDL_LAUNCHED();
  return MDL_OK;
}

extern "C" int mdl_nnconv_msg_fwd(const float* hid, const float* XT, const float* XB, const int32_t* src_ptr,
                                  const int32_t* src_eid, float* m, int64_t N, int32_t K, int32_t O,
                                  void* stream) {
  MDL_REQUIRE(hid && XT && XB && src_ptr && src_eid && m, "nnconv_msg_fwd: null pointer");
  return nnconv_launch(false, hid, XT, XB, nullptr, src_ptr, src_eid, m, nullptr, nullptr, nullptr, N, K, O, stream);
}

extern "C" int mdl_nnconv_msg_bwd(const float* hid, const float* XT, const float* dm, const int32_t* src_ptr,
                                  const int32_t* src_eid, float* dhid, float* dXT, float* dXB, int64_t N,
                                  int32_t K, int32_t O, void* stream) {
  MDL_REQUIRE(hid && XT && dm && src_ptr && src_eid && dhid && dXT && dXB, "nnconv_msg_bwd: null pointer");
  return nnconv_launch(true, hid, XT, nullptr, dm, src_ptr, src_eid, nullptr, dhid, dXT, dXB, N, K, O, stream);
}
