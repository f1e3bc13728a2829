// assemble.cu -- batch assembly from a device-resident dataset ("graph store").
//
// Replaces, on the reference's training path, the per-step CPU collate of the PyG DataLoader
// (training/training.py:300-307: Batch.from_data_list -> cat of every per-graph tensor,
// edge_index shifted by the running node count) and the `data.to(rank)` copy (training.py:39).
//
// The whole processed dataset sits in HBM as ONE block-diagonal graph: node rows, edges in the
// reference's order, and the destination-major layout (mdl_csr_from_coo) built once over all of
// it.  Graphs are contiguous in every one of those arrays, and shifting a graph by a node/edge
// offset keeps both orders, so a batch is a pure gather: no sort, no scan, no atomics.
//
//   grid = (B, P): CTA (b, p) copies part p of graph graph_ids[b].
#include "common.cuh"

namespace mdl {

struct AsmArgs {
  mdl_graph_store s;
  mdl_batch_out o;
};

template <typename T>
__device__ __forceinline__ void copy_span(T* __restrict__ dst, const T* __restrict__ src, int64_t n,
                                          int tid, int stride) {
  for (int64_t i = tid; i < n; i += stride) dst[i] = __ldg(src + i);
}

__device__ __forceinline__ void copy_shift_i32(int32_t* __restrict__ dst, const int32_t* __restrict__ src,
                                               int64_t n, int32_t shift, int tid, int stride) {
  for (int64_t i = tid; i < n; i += stride) dst[i] = __ldg(src + i) + shift;
}

template <typename T>
__device__ __forceinline__ void fill_span(T* __restrict__ dst, T v, int64_t n, int tid, int stride) {
  for (int64_t i = tid; i < n; i += stride) dst[i] = v;
}

// Rows [Nv, N) and edges/slots [Ev, E) of a capacity-padded batch: inert filler.  No segment of either
// pointer array covers a padded slot (dst_ptr/src_ptr stay at Ev for every padded node), so the
// segment-driven operators never touch them; padded nodes are empty segments outside every graph.
__device__ void fill_padding(const mdl_batch_out& o, int F, int G, int64_t Nv, int64_t Ev, int tid, int stride) {
  const int64_t np = o.N - Nv, ne = o.E - Ev;
  const int32_t last = (int32_t)(o.N - 1);
  fill_span(o.x + Nv * F, 0.0f, np * F, tid, stride);
  fill_span(o.batch + Nv, (int64_t)o.B, np, tid, stride);
  fill_span(o.edge_index + Ev, (int64_t)last, ne, tid, stride);
  fill_span(o.edge_index + o.E + Ev, (int64_t)last, ne, tid, stride);
  if (o.d_hat) fill_span(o.d_hat + Ev, 0.0f, ne, tid, stride);
  fill_span(o.edge_weight + Ev, 0.0f, ne, tid, stride);
  if (o.edge_attr) fill_span(o.edge_attr + Ev * G, 0.0f, ne * G, tid, stride);
  if (o.edge_attr_slots) fill_span(o.edge_attr_slots + Ev * G, 0.0f, ne * G, tid, stride);
  if (o.dst_ptr) {
    fill_span(o.dst_ptr + Nv, (int32_t)Ev, np + 1, tid, stride);
    fill_span(o.src_ptr + Nv, (int32_t)Ev, np + 1, tid, stride);
    fill_span(o.inv_deg_dst + Nv, 0.0f, np, tid, stride);
    fill_span(o.inv_deg_src + Nv, 0.0f, np, tid, stride);
    fill_span(o.dst_src + Ev, last, ne, tid, stride);
    fill_span(o.dst_dst + Ev, last, ne, tid, stride);
    for (int64_t i = tid; i < ne; i += stride) {
      o.dst_eid[Ev + i] = (int32_t)(Ev + i);
      o.src_slot[Ev + i] = (int32_t)(Ev + i);
    }
    if (tid == 0) o.graph_ptr[o.B] = (int32_t)Nv;
  }
}

__global__ void __launch_bounds__(256) k_assemble(const AsmArgs a) {
  const mdl_graph_store& s = a.s;
  const mdl_batch_out& o = a.o;
  const int b = blockIdx.x;
  const int tid = blockIdx.y * blockDim.x + threadIdx.x;
  const int stride = gridDim.y * blockDim.x;
  if (b == o.B) {   // the extra CTA row: padding and the closing pointer entries
    fill_padding(o, s.F, s.G, __ldg(o.node_off + o.B), __ldg(o.edge_off + o.B), tid, stride);
    return;
  }
  const int64_t g = __ldg(o.graph_ids + b);
  const int64_t np = __ldg(s.node_ptr + g), n = __ldg(s.node_ptr + g + 1) - np;
  const int64_t ep = __ldg(s.edge_ptr + g), e = __ldg(s.edge_ptr + g + 1) - ep;
  const int64_t no = __ldg(o.node_off + b), eo = __ldg(o.edge_off + b);
  const int32_t nshift = (int32_t)(no - np), eshift = (int32_t)(eo - ep);
  const int F = s.F, G = s.G;

  // ---- node-level rows
  copy_span(o.x + no * F, s.x + np * F, n * F, tid, stride);
  for (int64_t i = tid; i < n; i += stride) o.batch[no + i] = b;
  if (o.dst_ptr) {
    copy_shift_i32(o.dst_ptr + no, s.dst_ptr + np, n, eshift, tid, stride);
    copy_shift_i32(o.src_ptr + no, s.src_ptr + np, n, eshift, tid, stride);
    copy_span(o.inv_deg_dst + no, s.inv_deg_dst + np, n, tid, stride);
    copy_span(o.inv_deg_src + no, s.inv_deg_src + np, n, tid, stride);
  }
  // ---- graph-level rows
  if (tid < s.U) o.u[(int64_t)b * s.U + tid] = __ldg(s.u + g * s.U + tid);
  if (tid < s.Y) o.y[(int64_t)b * s.Y + tid] = __ldg(s.y + g * s.Y + tid);
  if (tid == 0 && o.graph_ptr) o.graph_ptr[b] = (int32_t)no;
  // ---- edges, reference order
  for (int64_t k = tid; k < e; k += stride) {
    o.edge_index[eo + k] = (int64_t)__ldg(s.src + ep + k) + (no - np);
    o.edge_index[o.E + eo + k] = (int64_t)__ldg(s.dst + ep + k) + (no - np);
  }
  if (o.d_hat) copy_span(o.d_hat + eo, s.d_hat + ep, e, tid, stride);
  copy_span(o.edge_weight + eo, s.edge_weight + ep, e, tid, stride);
  // ---- edges, destination-major slots
  if (o.dst_ptr) {
    copy_shift_i32(o.dst_src + eo, s.dst_src + ep, e, nshift, tid, stride);
    copy_shift_i32(o.dst_dst + eo, s.dst_dst + ep, e, nshift, tid, stride);
    copy_shift_i32(o.dst_eid + eo, s.dst_eid + ep, e, eshift, tid, stride);
    copy_shift_i32(o.src_slot + eo, s.src_slot + ep, e, eshift, tid, stride);
  }
  // ---- edge_attr: copied when the store holds it, else GaussianSmearing(d_hat) on the fly
  //      (same expression as k_gaussian_smear: expf(coeff * (diff*diff)), reference process.py:588-590).
  //      One warp per edge row, lanes across the G basis functions: no per-element division, the
  //      row's distance / source row is fetched once, stores are row-contiguous.
  const int lane = threadIdx.x & 31;
  const int64_t warp = tid >> 5, nwarps = stride >> 5;
  if (s.edge_attr) {
    if (o.edge_attr) copy_span(o.edge_attr + eo * G, s.edge_attr + ep * G, e * G, tid, stride);
    if (o.edge_attr_slots) {
      for (int64_t sl = warp; sl < e; sl += nwarps) {
        const float* __restrict__ from = s.edge_attr + (int64_t)__ldg(s.dst_eid + ep + sl) * G;
        float* __restrict__ to = o.edge_attr_slots + (eo + sl) * G;
        for (int j = lane; j < G; j += 32) to[j] = __ldg(from + j);
      }
    }
  } else if (o.edge_attr || o.edge_attr_slots) {  // neither: the batch keeps the 4 B/edge form (fused CGConv kernels)
    for (int64_t k = warp; k < e; k += nwarps) {
      const float d_ref = o.edge_attr ? __ldg(s.d_hat + ep + k) : 0.f;
      const float d_slot = o.edge_attr_slots ? __ldg(s.d_hat + __ldg(s.dst_eid + ep + k)) : 0.f;
      for (int j = lane; j < G; j += 32) {
        const float off = __ldg(o.smear_offset + j);
        if (o.edge_attr) {
          const float diff = d_ref - off;
          o.edge_attr[(eo + k) * G + j] = expf(o.smear_coeff * (diff * diff));
        }
        if (o.edge_attr_slots) {
          const float diff = d_slot - off;
          o.edge_attr_slots[(eo + k) * G + j] = expf(o.smear_coeff * (diff * diff));
        }
      }
    }
  }
}

}  // namespace mdl

extern "C" int mdl_assemble_batch(const mdl_graph_store* store, const mdl_batch_out* out, void* stream) {
  using namespace mdl;
  MDL_REQUIRE(store && out, "assemble_batch: null descriptor");
  MDL_REQUIRE(out->B >= 0 && out->N >= 0 && out->E >= 0, "assemble_batch: bad shape");
  MDL_REQUIRE(out->N < (1ll << 31) && out->E < (1ll << 31) && store->num_edges < (1ll << 31) &&
              store->num_nodes < (1ll << 31), "assemble_batch: sizes must fit int32");
  if (out->B == 0) return MDL_OK;
  MDL_REQUIRE(out->B < (1ll << 31) - 1, "assemble_batch: too many graphs");
  MDL_REQUIRE(store->F > 0 && store->G > 0 && store->U >= 0 && store->Y >= 0 && store->U <= 256 && store->Y <= 256,
              "assemble_batch: bad widths");
  MDL_REQUIRE(store->node_ptr && store->edge_ptr && store->x && store->src && store->dst && store->edge_weight &&
              store->u && store->y, "assemble_batch: store has null arrays");
  MDL_REQUIRE(store->edge_attr || store->d_hat, "assemble_batch: store needs edge_attr or d_hat");
  MDL_REQUIRE(out->graph_ids && out->node_off && out->edge_off && out->x && out->edge_index && out->edge_weight &&
              out->batch && out->u && out->y, "assemble_batch: batch has null arrays");
  MDL_REQUIRE(!out->d_hat || store->d_hat, "assemble_batch: d_hat requested but not stored");
  MDL_REQUIRE(store->edge_attr || (!out->edge_attr && !out->edge_attr_slots) || out->smear_offset,
              "assemble_batch: smear_offset needed to expand edge_attr from d_hat");
  if (out->dst_ptr) {
    MDL_REQUIRE(out->dst_src && out->dst_dst && out->dst_eid && out->src_ptr && out->src_slot &&
                out->inv_deg_dst && out->inv_deg_src && out->graph_ptr, "assemble_batch: incomplete layout outputs");
    MDL_REQUIRE(store->dst_ptr && store->dst_src && store->dst_dst && store->dst_eid && store->src_ptr &&
                store->src_slot && store->inv_deg_dst && store->inv_deg_src, "assemble_batch: store has no layout");
  } else {
    MDL_REQUIRE(!out->edge_attr_slots, "assemble_batch: edge_attr_slots needs the layout outputs");
  }
  MDL_REQUIRE(!out->edge_attr_slots || store->dst_eid, "assemble_batch: store has no layout");
  // enough CTAs per graph to fill the machine about four times over
  int parts = (int)std::min<int64_t>(16, std::max<int64_t>(1, ceil_div<int64_t>(4 * kNumSMs, out->B)));
  AsmArgs a{*store, *out};
  k_assemble<<<dim3((unsigned)out->B + 1, (unsigned)parts), 256, 0, as_stream(stream)>>>(a);
  MDL_LAUNCHED();
  return MDL_OK;
}
