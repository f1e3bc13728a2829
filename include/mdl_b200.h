/* mdl_b200.h -- C ABI of libmdl_b200.so, the sm_100a message-passing engine.
 *
 * The reference (Fung-Lab/MatDeepLearn) has no FFI: its hot path is reached by
 * Python attribute access into torch_geometric / torch_scatter.  Each entry
 * point below names the reference call site (file:line under /root/reference)
 * whose device work it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless marked "host"; the caller owns
 *     all memory (outputs and workspace included); the library allocates nothing
 *   - all floating point is fp32, all indices the library consumes are int32
 *     except the reference-layout inputs of mdl_csr_from_coo (int64)
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it
 *     and is CUDA-graph capturable (no syncs, no allocations inside)
 *   - return 0 on success; non-zero = error, text via mdl_last_error()
 *     (thread-local).  Nothing throws or exits across this boundary.
 *   - entry points are re-entrant (forward runs on the Python main thread,
 *     backward on autograd's worker thread)
 *
 * Engine edge order.  mdl_csr_from_coo sorts the E directed edges
 * (row=source j -> col=destination i, PyG flow source_to_target) stably by
 * destination.  "slot" s in [0,E) is a position in that order; edge-level
 * operands handed to the conv kernels (edge_attr, filters ...) are in slot
 * order (use mdl_gather_rows with dst_eid to permute reference-order tensors).
 */
#ifndef MDL_B200_H
#define MDL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MDL_API __attribute__((visibility("default")))
#else
#define MDL_API
#endif

#define MDL_OK 0
#define MDL_ERR_ARG 1      /* bad shape / null pointer / unsupported size */
#define MDL_ERR_CUDA 2     /* a CUDA runtime call or launch failed */
#define MDL_ERR_WORKSPACE 3

#define MDL_REDUCE_SUM 0
#define MDL_REDUCE_MEAN 1
#define MDL_REDUCE_MAX 2

MDL_API int mdl_version(void);
/* copies the calling thread's last error text into buf (NUL terminated) */
MDL_API int mdl_last_error(char* buf, size_t n);
/* number of kernels this library has launched since load (all threads);
 * bench.py reports it as gpu_launches */
MDL_API int64_t mdl_launch_count(void);

/* ---- graph layout: replaces PyG MessagePassing.__collect__'s per-layer
 * index_select/scatter bookkeeping (triggered at reference
 * matdeeplearn/models/cgcnn.py:142, schnet.py:140, mpnn.py:154) with a
 * once-per-batch sort.  Inputs are the reference's tensors as collated at
 * training/training.py:300-307: edge_index int64 [2,E], batch int64 [N]. ---- */
MDL_API size_t mdl_csr_workspace_bytes(int64_t num_nodes, int64_t num_edges);
MDL_API int mdl_csr_from_coo(const int64_t* edge_index, const int64_t* batch,
                     int64_t num_nodes, int64_t num_edges, int64_t num_graphs,
                     int32_t* dst_ptr,   /* [N+1] slots of destination n: [dst_ptr[n],dst_ptr[n+1]) */
                     int32_t* dst_src,   /* [E] source node of slot s */
                     int32_t* dst_dst,   /* [E] destination node of slot s */
                     int32_t* dst_eid,   /* [E] reference edge id of slot s */
                     int32_t* src_ptr,   /* [N+1] by-source segments over positions p */
                     int32_t* src_slot,  /* [E] slot s of by-source position p */
                     float* inv_deg_dst, /* [N] 1/max(1,in-degree)  */
                     float* inv_deg_src, /* [N] 1/max(1,out-degree) */
                     int32_t* graph_ptr, /* [B+1] node range of graph b */
                     void* workspace, size_t workspace_bytes, void* stream);

/* out[r,:] = src[idx[r],:]   (rows of `width` floats) */
MDL_API int mdl_gather_rows(const float* src, const int32_t* idx, float* out,
                    int64_t rows, int64_t width, void* stream);
/* out[idx[r],:] = src[r,:]   (inverse permutation of mdl_gather_rows) */
MDL_API int mdl_scatter_rows(const float* src, const int32_t* idx, float* out,
                     int64_t rows, int64_t width, void* stream);

/* ---- GaussianSmearing, reference matdeeplearn/process/process.py:580-590:
 * out[e,k] = exp(coeff * (d[e] - offset[k])^2); offset = the module's
 * registered linspace buffer (process.py:583), coeff per process.py:585 ---- */
MDL_API int mdl_gaussian_smear(const float* d, const float* offset, float* out, int64_t num_edges,
                       int32_t G, float coeff, void* stream);

/* ---- segmented reduce over contiguous segments: replaces torch_scatter
 * scatter(..., reduce=) / torch_geometric global_{mean,add,max}_pool
 * (reference cgcnn.py:154,169; megnet.py:86,130-132,346-348).
 * out[s,:] = reduce_{r in [ptr[s],ptr[s+1])} src[perm ? perm[r] : r, :]
 * empty segments give 0.  argmax (may be NULL unless MAX+backward) receives
 * the winning row.  bwd: grad_src[row,:] (+)= weight * grad_out[s,:]. ---- */
MDL_API int mdl_segment_reduce_fwd(const float* src, const int32_t* ptr, const int32_t* perm,
                           float* out, int32_t* argmax, int64_t num_segments,
                           int64_t width, int32_t reduce, void* stream);
MDL_API int mdl_segment_reduce_bwd(const float* grad_out, const int32_t* ptr, const int32_t* perm,
                           const int32_t* argmax, float* grad_src, int64_t num_segments,
                           int64_t num_rows, int64_t width, int32_t reduce, void* stream);

/* The reference layer's parameters -- lin_f / lin_s of torch_geometric CGConv, weight [C, 2C+G] over
 * cat[x_i, x_j, e] and bias [C] (NULL = no bias) -- repacked in one launch into the operands of the
 * hoisted form used below:  Wn [4C, C] (rows P_f | P_s | Q_f | Q_s: PQ = x Wn^T + bias),
 * bias [4C] (zero on the Q rows) and WeT [G, 2C]. */
MDL_API int mdl_cgconv_pack_weights(const float* w_f, const float* b_f, const float* w_s, const float* b_s,
                                    int32_t C, int32_t G, float* Wn, float* bias, float* WeT, void* stream);

/* ---- CGConv, PyG CGConv(channels=C, dim=G, aggr, batch_norm=False) as
 * constructed at reference matdeeplearn/models/cgcnn.py:80-82 and called at
 * cgcnn.py:136-145.  The caller splits lin_f/lin_s column-wise into
 * [W_i | W_j | W_e] and supplies the node-level projections
 *   PQ[n, 0:2C]  = x[n] W_i^T + b   (f rows then s rows)   "P", used at destination
 *   PQ[n, 2C:4C] = x[n] W_j^T       (f rows then s rows)   "Q", used at source
 * and We = [W_e of lin_f ; W_e of lin_s]  [2C, G].
 * fwd: out[i] = aggr_{s: dst=i} sigmoid(a_f) * softplus(a_s) + x[i],
 *      a = P[i] + Q[src(s)] + We . ea[s]
 * bwd: given grad_out [N,C] (for MEAN: already multiplied row-wise by inv_deg_dst -- a node-level
 *      elementwise op the caller fuses with its other node work) returns dPQ [N,4C] (dP = sum over in-edges of da,
 *      dQ = sum over out-edges of da), dWe [2C,G]; the caller finishes with
 *      dense node-level GEMMs (dx = grad_out + dPQ . Wn, dWn = dPQ^T x). ---- */
/* Which of the two entry points below run on the tcgen05 kernels for layer width C and edge width G under the default
 * dispatch: bit 0 = forward (cgconv_fwd_ws.cu), bit 1 = backward (cgconv_bwd.cu).  They serve C >= 64 (multiple of 4)
 * in 64-channel chunks and G <= 64; other shapes take the SIMT kernels (any C % 4 == 0, any G). */
MDL_API int mdl_cgconv_tc_supported(int32_t C, int32_t G);
MDL_API size_t mdl_cgconv_workspace_bytes(int64_t num_nodes, int64_t num_edges, int32_t C, int32_t G);
MDL_API int mdl_cgconv_fwd(const float* x, const float* PQ, const float* ea, const float* We,
                   const int32_t* dst_ptr, const int32_t* dst_src, const int32_t* dst_dst,
                   const float* inv_deg_dst, float* out,
                   int64_t num_nodes, int64_t num_edges, int32_t C, int32_t G,
                   int32_t reduce, void* stream);
MDL_API int mdl_cgconv_bwd(const float* grad_out, const float* PQ, const float* ea, const float* We,
                   const int32_t* dst_ptr, const int32_t* dst_src, const int32_t* dst_dst,
                   const int32_t* src_ptr, const int32_t* src_slot,
                   const float* inv_deg_dst, float* dPQ, float* dWe,
                   int64_t num_nodes, int64_t num_edges, int32_t C, int32_t G,
                   int32_t reduce, void* workspace, size_t workspace_bytes, void* stream);

/* ---- CGConv, smearing-fused form.  The reference expands every edge's normalised distance into a Gaussian basis
 * once, at dataset build (matdeeplearn/process/process.py:500-502 calling GaussianSmearing, process.py:580-590:
 * e[k] = exp(coeff * (d_hat - offset[k])^2), offset = linspace(start, stop, G)), and its CGConv layers then read the
 * [E, G] tensor from memory in every layer of every step (cgcnn.py:136-145).  These entry points take d_hat [E]
 * (slot order, i.e. permuted by dst_eid like `ea` above) plus the module's `offset` buffer [G] (device; uniform
 * spacing, G >= 2) and `coeff`, and expand the basis inside the fused edge kernels: 4 B instead of 4 G B per edge.
 * Same operator, outputs and gradients as mdl_cgconv_fwd / _bwd on ea = GaussianSmearing(d_hat) (basis within
 * 2.5e-6 of torch.exp's at the reference's parameters, 6e-6 for coarse or narrow bases).  mdl_cgconv_smear_supported: 1 if (C, G) is served by the tensor-core kernels that
 * implement this form (C = 64, G <= 64, no MDL_CGCONV_* kernel switch in the environment); otherwise the caller
 * materialises ea (mdl_gaussian_smear) and uses the entry points above. ---- */
MDL_API int mdl_cgconv_smear_supported(int32_t C, int32_t G);
MDL_API int mdl_cgconv_smear_fwd(const float* x, const float* PQ, const float* d_hat, const float* offset, float coeff,
                   const float* We, const int32_t* dst_ptr, const int32_t* dst_src, const int32_t* dst_dst,
                   const float* inv_deg_dst, float* out,
                   int64_t num_nodes, int64_t num_edges, int32_t C, int32_t G, int32_t reduce, void* stream);
MDL_API int mdl_cgconv_smear_bwd(const float* grad_out, const float* PQ, const float* d_hat, const float* offset,
                   float coeff, const float* We, const int32_t* dst_ptr, const int32_t* dst_src,
                   const int32_t* dst_dst, const float* inv_deg_dst, float* dPQ, float* dWe,
                   int64_t num_nodes, int64_t num_edges, int32_t C, int32_t G,
                   int32_t reduce, void* workspace, size_t workspace_bytes, void* stream);

/* ---- CFConv aggregate, PyG CFConv inside InteractionBlock as built at reference
 * matdeeplearn/models/schnet.py:81 and called schnet.py:134-143:
 *   out[i,:] = sum_{p in [ptr[i],ptr[i+1])} h[nbr[p],:] * w[eid[p],:]
 * forward: ptr=dst_ptr, nbr=dst_src, eid=dst_eid, w = filter [E,F] in REFERENCE edge order;
 * backward (dh): ptr=src_ptr, nbr = destination of each by-source position, eid = its edge id,
 * h := grad_out.  eid/nbr may be NULL (identity). ---- */
MDL_API int mdl_spmm_edge(const float* h, const float* w, const int32_t* ptr, const int32_t* nbr,
                          const int32_t* eid, float* out, int64_t num_segments, int64_t width,
                          void* stream);
/* out[eid[p],:] = a[ia[p],:] * b[ib[p],:]  -- CFConv filter gradient dW = grad_out[dst] * h[src] */
/* The same aggregate with ONE coefficient per edge (GCNConv's normalised adjacency D^-1/2 A D^-1/2 with
 * A_ij = edge_weight, reference matdeeplearn/models/gcn.py:80-82,141): out[i,:] = sum_p coef[eid[p]] * h[nbr[p],:];
 * and its coefficient gradient: out[eid[p]] = < a[ia[p],:], b[ib[p],:] >. */
MDL_API int mdl_spmm_edge_scalar(const float* h, const float* coef, const int32_t* ptr, const int32_t* nbr,
                                 const int32_t* eid, float* out, int64_t num_segments, int64_t width, void* stream);
MDL_API int mdl_edge_dot(const float* a, const float* b, const int32_t* ia, const int32_t* ib, const int32_t* eid,
                         float* out, int64_t num_edges, int64_t width, void* stream);
MDL_API int mdl_edge_mul(const float* a, const float* b, const int32_t* ia, const int32_t* ib,
                         const int32_t* eid, float* out, int64_t num_edges, int64_t width, void* stream);

/* ---- Megnet_EdgeModel's first Linear on cat[x[row], x[col], e, u[batch[row]]] (reference
 * matdeeplearn/models/megnet.py:41-47) with the weight split column-wise:
 *   out[e,:] = act(base[e,:] + A[ia[e],:] + B[ib[e],:] + U[ig[ia[e]],:] + bias)
 * base = e W_e^T (edge-level GEMM by the caller), A = x W_src^T, B = x W_dst^T, U = u W_u^T;
 * ia/ib = edge_index rows (int64, reference layout), ig = batch (int64). ---- */
MDL_API int mdl_edge_gather_add(const float* base, const float* A, const float* B, const float* U,
                                const int64_t* ia, const int64_t* ib, const int64_t* ig,
                                const float* bias, float* out, int64_t num_edges, int64_t width,
                                int32_t relu, void* stream);

/* ---- NNConv message, PyG NNConv as built at reference matdeeplearn/models/mpnn.py:83-88,
 * re-associated so that the [E, C*C] per-edge weight tensor is never formed:
 *   XT[j,k,:] = x[j] . T[k]   (T = second edge-network Linear reshaped [K, C, O]; caller's GEMM)
 *   m[e,:]    = sum_k hid[e,k] * XT[src(e),k,:] + XB[src(e),:]   (XB = x . reshape(bias2))
 * by-source CSR: src_ptr [N+1], src_eid [E] = reference edge id of each position. ---- */
MDL_API int mdl_nnconv_msg_fwd(const float* hid, const float* XT, const float* XB, const int32_t* src_ptr,
                               const int32_t* src_eid, float* m, int64_t num_nodes, int32_t K, int32_t O,
                               void* stream);
MDL_API int mdl_nnconv_msg_bwd(const float* hid, const float* XT, const float* dm, const int32_t* src_ptr,
                               const int32_t* src_eid, float* dhid, float* dXT, float* dXB,
                               int64_t num_nodes, int32_t K, int32_t O, void* stream);

/* ---- graph builder on the GPU --------------------------------------------------------------
 * The per-structure part of the reference's process_data (matdeeplearn/process/process.py:284-305
 * distances + threshold_sort + dense_to_sparse + add_self_loops, :540-560 threshold_sort, :365-388
 * and :594-605 node features).  Structures are concatenated: pos [num_nodes,3] f64, numbers
 * [num_nodes] i32, node_ptr [num_graphs+1] i64, cell [num_graphs,3] f64 orthorhombic box lengths
 * (NULL or 0 entries = not periodic along that axis).
 *
 * mdl_build_neighbors: one CTA per structure (max_nodes = the largest structure, <= ~2300 atoms).
 *   Row i keeps its neighbors+1 closest columns within `radius` by (distance, column) -- numpy's
 *   stable ordinal rank --, drops exact-zero distances, and writes them in ascending column order:
 *   nbr_col/nbr_w [num_nodes, neighbors+1] (column local to the structure, weight = float(distance)),
 *   cnt [num_nodes].  Distances are fp64 without FMA contraction (bit-equal to numpy).
 * mdl_build_emit: scatters the table into src/dst/edge_weight at first_edge[node] (the caller's
 *   prefix sum of cnt, laid out so that each structure's loops follow its edges), writes the loop
 *   (node,node,0) at loop_pos[node], and sets the two one-hot entries of x[node] (x zero-filled by
 *   the caller): column numbers[node]-1 and column z_width + cnt[node] + 1. */
MDL_API int mdl_build_neighbors(const double* pos, const double* cell, const int64_t* node_ptr,
                                int64_t num_graphs, int32_t max_nodes, double radius, int32_t neighbors,
                                int32_t* nbr_col, float* nbr_w, int32_t* cnt, void* stream);
/* The same with general (triclinic) cells: lattice [num_graphs, 28] f64 (NULL = none) holds, per structure, the
 * lattice vectors (rows, 9), their inverse (9), the image-shift range per axis (3), the per-axis periodicity flags (3)
 * and a "general cell" flag (entry 24; 0 = use the box lengths in `cell`) -- what the reference obtains from ASE's
 * get_all_distances(mic=True) on any cell (process.py:284-287).  The minimum-image search evaluates the host
 * builder's expressions (process._general_minimum_image) in the same order: bit-identical distances. */
MDL_API int mdl_build_neighbors_lattice(const double* pos, const double* cell, const double* lattice,
                                        const int64_t* node_ptr, int64_t num_graphs, int32_t max_nodes, double radius,
                                        int32_t neighbors, int32_t* nbr_col, float* nbr_w, int32_t* cnt, void* stream);
MDL_API int mdl_build_emit(const int32_t* nbr_col, const float* nbr_w, const int32_t* cnt,
                           const int64_t* first_edge, const int64_t* loop_pos,
                           const int64_t* node_graph_start, const int32_t* numbers, int64_t num_nodes,
                           int32_t neighbors, int32_t z_width, int32_t F, int32_t* src, int32_t* dst,
                           float* w, float* x, void* stream);

/* ---- batch assembly from a device-resident dataset -----------------------------------------
 * Replaces the reference's per-step CPU collate + H2D copy: PyG DataLoader -> Batch.from_data_list
 * (matdeeplearn/training/training.py:300-307) and data.to(rank) (training.py:39).
 *
 * mdl_graph_store: every graph of the processed dataset concatenated on the device, i.e. one
 * block-diagonal graph in the reference's tensor layout (process.py:504-523), plus the
 * destination-major layout mdl_csr_from_coo produces over ALL its edges (optional: NULL dst_ptr
 * means "no layout stored").  edge_attr may be NULL when d_hat (the normalised distance the
 * Gaussian basis is expanded from, process.py:486-502) is stored instead.
 * Node ids in src/dst/dst_src/dst_dst are store-global; dst_eid/src_slot are store-global edge /
 * slot numbers.  All pointers are device pointers; the two structs themselves live on the host. */
typedef struct mdl_graph_store {
  int64_t num_graphs, num_nodes, num_edges;
  int32_t F, G, U, Y;              /* widths of x, edge_attr, u, y rows */
  const int64_t* node_ptr;         /* [num_graphs+1] */
  const int64_t* edge_ptr;         /* [num_graphs+1] */
  const float* x;                  /* [num_nodes, F] */
  const int32_t* src;              /* [num_edges] edge_index[0] */
  const int32_t* dst;              /* [num_edges] edge_index[1] */
  const float* d_hat;              /* [num_edges] or NULL */
  const float* edge_weight;        /* [num_edges] */
  const float* edge_attr;          /* [num_edges, G] or NULL */
  const float* u;                  /* [num_graphs, U] */
  const float* y;                  /* [num_graphs, Y] */
  const int32_t* dst_ptr;          /* [num_nodes+1] */
  const int32_t* dst_src;          /* [num_edges] */
  const int32_t* dst_dst;          /* [num_edges] */
  const int32_t* dst_eid;          /* [num_edges] */
  const int32_t* src_ptr;          /* [num_nodes+1] */
  const int32_t* src_slot;         /* [num_edges] */
  const float* inv_deg_dst;        /* [num_nodes] */
  const float* inv_deg_src;        /* [num_nodes] */
} mdl_graph_store;

/* One batch = graphs graph_ids[0..B) in that order.  node_off/edge_off are the exclusive prefix
 * sums of their node/edge counts ([B+1] each, device); their last entries are the batch's node and
 * edge totals and are read ON THE DEVICE, so one captured launch serves batches of any size up to
 * the capacities N and E (the row counts of the output arrays; N = node_off[B], E = edge_off[B]
 * gives the exact batch).  Rows beyond the totals are filled with inert padding: zero features,
 * batch id B, empty segments in dst_ptr/src_ptr, edges/slots no segment refers to, graph_ptr[B] =
 * node_off[B] -- segment-driven operators and readouts therefore ignore them.  Outputs are exactly
 * Batch.from_data_list's tensors (x, edge_index int64 [2,E], edge_weight, edge_attr, batch, u, y;
 * d_hat and edge_attr optional) and, when dst_ptr != NULL, the arrays mdl_csr_from_coo would
 * produce for that batch (bit-identical) plus edge_attr in slot order (optional).  When the store
 * holds no edge_attr it is expanded as exp(smear_coeff * (d_hat - smear_offset[j])^2). */
typedef struct mdl_batch_out {
  int64_t B, N, E;
  const int64_t* graph_ids;
  const int64_t* node_off;
  const int64_t* edge_off;
  float* x;
  int64_t* edge_index;
  float* d_hat;                    /* optional */
  float* edge_weight;
  float* edge_attr;                /* optional */
  float* edge_attr_slots;          /* optional, needs the layout outputs */
  int64_t* batch;
  float* u;
  float* y;
  int32_t* dst_ptr;                /* NULL: skip every layout output below */
  int32_t* dst_src;
  int32_t* dst_dst;
  int32_t* dst_eid;
  int32_t* src_ptr;
  int32_t* src_slot;
  float* inv_deg_dst;
  float* inv_deg_src;
  int32_t* graph_ptr;
  const float* smear_offset;       /* [G] device, or NULL when the store holds edge_attr */
  float smear_coeff;
} mdl_batch_out;

MDL_API int mdl_assemble_batch(const mdl_graph_store* store, const mdl_batch_out* out, void* stream);

/* ---- masked training-mode BatchNorm1d -------------------------------------------------------
 * torch.nn.BatchNorm1d as the reference applies it to node features after every conv
 * (matdeeplearn/models/cgcnn.py:88-92,141-147; torch semantics: biased variance for the
 * normalisation, unbiased for running_var, running = (1-momentum)*running + momentum*batch),
 * with the statistics taken over the first *n_valid rows only (n_valid: device int32, NULL = all
 * N rows) and rows >= *n_valid written as zero (forward) / zero gradient (backward), so that a
 * capacity-padded batch normalises exactly like the unpadded one.  weight/bias/running_* may be
 * NULL.  workspace: mdl_batchnorm_workspace_bytes(N, C) bytes, zero-filled once by the caller
 * (the kernels leave its first word at zero). */
MDL_API size_t mdl_batchnorm_workspace_bytes(int64_t N, int32_t C);
MDL_API int mdl_batchnorm_fwd(const float* x, const int32_t* n_valid, int64_t N, int32_t C,
                              const float* weight, const float* bias, float* running_mean,
                              float* running_var, float momentum, float eps, float* out,
                              float* save_mean, float* save_invstd, void* workspace,
                              size_t workspace_bytes, void* stream);
MDL_API int mdl_batchnorm_bwd(const float* gout, const float* x, const int32_t* n_valid, int64_t N,
                              int32_t C, const float* weight, const float* save_mean,
                              const float* save_invstd, float* gx, float* gweight, float* gbias,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ---- dense-layer weight gradient over a tall batch, written in place -------------------------
 * dW[O,I] = G[N,O]^T X[N,I] and db[O] = column sums of G: what autograd's Linear backward computes
 * for torch.nn.Linear (reference cgcnn.py:64-77,97-111) and for lin_f / lin_s inside PyG's CGConv.
 * Rows are split over the grid, per-CTA partials are summed in CTA order (deterministic).
 * The result is delivered through a block map: output row o belongs to block o / block_rows; its
 * weights go to w[block] + (o % block_rows) * ldw + i, its bias sum to b[block][o % block_rows]
 * (NULL entries are skipped).  A plain Linear uses one block; CGConv's hoisted projections use
 * four (P_f, P_s, Q_f, Q_s -> column blocks of lin_f.weight.grad / lin_s.weight.grad), so every
 * gradient lands where the optimizer reads it (e.g. inside one flat gradient buffer) with no
 * concatenation.  mdl_copy_mapped scatters an already reduced [O,I] (or transposed [I,O]) result
 * the same way.  workspace: mdl_linear_wgrad_workspace_bytes(N, I, O) bytes. */
typedef struct mdl_wgrad_out {
  int32_t block_rows;
  int32_t num_blocks;              /* <= 8 */
  int64_t ldw;                     /* row stride of every w[] block, in floats */
  float* w[8];
  float* b[8];
} mdl_wgrad_out;
MDL_API size_t mdl_linear_wgrad_workspace_bytes(int64_t N, int32_t I, int32_t O);
MDL_API int mdl_linear_wgrad(const float* X, const float* G, int64_t N, int32_t I, int32_t O,
                             const mdl_wgrad_out* out, void* workspace, size_t workspace_bytes,
                             void* stream);
MDL_API int mdl_copy_mapped(const float* src, int32_t I, int32_t O, int32_t transposed,
                            const mdl_wgrad_out* out, void* stream);

/* Same, with the gradient rows scaled by rowscale[r] first (a layer whose output is multiplied per row afterwards:
 * SchNet's filter * cosine cutoff, reference schnet.py:81 -> PyG CFConv.forward).  N >= 2048 rows, I <= 256. */
MDL_API int mdl_linear_wgrad_rs(const float* X, const float* G, const float* rowscale, int64_t N, int32_t I, int32_t O,
                                const mdl_wgrad_out* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- dense layer over a long batch on tcgen05 (3xTF32): Y[r,:] = act(X[r,:] . B^T + bias), X [R,K] (K <= 128),
 * B[n][k] = W[n*ldn + k*ldk] (N <= 128).  Forward y = x W^T with W [O,I]: ldn = I, ldk = 1; input gradient
 * dx = g W: ldn = 1, ldk = I.  act: 0 none, 1 relu, 2 shifted softplus.  Replaces the cuBLAS SGEMMs over E rows of
 * the reference's edge-level Linear layers (megnet.py:28-56,222-247; mpnn.py:83-85) and of their backward. ---- */
MDL_API int mdl_linear_tc_supported(int64_t R, int32_t K, int32_t N);
MDL_API int mdl_linear_tc(const float* X, const float* W, int64_t ldn, int64_t ldk, const float* bias, float* Y,
                          int64_t R, int32_t K, int32_t N, int32_t act, void* stream);

/* ---- fused two-layer edge MLP (SchNet filter network + cosine cutoff), reference models/schnet.py:81,134-143 ->
 * PyG InteractionBlock.mlp = Sequential(Linear(G,F), ShiftedSoftplus, Linear(F,F)) and CFConv's `W = mlp(e) * C`:
 *   Y[e,:] = (act1(X[e,:] W1^T + b1) W2^T + b2) * rowscale[e]      X [E,G] (G <= 64), H = O = 128
 * act1: 0 = shifted softplus, 1 = relu; act2: 0 = none, 1 = relu (applied before the row scale).
 * T1 (optional, [E,H]) receives the hidden activations for the backward.  One pass on tcgen05 (3xTF32). ---- */
MDL_API int mdl_edge_mlp2_supported(int32_t G, int32_t H, int32_t O);
MDL_API int mdl_edge_mlp2_fwd(const float* X, const float* W1, const float* b1, const float* W2, const float* b2,
                              const float* rowscale, float* Y, float* T1, int64_t E, int32_t G, int32_t H, int32_t O,
                              int32_t act1, int32_t act2, void* stream);
/* dPre1[e,:] = ((dY[e,:] * rowscale[e]) W2) * act1'(pre1[e,:]), act1' recovered from T1.  The weight / bias
 * gradients are mdl_linear_wgrad_rs(T1, dY, rowscale) and mdl_linear_wgrad(X, dPre1). */
MDL_API int mdl_edge_mlp2_bwd(const float* dY, const float* rowscale, const float* W2, const float* T1, float* dPre1,
                              int64_t E, int32_t H, int32_t O, int32_t act1, void* stream);


/* ---- AdamW over one flat fp32 buffer: torch.optim.AdamW semantics (the reference's optimizer,
 * config.yml "optimizer: AdamW", matdeeplearn/training/training.py:429-432, step at :49).
 * hyper = device {lr, beta1, beta2, eps, weight_decay}; step = device float step count, advanced by
 * the call; grad_scale multiplies the gradient first (1/world after a sum all-reduce). ---- */
MDL_API int mdl_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                           const float* hyper, float* step, float grad_scale, int64_t n, void* stream);

/* ---- development aid: 32 x uint64 device counters that receive per-phase cycle sums
 * (thread 0 of every CTA) from the tensor-core CGConv kernels; NULL disables. ---- */
MDL_API int mdl_debug_set_phase_buffer(void* dev_ptr);

#ifdef __cplusplus
}
#endif
#endif /* MDL_B200_H */
