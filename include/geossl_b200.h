/*
 * geossl_b200.h -- C ABI of libgeossl_b200.so: the sm_100a kernels behind the GeoSSL-DDM
 * pretraining hot path (SchNet / PaiNN encoder fwd+bwd, DDM score loss).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked (host);
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work (no host sync);
 *   - node/edge features are row-major fp32; reference-facing indices arrive as int64 (PyG
 *     convention, schnet.py:86) and the internal CSR is int32;
 *   - edge / pair counts that are data dependent live in device memory (`n_edges_dev` is usually
 *     `rowptr + n_atoms`), buffers are sized at a host-known capacity, so a whole training step can be
 *     captured in a CUDA graph;
 *   - return value: 0 on success, a positive cudaError_t, or a negative GEOSSL_E* argument error.
 *     geossl_last_error() returns a thread-local message for the last failing call.
 *
 * Each entry cites the reference interface (file:line under chao1224/GeoSSL) it replaces.
 */
#ifndef GEOSSL_B200_H
#define GEOSSL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GEOSSL_ABI_VERSION 4   /* 2: filt_row / pair-index / batched-pack entry points; 3: round-2 entry points; 4: cfconv_pairs, pair product, chain_ex, wgrad batch, elementwise families */
#define GEOSSL_EINVAL (-1)   /* bad argument (null pointer, unsupported width, ...) */
#define GEOSSL_ECAP   (-2)   /* capacity too small */

int geossl_abi_version(void);
const char* geossl_last_error(void);
/* Number of kernel launches issued through this library by the calling process (bench.py's
 * "gpu_launches" evidence).  reset: non-zero zeroes the counter after reading. */
int64_t geossl_launch_count(int reset);

/* ------------------------------------------------------------------------------------------------
 * Graph construction.  Replaces torch_cluster.radius_graph as called at Geom3D/models/schnet.py:91
 * (every forward) and Geom3D/datasets/datasets_3D_Radius.py:120 (PaiNN, dataset time).
 * ---------------------------------------------------------------------------------------------- */

/* graph_ptr[g] = first atom of graph g, graph_ptr[n_graphs] = n_atoms, from the sorted `batch`
 * vector (dataloaders_AtomTuple.py:55,72).  Empty graphs are allowed. */
int geossl_graph_ptr(const int64_t* batch, int64_t n_atoms, int64_t n_graphs, int32_t* graph_ptr, void* stream);

/* Destination-sorted CSR of the fixed-radius graph with torch_cluster's semantics: candidates of the
 * same graph scanned in ascending atom index, dist = ((dx*dx)+dy*dy)+dz*dz in fp32 without FMA, keep
 * if dist < r*r, stop after max_num_neighbors+1 hits (self included), self removed.
 *   rowptr (n_atoms+1) int32, rowptr[n_atoms] = E;  src (capacity) int32 ascending within a row;
 *   edge_tgt (capacity) int32 = row of each edge;  edge_dist (capacity) fp32 = ||pos[src]-pos[tgt]||
 *   (schnet.py:92-93), may be NULL.  capacity >= E is required; (max_num_neighbors+1)*n_atoms always
 *   suffices.  scratch: (2*n_atoms + 2) int32.
 * flags: GEOSSL_RADIUS_FMA evaluates the distance as fma(dz,dz, fma(dy,dy, dx*dx)) -- what `dist += (x-y)*(x-y)` becomes
 *   under the compiler's default FMA contraction; torch_cluster is an un-vendored, unpinned dependency of the reference,
 *   so which rounding its binary used cannot be checked here: both are provided, the default is the uncontracted form
 *   (SURVEY.md Appendix B.1).  The two differ only when dist is within an ulp of r*r.
 * Cell list (optional, for graphs of >= cell_min_atoms atoms): cell_box / sorted_keys / sorted_atoms from
 *   geossl_radius_cell_keys + ONE stable ascending sort of the keys (sorted_atoms = the sort's permutation).  NULL
 *   cell_box => every graph takes the index-order scan.  The output is bit-identical either way. */
#define GEOSSL_RADIUS_FMA 1
int geossl_radius_csr(const float* pos, const int64_t* batch, const int32_t* graph_ptr, int64_t n_atoms,
                      float r, int max_num_neighbors, int64_t capacity, int32_t* scratch,
                      int32_t* rowptr, int32_t* src, int32_t* edge_tgt, float* edge_dist, int flags,
                      const float* cell_box, const int64_t* sorted_keys, const int64_t* sorted_atoms, int cell_min_atoms,
                      void* stream);
/* Per-graph bounding box -> uniform grid with cell edge 1.001 r (coarser if a box would need more than 1023 cells per
 * dimension): box (n_graphs,8) fp32 scratch, keys (n_atoms) int64 = graph << 30 | cx << 20 | cy << 10 | cz. */
int geossl_radius_cell_keys(const float* pos, const int64_t* batch, const int32_t* graph_ptr, int64_t n_atoms, int64_t n_graphs,
                            float r, float* box, int64_t* keys, void* stream);

/* (2,E) int64 `edge_index` = [source; target] exactly as radius_graph returns it (host already knows E). */
int geossl_csr_to_edge_index(const int32_t* src, const int32_t* edge_tgt, int64_t n_edges,
                             int64_t* edge_index, void* stream);

/* Source-sorted view of a destination-sorted CSR whose rows have ascending sources and whose graphs
 * occupy contiguous rows (what geossl_radius_csr / radius_graph emit):
 *   t_rowptr (n_atoms+1), t_eid (capacity) = edge ids grouped by source, ascending target within a
 *   source, t_tgt (capacity) = target of each of those edges.  Deterministic, atomic free. */
int geossl_csr_transpose(const int32_t* rowptr, const int32_t* src, const int64_t* batch,
                         const int32_t* graph_ptr, int64_t n_atoms, int32_t* scratch,
                         int32_t* t_rowptr, int32_t* t_eid, int32_t* t_tgt, void* stream);

/* Undirected-pair index of a destination-sorted CSR with ascending sources per row.  The interaction filter depends on an
 * edge only through its length (schnet.py:186-187) and |pos_j - pos_i| == |pos_i - pos_j| bit for bit, so both directions
 * of a pair can share one filter row: the filter network runs over U <= E pairs instead of E edges (U = E/2 when no row
 * is truncated).  A pair lives in the row of its smaller atom (owner_small = 1: canonical direction source > target; rows
 * walked in ascending order then stream their own block first and find shared rows in L2) or of its larger atom
 * (owner_small = 0); an edge whose reverse was cut by max_num_neighbors or is absent is its own pair ("orphan").  Row t's
 * pairs are its canonical edges in row order.
 *   pair_rowptr  (n_atoms+1)  pairs owned by each target row; pair_rowptr[n_atoms] = U stays on the device
 *   pair_of_edge (capacity)   pair id of every directed edge (the filt_row map of geossl_cfconv_*)
 *   pair_e1, pair_e2 (capacity) canonical edge id of pair u and its reverse edge id (or -1)
 *   pair_atoms   (2*capacity, 8-byte aligned) (s, t) of pair u's canonical edge s -> t; t is stored as ~t (negative)
 *                             when the reverse direction t -> s does not exist
 *   pair_dist    (capacity)   edge length of pair u
 * scratch: n_atoms int32.  Deterministic, atomic free, no host sync. */
int geossl_pair_index(const int32_t* rowptr, const int32_t* src, const float* edge_dist, int64_t n_atoms, int32_t* scratch,
                      int owner_small, int32_t* pair_rowptr, int32_t* pair_of_edge, int32_t* pair_e1, int32_t* pair_e2,
                      int32_t* pair_atoms, float* pair_dist, void* stream);

/* Device-side batch assembly: all ordered atom pairs of every molecule in itertools order (combination: i<j,
 * permutation: i!=j), offset by the cumulative atom count -- AtomTupleExtractor + BatchAtomTuple.from_data_list
 * (dataloaders_AtomTuple.py:15-37,45-73; ratio == 1).  pair_ptr (n_graphs+1) int64 = exclusive scan of the per-graph
 * pair counts; super_edge_index (2,n_pairs) int64; batch (n_atoms) int64 may be NULL. */
int geossl_super_edges(const int32_t* graph_ptr, const int64_t* pair_ptr, int64_t n_graphs, int64_t n_atoms, int permutation,
                       int64_t n_pairs, int64_t* super_edge_index, int64_t* batch, void* stream);

/* rowptr from a SORTED int64 key vector (row 1 of a radius_edge_index), keys in [0,n_rows). */
int geossl_rowptr_from_sorted(const int64_t* keys, int64_t n_keys, int64_t n_rows, int32_t* rowptr, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SchNet continuous-filter convolution.  Replaces GaussianSmearing.forward (schnet.py:205-207),
 * the filter network + cosine cutoff of CFConv.forward (schnet.py:141-145,186-187) and
 * MessagePassing.propagate/message (schnet.py:190,194-195; torch_scatter atomics) and their autograd.
 * F = num_filters in {32,64,128}; G = num_gaussians <= 64.
 * ---------------------------------------------------------------------------------------------- */

/* W_e = (Lin2(ssp(Lin1(rbf(d_e))))) * 0.5*(cos(d_e*pi/cutoff)+1)   -> filt (E,F)
 * w1 (F,G), b1 (F), w2 (F,F), b2 (F) are nn.Linear layouts (out,in). */
int geossl_filter_fwd(const float* edge_dist, const int32_t* n_edges_dev, int64_t capacity,
                      const float* offset, float coeff, float cutoff, int G, int F,
                      const float* w1, const float* b1, const float* w2, const float* b2,
                      float* filt, void* stream);

/* Same contract as geossl_filter_fwd, on the tcgen05 tensor cores (F = 128 only).  fp32 operands are split
 * into two 16-bit parts and each product is three MMAs with fp32 accumulation in TMEM: fp16 parts
 * (bf16_parts = 0; ~2^-22 per product, fp32 grade) or bf16 parts (bf16_parts = 1; ~2^-17, fp32 range). */
int geossl_filter_fwd_tc(const float* edge_dist, const int32_t* n_edges_dev, int64_t capacity,
                         const float* offset, float coeff, float cutoff, int G, int F,
                         const float* w1, const float* b1, const float* w2, const float* b2,
                         float* filt, int bf16_parts, void* stream);

/* Debug: per-phase clock64() trace of CTA 0 of the tensor-core filter kernels into a device buffer of 512
 * int64 slots (slot = tile*16 + event); NULL disables.  Used by profiles/trace_tc.py. */
int geossl_debug_set_trace(long long* device_buffer);       /* forward kernel */
int geossl_debug_set_trace_bwd(long long* device_buffer);   /* backward kernel */
int geossl_debug_set_trace_head(long long* device_buffer);  /* tensor-core DDM head kernels */
int geossl_debug_set_trace_linear(long long* device_buffer); /* atom-wise dense layer kernel (8 events) */

/* Self test of the tcgen05 plumbing (descriptors, swizzle, TMEM): one 128 x N x K split-precision GEMM.
 * mode 0: d[m][n] = sum_k a[m][k] b[n][k]  (a (128,K), b (128,K), K in {64,128}, N = 128; K-major operands)
 * mode 1: d[m][n] = sum_k a[k][m] b[k][n]  (a (128,128), b (128,N), N in {64,128}; MN-major operands)
 * mode 2: throughput probe (cycle counts in d[0..1]);  mode 3: as mode 0 with A resident in tensor memory, b (N,K) */
int geossl_tc_selftest(int mode, int fp16, const float* a, const float* b, int K, int N, float* d, void* stream);

/* m_i = sum_{e in row i} x[src_e] * W_row(e)   (atomic-free segmented reduction, one warp per row).
 * filt_row: NULL => row(e) = e (one filter row per directed edge); else row(e) = filt_row[e] (the pair_of_edge map of
 * geossl_pair_index: both directions of an undirected pair read ONE shared filter row). */
int geossl_cfconv_fwd(const float* x, const float* filt, const int32_t* filt_row, const int32_t* rowptr, const int32_t* src,
                      int64_t n_atoms, int F, float* out, void* stream);

/* dx_j = sum_{e: src_e = j} W_e * g[tgt_e]   over the source-sorted view. */
int geossl_cfconv_bwd_x(const float* filt, const int32_t* filt_row, const float* grad_out, const int32_t* t_rowptr,
                        const int32_t* t_eid, const int32_t* t_tgt, int64_t n_atoms, int F, float* grad_x, void* stream);

/* Pair-centric cfconv aggregate for batches of SMALL graphs (F = 128, one filter row per atom pair, geossl_pair_index):
 * one CTA per graph keeps the graph's operand rows and accumulators in shared memory and streams the graph's contiguous
 * block of filter rows ONCE, adding W_u * v[s] to row t and W_u * v[t] to row s for pair u = (s, t)  -- the same sums as
 * geossl_cfconv_fwd (transposed = 0, v = x) / geossl_cfconv_bwd_x (transposed = 1, v = grad_out) with filt_row =
 * pair_of_edge, in a different (fixed) fp32 summation order; every filter row crosses L2 -> SM once instead of twice
 * (schnet.py:190,194-195).  graph_ptr (n_graphs+1) int32 atom offsets, pair_rowptr (n_atoms+1), pair_atoms (capacity,2).
 * max_graph_atoms: host-known bound on the atoms of any graph that has edges (<= geossl_cfconv_pairs_max_atoms(groups));
 * a larger graph is still summed correctly, by one warp straight in global memory.  tuning: 0 = default, else
 * groups * 100 + rows (warps per graph; filter rows per register buffer, 8 or 16, two buffers in flight per warp). */
int geossl_cfconv_pairs_max_atoms(int groups);
int geossl_cfconv_pairs(const float* v, const float* filt, const int32_t* pair_atoms, const int32_t* pair_rowptr,
                        const int32_t* graph_ptr, int64_t n_graphs, int max_graph_atoms, int transposed, int tuning,
                        float* out, void* stream);

/* dW_e = x[src_e] * g[tgt_e]  materialised (E,F)  (second-order path and the unfused comparison). */
int geossl_cfconv_bwd_w(const float* x, const float* grad_out, const int32_t* rowptr, const int32_t* src,
                        int64_t n_atoms, int F, float* grad_filt, void* stream);

/* dW_u = x[s] * g[t] (+ x[t] * g[s] when the reverse edge exists) for every atom pair u = (s, t) of geossl_pair_index:
 * the edge product of the shared-row aggregate, materialised (n_pairs,F) (composed double-backward path with one filter
 * row per pair; schnet.py:190,194-195 differentiated w.r.t. W). */
int geossl_cfconv_pair_product(const float* x, const float* grad_out, const int32_t* pair_atoms, int64_t n_pairs, int F,
                               float* grad_filt, void* stream);

/* Elementwise families closed under differentiation (composed double-backward path, finetune_md17.py:32-54):
 *   geossl_ssp_family: out = g * s^(order)(a), s = shifted softplus (schnet.py:215-216), order 0..3 (order 0: the activation, g unused);
 *   geossl_row_scale:  out[r][f] = x[r][f] * c[r]  (the cutoff product of schnet.py:187);  geossl_row_dot: out[r] = sum_f a[r][f] b[r][f].
 * n = element count; operands 16-byte aligned, rows contiguous (F in {32, 64, 128}). */
int geossl_ssp_family(const float* a, const float* g, int64_t n, int order, float* out, void* stream);
int geossl_row_scale(const float* x, const float* c, int64_t n_rows, int F, float* out, void* stream);
int geossl_row_dot(const float* a, const float* b, int64_t n_rows, int F, float* out, void* stream);

/* Backward of the filter network fused with dW_e = x[src_e]*g[tgt_e] (never materialised):
 * recomputes rbf / Lin1 / ssp from d_e and accumulates gw1 (F,G), gb1 (F), gw2 (F,F), gb2 (F).
 * If grad_filt != NULL it is used as dW_e instead of x/g (then x, grad_out, src, edge_tgt may be NULL).
 * workspace: geossl_filter_bwd_workspace floats.  The gradient outputs are OVERWRITTEN (not accumulated).
 * (Force training, finetune_md17.py:46, differentiates through the geometry and uses the composable
 * primitives geossl_filter_fwd / geossl_cfconv_* from the host-side double-backward path instead.) */
int64_t geossl_filter_bwd_workspace(int G, int F);
int geossl_filter_bwd(const float* edge_dist, const int32_t* n_edges_dev, int64_t capacity,
                      const float* offset, float coeff, float cutoff, int G, int F,
                      const float* w1, const float* b1, const float* w2, const float* b2,
                      const float* x, const float* grad_out, const int32_t* src, const int32_t* edge_tgt,
                      const float* grad_filt, float* workspace,
                      float* gw1, float* gb1, float* gw2, float* gb2, void* stream);

/* Same contract as geossl_filter_bwd (x / grad_out / src / edge_tgt form), on the tcgen05 tensor cores:
 * F = 128, G <= 63, operands split into two bf16 parts, weight gradients accumulated in TMEM.
 * workspace: geossl_filter_bwd_tc_workspace() floats.
 * Pair form (pair_atoms non-NULL, from geossl_pair_index; src / edge_tgt are then unused and may be NULL):
 * edge_dist / n_edges_dev / capacity describe the undirected PAIRS, and the filter-output gradient of pair u is formed
 * as the sum over its one or two directions, x[s]*g[t] + x[t]*g[s]; the parameter gradients equal the per-edge form up
 * to fp32 summation order. */
int64_t geossl_filter_bwd_tc_workspace(void);
int geossl_filter_bwd_tc(const float* edge_dist, const int32_t* n_edges_dev, int64_t capacity,
                         const float* offset, float coeff, float cutoff, int G, int F,
                         const float* w1, const float* b1, const float* w2,
                         const float* x, const float* grad_out, const int32_t* src, const int32_t* edge_tgt,
                         const int32_t* pair_atoms,
                         float* workspace, float* gw1, float* gb1, float* gw2, float* gb2, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Atom-wise dense layers, 128 -> 128, on the tcgen05 tensor cores (split-precision operands).  Replace the
 * nn.Linear calls of schnet.py:99-101 (head), :165-166 (act + lin), :189,:191 (CFConv lin1 / lin2) and the
 * residual of :97.
 *   y[r][n] = epi( sum_k pre(x[r][k]) * Wm[n][k] ),  pre = identity | shifted softplus (pre_ssp),
 *   Wm = weight (out,in) or its transpose (data gradient) -- chosen when the weight is packed,
 *   epi: + bias[n] (may be NULL), * sigmoid(act_grad_input[r][n]) (may be NULL), + residual[r][n] (may be NULL).
 * ---------------------------------------------------------------------------------------------- */
/* `weight_image` = the weight packed by geossl_pack_weight (split + swizzled shared-memory operand image,
 * geossl_weight_image_bytes() bytes, 16-byte aligned) with the same transpose / parts flags; every CTA fetches it with
 * one bulk async copy (cp.async.bulk). */
int64_t geossl_weight_image_bytes(void);
int geossl_pack_weight(const float* weight, int transpose_weight, int bf16_parts, void* image, void* stream);
/* The same for a 128 x 128 block of a wider row-major matrix (ldw = its row stride in floats, any value >= 128). */
int geossl_pack_weight_ld(const float* weight, int ldw, int transpose_weight, int bf16_parts, void* image, void* stream);
/* Both images of one block in one launch: images[0] = forward (fp16 parts), images[1] = transposed (bf16 parts). */
int geossl_pack_weight_pair(const float* weight, int ldw, void* images, void* stream);
/* Every 128 x 128 weight BLOCK of a model in one launch: `weights` = DEVICE array of n_weights device pointers to the first
 * element of a block, `lds` = DEVICE array of the leading dimensions (in_features) of the matrices the blocks live in
 * (NULL => 128: plain 128 x 128 layers); images = n_weights x 2 x geossl_weight_image_bytes(): [b][0] forward image
 * (fp16 parts), [b][1] data-gradient image (transposed, bf16 parts). */
int geossl_pack_weights_batched(const float* const* weights, const int32_t* lds, int n_weights, void* images, void* stream);
int geossl_linear_tc(const float* x, int64_t n_rows, const void* weight_image, const float* bias, int pre_ssp,
                     const float* act_grad_input, const float* residual, float* y, int bf16_parts, void* stream);

/* One 128 x 128 BLOCK of a wider dense layer (PaiNN's Dense 128->384, 256->128, 128->256: painn.py:21-24,76-83,
 * painn_utils.py:9-35): the same kernel over column windows of wider row-major tensors.  x / act_grad_input / residual / y
 * point at the first element of their 128-column window; ld* are the row strides in floats.  act selects the
 * pre-activation (pre_act != 0) and its derivative (act_grad_input != NULL): 1 = shifted softplus, 2 = SiLU.  A K > 128
 * layer is the sum of its K-blocks: pass the partial result as `residual` (it may alias y).
 * k_cols (multiple of 16, <= 128): x has only k_cols live columns (a K-padded operand such as PaiNN's 20 rbf values + the
 * bias column): only those are read and only k_cols/16 MMA k-steps are issued. */
int geossl_linear_tc_block(const float* x, int64_t ldx, int64_t n_rows, const void* weight_image, const float* bias, int act,
                           int pre_act, const float* act_grad_input, int64_t ldz, const float* residual, int64_t ldr,
                           float* y, int64_t ldy, int bf16_parts, int k_cols, void* stream);
/* grad_weight block [o][i] (row stride ld_gw) = sum_r grad_y[r][o] * pre(x[r][i]), pre_act in {0, 1 = ssp, 2 = SiLU};
 * x_cols (multiple of 8, <= 128): live columns of x, the others count as zeros (their gradient columns come out zero). */
int geossl_linear_wgrad_tc_block(const float* grad_y, int64_t ld_dy, const float* x, int64_t ld_x, int64_t n_rows, int pre_act,
                                 float* workspace, float* grad_weight, int ld_gw, float* grad_bias, int x_cols, void* stream);

/* Up to four 128 -> 128 layers applied to the same row tile without leaving the SM (an interaction block's
 * conv.lin2 -> ssp -> lin (+ residual) -> next conv.lin1, schnet.py:191,165-166,97,189; the head lin1 -> ssp -> lin2, :99-101;
 * and the data-gradient chains of their backward with transposed images).  Stage s computes
 *   v = operand_s . W_s^T + bias_s;  v *= act'(act_grad_input_s);  v += residual_s;  store_s = v;  operand_{s+1} = act_next_s(v)
 * with every optional piece NULL / 0 when absent; operand_0 = x (n_rows,128).  act selects the activation family of
 * act' (1 = shifted softplus, 2 = SiLU); act_next is 0 (identity), 1 or 2 per stage.  The last stage must store. */
typedef struct {
    const void* weight_image;      /* packed by geossl_pack_weight(s_batched) with the orientation / parts this stage needs */
    const float* bias;             /* (128) or NULL */
    const float* act_grad_input;   /* (n_rows,128) or NULL */
    const float* residual;         /* (n_rows,128) or NULL */
    float* store;                  /* (n_rows,128) or NULL */
    int act_next;
} geossl_chain_stage;
int geossl_linear_chain_tc(const float* x, int64_t n_rows, const geossl_chain_stage* stages /*host*/, int n_stages, int bf16_parts,
                           int act, void* stream);

/* The same kernel with fan-out, fan-in and strided tensors, for the wider Dense layers of PaiNN (painn.py:21-24,76-83:
 * 128 -> 128 SiLU -> 384, 256 -> 128 SiLU -> 384, 128 -> 256 on the 3N vector rows) and their data gradients in ONE launch each:
 *   x != NULL : this stage loads its own operand tile (128 columns starting at x, row stride ldx, pre-activation x_act);
 *               required for the first stage.  keep = 1 : the previous stage's operand tile is used again (fan-out over
 *               several 128-column weight blocks).  partial = 1 / accumulate = 1 : consecutive stages sum into one
 *               accumulator (fan-in over the K-blocks of a K > 128 layer); only the last of them has an epilogue.
 *   z, residual, store: windows of wider tensors (row strides ldz, ldr, ld_store).  Up to 6 stages. */
typedef struct {
    const void* weight_image;
    const float* bias;
    const float* act_grad_input; int64_t ldz;
    const float* residual;       int64_t ldr;
    float* store;                int64_t ld_store;
    const float* x;              int64_t ldx;
    int act_next, x_act, keep, accumulate, partial;
} geossl_chain_stage_ex;
int geossl_linear_chain_ex(int64_t n_rows, const geossl_chain_stage_ex* stages /*host*/, int n_stages, int bf16_parts, int act,
                           void* stream);

/* grad_weight[o][i] = sum_r grad_y[r][o] * pre(x[r][i]);  grad_bias[o] = sum_r grad_y[r][o] (may be NULL). */
int64_t geossl_linear_wgrad_tc_workspace(int64_t n_rows);
int geossl_linear_wgrad_tc(const float* grad_y, const float* x, int64_t n_rows, int pre_ssp, float* workspace,
                           float* grad_weight, float* grad_bias, void* stream);

/* Up to 6 weight-gradient products over the SAME n_rows rows in one launch (+ one reduction launch): the three layers of an
 * interaction tail, the 128 x 128 blocks of a wider Dense layer, ...  Each problem has its own workspace
 * (geossl_linear_wgrad_tc_workspace(n_rows) floats) and the argument meaning of geossl_linear_wgrad_tc_block. */
typedef struct {
    const float* grad_y; int64_t ld_dy;
    const float* x;      int64_t ld_x;
    float* workspace;
    float* grad_weight;  int ld_gw;
    float* grad_bias;    /* may be NULL */
    int pre_act, x_cols;
} geossl_wgrad_problem;
int geossl_linear_wgrad_tc_batch(const geossl_wgrad_problem* problems /*host*/, int n_problems, int64_t n_rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * DDM head.  Replaces the distance block of do_DDM (examples/pretrain_GeoSSL.py:197-205) and
 * NCSN_version_03.forward (examples/NCSN.py:183-212) + MultiLayerPerceptron (NCSN.py:9-43) and their
 * autograd.  H = emb_dim in {32,64,128}.
 * ---------------------------------------------------------------------------------------------- */

/* d_p = sqrt(sum((pos[u_p]-pos[v_p])^2))  (P,) ; sei = super_edge_index (2,P) int64. */
int geossl_pair_distance(const float* pos, const int64_t* sei, int64_t n_pairs, float* dist, void* stream);

typedef struct {
    const float* in_w0;  /* input_distance_mlp.layers.0.weight (H,1) */
    const float* in_b0;  /* (H) */
    const float* in_w1;  /* input_distance_mlp.layers.1.weight (1,H) */
    const float* in_b1;  /* (1) */
    const float* out_w0; /* output_mlp.layers.0.weight (H,H+1) */
    const float* out_b0; /* (H) */
    const float* out_w1; /* output_mlp.layers.1.weight (H/2,H) */
    const float* out_b1; /* (H/2) */
    const float* out_w2; /* output_mlp.layers.2.weight (1,H/2) */
    const float* out_b2; /* (1) */
} geossl_ddm_params;

typedef struct {
    float* in_w0; float* in_b0; float* in_w1; float* in_b1;
    float* out_w0; float* out_b0; float* out_w1; float* out_b1; float* out_w2; float* out_b2;
} geossl_ddm_grads;

/* loss (1,) = mean over graphs of sum_p 0.5*(score_p - target_p)^2 * sigma_p^anneal_power with the
 * noise level (per graph) and the N(0,1) draw (per pair) supplied by the caller (RNG contract,
 * NCSN.py:190,194).  The mean is over max_p(batch[u_p])+1 graphs (torch_scatter dim_size).
 * workspace: geossl_ddm_workspace(H) floats.
 * n_pairs_live (all four entry points): NULL => every one of the n_pairs pairs is live.  Otherwise n_pairs is the
 * CAPACITY of a padded batch (row stride of sei, length of dist / noise) and *n_pairs_live, read on the device, is the
 * number of live pairs at the front: no host sync, so a CUDA graph captured once serves batches of any pair count up
 * to the capacity (variable-size Molecule3D batches, dataloaders_AtomTuple.py:45-78). */
int64_t geossl_ddm_workspace(int H);
/* workspace of the tensor-core editions (geossl_ddm_head_{fwd,bwd}_tc): the partial sums plus 8 per-pair scalars
 * (endpoints, graph id, sigma, perturbed distance, target, sigma^anneal, distance embedding) that a one-thread-per-pair
 * prologue kernel writes before the MMA pipeline starts. */
int64_t geossl_ddm_workspace_tc(int64_t n_pairs);
int geossl_ddm_head_fwd(const float* h, const int64_t* sei, const int64_t* batch, int64_t n_pairs, const int32_t* n_pairs_live,
                        const float* dist, const float* noise, const int64_t* noise_level,
                        const float* sigmas, int n_levels, float anneal_power, int H,
                        const geossl_ddm_params* params /*host*/, float* workspace, float* loss /*(2,): loss, #graphs*/,
                        void* stream);

/* Backward: grad_h (n_atoms,H) is ZEROED then accumulated; parameter gradients are overwritten.
 * grad_loss is the (1,) upstream gradient on the device. */
int geossl_ddm_head_bwd(const float* h, const int64_t* sei, const int64_t* batch, int64_t n_pairs, const int32_t* n_pairs_live,
                        int64_t n_atoms, const float* dist, const float* noise, const int64_t* noise_level,
                        const float* sigmas, int n_levels, float anneal_power, int H,
                        const geossl_ddm_params* params /*host*/, const float* loss_aux /*(2,) from fwd*/,
                        const float* grad_loss, float* workspace,
                        float* grad_h, const geossl_ddm_grads* grads /*host*/, void* stream);

/* Same contracts on the tcgen05 tensor cores (H = 128; fp16-split forward, bf16-split backward operands). */
int geossl_ddm_head_fwd_tc(const float* h, const int64_t* sei, const int64_t* batch, int64_t n_pairs, const int32_t* n_pairs_live,
                           const float* dist, const float* noise, const int64_t* noise_level,
                           const float* sigmas, int n_levels, float anneal_power, int H,
                           const geossl_ddm_params* params /*host*/, float* workspace, float* loss /*(2,)*/, void* stream);
int geossl_ddm_head_bwd_tc(const float* h, const int64_t* sei, const int64_t* batch, int64_t n_pairs, const int32_t* n_pairs_live,
                           int64_t n_atoms, const float* dist, const float* noise, const int64_t* noise_level,
                           const float* sigmas, int n_levels, float anneal_power, int H,
                           const geossl_ddm_params* params /*host*/, const float* loss_aux, const float* grad_loss,
                           float* workspace, float* grad_h, const geossl_ddm_grads* grads /*host*/, void* stream);
/* Forward AND backward of the head in ONE pass over the pairs (training: the backward kernel recomputes the forward per
 * tile anyway, so a separate forward launch is redundant work).  loss (2,) as geossl_ddm_head_fwd_tc.  grad_h and the
 * parameter gradients are the gradients of the UNSCALED sum over pairs (the number of graphs is only known once every
 * CTA has finished): the caller multiplies them by grad_loss / loss[1] (ops.DDMHead.backward: one multi-tensor scale).
 * The first layer of the score MLP is applied per ATOM (it is linear in h[u] + h[v]): A = h W0[:, :128]^T by the dense-layer
 * kernel, gathered per pair; its data / weight gradients are one per-atom GEMM each after the pair kernel has scattered dz1.
 * workspace: geossl_ddm_workspace_fused(n_pairs, n_atoms) floats. */
int64_t geossl_ddm_workspace_fused(int64_t n_pairs, int64_t n_atoms);
int geossl_ddm_head_fwd_bwd_tc(const float* h, const int64_t* sei, const int64_t* batch, int64_t n_pairs, const int32_t* n_pairs_live,
                               int64_t n_atoms, const float* dist, const float* noise, const int64_t* noise_level,
                               const float* sigmas, int n_levels, float anneal_power, int H,
                               const geossl_ddm_params* params /*host*/, float* workspace, float* loss /*(2,)*/, float* grad_h,
                               const geossl_ddm_grads* grads /*host*/, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PaiNN message block.  Replaces the per-edge part of PaiNN.forward (Geom3D/models/painn.py:232-245:
 * r_ij, d_ij, dir_ij, GaussianRBF painn_utils.py:99-103, CosineCutoff painn_utils.py:152-155,
 * filter_net * fcut) and PaiNNInteraction.forward's gather / Wij*xj / split / index_add (painn.py:53-64).
 * F = n_atom_basis in {32,64,128}; n_rbf <= 32.  The (E,3F) filter is rebuilt per edge, never stored.
 * Edge order: radius_edge_index (2,E) = [idx_i; idx_j] sorted by idx_j (row 1), idx_i ascending inside.
 * ---------------------------------------------------------------------------------------------- */

/* dist (E), dir (E,3) = (pos[idx_i]-pos[idx_j])/d, fcut (E) = 0.5(cos(d*pi/rc)+1)*[d<rc].
 * Capacity-padded lists (CUDA-graph replay): padding columns carry idx_j = n_atoms (one past the last atom, which keeps
 * the list idx_j-sorted and makes rowptr[n_atoms] of geossl_rowptr_from_sorted the LIVE edge count on the device); they
 * get fcut = 0 / dir = 0 here and are never visited by the message kernels (which walk the CSR rows). */
int geossl_painn_edge_geometry(const float* pos, const int64_t* radius_edge_index, int64_t n_edges, int64_t n_atoms, float cutoff,
                               float* dist, float* dir, float* fcut, void* stream);

/* q_out = q + sum_{e: idx_i=n} W_e[0:F]*ctx[j][0:F];
 * mu_out = mu + sum_e ( W_e[F:2F]*ctx[j][F:2F] * dir_e + W_e[2F:3F]*ctx[j][2F:3F] * mu[j] ),
 * W_e = (rbf(d_e) filter_w^T + filter_b) * fcut_e with filter_w (3F,n_rbf) the interaction's slice of
 * filter_net.weight.  ctx (N,3F) = interatomic_context_net(q); mu (N,3,F).
 * (i_rowptr, i_eid, i_nbr): edges grouped by idx_i -- edge id and idx_j of each. */
int geossl_painn_message_fwd(const float* q, const float* mu, const float* ctx, const float* filter_w, const float* filter_b,
                             const float* offsets, const float* widths, int n_rbf, int F,
                             const float* dist, const float* dir, const float* fcut,
                             const int32_t* i_rowptr, const int32_t* i_eid, const int32_t* i_nbr, int64_t n_atoms,
                             float* q_out, float* mu_out, const float* filter_pre, void* stream);
/* filter_pre (fwd and bwd): NULL => the 3F filter values of an edge are rebuilt from the rbf values against the
 * filter_net slice (any F).  Otherwise (E,3F): the PRE-cutoff filter rows phi_e W_f^T + b_f materialised by the tensor-core
 * filter GEMM (geossl_painn_rbf_pad + geossl_linear_tc_block over the K-padded rbf matrix); the kernels stream them and
 * multiply by fcut.  In the backward, edge_scratch then IS the gradient w.r.t. filter_pre and the filter_net gradients come
 * from that GEMM's weight-gradient kernel (geossl_linear_wgrad_tc_block); filter_w/b, workspace, grad_filter_* may be NULL. */
/* phi_pad (E,ld) = [rbf_0(d_e) .. rbf_{R-1}(d_e), 1, 0, ...], ld in {32,64,128} > n_rbf: the K-padded A operand of the filter
 * GEMM (column R carries the bias of filter_net).  GaussianRBF, painn_utils.py:99-103. */
int geossl_painn_rbf_pad(const float* dist, const float* fcut, int64_t n_edges, const float* offsets, const float* widths, int n_rbf,
                         int ld, float* phi_pad, void* stream);

/* Backward.  (j_rowptr, j_ctr): rowptr over the idx_j-sorted edge list and idx_i of each edge (int32).
 * Outputs: grad_ctx (N,3F), grad_mu_in (N,3,F) (includes the identity path), grad_filter_w (3F,n_rbf),
 * grad_filter_b (3F); dL/dq_in is grad_q_out itself.  edge_scratch: (E,3F) floats; workspace:
 * geossl_painn_workspace floats.  Positions receive no gradient (DDM pretraining does not need it). */
int64_t geossl_painn_workspace(int n_rbf, int F);
int geossl_painn_message_bwd(const float* grad_q_out, const float* grad_mu_out, const float* mu, const float* ctx,
                             const float* filter_w, const float* filter_b, const float* offsets, const float* widths,
                             int n_rbf, int F, const float* dist, const float* dir, const float* fcut,
                             const int32_t* j_rowptr, const int32_t* j_ctr, int64_t n_atoms, int64_t n_edges,
                             float* grad_ctx, float* grad_mu_in, float* edge_scratch, float* workspace,
                             float* grad_filter_w, float* grad_filter_b, const float* filter_pre, void* stream);

/* PaiNN update block, non-GEMM part (PaiNNMixing.forward, painn.py:100-113).  mu_mix (N,3,2F) = mu_channel_mix(mu).
 *   pre : ctx (N,2F) = [q, sqrt(sum_xyz mu_V^2 + eps)], dot (N,F) = sum_xyz mu_V*mu_W
 *   post: y (N,3F) = intraatomic_context_net(ctx) = [a|b|c];  q_out = q + a + c*dot;  mu_out = mu + b*mu_W
 * and their backward kernels (grad_mu_mix contributions of pre and post are summed by the caller). */
int geossl_painn_mix_pre(const float* q, const float* mu_mix, int64_t n_atoms, int F, float epsilon, float* ctx, float* dot, void* stream);
int geossl_painn_mix_pre_bwd(const float* mu_mix, const float* ctx, const float* grad_ctx, const float* grad_dot, int64_t n_atoms, int F,
                             float* grad_q, float* grad_mu_mix, void* stream);
int geossl_painn_mix_post(const float* q, const float* mu, const float* y, const float* mu_mix, const float* dot, int64_t n_atoms, int F,
                          float* q_out, float* mu_out, void* stream);
int geossl_painn_mix_post_bwd(const float* grad_q_out, const float* grad_mu_out, const float* y, const float* mu_mix, const float* dot,
                              int64_t n_atoms, int F, float* grad_y, float* grad_mu_mix, float* grad_dot, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOSSL_B200_H */
