/*
 * kagnn_b200.h -- C ABI of libkagnn_b200.so: the B200 (sm_100a) forward path of KAGNN's
 * KAN layers fused with the GCN / GIN / GINE neighbour aggregation that feeds them.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; memory is caller-owned;
 *     nothing is allocated or retained by the library (explicit workspace queries instead);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and re-entrant;
 *   - return value: 0 on success, a negative KAGNN_E* code otherwise (kagnn_strerror());
 *   - all floating-point data is fp32 row-major with an explicit leading dimension (elements);
 *     graph indices are int32 once in CSR form (int64 COO only at kagnn_csr_build);
 *   - there is no CPU fallback: an unsupported configuration is an error.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the KAGNN
 * repository root; nc = node_classification_clean, gc = graph_classification,
 * gr = graph_regression).
 */
#ifndef KAGNN_B200_H
#define KAGNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KAGNN_VERSION 100 /* 0.1.0 */

enum {
    KAGNN_OK = 0,
    KAGNN_EINVAL = -1,      /* bad shape / null pointer / inconsistent arguments          */
    KAGNN_EUNSUPPORTED = -2,/* valid request outside what the kernels implement           */
    KAGNN_EALIGN = -3,      /* pointer or leading dimension not aligned as required       */
    KAGNN_EWORKSPACE = -4,  /* workspace too small (see the *_workspace query)            */
    KAGNN_ECUDA = -5,       /* a CUDA runtime call / kernel launch failed                 */
    KAGNN_EINDEX = -6       /* edge_index / batch entry out of range                      */
};

/* Basis family of one KAN layer. */
enum {
    KAGNN_BASIS_BSPLINE = 0, /* ekan.KANLinear      (nc/ekan.py:7-162)        */
    KAGNN_BASIS_RBF = 1      /* fastkan.FastKANLayer (nc/fastkan.py:49-85)    */
};

/* Aggregation that produces the row tile fed to the KAN chain. */
enum {
    KAGNN_AGG_NONE = 0,     /* rows of x as they are (bare KANLinear / KAN / FastKAN)                           */
    KAGNN_AGG_GIN = 1,      /* self_scale*x_i + sum_{j->i} x_j             (PyG GINConv;  nc/models.py:48-56)   */
    KAGNN_AGG_GINE = 2,     /* self_scale*x_i + sum relu(x_j + e_ji)       (PyG GINEConv; gr/models.py:98)      */
    KAGNN_AGG_WEIGHTED = 3, /* self_weight[i]*x_i + sum w_e x_j            (PyG GCNConv;  nc/models.py:31-37)   */
    KAGNN_AGG_SEGMENT_SUM = 4,  /* sum of rows rowptr[i]..rowptr[i+1]      (global_add_pool;  gc/models.py:117) */
    KAGNN_AGG_SEGMENT_MEAN = 5  /* ... divided by max(count,1)             (global_mean_pool; gc/models.py:192) */
};

enum { KAGNN_ACT_NONE = 0, KAGNN_ACT_SILU = 1 };

/* y = act(x * scale[c] + shift[c]) per column c; NULL scale = 1, NULL shift = 0.
 * Carries GCNConv's bias, eval-mode BatchNorm1d (nc/models.py:197) and gc.KAGCN's silu (gc/models.py:190). */
typedef struct KagnnAffine {
    const float* scale;
    const float* shift;
    int32_t act;
    int32_t _pad;
} KagnnAffine;

/* One KAN layer, weights pre-packed by kagnn_pack_kan_weights(). */
typedef struct KagnnKanLayer {
    int32_t basis;          /* KAGNN_BASIS_*                                                                */
    int32_t in_features;
    int32_t out_features;
    int32_t grid_size;      /* B-spline: G.  RBF: num_grids                                                  */
    int32_t spline_order;   /* B-spline: k (1..4).  RBF: ignored                                             */
    float t0;               /* B-spline: first knot t_0 = lo - k*h (nc/ekan.py:28-37).  RBF: grid_min        */
    float h;                /* B-spline: knot spacing.  RBF: spacing of the centres                          */
    float inv_denominator;  /* RBF: 1/denominator (nc/fastkan.py:44).  B-spline: ignored                     */
    const float* packed_w;  /* [in][slots+1][out_pad4]: slots = G+k (B-spline) or G (RBF); last = base weight */
    const float* base_bias; /* RBF base_linear.bias (out) or NULL                                            */
    const float* ln_weight; /* RBF LayerNorm weight (in) or NULL = no LayerNorm                              */
    const float* ln_bias;   /* RBF LayerNorm bias (in) or NULL                                               */
    const void* packed_w_tc;/* optional: weights packed by kagnn_pack_kan_weights_tc() -> enables the tcgen05 path   */
    const float* ln_stats;  /* RBF, optional, FIRST layer of a launch in mode KAGNN_AGG_NONE only: per-row (mean, rstd) of
                             * its LayerNorm from kagnn_layernorm_stats(); lets the pipelined kernel take inputs wider than
                             * one 128-column tile unit (e.g. the skip-concat read-out).  NULL = computed in the kernel.   */
} KagnnKanLayer;

/* Input side of the fused layer: where rows come from and how they are aggregated. */
typedef struct KagnnAggregate {
    int32_t mode;             /* KAGNN_AGG_*                                                                 */
    int32_t num_cols;         /* feature width F of x                                                        */
    const float* x;           /* (num_src_rows, F), leading dimension ldx                                    */
    int64_t ldx;
    const int32_t* src_index; /* optional: row j of the logical x is x[src_index[j]] (embedding lookup)      */
    const int32_t* rowptr;    /* (num_rows+1) CSR over destination rows (or segment pointers for pooling)    */
    const int32_t* col;       /* (nnz) source row per CSR entry; NULL for the SEGMENT modes                  */
    const float* edge_weight; /* WEIGHTED: (nnz) in CSR order                                                */
    const float* self_weight; /* WEIGHTED: (num_rows) weight of x_i; NULL = self_scale                       */
    float self_scale;         /* GIN / GINE: 1 + eps                                                         */
    int32_t _pad;
    const float* edge_feat;   /* GINE: edge features, row r = edge_feat[edge_index_of(r)*ld_edge ...]         */
    int64_t ld_edge;
    const int32_t* edge_row;  /* GINE: (nnz) row of edge_feat for each CSR entry (perm or table code)        */
    /* Node-sharded graphs: source rows >= num_local_src are halo rows (copies of rows owned by other ranks,
     * filled by the per-layer all-to-all) and live in a second matrix: row j reads
     * x_halo[(j - num_local_src) * ld_halo ...].  x_halo == NULL: every source row is in x.                  */
    const float* x_halo;
    int64_t ld_halo;
    int64_t num_local_src;
    /* Node-sharded graphs, in-kernel NVLink gather (alternative to x_halo): `col` holds GLOBAL source ids and source row j lives
     * on rank j / rows_per_rank at peer_x[rank] + (j % rows_per_rank) * ldx, where peer_x is a DEVICE array of num_ranks
     * peer-mapped base pointers (same column slice and leading dimension on every rank; peer_x[my rank] == x).  The gather
     * warps of the fused kernel load remote rows directly over NVLink while the tensor-core pipeline works on the previous
     * tile: no pack kernel, no all-to-all, no halo matrix.  The caller orders layers across ranks (a barrier between the
     * producer launch of a matrix and the launches that read it remotely).  peer_x == NULL: not used.
     * Implemented by the pipelined tcgen05 kernel (128-bit path); other kernels return KAGNN_EUNSUPPORTED.                 */
    const float* const* peer_x;
    int64_t rows_per_rank;
    int32_t num_ranks;
    /* Two-part rows for KAGNN_AGG_NONE (the skip concat of nc/models.py:196-201 without the copy of x): logical row r is
     * [x_head[r, 0:num_head_cols] | x[r, 0:num_cols - num_head_cols]].  num_head_cols == 0: not used.  Implemented by the
     * pipelined tcgen05 kernel when num_head_cols is a multiple of 128; otherwise KAGNN_EUNSUPPORTED (the host side then
     * materialises the concatenation).                                                                                    */
    int32_t num_head_cols;
    const float* x_head;
    int64_t ld_head;
    /* Node-sharded graphs, halo rows that arrive WHILE the layer runs (kagnn_gather_rows_peer_ordered on a second stream): the
     * halo matrix is filled in the order the destination tiles first use its rows; halo_need[t] (one entry per 128-row tile) =
     * number of leading halo rows that tiles 0..t reference, halo_flags[c] >= 16 * halo_epoch <=> halo rows [256 c, 256 c + 256)
     * have landed (cumulative arrival counters of the pull kernel's 16 warps; epoch = 1, 2, ... per use, never reset).  The gather warps of the pipelined kernel wait for the flags of their tile's prefix before they read x_halo, so the
     * NVLink transfer overlaps the tensor-core pipeline tile by tile instead of preceding it.  reserve_sms: SMs the launch leaves
     * free (for the concurrently running pull kernel).  halo_flags == NULL: not used (x_halo must be complete at launch).     */
    const int32_t* halo_need;
    const int32_t* halo_flags;
    int32_t halo_epoch;
    int32_t reserve_sms;
    /* Node-sharded graphs, output rows PUSHED to the other ranks while the layer runs: besides y, every finished 128-row tile
     * is copied row by row into push_y[i] + row * ld_push for the num_push peer-mapped destinations (DEVICE array of pointers to
     * row 0 of THIS launch's rows inside each peer's replica of the layer output), by one otherwise idle warp per CTA that
     * relays the rows through shared memory with bulk copies (posted NVLink writes: no round trip, so the transfer rides
     * behind the tensor-core pipeline of the following tiles).  push_mask: optional, one byte per row, bit i set <=> peer i
     * needs the row (NULL: every row goes to every peer).  The caller orders the ranks (a barrier between this launch and the
     * peers' launches that read the replica).  Implemented by the pipelined tcgen05 kernel for outputs whose width is a multiple
     * of 4 columns and at most 128, 16-byte aligned y / ldy / destinations; other kernels return KAGNN_EUNSUPPORTED.
     * num_push == 0: not used.                                                                                              */
    float* const* push_y;
    const uint8_t* push_mask;
    int64_t ld_push;
    int32_t num_push;
    int32_t _pad2;
} KagnnAggregate;

/* ---- library ------------------------------------------------------------------------------------ */
int kagnn_version(void);
const char* kagnn_strerror(int code);
/* Number of SMs / max dynamic shared memory of the current device (for host-side planning). */
int kagnn_device_info(int32_t* num_sms, int32_t* max_smem_bytes, int32_t* cc_major, int32_t* cc_minor);

/* ---- graph ingestion (replaces what PyG's MessagePassing.propagate / gcn_norm redo per call) ------ */
/* COO int64 edge_index (2,E) row-major [src row; dst row] (PyG flow source_to_target) ->
 * destination-sorted CSR (stable: entries of one row keep their COO order).
 * perm[p] = original edge id of CSR entry p.  Workspace: kagnn_csr_build_workspace(E, N). */
size_t kagnn_csr_build_workspace(int64_t num_edges, int64_t num_nodes);
int kagnn_csr_build(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, int64_t num_src_nodes,
                    int32_t* rowptr, int32_t* col, int32_t* perm,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Sorted batch vector (N) int64 -> segment pointers (B+1) int32 (PyG global_*_pool's `batch`). */
int kagnn_segment_ptr(const int64_t* batch, int64_t num_nodes, int64_t num_graphs, int32_t* ptr, void* stream);

/* PyG gcn_norm(add_self_loops=True) on the CSR: existing self loops get weight 0, every node gets a
 * self weight dinv[i]^2 * loop_w, other entries dinv[src]*w*dinv[dst]; deg = in-degree at the target.
 * edge_weight_in: optional user weights in CSR order (NULL = ones).  dinv: (N) scratch/output. */
int kagnn_gcn_norm(const int32_t* rowptr, const int32_t* col, int64_t num_nodes, const float* edge_weight_in,
                   float* edge_weight_out, float* self_weight_out, float* dinv, void* stream);
/* The two halves of kagnn_gcn_norm, for node-sharded graphs where dinv of halo source rows comes from their owners:
 * kagnn_gcn_degree fills dinv / self_weight of the num_nodes destination rows; kagnn_gcn_edge_weight reads
 * dinv_src[col[e]] (col may address halo rows, dinv_src has num_src entries) and dinv_dst[row]. */
int kagnn_gcn_degree(const int32_t* rowptr, const int32_t* col, int64_t num_nodes, const float* edge_weight_in,
                     float* self_weight_out, float* dinv, void* stream);
int kagnn_gcn_edge_weight(const int32_t* rowptr, const int32_t* col, int64_t num_nodes, const float* edge_weight_in,
                          const float* dinv_src, const float* dinv_dst, float* edge_weight_out, void* stream);

/* ---- weights -------------------------------------------------------------------------------------- */
/* scaled_spline_weight (nc/ekan.py:146-152) folded while packing:
 * packed[(i*(S+1)+c)*out_pad4 + o] = spline_w[o,i,c]*scaler[o,i] (c<S), base_w[o,i] (c==S).
 * spline_w is (out,in,S) contiguous -- also the layout of fastkan's spline_linear.weight (out,in*G). */
size_t kagnn_packed_weight_elems(int32_t in_features, int32_t out_features, int32_t slots);
int kagnn_pack_kan_weights(const float* base_w, const float* spline_w, const float* spline_scaler_or_null,
                           int32_t in_features, int32_t out_features, int32_t slots,
                           float* packed, void* stream);

/* Tensor-core layout of the same weights: bf16 hi/lo pairs in the UMMA K-major canonical layout, chunked in the
 * order the tcgen05 kernel streams them (kagnn_b200/csrc/fused_tc.cu).  Supported: slots <= 8, out <= 256. */
size_t kagnn_packed_weight_tc_bytes(int32_t in_features, int32_t out_features);
int kagnn_pack_kan_weights_tc(const float* base_w, const float* spline_w, const float* spline_scaler_or_null,
                              int32_t in_features, int32_t out_features, int32_t slots, void* packed, void* stream);

/* ---- the hot path -------------------------------------------------------------------------------------
 * One launch: tile = aggregate(x) -> pre affine -> [optional store to agg_out] -> KAN layer 0 .. n_layers-1
 * (intermediate activations never leave the SM) -> post affine -> y.
 *   GIN  layer (nc/models.py:48-56 + :196-197):  mode GIN,  layers = the conv's KAN, post = eval BatchNorm
 *   GCN  layer (nc/models.py:31-37):             mode WEIGHTED over h = KAN(x), pre = bias (+BN / silu),
 *                                                layers = the NEXT conv's KAN (fusion boundary shifted half a layer)
 *   bare KANLinear.forward (nc/ekan.py:154-162): mode NONE, n_layers = 1
 *   readout (gc/models.py:117-118):              mode SEGMENT_SUM, layers = the readout KAN
 * n_layers may be 0 (pure aggregation into agg_out).  num_rows = destination rows (nodes or graphs). */
int kagnn_fused_layer_fwd(const KagnnAggregate* agg, int64_t num_rows,
                          const KagnnAffine* pre_or_null,
                          float* agg_out_or_null, int64_t ld_agg_out,
                          int32_t n_layers, const KagnnKanLayer* layers_host,
                          const KagnnAffine* post_or_null,
                          float* y, int64_t ldy, void* stream);

/* Which kernel kagnn_fused_layer_fwd may use: AUTO = tcgen05 path when every layer carries packed_w_tc and the
 * shapes fit (spline_order <= 3, G+k <= 8 or RBF G <= 8, out <= 256), else the general fp32 kernel.  Both are GPU paths. */
enum { KAGNN_PATH_AUTO = 0, KAGNN_PATH_FP32 = 1, KAGNN_PATH_TC = 2 };
int kagnn_set_path(int mode);
/* Which kernels the KAN-layer gradients (kagnn_kan_bwd_input / kagnn_kan_bwd_weights) may use: 0 (default) = the tcgen05
 * kernels when the shape fits (at least 128 rows, out <= 256; weights: G + k <= 8), else the fp32 CUDA-core kernels;
 * 1 = fp32 kernels only (tests compare the two); 2 = tcgen05 kernels, but d input splits the fp32 weights per tile instead of
 * reading the layer's tensor-core packing (packed_w_tc) as its operand, and d weights takes 128-row batches (tests);
 * 3 = tcgen05 kernels, d weights with one feature block per CTA instead of as many as tensor memory holds, d input without
 * look-ahead (tests). */
int kagnn_set_backward_path(int32_t mode);
int kagnn_get_launch_counters(int64_t* tc_launches, int64_t* fp32_launches);
/* Arithmetic of the tensor-core path.  KAGNN_PREC_FP32 (default): every product is formed from bf16 hi/lo pairs (three
 * tcgen05.mma per K step, fp32 accumulate) and matches the reference's fp32 forward within 1e-4.  KAGNN_PREC_BF16: operands are
 * rounded to bf16 once (ONE product per K step, fp32 accumulate) -- BASELINE config C5 ("fastkan ... hidden=256 grid=8 bf16");
 * about 2e-3 relative to the fp32 result per layer.  Applies to the pipelined kernel; the other kernels stay fp32. */
enum { KAGNN_PREC_FP32 = 0, KAGNN_PREC_BF16 = 1 };
int kagnn_set_precision(int mode);
int kagnn_get_precision(void);

/* The tensor-core path has two kernels: the pipelined one (A operand in tensor memory, fused_tc2.cu; B-spline chains up
 * to 128 wide) and the general one (A in shared memory, fused_tc.cu).  variant 0 = pipelined first (default),
 * 1 = general only.  kagnn_get_tc2_launches counts launches of the pipelined kernel. */
int kagnn_set_tc_variant(int variant);
int64_t kagnn_get_tc2_launches(void);

/* Self-test of the tcgen05 machinery (descriptor encodings, TMEM addressing, bulk TMA, bf16 hi/lo split):
 * D (128 x N) = A (128 x K) . B (N x K)^T, fp32 in/out, nprod = 1 (bf16 hi only) or 3 (hi/lo compensated).
 * N multiple of 16 in [16,256], K multiple of 16.  Used by tests only; no reference counterpart. */
size_t kagnn_tc_selftest_workspace(int32_t N, int32_t K);
int kagnn_tc_selftest(const float* A, const float* B, int32_t N, int32_t K, float* D, int32_t nprod,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Row gather out[r,:] = x[index[r],:] (halo send-buffer packing for the node-sharded multi-GPU path);
 * index == NULL copies rows 0..num_rows-1 (strided row copy). */
int kagnn_gather_rows(const float* x, int64_t ldx, const int32_t* index, int64_t num_rows, int32_t num_cols,
                      float* out, int64_t ld_out, void* stream);

/* Input of a B-spline layer with more than eight coefficients per (in, out) pair (G + k in 9 .. 40), laid out for the tensor-core
 * kernels: out[r, w*num_cols + c] = x[r, c] - w*shift, w = 0 .. windows-1.  Uniform B-splines are shift invariant
 * (B_{8w+j}(x) = B_j(x - 8wh)), so such a layer (ekan.py:154-162 with grid_size + spline_order > 8) IS a layer with
 * windows*in_features inputs, eight slots each and shift = 8h; the host side repacks the weights accordingly. */
int kagnn_expand_windows(const float* x, int64_t ldx, int64_t num_rows, int32_t num_cols, int32_t windows, float shift,
                         float* out, int64_t ld_out, void* stream);

/* Halo pull over NVLink (node-sharded graphs): out[r,:] = row (ids[r] % rows_per_rank) of rank (ids[r] / rows_per_rank)'s
 * matrix, read in place through peer_x, a DEVICE array of peer-mapped base pointers (see KagnnAggregate.peer_x).  Replaces
 * pack + all-to-all: only the distinct remote rows cross the link, no send lists, no NCCL.  16-byte aligned, cols % 4 == 0. */
int kagnn_gather_rows_peer(const float* const* peer_x, int64_t ldx, int64_t rows_per_rank, const int32_t* ids,
                           int64_t num_rows, int32_t num_cols, float* out, int64_t ld_out, void* stream);

/* The pull without an id list: every row id in [0, total_rows) with need[id] != 0 is copied from its owner into out[id] -- `out`
 * is a replica of the whole (total_rows x num_cols) matrix of which only the marked rows are filled.  Lets the sharded forward
 * plan a transfer with one scatter into a byte map instead of sort / unique / compaction of the remote ids. */
int kagnn_gather_rows_peer_masked(const float* const* peer_x, int64_t ldx, int64_t rows_per_rank, const uint8_t* need,
                                  int64_t total_rows, int32_t num_cols, float* out, int64_t ld_out, void* stream);

/* The same pull in first-use order with progress flags (see KagnnAggregate.halo_flags): a persistent kernel of num_ctas blocks
 * copies halo rows [256 c, 256 c + 256) chunk by chunk (block b takes chunks b, b + num_ctas, ...); each of its 16 warps adds 1
 * to chunk_flags[c] (release) when its rows are stored, so a chunk of the epoch-th use is complete at 16 * epoch.  chunk_flags
 * must be zero before the first use.  Blocks claim a whole SM each: num_ctas = the SMs given to the pull.  Launch it on
 * a second stream BEFORE the fused layer that consumes the flags. */
int kagnn_gather_rows_peer_ordered(const float* const* peer_x, int64_t ldx, int64_t rows_per_rank, const int32_t* ids,
                                   int64_t num_rows, int32_t num_cols, float* out, int64_t ld_out, int32_t* chunk_flags,
                                   int32_t epoch, int32_t num_ctas, void* stream);

/* LayerNorm row statistics (fastkan.py:66,78: biased variance, eps 1e-5) of two-part rows [x_head | x] (x_head may be NULL):
 * stats[2r] = mean, stats[2r+1] = 1/sqrt(var + eps).  See KagnnKanLayer.ln_stats. */
int kagnn_layernorm_stats(const float* x, int64_t ldx, int32_t num_cols, const float* x_head_or_null, int64_t ld_head,
                          int32_t num_head_cols, int64_t num_rows, float eps, float* stats, void* stream);

/* ---- GAT attention (PyG GATConv with the KAN as its shared projection: nc/models.py:39-46,76-83; gc/models.py:165-172,236-243) --
 * h = lin(x) is (num_rows, heads * channels).  kagnn_gat_scores: a_src / a_dst (num_rows, heads) = per-head dot products of h with
 * att_src / att_dst (heads * channels, PyG's (1, heads, channels) parameter flattened).  kagnn_gat_edge_softmax: PyG's attention
 * coefficients alpha = softmax_i(leaky_relu(a_src[j] + a_dst[i])) over the incoming CSR entries of i with existing self loops
 * removed and one self loop appended, written as edge_weight [heads][nnz] (CSR order, 0 for a removed loop) and self_weight
 * [heads][num_rows]: head by head the operands of kagnn_fused_layer_fwd in mode KAGNN_AGG_WEIGHTED on the head's column slice. */
int kagnn_gat_scores(const float* h, int64_t ldh, int64_t num_rows, int32_t heads, int32_t channels, const float* att_src,
                     const float* att_dst, float* a_src, float* a_dst, void* stream);
int kagnn_gat_edge_softmax(const int32_t* rowptr, const int32_t* col, int64_t num_rows, int64_t nnz, int32_t heads,
                           const float* a_src, const float* a_dst, float negative_slope, float* edge_weight, float* self_weight,
                           void* stream);

/* ---- small epilogues of the models (so that no step of a forward runs as framework math) ---------------------------- */
/* y[r,:] = log_softmax(x[r,:]) (gc/models.py:119,194). */
int kagnn_log_softmax_rows(const float* x, int64_t ldx, int64_t num_rows, int32_t num_cols, float* y, int64_t ldy, void* stream);
/* Training-mode BatchNorm1d forward (nc/models.py:197 under model.train(), also the no_grad validation pass of
 * nc/utils.py:172): batch mean / biased variance over all rows, y = act((x - mean) * rsqrt(var + eps) * weight + bias);
 * running_mean / running_var (may be NULL) get the momentum update with the unbiased variance.  x and y may alias. */
size_t kagnn_batchnorm_train_workspace(int32_t num_cols);
int kagnn_batchnorm_train_fwd(const float* x, int64_t ldx, int64_t num_rows, int32_t num_cols, const float* weight_or_null,
                              const float* bias_or_null, float eps, float momentum, float* running_mean_or_null,
                              float* running_var_or_null, int32_t act, float* y, int64_t ldy, void* workspace,
                              size_t workspace_bytes, void* stream);

/* Fused training-mode epilogue of a message-passing layer: y = dropout(bn(x)) of nc/models.py:196-198
 * with batch statistics -- column statistics, then ONE pass that normalises, applies the affine, updates the running estimates
 * and applies a Philox4x32-10 dropout mask keyed by (seed, element index); the mask is never stored, the backward regenerates it.
 * p_drop in [0, 1).  Workspace: kagnn_bn_dropout_train_workspace(cols) bytes for either direction. */
size_t kagnn_bn_dropout_train_workspace(int32_t num_cols);
int kagnn_bn_dropout_train_fwd(const float* x, int64_t ldx, int64_t num_rows, int32_t num_cols, const float* weight_or_null,
                               const float* bias_or_null, float eps, float momentum, float* running_mean_or_null,
                               float* running_var_or_null, float p_drop, uint64_t seed, float* y, int64_t ldy, void* workspace,
                               size_t workspace_bytes, void* stream);
int kagnn_bn_dropout_train_bwd(const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows, int32_t num_cols,
                               const float* weight_or_null, float eps, float p_drop, uint64_t seed, float* dx, int64_t ld_dx,
                               float* d_weight_or_null, float* d_bias_or_null, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Backward: what `loss.backward()` needs from the modules so that the reference's training
 * loops run (node_classification_clean/utils.py:125-132, graph_classification/graph_classification_utils.py:44-54).
 * B-spline layers only; the aggregation backward is kagnn_fused_layer_fwd with n_layers = 0 on the TRANSPOSED CSR.
 * ------------------------------------------------------------------------------------------------------------------- */

/* dx[n,i] = silu'(x) sum_o dy[n,o] Wb[o,i] + sum_s B_s'(x) sum_o dy[n,o] Ws[o,i,s] sc[o,i]: the input gradient of
 * KANLinear.forward (ekan.py:154-162) as autograd computes it through b_splines (ekan.py:79-112).  Reads layer->packed_w. */
int kagnn_kan_bwd_input(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* dy, int64_t ld_dy,
                        int64_t num_rows, float* dx, int64_t ld_dx, void* stream);

/* Gradient of the packed weights, same [in][slots+1][out_pad4] layout as kagnn_pack_kan_weights writes (slot `slots` = base
 * weight).  d_packed (kagnn_packed_weight_elems floats) is zeroed by the call; row sums use float atomics. */
int kagnn_kan_bwd_weights(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* dy, int64_t ld_dy,
                          int64_t num_rows, float* d_packed, void* stream);

/* Chain rule through scaled_spline_weight (ekan.py:146-152): d_packed -> d base_weight (out,in), d spline_weight (out,in,slots),
 * d spline_scaler (out,in).  d_base_w / d_scaler may be NULL. */
int kagnn_kan_unpack_weight_grads(const float* d_packed, const float* spline_w, const float* scaler_or_null, int32_t in_f,
                                  int32_t out_f, int32_t slots, float* d_base_w_or_null, float* d_spline_w,
                                  float* d_scaler_or_null, void* stream);

/* Backward of training-mode nn.BatchNorm1d (batch statistics recomputed from x in fp64). */
size_t kagnn_batchnorm_bwd_workspace(int32_t num_cols);
int kagnn_batchnorm_train_bwd(const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t num_rows, int32_t num_cols,
                              const float* weight_or_null, float eps, float* dx, int64_t ld_dx, float* d_weight_or_null,
                              float* d_bias_or_null, void* workspace, size_t workspace_bytes, void* stream);

/* out[c] = sum_r x[r,c] (GCNConv bias gradient); out is zeroed by the call. */
int kagnn_column_sums(const float* x, int64_t ldx, int64_t num_rows, int32_t num_cols, float* out, void* stream);

/* y = log_softmax(x) row-wise: dx = dy - exp(y) * sum_c dy. */
int kagnn_log_softmax_bwd(const float* y, int64_t ldy, const float* dy, int64_t ld_dy, int64_t rows, int32_t cols, float* dx,
                          int64_t ld_dx, void* stream);

/* y = silu(x), kept as its own launch when it must be differentiated (the activation between the GCN layers of
 * graph_classification/models.py:190), and dx = dy * silu'(x). */
int kagnn_silu_fwd(const float* x, int64_t ldx, int64_t rows, int32_t cols, float* y, int64_t ldy, void* stream);
int kagnn_silu_bwd(const float* x, int64_t ldx, const float* dy, int64_t ld_dy, int64_t rows, int32_t cols, float* dx,
                   int64_t ld_dx, void* stream);

/* global_add_pool / global_mean_pool backward: dx[n,:] = d_pooled[batch[n],:] (divided by the segment length when mean != 0). */
int kagnn_segment_pool_bwd(const float* d_pooled, int64_t ld_dp, const int32_t* segment_ptr, const int64_t* batch,
                           int64_t num_rows, int32_t num_cols, int32_t mean, float* dx, int64_t ld_dx, void* stream);

/* FastKAN layer (fastkan.py:76-85), z = LayerNorm(x):
 *   dz[n,i] = sum_g phi_g'(z) sum_o dy[n,o] Ws[o,i*G+g]      dx_base[n,i] = silu'(x) sum_o dy[n,o] Wb[o,i]
 * ln_stats = per-row (mean, rstd) from kagnn_layernorm_stats (NULL for a layer without LayerNorm: then dz receives the complete
 * input gradient dz + dx_base and dx_base is ignored).  Reads layer->packed_w, ln_weight, ln_bias. */
int kagnn_rbf_bwd_input(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* ln_stats_or_null, const float* dy,
                        int64_t ld_dy, int64_t num_rows, float* dz, int64_t ld_dz, float* dx_base, int64_t ld_dxb, void* stream);

/* Gradient of the packed FastKAN weights [in][G+1][out_pad4] (slot G = base_linear.weight); zeroed by the call. */
int kagnn_rbf_bwd_weights(const KagnnKanLayer* layer, const float* x, int64_t ldx, const float* ln_stats_or_null, const float* dy,
                          int64_t ld_dy, int64_t num_rows, float* d_packed, void* stream);

/* LayerNorm backward: dx = rstd (dz w - mean_i(dz w) - xhat mean_i(dz w xhat)) [+ dx_base]; d_weight = sum_n dz xhat,
 * d_bias = sum_n dz (either may be NULL; zeroed by the call). */
int kagnn_layernorm_bwd(const float* x, int64_t ldx, const float* ln_stats, const float* ln_weight_or_null, const float* dz,
                        int64_t ld_dz, const float* dx_base_or_null, int64_t ld_dxb, int64_t num_rows, int32_t num_cols, float* dx,
                        int64_t ld_dx, float* d_weight_or_null, float* d_bias_or_null, void* stream);

/* Backward of the GAT attention (autograd of PyG GATConv.forward as used by KAGATConv, nc/models.py:39-46): h (N, heads*C) is the
 * projected input, edge_weight (heads, nnz) / self_weight (heads, N) the coefficients the forward produced (kagnn_gat_edge_softmax).
 * On entry dh holds the aggregation part sum_i alpha_ij d out[i] (the WEIGHTED aggregation over the reversed edges, one launch per
 * head); the call adds the part that flows through the attention scores and writes d att_src / d att_dst (heads*C each).
 * Workspace: kagnn_gat_bwd_workspace(N, nnz, heads). */
size_t kagnn_gat_bwd_workspace(int64_t num_rows, int64_t nnz, int32_t heads);
int kagnn_gat_bwd(const int32_t* rowptr, const int32_t* col, int64_t num_rows, int64_t nnz, int32_t heads, int32_t channels,
                  const float* h, int64_t ldh, const float* dout, int64_t ld_dout, const float* att_src, const float* att_dst,
                  const float* edge_weight, const float* self_weight, float negative_slope, void* workspace, size_t workspace_bytes,
                  float* dh, int64_t ld_dh, float* d_att_src, float* d_att_dst, void* stream);

/* Backward of the GINE aggregation a_i = self_scale x_i + sum_{e: dst_e = i} relu(x_{src_e} + ef_e) (PyG GINEConv,
 * graph_regression/models.py:98) on the COO edge list: edge_index is the contiguous (2, E) int64 tensor, edge_feat has one row
 * per edge in the same order.  dx (num_nodes x num_cols) is overwritten; d_edge_feat (E x num_cols) may be NULL. */
int kagnn_gine_bwd(const float* x, int64_t ldx, const float* edge_feat, int64_t ld_edge, const int64_t* edge_index,
                   int64_t num_edges, int64_t num_nodes, int32_t num_cols, const float* da, int64_t ld_da, float self_scale,
                   float* dx, int64_t ld_dx, float* d_edge_feat_or_null, int64_t ld_de, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KAGNN_B200_H */
