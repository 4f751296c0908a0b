/*
 * gsat_b200.h -- C ABI of libgsat_b200.so, the B200 (sm_100a) implementation of GSAT's per-step
 * stochastic-attention message-passing path.
 *
 * The reference (mihikamd/DP-GSAT) is pure Python: there is no FFI in it.  Each entry point below replaces the
 * chain of PyTorch / torch_geometric / torch_scatter / torch_sparse calls at the cited reference file:line
 * (paths relative to the reference root); the Python package dp_gsat_b200 binds them with ctypes and re-exposes
 * the reference's own surfaces (GINConv, GIN, ExtractorMLP, GSAT.forward_pass, reorder_like, ...).
 * INTEGRATION.md shows the binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller (PyTorch);
 *    the library never allocates, frees or retains memory, and never synchronises the host with the device
 *  - fp32 values, int32 indices internally (N, E < 2^31); the reference's int64 edge_index is consumed once,
 *    by gsatb_index_build
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); calls are CUDA-graph capturable
 *  - return 0 on success, a negative GSATB_E* code otherwise; nothing throws across the ABI
 *  - nullable pointers are marked [nullable]
 *  - feature widths H, C must be multiples of 4 (128-bit vector access) unless stated
 */
#ifndef GSAT_B200_H
#define GSAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSATB_VERSION 100

#define GSATB_OK 0
#define GSATB_EINVAL -1        /* null pointer / negative size / unsupported flag */
#define GSATB_ESHAPE -2        /* unsupported width (H % 4 != 0, H too large, ...) */
#define GSATB_EALIGN -3        /* pointer not 16-byte aligned */
#define GSATB_EWS_TOO_SMALL -4 /* workspace smaller than gsatb_*_workspace() */
#define GSATB_EARCH -5         /* device is not compute capability 10.x */
#define GSATB_ENOT_SYMMETRIC -6
#define GSATB_ELAUNCH -7       /* CUDA launch error (see cudaGetLastError) */

typedef void* gsatb_stream_t; /* cudaStream_t */

int gsatb_version(void);
const char* gsatb_strerror(int code);
/* 0 when the current device is sm_100-class, GSATB_EARCH otherwise (no fallback path exists). */
int gsatb_check_device(void);
/* Optional DEVICE-resident step counter (uint64, owned by the caller, NULL switches it off).  Every counter-based
 * random stream of the library -- the dropout hashes of the tensor-core ops and the sampler's Philox offset -- adds
 * its current value inside the kernel, so that a CUDA graph that captured a whole training step (kernel arguments
 * frozen) still draws fresh randomness on each replay: the caller increments the counter once per step (in-graph).
 * Forward and backward of one step read the same value.  This and gsatb_tc_set_profile_buffer are the only pieces
 * of process-level state in the library. */
int gsatb_set_step_counter(const uint64_t* dev_counter);

/* ------------------------------------------------------------------------------------------------------------
 * K0  index builder.  Replaces, once per batch instead of every step:
 *   torch_geometric.utils.is_undirected            src/run_gsat.py:232,242   example/gsat.py:80
 *   torch_sparse.transpose(coalesced=False)        src/run_gsat.py:243       example/gsat.py:81
 *   reorder_like (sort_edge_index + 2 argsorts)    src/utils/utils.py:19-25
 *   the implicit COO->segment conversions inside MessagePassing.propagate / scatter / InstanceNorm / pooling
 *   int(batch.max())+1 host syncs                  (PyG InstanceNorm, global_add_pool)
 * Canonical orders are the stable ascending (dst,src) and (src,dst) orders; outputs are bit-exact against
 * oracle/gsat_oracle.py::build_index_oracle.
 *   flags[0] = number of rank positions whose (src,dst)/(dst,src) keys differ (0 <=> is_undirected)
 *   flags[1] = number of duplicate directed edges
 *   flags[2] = graph-contiguity violations (batch not sorted, edges not grouped by graph, edge crossing graphs)
 *   flags[3] = out-of-range node / graph ids
 * rev[e] = -1 where no reverse edge matches.
 * ---------------------------------------------------------------------------------------------------------- */
size_t gsatb_index_build_workspace(int64_t N, int64_t E, int64_t G);
int gsatb_index_build(const int64_t* edge_index /* [2,E] */, const int64_t* batch /* [N] */,
                      int64_t N, int64_t E, int64_t G,
                      int32_t* src, int32_t* dst, int32_t* rev,
                      int32_t* rowptr_dst /* [N+1] */, int32_t* eid_by_dst, int32_t* src_by_dst,
                      int32_t* rowptr_src /* [N+1] */, int32_t* eid_by_src, int32_t* dst_by_src,
                      int32_t* node_ptr /* [G+1] */, int32_t* edge_ptr /* [G+1] */,
                      int32_t* node_graph /* [N] */, int32_t* edge_graph /* [E] */,
                      int32_t* flags /* [4] */, void* ws, size_t ws_bytes, gsatb_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K3  attention-weighted GIN aggregation.  Replaces GINConv.forward/message up to (not including) self.nn:
 *   src/models/conv_layers.py:14-34  (+ PyG propagate: index_select, x_j*edge_atten, scatter-add, += (1+eps) x)
 *   out[i] = (sum over CSR row i, in order: att[eid] * x[src]) + (1+eps) * x[i]
 * bwd:  dx[j] = (sum over CSC row j: att[eid] * g[dst]) + (1+eps) g[j] ;  datt[e] = <x[src(e)], g[dst(e)]>
 * att [nullable] = edge_atten None (first GNN pass, run_gsat.py:191).  Deterministic: no atomics.
 * ---------------------------------------------------------------------------------------------------------- */
int gsatb_gin_aggregate_fwd(const float* x, const float* att, const int32_t* rowptr_dst, const int32_t* eid_by_dst,
                            const int32_t* src_by_dst, float eps, float* out, int64_t N, int64_t E, int H,
                            gsatb_stream_t stream);
/* Same forward with the result rounded to bf16 [N,H] (the operand layout of the tensor-core node MLP): saves a quarter
 * of the kernel's traffic and the separate cast pass in bf16 precision mode. */
int gsatb_gin_aggregate_fwd_bf16(const float* x, const float* att, const int32_t* rowptr_dst,
                                 const int32_t* eid_by_dst, const int32_t* src_by_dst, float eps, void* out_bf16,
                                 int64_t N, int64_t E, int H, gsatb_stream_t stream);
int gsatb_gin_aggregate_bwd(const float* gout, const float* x, const float* att, const int32_t* rowptr_src,
                            const int32_t* eid_by_src, const int32_t* dst_by_src, float eps, float* dx,
                            float* datt /* [nullable] */, int64_t N, int64_t E, int H, gsatb_stream_t stream);

/* Attention-aware GINEConv message passing (src/models/conv_layers.py:37-66 over PyG
 * GINEConv):  out[i] = sum_{e: dst(e)=i} relu(x[src(e)] + edge_feat[e]) * att[e] + (1+eps) x[i], with
 * edge_feat = lin(edge_attr) [E,H] computed by the caller.  bwd: t_e = att_e g[dst(e)] 1[x[src] + edge_feat[e] > 0];
 * dx[j] = sum_e t_e + (1+eps) g[j]; dedge_feat[e] = t_e [nullable]; datt[e] = <relu(x[src]+edge_feat[e]), g[dst]>
 * [nullable].  att [nullable].  Deterministic (CSR / CSC walks in edge order). */
int gsatb_gine_aggregate_fwd(const float* x, const float* edge_feat, const float* att, const int32_t* rowptr_dst,
                             const int32_t* eid_by_dst, const int32_t* src_by_dst, float eps, float* out, int64_t N,
                             int64_t E, int H, gsatb_stream_t stream);
int gsatb_gine_aggregate_bwd(const float* gout, const float* x, const float* edge_feat, const float* att,
                             const int32_t* rowptr_src, const int32_t* eid_by_src, const int32_t* dst_by_src, float eps,
                             float* dx, float* dedge_feat, float* datt, int64_t N, int64_t E, int H,
                             gsatb_stream_t stream);

/* Attention-aware LEConv message passing (src/models/conv_layers.py:69-92 over PyG LEConv,
 * the layer of SPMotifNet, src/models/spmotif_gnn.py:20-23,58-63):
 *   out[i] = sum_{e: dst(e)=i} ((a[src(e)] - b[i]) * edge_weight[e]) * att[e] + add[i]
 * with a = lin1(x), b = lin2(x), add = lin3(x) [N,H] computed by the caller; edge_weight [E], att [E], add [nullable]
 * (an absent factor is 1).  bwd: da[j] = sum_{e: src(e)=j} (g[dst(e)] att[e]) edge_weight[e];
 * db[i] = -sum_{e: dst(e)=i} (g[i] att[e]) edge_weight[e]; datt[e] = <(a[src]-b[dst]) edge_weight[e], g[dst]>
 * [nullable]; dedge_weight[e] = <a[src]-b[dst], g[dst] att[e]> [nullable]; d add = g (the caller's).
 * Deterministic (CSR / CSC walks in edge order). */
int gsatb_le_aggregate_fwd(const float* a, const float* b, const float* edge_weight, const float* att,
                           const int32_t* rowptr_dst, const int32_t* eid_by_dst, const int32_t* src_by_dst,
                           const float* add, float* out, int64_t N, int64_t E, int H, gsatb_stream_t stream);
int gsatb_le_aggregate_bwd(const float* gout, const float* a, const float* b, const float* edge_weight,
                           const float* att, const int32_t* rowptr_src, const int32_t* eid_by_src,
                           const int32_t* dst_by_src, const int32_t* rowptr_dst, const int32_t* eid_by_dst, float* da,
                           float* db, float* dedge_weight, float* datt, int64_t N, int64_t E, int H,
                           gsatb_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K5  graph readout.  Replaces global_add_pool / global_mean_pool (src/models/gin.py:34,53, pna.py:47,62).
 * ---------------------------------------------------------------------------------------------------------- */
int gsatb_pool_fwd(const float* x, const int32_t* node_ptr, float* out, int64_t N, int64_t G, int H, int mean,
                   gsatb_stream_t stream);
int gsatb_pool_bwd(const float* gout, const int32_t* node_ptr, const int32_t* node_graph, float* dx, int64_t N,
                   int64_t G, int H, int mean, gsatb_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Per-graph InstanceNorm over contiguous row segments (PyG InstanceNorm, eps=1e-5, no affine, biased variance of
 * the centred values), the norm inside the extractor MLP: src/utils/get_model.py:47-68, src/run_gsat.py:912-915.
 *   y = (x - mean_g) * rstd_g ; saves rstd [G,C] for backward.  Empty segments are skipped.
 * bwd:  gx = rstd * (gy - mean_g(gy) - y * mean_g(gy*y))
 * ---------------------------------------------------------------------------------------------------------- */
int gsatb_segnorm_fwd(const float* x, const int32_t* seg_ptr, float* y, float* rstd, int64_t M, int64_t G, int C,
                      float eps, gsatb_stream_t stream);
int gsatb_segnorm_bwd(const float* gy, const float* y, const float* rstd, const int32_t* seg_ptr, float* gx,
                      int64_t M, int64_t G, int C, gsatb_stream_t stream);

/* Edge feature gather of the extractor: out[e] = [emb[src(e)] , emb[dst(e)]]  (cat(emb[col], emb[row]),
 * src/run_gsat.py:912-914, example/gsat.py:133-135); bwd sums over CSC then CSR rows (deterministic, replaces the
 * index_add_ atomics of index_select's autograd). */
int gsatb_gather_concat_fwd(const float* emb, const int32_t* src, const int32_t* dst, float* out, int64_t E, int H,
                            gsatb_stream_t stream);
int gsatb_gather_concat_bwd(const float* g, const int32_t* rowptr_src, const int32_t* eid_by_src,
                            const int32_t* rowptr_dst, const int32_t* eid_by_dst, float* demb, int64_t N, int H,
                            gsatb_stream_t stream);
/* the same reduction over a bf16 g (fp32 accumulation in the same order) */
int gsatb_gather_concat_bwd_bf16(const void* g_bf16, const int32_t* rowptr_src, const int32_t* eid_by_src,
                                 const int32_t* rowptr_dst, const int32_t* eid_by_dst, float* demb, int64_t N, int H,
                                 gsatb_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Line-graph ("dual") builder of the fork.  Replaces the Python dict loops of
 *   src/datasets/mutag_dual.py:342-378 (group primal edges by first endpoint, all ordered pairs inside a group)
 *   and, with halve = 1, the relabelling of :536-548 (both directions of a primal edge share one dual node)
 * on top of K0's CSC row pointers of the PRIMAL graph and `members` = gsatb_stable_order(src): the edges grouped by
 * source node, each group in primal edge order (K0's own eid_by_src orders a group by destination).  One dual node per directed primal edge e (halve: per edge pair e >> 1);
 * groups in order of first appearance of their source node; inside a group with members m_0 < m_1 < ... the dual
 * edges (m_i, m_j), (m_j, m_i) for i < j, i outer.  Two calls because E_d = sum_v d(v)(d(v)-1) is data dependent:
 *   count: offs[e] [E] (int64 exclusive prefix of the group sizes, laid out by first appearance), total[0] = E_d
 *   fill : dual_edge_index int64 [2, Ed], dual_batch int64 [E] (halve: [E/2]) = node_graph[src[.]]   [nullable]
 * Bit-exact against oracle/gsat_oracle.py::line_graph_dual.
 * ---------------------------------------------------------------------------------------------------------- */
/* order[p] = id of the p-th edge in ascending `keys` order, ties in ascending id order (stable LSD radix sort of K0). */
size_t gsatb_stable_order_workspace(int64_t E);
int gsatb_stable_order(const int32_t* keys, int64_t E, int64_t key_range, int32_t* order, void* ws, size_t ws_bytes,
                       gsatb_stream_t stream);
size_t gsatb_line_graph_workspace(int64_t N, int64_t E);
int gsatb_line_graph_count(const int32_t* src, const int32_t* rowptr_src, const int32_t* members, int64_t N, int64_t E,
                           int64_t* offs, int64_t* total, void* ws, size_t ws_bytes, gsatb_stream_t stream);
int gsatb_line_graph_fill(const int32_t* src, const int32_t* rowptr_src, const int32_t* members, const int64_t* offs,
                          const int64_t* node_graph, int64_t N, int64_t E, int halve, int64_t* dual_edge_index,
                          int64_t Ed, int64_t* dual_batch, gsatb_stream_t stream);

/* On-device explanation metric: GSAT.get_precision_at_k, src/run_gsat.py:783-791 (a Python
 * loop over graphs with boolean masks over all edges and a numpy argsort each, on .cpu() copies).
 * precision[g] = (sum of exp_labels over the k highest-attention edges of graph g) / k, edges of a graph contiguous
 * (edge_ptr from K0); ties go to the smaller edge index (unspecified in the reference: numpy's default argsort is not
 * stable).  Graphs with fewer than k edges contribute all their edges and still divide by k, as the reference does. */
int gsatb_precision_at_k(const float* att, const float* exp_labels, const int32_t* edge_ptr, int64_t G, int k,
                         float* precision /* [G] */, gsatb_stream_t stream);

/* Weight / bias gradient of a Linear layer with a small input width (F + 1 <= 16): the node encoder Linear(x_dim, H)
 * of src/models/gin.py:22-25 / pna.py:20-25.  dW[h,f] = sum_n g[n,h] x[n,f], db[h] = sum_n g[n,h] (db nullable);
 * replaces the library's large-K fp32 sgemm in autograd; deterministic (per-CTA partials reduced in a fixed order). */
size_t gsatb_linear_small_dw_workspace(int64_t N, int H, int F);
int gsatb_linear_small_dw(const float* g, const float* x, float* dW /* [H,F] */, float* db /* [H] */, int64_t N, int H,
                          int F, void* ws, size_t ws_bytes, gsatb_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K2  concrete (Gumbel-sigmoid) sampling + undirected reverse-edge average + information loss.  Replaces
 *   concrete_sample / sampling          src/run_gsat.py:866-885   example/gsat.py:94-103
 *   (att + att[rev]) / 2                src/run_gsat.py:241-247   example/gsat.py:79-85
 *   info loss (KL to Bernoulli(r))      src/run_gsat.py:126-132   example/gsat.py:30-31
 * mode bits: 1 TRAINING (add logistic noise log u - log(1-u)); 2 AVERAGE (needs rev); 4 INFO_ON_EDGE_ATT (fork:
 * loss on the averaged attention; default upstream: on att); 8 NO_INFO (skip the reduction)
 * noise_u [nullable]: injected uniform draw in [1e-10, 1-1e-10]; when null and TRAINING, a counter-based
 * Philox4x32-10 stream (seed, offset + e) is used and can be regenerated in backward.
 * r_tensor [nullable]: per-edge r (fork: sigmoid(dual logits).detach()); else r_scalar.
 * info_mean[0] = mean_e f(a_e).  The reduction is two-stage with a fixed order (deterministic).
 * ---------------------------------------------------------------------------------------------------------- */
#define GSATB_MODE_TRAINING 1
#define GSATB_MODE_AVERAGE 2
#define GSATB_MODE_INFO_ON_EDGE_ATT 4
#define GSATB_MODE_NO_INFO 8
size_t gsatb_sample_workspace(int64_t E);
int gsatb_sample_avg_info_fwd(const float* logit, const float* noise_u, const int32_t* rev, const float* r_tensor,
                              float r_scalar, float temp, int mode, uint64_t seed, uint64_t offset, float* att,
                              float* edge_att, float* info_mean, int64_t E, void* ws, size_t ws_bytes,
                              gsatb_stream_t stream);
/* g_att, g_edge_att [nullable] upstream gradients; g_info: device scalar d loss / d info_mean [nullable]. */
int gsatb_sample_avg_info_bwd(const float* g_att, const float* g_edge_att, const float* g_info, const float* att,
                              const float* edge_att, const int32_t* rev, const float* r_tensor, float r_scalar,
                              float temp, int mode, float* dlogit, int64_t E, gsatb_stream_t stream);

/* out[e] = v[rev[e]]  (reorder_like(transpose(ei, v), ei, v), src/utils/utils.py:19-25); C values per edge. */
int gsatb_gather_rev(const float* v, const int32_t* rev, float* out, int64_t E, int C, gsatb_stream_t stream);

/* General reorder_like: position p of the stable (row,col) order of `to` is matched with position p of the stable
 * (row,col) order of `from` (src/utils/utils.py:20-22): map[order_to[p]] = order_from[p]; mismatch[0] counts
 * positions whose (row,col) differ (the reference raises ValueError when any does, utils.py:23-24). */
int gsatb_match_orders(const int32_t* order_to, const int32_t* order_from, const int32_t* src_to,
                       const int32_t* dst_to, const int32_t* src_from, const int32_t* dst_from, int32_t* map,
                       int32_t* mismatch, int64_t E, gsatb_stream_t stream);

/* K2' node->edge lift: edge_att[e] = a[src] * a[dst]  (src/run_gsat.py:870-875); bwd is row-parallel over
 * CSR + CSC (deterministic). */
int gsatb_lift_fwd(const float* node_att, const int32_t* src, const int32_t* dst, float* edge_att, int64_t E,
                   gsatb_stream_t stream);
int gsatb_lift_bwd(const float* g_edge, const float* node_att, const int32_t* rowptr_dst, const int32_t* eid_by_dst,
                   const int32_t* src_by_dst, const int32_t* rowptr_src, const int32_t* eid_by_src,
                   const int32_t* dst_by_src, float* d_node, int64_t N, gsatb_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K4  PNA multi-aggregator message passing.  Replaces PNAConvSimple.message/aggregate
 * (src/models/conv_layers.py:160-185; aggregators :193-226 on torch_scatter sum/mean/min/max):
 *   m_e = cat(x[dst(e)], x[src(e)], [edge_feat_e]) * att_e   (F = 2H + He),   out [N, n_aggs * F] in the order of
 *   agg_codes (0 sum, 1 mean, 2 min, 3 max, 4 var, 5 std = sqrt(relu(var) + 1e-5)); empty rows give 0 (std:
 *   sqrt(1e-5)), as torch_scatter does.  The variance is accumulated on the centred values (second walk over the
 *   row) instead of the reference's E[m^2]-E[m]^2, which cancels in fp32.  stat_mean / stat_msq (= that variance)
 *   [N,F] and argmin / argmax [N,F] (edge ids, ties to the smallest id) are saved for backward.  edge_feat [E,He] and att [E] are nullable.  Scalers are applied by the
 *   caller (the reference configs all use `scalers: false`).
 * bwd: dx [N,H] (x_i part by destination, x_j part by source -- two deterministic row-parallel passes),
 *      dedge_feat [E,He] [nullable], datt [E] [nullable].
 * ---------------------------------------------------------------------------------------------------------- */
int gsatb_pna_aggregate_fwd(const float* x, const float* edge_feat, const float* att, const int32_t* rowptr_dst,
                            const int32_t* eid_by_dst, const int32_t* src_by_dst, const int* agg_codes, int n_aggs,
                            float* out, float* stat_mean, float* stat_msq, int32_t* argmin, int32_t* argmax, int64_t N,
                            int64_t E, int H, int He, gsatb_stream_t stream);
int gsatb_pna_aggregate_bwd(const float* gout, const float* x, const float* edge_feat, const float* att,
                            const int32_t* rowptr_dst, const int32_t* eid_by_dst, const int32_t* src_by_dst,
                            const int32_t* rowptr_src, const int32_t* eid_by_src, const int32_t* dst_by_src,
                            const int* agg_codes, int n_aggs, const float* stat_mean, const float* stat_msq,
                            const int32_t* argmin, const int32_t* argmax, float* dx, float* dedge_feat, float* datt,
                            int64_t N, int64_t E, int H, int He, gsatb_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Dense layers on the tensor cores (tcgen05.mma kind::f16: bf16 operands, fp32 accumulation in TMEM; weights
 * streamed by TMA, activations staged through shared memory by producer warps with the prologue fused).
 * Replaces the nn.Linear / BatchNorm1d / ReLU chains of src/models/gin.py:55-62 and the Linear layers of the
 * extractor MLP (src/utils/get_model.py:57-68).  Tolerance: bf16 operand rounding (documented in the tests).
 *
 * gsatb_tc_prep_weight: fp32 W [OUT,K] (or its transpose) -> zero-padded bf16 [pad128(rows), pad64(cols)],
 * the layout the TMA descriptor of every tc op expects.
 * gsatb_tc_linear_fwd: out = act(pro(x) W^T + bias); pro(x) = relu(x*in_scale + in_shift) when in_scale is
 * given (BatchNorm+ReLU folded into the operand load), identity otherwise; act = ReLU when relu_out.
 * With stat_partials (OUT <= 128): stats[0:OUT] = sum_rows z, stats[OUT:2OUT] = sum_rows z^2 of the pre-activation
 * z (BatchNorm batch statistics), reduced in a fixed order in fp64.  pdrop > 0 applies dropout after the
 * activation (drop_mask [rows,OUT] uint8 keep mask, or the counter hash of drop_seed when NULL).
 * ---------------------------------------------------------------------------------------------------------- */
int gsatb_tc_prep_weight(const float* w, int OUT, int K, int transpose, void* w_bf16_padded, gsatb_stream_t stream);
size_t gsatb_tc_stat_partials_elems(int OUT);
/* Development aid: device buffer [148][16] int64 (zeroed by the caller) that the tensor-core kernels fill with
 * per-role cycle counters (see csrc/tc_ops.cu); NULL switches it off (the default). */
int gsatb_tc_set_profile_buffer(void* buf);
int gsatb_tc_linear_fwd(const void* x, int x_is_bf16, int ldx, const float* in_scale, const float* in_shift, const void* w_bf16_padded,
                        const float* bias, float* out, int ldo, int relu_out, float* stat_partials, double* stats,
                        const uint8_t* drop_mask, uint64_t drop_seed, float pdrop, int64_t rows, int K, int OUT,
                        gsatb_stream_t stream);

/* gsatb_tc_linear_bf16_fwd: out = drop(act(x_bf16 W^T + bias)) with the B operand loaded by TMA straight from the
 * row-major bf16 activations (both Linears of the GIN node MLP: K3's bf16 aggregation, then a1); out is bf16
 * (out_is_bf16) or fp32; optional BatchNorm statistics as above, taken from the fp32 accumulators before rounding.
 * posmask_out [nullable]: uint32 [rows, ceil(OUT/32)], bit c%32 of word (row, c/32) = (out[row, c] > 0): the only
 * thing the backward pass needs of the layer output h (gsatb_tc_gin_bwd2 then reads 4 bytes instead of 128 per row).
 * gsatb_bn_relu_bf16: a1 = ReLU(z1 * scale + shift) (BatchNorm folded per channel), bf16 in / bf16 out. */
int gsatb_tc_linear_bf16_fwd(const void* x_bf16, int ldx, const void* w_bf16_padded, const float* bias, void* out,
                             int out_is_bf16, int ldo, int relu_out, float* stat_partials, double* stats,
                             const uint8_t* drop_mask, uint64_t drop_seed, float pdrop, uint32_t* posmask_out,
                             int64_t rows, int K, int OUT, gsatb_stream_t stream);
int gsatb_bn_relu_bf16(const void* z_bf16, const float* scale, const float* shift, void* a_bf16, int64_t rows, int C,
                       gsatb_stream_t stream);

/* GIN node MLP backward (autograd of src/models/gin.py:55-62 + the ReLU / Dropout of gin.py:50-52):
 *   gin_bwd2: d2 = dh*(h>0)*drop_scale (written as bf16 [N,H]; the sign of h from posmask when given, else from h); da1 = d2 W2 on tcgen05 (w2t = prep(W2, transpose));
 *             g = da1*(ReLU(BN(z1))>0) and a1 = ReLU(BN(z1)) written as bf16 [N,H1]; stats[0:H1] = sum_rows g,
 *             stats[H1:2H1] = sum_rows g*xhat (BatchNorm backward), deterministic two-stage reduction
 *   gin_bwd1: dz1 = cA*g + cB*z1 + cC (BatchNorm backward folded per channel; written as bf16 [N,H1]);
 *             dx = dz1 W1 on tcgen05 (w1t = prep(W1, transpose)), fp32 [N,Kin] */
int gsatb_tc_gin_bwd2(const float* dh, const float* h /* [nullable if posmask] */, const uint32_t* posmask /* [nullable] */,
                      float drop_scale, const void* w2t_bf16_padded, const void* z1,
                      const float* bn_scale, const float* bn_shift, const float* mean, const float* rstd, void* d2,
                      void* g, void* a1, float* stat_partials, float* stats, int64_t N, int H, int H1,
                      gsatb_stream_t stream);
int gsatb_tc_gin_bwd1(const void* g, const void* z1, const float* cA, const float* cB, const float* cC,
                      const void* w1t_bf16_padded, void* dz1, float* dx, int64_t N, int H1, int Kin,
                      gsatb_stream_t stream);

/* GIN node MLP in the row-owner orientation (csrc/gin_rows.cu) for K = H1 = H in {64, 128} -- the shape of every GIN
 * layer of the reference (src/models/gin.py:28-35: MLP(hidden, hidden)).  Same contracts as the entries above, one
 * kernel per GEMM with every tile moved by TMA in both directions:
 *   gin_rows_lin1: z1 = x W1^T + b1 (bf16 [rows,H]) + BatchNorm batch statistics stats[0:H] = sum z, stats[H:2H] = sum z^2
 *                  from the fp32 accumulators (stat_partials: gsatb_tc_stat_partials_elems(H) floats; both nullable)
 *   gin_rows_lin2: a1 = ReLU(z1*scale + shift) formed in shared memory (stored as bf16 [rows,H] when a1_bf16 is given:
 *                  the operand of dW2), h = Dropout(ReLU(a1 W2^T + b2)) fp32 [rows,H], posmask_out as in
 *                  gsatb_tc_linear_bf16_fwd (same dropout word stream: identical masks for identical seeds)
 *   gin_rows_bwd1: dz1 = cA*g + cB*z1 + cC formed in shared memory (stored as bf16 when dz1_bf16 is given), dx = dz1 W1
 *                  fp32 [rows,H]   (autograd of gin.py:55-62; replaces gsatb_tc_gin_bwd1 at these shapes)
 *   gin_rows_bwd2: d2 = dh*(h>0)*drop_scale (sign bits from posmask; formed in place in the TMA-loaded fp32 tile, stored
 *                  as bf16 [rows,H]), da1 = d2 W2, g = da1*(ReLU(BN(z1))>0) as bf16 [rows,H]; stats[0:H] = sum_rows g,
 *                  stats[H:2H] = sum_rows g*xhat, taken from the bf16-rounded g that is stored, with a fixed-order
 *                  two-stage reduction (stat_partials: gsatb_gin_rows_stat_partials_elems(H) floats)   (replaces
 *                  gsatb_tc_gin_bwd2 at these shapes)
 * gsatb_gin_rows_supported -> 1 when the three widths qualify. */
int gsatb_gin_rows_supported(int K, int H1, int H);
size_t gsatb_gin_rows_stat_partials_elems(int H);
int gsatb_gin_rows_bwd2(const float* dh, const uint32_t* posmask, float drop_scale, const void* w2t_bf16_padded,
                        const void* z1_bf16, const float* bn_scale, const float* bn_shift, const float* mean,
                        const float* rstd, void* d2_bf16, void* g_bf16, float* stat_partials, float* stats, int64_t rows,
                        int H, gsatb_stream_t stream);
int gsatb_gin_rows_lin1(const void* x_bf16, const void* w1_bf16_padded, const float* bias /* [nullable] */, void* z1_bf16,
                        float* stat_partials /* [nullable] */, double* stats /* [nullable] */, int64_t rows, int H,
                        gsatb_stream_t stream);
int gsatb_gin_rows_lin2(const void* z1_bf16, const float* bn_scale, const float* bn_shift, const void* w2_bf16_padded,
                        const float* bias /* [nullable] */, void* a1_bf16 /* [nullable] */, float* h,
                        uint32_t* posmask_out /* [nullable] */, const uint8_t* drop_mask /* [nullable] */,
                        uint64_t drop_seed, float pdrop, int64_t rows, int H, gsatb_stream_t stream);
int gsatb_gin_rows_bwd1(const void* g_bf16, const void* z1_bf16, const float* cA, const float* cB, const float* cC,
                        const void* w1t_bf16_padded, void* dz1_bf16 /* [nullable] */, float* dx, int64_t rows, int H,
                        gsatb_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K1  extractor MLP forward, fused on the tensor cores.  Replaces ExtractorMLP.forward (src/run_gsat.py:909-927,
 * example/gsat.py:131-139) and the MLP/InstanceNorm stack of src/utils/get_model.py:47-68.
 * Tiles are graph-aligned: gsatb_tile_plan_host packs consecutive graphs into tiles of <= 128 rows / <= 32 graphs
 * from the HOST copy of edge_ptr (node_ptr in node mode); it returns GSATB_ESHAPE when one graph alone exceeds a
 * tile (the caller then uses the unfused segnorm path).
 *   make_f12: f12 = [emb[src] | emb[dst]] gathered on the fly and rounded to bf16 [rows, K = 2H] (src == NULL: node
 *             mode, f12 = emb rows, K = H) (gsatb_tc_ext_make_f12, declared with the backward ops: f12 is also the
 *             operand of dW1 = dz1^T f12)
 *   ext_fwd1: xhat1 = InstanceNorm(f12 W1^T) stored as bf16 [rows, C1]; rstd1 [G, C1]; f12 is the TMA-fed B operand
 *   make_h1 : h1 = Dropout(ReLU(xhat1)) as bf16 [rows, C1] (gsatb_tc_ext_make_h1, declared with the backward ops:
 *             h1 is also the operand of dW2 = dz2^T h1).  mask1 [rows, C1] (uint8 keep mask) is an optional injected
 *             dropout mask; otherwise a counter hash of (seed, element index) decides, and backward regenerates it.
 *   ext_fwd2: logit = (Dropout(ReLU(InstanceNorm(h1 W2^T))) . w3) + b3 with h1 as the TMA-fed B operand (four epilogue
 *             groups); also xhat2 bf16 [rows, H] and rstd2 [G, H] for backward.  mask2 [rows, H] optional as mask1.
 * ---------------------------------------------------------------------------------------------------------- */
int gsatb_tile_plan_host(const int32_t* seg_ptr_host, int64_t G, int max_rows, int max_seg, int32_t* tile_row,
                         int32_t* tile_seg, int32_t* num_tiles);
int gsatb_tc_ext_fwd1(const void* f12, const void* w1_bf16_padded, const int32_t* tile_row, const int32_t* tile_seg,
                      const int32_t* seg_ptr, int num_tiles, void* xhat1, float* rstd1, int64_t rows, int K, int C1,
                      float eps, gsatb_stream_t stream);
int gsatb_tc_ext_fwd2(const void* h1, const void* w2_bf16_padded, const float* w3, const float* b3,
                      const uint8_t* mask2, uint64_t seed, float pdrop, int training,
                      const int32_t* tile_row, const int32_t* tile_seg, const int32_t* seg_ptr, int num_tiles,
                      void* xhat2, float* rstd2, float* logit, int64_t rows, int C1, int H, float eps,
                      gsatb_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K1 backward (autograd of the extractor MLP).  seg_ptr / tiles as in the forward; seed, pdrop, training and the
 * optional injected masks must equal the forward call's so that the dropout masks are regenerated identically.
 *   ext_bwd_head : dlogit [rows] -> dz2 bf16 [rows,H] (through w3, Dropout2, ReLU2, InstanceNorm2);
 *                  dw3_part [G,H] per-graph partial sums of d w3 (summed by the caller)
 *   ext_bwd1     : dh1 = dz2 W2 on tcgen05 (w2t = gsatb_tc_prep_weight(W2, transpose=1)), epilogue = Dropout1 /
 *                  ReLU1 masks + InstanceNorm1 backward -> dz1 bf16 [rows,C1]
 *   linear_bf16in: out fp32 [rows,OUT] = x_bf16 [rows,K] W^T   (d f12 = dz1 W1 with w = prep(W1, transpose=1))
 *   make_h1/f12  : bf16 re-materialisation of Dropout(ReLU(xhat1)) and of the gathered input rows, the right-hand
 *                  operands of the weight-gradient GEMMs dW2 = dz2^T h1, dW1 = dz1^T f12 (plain library GEMMs)
 * ---------------------------------------------------------------------------------------------------------- */
int gsatb_tc_ext_bwd_head(const float* dlogit, const void* xhat2, const float* rstd2, const float* w3,
                          const int32_t* seg_ptr, const uint8_t* mask2, uint64_t seed, float pdrop, int training,
                          void* dz2, float* dw3_part, int64_t rows, int64_t G, int H, gsatb_stream_t stream);
int gsatb_tc_ext_bwd1(const void* dz2, const void* w2t_bf16_padded, const void* xhat1, const float* rstd1,
                      const uint8_t* mask1, uint64_t seed, float pdrop, int training, const int32_t* tile_row,
                      const int32_t* tile_seg, const int32_t* seg_ptr, int num_tiles, void* dz1, int64_t rows, int H,
                      int C1, gsatb_stream_t stream);
int gsatb_tc_linear_bf16in_fwd(const void* x_bf16, int ldx, const void* w_bf16_padded, float* out, int ldo,
                               int64_t rows, int K, int OUT, gsatb_stream_t stream);
int gsatb_tc_ext_make_h1(const void* xhat1, const uint8_t* mask1, uint64_t seed, float pdrop, int training, void* h1,
                         int64_t rows, int C1, gsatb_stream_t stream);
int gsatb_tc_ext_make_f12(const float* emb, const int32_t* src, const int32_t* dst, void* f12, int64_t rows, int H,
                          gsatb_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * K1 fused: the whole extractor MLP of src/run_gsat.py:909-927 (ExtractorMLP.forward: f12 = cat(emb[col], emb[row]),
 * col, row = edge_index) + src/utils/get_model.py:57-68 (Linear -> InstanceNorm -> ReLU -> Dropout, twice, then
 * Linear(H, 1)) as ONE persistent tcgen05 kernel per direction; no [rows, 4H] or [rows, 2H] tensor reaches HBM in the
 * forward pass.
 *
 * gsatb_ext_tile_plan: packs consecutive graphs into tiles of <= max_slots = gsatb_ext_tile_slots(H, edge_mode) slots
 *   (128, or 112 when 2H > 128; every graph padded to a multiple of 8 slots) and <= 16 graphs, on the device.  seg_ptr = edge_ptr (edge mode) or node_ptr (node mode) of the GraphIndex;
 *   tile_seg needs G + 1 entries; out2[0] = number of tiles T, out2[1] = graphs with more than max_slots rows (the fused
 *   kernels must not be run on such a batch: the caller takes the unfused tensor-core path).
 * gsatb_ext_fused_fwd: logit[r] for every row (edge, or node when src == dst == NULL).  w1 / w2 come from
 *   gsatb_tc_prep_weight.  The per-graph mean is removed from the gathered rows in fp32 BEFORE the bf16 rounding (the
 *   Linear biases in front of an InstanceNorm cancel exactly, so b1 / b2 are not read); the mean comes from the graph's
 *   contiguous node rows weighted by out- / in-degree (node_ptr, rowptr_src, rowptr_dst of the GraphIndex).  mask1 [rows, C1] / mask2
 *   [rows, H] inject dropout masks (parity tests); otherwise masks come from (seed, step counter, row, channel) and the
 *   effective seeds are written to seed_out[2] for the backward pass.  xhat2t (nullable): the InstanceNorm-2 output as
 *   bf16 in SLOT space, tile-major [ld_slots / 128 tiles][pad128(H) channels][128 slots] (one contiguous block per tile;
 *   channel rows >= H must be zero-initialised by the caller when H % 128 != 0) -- what the backward needs of the
 *   forward besides the logits.  max_tiles >= T bounds the grid (pass G).  H % 8 == 0, H <= 128. */
int gsatb_ext_tile_slots(int H, int edge_mode);
int gsatb_ext_tile_plan(const int32_t* seg_ptr, int64_t G, int max_slots, int32_t* tile_seg, int32_t* out2,
                        gsatb_stream_t stream);
int gsatb_ext_fused_fwd(const float* emb, const int32_t* src /* [nullable] */, const int32_t* dst /* [nullable] */,
                        const int32_t* node_ptr /* [nullable: node mode] */, const int32_t* rowptr_src /* [nullable: node mode] */,
                        const int32_t* rowptr_dst /* [nullable: node mode] */, const int32_t* seg_ptr,
                        const int32_t* tile_seg, const int32_t* num_tiles_dev, int max_tiles, int max_slots,
                        const void* w1_bf16_padded, const void* w2_bf16_padded, const float* w3,
                        const float* b3 /* [nullable] */, const uint8_t* mask1 /* [nullable] */,
                        const uint8_t* mask2 /* [nullable] */, uint64_t seed, float pdrop, int training, float* logit,
                        void* xhat2t /* [nullable] */, int64_t ld_slots, float* rstd2 /* [G, H], [nullable] */,
                        void* xs /* [ld_slots, pad64(Kin)] bf16, [nullable] */, uint32_t* seed_out /* [nullable] */, int64_t rows, int H, int C1, float eps, gsatb_stream_t stream);

/* gsatb_ext_fused_bwd: backward of gsatb_ext_fused_fwd over the same tile plan (autograd of src/run_gsat.py:909-927 +
 *   src/utils/get_model.py:57-68 at loss.backward(), src/run_gsat.py:634).  GEMM1 is recomputed per tile from xs, the
 *   centred bf16 input tiles the forward stored (TMA) in slot space [ld_slots = T * 128, pad64(Kin)]; the other inputs are
 *   d logit [rows], xhat2t / rstd2 / seeds as the forward wrote them, W2^T and W1^T from
 *   gsatb_tc_prep_weight(transpose = 1).  Outputs: d f12 [rows, Kin] fp32 (Kin = 2H, or H in node mode: then it IS d emb;
 *   edge mode: reduce with gsatb_gather_concat_bwd) -- or, with df12_is_bf16, the same tensor rounded to bf16 (edge mode of
 *   the bf16 precision mode: half the bytes of the step's largest intermediate; reduce with gsatb_gather_concat_bwd_bf16), dw3_part [min(max_tiles, 148) * 2, H] partial sums of d w3 (add
 *   them up), and the bf16 operands of the weight-gradient products in tile-major slot space, written with TMA stores:
 *   dz2t [tiles][pad128(H)][128], dz1t and h1t [tiles][pad128(C1)][128] (layout code 2 of gsatb_tc_dw: every 128-channel x
 *   64-slot box is a 32 KiB-local access instead of 128-byte pieces a whole row apart), so that dW2 = gsatb_tc_dw(dz2t, h1t) and
 *   dW1 = gsatb_tc_dw(dz1t, xs) with rows = ld_slots (slots outside the graphs are zero in dz1t / dz2t / h1t; xs must have
 *   been allocated zero-filled: rows [128 t + max_slots, 128 t + 128) are never written).  d b1 = d b2 = 0 exactly (the
 *   biases cancel in the InstanceNorms); d b3 = sum(d logit). */
int gsatb_ext_fused_bwd(const int32_t* seg_ptr, const int32_t* tile_seg, const int32_t* num_tiles_dev, int max_tiles,
                        int max_slots, int edge_mode, const void* w1_bf16_padded, const void* w2t_bf16_padded,
                        const void* w1t_bf16_padded, const float* w3, const float* dlogit, const void* xhat2t,
                        const float* rstd2, const void* xs, const uint8_t* mask1 /* [nullable] */,
                        const uint8_t* mask2 /* [nullable] */, const uint32_t* seeds /* [nullable] */, float pdrop, int training,
                        void* dz2t, void* dz1t, void* h1t, void* df12, int df12_is_bf16, float* dw3_part, int64_t ld_slots,
                        int64_t rows, int H, int C1, float eps, gsatb_stream_t stream);

/* gsatb_tc_dw: weight / bias gradients on the tensor cores:  dW[m, n] = sum_r A[r, m] * B[r, n],  db[m] = sum_r A[r, m]
 * (autograd of the Linear layers of src/utils/get_model.py:57-68 and src/models/gin.py:55-62, reached through
 * loss.backward() at src/run_gsat.py:634; round 1 ran them as library GEMMs).  A and B are bf16 activations, each either
 * row-major [rows, C] (layout 0, ld = elements per row), channel-major [C, rows] (layout 1, ld = elements per channel) or
 * tile-major [rows / 128][ld channels][128 rows] (layout 2, ld = channel rows per block, a multiple of 128): TMA loads them as
 * MN-major / K-major SWIZZLE_128B operand tiles, so no transposed copy is made.  Split-K over the rows with a
 * fixed-order reduction of the per-CTA partials: run-to-run deterministic.  dW is row-major [M, ldo] fp32 (the layout of
 * nn.Linear.weight.grad for A = dz, B = layer input); db [M] nullable; accumulate != 0 adds to dW / db in place.
 * workspace: gsatb_tc_dw_workspace(rows, M, N) bytes. */
size_t gsatb_tc_dw_workspace(int64_t rows, int M, int N);
int gsatb_tc_dw(const void* a_bf16, int a_channel_major, int64_t lda, const void* b_bf16, int b_channel_major,
                int64_t ldb, int64_t rows, int M, int N, float* dW, int ldo, float* db /* [nullable] */, int accumulate,
                void* workspace, size_t ws_bytes, gsatb_stream_t stream);

/* Dense-layer helpers around the tcgen05 GEMM kernels (csrc/dense.cu): every nn.Linear / BatchNorm1d of the path --
 * src/models/gin.py:22-25 (encoders), :55-62 (node MLP), :42 (fc_out); src/models/pna.py:20-50; src/utils/get_model.py:57-68;
 * src/models/conv_layers.py:49,77-79,149 (GINE / LE / PNA post_nn Linears) -- runs on this library's kernels in both
 * precision modes, forward and backward (autograd of F.linear / F.batch_norm at src/run_gsat.py:634).
 *
 * gsatb_split_bf16: fp32 x [rows, C] (ldx floats per row) -> bf16 GEMM operand.  nseg == 1: plain round-to-nearest
 * (precision 'bf16').  nseg == 6: strict mode (precision 'fp32'): x = h + m + l exactly with three bf16 parts, laid out as
 * six K-segments so that ONE bf16 GEMM with fp32 accumulation sums the six significant partial products (A side
 * pattern l,m,h,m,h,h; B side h,m,l,h,m,h -- smallest products first; pattern = 2 bits per segment, 0 = h, 1 = m, 2 = l, segment 0 in the low bits).
 * layout 0: out [rows, ld_out], segments along the feature axis (columns >= nseg*C zero-filled); layout 1: out
 * [nseg*rows, ld_out], segments stacked along the row axis (columns >= C zero-filled).  ld_out % 8 == 0. */
int gsatb_split_bf16(const float* x, int64_t rows, int C, int64_t ldx, int nseg, int pattern, int layout, void* out_bf16,
                     int64_t ld_out, gsatb_stream_t stream);
/* workspace bytes of the column reductions below */
size_t gsatb_col_workspace(int64_t rows, int C);
/* out[c] = sum_r x[r, c] (bias gradient of a Linear), fp64 accumulation in a fixed order */
int gsatb_colsum(const float* x, int64_t rows, int C, int64_t ld, float* out, void* workspace, size_t ws_bytes,
                 gsatb_stream_t stream);
/* BatchNorm1d (src/models/gin.py:59, pna.py:45) training-mode batch statistics of x [rows, C]: mean [C], rstd [C] =
 * 1/sqrt(biased var + eps); running_mean / running_var [nullable, both or neither] updated in place with `momentum` and
 * the unbiased variance as torch does; sums64 [nullable, 2C doubles] receives sum(x - x[0]), sum((x - x[0])^2). */
int gsatb_bn_stats(const float* x, int64_t rows, int C, float eps, float momentum, float* running_mean /* [nullable] */,
                   float* running_var /* [nullable] */, float* mean, float* rstd, double* sums64 /* [nullable] */,
                   void* workspace, size_t ws_bytes, gsatb_stream_t stream);
/* y = (x - mean) * rstd * gamma + beta, optionally followed by ReLU (gamma / beta nullable = 1 / 0) */
int gsatb_bn_apply(const float* x, const float* mean, const float* rstd, const float* gamma /* [nullable] */,
                   const float* beta /* [nullable] */, int relu, float* y, int64_t rows, int C, gsatb_stream_t stream);
/* dbeta[c] = sum_r g, dgamma[c] = sum_r g * xhat with g = dy (zeroed where y_relu <= 0 when y_relu is given) */
int gsatb_bn_bwd_stats(const float* dy, const float* x, const float* y_relu /* [nullable] */, const float* mean,
                       const float* rstd, int64_t rows, int C, float* dbeta, float* dgamma, double* sums64 /* [nullable] */,
                       void* workspace, size_t ws_bytes, gsatb_stream_t stream);
/* dx = gamma * rstd * (g - sum_g * inv_n - xhat * sum_gx * inv_n) in training mode, gamma * rstd * g in eval mode */
int gsatb_bn_bwd_apply(const float* dy, const float* x, const float* y_relu /* [nullable] */, const float* mean,
                       const float* rstd, const float* gamma /* [nullable] */, const float* sum_g /* [nullable] */,
                       const float* sum_gx /* [nullable] */, float inv_n, int training, float* dx, int64_t rows, int C,
                       gsatb_stream_t stream);

/* BatchNorm1d folding for the tensor-core GIN layer (src/models/gin.py:59 inside the node MLP of :55-62): one launch each way
 * instead of a chain of elementwise library kernels.  n = rows the statistics span: n_dev (device double, e.g. the
 * all-reduced count of the data-parallel group) when given, else n_host.
 * fwd: stats = [sum z, sum z^2] (2C doubles, from gsatb_tc_linear_bf16_fwd) -> mean, rstd, scale = gamma * rstd,
 *      shift = beta - mean * scale; training != 0 also updates running_mean / running_var (momentum, unbiased variance)
 *      and num_batches_tracked += 1; training == 0 folds the running statistics (stats unused).
 * bwd: dz1 = cA * g + cB * z1 + cC from sum_g = sum g, sum_gx = sum g * xhat (C floats each). */
int gsatb_bn_fold_fwd(const double* stats /* [nullable] */, double n_host, const double* n_dev /* [nullable] */,
                      const float* gamma /* [nullable] */, const float* beta /* [nullable] */, float eps, float momentum,
                      float* running_mean /* [nullable] */, float* running_var /* [nullable] */,
                      int64_t* num_batches_tracked /* [nullable] */, int training, float* mean, float* rstd, float* scale,
                      float* shift, int C, gsatb_stream_t stream);
int gsatb_bn_fold_bwd(const float* sum_g /* [nullable] */, const float* sum_gx /* [nullable] */, double n_host,
                      const double* n_dev /* [nullable] */, const float* gamma /* [nullable] */,
                      const float* mean /* [nullable] */, const float* rstd, int training, float* cA, float* cB, float* cC,
                      int C, gsatb_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * the step BEFORE the path -- feature encoders and batch collate on the device.
 *
 * Fused categorical encoders.  Replace ogb AtomEncoder / BondEncoder as called by src/models/gin.py:22-25,45-47 and
 * pna.py:20-23,53-55 (K embedding gathers + K-1 adds; K scatter-adds backward):
 *   out[m,:] = sum_{k<K} table_cat[feat_row_offset[k] + idx[m,k], :]      idx int64 [M,K], table_cat [R,H], out [M,H]
 * table_cat = the K embedding tables concatenated row-wise; feat_row_offset_host = K+1 HOST ints (first row of each
 * table, last = R), K <= 16.  Sums run in feature order (bit-identical to the reference's out = 0 + e_0 + e_1 ...).
 * An index outside its table is clamped and reported by OR-ing 1 into *oob_flag [nullable].
 * bwd: dtable_cat[r,:] = sum of gout rows whose index hits r -- deterministic (per-CTA shared-memory slabs, one thread
 * per channel, then a fixed-order reduction); ws >= gsatb_embedding_sum_bwd_workspace(M, R, H) bytes.
 * ---------------------------------------------------------------------------------------------------------- */
int gsatb_embedding_sum_fwd(const int64_t* idx, const float* table_cat, const int32_t* feat_row_offset_host, float* out,
                            int32_t* oob_flag, int64_t M, int K, int H, gsatb_stream_t stream);
size_t gsatb_embedding_sum_bwd_workspace(int64_t M, int R, int H);
int gsatb_embedding_sum_bwd(const float* gout, const int64_t* idx, const int32_t* feat_row_offset_host,
                            float* dtable_cat, int64_t M, int K, int H, void* ws, size_t ws_bytes,
                            gsatb_stream_t stream);

/* Device-side batch collate.  Replaces the host-side torch_geometric DataLoader collate the reference runs for every
 * batch (src/utils/get_data_loaders.py:130-145; Batch.from_data_list semantics) for a dataset held
 * in HBM in packed form: graph g owns rows [ds_ptr[g], ds_ptr[g+1]) of each per-node / per-edge tensor and
 * ds_edge_index [2, E_ds] holds graph-local node ids.  ids [B] = the graphs of the batch in order; out_ptr [B+1] =
 * exclusive prefix sum of their row counts (out_ptr[B] == rows_out).  All pointers are device pointers.
 *   collate_rows:       out[out_ptr[b] + r, :] = src[ds_ptr[ids[b]] + r, :], rows of row_bytes (multiple of 4) bytes;
 *                       out_batch[out row] = b [nullable]; ds_ptr == NULL: one row per graph (per-graph labels);
 *                       row_bytes == 0: only out_batch is written.
 *   collate_edge_index: out[:, out_edge_ptr[b] + e] = ds_edge_index[:, ds_edge_ptr[ids[b]] + e] + out_node_ptr[b].
 * ---------------------------------------------------------------------------------------------------------- */
int gsatb_collate_rows(const void* src, int64_t row_bytes, const int64_t* ds_ptr, const int64_t* ids,
                       const int64_t* out_ptr, int64_t B, int64_t rows_out, void* out, int64_t* out_batch,
                       gsatb_stream_t stream);
int gsatb_collate_edge_index(const int64_t* ds_edge_index, int64_t E_ds, const int64_t* ds_edge_ptr, const int64_t* ids,
                             const int64_t* out_edge_ptr, const int64_t* out_node_ptr, int64_t B, int64_t E_out,
                             int64_t* out_edge_index, gsatb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GSAT_B200_H */
