/*
 * lgx.h -- C ABI of the B200-native LightGCN propagation + scoring engine (liblgx.so).
 *
 * This is the drop-in boundary for the hot path of the reference's PyTorch LightGCN
 * (PT/ = /root/reference/lightGCN/LightGCN-PyTorch-master/code/).  The reference has no
 * FFI for this path -- every device op is a PyTorch library call -- so each entry point
 * below cites the reference call site it replaces.  Host code (Python/PyTorch) binds these
 * with ctypes (factors_of_serendipity_recommendation_b200/_lgx.py); INTEGRATION.md shows
 * the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller (PyTorch) allocates all tensors and workspaces; the library owns only the
 *     opaque lgx_graph handle (freed by lgx_graph_destroy);
 *   - every call takes the CUDA stream to enqueue on and never synchronises implicitly,
 *     except lgx_graph_build* / lgx_graph_from_csr (one-time set-up, they sync the stream);
 *   - return 0 on success, non-zero on error; lgx_last_error() gives the thread-local message;
 *   - there is NO CPU fallback: without an sm_100 device every compute call returns LGX_ERR_DEVICE.
 */
#ifndef LGX_H_
#define LGX_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#pragma GCC visibility push(default)
#endif

#define LGX_OK 0
#define LGX_ERR_INVALID 1   /* bad argument */
#define LGX_ERR_CUDA 2      /* CUDA runtime / driver error */
#define LGX_ERR_DEVICE 3    /* no sm_100 device */
#define LGX_ERR_WORKSPACE 4 /* workspace too small */

/* scoring arithmetic modes (lgx_score_topk, lgx_pack_operand) */
#define LGX_SCORE_FP32 0    /* CUDA-core fp32 dot products (exact mode, same arithmetic as torch.matmul fp32) */
#define LGX_SCORE_BF16 1    /* tcgen05 bf16 x bf16 -> fp32 accumulate in TMEM */
#define LGX_SCORE_BF16X3 2  /* tcgen05, operands split hi+lo: hi*hi + hi*lo + lo*hi (~fp32 accuracy) */

typedef struct lgx_graph lgx_graph; /* opaque: CSR of D^-1/2 A D^-1/2 + degree-sorted schedule */
typedef void* lgx_stream;           /* cudaStream_t */

/* ------------------------------------------------------------------------------------------ misc */
const char* lgx_last_error(void);
int lgx_version(void);
/* 0 iff the current device is sm_100 (B200).  Fills sm_count / l2_bytes when non-NULL. */
int lgx_device_check(int* sm_count, int64_t* l2_bytes);

/* ----------------------------------------------------------------------------------------- graph
 * Replaces Loader.__init__ graph part (PT/dataloader.py:288-293: UserItemNet, users_D, items_D),
 * Loader.getSparseGraph (PT/dataloader.py:339-376: A=[[0,R],[R^T,0]], D^-1/2 A D^-1/2, tocsr) and
 * _convert_sp_mat_to_sp_tensor (:331-337).  Contract: canonical CSR (rows ascending, columns
 * ascending, duplicate (u,i) pairs merged with value = multiplicity), int64 indptr, int32 indices,
 * fp32 values = fl(fl(dinv[r]*mult)*dinv[c]) with dinv = correctly rounded fp32 deg^-1/2 (0 for
 * isolated nodes), plus a stable degree-descending row order used to schedule the SpMM.
 */
int lgx_graph_build(int32_t n_users, int32_t m_items, int64_t n_edges,
                    const int32_t* users, const int32_t* items, /* device, length n_edges */
                    int32_t chunk_nnz,                            /* long-row split size, 0 = default */
                    lgx_stream stream, lgx_graph** out);
/* Same from HOST arrays (copies inside; the train.txt parser's output). */
int lgx_graph_build_host(int32_t n_users, int32_t m_items, int64_t n_edges,
                         const int32_t* users_host, const int32_t* items_host,
                         int32_t chunk_nnz, lgx_stream stream, lgx_graph** out);
/* Adopt an existing CSR (device arrays are copied): the s_pre_adj_mat.npz cache read at
 * PT/dataloader.py:343, or one rank's row block for row-sharded propagation
 * (_split_A_hat, PT/dataloader.py:319-329).  n_rows may be < n_cols (a row shard). */
int lgx_graph_from_csr(int64_t n_rows, int64_t n_cols, int64_t nnz,
                       const int64_t* indptr, const int32_t* indices, const float* values,
                       int32_t n_users, int32_t m_items, int32_t chunk_nnz,
                       lgx_stream stream, lgx_graph** out);
/* info[0..7] = n_rows, n_cols, nnz, n_users, m_items, n_work_items, n_long_rows, max_row_nnz;
 * info[8] = partial scratch floats per unit d (n_partials); info[9] = chunk_nnz. */
int lgx_graph_info(const lgx_graph* g, int64_t* info_host /* [10] */);
/* Copy out the canonical CSR / degree tables (any pointer may be NULL). */
int lgx_graph_export(const lgx_graph* g, int64_t* indptr, int32_t* indices, float* values,
                     int32_t* degree, float* dinv, int32_t* row_order, lgx_stream stream);
/* Borrow the handle's own device arrays (valid until destroy) so torch can alias them as a
 * sparse_csr tensor for getSparseGraph() without a copy. */
int lgx_graph_pointers(const lgx_graph* g, const int64_t** indptr, const int32_t** indices,
                       const float** values);
/* LGX_GRAPH_NORMALIZED: every stored value equals dinv[row] * dinv[col] with dinv = (row length)^-1/2 -- true for
 * lgx_graph_build graphs without duplicate pairs.  lgx_graph_from_csr cannot know it: a caller that adopts a row block
 * of such a graph (row-sharded propagation) sets the flag, which lets the SpMM classify hot columns from the value
 * alone (L2 residency hints on graphs whose embedding table is far larger than L2). */
#define LGX_GRAPH_NORMALIZED 1
int lgx_graph_get_flags(const lgx_graph* g);
int lgx_graph_set_flags(lgx_graph* g, int32_t flags);
int lgx_graph_destroy(lgx_graph* g);

/* ----------------------------------------------------------------------------------- propagation
 * One sparse layer with the fused epilogue (replaces torch.sparse.mm at PT/model.py:171 and the
 * stack/mean at :173-175):
 *     acc   = A_hat[rows] * X                  (X: [n_cols, d] fp32 row-major)
 *     Y     = acc                              (if Y != NULL;    [n_rows, d])
 *     S_out = (S_in + acc) / div               (if S_out != NULL; S_in may alias S_out)
 * workspace: lgx_spmm_workspace_bytes(g, d) bytes (partials of split long rows), may be NULL if 0.
 */
size_t lgx_spmm_workspace_bytes(const lgx_graph* g, int32_t d);
int lgx_spmm(const lgx_graph* g, const float* X, const float* S_in, float* Y, float* S_out,
             float div, int32_t d, void* workspace, lgx_stream stream);

/* Fused SpMM + all-gather for row-sharded propagation (the reference's fold loop + torch.cat,
 * PT/model.py:164-169, across GPUs): the epilogue stores this rank's output rows straight into EVERY
 * rank's copy of the gathered layer over NVLink peer memory, at rows [row_offset, row_offset + n_rows).
 *   peers_host : HOST array of n_peers device pointers (own buffer included), each [n_total_rows, d] fp32,
 *                obtained with lgx_peer_alloc / lgx_peer_open (CUDA IPC, one process per GPU, one node);
 *   store_mean : 0 -> peers receive acc (next layer's input); 1 -> peers receive (S_in + acc) / div.
 * The caller separates layers with a cross-rank barrier (all writers done before anyone reads). */
int lgx_spmm_peers(const lgx_graph* g, const float* X, const float* S_in, float* const* peers_host,
                   int32_t n_peers, int64_t row_offset, int32_t store_mean, float* S_out, float div,
                   int32_t d, void* workspace, lgx_stream stream);
/* Peer-visible device allocations: alloc returns the pointer and a 64-byte IPC handle to send to the
 * other ranks, which map it with lgx_peer_open (and unmap with lgx_peer_close). */
int lgx_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int lgx_peer_open(const unsigned char* handle64, void** ptr);
int lgx_peer_close(void* ptr);
/* Copy-engine variant of the exchange: copy bytes [offset, offset + bytes) of this rank's buffer
 * (peers_host[self]) into the same range of every other rank's buffer with one asynchronous P2P copy
 * per peer on `stream` (run it on a side stream to overlap with the next chunk's SpMM). */
int lgx_peer_copy(void* const* peers_host, int32_t n_peers, int32_t self, size_t offset_bytes, size_t bytes,
                  lgx_stream stream);
int lgx_peer_free(void* ptr);
/* Cross-rank barrier between two layers of the fused exchange, on the device: every rank owns a peer-visible array of
 * n_peers uint32 flags (lgx_peer_alloc, zero-initialised); the kernel writes `epoch` (> 0, growing by one per call on
 * all ranks) into slot [self] of every rank's array and waits until its own array has reached `epoch`.  Stream-
 * ordered after the layer's SpMM; a rank that never arrives makes the others trap after ~10 s instead of hanging. */
int lgx_peer_barrier(void* const* flag_peers_host, int32_t n_peers, int32_t self, uint32_t epoch,
                     lgx_stream stream);

/* LightGCN.computer() (PT/model.py:145-177): out = mean(E0, A E0, ..., A^L E0), E0 = cat(users, items).
 * workspace: lgx_propagate_workspace_bytes(g, d, L) (two [n,d] ping-pong layers + spmm workspace).
 * layers_out (optional): [L, n, d] receives every layer's embeddings. Square graphs only. */
size_t lgx_propagate_workspace_bytes(const lgx_graph* g, int32_t d, int32_t n_layers);
int lgx_propagate_fwd(const lgx_graph* g, const float* E0, float* out_mean, float* layers_out,
                      int32_t n_layers, int32_t d, void* workspace, lgx_stream stream);
/* Backward of computer() (autograd of PT/model.py:163-175 triggered at PT/utils.py:49).
 * g_scaled = dL/d(out_mean) / (L+1).  dE0 = g + A(g + A(g + ...)) (Horner; A_hat is symmetric). */
int lgx_propagate_bwd(const lgx_graph* g, const float* g_scaled, float* dE0,
                      int32_t n_layers, int32_t d, void* workspace, lgx_stream stream);

/* Edge dropout (LightGCN.__dropout_x / __dropout, PT/model.py:125-143; --dropout 1, training only).
 * Every stored entry k of A_hat is kept with probability keep_prob and scaled by 1/keep_prob; the keep
 * decision is a counter-based hash of (seed, position of the entry), evaluated inside the SpMM -- no
 * COO rebuild and no CPU torch.rand.  One seed per computer() call: all L layers and the backward of
 * that call see the same dropped graph.  The dropped graph is not symmetric, so the backward reads the
 * keep decision of the mirrored entry (positions built once by lgx_graph_enable_dropout).
 * lgx_dropout_mask exports the mask (uint8[nnz]; transpose = 1: mask of the mirrored entries). */
int lgx_graph_enable_dropout(lgx_graph* g, lgx_stream stream);
int lgx_dropout_mask(const lgx_graph* g, float keep_prob, uint64_t seed, int32_t transpose,
                     uint8_t* mask, lgx_stream stream);
int lgx_propagate_fwd_dropout(const lgx_graph* g, const float* E0, float* out_mean, int32_t n_layers,
                              int32_t d, float keep_prob, uint64_t seed, void* workspace,
                              lgx_stream stream);
int lgx_propagate_bwd_dropout(const lgx_graph* g, const float* g_scaled, float* dE0, int32_t n_layers,
                              int32_t d, float keep_prob, uint64_t seed, void* workspace,
                              lgx_stream stream);

/* --------------------------------------------------------------------------------------- scoring
 * getUsersRating (PT/model.py:179-184): out[b, j] = f(<U[users[b]], I[j]>), f = sigmoid if
 * apply_sigmoid; fp32 CUDA-core arithmetic like the reference's SGEMM.  users may be NULL (rows 0..B-1). */
int lgx_score_dense(const float* U, const int64_t* users, int32_t B, const float* I, int32_t M,
                    int32_t d, float* out, int32_t apply_sigmoid, lgx_stream stream);

/* Candidate bucketing of the serendipity pipeline (/root/reference/recommend.py:375-380, the same min / max scan over
 * item-item products at /root/reference/utils.py:496-519): scores are the fp32 dot products rounded to fp16 like the
 * reference's .astype(np.float16).
 *   lgx_score_minmax : out2 (device float[2]) = {min, max} of fp16(<U[b], I[j]>) over all b < B, j < M; the [B, M] matrix
 *                      is never materialised.  workspace8: 8 bytes of device scratch.
 *   lgx_score_bucket : labels[b, j] = int8(floor(fp16(fp16(score - min_dis) / inter))) for a user batch (users may be
 *                      NULL = rows 0..B-1 of U); min_dis and inter are rounded to fp16 like numpy's float16 ufuncs do. */
int lgx_score_minmax(const float* U, int32_t B, const float* I, int32_t M, int32_t d, float* out2,
                     void* workspace8, lgx_stream stream);
int lgx_score_bucket(const float* U, const int64_t* users, int32_t B, const float* I, int32_t M, int32_t d,
                     float min_dis, float inter, int8_t* labels, lgx_stream stream);

/* Pack fp32 rows into the bf16 operand layout the tcgen05 kernel reads: [rows, K] bf16 with
 * K = d (LGX_SCORE_BF16) or 3d (LGX_SCORE_BF16X3; is_items picks [hi|lo|hi] vs users' [hi|hi|lo]).
 * row_ids (optional int64[rows]) gathers rows of src. */
size_t lgx_pack_bytes(int32_t rows, int32_t d, int32_t mode);
int lgx_pack_operand(const float* src, const int64_t* row_ids, int32_t rows, int32_t d, int32_t mode,
                     int32_t is_items, void* dst, lgx_stream stream);

/* Fused score + train-mask + top-K (replaces getUsersRating + the mask at PT/Procedure.py:129-134 +
 * torch.topk at :135; the [B, M] score matrix never reaches HBM).
 *   U_op / I_op : fp32 [B,d] / [M,d] for LGX_SCORE_FP32 (U_op already gathered to the batch),
 *                 packed bf16 operands (lgx_pack_operand) for the tcgen05 modes;
 *   users       : int64[B] global user ids of the batch rows (mask lookup), or NULL = the identity batch (row u is
 *                 user u, e.g. "all users in order").  For the identity batch the tcgen05 modes keep the graph's
 *                 train mask in (user tile, item tile) buckets with the graph handle -- built by the first call,
 *                 reused by the following ones, like the reference's dataset builds allPos once
 *                 (PT/dataloader.py); LGX_SCORE_MASK_CACHE=0 buckets on every call as for explicit batches;
 *   g           : graph whose user rows hold the train items to exclude, or NULL for no mask;
 *   item_offset : global id of local item 0 (item-sharded catalogue); indices returned are global;
 *   out_idx/out_val : int64 / fp32 [B, k], sorted by score descending, ties by ascending item id.
 *                 Scores are RAW dot products (sigmoid is monotone; apply it to out_val if needed).
 *                 If fewer than k unmasked items exist the tail is filled with masked train items
 *                 carrying value -1024 like the reference (PT/Procedure.py:134).
 * The workspace holds the per-split partial lists and, when a user tile's catalogue is split over several
 * CTAs, one 32-bit bound per batch row through which those CTAs share the row's running threshold.
 */
size_t lgx_score_topk_workspace_bytes(int32_t B, int32_t M, int32_t d, int32_t k, int32_t mode);
/* Host-only: how lgx_score_topk decomposes a call on a device with `sms` SMs (0 = the current device, 148 when
 * there is none).  plan4 = {user tiles, item tiles, splits per user tile, item tiles per split}.  The tcgen05
 * modes pick the split count that minimises waves * (item tiles / splits + per-unit overhead); needs no GPU. */
int lgx_score_plan(int32_t B, int32_t M, int32_t d, int32_t k, int32_t mode, int32_t sms, int32_t* plan4);
int lgx_score_topk(const lgx_graph* g, const void* U_op, const int64_t* users, int32_t B,
                   const void* I_op, int32_t M, int32_t d, int32_t k, int32_t mode,
                   int64_t item_offset, int64_t* out_idx, float* out_val,
                   void* workspace, size_t workspace_bytes, lgx_stream stream);
/* Merge P per-shard candidate lists [P, B, k] (after an all-gather) into the global top-k. */
int lgx_topk_merge(const int64_t* cand_idx, const float* cand_val, int32_t P, int32_t B, int32_t k,
                   int64_t* out_idx, float* out_val, lgx_stream stream);

/* ------------------------------------------------------------------------------------------ BPR
 * getEmbedding + bpr_loss (PT/model.py:186-209): light = computer() output [N,d], E0 = raw tables.
 *   out2[0] = mean softplus(<u,n> - <u,p>), out2[1] = 0.5*(|u0|^2+|p0|^2+|n0|^2)/B;
 *   coef     : float[3*B] scratch; coef[0:B] = sigmoid(<u,n> - <u,p>) / B (d loss / d (neg-pos), kept
 *              for the backward), coef[B:3B] = per-sample loss / reg terms (reduced in a fixed order).
 */
int lgx_bpr_forward(const float* light, const float* E0, const int64_t* users, const int64_t* pos,
                    const int64_t* neg, int32_t B, int32_t n_users, int32_t d,
                    float* out2, float* coef, lgx_stream stream);
/* Scatter-add d loss/d light, scaled by grad_scale * (*grad_scale_dev), into G [N,d] (caller zeroes
 * G).  grad_scale_dev (optional device scalar) carries the upstream autograd gradient without a
 * host sync. */
int lgx_bpr_backward_light(const float* light, const int64_t* users, const int64_t* pos,
                           const int64_t* neg, const float* coef, int32_t B, int32_t n_users,
                           int32_t d, float grad_scale, const float* grad_scale_dev, float* G,
                           lgx_stream stream);
/* Scatter-add d reg/d E0 = grad_scale * (*grad_scale_dev) * E0[row] / B into dE0. */
int lgx_bpr_backward_reg(const float* E0, const int64_t* users, const int64_t* pos,
                         const int64_t* neg, int32_t B, int32_t n_users, int32_t d,
                         float grad_scale, const float* grad_scale_dev, float* dE0,
                         lgx_stream stream);
/* torch.optim.Adam step (PT/utils.py:41,50), one fused pass over a contiguous table. */
int lgx_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, int32_t step, lgx_stream stream);

/* Same update with the step counter on the device (state4: int32[4] = {step, step_size, bc2_sqrt, pad},
 * zero-initialised by the caller): nothing in the call depends on a host value that changes per step,
 * so the whole BPR step can be captured in a CUDA graph and replayed. */
int lgx_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                      float lr, float beta1, float beta2, float eps, int32_t* state4, lgx_stream stream);

/* -------------------------------------------------------------------------------------- sampler
 * BPR triples (PT/utils.py:55-99 / PT/sources/sampling.cpp:27-56), counter-based RNG on device.
 *   per_user == 0: reference Python semantics (user ~ U[0,n_users) with replacement; users without
 *                  positives are re-drawn), n_samples triples;
 *   per_user  > 0: sampling.cpp semantics, every user exactly per_user triples (n_samples ignored).
 * out: int64 [S, 3] (user, pos, neg); neg is rejected against the user's train row by binary search.
 */
int lgx_sample_bpr(const lgx_graph* g, int64_t n_samples, int32_t per_user, uint64_t seed,
                   int64_t* out, lgx_stream stream);

/* -------------------------------------------------------------------------------------- metrics
 * utils.getLabel + RecallPrecision_ATk + NDCGatK_r (PT/utils.py:218-285) for one k, summed over the
 * batch: sums3[0] += sum recall, sums3[1] += sum hits (precision * k), sums3[2] += sum ndcg  (fp64).
 * Ground truth as CSR over the batch rows: gt_ptr int64[B+1], gt_items int64.
 */
int lgx_rank_metrics(const int64_t* topk_idx, int32_t B, int32_t k_stride, int32_t k,
                     const int64_t* gt_ptr, const int64_t* gt_items, double* sums3, lgx_stream stream);

#ifdef __cplusplus
#pragma GCC visibility pop
}
#endif
#endif /* LGX_H_ */
