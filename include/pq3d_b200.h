/* pq3d_b200 — C ABI of the B200-native promptable-query-decoder kernels.
 *
 * The reference (PQ3D) has no FFI on this path: the decoder is PyTorch eager code
 * (modules/grounding/query_encoder.py, modules/layers/transformers.py, modules/heads/mask_head.py)
 * that bottoms out in torch.nn.functional.  Each entry point below names the reference call site it
 * replaces.  A host in any language binds these with plain pointers and sizes (ctypes in this repo:
 * pq3d_b200/_lib.py; INTEGRATION.md shows the reference-side stub).
 *
 * Conventions
 *   - every pointer except `stream` and the pointer TABLES (`K`, `Vt`, `mask_bits` in
 *     pq3d_attention_fwd are host arrays of device pointers; `mem_masks_dev` in
 *     pq3d_mask_head_finalize is a DEVICE array of device pointers) is a device pointer owned by the
 *     caller; nothing is allocated or freed by the library and no global state is kept
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it
 *   - bf16 tensors are `void*`; "ld*" are leading dimensions in ELEMENTS
 *   - bool masks are uint8 with 1 = ignore (PyTorch convention, model/query3d_unified.py:125)
 *   - return 0 on success, <0 on error (-1 invalid argument, -2 CUDA error, -3 unsupported);
 *     pq3d_last_error() returns a thread-local message.  The library never calls exit().
 *   - requires an sm_100a device (tcgen05 / TMEM / TMA); there is no fallback path.
 */
#ifndef PQ3D_B200_H_
#define PQ3D_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* pq3d_last_error(void);
int pq3d_abi_version(void);
/* Launch priority of the calling thread's subsequent kernel launches (cudaLaunchAttributePriority; 0 = default = least
 * urgent, negative = more urgent, clamped to the device's range).  Recorded per kernel node when a CUDA graph captures the
 * launches: with several batches in flight the latency-bound query-side kernels are marked urgent so that the long
 * K / V^T projections of other batches do not hold them back.  (Scheduling only; no reference counterpart.) */
int pq3d_set_launch_priority(int priority);
/* Debug only: per-CTA timeline of pq3d_linear_bf16 (8 x uint64 per CTA) into a device buffer; NULL disables. */
int pq3d_debug_set_timeline(void* buf);
int pq3d_debug_set_attention_timeline(void* buf);
/* Testing hook: 1 = long memories always take the running-max (two-pass) attention schedule. */
int pq3d_debug_force_two_pass(int on);

/* C[g] = epilogue(A[g] · W[g]ᵀ), bf16 operands, fp32 accumulation on tcgen05 tensor cores.
 *   A: [a_rows_total, lda] bf16, group g starts at row g*a_group_rows, uses M rows, K columns
 *   W: [w_rows_total, ldw] bf16, group g starts at row g*w_group_rows, uses N rows (nn.Linear layout)
 *   C: out_fp32 ? float : bf16, element (g, m, n) at C[g*c_group_stride + m*ldc + n]
 *   epilogue: x = acc + bias[g*bias_group_stride + (bias_along_m ? m : n)];
 *             if (n < alpha_ncols) x *= alpha;  if (relu) x = max(x, 0);
 *             if (row_zero && row_zero[g*row_zero_group_stride + m]) x = 0
 *   block_n: 0 = auto, or 64 / 128 / 256 (tile width).  K must be a multiple of 64.
 * Replaces every nn.Linear on the path: MHA in/out projections (torch/nn/functional.py:5867-5873,
 * :6653), w_qs/w_ks/w_vs/fc (modules/layers/transformers.py:180-185,190-192,238), FFN
 * (modules/grounding/query_encoder.py:384), mask-head q_proj/k_proj and their einsum
 * (modules/heads/mask_head.py:52-56), cls MLP (modules/utils.py:18-25). */
int pq3d_linear_bf16(const void* A, int64_t lda, int64_t a_rows_total, int64_t a_group_rows,
                     const void* W, int64_t ldw, int64_t w_rows_total, int64_t w_group_rows,
                     void* C, int64_t ldc, int64_t c_group_stride, int out_fp32,
                     const float* bias, int64_t bias_group_stride, int bias_along_m,
                     const uint8_t* row_zero, int64_t row_zero_group_stride,
                     int M, int N, int K, int groups, float alpha, int alpha_ncols, int relu,
                     int block_n, void* stream);

/* pq3d_linear_bf16 with two extras.  max_ctas > 0: the persistent grid uses at most that many CTAs (SMs), so a long
 * projection can run NEXT TO a chain of small latency-bound kernels on another stream instead of holding every SM.
 * a_row_offsets (HOST array [groups], groups <= 8, or NULL): group g's A rows start at a_row_offsets[g] instead of
 * g*a_group_rows — groups that share or permute their A operand (self-attention q / k / v reading x+pos, x+pos, x:
 * modules/layers/transformers.py:189-192, torch/nn/functional.py:5867-5873).
 * flags bit 0: W holds weights no kernel of the running dependency chain writes; its first tiles are fetched before the
 * programmatic-dependent-launch wait on the previous kernel (only A depends on that kernel).  flags bit 1: never use CTA
 * pairs (cta_group::2) — set by callers that keep several graphs in flight on different streams (see csrc/gemm.cu). */
int pq3d_linear_bf16_ex(const void* A, int64_t lda, int64_t a_rows_total, int64_t a_group_rows,
                        const void* W, int64_t ldw, int64_t w_rows_total, int64_t w_group_rows,
                        void* C, int64_t ldc, int64_t c_group_stride, int out_fp32,
                        const float* bias, int64_t bias_group_stride, int bias_along_m,
                        const uint8_t* row_zero, int64_t row_zero_group_stride,
                        int M, int N, int K, int groups, float alpha, int alpha_ncols, int relu,
                        int block_n, int max_ctas, const int32_t* a_row_offsets, int flags, void* stream);

/* Strided batched GEMM on the same kernel: C[g1,g2] = alpha * A[g1,g2] · W[g1,g2]ᵀ for G1 x G2 problems whose operands
 * are strided views, e.g. the per-(scene, head) slices of [tokens, heads*64] tensors.  Element (g1, g2, row, k) of A at
 * A[g1*a_g1_stride + g2*a_g2_stride + row*a_row_stride + k] (strides in elements, all 16-byte granular, non-zero);
 * same for W (rows = N) and C.  K multiple of 64.  Used by the attention backward products (dP = dO·Vᵀ, dV = Pᵀ·dO,
 * dQ = dS·K, dK = dSᵀ·Q — torch autograd of torch/nn/functional.py:6630-6647 in the reference's training step). */
int pq3d_bgemm_bf16(const void* A, int64_t a_row_stride, int64_t a_g2_stride, int64_t a_g1_stride,
                    const void* W, int64_t w_row_stride, int64_t w_g2_stride, int64_t w_g1_stride,
                    void* C, int64_t c_row_stride, int64_t c_g2_stride, int64_t c_g1_stride, int out_fp32,
                    int M, int N, int K, int G2, int G1, float alpha, int block_n, void* stream);

/* Masked softmax attention for n_mem (<= 4) memories in one launch, head_dim = 64.
 *   Q : bf16 [B*Nq, ldq]; memory i / head h at columns i*q_mem_stride + h*64; already scaled by
 *       log2(e)/sqrt(64) — scores are handled in the log2 domain (a probability is one ex2)
 *   K[i] : bf16 [B*S_pitch[i], ldk[i]]; head h at columns k_col0[i] + h*64
 *   Vt[i]: bf16 [vt_rows[i], ldvt[i]] — V TRANSPOSED: row vt_row0[i] + h*64 + d, column b*Vt_pitch[i] + s
 *          (Vt_pitch multiple of 8: TMA strides are 16-byte granular; pad columns must hold finite values)
 *   mask_bits[i]: packed by pq3d_pack_mask (1 = ignore), word address
 *                 b*mask_b_stride + h*mask_h_stride + n*mask_q_stride + s/32; NULL = nothing masked
 *   kv_tiles[i]:  optional device int32 [B] from pq3d_pack_mask(active_tiles): key tiles past it are skipped
 *   O : bf16, element (i, b, n, h*64+d) at O[i*o_mem_stride + (b*Nq+n)*ldo + h*64 + d]
 *   zero_attn: nn.MultiheadAttention(add_zero_attn=True) — one extra never-masked key with score 0
 *              and value 0 (torch/nn/functional.py:6585-6602), handled analytically
 *   score_bias (or NULL): fp32 [B,H,Nq,bias_ld], log2 domain, added to the scores before masking, rows padded to a
 *              multiple of 128 keys — the spatial term of MultiHeadAttentionSpatial 'mul'
 *              (modules/layers/transformers.py:231-233), produced by pq3d_spatial_bias.
 *   stat_m / stat_l (or NULL): fp32 [n_mem][B][H][Nq] (memory stride stat_mem_stride) receive, per row, the softmax
 *              reference m (log2 domain) and the denominator l (zero-attn term included): P = ex2(s - m) / l —
 *              what the backward needs to recompute the probabilities.
 * Replaces torch/nn/functional.py:6630-6647 as called from CrossAttentionLayer.forward_post
 * (modules/grounding/query_encoder.py:297-303) and modules/layers/transformers.py:193-237. */
int pq3d_attention_fwd(int n_mem, const void* Q, int64_t ldq, int64_t q_mem_stride,
                       const void* const* K, const int64_t* ldk, const int64_t* k_col0,
                       const void* const* Vt, const int64_t* ldvt, const int64_t* vt_row0, const int64_t* vt_rows,
                       const int32_t* S, const int32_t* S_pitch, const int32_t* Vt_pitch,
                       const uint32_t* const* mask_bits, const int64_t* mask_b_stride,
                       const int64_t* mask_h_stride, const int64_t* mask_q_stride, const int32_t* const* kv_tiles,
                       void* O, int64_t ldo, int64_t o_mem_stride, int B, int H, int Nq, int zero_attn,
                       const float* score_bias, int64_t bias_ld, float* stat_m, float* stat_l,
                       int64_t stat_mem_stride, void* stream);

/* Training-mode forward: pq3d_attention_fwd plus dropout on the attention probabilities (nn.MultiheadAttention(dropout=p):
 * applied after the softmax normalisation).  Probability (b, h, n, key) of memory i is kept iff the counter RNG says so
 * for element ((b*H + h)*Nq + n)*ceil128(S[i]) + key of stream sites[i] (host array) in the step seeded by *seed_dev. */
int pq3d_attention_fwd_train(int n_mem, const void* Q, int64_t ldq, int64_t q_mem_stride,
                       const void* const* K, const int64_t* ldk, const int64_t* k_col0,
                       const void* const* Vt, const int64_t* ldvt, const int64_t* vt_row0, const int64_t* vt_rows,
                       const int32_t* S, const int32_t* S_pitch, const int32_t* Vt_pitch,
                       const uint32_t* const* mask_bits, const int64_t* mask_b_stride,
                       const int64_t* mask_h_stride, const int64_t* mask_q_stride, const int32_t* const* kv_tiles,
                       void* O, int64_t ldo, int64_t o_mem_stride, int B, int H, int Nq, int zero_attn,
                       const float* score_bias, int64_t bias_ld, float* stat_m, float* stat_l,
                       int64_t stat_mem_stride, float drop_p,
                       const uint32_t* seed_dev, const uint32_t* sites, void* stream);

/* Score bias of MultiHeadAttentionSpatial 'mul' for L layers at once:
 * out[l,b,h,n,m] = log2(max(relu(pairwise_locs[b,n,m,:] · loc_w[l,h,:] + loc_b[l,h]), 1e-6)), rows padded to ld
 * (pad columns are left untouched).  Replaces modules/layers/transformers.py:196-199,231-232. */
int pq3d_spatial_bias(const float* pairwise_locs, const float* loc_w, const float* loc_b, float* out, int L, int B,
                      int H, int N, int64_t ld, void* stream);

/* xv = bf16(feat), xk = bf16(feat + pos) with rows padded to S_pitch (zero-filled); pos may be NULL
 * (then xk = xv values), either output may be NULL.  feat/pos: fp32 [B,S,D].
 * Replaces CrossAttentionLayer.with_pos_embed + the autocast input casts
 * (modules/grounding/query_encoder.py:285-286,298-300). */
int pq3d_ingest_memory(const float* feat, const float* pos, void* xk, void* xv, int B, int S, int S_pitch, int D,
                       void* stream);

/* The same for n_mem (1..4) memories of identical shape that share ONE positional table (the scene memories all add
 * fts_pos, model/query3d_unified.py:139-155): pos is read once.  feats: HOST array of device pointers; memory m's
 * operands land at xk + m*mem_stride, xv + m*mem_stride (elements). */
int pq3d_ingest_memories(int n_mem, const float* const* feats, const float* pos, void* xk, void* xv, int64_t mem_stride,
                         int B, int S, int S_pitch, int D, void* stream);

/* out = (1/G) * sum_g LayerNorm_g(residual + y[g]); y: fp32 [G][R,D] (group stride in elements) or NULL,
 * residual fp32 [R,D] or NULL, gamma/beta fp32 [G,D].  Optional outputs (NULL to skip): out_f32,
 * out_bf16 = bf16(out), out_pos_bf16 = bf16(out + pos).  D multiple of 128, <= 1024.
 * Replaces `tgt = norm(tgt + dropout(tgt2))` (modules/grounding/query_encoder.py:304-305,386-387,
 * 449-450) and parallel_ca's mean over memories (:153). */
int pq3d_add_layernorm(const float* y, int64_t y_group_stride, const float* residual, const float* gamma,
                       const float* beta, int G, float eps, int R, int D, const float* pos, float* out_f32,
                       void* out_bf16, void* out_pos_bf16, void* stream);

/* bits[row, w] packs mask[row, 32w .. 32w+31] (1 = ignore), W = 4*ceil(S/128) words per row, bits past S set.
 * unmask_full_rows: rows that are entirely masked become entirely visible
 * (`attn_mask[attn_mask.all(-1)] = False`, modules/grounding/query_encoder.py:83); mask_fixed (optional,
 * [rows,S]) receives the bool mask after that fix-up.  active_tiles (optional, int32 [rows/rows_per_batch]) receives,
 * per batch entry, the number of leading 128-key tiles that contain a visible key for at least one of its rows —
 * pq3d_attention_fwd skips the rest (trailing padding of ragged scenes). */
int pq3d_pack_mask(const uint8_t* mask, uint32_t* bits, int64_t rows, int S, int unmask_full_rows,
                   uint8_t* mask_fixed, int32_t* active_tiles, int64_t rows_per_batch, void* stream);

/* mask_logits[b,s,n] = raw[b,s,n] / (#memories valid at (b,s) + 1e-8), -1e6 where seg_masks[b,s];
 * attn_mask[b,n,s] = sigmoid(mask_logits[b,s,n]) < 0.5.  mem_masks_dev: device array of n_mem device
 * pointers to uint8 [B,S] (1 = ignore).  Replaces modules/heads/mask_head.py:36-43. */
int pq3d_mask_head_finalize(const float* raw, const uint8_t* const* mem_masks_dev, int n_mem,
                            const uint8_t* seg_masks, float* mask_logits, uint8_t* attn_mask, int B, int S, int N,
                            void* stream);

/* out = (1 - sigmoid(g)) * query + sigmoid(g) * update  (structure 'gate', query_encoder.py:166-170). */
int pq3d_gate_mix(const float* gate_logits, const float* query, const float* update, float* out, int64_t n,
                  void* stream);

/* out = bf16(x + add) (add may be NULL); n multiple of 4. */
int pq3d_cast_bf16(const float* x, const float* add, void* out, int64_t n, void* stream);

/* Fourier positional features as the bf16 operand of CoordinateEncoder.feat_proj:
 * xyz' = (xyz - min)/(max - min); out[b,l,:] = [sin(2*pi*xyz'·gauss_B), cos(...)], gauss_B fp32 [3, d_pos/2].
 * xyz: fp32, point (b,l) at xyz + (b*L+l)*xyz_stride.  Replaces model/query3d_unified.py:22-24 ->
 * modules/third_party/mask3d/position_embedding.py:38-43,127-156. */
int pq3d_fourier_pos(const float* xyz, int xyz_stride, const float* coord_min, const float* coord_max,
                     const float* gauss_B, void* out_bf16, int B, int L, int d_pos, void* stream);

/* out[b,i,j,:] = [d/max_b d, dz/d, d2/d, dy/d2, dx/d2] for query centres (fp32, row stride c_stride),
 * d = sqrt(|ci-cj|^2 + eps), d2 over xy.  Replaces calc_pairwise_locs(..., 'center', spatial_dist_norm=True,
 * spatial_dim=5), modules/utils.py:38-68. */
int pq3d_pairwise_locs(const float* centers, int c_stride, float* out, int B, int N, float eps, void* stream);

/* ---- §8f-2: voxel -> segment pooling (modules/vision/pcd_mask3d_encoder.py:144-154, torch_scatter.scatter_mean) ---- */

/* Bytes of caller-owned scratch pq3d_segment_csr needs; voxel_offsets_host is a HOST array of B+1 prefix sums (scene b
 * owns voxels [off[b], off[b+1])).  Returns -1 on bad arguments. */
int64_t pq3d_segment_csr_workspace_bytes(const int64_t* voxel_offsets_host, int B, int max_seg);

/* Stable counting sort of the voxels by (scene, segment): perm (int32 [Nv_total]) lists the voxel ids segment by
 * segment, ascending inside a segment; offsets (int32 [B*max_seg + 1]) delimit the segments.  p2s: int64 [Nv_total]
 * per-scene segment ids (the reference's `point2segment`, scenes concatenated); ids outside [0, max_seg) are skipped.
 * Integer work only, deterministic; shared by the five feature scales of a batch. */
int pq3d_segment_csr(const int64_t* p2s, const int64_t* voxel_offsets_host, int B, int max_seg, int32_t* perm,
                     int32_t* offsets, void* workspace, int64_t workspace_bytes, void* stream);

/* out[g, :] = sum_{v in segment g, ascending v} feat[v, :] / max(count_g, 1)   (fp32 adds in voxel order: bit-exact
 * against torch's CPU scatter_add; empty segments give 0 like scatter_mean).  feat fp32 [Nv_total, C] (ld); G = B*max_seg.
 * out32 (optional) fp32 [G, C] (ld32); out16 (optional) bf16 [G, K16] (ld16), columns [C, K16) zero — the operand of the
 * per-scale Linear(C -> hidden) GEMM that follows (pcd_mask3d_encoder.py:125-130,151). */
int pq3d_segment_mean(const float* feat, int64_t ld, const int32_t* perm, const int32_t* offsets, int G, int C,
                      float* out32, int64_t ld32, void* out16, int64_t ld16, int K16, void* stream);

/* ---- §8f-3: matcher cost matrices and matched mask losses (modules/third_party/mask3d/matcher.py:12-64,103-181,
 *      criterion.py:26-75,163-196) ---- */

/* cost[b, n, m] = w_mask * batch_sigmoid_ce(n, m) + w_class * (-softmax(pred_logits[b, n])[label_m], -1 for
 * ignore_label) + w_dice * batch_dice(n, m) for m < tgt_count[b], 0 beyond; fp32 throughout like the reference
 * (autocast disabled, matcher.py:160).  pred_masks fp32 [B, S, N] (a `predictions_mask` entry), pred_logits fp32
 * [B, N, C], tgt_masks uint8 [B, Mmax, S] (1 = point in instance), tgt_labels int64 [B, Mmax], tgt_count int32 [B].
 * All S points are used (num_points = -1, the shipped configuration). */
int pq3d_match_cost(const float* pred_masks, const float* pred_logits, const uint8_t* tgt_masks, const int64_t* tgt_labels,
                    const int32_t* tgt_count, float* cost, int B, int N, int S, int C, int Mmax, float w_class,
                    float w_mask, float w_dice, int64_t ignore_label, void* stream);

/* Per matched pair k = pairs[k] = (scene, query, target) (int32 [n_pairs, 3], device): ce[k] = mean_c BCE(x, t),
 * dice[k] = 1 - (2 sum sig*t + 1) / (sum sig + sum t + 1); sums fp32 [n_pairs, 2] is saved for the backward. */
int pq3d_matched_mask_loss_fwd(const float* pred_masks, const uint8_t* tgt_masks, const int32_t* pairs, int n_pairs, int B,
                               int N, int S, int Mmax, float* ce, float* dice, float* sums, void* stream);
/* d_pred (fp32 [B, S, N], zero-filled by the caller) receives g_ce[k] * d ce[k] + g_dice[k] * d dice[k]. */
int pq3d_matched_mask_loss_bwd(const float* pred_masks, const uint8_t* tgt_masks, const int32_t* pairs, int n_pairs, int B,
                               int N, int S, int Mmax, const float* g_ce, const float* g_dice, const float* sums,
                               float* d_pred, void* stream);

/* ---- §8f-4: inference post-processing of the instance predictions, per scene
 *      (evaluator/instseg_eval.py:85-150 eval_instance_step, :272-305 get_full_res_mask / get_mask_and_scores) ---- */

/* probs[q, c] = softmax(logits[q, :])[c] for c < C1 - 1 (the last, no-object class is dropped, :106). */
int pq3d_class_probs(const float* logits, float* probs, int Q, int C1, void* stream);
/* Exact top-K of n floats, sorted descending, ties: lower index first (torch.topk(sorted=True), :287-289).
 * K <= min(n, 1024), n <= 2^20. */
int pq3d_topk(const float* vals, int n, int K, float* out_val, int32_t* out_idx, void* stream);
/* out[b] = #{i : idx[i] == b} (int32, zero-filled here): voxels per segment, points per full-resolution segment. */
int pq3d_bincount(const int64_t* idx, int64_t n, int32_t* out, int bins, void* stream);
/* score[k] = cls_score[k] * sum_v sig(m) [m > 0] / (sum_v [m > 0] + 1e-6) for query q_k = sel_flat[k] / C, evaluated on
 * SEGMENT logits pred_masks [S, Q] weighted by the voxel count of each segment (masks[voxel2segment] is never built). */
int pq3d_instseg_scores(const float* pred_masks, const int32_t* seg_count, const int32_t* sel_flat, int C, int S, int Q,
                        int K, const float* cls_score, float* score, void* stream);
/* q_of[k] = flat[order[k]] / C, cls_of[k] = flat[order[k]] % C (order NULL = identity). */
int pq3d_split_index(const int32_t* flat, const int32_t* order, int K, int C, int32_t* q_of, int32_t* cls_of, void* stream);
/* Full-resolution masks [P, K] (float 0/1: integer majority vote of the points of each full-resolution segment, equal to
 * scatter_mean(...) > 0.5 bit for bit) and heatmaps [P, K] = sigmoid of the point's segment logit, for the K queries q_of
 * (already in output order).  votes int32 [n_fullseg, K] scratch, points = bincount(segment_to_full). */
int pq3d_instseg_fullres(const float* pred_masks, const int32_t* q_of, const int64_t* voxel2segment,
                         const int64_t* voxel_to_full, const int64_t* segment_to_full, int64_t P, int S, int Q, int K,
                         int n_fullseg, int32_t* votes, const int32_t* points, float* mask, float* heat, void* stream);

/* ---- backward companions (training step: autograd of the reference's decoder, trainer/query3d_trainer.py:18-28) ---- */

/* out_t[b1,b2][c][r] = bf16(scale * in[b1,b2][r][c] * (gate > 0 ? 1 : 0)) for r < R, zero for R <= r < Rp; optional
 * un-transposed bf16 copy out_c.  in: fp32 (in_fp32) or bf16; gate (optional, bf16, same indexing as in): ReLU backward.
 * Puts activations / gradients into the K-major layout pq3d_linear_bf16 needs for wgrad (dW = dYᵀ·X). */
int pq3d_transpose_cast(const void* in, int in_fp32, int64_t ld_in, int64_t in_b1, int64_t in_b2,
                        const void* gate, int64_t ld_gate, int64_t gate_b1, int64_t gate_b2,
                        void* out_t, int64_t ld_t, int64_t t_b1, int64_t t_b2,
                        void* out_c, int64_t ld_c, int64_t c_b1, int64_t c_b2,
                        int R, int C, int Rp, int B1, int B2, float scale, void* stream);

/* out[c] (+)= scale * sum_r in[r][c] * (gate[r][c] > 0 ? 1 : 0) — bias gradients; in fp32 or bf16, gate optional (bf16). */
int pq3d_colsum(const void* in, int in_fp32, int64_t ld, const void* gate, int64_t ld_gate, float* out, int R, int C,
                int accumulate, float scale, void* stream);

/* Backward of pq3d_add_layernorm / pq3d_add_layernorm_train: d_x[g] = gradient of the branch input y[g] (fp32 and / or
 * its bf16 copy d_x_bf16, the operand of the dgrad GEMM that follows, same group stride; with dropout it carries the
 * forward's keep mask and 1/(1-p)), d_res = sum_g (gradient of residual + dropout(y[g])), d_gamma / d_beta accumulated
 * with atomics into zero-initialised fp32 [G,D].  drop_p / seed_dev / site / row_w / rows_per_scene: exactly the
 * forward call's (0 / NULL without dropout or memory dropout).  Any output may be NULL. */
int pq3d_layernorm_bwd(const float* y, int64_t y_group_stride, const float* residual, const float* gamma,
                       const float* d_out, int G, float eps, int R, int D, float* d_x, int64_t dx_group_stride,
                       void* d_x_bf16, float* d_res, float* d_gamma, float* d_beta, float drop_p,
                       const uint32_t* seed_dev, uint32_t site, const float* row_w, int rows_per_scene, void* stream);

/* Training-mode tail of a residual block: out = sum_g w[b,g] * LN_g(residual + dropout_p(y[g])), w = 1/G or the
 * memory-dropout weights row_w [R/rows_per_scene, G] (keep / #kept per scene; query_encoder.py:145-153).  Dropout keeps
 * element e = (g*R + row)*D + col iff hash32(e ^ hash32(*seed_dev + site*0x9E3779B9)) >= p*2^32 (csrc/ptx.cuh,
 * restated in pq3d_b200/rng.py); the backward regenerates the mask from the same (seed, site).  Outputs as in
 * pq3d_add_layernorm.  Replaces `tgt = norm(tgt + dropout(tgt2))` (query_encoder.py:304-305, 386-387, 449-450). */
int pq3d_add_layernorm_train(const float* y, int64_t y_group_stride, const float* residual, const float* gamma,
                             const float* beta, int G, float eps, int R, int D, const float* pos, float* out_f32,
                             void* out_bf16, void* out_pos_bf16, float drop_p, const uint32_t* seed_dev, uint32_t site,
                             const float* row_w, int rows_per_scene, void* stream);

/* In-place dropout of a bf16 array (FFN hidden activations, query_encoder.py:384): element e kept iff the counter RNG
 * says so (see pq3d_add_layernorm_train), kept values scaled by 1/(1-p).  n multiple of 8. */
int pq3d_dropout_bf16(void* x, int64_t n, float drop_p, const uint32_t* seed_dev, uint32_t site, void* stream);

/* delta[b,h,n] = sum_d dO[b*N+n, h*64+d] * O[b*N+n, h*64+d] (bf16 in, fp32 out). */
int pq3d_attn_delta(const void* dO, const void* O, int64_t ld, float* delta, int B, int H, int N, void* stream);

/* Softmax backward on recomputed scores: P = ex2(S2 + bias - m)/l (0 where masked), dS2 = ln2 * P * (dP - delta);
 * S2, dP fp32 [B,H,N,ld]; outputs bf16 P (optional, may be NULL), dS [B,H,N,ld] and transposed Pt, dSt [B,H,ld,Np]
 * (columns n >= N zero).  ld multiple of 32, Np of 8. */
int pq3d_softmax_bwd(const float* S2, const float* dP, const float* delta, const float* m, const float* l,
                     const float* bias, int64_t bias_ld, const uint32_t* mask_bits, int64_t mask_b_stride,
                     int64_t mask_h_stride, int64_t mask_q_stride, void* P, void* dS, void* Pt, void* dSt,
                     int B, int H, int N, int S, int ld, int Np, void* stream);

/* Gradients of pairwise_loc_fc (modules/layers/transformers.py:196-199) from dS2 of the spatial self-attention:
 * d_w [H,5], d_b [H] accumulated with atomics (zero-initialise). */
int pq3d_spatial_bias_bwd(const float* pairwise_locs, const float* loc_w, const float* loc_b, const void* dS,
                          int64_t ld, float* d_w, float* d_b, int B, int H, int N, void* stream);

/* out = a + b (+ c), fp32, n multiple of 4. */
int pq3d_add3(const float* a, const float* b, const float* c, float* out, int64_t n, void* stream);

/* Per-step refresh of the kernels' operand copies of the fp32 parameters, ONE launch for the whole decoder.
 * segs_dev: DEVICE array of n_seg x 8 int64 words {src, dst_c, dst_t, rows, cols, ld_c, ld_t, flags}: src fp32
 * [rows, cols] dense -> dst_c[r*ld_c + c] (bf16, or fp32 when flags & 1; may be 0) and dst_t[c*ld_t + r] (bf16
 * transposed copy for the dgrad GEMMs; may be 0).  tile_start_dev: DEVICE int32 [n_seg] — index of each segment's
 * first 64x64 tile; total_tiles = number of tiles over all segments.  Replaces autocast's per-call weight casts
 * (every nn.Linear / MHA projection under torch.autocast in the reference's training step). */
int pq3d_pack_segments(const int64_t* segs_dev, const int32_t* tile_start_dev, int n_seg, int total_tiles,
                       void* stream);

/* Backward of the attention core in one kernel (scores never leave the SM).  Per (scene b, head h):
 *   S2 = Q2 K^T + bias (log2 domain, Q2 = Q pre-scaled as in pq3d_attention_fwd), P = ex2(S2 - m) / l (0 where masked),
 *   dP = dO V^T, dS2 = ln2 * P o (dP - delta);   dV = P^T dO,  dK = dS2^T Q2,  dQ += q_scale * dS2 K.
 *   Q  : bf16 [B*Nq, ldq], head h at columns q_col0 + h*64;  dO likewise (lddo, do_col0)
 *   K,V: bf16 row-major [B*S_pitch, ld], scene b at rows b*S_pitch, S valid rows; head h at columns col0 + h*64
 *   mask_bits / strides, bias / bias_ld: as in pq3d_attention_fwd (bias: fp32 [B, H, Nq, bias_ld])
 *   stat_m, stat_l: [B, H, Nq] saved by the forward; delta: [B, H, Nq] from pq3d_attn_delta
 *   dK, dV: bf16, same row layout as K / V (ld_dk, dk_col0, ...): rows < S of every head block are overwritten
 *   dQ: fp32 [B*Nq, ld_dq], ACCUMULATED with atomics at columns dq_col0 + h*64 (zero-fill before the call)
 *   dS_out: optional bf16 [B, H, Nq, ds_ld] (for the spatial-bias backward)
 *   drop_p / seed_dev / site: the forward's attention-probability dropout (pq3d_attention_fwd_train), regenerated here
 * Requires Nq <= 128.  Replaces autograd through torch/nn/functional.py:6630-6647 and
 * modules/layers/transformers.py:224-236 in the reference's training step. */
int pq3d_attention_bwd(const void* Q, int64_t ldq, int q_col0, const void* dO, int64_t lddo, int do_col0,
                       const void* K, int64_t ldk, int k_col0, const void* V, int64_t ldv, int v_col0, int S,
                       int S_pitch, const uint32_t* mask_bits, int64_t mask_b_stride, int64_t mask_h_stride,
                       int64_t mask_q_stride, const float* bias, int64_t bias_ld, const float* stat_m,
                       const float* stat_l, const float* delta, void* dK, int64_t ld_dk, int dk_col0, void* dV,
                       int64_t ld_dv, int dv_col0, float* dQ, int64_t ld_dq, int dq_col0, void* dS_out,
                       int64_t ds_ld, int B, int H, int Nq, float q_scale, float drop_p, const uint32_t* seed_dev,
                       uint32_t site, void* stream);

/* Backward of pq3d_mask_head_finalize: d_raw[b,s,n] = d_logits[b,s,n] / (#valid memories at (b,s) + 1e-8), 0 where the
 * segment is padded (its logit was overwritten by -1e6) — written as the bf16 operand [B*S, Np] (columns N..Np zero) of
 * the products d_q = d_raw^T k and d_k = d_raw q.  masks: uint8 [n_mem + 1][B*S], memory masks first (1 = ignore), the
 * segment padding mask last.  Replaces autograd through modules/heads/mask_head.py:36-40. */
int pq3d_mask_head_finalize_bwd(const float* d_logits, const uint8_t* masks, int n_mem, void* d_raw_bf16, int B, int S,
                                int N, int Np, void* stream);

/* Backward of pq3d_gate_mix: g = sigmoid(gate_logits); d_update = d_out*g, d_query = d_out*(1-g),
 * d_gate_logits = d_out*(update - query)*g*(1-g) (fp32 and its bf16 copy).  Replaces autograd through
 * modules/grounding/query_encoder.py:168-170. */
int pq3d_gate_mix_bwd(const float* gate_logits, const float* query, const float* update, const float* d_out,
                      float* d_gate_logits, void* d_gate_logits_bf16, float* d_update, float* d_query, int64_t n,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PQ3D_B200_H_ */
