/*
 * u2pl_b200.h -- C ABI of libu2pl_b200.so (sm_100a).
 *
 * The reference (Haochen-Wang409/U2PL) is pure Python; its "FFI" for the hot
 * path is the Python call boundary u2pl.utils.loss_helper / u2pl.utils.utils /
 * u2pl.models.model_helper.  This header is what the
 * host-side mirror of that boundary (u2pl_b200/u2pl/..., loaded with ctypes)
 * binds.  Conventions:
 *   - every pointer is a DEVICE pointer unless the parameter name starts with h_;
 *   - tensors are contiguous; layouts are spelled per function;
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises the
 *     host, everything is enqueued on `stream`;
 *   - return value: 0 on success, otherwise a cudaError_t (>0) or a negative
 *     U2PL_E_* code; u2pl_last_error() returns a static description.
 *   - no torch types, no C++ types.
 */
#ifndef U2PL_B200_H_
#define U2PL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define U2PL_ABI_VERSION 1

#define U2PL_E_BADARG   (-1)   /* shape / size outside what the kernels support      */
#define U2PL_E_WS_SMALL (-2)   /* workspace smaller than the *_ws_bytes() answer      */

#define U2PL_MAX_QUANTILES 4   /* percentiles resolved by one u2pl_entropy_thresholds */

int         u2pl_abi_version(void);
const char *u2pl_last_error(void);
/* number of kernels this library has launched in this process (bench.py gpu_launches) */
int64_t     u2pl_launch_count(void);

/* ------------------------------------------------------------------------
 * A6/A7/A8  softmax-entropy + adaptive percentile thresholds
 * replaces: u2pl/utils/loss_helper.py:35-40 (prob, entropy, np.percentile) and
 *           train_semi.py:402-415 (same entropy, two more percentiles).
 *
 * logits  [B, C, HW] fp32 (NCHW contiguous, HW = H*W); target [B, HW] int64.
 * entropy [B, HW] fp32 out.  Pixels with target == ignore are excluded from the
 * percentile population.  h_percents: nq (<= U2PL_MAX_QUANTILES) percentiles in
 * [0,100] (HOST array, read at call time); thresh: nq fp32 out (device);
 * n_valid: one int64 out (device) = size of the population.
 * Percentile = numpy 2.x `np.percentile(float32 data, q)`: float32 virtual
 * index, linear interpolation, two-sided lerp; resolved on device by a 3-pass
 * (12/10/10 bit) radix select, no host round trip.
 * If the population is empty thresh[] = NaN.
 * ---------------------------------------------------------------------- */
size_t u2pl_entropy_ws_bytes(int64_t B, int64_t HW);
int u2pl_entropy_thresholds(const float *logits, const int64_t *target,
                            int64_t B, int64_t C, int64_t HW, int64_t ignore,
                            const float *h_percents, int nq,
                            float *entropy, float *thresh, int64_t *n_valid,
                            void *ws, size_t ws_bytes, void *stream);

/* Two-level variant ("exact where it matters"): same thresholds (bit-identical) and therefore the same
 * masks as u2pl_entropy_thresholds, but the entropy map is evaluated with hardware ex2/lg2
 * (|error| <= 1e-4 against the contract arithmetic) and re-evaluated under the contract only for pixels
 * within 3e-4 of a target order statistic's 22-bit key bin; those exact values are stored back, so
 * every later `entropy <=/>= thresh[j]` comparison on the returned map is exact.  The heavy pass is
 * HBM-bound instead of issue-bound.  C in {19, 21}; other C fall through to the exact-everywhere path.
 * Workspace: u2pl_entropy_fast_ws_bytes (holds the per-target candidate lists). */
size_t u2pl_entropy_fast_ws_bytes(int64_t B, int64_t HW);
int u2pl_entropy_thresholds_fast(const float *logits, const int64_t *target,
                                 int64_t B, int64_t C, int64_t HW, int64_t ignore,
                                 const float *h_percents, int nq,
                                 float *entropy, float *thresh, int64_t *n_valid,
                                 void *ws, size_t ws_bytes, void *stream);

/* ------------------------------------------------------------------------
 * A6  reliable/unreliable partition of the pseudo-label target
 * replaces: loss_helper.py:41-44
 *   drop = (entropy >= thresh[thresh_idx]) & (target != ignore); target[drop] = ignore
 * target is rewritten in place; drop_mask (uint8, may be NULL) receives the
 * mask; n_kept (int64, device) = #(target != ignore) after the rewrite.
 * ---------------------------------------------------------------------- */
int u2pl_partition_target(const float *entropy, int64_t *target, int64_t n, int64_t ignore,
                          const float *thresh, int thresh_idx,
                          uint8_t *drop_mask, int64_t *n_kept, void *stream);

/* ------------------------------------------------------------------------
 * A6/A7  the whole chain in ONE launch: entropy + percentile thresholds + partition
 * replaces: loss_helper.py:35-44 (and train_semi.py:402-415 for the extra percentiles) in a single persistent
 * cooperative kernel (one CTA per SM, the CTA's slice of entropy keys stays in shared memory between the
 * passes; csrc/entropy_partition.cu `entropy_chain_kernel`).  Same results, bit for bit, as
 * u2pl_entropy_thresholds_fast followed by u2pl_partition_target on a copy of target_in:
 *   entropy [B,HW], thresh [nq], n_valid as above;
 *   target_out[i] = (entropy[i] >= thresh[part_idx] && target_in[i] != ignore) ? ignore : target_in[i]
 *   (target_in is NOT modified -- the caller's clone of loss_helper.py:381 is the output buffer),
 *   drop_mask (uint8, may be NULL), n_kept = #(target_out != ignore).
 * Falls back internally to the multi-launch kernels when B*HW exceeds what the CTAs' shared memory holds
 * (> 36864 pixels per SM) or C is not 19 / 21.  Workspace: u2pl_entropy_fast_ws_bytes.
 * ---------------------------------------------------------------------- */
int u2pl_entropy_partition_fused(const float *logits, const int64_t *target_in,
                                 int64_t B, int64_t C, int64_t HW, int64_t ignore,
                                 const float *h_percents, int nq, int part_idx,
                                 float *entropy, float *thresh, int64_t *n_valid,
                                 int64_t *target_out, uint8_t *drop_mask, int64_t *n_kept,
                                 void *ws, size_t ws_bytes, void *stream);

/* ------------------------------------------------------------------------
 * A8  low / high entropy masks (train_semi.py:408-418), evaluated at arbitrary
 * pixel positions so the nearest-neighbour down-sample of :427-453 can be fused:
 *   out_low[j]  = (entropy[idx[j]] <= thresh[lo_idx]) & (target[idx[j]] != ignore)
 *   out_high[j] = (entropy[idx[j]] >= thresh[hi_idx]) & (target[idx[j]] != ignore)
 * idx == NULL means identity (j = pixel).  Outputs are fp32 0/1 like the reference.
 * ---------------------------------------------------------------------- */
int u2pl_entropy_masks(const float *entropy, const int64_t *target, const int64_t *idx,
                       int64_t n_out, int64_t ignore, const float *thresh, int lo_idx, int hi_idx,
                       float *out_low, float *out_high, void *stream);

/* ------------------------------------------------------------------------
 * A3 / N2  consumers of the x4 bilinear up-sampling (align_corners=True) fused with it (csrc/upsample_ce.cu): the
 * full-resolution [B,C,H,W] logits of train_semi.py:317-324,344-358 are never materialised.
 * low [B,C,h,w] fp32 NCHW contiguous; target [B,H,W] int64; C in {19, 21} (u2pl_upsample_fused_supported).
 *   u2pl_up_softmax_max   replaces F.interpolate -> F.softmax(dim=1) -> torch.max(dim=1)   (train_semi.py:318-324)
 *   u2pl_upce_forward     replaces F.interpolate -> F.cross_entropy(ignore_index): nll_sum over valid pixels + their count
 *                         (train_semi.py:344-358 with loss_helper.py:313-319 / :44-46)
 *   u2pl_upce_backward    grad_low [B,C,h,w] = interpolation^T((softmax - onehot) * scale[0]) -- the CE backward and
 *                         upsample_bilinear2d_backward in one gather pass (no atomics, deterministic)
 * ---------------------------------------------------------------------- */
int u2pl_upsample_fused_supported(int64_t C);
size_t u2pl_upce_ws_bytes(void);
int u2pl_up_softmax_max(const float *low, int64_t B, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W,
                        float *out_prob, int64_t *out_label, void *stream);
int u2pl_upce_forward(const float *low, const int64_t *target, int64_t B, int64_t C, int64_t h, int64_t w,
                      int64_t H, int64_t W, int64_t ignore, float *nll_sum, int64_t *n_used,
                      void *ws, size_t ws_bytes, void *stream);
int u2pl_upce_backward(const float *low, const int64_t *target, int64_t B, int64_t C, int64_t h, int64_t w,
                       int64_t H, int64_t W, int64_t ignore, const float *scale, float *grad_low, void *stream);

/* ------------------------------------------------------------------------
 * A6/A12  cross entropy with ignore_index, forward and backward
 * replaces: F.cross_entropy(predict, target, ignore_index=255) at
 *           loss_helper.py:46 and nn.CrossEntropyLoss at :265,313-319.
 * logits [B, C, HW] fp32, target [B, HW] int64.
 * fwd: nll_sum (fp32, device) = sum over non-ignored pixels of -log_softmax[target];
 *      n_used (int64, device) = number of non-ignored pixels.  (deterministic:
 *      per-block partials in ws, reduced in fixed order.)
 * bwd: grad[b,c,i] = scale[0] * (softmax_c - [c == target]) for non-ignored
 *      pixels, 0 otherwise.  scale is a device fp32 scalar (so that
 *      weight / n_used * upstream never visits the host).
 * ---------------------------------------------------------------------- */
size_t u2pl_ce_ws_bytes(int64_t B, int64_t HW);
int u2pl_ce_forward(const float *logits, const int64_t *target, int64_t B, int64_t C, int64_t HW,
                    int64_t ignore, float *nll_sum, int64_t *n_used,
                    void *ws, size_t ws_bytes, void *stream);
int u2pl_ce_backward(const float *logits, const int64_t *target, int64_t B, int64_t C, int64_t HW,
                     int64_t ignore, const float *scale, float *grad, void *stream);

/* ------------------------------------------------------------------------
 * A12  OHEM pixel selection (OhemCrossEntropy2dTensor.forward, loss_helper.py:502-531)
 *   mask_prob = softmax(pred)[target]  (1 where target == ignore)
 *   kth = min(N, min_kept)-th smallest mask_prob (radix select instead of the reference's full sort)
 *   threshold = max(thresh, kth); kept = mask_prob <= threshold     (skipped if min_kept > #valid)
 *   new_target = kept ? target : ignore        (the CE itself is u2pl_ce_forward / backward)
 * kth_value (fp32) and n_valid (int64) are device outputs; workspace as u2pl_entropy_ws_bytes.
 * ---------------------------------------------------------------------- */
int u2pl_ohem_select(const float *logits, const int64_t *target, int64_t B, int64_t C, int64_t HW,
                     int64_t ignore, float thresh, int64_t min_kept,
                     int64_t *new_target, float *kth_value, int64_t *n_valid,
                     void *ws, size_t ws_bytes, void *stream);

/* unsup loss scalar of loss_helper.py:44-46 on device:
 *   loss = (total_pixels / n_kept) * (nll_sum / n_kept);   bwd_scale = upstream * total_pixels / n_kept^2
 * n_kept == 0 gives loss = NaN like the reference (0/0). upstream may be NULL (= 1). */
int u2pl_unsup_finalize(const float *nll_sum, const int64_t *n_kept, int64_t total_pixels,
                        const float *upstream, float *loss, float *bwd_scale, void *stream);

/* ------------------------------------------------------------------------
 * A9/A10  class-wise memory-bank contrastive loss
 * replaces: u2pl/utils/loss_helper.py:51-235 (compute_contra_memobank_loss) and
 *           u2pl/utils/utils.py:28-47 (dequeue_and_enqueue).
 *
 * Pixels are numbered row-major over [Bl+Bu, h, w] (labelled images first), P of them.
 * Class membership is one uint32 bitmask per pixel (C <= 32); "blocks" are runs of 256
 * pixels (u2pl_contra_num_blocks).  Feature tensors [N, D, h, w] are addressed by element
 * strides: element (n, d, pixel p) at n*sn + d*sd + p*sp  (NCHW: sd = h*w, sp = 1;
 * channels-last: sd = 1, sp = D).
 * ---------------------------------------------------------------------- */

/* label[:, c] != 0  ->  bit c.   onehot [B, C, hw] int64, bits [B*hw].  (loss_helper.py:80-81 operands) */
int u2pl_onehot_to_bits(const int64_t *onehot, int64_t B, int64_t C, int64_t hw, uint32_t *bits, void *stream);

/* A8 fused: nearest-neighbour down-sampled low/high masks + class bitmask straight from the label
 * maps and the full-resolution entropy -- what train_semi.py:408-465 builds through label_onehot
 * (utils.py:50-59, scatter quirk Q8 reproduced) and four F.interpolate(mode="nearest") calls,
 * without the two [B,C,H,W] fp32 one-hot temporaries.  label_l [Bl,H,W], label_u [Bu,H,W] int64;
 * entropy [Bu,H,W]; outputs indexed over [(Bl+Bu), h, w]. */
int u2pl_contra_prep_lowres(const int64_t *label_l, const int64_t *label_u, const float *entropy,
                            const float *thresh, int lo_idx, int hi_idx,
                            int64_t Bl, int64_t Bu, int64_t H, int64_t W, int64_t h, int64_t w,
                            int64_t C, int64_t ignore, int negative_high_entropy,
                            uint32_t *label_bits, float *low_mask, float *high_mask, void *stream);

int64_t u2pl_contra_num_blocks(int64_t P);

/* Loop 1 of the reference (loss_helper.py:103-154) for every class at once.
 * bits3 [3][P]: 0 low-valid (:104), 1 anchor candidates (:108-110), 2 negative keys (:111-140).
 * blockcnt/blockoff [3][C][nb]: per-block member counts and their exclusive scan per class;
 * totals [3][C]: low_valid.sum() (:152), len(seg_feat_low_entropy_list[i]) (:175), #keys (:142). */
int u2pl_contra_classify(const uint32_t *label_bits, const float *prob_l, const float *prob_u,
                         const float *low_mask, const float *high_mask,
                         int64_t Bl, int64_t Bu, int64_t C, int64_t hw,
                         float current_class_threshold, float current_class_negative_threshold,
                         int low_rank, int high_rank,
                         uint32_t *bits3, uint32_t *blockcnt, uint32_t *blockoff, uint32_t *totals, void *stream);

/* class prototypes: mean of rep_teacher rows over low-valid pixels (:119-123).
 * partial: scratch [u2pl_contra_proto_parts()][C][D] fp32; proto [C][D] (NaN rows for empty classes). */
int64_t u2pl_contra_proto_parts(void);
int u2pl_contra_proto(const float *rep_teacher, int64_t sn, int64_t sd, int64_t sp,
                      int64_t P, int64_t C, int64_t D, int64_t hw,
                      const uint32_t *lv_bits, const uint32_t *lv_totals,
                      float *partial, float *proto, void *stream);

/* keys = rep_teacher[negative_mask] (:142) for all classes into packed [sum_c k_c, D], class-major,
 * pixel order inside a class.  class_base [C] = exclusive scan of the key counts (device). */
int u2pl_contra_pack_keys(const float *rep_teacher, int64_t sn, int64_t sd, int64_t sp,
                          int64_t P, int64_t C, int64_t D, int64_t hw,
                          const uint32_t *ng_bits, const uint32_t *blockoff_ng, const uint32_t *class_base,
                          float *packed, void *stream);

/* dequeue_and_enqueue (utils.py:28-47) on a device-resident ring buffer.  bank [rows][D];
 * desc: ndesc x 5 uint32 {src_first_row, dst_row_base, dst_first_pos, capacity, count}: row r of the
 * segment goes to bank[dst_row_base + (dst_first_pos + r) % capacity].  The FIFO arithmetic
 * (newest queue_size rows survive) is host bookkeeping: u2pl_b200/bank.py. */
int u2pl_bank_append(const float *src_rows, float *bank, int64_t D, const uint32_t *desc, int ndesc,
                     int64_t max_count, void *stream);

/* Loop 2 (:173-230) for all active classes in one launch: one warp per (class, query).
 * act_class [nact]: list position j whose anchors/prototype are used (quirk Q1: the bank rows in
 * neg_rows were sampled from class valid_classes[j]); a_ord [nact][nq]: anchor ordinals drawn by
 * torch.randint on the host (:179-181); neg_rows [nact][nq][nneg]: physical bank rows (:194-197).
 * Outputs: loss_q [nact*nq], grad_rows [nact*nq][D] (= d loss / d anchor row, already scaled by
 * 1/(nq*valid_seg)), anchor_pix [nact*nq], loss (scalar) = sum_q CE_q / (nq*valid_seg)  (:228-233). */
int u2pl_infonce_forward(const float *rep, int64_t sn, int64_t sd, int64_t sp,
                         int64_t P, int64_t D, int64_t hw,
                         const uint32_t *an_bits, const uint32_t *blockoff_an,
                         const int32_t *act_class, const int32_t *a_ord, const int32_t *neg_rows,
                         const float *proto, const float *bank,
                         int nact, int nq, int nneg, float temperature, int valid_seg,
                         float *loss_q, float *grad_rows, int32_t *anchor_pix, float *loss, void *stream);

/* Same loss with the bank SHARDED BY CLASS across the GPUs of one box (owner(c) = c mod world): the
 * negatives of an active class are read by the loss kernel itself straight from the owner's shard -- local memory or
 * a peer GPU's memory mapped with u2pl_shard_open, i.e. 1 KB row loads over NVLink -- instead of being gathered,
 * exchanged and re-read.  class_bank: DEVICE array [nact] of shard base pointers (one per active class);
 * neg_rows: rows inside that shard.  Everything else as u2pl_infonce_forward.
 * replaces: `negative_feat = memobank[valid_classes[i]][0].clone().cuda()` + index (loss_helper.py:192-200). */
int u2pl_infonce_forward_sharded(const float *rep, int64_t sn, int64_t sd, int64_t sp,
                                 int64_t P, int64_t D, int64_t hw,
                                 const uint32_t *an_bits, const uint32_t *blockoff_an,
                                 const int32_t *act_class, const int32_t *a_ord, const int32_t *neg_rows,
                                 const float *proto, const float *const *class_bank,
                                 int nact, int nq, int nneg, float temperature, int valid_seg,
                                 float *loss_q, float *grad_rows, int32_t *anchor_pix, float *loss, void *stream);

/* Bank shards that peers can map (CUDA IPC; same box, NVLink).  u2pl_shard_alloc: cudaMalloc + zero + export a
 * 64-byte handle; u2pl_shard_open: map a peer's shard from its handle; u2pl_shard_close: cudaFree (owned != 0) or
 * unmap.  The handles are exchanged by the host (u2pl_b200/bank.py, one all_gather_object at start-up).
 * replaces: train_semi.py:161-169 (bank construction) for world size > 1. */
int u2pl_shard_alloc(int64_t bytes, void **dptr, unsigned char *handle64);
int u2pl_shard_open(const unsigned char *handle64, void **dptr);
int u2pl_shard_close(void *dptr, int owned);

/* grad_rep[anchor pixel, :] += upstream * grad_rows (atomics: queries are sampled with replacement). */
int u2pl_infonce_backward(const float *grad_rows, const int32_t *anchor_pix, int nrows,
                          int64_t D, int64_t hw, int64_t sn, int64_t sd, int64_t sp,
                          const float *upstream, float *grad_rep, void *stream);

/* ------------------------------------------------------------------------
 * A1/A2  batch normalisation (+ReLU, +residual) on channels-last bf16 activations
 * replaces: nn.BatchNorm2d / nn.SyncBatchNorm + nn.ReLU + residual add as used in
 *           u2pl/models/resnet.py:120-140, base.py:22-76, decoder.py:57-100.
 * x, y, residual, dy, dx, dres: [M, C] bf16 row-major (= NHWC with M = N*H*W), 16-byte aligned;
 * C % 8 == 0, C/8 a power of two, C <= 2048.  Statistics, gamma/beta, running stats: fp32.
 *   u2pl_bn_stats            sums[0][c] = sum_m x, sums[1][c] = sum_m x^2   (partial: scratch
 *                            [u2pl_bn_parts()][2][C]); all-reduce `sums` across ranks for SyncBN
 *   u2pl_bn_finalize         mean, invstd (biased var + eps), running-stat update (unbiased var,
 *                            momentum; pass NULL to skip), scale = gamma*invstd, shift = beta - mean*scale
 *   u2pl_bn_fold             eval mode: scale/shift from the running statistics
 *   u2pl_bn_apply            y = [relu]( x*scale + shift [+ residual] )
 *   u2pl_bn_backward_reduce  sums[0] = sum g, sums[1] = sum g*xhat,  g = dy * (y > 0)  (y NULL: no ReLU)
 *   u2pl_bn_backward_elemt   dx = gamma*invstd*(g - sums[0]/count - xhat*sums[1]/count); dres = g (NULL: skip)
 * ---------------------------------------------------------------------- */
int64_t u2pl_bn_parts(void);
int u2pl_bn_stats(const void *x, int64_t M, int64_t C, float *partial, float *sums, void *stream);
int u2pl_bn_finalize(const float *sums, int64_t C, double count, const float *gamma, const float *beta,
                     float *running_mean, float *running_var, float momentum, float eps,
                     float *mean, float *invstd, float *scale, float *shift, void *stream);
int u2pl_bn_fold(int64_t C, const float *gamma, const float *beta, const float *running_mean,
                 const float *running_var, float eps, float *scale, float *shift, void *stream);
int u2pl_bn_apply(const void *x, const void *residual, const float *scale, const float *shift,
                  int64_t M, int64_t C, int relu, void *y, void *stream);
int u2pl_bn_backward_reduce(const void *dy, const void *x, const void *y, const float *mean, const float *invstd,
                            int64_t M, int64_t C, float *partial, float *sums, void *stream);
int u2pl_bn_backward_elemt(const void *dy, const void *x, const void *y, const float *mean, const float *invstd,
                           const float *gamma, const float *sums, double count, int64_t M, int64_t C,
                           float *coef /* scratch [3][C] */, void *dx, void *dres, void *stream);

/* ------------------------------------------------------------------------
 * A1  stem max-pooling: nn.MaxPool2d(kernel_size=3, stride=2, padding=1, ceil_mode=True) (resnet.py:185,282) on a
 *     channels-last bf16 tensor.  x [n,h,w,c] -> y [n,ho,wo,c] with ho = u2pl_maxpool3s2_out(h); tap [n,ho,wo,c] uint8 =
 *     winning window position kh*3+kw (ATen's tie rule: row-major scan, a later tap wins only if strictly greater or
 *     NaN), may be NULL when no backward follows.  backward: dx[n,h,w,c] = sum of dy over the windows whose tap points
 *     at (h,w) -- a gather, no atomics.  c % 8 == 0.
 * ---------------------------------------------------------------------- */
int64_t u2pl_maxpool3s2_out(int64_t n);
int u2pl_maxpool3s2_forward(const void *x, void *y, void *tap, int64_t n, int64_t h, int64_t w, int64_t c, void *stream);
int u2pl_maxpool3s2_backward(const void *dy, const void *tap, void *dx, int64_t n, int64_t h, int64_t w, int64_t c, void *stream);

/* ------------------------------------------------------------------------
 * A1  1x1 convolution as a tensor-core GEMM (tcgen05 + TMA), optional folded-BN + ReLU epilogue
 * replaces: conv1x1 (resnet.py:39-41) [+ eval-mode BatchNorm + ReLU] on channels-last activations.
 *   D[M,N] = act( (A[M,K] . B[N,K]^T) * scale[n] + shift[n] ),  A/B/D bf16 row-major, fp32 accumulation;
 *   scale/shift fp32 (NULL: plain product), relu != 0 applies max(.,0).  K % 8 == 0, N % 8 == 0,
 *   pointers 16-byte aligned.  M = N*H*W pixels, K = Cin, N = Cout.
 * ---------------------------------------------------------------------- */
int u2pl_gemm_bf16_tn(const void *A, const void *B, void *D, int64_t M, int64_t N, int64_t K,
                      const float *scale, const float *shift, int relu, void *stream);

/* ------------------------------------------------------------------------
 * A1  stride-1 "same" convolution as an implicit GEMM on the tensor cores (tcgen05 + 4-D TMA boxes, no im2col),
 *     epilogue = folded eval-mode BatchNorm scale/shift, residual add, ReLU
 * replaces: conv3x3 / conv1x1 (resnet.py:25-41) + BatchNorm (eval) + `out += identity` + ReLU (resnet.py:118-140),
 *           the dilated ASPP branches (base.py:38-75) and the decoder heads' convs (decoder.py:60-113) on the
 *           teacher's pseudo-label forward (train_semi.py:318-319).
 *   x [n,h,w,cin], out / residual [n,h,w,cout] bf16 channels-last dense; wgt [cout, ksize, ksize, cin] bf16 (the
 *   memory of a channels-last weight tensor); padding = dilation * (ksize / 2); ksize 1 or 3; fp32 accumulation;
 *   out = act((conv) * scale[co] + shift[co] + residual); scale/shift/residual may be NULL.  cin % 8 == 0,
 *   cout % 8 == 0, pointers 16-byte aligned.
 * ---------------------------------------------------------------------- */
int u2pl_conv_bf16_nhwc(const void *x, const void *wgt, void *out, int64_t n, int64_t h, int64_t w, int64_t cin,
                        int64_t cout, int ksize, int dilation, const float *scale, const float *shift,
                        const void *residual, int relu, void *stream);

/* Train-mode variant: raw convolution output (bf16) PLUS the per-channel batch statistics of the stored values in the
 * same pass -- the epilogue accumulates sum and sum of squares per (pixel tile, channel) into stat_part
 * [u2pl_conv_stat_parts(n,h,w,ksize)][2][cout] and a second tiny kernel reduces them to sums [2][cout], the input
 * u2pl_bn_finalize expects.  replaces: conv (resnet.py:25-41) + the statistics pass of F.batch_norm(training=True)
 * (u2pl_bn_stats) on the student / train-mode teacher forwards (train_semi.py:340-341, 362-364). */
int64_t u2pl_conv_stat_parts(int64_t n, int64_t h, int64_t w, int ksize);
int u2pl_conv_bf16_nhwc_stats(const void *x, const void *wgt, void *out, int64_t n, int64_t h, int64_t w,
                              int64_t cin, int64_t cout, int ksize, int dilation, float *stat_part, float *sums,
                              void *stream);

/* General form: additionally applies act_in(x * in_scale[ci] + in_shift[ci]) to the INPUT while its tiles sit in shared
 * memory (in_scale / in_shift [cin] or NULL, in_relu) -- i.e. the previous layer's train-mode BatchNorm + ReLU is folded
 * into this convolution's operand load, so the normalised activation never exists in HBM; zero padding applies to the
 * activated tensor.  Any combination with the output epilogue (scale/shift/residual/relu) and the statistics epilogue
 * (stat_part + sums, both or neither) is allowed.
 * replaces: bn(x) -> relu -> conv chains inside Bottleneck.forward (resnet.py:118-131) on the train-mode forwards. */
int u2pl_conv_bf16_nhwc_ex(const void *x, const void *wgt, void *out, int64_t n, int64_t h, int64_t w, int64_t cin,
                           int64_t cout, int ksize, int dilation, const float *in_scale, const float *in_shift,
                           int in_relu, const float *scale, const float *shift, const void *residual, int relu,
                           float *stat_part, float *sums, void *stream);

/* Weight gradient of the same stride-1 "same" 3x3 (dilated) convolution on the tensor cores, operands read in place
 * from the channels-last tensors (both are MN-major for this contraction; no transposes, no cropped copies):
 *   partial[s][r*3+s'][co][ci] = sum over the pixels of K-split s of gout[n,h,w,co] * x[n,h+(r-1)d,w+(s'-1)d,ci]
 * x [n,h,w,cin], gout [n,h,w,cout] bf16 dense; partial fp32 [u2pl_conv_wgrad_splits(...)][9][cout][cin] -- the caller
 * sums over the first axis and permutes to the weight layout.  cin % 8 == 0, cout % 8 == 0.
 * replaces: the weight-gradient half of convolution_backward for conv3x3 (resnet.py:25-36), ASPP (base.py:38-75) and
 * decoder convs (decoder.py:60-113) in loss.backward() (train_semi.py:527). */
int u2pl_conv_wgrad_splits(int64_t n, int64_t h, int64_t w, int64_t cin, int64_t cout);
int u2pl_conv_wgrad_bf16_nhwc(const void *x, const void *gout, float *partial, int64_t n, int64_t h, int64_t w,
                              int64_t cin, int64_t cout, int dilation, void *stream);

/* ------------------------------------------------------------------------
 * A2  SyncBatchNorm statistics exchange over peer-mapped memory (one NVSwitch box)
 * replaces: the all_gather / all_reduce inside nn.SyncBatchNorm (reference base.py:6-8), ~350 per training step.
 * Every rank allocates an exchange region of u2pl_peer_region_bytes() with u2pl_shard_alloc (zero-filled), exports it
 * through CUDA IPC and maps every peer's region with u2pl_shard_open.  u2pl_peer_allreduce_f32 then sums buf[0..n) in
 * place across the `world` ranks in ONE single-CTA kernel: push into every peer's slot, system-scope fence + flag,
 * wait for all flags, sum in rank order (bitwise identical on every rank).  peer_bases: HOST array of `world` device
 * pointers (own region at index `rank`).  seq: 1, 2, 3, ... identical on all ranks (double-buffering by parity).
 * n <= u2pl_peer_max_floats() (4096).  All ranks must issue the same sequence of calls.
 * ---------------------------------------------------------------------- */
int64_t u2pl_peer_region_bytes(void);
int64_t u2pl_peer_max_floats(void);
int u2pl_peer_allreduce_f32(float *buf, int64_t n, void *const *peer_bases, int rank, int world, uint32_t seq, void *stream);

/* ------------------------------------------------------------------------
 * A13 / N3  SGD (momentum, weight decay) + EMA of the teacher parameters, one multi-tensor kernel
 * replaces: optimizer.step() (torch.optim.SGD built by lr_helper.py:12-27) and the EMA loop of train_semi.py:531-548
 * (~1.5k tiny launches).  tensor_table: DEVICE array of records of u2pl_sgd_tensor_bytes() (= 56) bytes each:
 *   { float *p; const float *g; float *m; float *t (or NULL); int64 n; float lr, wd; int32 first; int32 pad }
 * chunk_table: DEVICE array of n_chunks {uint32 tensor index, uint32 chunk index}; chunk = u2pl_sgd_chunk_elems() elements.
 *   d = g + wd*p;  m = first ? d : momentum*m + d;  p -= lr*m;  if do_ema and t: t = ema_decay*t + ema_one_minus*p
 * rounded where torch.optim.SGD (multi-tensor) and the reference's EMA expression round; ema_one_minus is the caller's
 * float32(1 - decay) evaluated in double, as Python evaluates `(1 - ema_decay)` (train_semi.py:546) -- not 1.0f - ema_decay.
 * ---------------------------------------------------------------------- */
int64_t u2pl_sgd_tensor_bytes(void);
int64_t u2pl_sgd_chunk_elems(void);
int u2pl_sgd_ema_step(const void *tensor_table, const void *chunk_table, int64_t n_chunks, float momentum, float ema_decay,
                      float ema_one_minus, int do_ema, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* U2PL_B200_H_ */
