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

/* unsup loss scalar of loss_helper.py:44-46 on device:
 *   loss = (total_pixels / n_kept) * (nll_sum / n_kept);   bwd_scale = upstream * total_pixels / n_kept^2
 * n_kept == 0 gives loss = NaN like the reference (0/0). upstream may be NULL (= 1). */
int u2pl_unsup_finalize(const float *nll_sum, const int64_t *n_kept, int64_t total_pixels,
                        const float *upstream, float *loss, float *bwd_scale, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* U2PL_B200_H_ */
