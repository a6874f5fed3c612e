// masked_ce.cu -- cross entropy with ignore_index over [B,C,HW] fp32 logits, forward
// and backward, plus the scalar algebra of the unsupervised loss.  Replaces
// F.cross_entropy(predict, target, ignore_index=255) at loss_helper.py:46 and the
// nn.CrossEntropyLoss inside Criterion (loss_helper.py:265,313-319): log_softmax,
// nll_loss2d and their two backward kernels (four full passes over a 354 MB tensor
// for V16) become one read pass (fwd) and one read + one write pass (bwd).
// HBM-bound: fwd 4C+8 B/pixel, bwd 8C+8 B/pixel.  Ignored pixels skip the logit reads.
#include "common.cuh"

namespace u2pl {

template <int C>
__global__ void __launch_bounds__(256)
ce_fwd_kernel(const float *__restrict__ logits, const int64_t *__restrict__ target,
              uint32_t HW, uint32_t N, int64_t ignore,
              float *__restrict__ part_sum, uint32_t *__restrict__ part_cnt)
{
    float acc = 0.0f;
    uint32_t cnt = 0;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < N; i += gridDim.x * 256u) {
        const int64_t t = __ldg(target + i);
        if (t == ignore || t < 0 || t >= C) continue;
        const uint32_t b = i / HW, p = i - b * HW;
        const float *x = logits + static_cast<size_t>(b) * C * HW + p;
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = __ldg(x + static_cast<size_t>(c) * HW);
        float m = v[0];
#pragma unroll
        for (int c = 1; c < C; ++c) m = fmaxf(m, v[c]);
        float S = 0.0f, xt = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            S += expf(v[c] - m);
            xt = (c == static_cast<int>(t)) ? v[c] : xt;
        }
        acc += logf(S) - (xt - m);
        ++cnt;
    }
    acc = warp_sum(acc);
    cnt = static_cast<uint32_t>(warp_sum_i(static_cast<int>(cnt)));
    __shared__ float ws[8];
    __shared__ uint32_t wc[8];
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = acc; wc[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        uint32_t c = 0;
        for (int w = 0; w < 8; ++w) { s += ws[w]; c += wc[w]; }
        part_sum[blockIdx.x] = s;
        part_cnt[blockIdx.x] = c;
    }
}

__global__ void __launch_bounds__(256)
ce_fwd_kernel_anyC(const float *__restrict__ logits, const int64_t *__restrict__ target,
                   uint32_t C, uint32_t HW, uint32_t N, int64_t ignore,
                   float *__restrict__ part_sum, uint32_t *__restrict__ part_cnt)
{
    float acc = 0.0f;
    uint32_t cnt = 0;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < N; i += gridDim.x * 256u) {
        const int64_t t = __ldg(target + i);
        if (t == ignore || t < 0 || t >= C) continue;
        const uint32_t b = i / HW, p = i - b * HW;
        const float *x = logits + static_cast<size_t>(b) * C * HW + p;
        float m = __ldg(x);
        for (uint32_t c = 1; c < C; ++c) m = fmaxf(m, __ldg(x + static_cast<size_t>(c) * HW));
        float S = 0.0f;
        for (uint32_t c = 0; c < C; ++c) S += expf(__ldg(x + static_cast<size_t>(c) * HW) - m);
        acc += logf(S) - (__ldg(x + static_cast<size_t>(t) * HW) - m);
        ++cnt;
    }
    acc = warp_sum(acc);
    cnt = static_cast<uint32_t>(warp_sum_i(static_cast<int>(cnt)));
    __shared__ float ws[8];
    __shared__ uint32_t wc[8];
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = acc; wc[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        uint32_t c = 0;
        for (int w = 0; w < 8; ++w) { s += ws[w]; c += wc[w]; }
        part_sum[blockIdx.x] = s;
        part_cnt[blockIdx.x] = c;
    }
}

// fixed-order reduction of the per-block partials (deterministic run to run)
__global__ void __launch_bounds__(256)
ce_reduce_kernel(const float *__restrict__ part_sum, const uint32_t *__restrict__ part_cnt, int nblocks,
                 float *__restrict__ nll_sum, int64_t *__restrict__ n_used)
{
    __shared__ double sd[256];
    __shared__ unsigned long long sc[256];
    double s = 0.0;
    unsigned long long c = 0;
    for (int j = threadIdx.x; j < nblocks; j += 256) { s += static_cast<double>(part_sum[j]); c += part_cnt[j]; }
    sd[threadIdx.x] = s;
    sc[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sd[threadIdx.x] += sd[threadIdx.x + o]; sc[threadIdx.x] += sc[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { *nll_sum = static_cast<float>(sd[0]); *n_used = static_cast<int64_t>(sc[0]); }
}

template <int C>
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const float *__restrict__ logits, const int64_t *__restrict__ target,
              uint32_t HW, uint32_t N, int64_t ignore, const float *__restrict__ scale,
              float *__restrict__ grad)
{
    const float sc = __ldg(scale);
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < N; i += gridDim.x * 256u) {
        const int64_t t = __ldg(target + i);
        const uint32_t b = i / HW, p = i - b * HW;
        const size_t off = static_cast<size_t>(b) * C * HW + p;
        float *g = grad + off;
        if (t == ignore || t < 0 || t >= C) {
#pragma unroll
            for (int c = 0; c < C; ++c) g[static_cast<size_t>(c) * HW] = 0.0f;
            continue;
        }
        const float *x = logits + off;
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = __ldg(x + static_cast<size_t>(c) * HW);
        float m = v[0];
#pragma unroll
        for (int c = 1; c < C; ++c) m = fmaxf(m, v[c]);
        float S = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) { v[c] = expf(v[c] - m); S += v[c]; }
        const float r = sc / S;
#pragma unroll
        for (int c = 0; c < C; ++c)
            g[static_cast<size_t>(c) * HW] = v[c] * r - ((c == static_cast<int>(t)) ? sc : 0.0f);
    }
}

__global__ void __launch_bounds__(256)
ce_bwd_kernel_anyC(const float *__restrict__ logits, const int64_t *__restrict__ target,
                   uint32_t C, uint32_t HW, uint32_t N, int64_t ignore, const float *__restrict__ scale,
                   float *__restrict__ grad)
{
    const float sc = __ldg(scale);
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < N; i += gridDim.x * 256u) {
        const int64_t t = __ldg(target + i);
        const uint32_t b = i / HW, p = i - b * HW;
        const size_t off = static_cast<size_t>(b) * C * HW + p;
        float *g = grad + off;
        if (t == ignore || t < 0 || t >= C) {
            for (uint32_t c = 0; c < C; ++c) g[static_cast<size_t>(c) * HW] = 0.0f;
            continue;
        }
        const float *x = logits + off;
        float m = __ldg(x);
        for (uint32_t c = 1; c < C; ++c) m = fmaxf(m, __ldg(x + static_cast<size_t>(c) * HW));
        float S = 0.0f;
        for (uint32_t c = 0; c < C; ++c) S += expf(__ldg(x + static_cast<size_t>(c) * HW) - m);
        const float r = sc / S;
        for (uint32_t c = 0; c < C; ++c)
            g[static_cast<size_t>(c) * HW] =
                expf(__ldg(x + static_cast<size_t>(c) * HW) - m) * r - ((c == static_cast<uint32_t>(t)) ? sc : 0.0f);
    }
}

__global__ void unsup_finalize_kernel(const float *__restrict__ nll_sum, const int64_t *__restrict__ n_kept,
                                      float total_pixels, const float *__restrict__ upstream,
                                      float *__restrict__ loss, float *__restrict__ bwd_scale)
{
    // loss_helper.py:44-46: weight = B*H*W / #kept ; loss = weight * mean_kept(nll)
    const float k = static_cast<float>(*n_kept);
    const float weight = total_pixels / k;
    if (loss) *loss = weight * (*nll_sum / k);
    if (bwd_scale) *bwd_scale = (upstream ? *upstream : 1.0f) * weight / k;
}

constexpr int kCeMaxBlocks = 148 * 8;

static int ce_grid(uint32_t N)
{
    const long long need = (static_cast<long long>(N) + 255) / 256;
    return static_cast<int>(need < kCeMaxBlocks ? need : kCeMaxBlocks);
}

}  // namespace u2pl

using namespace u2pl;

extern "C" size_t u2pl_ce_ws_bytes(int64_t, int64_t) { return static_cast<size_t>(kCeMaxBlocks) * 8; }

extern "C" int u2pl_ce_forward(const float *logits, const int64_t *target, int64_t B, int64_t C, int64_t HW,
                               int64_t ignore, float *nll_sum, int64_t *n_used,
                               void *ws, size_t ws_bytes, void *stream)
{
    if (B <= 0 || C <= 0 || HW <= 0 || B * HW >= (1LL << 31)) return bad_arg("ce_forward: bad shape");
    if (ws_bytes < static_cast<size_t>(kCeMaxBlocks) * 8) { set_error("ce_forward: workspace too small"); return U2PL_E_WS_SMALL; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t N = static_cast<uint32_t>(B * HW), hw = static_cast<uint32_t>(HW);
    float *ps = static_cast<float *>(ws);
    uint32_t *pcnt = reinterpret_cast<uint32_t *>(ps + kCeMaxBlocks);
    const int grid = ce_grid(N);
    switch (C) {
        case 19: ce_fwd_kernel<19><<<grid, 256, 0, s>>>(logits, target, hw, N, ignore, ps, pcnt); break;
        case 21: ce_fwd_kernel<21><<<grid, 256, 0, s>>>(logits, target, hw, N, ignore, ps, pcnt); break;
        default: ce_fwd_kernel_anyC<<<grid, 256, 0, s>>>(logits, target, static_cast<uint32_t>(C), hw, N, ignore, ps, pcnt);
    }
    ce_reduce_kernel<<<1, 256, 0, s>>>(ps, pcnt, grid, nll_sum, n_used);
    return check_launch("ce_forward", 2);
}

extern "C" int u2pl_ce_backward(const float *logits, const int64_t *target, int64_t B, int64_t C, int64_t HW,
                                int64_t ignore, const float *scale, float *grad, void *stream)
{
    if (B <= 0 || C <= 0 || HW <= 0 || B * HW >= (1LL << 31)) return bad_arg("ce_backward: bad shape");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t N = static_cast<uint32_t>(B * HW), hw = static_cast<uint32_t>(HW);
    const int grid = ce_grid(N);
    switch (C) {
        case 19: ce_bwd_kernel<19><<<grid, 256, 0, s>>>(logits, target, hw, N, ignore, scale, grad); break;
        case 21: ce_bwd_kernel<21><<<grid, 256, 0, s>>>(logits, target, hw, N, ignore, scale, grad); break;
        default: ce_bwd_kernel_anyC<<<grid, 256, 0, s>>>(logits, target, static_cast<uint32_t>(C), hw, N, ignore, scale, grad);
    }
    return check_launch("ce_backward");
}

extern "C" int u2pl_unsup_finalize(const float *nll_sum, const int64_t *n_kept, int64_t total_pixels,
                                   const float *upstream, float *loss, float *bwd_scale, void *stream)
{
    unsup_finalize_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(nll_sum, n_kept, static_cast<float>(total_pixels),
                                                                        upstream, loss, bwd_scale);
    return check_launch("unsup_finalize");
}
