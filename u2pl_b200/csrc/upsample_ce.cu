// upsample_ce.cu -- the x4 bilinear up-sampling (align_corners=True) of the low-resolution logits FUSED with what
// consumes it, so the 354 MB full-resolution [B,C,H,W] tensors of train_semi.py:317-324,344-358 never exist:
//
//   up_softmax_max   teacher pseudo labels (train_semi.py:318-324): F.interpolate -> softmax(dim=1) -> max(dim=1)
//                    reads [B,C,h,w] (22 MB, L2 resident), writes max-prob fp32 + arg-max int64 at full resolution
//   upce_forward     F.interpolate -> F.cross_entropy(ignore_index) (train_semi.py:344-358, loss_helper.py:46,313-319):
//                    per full-resolution pixel interpolate the C logits from their four low-resolution taps, log-sum-exp,
//                    nll; fixed-order two-stage reduction
//   upce_backward    gradient w.r.t. the LOW-resolution logits directly (the transpose of the interpolation applied to
//                    (softmax - onehot) * scale) as a GATHER: one thread per low-resolution pixel walks the <= 7x7
//                    full-resolution pixels whose taps include it.  No atomics (deterministic), no full-resolution
//                    gradient tensor (ATen: 354 MB written by the CE backward, re-read by upsample_bilinear2d_backward's
//                    atomic scatter, 2.1 ms per call on B200).
//
// Interpolation arithmetic = ATen's upsample_bilinear2d (align_corners=True): scale = (in-1)/(out-1) in fp32,
// src = scale*dst, i0 = (int)src, l1 = src - i0, l0 = 1 - l1, i1 = i0 + (i0 < in-1);
// v = l0y*(l0x*a + l1x*b) + l1y*(l0x*c + l1x*d).  Floating-point results (losses 1e-5, gradients 1e-6 vs torch).
#include "common.cuh"

namespace u2pl {

struct UpGeom {
    int h, w, H, W;
    float sy, sx;                      // (h-1)/(H-1), (w-1)/(W-1); 0 when the output extent is 1
};

__device__ __forceinline__ void src_coord(int dst, float scale, int in_size, int &i0, int &i1, float &l0, float &l1)
{
    const float s = scale * static_cast<float>(dst);
    i0 = static_cast<int>(s);
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    l1 = s - static_cast<float>(i0);
    l0 = 1.0f - l1;
}

template <int C>
__device__ __forceinline__ void interp_logits(const float *__restrict__ low, const UpGeom &g, int b, int i, int j, float (&v)[C])
{
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    src_coord(i, g.sy, g.h, y0, y1, ly0, ly1);
    src_coord(j, g.sx, g.w, x0, x1, lx0, lx1);
    const size_t plane = static_cast<size_t>(g.h) * g.w;
    const float *p = low + static_cast<size_t>(b) * C * plane;
    const int o00 = y0 * g.w + x0, o01 = y0 * g.w + x1, o10 = y1 * g.w + x0, o11 = y1 * g.w + x1;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float *q = p + static_cast<size_t>(c) * plane;
        v[c] = ly0 * (lx0 * __ldg(q + o00) + lx1 * __ldg(q + o01)) + ly1 * (lx0 * __ldg(q + o10) + lx1 * __ldg(q + o11));
    }
}

// ------------------------------------------------------------------ A3: bilinear -> softmax -> max
template <int C>
__global__ void __launch_bounds__(256)
up_softmax_max_kernel(const float *__restrict__ low, int B, UpGeom g, float *__restrict__ out_prob, int64_t *__restrict__ out_label)
{
    const uint32_t HW = static_cast<uint32_t>(g.H) * g.W, N = static_cast<uint32_t>(B) * HW;
    for (uint32_t idx = blockIdx.x * 256u + threadIdx.x; idx < N; idx += gridDim.x * 256u) {
        const int b = idx / HW, r = idx - b * HW, i = r / g.W, j = r - i * g.W;
        float v[C];
        interp_logits<C>(low, g, b, i, j, v);
        float m = v[0];
        int am = 0;
#pragma unroll
        for (int c = 1; c < C; ++c) { if (v[c] > m) { m = v[c]; am = c; } }       // first maximum, like torch.max
        float S = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) S += expf(v[c] - m);
        out_prob[idx] = 1.0f / S;                                                 // exp(m - m) / S
        out_label[idx] = am;
    }
}

// ------------------------------------------------------------------ N2: bilinear -> cross entropy, forward
template <int C>
__global__ void __launch_bounds__(256)
upce_fwd_kernel(const float *__restrict__ low, const int64_t *__restrict__ target, int B, UpGeom g, int64_t ignore,
                float *__restrict__ part_sum, uint32_t *__restrict__ part_cnt)
{
    const uint32_t HW = static_cast<uint32_t>(g.H) * g.W, N = static_cast<uint32_t>(B) * HW;
    float acc = 0.0f;
    uint32_t cnt = 0;
    for (uint32_t idx = blockIdx.x * 256u + threadIdx.x; idx < N; idx += gridDim.x * 256u) {
        const int64_t t = __ldg(target + idx);
        if (t == ignore || t < 0 || t >= C) continue;
        const int b = idx / HW, r = idx - b * HW, i = r / g.W, j = r - i * g.W;
        float v[C];
        interp_logits<C>(low, g, b, i, j, v);
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
        for (int k = 0; k < 8; ++k) { s += ws[k]; c += wc[k]; }
        part_sum[blockIdx.x] = s;
        part_cnt[blockIdx.x] = c;
    }
}

__global__ void __launch_bounds__(256)
upce_reduce_kernel(const float *__restrict__ part_sum, const uint32_t *__restrict__ part_cnt, int nblocks,
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

// ------------------------------------------------------------------ N2: backward, gather form
// Low-resolution pixel (y,x) receives from full-resolution pixel (i,j) the weight wy(i,y)*wx(j,x) where wy(i,y) = l0 if
// y == y0(i), plus l1 if y == y1(i) (both can hold at the clamped last row).  Rows i with y0(i) in {y-1, y} are the only
// candidates: i in [ceil((y-1)/sy), floor((y+1)/sy)] -- a window of <= 2/sy + 1 rows, searched exactly by evaluating
// src_coord on a conservatively widened range.
__device__ __forceinline__ void window(int y, float scale, int out_size, int &lo, int &hi)
{
    if (scale <= 0.0f) { lo = 0; hi = out_size - 1; return; }
    const float inv = 1.0f / scale;
    lo = max(0, static_cast<int>(floorf((static_cast<float>(y) - 1.0f) * inv)) - 1);
    hi = min(out_size - 1, static_cast<int>(ceilf((static_cast<float>(y) + 1.0f) * inv)) + 1);
}

template <int C>
__global__ void __launch_bounds__(128)
upce_bwd_kernel(const float *__restrict__ low, const int64_t *__restrict__ target, int B, UpGeom g, int64_t ignore,
                const float *__restrict__ scale, float *__restrict__ grad_low)
{
    const uint32_t hw = static_cast<uint32_t>(g.h) * g.w, P = static_cast<uint32_t>(B) * hw;
    const uint32_t idx = blockIdx.x * 128u + threadIdx.x;
    if (idx >= P) return;
    const int b = idx / hw, r = idx - b * hw, y = r / g.w, x = r - y * g.w;
    const float sc = __ldg(scale);
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.0f;
    int ilo, ihi, jlo, jhi;
    window(y, g.sy, g.H, ilo, ihi);
    window(x, g.sx, g.W, jlo, jhi);
    const int64_t *tb = target + static_cast<size_t>(b) * g.H * g.W;
    for (int i = ilo; i <= ihi; ++i) {
        int y0, y1;
        float ly0, ly1;
        src_coord(i, g.sy, g.h, y0, y1, ly0, ly1);
        const float wy = ((y0 == y) ? ly0 : 0.0f) + ((y1 == y) ? ly1 : 0.0f);
        if (wy == 0.0f) continue;
        for (int j = jlo; j <= jhi; ++j) {
            int x0, x1;
            float lx0, lx1;
            src_coord(j, g.sx, g.w, x0, x1, lx0, lx1);
            const float wx = ((x0 == x) ? lx0 : 0.0f) + ((x1 == x) ? lx1 : 0.0f);
            if (wx == 0.0f) continue;
            const int64_t t = __ldg(tb + static_cast<size_t>(i) * g.W + j);
            if (t == ignore || t < 0 || t >= C) continue;
            float v[C];
            interp_logits<C>(low, g, b, i, j, v);
            float m = v[0];
#pragma unroll
            for (int c = 1; c < C; ++c) m = fmaxf(m, v[c]);
            float S = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) { v[c] = expf(v[c] - m); S += v[c]; }
            const float wgt = wy * wx * sc;
            const float rr = wgt / S;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] += v[c] * rr - ((c == static_cast<int>(t)) ? wgt : 0.0f);
        }
    }
    float *gp = grad_low + static_cast<size_t>(b) * C * hw + r;
#pragma unroll
    for (int c = 0; c < C; ++c) gp[static_cast<size_t>(c) * hw] = acc[c];
}

constexpr int kUpMaxBlocks = 148 * 8;

static UpGeom make_geom(int64_t h, int64_t w, int64_t H, int64_t W)
{
    UpGeom g;
    g.h = static_cast<int>(h); g.w = static_cast<int>(w); g.H = static_cast<int>(H); g.W = static_cast<int>(W);
    g.sy = (H > 1) ? static_cast<float>(h - 1) / static_cast<float>(H - 1) : 0.0f;       // ATen area_pixel_compute_scale, align_corners
    g.sx = (W > 1) ? static_cast<float>(w - 1) / static_cast<float>(W - 1) : 0.0f;
    return g;
}

static bool shape_ok(int64_t B, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W)
{
    return B > 0 && (C == 19 || C == 21) && h > 0 && w > 0 && H >= h && W >= w && B * H * W < (1LL << 31);
}

}  // namespace u2pl

using namespace u2pl;

extern "C" int u2pl_upsample_fused_supported(int64_t C) { return (C == 19 || C == 21) ? 1 : 0; }

extern "C" size_t u2pl_upce_ws_bytes(void) { return static_cast<size_t>(kUpMaxBlocks) * 8; }

extern "C" int u2pl_up_softmax_max(const float *low, int64_t B, int64_t C, int64_t h, int64_t w, int64_t H, int64_t W,
                                   float *out_prob, int64_t *out_label, void *stream)
{
    if (!shape_ok(B, C, h, w, H, W)) return bad_arg("up_softmax_max: need C in {19,21}, H >= h, W >= w, B*H*W < 2^31");
    const UpGeom g = make_geom(h, w, H, W);
    const long long need = (B * H * W + 255) / 256;
    const int grid = static_cast<int>(need < kUpMaxBlocks ? need : kUpMaxBlocks);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (C == 19) up_softmax_max_kernel<19><<<grid, 256, 0, s>>>(low, static_cast<int>(B), g, out_prob, out_label);
    else up_softmax_max_kernel<21><<<grid, 256, 0, s>>>(low, static_cast<int>(B), g, out_prob, out_label);
    return check_launch("up_softmax_max");
}

extern "C" int u2pl_upce_forward(const float *low, const int64_t *target, int64_t B, int64_t C, int64_t h, int64_t w,
                                 int64_t H, int64_t W, int64_t ignore, float *nll_sum, int64_t *n_used,
                                 void *ws, size_t ws_bytes, void *stream)
{
    if (!shape_ok(B, C, h, w, H, W)) return bad_arg("upce_forward: need C in {19,21}, H >= h, W >= w, B*H*W < 2^31");
    if (ws_bytes < static_cast<size_t>(kUpMaxBlocks) * 8) { set_error("upce_forward: workspace too small"); return U2PL_E_WS_SMALL; }
    const UpGeom g = make_geom(h, w, H, W);
    const long long need = (B * H * W + 255) / 256;
    const int grid = static_cast<int>(need < kUpMaxBlocks ? need : kUpMaxBlocks);
    float *ps = static_cast<float *>(ws);
    uint32_t *pc = reinterpret_cast<uint32_t *>(ps + kUpMaxBlocks);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (C == 19) upce_fwd_kernel<19><<<grid, 256, 0, s>>>(low, target, static_cast<int>(B), g, ignore, ps, pc);
    else upce_fwd_kernel<21><<<grid, 256, 0, s>>>(low, target, static_cast<int>(B), g, ignore, ps, pc);
    upce_reduce_kernel<<<1, 256, 0, s>>>(ps, pc, grid, nll_sum, n_used);
    return check_launch("upce_forward", 2);
}

extern "C" int u2pl_upce_backward(const float *low, const int64_t *target, int64_t B, int64_t C, int64_t h, int64_t w,
                                  int64_t H, int64_t W, int64_t ignore, const float *scale, float *grad_low, void *stream)
{
    if (!shape_ok(B, C, h, w, H, W)) return bad_arg("upce_backward: need C in {19,21}, H >= h, W >= w, B*H*W < 2^31");
    const UpGeom g = make_geom(h, w, H, W);
    const long long P = B * h * w;
    const int grid = static_cast<int>((P + 127) / 128);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (C == 19) upce_bwd_kernel<19><<<grid, 128, 0, s>>>(low, target, static_cast<int>(B), g, ignore, scale, grad_low);
    else upce_bwd_kernel<21><<<grid, 128, 0, s>>>(low, target, static_cast<int>(B), g, ignore, scale, grad_low);
    return check_launch("upce_backward");
}
