// pool.cu -- 3x3 / stride 2 / padding 1 / ceil_mode max-pooling of a channels-last bf16 tensor, forward and backward.
//
// replaces: nn.MaxPool2d(kernel_size=3, stride=2, padding=1, ceil_mode=True) of the ResNet stem (resnet.py:185,282) on the
// three network passes of a step.  ATen's channels-last kernels move this 257x257x128 tensor at ~0.6 TB/s (2.7 + 2.1 ms per
// step, profiles/r01_step_profile_torchprof_v4.txt); it is a pure streaming op: 16-byte vectors over the channel axis,
// one thread per (output pixel, 8 channels), the winning tap stored as one byte per element so that the backward pass is
// a gather (each input pixel looks at the <= 4 windows that contain it) without atomics.
// Tie rule = ATen's: windows are scanned row-major and a later tap wins only if it is strictly greater (or NaN).
#include <algorithm>
#include <cuda_bf16.h>
#include "common.cuh"

namespace u2pl {

__device__ __forceinline__ void unpack8_bf16(const uint4 &v, float (&f)[8])
{
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 t = __bfloat1622float2(h[e]);
        f[2 * e] = t.x;
        f[2 * e + 1] = t.y;
    }
}

__global__ void __launch_bounds__(256)
maxpool3s2_fwd_kernel(const uint4 *__restrict__ x, int N, int H, int W, int C8, int Ho, int Wo,
                      uint4 *__restrict__ y, uint2 *__restrict__ tap)
{
    const long long total = static_cast<long long>(N) * Ho * Wo * C8;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
        const int c = static_cast<int>(i % C8);
        long long p = i / C8;
        const int wo = static_cast<int>(p % Wo); p /= Wo;
        const int ho = static_cast<int>(p % Ho);
        const int n = static_cast<int>(p / Ho);
        float best[8];
        uint32_t idx[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; idx[e] = 0xffu; }
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int h = 2 * ho - 1 + kh;
            if (h < 0 || h >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int w = 2 * wo - 1 + kw;
                if (w < 0 || w >= W) continue;
                float v[8];
                unpack8_bf16(__ldg(x + ((static_cast<long long>(n) * H + h) * W + w) * C8 + c), v);
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (idx[e] == 0xffu || v[e] > best[e] || v[e] != v[e]) { best[e] = v[e]; idx[e] = kh * 3 + kw; }
            }
        }
        uint4 o;
        __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) oh[e] = __floats2bfloat162_rn(best[2 * e], best[2 * e + 1]);
        y[i] = o;
        if (tap) {
            uint2 t;
            t.x = idx[0] | (idx[1] << 8) | (idx[2] << 16) | (idx[3] << 24);
            t.y = idx[4] | (idx[5] << 8) | (idx[6] << 16) | (idx[7] << 24);
            tap[i] = t;
        }
    }
}

__global__ void __launch_bounds__(256)
maxpool3s2_bwd_kernel(const uint4 *__restrict__ dy, const uint2 *__restrict__ tap, int N, int H, int W, int C8, int Ho, int Wo,
                      uint4 *__restrict__ dx)
{
    const long long total = static_cast<long long>(N) * H * W * C8;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
        const int c = static_cast<int>(i % C8);
        long long p = i / C8;
        const int w = static_cast<int>(p % W); p /= W;
        const int h = static_cast<int>(p % H);
        const int n = static_cast<int>(p / H);
        float g[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = 0.0f;
        // windows (ho, wo) with 2*ho - 1 <= h <= 2*ho + 1: ho in {h/2, (h+1)/2} (equal when h is even)
        const int ho0 = h >> 1, ho1 = (h + 1) >> 1, wo0 = w >> 1, wo1 = (w + 1) >> 1;
        for (int ho = ho0; ho <= ho1; ++ho) {
            if (ho >= Ho) continue;
            const int kh = h - (2 * ho - 1);
            for (int wo = wo0; wo <= wo1; ++wo) {
                if (wo >= Wo) continue;
                const int kw = w - (2 * wo - 1);
                const uint32_t mine = static_cast<uint32_t>(kh * 3 + kw);
                const long long o = ((static_cast<long long>(n) * Ho + ho) * Wo + wo) * C8 + c;
                const uint2 t = __ldg(tap + o);
                float d[8];
                unpack8_bf16(__ldg(dy + o), d);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const uint32_t te = ((e < 4 ? t.x : t.y) >> (8 * (e & 3))) & 0xffu;
                    if (te == mine) g[e] += d[e];
                }
            }
        }
        uint4 o;
        __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) oh[e] = __floats2bfloat162_rn(g[2 * e], g[2 * e + 1]);
        dx[i] = o;
    }
}

static inline int pooled(int64_t n) { return static_cast<int>((n + 2 - 3 + 1) / 2 + 1 - (((n + 2 - 3 + 1) / 2) * 2 >= n + 1 ? 1 : 0)); }

}  // namespace u2pl

using namespace u2pl;

extern "C" int64_t u2pl_maxpool3s2_out(int64_t n)      // ceil((n + 2*1 - 3) / 2) + 1, last window must start inside the input
{
    return pooled(n);
}

extern "C" int u2pl_maxpool3s2_forward(const void *x, void *y, void *tap, int64_t n, int64_t h, int64_t w, int64_t c, void *stream)
{
    if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c % 8)) return bad_arg("maxpool3s2_forward: need C % 8 == 0");
    const int ho = pooled(h), wo = pooled(w);
    const long long total = n * ho * wo * (c / 8);
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(kNumSMs) * 16));
    maxpool3s2_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4 *>(x), static_cast<int>(n), static_cast<int>(h), static_cast<int>(w), static_cast<int>(c / 8), ho, wo,
        static_cast<uint4 *>(y), static_cast<uint2 *>(tap));
    return check_launch("maxpool3s2_forward");
}

extern "C" int u2pl_maxpool3s2_backward(const void *dy, const void *tap, void *dx, int64_t n, int64_t h, int64_t w, int64_t c, void *stream)
{
    if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c % 8) || !tap) return bad_arg("maxpool3s2_backward: need C % 8 == 0 and the tap map");
    const int ho = pooled(h), wo = pooled(w);
    const long long total = n * h * w * (c / 8);
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(kNumSMs) * 16));
    maxpool3s2_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4 *>(dy), static_cast<const uint2 *>(tap), static_cast<int>(n), static_cast<int>(h), static_cast<int>(w),
        static_cast<int>(c / 8), ho, wo, static_cast<uint4 *>(dx));
    return check_launch("maxpool3s2_backward");
}
