// bn.cu -- batch normalisation (+ReLU, +residual add) for channels-last bf16 activations.
//
// The reference runs 117 nn.(Sync)BatchNorm2d layers three times per step (student, teacher-train,
// teacher-eval; base.py:6-8, resnet.py:93-140, train_semi.py:318-364) through ATen's generic
// kernels, each followed by a separate ReLU (and a separate residual add).  Profiled on B200 the
// ATen channels-last BN kernels are ~35 % of the whole step (batch_norm_collect_statistics at
// ~0.5 TB/s).  These kernels do the same arithmetic as pure streaming passes:
//
//   stats_partial   x[M,C] bf16 -> per-block per-channel (sum, sum of squares)   1 read
//   reduce          partials -> sums[2][C]               (all-reduced across ranks for SyncBN)
//   finalize        mean / biased var / invstd, running-stat update (unbiased var, momentum),
//                   folded scale = gamma*invstd, shift = beta - mean*scale
//   apply           y = relu(x*scale + shift [+ residual])                        1 read (+1), 1 write
//   bwd_partial     s1 = sum g, s2 = sum g*xhat with g = dy * (y > 0)             3 reads
//   bwd_elemt       dx = gamma*invstd*(g - s1/n - xhat*s2/n), dres = g            3 reads, 1-2 writes
//
// All HBM-bound; every thread moves 16-byte vectors (8 bf16 channels), rows are contiguous so a
// block always touches one contiguous span.  fp32 accumulation per thread, double in the
// cross-block reduction.  C must be a multiple of 8 and <= 2048 (true for every BN in the network).
#include <cuda_bf16.h>
#include "common.cuh"

namespace u2pl {

__device__ __forceinline__ void unpack8(const uint4 &q, float (&v)[8])
{
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h[j]);
        v[2 * j] = f.x;
        v[2 * j + 1] = f.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8])
{
    uint4 q;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    return q;
}

constexpr int kBnThreads = 256;

// block-level reduction of per-thread (a[8], b[8]) over the row slots; result for channel vector
// `cvec` is written by the threads of slot 0.
__device__ __forceinline__ void reduce_slots_store(float (&a)[8], float (&b)[8], int tpr, int C,
                                                   float *__restrict__ out_a, float *__restrict__ out_b)
{
    __shared__ float sm[kBnThreads * 16];
    const int tid = threadIdx.x, slot = tid / tpr, cvec = tid - slot * tpr, slots = kBnThreads / tpr;
#pragma unroll
    for (int j = 0; j < 8; ++j) { sm[tid * 16 + j] = a[j]; sm[tid * 16 + 8 + j] = b[j]; }
    __syncthreads();
    if (slot == 0) {
        for (int s = 1; s < slots; ++s) {
            const float *o = sm + (s * tpr + cvec) * 16;
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] += o[j]; b[j] += o[8 + j]; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { out_a[cvec * 8 + j] = a[j]; out_b[cvec * 8 + j] = b[j]; }
    }
}

__global__ void __launch_bounds__(kBnThreads)
bn_stats_partial_kernel(const uint4 *__restrict__ x, long long M, int C, float *__restrict__ partial)
{
    const int tpr = C >> 3, slots = kBnThreads / tpr;
    const int slot = threadIdx.x / tpr, cvec = threadIdx.x - slot * tpr;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    const long long step = static_cast<long long>(gridDim.x) * slots;
    long long row = static_cast<long long>(blockIdx.x) * slots + slot;
    if (slot < slots) {
        for (; row + 3 * step < M; row += 4 * step) {                 // 4 independent 16 B loads in flight
            uint4 r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) r[u] = __ldg(x + (row + u * step) * tpr + cvec);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float v[8];
                unpack8(r[u], v);
#pragma unroll
                for (int j = 0; j < 8; ++j) { s[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
            }
        }
        for (; row < M; row += step) {
            float v[8];
            unpack8(__ldg(x + row * tpr + cvec), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
        }
    }
    float *out = partial + static_cast<size_t>(blockIdx.x) * 2 * C;
    reduce_slots_store(s, q, tpr, C, out, out + C);
}

// sums[2][C] = sum over blocks of partial[b][2][C]   (double accumulation, fixed order).
// 32 consecutive columns x 8 part-slices per block: every load is a coalesced 128 B row segment and
// the serial chain is nparts/8 long.
__global__ void __launch_bounds__(256)
bn_reduce_kernel(const float *__restrict__ partial, int nparts, int C2, float *__restrict__ sums)
{
    __shared__ double sm[8][32];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31), sl = threadIdx.x >> 5;
    double a = 0.0;
    if (col < C2)
        for (int b = sl; b < nparts; b += 8) a += static_cast<double>(__ldg(partial + static_cast<size_t>(b) * C2 + col));
    sm[sl][threadIdx.x & 31] = a;
    __syncthreads();
    if (sl == 0 && col < C2) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
        sums[col] = static_cast<float>(t);
    }
}

__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float *__restrict__ sums, int C, double count, const float *__restrict__ gamma,
                   const float *__restrict__ beta, float *__restrict__ running_mean, float *__restrict__ running_var,
                   float momentum, float eps, float *__restrict__ mean_out, float *__restrict__ invstd_out,
                   float *__restrict__ scale, float *__restrict__ shift)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    const double mean = static_cast<double>(sums[c]) / count;
    double var = static_cast<double>(sums[C + c]) / count - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    mean_out[c] = static_cast<float>(mean);
    invstd_out[c] = invstd;
    const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
    scale[c] = g * invstd;
    shift[c] = b - static_cast<float>(mean) * g * invstd;
    if (running_mean) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
    }
}

// eval mode: fold the running statistics
__global__ void __launch_bounds__(256)
bn_fold_kernel(int C, const float *__restrict__ gamma, const float *__restrict__ beta,
               const float *__restrict__ running_mean, const float *__restrict__ running_var, float eps,
               float *__restrict__ scale, float *__restrict__ shift)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    const float invstd = rsqrtf(running_var[c] + eps);
    const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
    scale[c] = g * invstd;
    shift[c] = b - running_mean[c] * g * invstd;
}

// Thread (slot, cvec) owns the same 8 channels for every row it visits, so scale / shift sit in registers
// and each 16-byte activation vector costs one load (+1 for the residual) and one store.
__global__ void __launch_bounds__(kBnThreads)
bn_apply_kernel(const uint4 *__restrict__ x, const uint4 *__restrict__ res, const float *__restrict__ scale,
                const float *__restrict__ shift, long long M, int tpr, int relu, uint4 *__restrict__ y)
{
    const int slots = kBnThreads / tpr, slot = threadIdx.x / tpr, cvec = threadIdx.x - slot * tpr;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = __ldg(scale + cvec * 8 + j); sh[j] = __ldg(shift + cvec * 8 + j); }
    const long long step = static_cast<long long>(gridDim.x) * slots;
    long long row = static_cast<long long>(blockIdx.x) * slots + slot;
    for (; row + 3 * step < M; row += 4 * step) {                     // 4 (8 with residual) independent loads in flight
        uint4 xv[4], rv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long i = (row + u * step) * tpr + cvec;
            xv[u] = __ldg(x + i);
            if (res) rv[u] = __ldg(res + i);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float v[8];
            unpack8(xv[u], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
            if (res) {
                float r[8];
                unpack8(rv[u], r);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += r[j];
            }
            if (relu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
            }
            y[(row + u * step) * tpr + cvec] = pack8(v);
        }
    }
    for (; row < M; row += step) {
        const long long i = row * tpr + cvec;
        float v[8];
        unpack8(__ldg(x + i), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
        if (res) {
            float r[8];
            unpack8(__ldg(res + i), r);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += r[j];
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        y[i] = pack8(v);
    }
}

__global__ void __launch_bounds__(kBnThreads)
bn_bwd_partial_kernel(const uint4 *__restrict__ dy, const uint4 *__restrict__ x, const uint4 *__restrict__ y,
                      const float *__restrict__ mean, const float *__restrict__ invstd, long long M, int C,
                      float *__restrict__ partial)
{
    const int tpr = C >> 3, slots = kBnThreads / tpr;
    const int slot = threadIdx.x / tpr, cvec = threadIdx.x - slot * tpr;
    float s1[8], s2[8], mu[8], is[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; mu[j] = mean[cvec * 8 + j]; is[j] = invstd[cvec * 8 + j]; }
    const long long step = static_cast<long long>(gridDim.x) * slots;
    if (slot < slots) {
        long long row = static_cast<long long>(blockIdx.x) * slots + slot;
        for (; row + step < M; row += 2 * step) {                      // two rows (6 independent 16 B loads) in flight
            const long long i0 = row * tpr + cvec, i1 = (row + step) * tpr + cvec;
            const uint4 a0 = __ldg(dy + i0), b0 = __ldg(x + i0), a1 = __ldg(dy + i1), b1 = __ldg(x + i1);
            uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
            if (y) { c0 = __ldg(y + i0); c1 = __ldg(y + i1); }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float g[8], xv[8];
                unpack8(u ? a1 : a0, g);
                unpack8(u ? b1 : b0, xv);
                if (y) {
                    float yv[8];
                    unpack8(u ? c1 : c0, yv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.0f ? g[j] : 0.0f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) { s1[j] += g[j]; s2[j] = fmaf(g[j], (xv[j] - mu[j]) * is[j], s2[j]); }
            }
        }
        for (; row < M; row += step) {
            const long long i = row * tpr + cvec;
            float g[8], xv[8];
            unpack8(__ldg(dy + i), g);
            unpack8(__ldg(x + i), xv);
            if (y) {
                float yv[8];
                unpack8(__ldg(y + i), yv);
#pragma unroll
                for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.0f ? g[j] : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { s1[j] += g[j]; s2[j] = fmaf(g[j], (xv[j] - mu[j]) * is[j], s2[j]); }
        }
    }
    float *out = partial + static_cast<size_t>(blockIdx.x) * 2 * C;
    reduce_slots_store(s1, s2, tpr, C, out, out + C);
}

// dx = gamma*invstd*(g - s1/n - xhat*s2/n) folded into dx = A*g + B*x + D per channel:
//   A = gamma*invstd,  B = -gamma*invstd^2*s2/n,  D = -A*s1/n - B*mean
__global__ void __launch_bounds__(256)
bn_bwd_coef_kernel(const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ gamma,
                   const float *__restrict__ sums, float inv_count, int C, float *__restrict__ coef)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    const float is = invstd[c], ga = gamma ? gamma[c] : 1.0f;
    const float A = ga * is, B = -ga * is * is * sums[C + c] * inv_count;
    coef[c] = A;
    coef[C + c] = B;
    coef[2 * C + c] = -A * sums[c] * inv_count - B * mean[c];
}

__global__ void __launch_bounds__(kBnThreads)
bn_bwd_elemt_kernel(const uint4 *__restrict__ dy, const uint4 *__restrict__ x, const uint4 *__restrict__ y,
                    const float *__restrict__ coef, long long M, int C, uint4 *__restrict__ dx, uint4 *__restrict__ dres)
{
    const int tpr = C >> 3, slots = kBnThreads / tpr, slot = threadIdx.x / tpr, cvec = threadIdx.x - slot * tpr;
    float A[8], B[8], D[8];                                            // per-channel coefficients live in registers
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        A[j] = __ldg(coef + cvec * 8 + j);
        B[j] = __ldg(coef + C + cvec * 8 + j);
        D[j] = __ldg(coef + 2 * C + cvec * 8 + j);
    }
    const long long step = static_cast<long long>(gridDim.x) * slots;
    long long row = static_cast<long long>(blockIdx.x) * slots + slot;
    for (; row + step < M; row += 2 * step) {                          // two rows = 4-6 independent 16 B loads in flight
        const long long i0 = row * tpr + cvec, i1 = (row + step) * tpr + cvec;
        const uint4 a0 = __ldg(dy + i0), b0 = __ldg(x + i0), a1 = __ldg(dy + i1), b1 = __ldg(x + i1);
        uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
        if (y) { c0 = __ldg(y + i0); c1 = __ldg(y + i1); }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float g[8], xv[8], o[8];
            unpack8(u ? a1 : a0, g);
            unpack8(u ? b1 : b0, xv);
            if (y) {
                float yv[8];
                unpack8(u ? c1 : c0, yv);
#pragma unroll
                for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.0f ? g[j] : 0.0f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(A[j], g[j], fmaf(B[j], xv[j], D[j]));
            const long long i = u ? i1 : i0;
            dx[i] = pack8(o);
            if (dres) dres[i] = pack8(g);
        }
    }
    for (; row < M; row += step) {
        const long long i = row * tpr + cvec;
        float g[8], xv[8], o[8];
        unpack8(__ldg(dy + i), g);
        unpack8(__ldg(x + i), xv);
        if (y) {
            float yv[8];
            unpack8(__ldg(y + i), yv);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.0f ? g[j] : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(A[j], g[j], fmaf(B[j], xv[j], D[j]));
        dx[i] = pack8(o);
        if (dres) dres[i] = pack8(g);
    }
}

// eval-mode backward is never needed (the eval teacher runs under no_grad).

static int bn_grid(long long work_items)
{
    const long long need = (work_items + kBnThreads - 1) / kBnThreads;
    const long long cap = 148LL * 8;
    return static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
}

static bool bn_shape_ok(long long M, long long C) { return M > 0 && C >= 8 && C <= 2048 && (C % 8) == 0 && (256 % (C / 8)) == 0; }

}  // namespace u2pl

using namespace u2pl;

constexpr int kBnParts = 148 * 4;        // partial-sum blocks: enough 16 B loads in flight to cover HBM latency

extern "C" int64_t u2pl_bn_parts(void) { return kBnParts; }

int u2pl::bn_reduce_parts(const float *partial, int nparts, int c2, float *sums, void *stream)
{
    bn_reduce_kernel<<<(c2 + 31) / 32, 256, 0, static_cast<cudaStream_t>(stream)>>>(partial, nparts, c2, sums);
    return check_launch("bn_reduce_parts");
}

extern "C" int u2pl_bn_stats(const void *x, int64_t M, int64_t C, float *partial, float *sums, void *stream)
{
    if (!bn_shape_ok(M, C)) return bad_arg("bn_stats: need C % 8 == 0, C/8 a divisor of 256, C <= 2048");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int parts = kBnParts;
    bn_stats_partial_kernel<<<parts, kBnThreads, 0, s>>>(static_cast<const uint4 *>(x), M, static_cast<int>(C), partial);
    bn_reduce_kernel<<<static_cast<int>((2 * C + 31) / 32), 256, 0, s>>>(partial, parts, static_cast<int>(2 * C), sums);
    return check_launch("bn_stats", 2);
}

extern "C" int u2pl_bn_finalize(const float *sums, int64_t C, double count, const float *gamma, const float *beta,
                                float *running_mean, float *running_var, float momentum, float eps,
                                float *mean, float *invstd, float *scale, float *shift, void *stream)
{
    bn_finalize_kernel<<<static_cast<int>((C + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        sums, static_cast<int>(C), count, gamma, beta, running_mean, running_var, momentum, eps, mean, invstd, scale, shift);
    return check_launch("bn_finalize");
}

extern "C" int u2pl_bn_fold(int64_t C, const float *gamma, const float *beta, const float *running_mean,
                            const float *running_var, float eps, float *scale, float *shift, void *stream)
{
    bn_fold_kernel<<<static_cast<int>((C + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<int>(C), gamma, beta, running_mean, running_var, eps, scale, shift);
    return check_launch("bn_fold");
}

extern "C" int u2pl_bn_apply(const void *x, const void *residual, const float *scale, const float *shift,
                             int64_t M, int64_t C, int relu, void *y, void *stream)
{
    if (!bn_shape_ok(M, C)) return bad_arg("bn_apply: unsupported channel count");
    const long long nvec = M * (C / 8);
    bn_apply_kernel<<<bn_grid(nvec / 2), kBnThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4 *>(x), static_cast<const uint4 *>(residual), scale, shift, M, static_cast<int>(C / 8), relu,
        static_cast<uint4 *>(y));
    return check_launch("bn_apply");
}

extern "C" int u2pl_bn_backward_reduce(const void *dy, const void *x, const void *y, const float *mean, const float *invstd,
                                       int64_t M, int64_t C, float *partial, float *sums, void *stream)
{
    if (!bn_shape_ok(M, C)) return bad_arg("bn_backward_reduce: unsupported channel count");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int parts = kBnParts;
    bn_bwd_partial_kernel<<<parts, kBnThreads, 0, s>>>(static_cast<const uint4 *>(dy), static_cast<const uint4 *>(x),
                                                        static_cast<const uint4 *>(y), mean, invstd, M, static_cast<int>(C), partial);
    bn_reduce_kernel<<<static_cast<int>((2 * C + 31) / 32), 256, 0, s>>>(partial, parts, static_cast<int>(2 * C), sums);
    return check_launch("bn_backward_reduce", 2);
}

extern "C" int u2pl_bn_backward_elemt(const void *dy, const void *x, const void *y, const float *mean, const float *invstd,
                                      const float *gamma, const float *sums, double count, int64_t M, int64_t C,
                                      float *coef, void *dx, void *dres, void *stream)
{
    if (!bn_shape_ok(M, C)) return bad_arg("bn_backward_elemt: unsupported channel count");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long nvec = M * (C / 8);
    bn_bwd_coef_kernel<<<static_cast<int>((C + 255) / 256), 256, 0, s>>>(mean, invstd, gamma, sums, static_cast<float>(1.0 / count),
                                                                       static_cast<int>(C), coef);
    bn_bwd_elemt_kernel<<<bn_grid(nvec / 2), kBnThreads, 0, s>>>(static_cast<const uint4 *>(dy), static_cast<const uint4 *>(x),
                                                                 static_cast<const uint4 *>(y), coef, M, static_cast<int>(C),
                                                             static_cast<uint4 *>(dx), static_cast<uint4 *>(dres));
    return check_launch("bn_backward_elemt", 2);
}
