// tc_selftest.cu -- torch-free check of the tensor-core entry points of libu2pl_b200.so through the C ABI:
// u2pl_conv_bf16_nhwc (plain / epilogue / statistics) and u2pl_conv_wgrad_bf16_nhwc against CPU loops.
// Starts in about a second (no Python, no torch import), so it fits in the smallest GPU slot:
//   nvcc -O2 -std=c++17 -o tools/cu/tc_selftest.bin tools/cu/tc_selftest.cu -ldl     (built here, runs on the box)
//   ./tools/cu/tc_selftest.bin [conv|stats|xform|pool|wgrad|all|perf]      (perf: CUDA-event timings at the network's layer shapes)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <vector>
#include <cuda_runtime.h>

typedef int (*conv_fn)(const void *, const void *, void *, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, const float *,
                       const float *, const void *, int, void *);
typedef int (*conv_stats_fn)(const void *, const void *, void *, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, float *, float *, void *);
typedef int (*conv_ex_fn)(const void *, const void *, void *, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, const float *,
                          const float *, int, const float *, const float *, const void *, int, float *, float *, void *);
typedef int64_t (*parts_fn)(int64_t, int64_t, int64_t, int);
typedef int (*splits_fn)(int64_t, int64_t, int64_t, int64_t, int64_t);
typedef int (*wgrad_fn)(const void *, const void *, float *, int64_t, int64_t, int64_t, int64_t, int64_t, int, void *);
typedef int64_t (*pool_out_fn)(int64_t);
typedef int (*pool_fn)(const void *, void *, void *, int64_t, int64_t, int64_t, int64_t, void *);
typedef int (*pool_bwd_fn)(const void *, const void *, void *, int64_t, int64_t, int64_t, int64_t, void *);
typedef const char *(*err_fn)(void);

static uint16_t f2bf(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7fffu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
}
static float bf2f(uint16_t h)
{
    uint32_t u = static_cast<uint32_t>(h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static uint32_t g_seed = 777u;
static float rnd() { g_seed = g_seed * 1664525u + 1013904223u; return (static_cast<int>((g_seed >> 20) % 2001) - 1000) / 1000.0f; }

template <typename T> static T *to_dev(const std::vector<T> &h)
{
    T *d = nullptr;
    cudaMalloc(&d, h.size() * sizeof(T));
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

struct Lib { conv_fn conv; conv_ex_fn conv_ex; conv_stats_fn conv_stats; parts_fn parts; splits_fn splits; wgrad_fn wgrad; err_fn err; pool_out_fn pool_out; pool_fn pool; pool_bwd_fn pool_bwd; };

static int check_conv(const Lib &L, int N, int Cin, int H, int W, int Cout, int k, int d, bool epi, bool stats)
{
    const size_t nx = static_cast<size_t>(N) * H * W * Cin, nw = static_cast<size_t>(Cout) * k * k * Cin, ny = static_cast<size_t>(N) * H * W * Cout;
    std::vector<uint16_t> x(nx), w(nw), res(ny);
    std::vector<float> scale(Cout), shift(Cout);
    for (auto &v : x) v = f2bf(rnd());
    for (auto &v : w) v = f2bf(rnd() / sqrtf(static_cast<float>(Cin * k * k)) * 3.0f);
    for (auto &v : res) v = f2bf(rnd());
    for (int c = 0; c < Cout; ++c) { scale[c] = 1.0f + 0.5f * rnd(); shift[c] = rnd(); }
    std::vector<float> ref(ny);
    const int pad = d * (k / 2);
    for (int n = 0; n < N; ++n)
        for (int h = 0; h < H; ++h)
            for (int ww = 0; ww < W; ++ww)
                for (int co = 0; co < Cout; ++co) {
                    float acc = 0.0f;
                    for (int r = 0; r < k; ++r)
                        for (int s = 0; s < k; ++s) {
                            const int hi = h + r * d - pad, wi = ww + s * d - pad;
                            if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                            const uint16_t *xp = &x[((static_cast<size_t>(n) * H + hi) * W + wi) * Cin];
                            const uint16_t *wp = &w[((static_cast<size_t>(co) * k + r) * k + s) * Cin];
                            for (int ci = 0; ci < Cin; ++ci) acc += bf2f(xp[ci]) * bf2f(wp[ci]);
                        }
                    const size_t o = ((static_cast<size_t>(n) * H + h) * W + ww) * Cout + co;
                    if (epi) { acc = acc * scale[co] + shift[co] + bf2f(res[o]); acc = acc > 0.0f ? acc : 0.0f; }
                    ref[o] = acc;
                }
    uint16_t *dx = to_dev(x), *dw = to_dev(w), *dres = to_dev(res), *dy = nullptr;
    float *dscale = to_dev(scale), *dshift = to_dev(shift);
    cudaMalloc(&dy, ny * 2);
    cudaMemset(dy, 0xff, ny * 2);
    int rc;
    std::vector<float> sums(2 * Cout, 0.0f);
    if (stats) {
        const int64_t parts = L.parts(N, H, W, k);
        float *dpart = nullptr, *dsums = nullptr;
        cudaMalloc(&dpart, parts * 2 * Cout * 4);
        cudaMalloc(&dsums, 2 * Cout * 4);
        rc = L.conv_stats(dx, dw, dy, N, H, W, Cin, Cout, k, d, dpart, dsums, nullptr);
        cudaDeviceSynchronize();
        cudaMemcpy(sums.data(), dsums, 2 * Cout * 4, cudaMemcpyDeviceToHost);
    } else {
        rc = L.conv(dx, dw, dy, N, H, W, Cin, Cout, k, d, epi ? dscale : nullptr, epi ? dshift : nullptr, epi ? dres : nullptr, epi ? 1 : 0, nullptr);
    }
    const cudaError_t e = cudaDeviceSynchronize();
    if (rc != 0 || e != cudaSuccess) {
        printf("conv N=%d Cin=%d %dx%d Cout=%d k=%d d=%d epi=%d stats=%d: rc=%d cuda=%s err=%s\n", N, Cin, H, W, Cout, k, d, epi, stats, rc,
               cudaGetErrorString(e), L.err());
        return 1;
    }
    std::vector<uint16_t> y(ny);
    cudaMemcpy(y.data(), dy, ny * 2, cudaMemcpyDeviceToHost);
    float worst = 0.0f, scale_ref = 1.0f;
    for (size_t i = 0; i < ny; ++i) scale_ref = fmaxf(scale_ref, fabsf(ref[i]));
    size_t bad = 0;
    for (size_t i = 0; i < ny; ++i) {
        const float dlt = fabsf(bf2f(y[i]) - ref[i]);
        if (!(dlt <= 1e-2f * scale_ref)) ++bad;
        if (dlt > worst || dlt != dlt) worst = dlt;
    }
    double s_err = 0.0;
    if (stats) {
        for (int c = 0; c < Cout; ++c) {
            double a = 0.0, b = 0.0;
            for (size_t p = 0; p < static_cast<size_t>(N) * H * W; ++p) { const double v = bf2f(y[p * Cout + c]); a += v; b += v * v; }
            s_err = fmax(s_err, fabs(a - sums[c]) / (1.0 + fabs(a)));
            s_err = fmax(s_err, fabs(b - sums[Cout + c]) / (1.0 + fabs(b)));
        }
        if (s_err > 1e-3) ++bad;
    }
    printf("conv N=%d Cin=%d %dx%d Cout=%d k=%d d=%d epi=%d stats=%d: mismatches=%zu/%zu max_err=%g (ref max %g) stats_rel_err=%g  %s\n", N, Cin,
           H, W, Cout, k, d, epi, stats, bad, ny, worst, scale_ref, s_err, bad == 0 ? "OK" : "WRONG");
    return bad == 0 ? 0 : 1;
}

// in-transform variant: relu(x * s + t) applied to the input in shared memory, plain output, optional statistics
static int check_conv_xform(const Lib &L, int N, int Cin, int H, int W, int Cout, int k, int d, bool stats)
{
    const size_t nx = static_cast<size_t>(N) * H * W * Cin, nw = static_cast<size_t>(Cout) * k * k * Cin, ny = static_cast<size_t>(N) * H * W * Cout;
    std::vector<uint16_t> x(nx), w(nw);
    std::vector<float> is(Cin), it(Cin);
    for (auto &v : x) v = f2bf(rnd());
    for (auto &v : w) v = f2bf(rnd() / sqrtf(static_cast<float>(Cin * k * k)) * 3.0f);
    for (int c = 0; c < Cin; ++c) { is[c] = 1.0f + 0.5f * rnd(); it[c] = 0.5f * rnd(); }
    std::vector<float> z(nx), ref(ny);
    for (size_t i = 0; i < nx; ++i) {                              // activated input, rounded to bf16 like the kernel's rewrite
        float v = bf2f(x[i]) * is[i % Cin] + it[i % Cin];
        z[i] = bf2f(f2bf(v > 0.0f ? v : 0.0f));
    }
    const int pad = d * (k / 2);
    for (int n = 0; n < N; ++n)
        for (int h = 0; h < H; ++h)
            for (int ww = 0; ww < W; ++ww)
                for (int co = 0; co < Cout; ++co) {
                    float acc = 0.0f;
                    for (int r = 0; r < k; ++r)
                        for (int s = 0; s < k; ++s) {
                            const int hi = h + r * d - pad, wi = ww + s * d - pad;
                            if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                            const float *zp = &z[((static_cast<size_t>(n) * H + hi) * W + wi) * Cin];
                            const uint16_t *wp = &w[((static_cast<size_t>(co) * k + r) * k + s) * Cin];
                            for (int ci = 0; ci < Cin; ++ci) acc += zp[ci] * bf2f(wp[ci]);
                        }
                    ref[((static_cast<size_t>(n) * H + h) * W + ww) * Cout + co] = acc;
                }
    uint16_t *dx = to_dev(x), *dw = to_dev(w), *dy = nullptr;
    float *dis = to_dev(is), *dit = to_dev(it), *dpart = nullptr, *dsums = nullptr;
    cudaMalloc(&dy, ny * 2);
    cudaMemset(dy, 0xff, ny * 2);
    if (stats) { cudaMalloc(&dpart, L.parts(N, H, W, k) * 2 * Cout * 4); cudaMalloc(&dsums, 2 * Cout * 4); }
    const int rc = L.conv_ex(dx, dw, dy, N, H, W, Cin, Cout, k, d, dis, dit, 1, nullptr, nullptr, nullptr, 0, dpart, dsums, nullptr);
    const cudaError_t e = cudaDeviceSynchronize();
    if (rc != 0 || e != cudaSuccess) { printf("conv_xform: rc=%d cuda=%s err=%s\n", rc, cudaGetErrorString(e), L.err()); return 1; }
    std::vector<uint16_t> y(ny);
    cudaMemcpy(y.data(), dy, ny * 2, cudaMemcpyDeviceToHost);
    float worst = 0.0f, mx = 1.0f;
    for (size_t i = 0; i < ny; ++i) mx = fmaxf(mx, fabsf(ref[i]));
    size_t bad = 0;
    for (size_t i = 0; i < ny; ++i) {
        const float dlt = fabsf(bf2f(y[i]) - ref[i]);
        if (!(dlt <= 1e-2f * mx)) ++bad;
        if (dlt > worst || dlt != dlt) worst = dlt;
    }
    double s_err = 0.0;
    if (stats) {
        std::vector<float> sums(2 * Cout);
        cudaMemcpy(sums.data(), dsums, 2 * Cout * 4, cudaMemcpyDeviceToHost);
        for (int c = 0; c < Cout; ++c) {
            double a = 0.0, b = 0.0;
            for (size_t p = 0; p < static_cast<size_t>(N) * H * W; ++p) { const double v = bf2f(y[p * Cout + c]); a += v; b += v * v; }
            s_err = fmax(s_err, fmax(fabs(a - sums[c]) / (1.0 + fabs(a)), fabs(b - sums[Cout + c]) / (1.0 + fabs(b))));
        }
        if (s_err > 1e-3) ++bad;
    }
    printf("conv_xform N=%d Cin=%d %dx%d Cout=%d k=%d d=%d stats=%d: mismatches=%zu/%zu max_err=%g (ref max %g) stats_rel_err=%g  %s\n", N, Cin, H, W,
           Cout, k, d, stats, bad, ny, worst, mx, s_err, bad == 0 ? "OK" : "WRONG");
    return bad == 0 ? 0 : 1;
}

static int check_wgrad(const Lib &L, int N, int Cin, int H, int W, int Cout, int d)
{
    const size_t nx = static_cast<size_t>(N) * H * W * Cin, ng = static_cast<size_t>(N) * H * W * Cout, nd = static_cast<size_t>(9) * Cout * Cin;
    std::vector<uint16_t> x(nx), g(ng);
    for (auto &v : x) v = f2bf(rnd());
    for (auto &v : g) v = f2bf(rnd());
    std::vector<float> ref(nd, 0.0f);
    for (int tap = 0; tap < 9; ++tap) {
        const int dh = (tap / 3 - 1) * d, dw = (tap % 3 - 1) * d;
        for (int n = 0; n < N; ++n)
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w) {
                    const int hi = h + dh, wi = w + dw;
                    if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                    const uint16_t *gp = &g[((static_cast<size_t>(n) * H + h) * W + w) * Cout];
                    const uint16_t *xp = &x[((static_cast<size_t>(n) * H + hi) * W + wi) * Cin];
                    for (int co = 0; co < Cout; ++co) {
                        const float gv = bf2f(gp[co]);
                        float *row = &ref[(static_cast<size_t>(tap) * Cout + co) * Cin];
                        for (int ci = 0; ci < Cin; ++ci) row[ci] += gv * bf2f(xp[ci]);
                    }
                }
    }
    const int splits = L.splits(N, H, W, Cin, Cout);
    uint16_t *dx = to_dev(x), *dg = to_dev(g);
    float *dpart = nullptr;
    cudaMalloc(&dpart, static_cast<size_t>(splits) * nd * 4);
    cudaMemset(dpart, 0xff, static_cast<size_t>(splits) * nd * 4);
    const int rc = L.wgrad(dx, dg, dpart, N, H, W, Cin, Cout, d, nullptr);
    const cudaError_t e = cudaDeviceSynchronize();
    if (rc != 0 || e != cudaSuccess) { printf("wgrad: rc=%d cuda=%s err=%s\n", rc, cudaGetErrorString(e), L.err()); return 1; }
    std::vector<float> part(static_cast<size_t>(splits) * nd);
    cudaMemcpy(part.data(), dpart, part.size() * 4, cudaMemcpyDeviceToHost);
    float worst = 0.0f, mx = 1.0f;
    size_t bad = 0;
    for (size_t i = 0; i < nd; ++i) {
        float a = 0.0f;
        for (int s = 0; s < splits; ++s) a += part[s * nd + i];
        const float dlt = fabsf(a - ref[i]);
        mx = fmaxf(mx, fabsf(ref[i]));
        if (!(dlt <= 2e-3f * fmaxf(1.0f, fabsf(ref[i])) + 1e-2f)) ++bad;
        if (dlt > worst || dlt != dlt) worst = dlt;
    }
    printf("wgrad N=%d Cin=%d %dx%d Cout=%d d=%d splits=%d: mismatches=%zu/%zu max_err=%g (ref max %g)  %s\n", N, Cin, H, W, Cout, d, splits, bad,
           nd, worst, mx, bad == 0 ? "OK" : "WRONG");
    return bad == 0 ? 0 : 1;
}

// 3x3 / stride 2 / pad 1 / ceil_mode max-pooling: forward values + taps, backward gather, against CPU loops (ATen's tie rule)
static int check_pool(const Lib &L, int N, int H, int W, int C)
{
    const int Ho = static_cast<int>(L.pool_out(H)), Wo = static_cast<int>(L.pool_out(W));
    const size_t nx = static_cast<size_t>(N) * H * W * C, ny = static_cast<size_t>(N) * Ho * Wo * C;
    std::vector<uint16_t> x(nx), g(ny);
    for (auto &v : x) { const float r = rnd(); v = f2bf(r > 0.0f ? roundf(r * 8.0f) / 8.0f : 0.0f); }    // ReLU-like: zeros and ties
    for (auto &v : g) v = f2bf(rnd());
    std::vector<float> yref(ny), dxref(nx, 0.0f);
    std::vector<uint8_t> tref(ny);
    for (int n = 0; n < N; ++n)
        for (int ho = 0; ho < Ho; ++ho)
            for (int wo = 0; wo < Wo; ++wo)
                for (int c = 0; c < C; ++c) {
                    float best = 0.0f;
                    int tap = -1, bh = 0, bw = 0;
                    for (int kh = 0; kh < 3; ++kh)
                        for (int kw = 0; kw < 3; ++kw) {
                            const int h = 2 * ho - 1 + kh, w = 2 * wo - 1 + kw;
                            if (h < 0 || h >= H || w < 0 || w >= W) continue;
                            const float v = bf2f(x[((static_cast<size_t>(n) * H + h) * W + w) * C + c]);
                            if (tap < 0 || v > best) { best = v; tap = kh * 3 + kw; bh = h; bw = w; }
                        }
                    const size_t o = ((static_cast<size_t>(n) * Ho + ho) * Wo + wo) * C + c;
                    yref[o] = best; tref[o] = static_cast<uint8_t>(tap);
                    dxref[((static_cast<size_t>(n) * H + bh) * W + bw) * C + c] += bf2f(g[o]);
                }
    uint16_t *dx_in = to_dev(x), *dg = to_dev(g), *dy = nullptr, *ddx = nullptr;
    uint8_t *dtap = nullptr;
    cudaMalloc(&dy, ny * 2); cudaMalloc(&ddx, nx * 2); cudaMalloc(&dtap, ny);
    int rc = L.pool(dx_in, dy, dtap, N, H, W, C, nullptr);
    if (rc == 0) rc = L.pool_bwd(dg, dtap, ddx, N, H, W, C, nullptr);
    const cudaError_t e = cudaDeviceSynchronize();
    if (rc != 0 || e != cudaSuccess) { printf("pool: rc=%d cuda=%s err=%s\n", rc, cudaGetErrorString(e), L.err()); return 1; }
    std::vector<uint16_t> y(ny), dxo(nx);
    std::vector<uint8_t> tap(ny);
    cudaMemcpy(y.data(), dy, ny * 2, cudaMemcpyDeviceToHost);
    cudaMemcpy(tap.data(), dtap, ny, cudaMemcpyDeviceToHost);
    cudaMemcpy(dxo.data(), ddx, nx * 2, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (size_t i = 0; i < ny; ++i) if (bf2f(y[i]) != yref[i] || tap[i] != tref[i]) ++bad;
    for (size_t i = 0; i < nx; ++i) if (!(fabsf(bf2f(dxo[i]) - dxref[i]) <= 2e-2f * fmaxf(1.0f, fabsf(dxref[i])))) ++bad;
    printf("pool N=%d %dx%d -> %dx%d C=%d: mismatches=%zu  %s\n", N, H, W, Ho, Wo, C, bad, bad == 0 ? "OK" : "WRONG");
    return bad == 0 ? 0 : 1;
}

// ---- timings at the network's layer shapes (T1 batch: 16 images), no CPU reference: TFLOP/s per kernel
static uint16_t *dev_pattern(size_t n)
{
    std::vector<uint16_t> h(1 << 20);
    for (auto &v : h) v = f2bf(rnd());
    uint16_t *d = nullptr;
    cudaMalloc(&d, n * 2);
    for (size_t o = 0; o < n; o += h.size())
        cudaMemcpy(d + o, h.data(), (n - o < h.size() ? n - o : h.size()) * 2, cudaMemcpyHostToDevice);
    return d;
}

static void perf(const Lib &L, int only)
{
    if (only < 0) {   // stem max-pooling at the student batch: 32 x 257 x 257 x 128
        const int N = 32, H = 257, W = 257, C = 128, Ho = static_cast<int>(L.pool_out(H)), Wo = static_cast<int>(L.pool_out(W));
        const size_t nx = static_cast<size_t>(N) * H * W * C, ny = static_cast<size_t>(N) * Ho * Wo * C;
        uint16_t *x = dev_pattern(nx), *y = dev_pattern(ny), *dx = nullptr;
        uint8_t *tap = nullptr;
        cudaMalloc(&dx, nx * 2); cudaMalloc(&tap, ny);
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        float ms[2];
        for (int pass = 0; pass < 2; ++pass) {
            for (int i = 0; i < 13; ++i) {
                if (i == 3) cudaEventRecord(a);
                if (pass == 0) L.pool(x, y, tap, N, H, W, C, nullptr); else L.pool_bwd(y, tap, dx, N, H, W, C, nullptr);
            }
            cudaEventRecord(b); cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms[pass], a, b);
            ms[pass] /= 10.0f;
        }
        printf("perf maxpool 32x257x257x128: fwd %.3f ms (%.0f GB/s), bwd %.3f ms (%.0f GB/s)\n", ms[0], (nx * 2 + ny * 3) / ms[0] / 1e6, ms[1],
               (nx * 2 + ny * 3) / ms[1] / 1e6);
        cudaFree(x); cudaFree(y); cudaFree(dx); cudaFree(tap);
    }
    struct Shape { const char *name; int N, Cin, H, W, Cout, k, d; };
    const Shape shapes[] = {{"layer3.conv1 1x1 1024->256", 16, 1024, 65, 65, 256, 1, 1}, {"layer3.conv2 3x3 d2 256->256", 16, 256, 65, 65, 256, 3, 2},
                            {"layer3.conv3 1x1 256->1024", 16, 256, 65, 65, 1024, 1, 1}, {"layer4.conv2 3x3 d8 512->512", 16, 512, 65, 65, 512, 3, 8},
                            {"aspp 3x3 d12 2048->256", 16, 2048, 65, 65, 256, 3, 12}, {"head 3x3 1280->256", 16, 1280, 65, 65, 256, 3, 1},
                            {"decoder 3x3 256->256 @129", 16, 256, 129, 129, 256, 3, 1}, {"8192^3 as 1x1", 1, 8192, 8192, 1, 8192, 1, 1}};
    int shape_i = -1;
    for (const Shape &sh : shapes) {
        if (++shape_i != only && only >= 0) continue;
        const size_t nx = static_cast<size_t>(sh.N) * sh.H * sh.W * sh.Cin, ny = static_cast<size_t>(sh.N) * sh.H * sh.W * sh.Cout;
        const size_t nw = static_cast<size_t>(sh.Cout) * sh.k * sh.k * sh.Cin;
        uint16_t *x = dev_pattern(nx), *w = dev_pattern(nw), *y = dev_pattern(ny);
        std::vector<float> ones(sh.Cout, 1.0f);
        float *sc = to_dev(ones), *sf = to_dev(ones);
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        auto time_it = [&](auto fn) {
            for (int i = 0; i < 3; ++i) fn();
            cudaEventRecord(a);
            for (int i = 0; i < 10; ++i) fn();
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, a, b);
            return ms / 10.0f;
        };
        const double flop = 2.0 * sh.N * sh.H * sh.W * static_cast<double>(sh.Cout) * sh.Cin * sh.k * sh.k;
        const float t_conv = time_it([&] { L.conv(x, w, y, sh.N, sh.H, sh.W, sh.Cin, sh.Cout, sh.k, sh.d, sc, sf, nullptr, 1, nullptr); });
        printf("perf %-30s conv+bn+relu %8.3f ms %7.1f TFLOP/s", sh.name, t_conv, flop / t_conv / 1e9);
        if (sh.k == 1 && sh.N > 1) {                   // bottleneck exit: + residual (read) in the epilogue
            uint16_t *res = dev_pattern(ny);
            const float t_res = time_it([&] { L.conv(x, w, y, sh.N, sh.H, sh.W, sh.Cin, sh.Cout, sh.k, sh.d, sc, sf, res, 1, nullptr); });
            printf(" | +residual %8.3f ms (%.0f GB/s of x+res+y)", t_res, (nx + 2 * ny) * 2 / t_res / 1e6);
            cudaFree(res);
        }
        if (sh.k == 3) {
            const int64_t parts = L.parts(sh.N, sh.H, sh.W, sh.k);
            float *part = nullptr, *sums = nullptr, *wpart = nullptr;
            cudaMalloc(&part, parts * 2 * sh.Cout * 4); cudaMalloc(&sums, 2 * sh.Cout * 4);
            const float t_st = time_it([&] { L.conv_stats(x, w, y, sh.N, sh.H, sh.W, sh.Cin, sh.Cout, sh.k, sh.d, part, sums, nullptr); });
            const int splits = L.splits(sh.N, sh.H, sh.W, sh.Cin, sh.Cout);
            cudaMalloc(&wpart, static_cast<size_t>(splits) * 9 * sh.Cout * sh.Cin * 4);
            const float t_wg = time_it([&] { L.wgrad(x, y, wpart, sh.N, sh.H, sh.W, sh.Cin, sh.Cout, sh.d, nullptr); });
            printf(" | conv+stats %8.3f ms | wgrad (splits %d) %8.3f ms %7.1f TFLOP/s", t_st, splits, t_wg, flop / t_wg / 1e9);
            cudaFree(part); cudaFree(sums); cudaFree(wpart);
        }
        const cudaError_t e = cudaDeviceSynchronize();
        printf("%s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
        cudaFree(x); cudaFree(w); cudaFree(y); cudaFree(sc); cudaFree(sf);
    }
}

int main(int argc, char **argv)
{
    const char *what = argc > 1 ? argv[1] : "all";
    void *h = dlopen("u2pl_b200/libu2pl_b200.so", RTLD_NOW);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    Lib L;
    L.conv = reinterpret_cast<conv_fn>(dlsym(h, "u2pl_conv_bf16_nhwc"));
    L.conv_ex = reinterpret_cast<conv_ex_fn>(dlsym(h, "u2pl_conv_bf16_nhwc_ex"));
    L.conv_stats = reinterpret_cast<conv_stats_fn>(dlsym(h, "u2pl_conv_bf16_nhwc_stats"));
    L.parts = reinterpret_cast<parts_fn>(dlsym(h, "u2pl_conv_stat_parts"));
    L.splits = reinterpret_cast<splits_fn>(dlsym(h, "u2pl_conv_wgrad_splits"));
    L.wgrad = reinterpret_cast<wgrad_fn>(dlsym(h, "u2pl_conv_wgrad_bf16_nhwc"));
    L.err = reinterpret_cast<err_fn>(dlsym(h, "u2pl_last_error"));
    L.pool_out = reinterpret_cast<pool_out_fn>(dlsym(h, "u2pl_maxpool3s2_out"));
    L.pool = reinterpret_cast<pool_fn>(dlsym(h, "u2pl_maxpool3s2_forward"));
    L.pool_bwd = reinterpret_cast<pool_bwd_fn>(dlsym(h, "u2pl_maxpool3s2_backward"));
    if (!L.pool_out || !L.pool || !L.pool_bwd) { fprintf(stderr, "missing pool symbol\n"); return 2; }
    if (!L.conv || !L.conv_ex || !L.conv_stats || !L.parts || !L.splits || !L.wgrad || !L.err) { fprintf(stderr, "missing symbol\n"); return 2; }
    int fails = 0;
    const bool all = !strcmp(what, "all");
    if (all || !strcmp(what, "conv")) {
        fails += check_conv(L, 1, 64, 13, 11, 136, 1, 1, false, false);      // flat 1x1, 256-wide channel tile, partial tiles
        fails += check_conv(L, 1, 72, 20, 24, 40, 3, 3, true, false);        // Cin % 64 != 0, epilogue, 128-wide tile
        fails += check_conv(L, 2, 128, 17, 33, 256, 3, 2, true, false);      // two K blocks per tap, 256-wide tile
        // CTA-pair kernel (conv_tc2.cu, Cout > 128): several channel tiles with a partial last one + residual through TMA,
        // odd number of pixel tiles (the peer's last tile is beyond the tensor), flat 1x1 with an even tile count
        fails += check_conv(L, 2, 128, 17, 33, 520, 3, 2, true, false);
        fails += check_conv(L, 1, 64, 23, 17, 256, 3, 5, true, false);
        fails += check_conv(L, 1, 256, 30, 40, 512, 1, 1, true, false);
        fails += check_conv(L, 3, 192, 9, 40, 264, 3, 1, false, false);
        // flat-tile kernel (conv_tc3.cu): tiles that straddle image boundaries, taps entirely outside the map, Cin % 64 != 0
        // with several channel blocks, an M tail shorter than one epilogue warp's 32 rows
        fails += check_conv(L, 3, 304, 13, 15, 256, 3, 12, true, false);
        fails += check_conv(L, 2, 64, 33, 31, 48, 3, 1, true, false);
        fails += check_conv(L, 5, 128, 9, 9, 1024, 1, 1, true, false);
    }
    if (all || !strcmp(what, "stats")) {
        fails += check_conv(L, 2, 64, 17, 19, 128, 3, 1, false, true);
        fails += check_conv(L, 1, 128, 20, 35, 256, 3, 2, false, true);      // 256-wide tile (largest shared-memory footprint)
        fails += check_conv(L, 3, 64, 5, 7, 264, 1, 1, false, true);         // flat 1x1, partial channel tile
        fails += check_conv(L, 3, 64, 23, 17, 264, 3, 2, false, true);       // pair kernel: odd tile count, partial channel tile
        fails += check_conv(L, 2, 64, 16, 32, 512, 1, 1, false, true);       // pair kernel: flat, even tile count, two channel tiles
    }
    if (all || !strcmp(what, "xform")) {
        fails += check_conv_xform(L, 2, 72, 11, 19, 40, 3, 2, false);        // Cin % 64 != 0, padding rows, 128-wide tile
        fails += check_conv_xform(L, 1, 128, 20, 35, 256, 3, 1, true);       // two channel blocks x nine taps, statistics
        fails += check_conv_xform(L, 3, 64, 5, 7, 136, 1, 1, false);         // flat 1x1, rows beyond the tensor
    }
    if (all || !strcmp(what, "pool")) {
        fails += check_pool(L, 2, 17, 19, 16);
        fails += check_pool(L, 1, 16, 12, 128);
        fails += check_pool(L, 1, 257, 257, 8);
    }
    if (all || !strcmp(what, "wgrad")) {
        fails += check_wgrad(L, 1, 264, 7, 17, 136, 2);
        fails += check_wgrad(L, 2, 512, 9, 20, 256, 12);                     // taps that fall entirely outside the map, 2 ci tiles
    }
    if (!strcmp(what, "perf")) { perf(L, argc > 2 ? atoi(argv[2]) : -1); return 0; }   // perf [shape index]
    printf("%s\n", fails ? "SELFTEST FAILED" : "SELFTEST PASSED");
    return fails ? 1 : 0;
}
