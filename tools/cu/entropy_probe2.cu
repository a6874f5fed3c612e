#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../u2pl_b200/csrc/arith.cuh"
using namespace u2pl;
// Packed f32x2 variant of the contract arithmetic, kept only in this probe (see the note in arith.cuh):
// on B200 with nvcc 12.9 the .x half of det_logf2 comes out wrong inside the unrolled class loop.
struct PackedConsts { float2 magic, neg_magic, log2e, neg_ln2, ln2, one, neg_one, neg_half, tiny; float2 q[5]; float2 r[7]; };
#define P2(v) {v, v}
__constant__ PackedConsts kPk = {P2(12582912.0f), P2(-12582912.0f), P2(1.44269502162933349609375f), P2(-0.693147182464599609375f),
    P2(0.693147182464599609375f), P2(1.0f), P2(-1.0f), P2(-0.5f), P2(1e-10f),
    {P2(0.0013933652080595493f), P2(0.008363181725144386f), P2(0.04166646674275398f), P2(0.16666576266288757f), P2(0.5f)},
    {P2(0.08507229387760162f), P2(-0.14198024570941925f), P2(0.1495114266872406f), P2(-0.16587895154953003f),
     P2(0.1996057629585266f), P2(-0.2500097155570984f), P2(0.33333972096443176f)}};
__device__ __forceinline__ float2 det_expf2(float2 d)
{
    d.x = fmaxf(d.x, -87.0f); d.y = fmaxf(d.y, -87.0f);
    const float2 t = __ffma2_rn(d, kPk.log2e, kPk.magic);
    const float2 kf = __fadd2_rn(t, kPk.neg_magic);
    const float2 r = __ffma2_rn(kf, kPk.neg_ln2, d);
    float2 q = kPk.q[0];
    for (int i = 1; i < 5; ++i) q = __ffma2_rn(q, r, kPk.q[i]);
    float2 p = __fadd2_rn(__ffma2_rn(__fmul2_rn(r, r), q, r), kPk.one);
    p.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
    p.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
    return p;
}
__device__ __forceinline__ float2 det_logf2(float2 y)
{
    const uint32_t ix = __float_as_uint(y.x), iy = __float_as_uint(y.y);
    const int32_t ex = static_cast<int32_t>(ix - 0x3f3504f3u) >> 23, ey = static_cast<int32_t>(iy - 0x3f3504f3u) >> 23;
    const float2 m = make_float2(__uint_as_float(ix - (static_cast<uint32_t>(ex) << 23)), __uint_as_float(iy - (static_cast<uint32_t>(ey) << 23)));
    const float2 ef = __fadd2_rn(make_float2(__uint_as_float(0x4B400000u + static_cast<uint32_t>(ex)),
                                             __uint_as_float(0x4B400000u + static_cast<uint32_t>(ey))), kPk.neg_magic);
    const float2 f = __fadd2_rn(m, kPk.neg_one);
    float2 R = kPk.r[0];
    for (int i = 1; i < 7; ++i) R = __ffma2_rn(R, f, kPk.r[i]);
    const float2 l = __fadd2_rn(f, __fmul2_rn(__fmul2_rn(f, f), __ffma2_rn(f, R, kPk.neg_half)));
    return __ffma2_rn(ef, kPk.ln2, l);
}
constexpr int C = 21;
// stage dump: for thread 0 only; packed vs scalar for lane x
__global__ void probe(const float *in, float *out)
{
    float a[C], b[C];
    float2 v[C];
    for (int c = 0; c < C; ++c) { a[c] = in[c]; b[c] = in[C + c]; v[c] = make_float2(a[c], b[c]); }
    // packed
    float2 m = v[0];
    for (int c = 1; c < C; ++c) { m.x = fmaxf(m.x, v[c].x); m.y = fmaxf(m.y, v[c].y); }
    const float2 nm = __fmul2_rn(m, kPk.neg_one);
    float2 S = make_float2(0.f, 0.f);
    float2 d0 = __fadd2_rn(v[0], nm);
    for (int c = 0; c < C; ++c) { v[c] = det_expf2(__fadd2_rn(v[c], nm)); S = __fadd2_rn(S, v[c]); }
    const float2 rinv = make_float2(__fdiv_rn(1.0f, S.x), __fdiv_rn(1.0f, S.y));
    float2 acc = make_float2(0.f, 0.f);
    float2 p0, l0, y0;
    for (int c = 0; c < C; ++c) {
        const float2 p = __fmul2_rn(v[c], rinv);
        const float2 y = __fadd2_rn(p, kPk.tiny);
        const float2 l = det_logf2(y);
        if (c == 0) { p0 = p; l0 = l; y0 = y; }
        acc = __ffma2_rn(p, l, acc);
    }
    // scalar lane x
    float ms = a[0];
    for (int c = 1; c < C; ++c) ms = fmaxf(ms, a[c]);
    float Ss = 0.f, e0s = 0.f;
    float es[C];
    for (int c = 0; c < C; ++c) { es[c] = det_expf(__fadd_rn(a[c], -ms)); Ss = __fadd_rn(Ss, es[c]); }
    e0s = es[0];
    float rs = __fdiv_rn(1.0f, Ss);
    float accs = 0.f, p0s = 0, l0s = 0;
    for (int c = 0; c < C; ++c) { float p = __fmul_rn(es[c], rs); float l = det_logf(__fadd_rn(p, 1e-10f)); if (c == 0) { p0s = p; l0s = l; } accs = __fmaf_rn(p, l, accs); }
    float vals[] = {m.x, ms, nm.x, -ms, d0.x, a[0] - ms, v[0].x, e0s, S.x, Ss, rinv.x, rs, p0.x, p0s, y0.x, p0s + 1e-10f, l0.x, l0s, acc.x, accs,
                    m.y, nm.y, S.y, rinv.y, p0.y, l0.y, acc.y, 0};
    for (int i = 0; i < 28; ++i) out[i] = vals[i];
}
int main()
{
    float h[2 * C], *d, *o, r[28];
    srand(1);
    for (int i = 0; i < 2 * C; ++i) h[i] = 6.f * (rand() / (float)RAND_MAX - 0.5f);
    cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, sizeof(r));
    cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    probe<<<1, 1>>>(d, o);
    cudaMemcpy(r, o, sizeof(r), cudaMemcpyDeviceToHost);
    const char *names[] = {"m", "nm", "d0", "e0", "S", "rinv", "p0", "y0", "l0", "acc"};
    for (int i = 0; i < 10; ++i) printf("%-5s packed.x=%-14.8g scalar=%-14.8g %s\n", names[i], r[2 * i], r[2 * i + 1], r[2 * i] == r[2 * i + 1] ? "" : "<<<");
    printf("lane y: m=%g nm=%g S=%g rinv=%g p0=%g l0=%g acc=%g\n", r[20], r[21], r[22], r[23], r[24], r[25], r[26]);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
