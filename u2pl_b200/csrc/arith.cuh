// arith.cuh -- the device side of the arithmetic contract.
//
// exp and log used by every kernel whose result feeds an ordering decision
// (entropy percentiles, reliable/unreliable masks, rank windows) are defined
// here ONLY through IEEE-754 binary32 round-to-nearest operations
// (__fmaf_rn/__fmul_rn/__fadd_rn/__fdiv_rn never get re-associated or fused by
// nvcc) plus integer bit manipulation, so that the result is a pure function
// of the input bits and is reproducible on any IEEE machine.  The CPU oracle
// (oracle/u2pl_oracle.c) implements the same contract independently; parity
// tests require bit equality between the two.
//
// Accuracy against the correctly rounded functions: exp <= 1.5 ulp for
// d >= -20, log <= 1.2 ulp on [1e-10, 1]; this is the same class as CUDA's and
// ATen's own expf/logf which the reference reaches through torch.softmax /
// torch.log (loss_helper.py:35-36).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace u2pl {

constexpr float kMagic = 12582912.0f;                 // 1.5 * 2^23
constexpr float kLog2e = 1.44269502162933349609375f;
constexpr float kLn2   = 0.693147182464599609375f;

// exp(d), d <= 0 (clamped at -87).
__device__ __forceinline__ float det_expf(float d)
{
    d = fmaxf(d, -87.0f);
    const float t  = __fmaf_rn(d, kLog2e, kMagic);    // k = rint(d*log2e) sits in the low mantissa bits
    const float kf = __fadd_rn(t, -kMagic);
    const float r  = __fmaf_rn(kf, -kLn2, d);
    float q = 0.0013933652080595493f;
    q = __fmaf_rn(q, r, 0.008363181725144386f);
    q = __fmaf_rn(q, r, 0.04166646674275398f);
    q = __fmaf_rn(q, r, 0.16666576266288757f);
    q = __fmaf_rn(q, r, 0.5f);
    const float r2 = __fmul_rn(r, r);
    float p = __fmaf_rn(r2, q, r);
    p = __fadd_rn(p, 1.0f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

// log(y), y positive normal.
__device__ __forceinline__ float det_logf(float y)
{
    const uint32_t ix = __float_as_uint(y);
    const int32_t  e  = static_cast<int32_t>(ix - 0x3f3504f3u) >> 23;
    const float m  = __uint_as_float(ix - (static_cast<uint32_t>(e) << 23));
    const float ef = __fadd_rn(__uint_as_float(0x4B400000u + static_cast<uint32_t>(e)), -kMagic);
    const float f  = __fadd_rn(m, -1.0f);
    float R = 0.08507229387760162f;
    R = __fmaf_rn(R, f, -0.14198024570941925f);
    R = __fmaf_rn(R, f, 0.1495114266872406f);
    R = __fmaf_rn(R, f, -0.16587895154953003f);
    R = __fmaf_rn(R, f, 0.1996057629585266f);
    R = __fmaf_rn(R, f, -0.2500097155570984f);
    R = __fmaf_rn(R, f, 0.33333972096443176f);
    const float f2 = __fmul_rn(f, f);
    const float u  = __fmaf_rn(f, R, -0.5f);
    const float tt = __fmul_rn(f2, u);
    const float l  = __fadd_rn(f, tt);
    return __fmaf_rn(ef, kLn2, l);
}


// ---- packed (f32x2) forms: sm_100 issues two IEEE binary32 operations per instruction
// (FFMA2 / FMUL2 / FADD2).  Every lane-half performs exactly the scalar sequence above, so the
// results are bit-identical to det_expf / det_logf; only the issue-slot count halves.
__device__ __forceinline__ float2 f2(float v) { return make_float2(v, v); }

__device__ __forceinline__ float2 det_expf2(float2 d)
{
    d.x = fmaxf(d.x, -87.0f);
    d.y = fmaxf(d.y, -87.0f);
    const float2 t  = __ffma2_rn(d, f2(kLog2e), f2(kMagic));
    const float2 kf = __fadd2_rn(t, f2(-kMagic));
    const float2 r  = __ffma2_rn(kf, f2(-kLn2), d);
    float2 q = f2(0.0013933652080595493f);
    q = __ffma2_rn(q, r, f2(0.008363181725144386f));
    q = __ffma2_rn(q, r, f2(0.04166646674275398f));
    q = __ffma2_rn(q, r, f2(0.16666576266288757f));
    q = __ffma2_rn(q, r, f2(0.5f));
    const float2 r2 = __fmul2_rn(r, r);
    float2 p = __ffma2_rn(r2, q, r);
    p = __fadd2_rn(p, f2(1.0f));
    p.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
    p.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
    return p;
}

__device__ __forceinline__ float2 det_logf2(float2 y)
{
    const uint32_t ix = __float_as_uint(y.x), iy = __float_as_uint(y.y);
    const int32_t ex = static_cast<int32_t>(ix - 0x3f3504f3u) >> 23;
    const int32_t ey = static_cast<int32_t>(iy - 0x3f3504f3u) >> 23;
    const float2 m = make_float2(__uint_as_float(ix - (static_cast<uint32_t>(ex) << 23)),
                                 __uint_as_float(iy - (static_cast<uint32_t>(ey) << 23)));
    const float2 ef = __fadd2_rn(make_float2(__uint_as_float(0x4B400000u + static_cast<uint32_t>(ex)),
                                             __uint_as_float(0x4B400000u + static_cast<uint32_t>(ey))), f2(-kMagic));
    const float2 f = __fadd2_rn(m, f2(-1.0f));
    float2 R = f2(0.08507229387760162f);
    R = __ffma2_rn(R, f, f2(-0.14198024570941925f));
    R = __ffma2_rn(R, f, f2(0.1495114266872406f));
    R = __ffma2_rn(R, f, f2(-0.16587895154953003f));
    R = __ffma2_rn(R, f, f2(0.1996057629585266f));
    R = __ffma2_rn(R, f, f2(-0.2500097155570984f));
    R = __ffma2_rn(R, f, f2(0.33333972096443176f));
    const float2 fsq = __fmul2_rn(f, f);
    const float2 u  = __ffma2_rn(f, R, f2(-0.5f));
    const float2 tt = __fmul2_rn(fsq, u);
    const float2 l  = __fadd2_rn(f, tt);
    return __ffma2_rn(ef, f2(kLn2), l);
}

// Order-preserving float -> uint32 key (total order, -0 < +0).
__device__ __forceinline__ uint32_t float_key(float v)
{
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k)
{
    const uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

}  // namespace u2pl
