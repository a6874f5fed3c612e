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


// NOTE on packed f32x2 (FFMA2/FADD2/FMUL2, sm_100): a two-pixels-per-thread variant of the entropy
// arithmetic was built and measured in round 1.  (1) It gave no throughput: FFMA2 issues at half the
// FFMA rate on B200, and the kernel is bound by the integer (ALU-pipe) half of exp/log anyway.
// (2) With nvcc/ptxas 12.9 the unrolled 21-class loop produced wrong .x (LO-half) results for
// det_logf2 while the same function was correct in isolation (tools/cu/entropy_probe2.cu dumps the
// stages) -- a miscompile or a missing hazard on register pairs fed by integer ops.  The scalar
// form below is therefore the only one shipped.

// Scalar per-pixel entropy under the contract (C register-resident logits in v[]).
template <int C>
__device__ __forceinline__ float entropy_of(float (&v)[C])
{
    float m = v[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, v[c]);
    float S = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        v[c] = det_expf(__fadd_rn(v[c], -m));
        S = __fadd_rn(S, v[c]);
    }
    const float rinv = __fdiv_rn(1.0f, S);
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float p = __fmul_rn(v[c], rinv);
        const float l = det_logf(__fadd_rn(p, 1e-10f));
        acc = __fmaf_rn(p, l, acc);
    }
    return -acc;
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
