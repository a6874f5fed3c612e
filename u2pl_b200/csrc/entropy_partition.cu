// entropy_partition.cu -- fused softmax-entropy + on-device percentile thresholds +
// reliable/unreliable partition.  Replaces loss_helper.py:35-44 and
// train_semi.py:402-418 of the reference (ATen softmax/log/sum, boolean gather,
// D2H copy, np.percentile on the host, mask algebra) with
//
//   K1  entropy_hist    one pass over the [B,C,HW] logits: entropy out, order-preserving
//                       uint32 keys out, 12-bit radix histogram of the valid keys
//   K2  select1         ranks lo/hi of every requested percentile (numpy float32 index
//                       arithmetic) + the 12-bit bin each rank falls in
//   K3  hist_refine(2)  10-bit histogram inside those bins      (keys are L2 resident)
//   K4  select_refine   K5 hist_refine(3)   K6 select_refine -> thresholds (two-sided lerp)
//   K7  partition       target rewrite + drop mask + kept count
//
// HBM-bound by design: K1 moves (4C + 8 + 4 + 4) B/pixel, K7 moves 4+8+8+1 B/pixel,
// K3/K5 re-read 4 B/pixel of keys out of the 126 MB L2.  Nothing synchronises the host.
#include <algorithm>
#include <cstdlib>
#include "arith.cuh"
#include "common.cuh"

namespace u2pl {

constexpr int kMaxQ = U2PL_MAX_QUANTILES;
constexpr int kMaxT = 2 * kMaxQ;          // (lo, hi) order statistic per percentile
constexpr int kBins1 = 4096;              // pass 1: key bits 31..20
constexpr int kBinsR = 1024;              // pass 2: bits 19..10, pass 3: bits 9..0
constexpr uint32_t kInvalidKey = 0xFFFFFFFFu;

struct SelState {
    uint32_t prefix[kMaxT];
    uint32_t rank[kMaxT];           // rank inside the current prefix bin
    float    gamma[kMaxQ];
    uint32_t n;
    uint32_t done;                  // two-level path: blocks of exact_select that have finished
    uint32_t grank[kMaxT];          // global rank of every target order statistic
    uint32_t cnt[kMaxT];            // two-level path: candidates gathered per target
    uint32_t below[kMaxT];          //                 valid pixels surely below the candidate band
    float    val[kMaxT];            //                 exact order statistics
    uint32_t band[kMaxT];           //                 candidate band (list / counter index) of every target
    uint32_t bar;                   // fused chain: grid-barrier arrival counter (zeroed with the rest of the state)
    uint32_t next_block;            // fused chain: dynamic pixel-block scheduler
    uint32_t nkmin[kMaxT];          // fused chain: max over the band's candidates of ~key (zero-initialised minimum)
    uint32_t kmax[kMaxT];           //              max key
    unsigned long long stamp[12];    // fused chain: %globaltimer of CTA 0 at the phase boundaries (U2PL_CHAIN_TIMING=1 prints them)
};

struct Percents { float q[kMaxQ]; int use_rank; uint32_t rank; };   // use_rank: one explicit order statistic instead

// ------------------------------------------------------------------ K1
// warp-aggregated shared-memory histogram increment: lanes that hit the same bin elect one leader
// (real teacher logits put most pixels into a handful of bins; plain atomics would serialise 32-way)
__device__ __forceinline__ void hist_add(uint32_t *sh, uint32_t bin)
{
    const uint32_t m = __activemask();
    const uint32_t peers = __match_any_sync(m, bin);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sh[bin], __popc(peers));
}

constexpr int kEntThreads = 256;

template <int C>
__global__ void __launch_bounds__(kEntThreads)
entropy_hist_kernel(const float *__restrict__ logits, const int64_t *__restrict__ target,
                    uint32_t HW, uint32_t N, int64_t ignore,
                    float *__restrict__ ent, uint32_t *__restrict__ keys,
                    uint32_t *__restrict__ hist1)
{
    __shared__ uint32_t sh[kBins1];
    for (int j = threadIdx.x; j < kBins1; j += kEntThreads) sh[j] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * kEntThreads + threadIdx.x; i < N; i += gridDim.x * kEntThreads) {
        const uint32_t b = i / HW, p = i - b * HW;
        const float *x = logits + static_cast<size_t>(b) * C * HW + p;
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = __ldg(x + static_cast<size_t>(c) * HW);     // C coalesced loads in flight
        const int64_t t = __ldg(target + i);
        const float h = entropy_of<C>(v);
        ent[i] = h;
        const bool valid = (t != ignore);
        const uint32_t key = valid ? float_key(h) : kInvalidKey;
        keys[i] = key;
        if (valid) hist_add(sh, key >> 20);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < kBins1; j += kEntThreads)
        if (sh[j]) atomicAdd(&hist1[j], sh[j]);
}

// Any class count: three passes over the C axis per pixel (max, sum, entropy); same
// arithmetic, same operation order, identical bits -- only slower (L1/L2 re-reads).
__global__ void __launch_bounds__(256)
entropy_hist_kernel_anyC(const float *__restrict__ logits, const int64_t *__restrict__ target,
                         uint32_t C, uint32_t HW, uint32_t N, int64_t ignore,
                         float *__restrict__ ent, uint32_t *__restrict__ keys,
                         uint32_t *__restrict__ hist1)
{
    __shared__ uint32_t sh[kBins1];
    for (int j = threadIdx.x; j < kBins1; j += 256) sh[j] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < N; i += gridDim.x * 256u) {
        const uint32_t b = i / HW, p = i - b * HW;
        const float *x = logits + static_cast<size_t>(b) * C * HW + p;
        float m = __ldg(x);
        for (uint32_t c = 1; c < C; ++c) m = fmaxf(m, __ldg(x + static_cast<size_t>(c) * HW));
        float S = 0.0f;
        for (uint32_t c = 0; c < C; ++c)
            S = __fadd_rn(S, det_expf(__fadd_rn(__ldg(x + static_cast<size_t>(c) * HW), -m)));
        const float rinv = __fdiv_rn(1.0f, S);
        float acc = 0.0f;
        for (uint32_t c = 0; c < C; ++c) {
            const float e = det_expf(__fadd_rn(__ldg(x + static_cast<size_t>(c) * HW), -m));
            const float pr = __fmul_rn(e, rinv);
            acc = __fmaf_rn(pr, det_logf(__fadd_rn(pr, 1e-10f)), acc);
        }
        const float h = -acc;
        ent[i] = h;
        const bool valid = (__ldg(target + i) != ignore);
        const uint32_t key = valid ? float_key(h) : kInvalidKey;
        keys[i] = key;
        if (valid) hist_add(sh, key >> 20);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < kBins1; j += 256)
        if (sh[j]) atomicAdd(&hist1[j], sh[j]);
}

// ------------------------------------------------------------------ K2
// One block of 256 threads, thread t owns bins [16t, 16t+16).
__global__ void __launch_bounds__(256)
select1_kernel(const uint32_t *__restrict__ hist1, SelState *__restrict__ st, Percents pc, int nq)
{
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t s_rank[kMaxT];
    __shared__ uint32_t s_n;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t loc[16];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) { loc[j] = hist1[tid * 16 + j]; sum += loc[j]; }
    uint32_t inc = sum;                                    // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < wid; ++w) base += warp_tot[w];
    const uint32_t excl = base + inc - sum;
    if (tid == 255) s_n = excl + sum;
    __syncthreads();
    const uint32_t n = s_n;
    if (tid == 0) {
        st->n = n;
        // numpy 2.x: q32 = q / float32(100); v = float32(n-1) * q32  (all binary32)
        const float nm1 = __uint2float_rn(n ? n - 1u : 0u);
        if (pc.use_rank) {                                  // OHEM: the k-th smallest value itself (no interpolation)
            const uint32_t r = n ? min(pc.rank, n - 1u) : 0u;
            st->gamma[0] = 0.0f;
            s_rank[0] = s_rank[1] = r;
        }
        for (int j = 0; j < (pc.use_rank ? 0 : nq); ++j) {
            const float q32 = __fdiv_rn(pc.q[j], 100.0f);
            const float v = __fmul_rn(nm1, q32);
            const float fl = floorf(v);
            uint32_t lo, hi;
            if (n == 0) { lo = hi = 0; }
            else if (v >= nm1) { lo = hi = n - 1u; }
            else { lo = static_cast<uint32_t>(fl); hi = lo + 1u; }
            st->gamma[j] = __fadd_rn(v, -fl);
            s_rank[2 * j] = lo;
            s_rank[2 * j + 1] = hi;
        }
    }
    __syncthreads();
    if (n == 0) return;
    for (int t = 0; t < 2 * nq; ++t) {
        const uint32_t r = s_rank[t];
        if (r >= excl && r < excl + sum) {
            uint32_t cum = excl;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (r < cum + loc[j]) { st->prefix[t] = tid * 16 + j; st->rank[t] = r - cum; st->grank[t] = r; break; }
                cum += loc[j];
            }
        }
    }
}

// ------------------------------------------------------------------ K3 / K5
// pass 2: match key>>20 against the 12-bit prefix, bin on bits 19..10
// pass 3: match key>>10 against the 22-bit prefix, bin on bits  9..0
template <int PASS>
__global__ void __launch_bounds__(256)
hist_refine_kernel(const uint32_t *__restrict__ keys, uint32_t N,
                   const SelState *__restrict__ st, uint32_t *__restrict__ hist, int T)
{
    constexpr int kMatchShift = (PASS == 2) ? 20 : 10;
    constexpr int kBinShift = (PASS == 2) ? 10 : 0;
    extern __shared__ uint32_t sh[];                      // [T][1024]
    for (int j = threadIdx.x; j < T * kBinsR; j += 256) sh[j] = 0;
    uint32_t pre[kMaxT];
#pragma unroll
    for (int t = 0; t < kMaxT; ++t) pre[t] = (t < T) ? st->prefix[t] : kInvalidKey;
    const bool empty = (st->n == 0);
    __syncthreads();
    if (!empty) {
        const uint32_t n4 = N >> 2;
        const uint4 *k4 = reinterpret_cast<const uint4 *>(keys);
        for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n4; i += gridDim.x * 256u) {
            const uint4 q = __ldg(k4 + i);
            const uint32_t kk[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t key = kk[e];
                if (key == kInvalidKey) continue;
                const uint32_t hi = key >> kMatchShift;
                uint32_t mt = 0;                                   // targets whose prefix this key extends
#pragma unroll
                for (int t = 0; t < kMaxT; ++t) mt |= (t < T && hi == pre[t]) ? (1u << t) : 0u;
                if (mt) {
                    const uint32_t bin = (key >> kBinShift) & (kBinsR - 1);
                    const uint32_t am = __activemask();
                    const uint32_t peers = __match_any_sync(am, (mt << 10) | bin);
                    if ((threadIdx.x & 31) == __ffs(peers) - 1) {
                        const uint32_t cnt = __popc(peers);
                        while (mt) {
                            const int t = __ffs(mt) - 1;
                            mt &= mt - 1;
                            atomicAdd(&sh[t * kBinsR + bin], cnt);
                        }
                    }
                }
            }
        }
        if (blockIdx.x == 0 && threadIdx.x < (N & 3u)) {
            const uint32_t key = keys[(n4 << 2) + threadIdx.x];
            if (key != kInvalidKey) {
                const uint32_t hi = key >> kMatchShift;
                const uint32_t bin = (key >> kBinShift) & (kBinsR - 1);
                for (int t = 0; t < T; ++t)
                    if (hi == pre[t]) atomicAdd(&sh[t * kBinsR + bin], 1u);
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < T * kBinsR; j += 256)
        if (sh[j]) atomicAdd(&hist[j], sh[j]);
}

// ------------------------------------------------------------------ K4 / K6
// One block of 1024 threads; thread j owns bin j of the current target's histogram.
template <bool FINAL>
__global__ void __launch_bounds__(1024)
select_refine_kernel(const uint32_t *__restrict__ hist, SelState *__restrict__ st, int nq,
                     float *__restrict__ thresh, int64_t *__restrict__ n_valid)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ float s_val[kMaxT];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t n = st->n;
    if (n == 0) {
        if (FINAL && tid == 0) {
            for (int j = 0; j < nq; ++j) thresh[j] = __uint_as_float(0x7fc00000u);
            if (n_valid) *n_valid = 0;
        }
        return;
    }
    for (int t = 0; t < 2 * nq; ++t) {
        const uint32_t cnt = hist[t * kBinsR + tid];
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < wid; ++w) base += warp_tot[w];
        const uint32_t excl = base + inc - cnt;
        const uint32_t r = st->rank[t];
        const uint32_t pre = st->prefix[t];
        __syncthreads();                                  // everyone has read rank/prefix/warp_tot
        if (r >= excl && r < excl + cnt) {
            const uint32_t np = (pre << 10) | static_cast<uint32_t>(tid);
            st->prefix[t] = np;
            st->rank[t] = r - excl;
            if (FINAL) s_val[t] = key_float(np);
        }
        __syncthreads();
    }
    if (FINAL && tid == 0) {
        for (int j = 0; j < nq; ++j) {
            // numpy _lerp: a + (b-a)*g, and b - (b-a)*(1-g) where g >= 0.5  (mul and add rounded separately)
            const float a = s_val[2 * j], b = s_val[2 * j + 1], g = st->gamma[j];
            const float d = __fadd_rn(b, -a);
            float r = __fadd_rn(a, __fmul_rn(d, g));
            if (g >= 0.5f) r = __fadd_rn(b, -__fmul_rn(d, __fadd_rn(1.0f, -g)));
            thresh[j] = r;
        }
        if (n_valid) *n_valid = static_cast<int64_t>(n);
    }
}

// ------------------------------------------------------------------ K7
__global__ void __launch_bounds__(256)
partition_kernel(const float *__restrict__ ent, int64_t *__restrict__ target, uint32_t N,
                 int64_t ignore, const float *__restrict__ thresh, int idx,
                 uint8_t *__restrict__ drop_mask, unsigned long long *__restrict__ n_kept)
{
    const float th = __ldg(thresh + idx);
    int kept = 0;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < N; i += gridDim.x * 256u) {
        const int64_t t = target[i];
        const bool valid = (t != ignore);
        const bool drop = valid && (ent[i] >= th);
        if (drop) target[i] = ignore;
        if (drop_mask) drop_mask[i] = drop ? 1 : 0;
        kept += (valid && !drop) ? 1 : 0;
    }
    kept = warp_sum_i(kept);
    __shared__ int wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = kept;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < 8; ++w) s += wsum[w];
        if (s) atomicAdd(n_kept, static_cast<unsigned long long>(s));
    }
}

__global__ void __launch_bounds__(256)
entropy_masks_kernel(const float *__restrict__ ent, const int64_t *__restrict__ target,
                     const int64_t *__restrict__ idx, uint32_t n_out, int64_t ignore,
                     const float *__restrict__ thresh, int lo_idx, int hi_idx,
                     float *__restrict__ out_low, float *__restrict__ out_high)
{
    const float tl = __ldg(thresh + lo_idx), thh = __ldg(thresh + hi_idx);
    for (uint32_t j = blockIdx.x * 256u + threadIdx.x; j < n_out; j += gridDim.x * 256u) {
        const int64_t s = idx ? __ldg(idx + j) : static_cast<int64_t>(j);
        const float e = __ldg(ent + s);
        const bool valid = (__ldg(target + s) != ignore);
        if (out_low) out_low[j] = (valid && e <= tl) ? 1.0f : 0.0f;
        if (out_high) out_high[j] = (valid && e >= thh) ? 1.0f : 0.0f;
    }
}



// ====================================================================== two-level path
// "Exact where it matters".  The masks only depend on how every entropy compares with the
// thresholds, and the thresholds only depend on a few order statistics.  So:
//   F1 entropy_fast_hist   entropies with hardware ex2/lg2 (2 MUFU + ~6 FP32 ops per class instead of
//                          ~43 issue slots): HBM-bound.  |fast - contract| <= kDelta (see below).
//   F2 fast_refine         select1 + 10-bit histogram on the FAST keys: a 22-bit key bin per target rank
//   F3 fast_candidate      every valid pixel whose fast entropy lies within 3*kDelta of a target's bin is
//                          re-evaluated under the arithmetic contract (exact value stored back into
//                          `entropy`, exact key appended to the target's candidate list); pixels surely
//                          below the band are only counted
//   F4 exact_select        per target: radix select of rank (global rank - #below) among its candidates;
//                          last block: numpy lerp -> thresholds
// Soundness: order statistics are 1-Lipschitz in the sup norm, so the exact r-th value is within kDelta
// of the fast r-th value, which lies in the target's bin.  Non-candidates below (above) the band have
// exact values < bin.lo - 2*kDelta (> bin.hi + 2*kDelta): their order relative to the exact statistic
// is known, hence rank_in_candidates = r - #below.  Every later comparison `entropy <=/>= threshold`
// is exact for candidates (exact values stored) and decided by a margin > kDelta for all others.
// If the invariant 0 <= rank_in_candidates < #candidates is ever violated the thresholds are set to NaN.
constexpr float kDelta = 1.0e-4f;     // bound on |fast - contract| (measured max ~3e-6; tests assert < kDelta/4)

// H = ln S - (sum_c e_c d_c) / S  with d_c = x_c - max, e_c = exp(d_c), S = sum_c e_c: the identity for
// -sum p ln p.  It drops the reference's "+1e-10" inside the log, which moves H by at most C*1e-10, far
// inside kDelta; one MUFU (ex2) + 3 FP32 ops per class and a single lg2 per pixel.
template <int C>
__device__ __forceinline__ float entropy_fast_of(float (&v)[C])
{
    float m = v[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, v[c]);
    float S = 0.0f, Wd = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float d = v[c] - m;
        float e;                                           // MUFU.EX2 directly (exp2f adds denormal-range handling)
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(d * 1.4426950408889634f));
        S += e;
        Wd = fmaf(e, d, Wd);
    }
    return fmaf(0.6931471805599453f, __log2f(S), -__fdividef(Wd, S));
}

template <int C>
__global__ void __launch_bounds__(kEntThreads)
entropy_fast_hist_kernel(const float *__restrict__ logits, const int64_t *__restrict__ target,
                         uint32_t HW, uint32_t N, int64_t ignore,
                         float *__restrict__ ent, uint32_t *__restrict__ keys, uint32_t *__restrict__ hist1)
{
    __shared__ uint32_t sh[kBins1];
    for (int j = threadIdx.x; j < kBins1; j += kEntThreads) sh[j] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * kEntThreads + threadIdx.x; i < N; i += gridDim.x * kEntThreads) {
        const uint32_t b = i / HW, p = i - b * HW;
        const float *x = logits + static_cast<size_t>(b) * C * HW + p;
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = __ldg(x + static_cast<size_t>(c) * HW);
        const int64_t t = __ldg(target + i);
        const float h = entropy_fast_of<C>(v);
        ent[i] = h;
        const bool valid = (t != ignore);
        const uint32_t key = valid ? float_key(h) : kInvalidKey;
        keys[i] = key;
        if (valid) hist_add(sh, key >> 20);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < kBins1; j += kEntThreads)
        if (sh[j]) atomicAdd(&hist1[j], sh[j]);
}

// one block per target: 11/11/10-bit radix select inside the candidate list; the last block to finish
// turns the exact order statistics into thresholds (numpy's two-sided lerp).
__global__ void __launch_bounds__(1024)
exact_select_kernel(const uint32_t *__restrict__ lists, uint32_t N, SelState *__restrict__ st, int nq,
                    float *__restrict__ thresh, int64_t *__restrict__ n_valid)
{
    __shared__ uint32_t hist[2048];
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_prefix, s_rank;
    __shared__ int s_last;
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t n_all = st->n;
    if (n_all != 0) {
        const uint32_t u = st->band[t];
        const uint32_t n = st->cnt[u];
        const uint32_t r0 = st->grank[t] - st->below[u];              // unsigned: a violated invariant shows as r0 >= n
        const uint32_t *list = lists + static_cast<size_t>(u) * N;
        if (r0 >= n) {
            if (tid == 0) st->val[t] = __uint_as_float(0x7fc00000u);
        } else {
            if (tid == 0) { s_prefix = 0; s_rank = r0; }
            const int shifts[3] = {21, 10, 0}, nbits[3] = {11, 11, 10};
            for (int pass = 0; pass < 3; ++pass) {
                hist[tid] = 0; hist[tid + 1024] = 0;
                __syncthreads();
                const uint32_t pre = s_prefix, r = s_rank;
                const int sh = shifts[pass], hb = sh + nbits[pass];
                const uint32_t mask = (1u << nbits[pass]) - 1u;
                for (uint32_t j = tid; j < n; j += 1024) {
                    const uint32_t k = __ldg(list + j);
                    if (pass == 0 || (k >> hb) == pre) atomicAdd(&hist[(k >> sh) & mask], 1u);
                }
                __syncthreads();
                const uint32_t c0 = hist[2 * tid], c1 = hist[2 * tid + 1], sum = c0 + c1;
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += y;
                }
                if (lane == 31) warp_tot[wid] = inc;
                __syncthreads();
                uint32_t base = 0;
                for (int w = 0; w < wid; ++w) base += warp_tot[w];
                const uint32_t excl = base + inc - sum;
                __syncthreads();
                if (r >= excl && r < excl + sum) {
                    const uint32_t bin = (r < excl + c0) ? 2 * tid : 2 * tid + 1;
                    s_prefix = (pre << nbits[pass]) | bin;
                    s_rank = r - ((r < excl + c0) ? excl : excl + c0);
                }
                __syncthreads();
            }
            if (tid == 0) st->val[t] = key_float(s_prefix);
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = (atomicAdd(&st->done, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && tid == 0) {
        __threadfence();
        for (int j = 0; j < nq; ++j) {
            if (n_all == 0) { thresh[j] = __uint_as_float(0x7fc00000u); continue; }
            const volatile float *vv = st->val;
            const float a = vv[2 * j], b = vv[2 * j + 1], g = st->gamma[j];
            const float d = __fadd_rn(b, -a);
            float r = __fadd_rn(a, __fmul_rn(d, g));
            if (g >= 0.5f) r = __fadd_rn(b, -__fmul_rn(d, __fadd_rn(1.0f, -g)));
            thresh[j] = r;
        }
        if (n_valid) *n_valid = static_cast<int64_t>(n_all);
    }
}


// ---- fused launches of the two-level path -------------------------------------------------------------
// The tiny single-block select kernels are folded into the prologue of the pass that consumes them (every
// block repeats the ~16 KB histogram scan; block 0 publishes the result), so the chain is
//   entropy_fast_hist -> fast_refine (select1 + 10-bit histogram) -> fast_candidate (select + band test +
//   exact re-evaluation) -> exact_select (+ lerp).

// select1 for all targets, executed by one 256-thread block; results in shared arrays.
__device__ __forceinline__ void block_select1(const uint32_t *__restrict__ hist1, const Percents &pc, int nq,
                                              uint32_t *s_prefix, uint32_t *s_rank, uint32_t *s_grank, float *s_gamma,
                                              uint32_t *s_n, uint32_t *warp_tot)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t loc[16], sum = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) { loc[j] = __ldg(hist1 + tid * 16 + j); sum += loc[j]; }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < wid; ++w) base += warp_tot[w];
    const uint32_t excl = base + inc - sum;
    if (tid == 255) *s_n = excl + sum;
    __syncthreads();
    const uint32_t n = *s_n;
    if (tid == 0) {
        const float nm1 = __uint2float_rn(n ? n - 1u : 0u);
        for (int j = 0; j < nq; ++j) {                    // numpy 2.x float32 virtual index (see select1_kernel)
            const float q32 = __fdiv_rn(pc.q[j], 100.0f);
            const float v = __fmul_rn(nm1, q32);
            const float fl = floorf(v);
            uint32_t lo, hi;
            if (n == 0) { lo = hi = 0; }
            else if (v >= nm1) { lo = hi = n - 1u; }
            else { lo = static_cast<uint32_t>(fl); hi = lo + 1u; }
            s_gamma[j] = __fadd_rn(v, -fl);
            s_grank[2 * j] = lo;
            s_grank[2 * j + 1] = hi;
        }
    }
    __syncthreads();
    if (n == 0) return;
    for (int t = 0; t < 2 * nq; ++t) {
        const uint32_t r = s_grank[t];
        if (r >= excl && r < excl + sum) {
            uint32_t cum = excl;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (r < cum + loc[j]) { s_prefix[t] = tid * 16 + j; s_rank[t] = r - cum; break; }
                cum += loc[j];
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
fast_refine_kernel(const uint32_t *__restrict__ keys, uint32_t N, const uint32_t *__restrict__ hist1,
                   SelState *__restrict__ st, Percents pc, int nq, uint32_t *__restrict__ hist2)
{
    extern __shared__ uint32_t sh[];                      // [T][1024]
    __shared__ uint32_t s_prefix[kMaxT], s_rank[kMaxT], s_grank[kMaxT], s_n, warp_tot[8];
    __shared__ float s_gamma[kMaxQ];
    const int T = 2 * nq;
    for (int j = threadIdx.x; j < T * kBinsR; j += 256) sh[j] = 0;
    block_select1(hist1, pc, nq, s_prefix, s_rank, s_grank, s_gamma, &s_n, warp_tot);
    const uint32_t n = s_n;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->n = n;
        for (int t = 0; t < T; ++t) { st->prefix[t] = s_prefix[t]; st->rank[t] = s_rank[t]; st->grank[t] = s_grank[t]; }
        for (int j = 0; j < nq; ++j) st->gamma[j] = s_gamma[j];
    }
    if (n == 0) return;
    uint32_t pre[kMaxT];                                  // targets with equal prefixes share the first one's histogram
#pragma unroll
    for (int t = 0; t < kMaxT; ++t) {
        pre[t] = (t < T) ? s_prefix[t] : kInvalidKey;
#pragma unroll
        for (int u = 0; u < t; ++u)
            if (t < T && s_prefix[u] == s_prefix[t]) pre[t] = kInvalidKey;     // key >> 20 never equals 0xFFFFFFFF
    }
    __syncthreads();
    const uint32_t n4 = N >> 2;
    const uint4 *k4 = reinterpret_cast<const uint4 *>(keys);
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n4 + 1; i += gridDim.x * 256u) {
        uint32_t kk[4];
        if (i < n4) { const uint4 q = __ldg(k4 + i); kk[0] = q.x; kk[1] = q.y; kk[2] = q.z; kk[3] = q.w; }
        else { for (int e = 0; e < 4; ++e) kk[e] = ((n4 << 2) + e < N) ? keys[(n4 << 2) + e] : kInvalidKey; }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t key = kk[e];
            if (key == kInvalidKey) continue;
            const uint32_t hi = key >> 20;
            uint32_t mt = 0;
#pragma unroll
            for (int t = 0; t < kMaxT; ++t) mt |= (hi == pre[t]) ? (1u << t) : 0u;
            if (mt) {
                const uint32_t bin = (key >> 10) & (kBinsR - 1);
                const uint32_t peers = __match_any_sync(__activemask(), (mt << 10) | bin);
                if ((threadIdx.x & 31) == __ffs(peers) - 1) {
                    const uint32_t cnt = __popc(peers);
                    while (mt) { const int t = __ffs(mt) - 1; mt &= mt - 1; atomicAdd(&sh[t * kBinsR + bin], cnt); }
                }
            }
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < T * kBinsR; j += 256)
        if (sh[j]) atomicAdd(&hist2[j], sh[j]);
}

constexpr int kCandTile = 4096;                            // pixels scanned per block iteration

template <int C>
__global__ void __launch_bounds__(256)
fast_candidate_kernel(const float *__restrict__ logits, const uint32_t *__restrict__ keys, uint32_t HW, uint32_t N,
                      const uint32_t *__restrict__ hist2, SelState *__restrict__ st, int T,
                      float *__restrict__ ent, uint32_t *__restrict__ lists)
{
    __shared__ uint32_t s_pix[kCandTile];                 // candidates of the current tile: pixel index ...
    __shared__ uint8_t s_hit[kCandTile];                  // ... and the bands it falls into
    __shared__ uint32_t s_cnt, warp_tot[8], s_pre22[kMaxT], s_below[kMaxT], s_band[kMaxT], s_bcnt[kMaxT], s_bbase[kMaxT], s_bpos[kMaxT];
    __shared__ float s_lo[kMaxT], s_hi[kMaxT];
    __shared__ int s_nband;
    const uint32_t n_all = st->n;
    if (n_all == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // ---- select inside the 10-bit histograms: 22-bit fast-key bin per target.  All loads are issued before the
    // first use (one L2 round trip for the whole prologue; every block repeats it).
    __shared__ uint32_t s_pre12[kMaxT], s_rk[kMaxT];
    if (tid < T) { s_pre12[tid] = st->prefix[tid]; s_rk[tid] = st->rank[tid]; }
    __syncthreads();
    uint32_t cc[kMaxT][4];
#pragma unroll
    for (int t = 0; t < kMaxT; ++t) {
        if (t < T) {
            int owner = t;
            for (int u = t - 1; u >= 0; --u) if (s_pre12[u] == s_pre12[t]) owner = u;
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(hist2 + owner * kBinsR) + tid);
            cc[t][0] = q.x; cc[t][1] = q.y; cc[t][2] = q.z; cc[t][3] = q.w;
        }
    }
#pragma unroll
    for (int t = 0; t < kMaxT; ++t) {
        if (t < T) {                                       // T is uniform across the block
            const uint32_t sum = cc[t][0] + cc[t][1] + cc[t][2] + cc[t][3];
            uint32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            if (lane == 31) warp_tot[wid] = inc;
            __syncthreads();
            uint32_t base = 0;
            for (int w = 0; w < wid; ++w) base += warp_tot[w];
            const uint32_t excl = base + inc - sum, r = s_rk[t];
            if (r >= excl && r < excl + sum) {
                uint32_t cum = excl;
                int bin = 4 * tid;
#pragma unroll
                for (int j = 0; j < 4; ++j) { if (r >= cum && r < cum + cc[t][j]) bin = 4 * tid + j; cum += cc[t][j]; }
                s_pre22[t] = (s_pre12[t] << 10) | static_cast<uint32_t>(bin);
            }
            __syncthreads();
        }
    }
    // ---- merge targets with the same bin into one candidate band (lo/hi ranks of a percentile, repeated percentiles)
    if (tid == 0) {
        int nb = 0;
        for (int t = 0; t < T; ++t) {
            int b = -1;
            for (int u = 0; u < t; ++u) if (s_pre22[u] == s_pre22[t]) { b = static_cast<int>(s_band[u]); break; }
            if (b < 0) {
                b = nb++;
                const uint32_t pre = s_pre22[t];
                s_lo[b] = key_float(pre << 10) - 3.0f * kDelta;
                s_hi[b] = (pre == 0x3FFFFFu) ? __uint_as_float(0x7f800000u) : key_float((pre + 1u) << 10) + 3.0f * kDelta;
            }
            s_band[t] = static_cast<uint32_t>(b);
            if (blockIdx.x == 0) st->band[t] = static_cast<uint32_t>(b);
        }
        s_nband = nb;
    }
    if (tid < kMaxT) s_below[tid] = 0;
    __syncthreads();
    const int U = s_nband;
    float lo[kMaxT], hi[kMaxT];
    uint32_t below[kMaxT];
#pragma unroll
    for (int u = 0; u < kMaxT; ++u) {
        below[u] = 0;
        lo[u] = (u < U) ? s_lo[u] : __uint_as_float(0xff800000u);     // -inf: never below, never inside
        hi[u] = (u < U) ? s_hi[u] : lo[u];
    }
    for (uint32_t base = blockIdx.x * kCandTile; base < N; base += gridDim.x * kCandTile) {
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        // ---- phase 1: band test on the fast keys (16 per thread)
#pragma unroll 4
        for (int it = 0; it < kCandTile / 256; ++it) {
            const uint32_t i = base + it * 256 + tid;
            const uint32_t k = (i < N) ? __ldg(keys + i) : kInvalidKey;
            if (k == kInvalidKey) continue;
            const float h = key_float(k);
            uint32_t hit = 0;
#pragma unroll
            for (int u = 0; u < kMaxT; ++u) {
                below[u] += (h < lo[u]) ? 1u : 0u;
                hit |= (h >= lo[u] && h < hi[u]) ? (1u << u) : 0u;
            }
            if (hit) {
                const uint32_t pos = atomicAdd(&s_cnt, 1u);
                s_pix[pos] = i;
                s_hit[pos] = static_cast<uint8_t>(hit);
            }
        }
        __syncthreads();
        // ---- phase 2: exact (contract) entropy of the candidates, one per thread; the tile reserves its range of every
        // band's list with ONE global atomic (an atomic per candidate put ~1e5 same-address operations through L2: 40 us)
        const uint32_t nc = s_cnt;
        if (tid < kMaxT) { s_bcnt[tid] = 0; s_bpos[tid] = 0; }
        __syncthreads();
        for (uint32_t j = tid; j < nc; j += 256) {
            const uint32_t i = s_pix[j];
            uint32_t hit = s_hit[j];
            const uint32_t b = i / HW, p = i - b * HW;
            const float *x = logits + static_cast<size_t>(b) * C * HW + p;
            float v[C];
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = __ldg(x + static_cast<size_t>(c) * HW);
            const float e = entropy_of<C>(v);
            ent[i] = e;
            s_pix[j] = float_key(e);                       // the pixel index is no longer needed: keep the exact key instead
            while (hit) {
                const int u = __ffs(hit) - 1;
                hit &= hit - 1;
                atomicAdd(&s_bcnt[u], 1u);
            }
        }
        __syncthreads();
        if (tid < kMaxT && s_bcnt[tid]) s_bbase[tid] = atomicAdd(&st->cnt[tid], s_bcnt[tid]);
        __syncthreads();
        for (uint32_t j = tid; j < nc; j += 256) {
            uint32_t hit = s_hit[j];
            const uint32_t ek = s_pix[j];
            while (hit) {
                const int u = __ffs(hit) - 1;
                hit &= hit - 1;
                lists[static_cast<size_t>(u) * N + s_bbase[u] + atomicAdd(&s_bpos[u], 1u)] = ek;
            }
        }
        __syncthreads();
    }
    // one global atomic per band per BLOCK (all blocks hit the same addresses)
#pragma unroll
    for (int u = 0; u < kMaxT; ++u) {
        const uint32_t sm = static_cast<uint32_t>(warp_sum_i(static_cast<int>(below[u])));
        if (lane == 0 && sm) atomicAdd(&s_below[u], sm);
    }
    __syncthreads();
    if (tid < U && s_below[tid]) atomicAdd(&st->below[tid], s_below[tid]);
}


// ====================================================================== the whole chain as ONE cooperative kernel
// entropy -> percentile thresholds -> reliable/unreliable partition in a single persistent launch (two 768-thread CTAs per
// SM, cudaLaunchCooperativeKernel so that the software grid barrier below cannot deadlock).  The logits are streamed from
// HBM exactly once; per pixel a CTA keeps 16 bits in shared memory (the bin of its fast entropy in a fine histogram), the
// entropy map itself stays in L2-resident global memory and is touched again only for the ~1 % of pixels near a threshold.
// Measured history of the design (B200, V16 size, profiles/r02_chain_*): four launches 204 us -> this kernel 120 us; what
// mattered, in order: linear fine bins (candidates 1e5 -> 5e3), one list reservation per CTA instead of one global atomic
// per candidate (40 us), 2 B instead of 5 B of shared memory per pixel (L1 kept ~100 KB: streaming pass 111 -> 79 us),
// bitmap + 16-byte loads in the near scan (21 -> 4.5 us), bulk partition overlapped with the candidate warps.
constexpr int kChainThreads = 768;                        // two CTAs per SM: 1536 threads keep ~40 registers each, like entropy_fast_hist
constexpr int kChainCtasPerSm = 2;
constexpr int kChainMaxSlice = 24576;                     // pixels per CTA (capacity of the dynamic deal): 48 KB of 16-bit bins
// Fine histogram: LINEAR in the entropy value, 1024 bins per unit over [0, 4) (an entropy is at most ln C < 4 for
// C <= 54; the last bin is open-ended).  The error bound kDelta of the fast evaluation is absolute, so is the bin width:
// a band (bin + 3 kDelta either side) is ~1.6e-3 wide and holds a few thousand of the 4.2 M pixels, where 128 bins per
// OCTAVE made bands of 0.016 around the 90th percentile and ~1e5 candidates.
constexpr int kFineBins = 4096;
constexpr float kFineScale = 1024.0f;
constexpr int kFinePad = 4608;                            // histogram words: kChainThreads x 6 bins per thread in the scan
constexpr uint32_t kCandCap = 4096;                       // candidates compacted per round (16-bit index + band mask)
constexpr int kSelBins = 2048;                            // P3 radix digit: 11 bits

// monotone in h: bin b holds [b/1024, (b+1)/1024); bin 0 also takes everything below 0, the last bin everything above
__device__ __forceinline__ uint32_t fine_bin(float h)
{
    const float x = h * kFineScale;                        // exact (power of two)
    return x < 1.0f ? 0u : min(static_cast<uint32_t>(x), static_cast<uint32_t>(kFineBins - 1));
}
__device__ __forceinline__ float fine_lo(uint32_t b) { return static_cast<float>(b) * (1.0f / kFineScale); }

__device__ __forceinline__ void phase_stamp(SelState *st, int k)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        st->stamp[k] = t;
    }
}

__device__ __forceinline__ void grid_barrier(uint32_t *ctr, uint32_t goal)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        uint32_t v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < goal);
        __threadfence();
    }
    __syncthreads();
}

// block-wide exclusive scan position of `cnt` (one value per thread); warp_tot: one word per warp
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t cnt, uint32_t *warp_tot)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    __syncthreads();                                       // warp_tot free (previous use read)
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < wid; ++w) base += warp_tot[w];
    return base + inc - cnt;
}

// P1  pixel blocks of 768 pixels are dealt dynamically (atomic counter, no block barrier); per pixel: C coalesced loads in
//     flight, fast entropy (MUFU ex2/lg2) -> ent[], its bin in a LINEAR fine histogram (1024 bins per unit over [0,4))
//     -> s_bin[] (16 bit) and the CTA's shared-memory histogram, merged into the global one with atomics
// P2  every CTA (redundantly) scans the merged histogram: the bin of each target rank -- which also tells how many valid
//     pixels lie in LOWER bins, so nothing is counted per pixel -- and from the partition percentile's bins the range
//     outside of which a pixel's side of the cut is already certain.  All warps list the pixels whose bin can reach a
//     candidate band (bin +- 3 kDelta); then warps 0..7 read those entropies back, re-evaluate the candidates AT ONCE under
//     the arithmetic contract (exact value back into ent[]), reserve the CTA's range of each band's global list with ONE
//     atomic per band and append the exact keys (+ per-band min / max), while warps 8..23 write target_out / drop_mask
//     for the 99 % of pixels that are ignored or at least two bins away from the threshold's bins
// P3  CTA t mod G: exact select of target t inside its candidate list: 11-bit radix digits over the bits in which the
//     candidates differ (they share their leading ~15 bits; a digit they all share would serialise the shared atomics)
// P4  numpy lerp -> thresholds (every CTA, redundantly; CTA 0 publishes); the remaining ~1 % of pixels compare their stored
//     (exact) entropy with the threshold; kept count
// Three grid barriers replace four launches + three scans.  Soundness argument: see the two-level path (kDelta).
template <int C>
__global__ void __launch_bounds__(kChainThreads, kChainCtasPerSm)
entropy_chain_kernel(const float *__restrict__ logits, const int64_t *__restrict__ target_in, uint32_t HW, uint32_t N,
                     uint32_t slice, int64_t ignore, Percents pc, int nq, int part_idx, int dbg,
                     float *__restrict__ ent, float *__restrict__ thresh, int64_t *__restrict__ n_valid,
                     int64_t *__restrict__ target_out, uint8_t *__restrict__ drop_mask, unsigned long long *__restrict__ n_kept,
                     uint32_t *__restrict__ hist, SelState *__restrict__ st, uint32_t *__restrict__ lists)
{
    extern __shared__ uint32_t chain_smem[];
    // Per pixel the CTA keeps only its 16-bit fine-histogram bin (0xFFFF = ignored).  Full keys and class ids in shared
    // memory (5 B per pixel, 220 KB per SM) left ~6 KB of L1 and the streaming pass ran at 3.6 TB/s; with 2 B per pixel the
    // L1 keeps ~100 KB and the pass runs like the stand-alone entropy kernel.  Later phases read ent[] (L2) only for the
    // ~1 % of pixels whose bin is within one bin of a target / threshold, and re-read the labels once in P4.
    uint32_t *s_hist = chain_smem;                         // [kFinePad]: fine histogram; later the select histogram
    uint16_t *s_pix = reinterpret_cast<uint16_t *>(s_hist + kFinePad);   // [kCandCap] candidate slice-local index
    uint8_t *s_hit = reinterpret_cast<uint8_t *>(s_pix + kCandCap);      // [kCandCap] candidate band mask (kMaxT <= 8 bands)
    uint16_t *s_bin = reinterpret_cast<uint16_t *>(s_hit + kCandCap);    // [slice] fine bin of the pixel's FAST entropy
    __shared__ uint32_t warp_tot[32], s_grank[kMaxT], s_tbin[kMaxT], s_tcum[kMaxT], s_band[kMaxT];
    __shared__ uint32_t s_near[kFineBins / 32], s_bnmin[kMaxT], s_bmax[kMaxT], s_bbin[kMaxT], s_bcum[kMaxT], s_bcnt[kMaxT], s_blow[kMaxT], s_bbase[kMaxT], s_bpos[kMaxT];
    __shared__ float s_gamma[kMaxQ], s_lo[kMaxT], s_hi[kMaxT], s_binlo[kMaxT], s_thr[kMaxQ];
    __shared__ uint32_t s_nnear, s_n, s_cnt, s_sel_prefix, s_sel_rank, s_kmin, s_kmax, s_plo, s_phi, s_kept;
    __shared__ int s_nband;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int T = 2 * nq;
    const uint32_t G = gridDim.x;
    // Pixel blocks of kChainThreads pixels are handed out DYNAMICALLY (atomic counter): all CTAs sweep the same region of
    // every class plane at the same time, like a grid-stride kernel (contiguous per-CTA slices made 21 x 296 separate DRAM
    // streams), and an SM that gets less memory bandwidth simply takes fewer blocks (with a static round-robin deal the
    // slowest CTA finished its 19 blocks 30 us after the fastest).  A CTA keeps at most kMaxBlk blocks (its shared-memory
    // slice); slice-local index j = k * kChainThreads + tid for the k-th block it took.
    const uint32_t kMaxBlk = slice / kChainThreads;        // capacity (slice is a multiple of kChainThreads)
    const uint32_t total_blocks = (N + kChainThreads - 1) / kChainThreads;
    __shared__ uint32_t s_blk[kChainMaxSlice / kChainThreads + 1];
    auto global_of = [&](uint32_t j) { return s_blk[j / kChainThreads] * kChainThreads + (j % kChainThreads); };

    // ---------------------------------------------------------------- P1
    phase_stamp(st, 0);
    if (blockIdx.x == 0 && tid == 0 && n_kept) *n_kept = 0;     // every CTA adds its count after the third grid barrier
    unsigned long long *dbg_t = reinterpret_cast<unsigned long long *>(hist + 5120) + blockIdx.x * 2;   // (U2PL_CHAIN_TIMING) free words of hist2
    if (dbg && tid == 0) { unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); dbg_t[0] = t0; }
    for (int j = tid; j < kFinePad; j += kChainThreads) s_hist[j] = 0;
    __syncthreads();
    // Block ids travel through s_blk[] WITHOUT block barriers: warp 0 claims slot k + 2 when it starts its k-th block and
    // the other warps spin on a slot only if they run more than two blocks ahead of warp 0 (a barrier per block made all
    // 24 warps wait for the slowest one 19 times per CTA).  A claim beyond the last block ends the deal: warp 0 marks
    // every remaining slot.
    constexpr uint32_t kEmpty = 0xFFFFFFFFu;
    volatile uint32_t *v_blk = s_blk;
    for (int k2 = tid; k2 <= static_cast<int>(kMaxBlk); k2 += kChainThreads) s_blk[k2] = kEmpty;
    __syncthreads();
    auto claim = [&](uint32_t slot) {                      // warp 0, lane 0
        if (slot >= kMaxBlk) return;
        const uint32_t got = atomicAdd(&st->next_block, 1u);
        if (got >= total_blocks) { for (uint32_t z = slot; z < kMaxBlk; ++z) v_blk[z] = total_blocks; }
        else v_blk[slot] = got;
    };
    if (tid == 0) { claim(0); if (v_blk[0] < total_blocks) claim(1); }
    for (uint32_t k = 0; k < kMaxBlk; ++k) {
        uint32_t blk = 0;
        if (lane == 0) {
            if (wid == 0 && k + 2 < kMaxBlk && v_blk[k + 1] < total_blocks) claim(k + 2);   // (slot k + 1 is warp 0's own, already written)
            while ((blk = v_blk[k]) == kEmpty) { }
        }
        blk = __shfl_sync(0xffffffffu, blk, 0);
        if (blk >= total_blocks) break;
        const uint32_t j = k * kChainThreads + tid;
        const uint32_t i = blk * kChainThreads + tid;
        if (i >= N) {
            s_bin[j] = 0xFFFFu;
        } else {
            const uint32_t b = i / HW, p = i - b * HW;
            const float *x = logits + static_cast<size_t>(b) * C * HW + p;
            float v[C];
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = __ldg(x + static_cast<size_t>(c) * HW);
            const int64_t t = __ldg(target_in + i);
            const float h = entropy_fast_of<C>(v);
            ent[i] = h;
            const bool valid = (t != ignore);
            const uint32_t fb = fine_bin(h);
            s_bin[j] = valid ? static_cast<uint16_t>(fb) : static_cast<uint16_t>(0xFFFFu);
            if (valid && !(dbg & 2)) hist_add(s_hist, fb);
        }
    }
    __syncthreads();
    uint32_t nblk = 0;
    while (nblk < kMaxBlk && s_blk[nblk] < total_blocks) ++nblk;
    const uint32_t cnt = nblk * kChainThreads;             // local indices scanned by the later phases; pixels beyond N hold kInvalidKey
    __syncthreads();
    for (int j = tid; j < kFineBins; j += kChainThreads)
        if (s_hist[j]) atomicAdd(&hist[j], s_hist[j]);
    phase_stamp(st, 1);
    if (dbg && tid == 0) { unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); dbg_t[1] = t1; }
    grid_barrier(&st->bar, G);
    phase_stamp(st, 2);

    // ---------------------------------------------------------------- P2: bins of the target ranks, bands, candidates
    {
        // The merged histogram is read by all 296 CTAs at the same moment: staged with coalesced 16-byte loads (144 line
        // requests per CTA; six strided 4-byte loads per thread asked L2 for the same 144 lines 864 times per CTA).
        for (int k4 = tid; k4 < kFinePad / 4; k4 += kChainThreads)
            reinterpret_cast<uint4 *>(s_hist)[k4] = __ldcg(reinterpret_cast<const uint4 *>(hist) + k4);
        __syncthreads();
        uint32_t loc[6], sum = 0;                          // thread t owns bins 6t..6t+5 (the global array is padded to kFinePad)
#pragma unroll
        for (int k = 0; k < 6; ++k) { loc[k] = s_hist[tid * 6 + k]; sum += loc[k]; }
        const uint32_t excl = block_excl_scan(sum, warp_tot);
        if (tid == kChainThreads - 1) s_n = excl + sum;
        __syncthreads();
        const uint32_t n = s_n;
        if (tid == 0) {
            const float nm1 = __uint2float_rn(n ? n - 1u : 0u);
            for (int j = 0; j < nq; ++j) {                 // numpy 2.x float32 virtual index (see select1_kernel)
                const float q32 = __fdiv_rn(pc.q[j], 100.0f);
                const float v = __fmul_rn(nm1, q32);
                const float fl = floorf(v);
                uint32_t lo, hi;
                if (n == 0) { lo = hi = 0; }
                else if (v >= nm1) { lo = hi = n - 1u; }
                else { lo = static_cast<uint32_t>(fl); hi = lo + 1u; }
                s_gamma[j] = __fadd_rn(v, -fl);
                s_grank[2 * j] = lo;
                s_grank[2 * j + 1] = hi;
            }
        }
        __syncthreads();
        if (n != 0) {
            for (int t = 0; t < T; ++t) {
                const uint32_t r = s_grank[t];
                if (r >= excl && r < excl + sum) {
                    uint32_t cum = excl;
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        if (r < cum + loc[k]) { s_tbin[t] = tid * 6 + k; s_tcum[t] = cum; break; }
                        cum += loc[k];
                    }
                }
            }
        }
        __syncthreads();
    }
    const uint32_t n_all = s_n;                            // uniform across the grid (same histogram everywhere)
    if (tid == 0) {
        int nb = 0;
        if (n_all != 0) {                                  // merge targets with the same bin into one candidate band
            for (int t = 0; t < T; ++t) {
                int b = -1;
                for (int u = 0; u < t; ++u) if (s_tbin[u] == s_tbin[t]) { b = static_cast<int>(s_band[u]); break; }
                if (b < 0) {
                    b = nb++;
                    const uint32_t bin = s_tbin[t];
                    s_binlo[b] = (bin == 0) ? __uint_as_float(0xff800000u) : fine_lo(bin);
                    s_lo[b] = (bin == 0) ? __uint_as_float(0xff800000u) : fine_lo(bin) - 3.0f * kDelta;
                    s_hi[b] = (bin == kFineBins - 1) ? __uint_as_float(0x7f800000u) : fine_lo(bin + 1) + 3.0f * kDelta;
                    s_bcum[b] = s_tcum[t];                 // valid pixels in lower bins: fast value < s_binlo[b]
                    s_bbin[b] = bin;
                }
                s_band[t] = static_cast<uint32_t>(b);
            }
            if (blockIdx.x == 0) {
                st->n = n_all;
                for (int t = 0; t < T; ++t) { st->grank[t] = s_grank[t]; st->band[t] = s_band[t]; }
                for (int j = 0; j < nq; ++j) st->gamma[j] = s_gamma[j];
            }
            // The partition threshold lies between the exact order statistics of targets 2p and 2p+1, each within kDelta of
            // its bin: every pixel two or more bins outside [plo, phi] is decided now, before the threshold is known.
            s_plo = min(s_tbin[2 * part_idx], s_tbin[2 * part_idx + 1]);
            s_phi = max(s_tbin[2 * part_idx], s_tbin[2 * part_idx + 1]);
        } else {
            s_plo = 0; s_phi = 0;
        }
        s_nband = nb;
        s_kept = 0;
        s_nnear = 0;
        for (int k = 0; k < kFineBins / 32; ++k) s_near[k] = 0;
        for (int u = 0; u < nb; ++u)
            for (uint32_t bb = (s_bbin[u] ? s_bbin[u] - 1u : 0u); bb <= min(s_bbin[u] + 1u, static_cast<uint32_t>(kFineBins - 1)); ++bb)
                s_near[bb >> 5] |= 1u << (bb & 31u);
    }
    __syncthreads();
    if (dbg) phase_stamp(st, 8);
    // Near scan by ALL warps: eight bins per 16-byte shared-memory load, one bitmap probe each (s_near: the bins a band can
    // reach -- its own and the two neighbours, 3 kDelta < bin width).  The few hundred slice-local indices it finds go to a
    // list (the histogram staging area is free again); only those pixels are looked at by the candidate warps below.
    uint16_t *s_nearlist = reinterpret_cast<uint16_t *>(s_hist);
    constexpr uint32_t kNearCap = kCandCap;                // so that the candidates of a list round always fit
    if (s_nband > 0) {
        for (uint32_t j0 = tid * 8; j0 < cnt; j0 += kChainThreads * 8) {
            const uint4 pk = *reinterpret_cast<const uint4 *>(s_bin + j0);
            const uint32_t wv[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint32_t fb = (wv[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
                if (fb >= static_cast<uint32_t>(kFineBins)) continue;                 // ignored pixel
                if (!((s_near[fb >> 5] >> (fb & 31u)) & 1u)) continue;
                const uint32_t pos = atomicAdd(&s_nnear, 1u);
                if (pos < kNearCap) s_nearlist[pos] = static_cast<uint16_t>(j0 + e);
            }
        }
    }
    __syncthreads();
    constexpr int kCandThreads = 256;                      // warps 0..7: candidates; warps 8..23: bulk partition
    if (tid < kCandThreads) {
        const int U = s_nband;
        const uint32_t nnear = s_nnear;
        const bool use_list = nnear <= kNearCap;           // else (massive ties): scan the slice in bounded rounds, as below
        // Candidates are compacted (16-bit slice-local index + band mask) and then re-evaluated all at once, one per
        // thread, so the ~3 us latency of an exact evaluation (C strided loads + ~900 dependent issue slots) is paid once.
        // A slice holding more than kCandCap candidates (massive ties) is processed in several rounds of kCandCap keys.
        bool bounded = false;
        for (uint32_t tb = 0; U > 0 && tb < cnt && !(dbg & 8); ) {       // (dbg & 8: timing experiment without the candidate work -- wrong results)
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) s_cnt = 0;
            if (tid < kMaxT) { s_bcnt[tid] = 0; s_blow[tid] = 0; s_bpos[tid] = 0; s_bnmin[tid] = 0; s_bmax[tid] = 0; }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const uint32_t te = (bounded && !use_list) ? min(cnt, tb + kCandCap) : cnt;
            auto test_pixel = [&](uint32_t j) {
                const float h = __ldcg(ent + global_of(j));
                uint32_t hit = 0;
                for (int u = 0; u < U; ++u) hit |= (h >= s_lo[u] && h < s_hi[u]) ? (1u << u) : 0u;
                if (hit) {
                    const uint32_t pos = atomicAdd(&s_cnt, 1u);
                    if (pos < kCandCap) { s_pix[pos] = static_cast<uint16_t>(j); s_hit[pos] = static_cast<uint8_t>(hit); }
                }
            };
            if (use_list) {
                for (uint32_t k2 = tid; k2 < nnear; k2 += kCandThreads) test_pixel(s_nearlist[k2]);
            } else {
                for (uint32_t j0 = tb + tid * 8; j0 < te; j0 += kCandThreads * 8) {
                    const uint4 pk = *reinterpret_cast<const uint4 *>(s_bin + j0);
                    const uint32_t wv[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const uint32_t fb = (wv[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
                        if (fb >= static_cast<uint32_t>(kFineBins)) continue;
                        if (!((s_near[fb >> 5] >> (fb & 31u)) & 1u)) continue;
                        test_pixel(j0 + e);
                    }
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (dbg) phase_stamp(st, 6);
            const uint32_t nc = s_cnt;
            if (nc > kCandCap) { bounded = true; continue; }   // did not fit (only possible for a whole-slice scan): redo in bounded rounds
            for (uint32_t c2 = tid; c2 < nc; c2 += kCandThreads) {        // exact (contract) entropy of the candidates
                const uint32_t j = s_pix[c2];
                uint32_t hit = s_hit[c2];
                const uint32_t i = global_of(j);
                const float hfast = __ldcg(ent + i);
                const uint32_t b = i / HW, p = i - b * HW;
                const float *x = logits + static_cast<size_t>(b) * C * HW + p;
                float v[C];
#pragma unroll
                for (int c = 0; c < C; ++c) v[c] = __ldg(x + static_cast<size_t>(c) * HW);
                const float e = entropy_of<C>(v);
                __stcg(ent + i, e);
                const uint32_t ek0 = float_key(e);
                while (hit) {
                    const int u = __ffs(hit) - 1;
                    hit &= hit - 1;
                    atomicAdd(&s_bcnt[u], 1u);
                    atomicMax(&s_bnmin[u], ~ek0);
                    atomicMax(&s_bmax[u], ek0);
                    if (hfast < s_binlo[u]) atomicAdd(&s_blow[u], 1u);     // counted in s_bcum although it is a candidate
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid < U && s_bcnt[tid]) {                  // ONE global atomic per band per CTA reserves the list range
                s_bbase[tid] = atomicAdd(&st->cnt[tid], s_bcnt[tid]);
                if (s_blow[tid]) atomicAdd(&st->below[tid], s_blow[tid]);
                atomicMax(&st->nkmin[tid], s_bnmin[tid]);
                atomicMax(&st->kmax[tid], s_bmax[tid]);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (uint32_t c2 = tid; c2 < nc; c2 += kCandThreads) {
                const uint32_t j = s_pix[c2];
                uint32_t hit = s_hit[c2];
                const uint32_t ek = float_key(__ldcg(ent + global_of(j)));    // this thread's own store of the first pass
                while (hit) {
                    const int u = __ffs(hit) - 1;
                    hit &= hit - 1;
                    lists[static_cast<size_t>(u) * N + s_bbase[u] + atomicAdd(&s_bpos[u], 1u)] = ek;
                }
            }
            tb = te;
        }
        if (dbg) phase_stamp(st, 7);
    } else if (target_out != nullptr && !(dbg & 4)) {      // (dbg & 4: timing experiment without the bulk partition -- wrong results)
        // Bulk partition, overlapped with the latency-bound candidate work of warps 0..7: every pixel that is ignored or
        // lies two or more bins away from the threshold's bins (99 % of them) gets its output now.
        const uint32_t plo = s_plo, phi = s_phi;
        const int bt = tid - kCandThreads, nbt = kChainThreads - kCandThreads;
        int kept = 0;
        for (uint32_t j0 = bt; j0 < cnt; j0 += 4 * nbt) {  // four independent label loads in flight per thread
            uint32_t ii[4];
            int dec[4];                                    // 0 = ignored / dropped, 1 = kept, 2 = not decided here
            int64_t lab[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t j = j0 + e * nbt;
                dec[e] = 2; ii[e] = 0; lab[e] = ignore;
                if (j < cnt) {
                    ii[e] = global_of(j);
                    const uint32_t fb = s_bin[j];
                    if (ii[e] >= N) dec[e] = 2;
                    else if (fb == 0xFFFFu) dec[e] = 3;    // ignored pixel: output `ignore`, not counted as dropped
                    else if (fb > phi + 1u) dec[e] = 0;
                    else if (fb + 1u < plo) dec[e] = 1;
                }
                if (dec[e] == 1) lab[e] = __ldg(target_in + ii[e]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (dec[e] == 2) continue;
                target_out[ii[e]] = lab[e];
                if (drop_mask) drop_mask[ii[e]] = dec[e] == 0 ? 1 : 0;
                kept += dec[e] == 1 ? 1 : 0;
            }
        }
        kept = warp_sum_i(kept);
        if (lane == 0 && kept) atomicAdd(&s_kept, static_cast<uint32_t>(kept));
    }
    grid_barrier(&st->bar, 2 * G);
    phase_stamp(st, 3);

    // ---------------------------------------------------------------- P3: exact select, target t on CTA t mod G
    for (int t = blockIdx.x; n_all != 0 && t < T; t += static_cast<int>(G)) {       // (grids smaller than T loop)
        uint32_t *sel = s_hist;                            // [kSelBins]
        __syncthreads();
        const uint32_t u = s_band[t];
        const uint32_t n = __ldcg(&st->cnt[u]);
        // rank among the candidates: pixels surely below = those in lower bins that are not candidates themselves
        const uint32_t r0 = s_grank[t] - (s_bcum[u] - __ldcg(&st->below[u]));      // unsigned: a violated invariant shows as r0 >= n
        const uint32_t *list = lists + static_cast<size_t>(u) * N;
        if (r0 >= n) {
            if (tid == 0) st->val[t] = __uint_as_float(0x7fc00000u);
            continue;
        }
        if (tid == 0) { s_kmin = ~__ldcg(&st->nkmin[u]); s_kmax = __ldcg(&st->kmax[u]); s_sel_prefix = 0; s_sel_rank = r0; }   // (min / max gathered in P2)
        __syncthreads();
        const uint32_t diff = s_kmin ^ s_kmax;
        const int nb = diff ? 32 - __clz(diff) : 0;        // the candidates agree on their leading 32 - nb bits
        const uint32_t low_mask = (nb == 32) ? 0xffffffffu : ((1u << nb) - 1u);
        for (int done = 0; done < nb; ) {
            const int w = min(11, nb - done), sh = nb - done - w;
            for (int j = tid; j < kSelBins; j += kChainThreads) sel[j] = 0;
            __syncthreads();
            const uint32_t pre = s_sel_prefix, r = s_sel_rank;
            const uint32_t dmask = (1u << w) - 1u;
            for (uint32_t j0 = tid; j0 < n; j0 += 4 * kChainThreads) {
                uint32_t kk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) kk[e] = (j0 + e * kChainThreads < n) ? __ldcg(list + j0 + e * kChainThreads) : 0u;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t k = kk[e] & low_mask;
                    if (j0 + e * kChainThreads < n && (done == 0 || (k >> (sh + w)) == pre)) hist_add(sel, (k >> sh) & dmask);
                }
            }
            __syncthreads();
            uint32_t loc[3], sum = 0;                      // thread t owns bins 3t..3t+2 (768 x 3 >= 2048)
#pragma unroll
            for (int k = 0; k < 3; ++k) { loc[k] = (tid * 3 + k < kSelBins) ? sel[tid * 3 + k] : 0u; sum += loc[k]; }
            const uint32_t excl = block_excl_scan(sum, warp_tot);
            __syncthreads();
            if (r >= excl && r < excl + sum) {
                uint32_t cum = excl;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (r >= cum && r < cum + loc[k]) { s_sel_prefix = (pre << w) | static_cast<uint32_t>(tid * 3 + k); s_sel_rank = r - cum; }
                    cum += loc[k];
                }
            }
            __syncthreads();
            done += w;
        }
        if (tid == 0) st->val[t] = key_float((s_kmin & ~low_mask) | s_sel_prefix);
    }
    grid_barrier(&st->bar, 3 * G);
    phase_stamp(st, 4);

    // ---------------------------------------------------------------- P4: thresholds + partition of the slice
    if (tid == 0) {
        for (int j = 0; j < nq; ++j) {
            float r = __uint_as_float(0x7fc00000u);
            if (n_all != 0) {                              // numpy _lerp (see select_refine_kernel)
                const float a = __ldcg(&st->val[2 * j]), b = __ldcg(&st->val[2 * j + 1]), g = s_gamma[j];
                const float d = __fadd_rn(b, -a);
                r = __fadd_rn(a, __fmul_rn(d, g));
                if (g >= 0.5f) r = __fadd_rn(b, -__fmul_rn(d, __fadd_rn(1.0f, -g)));
            }
            s_thr[j] = r;
            if (blockIdx.x == 0) thresh[j] = r;
        }
        if (blockIdx.x == 0 && n_valid) *n_valid = static_cast<int64_t>(n_all);
    }
    __syncthreads();
    if (target_out != nullptr) {
        const float th = s_thr[part_idx];
        const bool th_nan = !(th == th);                   // violated invariant (thresholds are NaN): nothing more is dropped
        const uint32_t plo = s_plo, phi = s_phi;
        int kept = 0;
        for (uint32_t j = tid; n_all != 0 && j < cnt; j += kChainThreads) {
            const uint32_t fb = s_bin[j];
            if (fb == 0xFFFFu || fb > phi + 1u || fb + 1u < plo) continue;       // decided in P2
            const uint32_t i = global_of(j);
            if (i >= N) continue;
            const bool drop = !th_nan && (__ldcg(ent + i) >= th);                // the stored value is exact for every candidate
            target_out[i] = drop ? ignore : __ldg(target_in + i);
            if (drop_mask) drop_mask[i] = drop ? 1 : 0;
            kept += drop ? 0 : 1;
        }
        kept = warp_sum_i(kept);
        if (lane == 0 && kept) atomicAdd(&s_kept, static_cast<uint32_t>(kept));
        __syncthreads();
        if (tid == 0 && s_kept) atomicAdd(n_kept, static_cast<unsigned long long>(s_kept));
    }
    phase_stamp(st, 5);
}

// ------------------------------------------------------------------ OHEM (loss_helper.py:502-531)
// mask_prob = softmax(pred)[target] (1.0 where target == ignore), computed under the arithmetic
// contract because the kept set is cut by an order statistic of it (:520-524).
template <int C>
__global__ void __launch_bounds__(256)
ohem_prob_hist_kernel(const float *__restrict__ logits, const int64_t *__restrict__ target,
                      uint32_t HW, uint32_t N, int64_t ignore, uint32_t *__restrict__ keys,
                      uint32_t *__restrict__ hist1, unsigned long long *__restrict__ n_valid)
{
    __shared__ uint32_t sh[kBins1];
    for (int j = threadIdx.x; j < kBins1; j += 256) sh[j] = 0;
    __syncthreads();
    int nv = 0;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < N; i += gridDim.x * 256u) {
        const int64_t t = __ldg(target + i);
        float mp = 1.0f;                                            // :516 masked_fill_(~valid, 1)
        if (t != ignore) {
            ++nv;
            const uint32_t b = i / HW, p = i - b * HW;
            const float *x = logits + static_cast<size_t>(b) * C * HW + p;
            float v[C];
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = __ldg(x + static_cast<size_t>(c) * HW);
            float m = v[0];
#pragma unroll
            for (int c = 1; c < C; ++c) m = fmaxf(m, v[c]);
            float S = 0.0f, et = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float e = det_expf(__fadd_rn(v[c], -m));
                S = __fadd_rn(S, e);
                et = (c == static_cast<int>(t)) ? e : et;
            }
            mp = __fmul_rn(et, __fdiv_rn(1.0f, S));
        }
        const uint32_t key = float_key(mp);
        keys[i] = key;
        hist_add(sh, key >> 20);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < kBins1; j += 256)
        if (sh[j]) atomicAdd(&hist1[j], sh[j]);
    nv = warp_sum_i(nv);
    if ((threadIdx.x & 31) == 0 && nv) atomicAdd(n_valid, static_cast<unsigned long long>(nv));
}

__global__ void __launch_bounds__(256)
ohem_prob_hist_kernel_anyC(const float *__restrict__ logits, const int64_t *__restrict__ target, uint32_t C,
                           uint32_t HW, uint32_t N, int64_t ignore, uint32_t *__restrict__ keys,
                           uint32_t *__restrict__ hist1, unsigned long long *__restrict__ n_valid)
{
    __shared__ uint32_t sh[kBins1];
    for (int j = threadIdx.x; j < kBins1; j += 256) sh[j] = 0;
    __syncthreads();
    int nv = 0;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < N; i += gridDim.x * 256u) {
        const int64_t t = __ldg(target + i);
        float mp = 1.0f;
        if (t != ignore) {
            ++nv;
            const uint32_t b = i / HW, p = i - b * HW;
            const float *x = logits + static_cast<size_t>(b) * C * HW + p;
            float m = __ldg(x);
            for (uint32_t c = 1; c < C; ++c) m = fmaxf(m, __ldg(x + static_cast<size_t>(c) * HW));
            float S = 0.0f, et = 0.0f;
            for (uint32_t c = 0; c < C; ++c) {
                const float e = det_expf(__fadd_rn(__ldg(x + static_cast<size_t>(c) * HW), -m));
                S = __fadd_rn(S, e);
                et = (c == static_cast<uint32_t>(t)) ? e : et;
            }
            mp = __fmul_rn(et, __fdiv_rn(1.0f, S));
        }
        const uint32_t key = float_key(mp);
        keys[i] = key;
        hist_add(sh, key >> 20);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < kBins1; j += 256)
        if (sh[j]) atomicAdd(&hist1[j], sh[j]);
    nv = warp_sum_i(nv);
    if ((threadIdx.x & 31) == 0 && nv) atomicAdd(n_valid, static_cast<unsigned long long>(nv));
}

__global__ void __launch_bounds__(256)
ohem_partition_kernel(const uint32_t *__restrict__ keys, const int64_t *__restrict__ target, uint32_t N,
                      int64_t ignore, float thresh, long long min_kept, const float *__restrict__ kth,
                      const unsigned long long *__restrict__ n_valid, int64_t *__restrict__ new_target)
{
    const long long nv = static_cast<long long>(*n_valid);
    const bool filter = !(min_kept > nv) && nv > 0 && min_kept > 0;       // :512-519
    const float k = __ldg(kth);
    const float threshold = (k > thresh) ? k : thresh;                      // :518,522-523
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < N; i += gridDim.x * 256u) {
        const int64_t t = __ldg(target + i);
        bool keep = (t != ignore);
        if (keep && filter) keep = key_float(__ldg(keys + i)) <= threshold; // :524-526
        new_target[i] = keep ? t : ignore;                                   // :528
    }
}

// ------------------------------------------------------------------ host side
struct EntropyWs {
    uint32_t *keys, *hist1, *hist2, *hist3;
    SelState *st;
    size_t zero_bytes;          // hist1..st are contiguous, zeroed per call
};

static size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

static size_t ws_layout(int64_t N, void *base, EntropyWs *out)
{
    size_t off = 0;
    const size_t keys_b = align256(static_cast<size_t>(N) * 4 + 16);
    const size_t h1 = align256(kBins1 * 4), hr = align256(kMaxT * kBinsR * 4), stb = align256(sizeof(SelState));
    if (out) {
        char *p = static_cast<char *>(base);
        out->keys = reinterpret_cast<uint32_t *>(p + off);
        out->hist1 = reinterpret_cast<uint32_t *>(p + keys_b);
        out->hist2 = reinterpret_cast<uint32_t *>(p + keys_b + h1);
        out->hist3 = reinterpret_cast<uint32_t *>(p + keys_b + h1 + hr);
        out->st = reinterpret_cast<SelState *>(p + keys_b + h1 + 2 * hr);
        out->zero_bytes = h1 + 2 * hr + stb;
    }
    off = keys_b + h1 + 2 * hr + stb;
    return off;
}

static int grid_for(const void *kernel, int threads, size_t smem, uint32_t N, int per_thread = 1)
{
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (per_sm < 1) per_sm = 1;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long need = (static_cast<long long>(N) + threads * per_thread - 1) / (threads * per_thread);
    const long long cap = static_cast<long long>(sms) * per_sm;       // one resident wave, grid-stride
    return static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
}

template <int C>
static void launch_entropy(const float *logits, const int64_t *target, uint32_t HW, uint32_t N,
                           int64_t ignore, float *ent, const EntropyWs &w, cudaStream_t s)
{
    const int grid = grid_for(reinterpret_cast<const void *>(entropy_hist_kernel<C>), kEntThreads, 0, N);
    entropy_hist_kernel<C><<<grid, kEntThreads, 0, s>>>(logits, target, HW, N, ignore, ent, w.keys, w.hist1);
}

}  // namespace u2pl

using namespace u2pl;

extern "C" size_t u2pl_entropy_ws_bytes(int64_t B, int64_t HW)
{
    return ws_layout(B * HW, nullptr, nullptr);
}

extern "C" int u2pl_entropy_thresholds(const float *logits, const int64_t *target,
                                       int64_t B, int64_t C, int64_t HW, int64_t ignore,
                                       const float *h_percents, int nq,
                                       float *entropy, float *thresh, int64_t *n_valid,
                                       void *ws, size_t ws_bytes, void *stream)
{
    if (B <= 0 || C <= 0 || HW <= 0) return bad_arg("entropy_thresholds: empty shape");
    if (nq < 1 || nq > kMaxQ) return bad_arg("entropy_thresholds: nq must be in [1, U2PL_MAX_QUANTILES]");
    if (B * HW >= (1LL << 31) || B * C * HW >= (1LL << 40)) return bad_arg("entropy_thresholds: B*HW must be < 2^31");
    if (ws_bytes < ws_layout(B * HW, nullptr, nullptr)) { set_error("entropy_thresholds: workspace too small"); return U2PL_E_WS_SMALL; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t N = static_cast<uint32_t>(B * HW), hw = static_cast<uint32_t>(HW);
    EntropyWs w;
    ws_layout(B * HW, ws, &w);
    cudaError_t e = cudaMemsetAsync(w.hist1, 0, w.zero_bytes, s);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
    Percents pc;
    pc.use_rank = 0; pc.rank = 0;
    for (int j = 0; j < kMaxQ; ++j) pc.q[j] = (j < nq) ? h_percents[j] : 0.0f;

    switch (C) {
        case 19: launch_entropy<19>(logits, target, hw, N, ignore, entropy, w, s); break;
        case 21: launch_entropy<21>(logits, target, hw, N, ignore, entropy, w, s); break;
        default: {
            const int grid = grid_for(reinterpret_cast<const void *>(entropy_hist_kernel_anyC), 256, 0, N);
            entropy_hist_kernel_anyC<<<grid, 256, 0, s>>>(logits, target, static_cast<uint32_t>(C), hw, N,
                                                          ignore, entropy, w.keys, w.hist1);
        }
    }
    int rc = check_launch("entropy_hist");
    if (rc) return rc;
    select1_kernel<<<1, 256, 0, s>>>(w.hist1, w.st, pc, nq);
    const int T = 2 * nq;
    const size_t smem = static_cast<size_t>(T) * kBinsR * 4;
    const int g2 = grid_for(reinterpret_cast<const void *>(hist_refine_kernel<2>), 256, smem, N, 4);
    hist_refine_kernel<2><<<g2, 256, smem, s>>>(w.keys, N, w.st, w.hist2, T);
    select_refine_kernel<false><<<1, 1024, 0, s>>>(w.hist2, w.st, nq, thresh, n_valid);
    hist_refine_kernel<3><<<g2, 256, smem, s>>>(w.keys, N, w.st, w.hist3, T);
    select_refine_kernel<true><<<1, 1024, 0, s>>>(w.hist3, w.st, nq, thresh, n_valid);
    return check_launch("entropy_thresholds select chain", 5);
}


static size_t fast_ws_layout(int64_t N, void *base, EntropyWs *out, uint32_t **lists)
{
    const size_t head = ws_layout(N, base, out);
    if (lists) *lists = reinterpret_cast<uint32_t *>(static_cast<char *>(base) + head);
    return head + align256(static_cast<size_t>(kMaxT) * static_cast<size_t>(N) * 4);
}

extern "C" size_t u2pl_entropy_fast_ws_bytes(int64_t B, int64_t HW) { return fast_ws_layout(B * HW, nullptr, nullptr, nullptr); }

extern "C" int u2pl_entropy_thresholds_fast(const float *logits, const int64_t *target,
                                            int64_t B, int64_t C, int64_t HW, int64_t ignore,
                                            const float *h_percents, int nq,
                                            float *entropy, float *thresh, int64_t *n_valid,
                                            void *ws, size_t ws_bytes, void *stream)
{
    if (C != 19 && C != 21)      // the two-level path is specialised for the class counts of the shipped configs
        return u2pl_entropy_thresholds(logits, target, B, C, HW, ignore, h_percents, nq, entropy, thresh, n_valid, ws, ws_bytes, stream);
    if (B <= 0 || HW <= 0) return bad_arg("entropy_thresholds_fast: empty shape");
    if (nq < 1 || nq > kMaxQ) return bad_arg("entropy_thresholds_fast: nq must be in [1, U2PL_MAX_QUANTILES]");
    if (B * HW >= (1LL << 31)) return bad_arg("entropy_thresholds_fast: B*HW must be < 2^31");
    if (ws_bytes < fast_ws_layout(B * HW, nullptr, nullptr, nullptr)) { set_error("entropy_thresholds_fast: workspace too small"); return U2PL_E_WS_SMALL; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t N = static_cast<uint32_t>(B * HW), hw = static_cast<uint32_t>(HW);
    EntropyWs w;
    uint32_t *lists = nullptr;
    fast_ws_layout(B * HW, ws, &w, &lists);
    cudaError_t e = cudaMemsetAsync(w.hist1, 0, w.zero_bytes, s);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
    Percents pc;
    pc.use_rank = 0; pc.rank = 0;
    for (int j = 0; j < kMaxQ; ++j) pc.q[j] = (j < nq) ? h_percents[j] : 0.0f;
    const int T = 2 * nq;
    if (C == 19) {
        const int g = grid_for(reinterpret_cast<const void *>(entropy_fast_hist_kernel<19>), kEntThreads, 0, N);
        entropy_fast_hist_kernel<19><<<g, kEntThreads, 0, s>>>(logits, target, hw, N, ignore, entropy, w.keys, w.hist1);
    } else {
        const int g = grid_for(reinterpret_cast<const void *>(entropy_fast_hist_kernel<21>), kEntThreads, 0, N);
        entropy_fast_hist_kernel<21><<<g, kEntThreads, 0, s>>>(logits, target, hw, N, ignore, entropy, w.keys, w.hist1);
    }
    const size_t smem = static_cast<size_t>(T) * kBinsR * 4;
    fast_refine_kernel<<<4 * kNumSMs, 256, smem, s>>>(w.keys, N, w.hist1, w.st, pc, nq, w.hist2);
    const int gc = static_cast<int>(std::min<long long>((static_cast<long long>(N) + kCandTile - 1) / kCandTile, 4LL * kNumSMs));
    if (C == 19) fast_candidate_kernel<19><<<gc, 256, 0, s>>>(logits, w.keys, hw, N, w.hist2, w.st, T, entropy, lists);
    else         fast_candidate_kernel<21><<<gc, 256, 0, s>>>(logits, w.keys, hw, N, w.hist2, w.st, T, entropy, lists);
    exact_select_kernel<<<T, 1024, 0, s>>>(lists, N, w.st, nq, thresh, n_valid);
    return check_launch("entropy_thresholds_fast", 4);
}

extern "C" int u2pl_partition_target(const float *entropy, int64_t *target, int64_t n, int64_t ignore,
                                     const float *thresh, int thresh_idx,
                                     uint8_t *drop_mask, int64_t *n_kept, void *stream)
{
    if (n <= 0 || n >= (1LL << 31)) return bad_arg("partition_target: n must be in (0, 2^31)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(n_kept, 0, sizeof(int64_t), s);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
    const int grid = grid_for(reinterpret_cast<const void *>(partition_kernel), 256, 0, static_cast<uint32_t>(n), 2);
    partition_kernel<<<grid, 256, 0, s>>>(entropy, target, static_cast<uint32_t>(n), ignore, thresh, thresh_idx,
                                          drop_mask, reinterpret_cast<unsigned long long *>(n_kept));
    return check_launch("partition");
}

// One cooperative launch for the whole chain (entropy_chain_kernel); falls back to the multi-launch two-level path +
// partition_kernel when the pixel slice of a CTA would not fit in shared memory or the class count is not specialised.
template <int C>
static int launch_chain(const float *logits, const int64_t *target_in, uint32_t hw, uint32_t N, int64_t ignore, const Percents &pc,
                        int nq, int part_idx, float *entropy, float *thresh, int64_t *n_valid, int64_t *target_out,
                        uint8_t *drop_mask, int64_t *n_kept, const EntropyWs &w, uint32_t *lists, cudaStream_t s, bool *launched)
{
    *launched = false;
    int dev = 0, sms = kNumSMs, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) return 0;
    uint32_t grid = static_cast<uint32_t>(std::min<long long>(static_cast<long long>(kChainCtasPerSm) * sms,
                                                               std::max<long long>(1, (static_cast<long long>(N) + 4095) / 4096)));
    const uint32_t blocks = (N + kChainThreads - 1) / kChainThreads;       // pixel blocks, dealt round-robin to the CTAs
    if (grid > blocks) grid = blocks;
    // capacity of a CTA: its even share + 25 % (dynamic deal), at most kChainMaxSlice
    uint32_t need = (blocks + grid - 1) / grid;
    uint32_t slice = std::min<uint32_t>((need + (need + 3) / 4 + 1) * kChainThreads, (kChainMaxSlice / kChainThreads) * kChainThreads);
    if (static_cast<unsigned long long>(slice / kChainThreads) * grid < blocks) return 0;       // would not cover the tensor
    if (slice > static_cast<uint32_t>(kChainMaxSlice)) return 0;
    auto smem_for = [](size_t sl) { return kFinePad * 4 + kCandCap * 3 + sl * 2; };
    const size_t smem = smem_for(slice);
    static bool configured = false;
    if (!configured) {
        const size_t max_smem = smem_for(kChainMaxSlice);
        cudaError_t e = cudaFuncSetAttribute(entropy_chain_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(max_smem));
        if (e != cudaSuccess) { cudaGetLastError(); return 0; }
        configured = true;
    }
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, entropy_chain_kernel<C>, kChainThreads, smem);
    if (per_sm < 1 || static_cast<long long>(per_sm) * sms < grid) return 0;
    Percents pcc = pc;
    unsigned long long *nk = reinterpret_cast<unsigned long long *>(n_kept);
    uint32_t hw_ = hw, N_ = N;
    static const int dbg = [] { const char *e = getenv("U2PL_CHAIN_TIMING"); return e ? atoi(e) : 0; }();
    int dbg_ = dbg;
    void *args[] = {&logits, &target_in, &hw_, &N_, &slice, &ignore, &pcc, &nq, &part_idx, &dbg_, &entropy, &thresh, &n_valid,
                    &target_out, &drop_mask, &nk, const_cast<uint32_t **>(&w.hist1),      // hist1: 4096 words, zeroed: the fine histogram
                    const_cast<SelState **>(&w.st), &lists};
    cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(entropy_chain_kernel<C>), dim3(grid), dim3(kChainThreads),
                                                args, smem, s);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
    *launched = true;
    static const bool timing = getenv("U2PL_CHAIN_TIMING") != nullptr;
    if (timing) {                                          // debug only: synchronises the stream
        SelState h;
        cudaStreamSynchronize(s);
        cudaMemcpy(&h, w.st, sizeof(SelState), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[entropy_chain] P2 of CTA 0: histogram scan %.1f us | band test %.1f us | exact evaluation + list append %.1f us | wait at barrier 2 %.1f us\n",
                (h.stamp[8] - h.stamp[2]) * 1e-3, (h.stamp[6] - h.stamp[8]) * 1e-3, (h.stamp[7] - h.stamp[6]) * 1e-3, (h.stamp[3] - h.stamp[7]) * 1e-3);
        fprintf(stderr, "[entropy_chain] grid %u slice %u | P1 %.1f us | bar1 %.1f | P2+bar2 %.1f | P3+bar3 %.1f | P4 %.1f | total %.1f us (CTA 0)\n",
                grid, slice, (h.stamp[1] - h.stamp[0]) * 1e-3, (h.stamp[2] - h.stamp[1]) * 1e-3, (h.stamp[3] - h.stamp[2]) * 1e-3,
                (h.stamp[4] - h.stamp[3]) * 1e-3, (h.stamp[5] - h.stamp[4]) * 1e-3, (h.stamp[5] - h.stamp[0]) * 1e-3);
        static unsigned long long tt[2 * 1024];
        if (grid <= 1024) {
            cudaMemcpy(tt, w.hist1 + 5120, grid * 16, cudaMemcpyDeviceToHost);
            unsigned long long s0 = ~0ull, s1 = 0, e0 = ~0ull, e1 = 0; double dsum = 0, dmax = 0, dmin = 1e30;
            for (uint32_t b = 0; b < grid; ++b) {
                s0 = std::min(s0, tt[2 * b]); s1 = std::max(s1, tt[2 * b]); e0 = std::min(e0, tt[2 * b + 1]); e1 = std::max(e1, tt[2 * b + 1]);
                const double d = (tt[2 * b + 1] - tt[2 * b]) * 1e-3; dsum += d; dmax = std::max(dmax, d); dmin = std::min(dmin, d);
            }
            fprintf(stderr, "[entropy_chain] P1 per CTA: start spread %.1f us, end spread %.1f us (first end %.1f, last end %.1f after first start), duration min %.1f avg %.1f max %.1f us\n",
                    (s1 - s0) * 1e-3, (e1 - e0) * 1e-3, (e0 - s0) * 1e-3, (e1 - s0) * 1e-3, dmin, dsum / grid, dmax);
        }
    }
    return check_launch("entropy_chain");
}

extern "C" int u2pl_entropy_partition_fused(const float *logits, const int64_t *target_in,
                                            int64_t B, int64_t C, int64_t HW, int64_t ignore,
                                            const float *h_percents, int nq, int part_idx,
                                            float *entropy, float *thresh, int64_t *n_valid,
                                            int64_t *target_out, uint8_t *drop_mask, int64_t *n_kept,
                                            void *ws, size_t ws_bytes, void *stream)
{
    if (B <= 0 || C <= 0 || HW <= 0) return bad_arg("entropy_partition_fused: empty shape");
    if (nq < 1 || nq > kMaxQ || part_idx < 0 || part_idx >= nq) return bad_arg("entropy_partition_fused: bad nq / part_idx");
    if (B * HW >= (1LL << 31)) return bad_arg("entropy_partition_fused: B*HW must be < 2^31");
    if (!target_out || !n_kept) return bad_arg("entropy_partition_fused: target_out and n_kept are required");
    if (ws_bytes < fast_ws_layout(B * HW, nullptr, nullptr, nullptr)) { set_error("entropy_partition_fused: workspace too small"); return U2PL_E_WS_SMALL; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t N = static_cast<uint32_t>(B * HW), hw = static_cast<uint32_t>(HW);
    static const bool disabled = [] { const char *e = getenv("U2PL_ENTROPY_CHAIN"); return e && e[0] == '0'; }();
    if ((C == 19 || C == 21) && !disabled) {
        EntropyWs w;
        uint32_t *lists = nullptr;
        fast_ws_layout(B * HW, ws, &w, &lists);
        cudaError_t e = cudaMemsetAsync(w.hist1, 0, w.zero_bytes, s);      // (n_kept is zeroed by the kernel itself / the fallback path)
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
        Percents pc;
        pc.use_rank = 0; pc.rank = 0;
        for (int j = 0; j < kMaxQ; ++j) pc.q[j] = (j < nq) ? h_percents[j] : 0.0f;
        bool launched = false;
        int rc = (C == 19) ? launch_chain<19>(logits, target_in, hw, N, ignore, pc, nq, part_idx, entropy, thresh, n_valid, target_out,
                                              drop_mask, n_kept, w, lists, s, &launched)
                           : launch_chain<21>(logits, target_in, hw, N, ignore, pc, nq, part_idx, entropy, thresh, n_valid, target_out,
                                              drop_mask, n_kept, w, lists, s, &launched);
        if (rc != 0 || launched) return rc;
    }
    // multi-launch path: thresholds, then copy + partition
    int rc = u2pl_entropy_thresholds_fast(logits, target_in, B, C, HW, ignore, h_percents, nq, entropy, thresh, n_valid, ws, ws_bytes, stream);
    if (rc != 0) return rc;
    cudaError_t e = cudaMemcpyAsync(target_out, target_in, static_cast<size_t>(N) * sizeof(int64_t), cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
    return u2pl_partition_target(entropy, target_out, static_cast<int64_t>(N), ignore, thresh, part_idx, drop_mask, n_kept, stream);
}

extern "C" int u2pl_entropy_masks(const float *entropy, const int64_t *target, const int64_t *idx,
                                  int64_t n_out, int64_t ignore, const float *thresh, int lo_idx, int hi_idx,
                                  float *out_low, float *out_high, void *stream)
{
    if (n_out <= 0 || n_out >= (1LL << 31)) return bad_arg("entropy_masks: n_out must be in (0, 2^31)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(reinterpret_cast<const void *>(entropy_masks_kernel), 256, 0, static_cast<uint32_t>(n_out));
    entropy_masks_kernel<<<grid, 256, 0, s>>>(entropy, target, idx, static_cast<uint32_t>(n_out), ignore, thresh,
                                              lo_idx, hi_idx, out_low, out_high);
    return check_launch("entropy_masks");
}

extern "C" int u2pl_ohem_select(const float *logits, const int64_t *target, int64_t B, int64_t C, int64_t HW,
                                int64_t ignore, float thresh, int64_t min_kept,
                                int64_t *new_target, float *kth_value, int64_t *n_valid,
                                void *ws, size_t ws_bytes, void *stream)
{
    if (B <= 0 || C <= 0 || HW <= 0 || B * HW >= (1LL << 31)) return bad_arg("ohem_select: bad shape");
    if (ws_bytes < ws_layout(B * HW, nullptr, nullptr)) { set_error("ohem_select: workspace too small"); return U2PL_E_WS_SMALL; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t N = static_cast<uint32_t>(B * HW), hw = static_cast<uint32_t>(HW);
    EntropyWs w;
    ws_layout(B * HW, ws, &w);
    cudaError_t e = cudaMemsetAsync(w.hist1, 0, w.zero_bytes, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(n_valid, 0, sizeof(int64_t), s);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
    unsigned long long *nv = reinterpret_cast<unsigned long long *>(n_valid);
    const int grid = static_cast<int>(std::min<long long>((static_cast<long long>(N) + 255) / 256, 148 * 6));
    switch (C) {
        case 19: ohem_prob_hist_kernel<19><<<grid, 256, 0, s>>>(logits, target, hw, N, ignore, w.keys, w.hist1, nv); break;
        case 21: ohem_prob_hist_kernel<21><<<grid, 256, 0, s>>>(logits, target, hw, N, ignore, w.keys, w.hist1, nv); break;
        default: ohem_prob_hist_kernel_anyC<<<grid, 256, 0, s>>>(logits, target, static_cast<uint32_t>(C), hw, N, ignore, w.keys, w.hist1, nv);
    }
    Percents pc;
    for (int j = 0; j < kMaxQ; ++j) pc.q[j] = 0.0f;
    pc.use_rank = 1;
    const long long k = std::min<long long>(static_cast<long long>(N), std::max<long long>(min_kept, 1)) - 1;   // :521
    pc.rank = static_cast<uint32_t>(k);
    select1_kernel<<<1, 256, 0, s>>>(w.hist1, w.st, pc, 1);
    const size_t smem = 2 * kBinsR * 4;
    const int g2 = grid_for(reinterpret_cast<const void *>(hist_refine_kernel<2>), 256, smem, N, 4);
    hist_refine_kernel<2><<<g2, 256, smem, s>>>(w.keys, N, w.st, w.hist2, 2);
    select_refine_kernel<false><<<1, 1024, 0, s>>>(w.hist2, w.st, 1, kth_value, nullptr);
    hist_refine_kernel<3><<<g2, 256, smem, s>>>(w.keys, N, w.st, w.hist3, 2);
    select_refine_kernel<true><<<1, 1024, 0, s>>>(w.hist3, w.st, 1, kth_value, nullptr);
    ohem_partition_kernel<<<grid, 256, 0, s>>>(w.keys, target, N, ignore, thresh, static_cast<long long>(min_kept), kth_value, nv, new_target);
    return check_launch("ohem_select", 7);
}
