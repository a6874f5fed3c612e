// contra.cu -- class-wise memory-bank contrastive loss (reference loss_helper.py:51-235 and
// utils.py:28-47) as a handful of kernels over compact per-pixel class bitmasks:
//
//   onehot_to_bits   [B,C,h,w] int64 (multi-)hot labels -> one uint32 bitmask per pixel
//   classify         loop 1 of the reference (:103-154) for all classes at once: low-valid /
//                    anchor / negative-key membership bits per pixel + per-256-pixel-block
//                    per-class counts            (no sort: rank of a class = #probs above it)
//   scan             exclusive scan of those counts per class -> ordinal of every member pixel
//   proto            class prototypes = mean of teacher features over low-valid pixels (:119-123)
//   pack_keys        negative keys rep_teacher[negative_mask] (:142) gathered, in the
//                    reference's row-major order, into one packed [sum_k, D] buffer
//   bank_append      dequeue_and_enqueue (utils.py:28-47) on a device-resident ring buffer
//   infonce_fwd      loop 2 (:173-230): anchor gather, cosine logits against prototype + sampled
//                    bank rows, CE(.,0)/temp; one warp per query, warp-shuffle reductions,
//                    online softmax that also yields d loss / d anchor in the same pass
//   infonce_bwd      scatter-add of those rows into rep.grad
//
// Everything is gather / scan / copy work: HBM- (or L2-) bound, no tensor cores.
// Feature tensors are addressed through element strides (sn, sd, sp) so both NCHW
// (sd = h*w, sp = 1) and channels-last (sd = 1, sp = D) layouts are read in place.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "common.cuh"

namespace u2pl {

constexpr int kBlk = 256;          // pixels per classify block == scan granularity
constexpr int kMaxC = 32;          // classes per bitmask
constexpr int kMaxD = 256;         // feature dim handled by one warp with 2 float4 per lane

// ------------------------------------------------------------------ onehot -> bits
__global__ void __launch_bounds__(256)
onehot_to_bits_kernel(const int64_t *__restrict__ onehot, uint32_t B, uint32_t C, uint32_t hw,
                      uint32_t *__restrict__ bits)
{
    const uint32_t P = B * hw;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < P; i += gridDim.x * 256u) {
        const uint32_t b = i / hw, p = i - b * hw;
        const int64_t *x = onehot + static_cast<size_t>(b) * C * hw + p;
        uint32_t m = 0;
        for (uint32_t c = 0; c < C; ++c)
            if (__ldg(x + static_cast<size_t>(c) * hw) != 0) m |= (1u << c);
        bits[i] = m;
    }
}


// ------------------------------------------------------------------ fused low-res prep
// train_semi.py:408-465 without the two [B,C,H,W] one-hot temporaries: nearest-neighbour
// down-sampled low/high entropy masks and the class bitmask label_onehot()+interpolate would give.
// label_onehot's scatter quirk (utils.py:50-59) is reproduced: only batch slot 0 of
// each group carries classes -- the union over all images of the group, ignored pixels counting as
// class 0 -- and it is cleared where image 0 itself is ignored.
struct PrepArgs {
    const int64_t *label_l, *label_u;   // [Bl,H,W], [Bu,H,W]
    const float *entropy;               // [Bu,H,W]
    const float *thresh;
    int lo_idx, hi_idx;
    uint32_t Bl, Bu, H, W, h, w;
    int64_t ignore;
    int negative_high_entropy;
    uint32_t *bits;                     // [(Bl+Bu)*h*w]
    float *low_mask, *high_mask;        // [(Bl+Bu)*h*w]
};

__global__ void __launch_bounds__(256)
prep_lowres_kernel(PrepArgs a)
{
    const uint32_t hw = a.h * a.w, P = (a.Bl + a.Bu) * hw;
    const float sy = __fdiv_rn(static_cast<float>(a.H), static_cast<float>(a.h));   // ATen nearest: scale = in/out (fp32)
    const float sx = __fdiv_rn(static_cast<float>(a.W), static_cast<float>(a.w));
    const float tl = __ldg(a.thresh + a.lo_idx), th = __ldg(a.thresh + a.hi_idx);
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < P; i += gridDim.x * 256u) {
        const uint32_t n = i / hw, r = i - n * hw, y = r / a.w, x = r - y * a.w;
        const uint32_t yy = min(static_cast<uint32_t>(floorf(__fmul_rn(static_cast<float>(y), sy))), a.H - 1);
        const uint32_t xx = min(static_cast<uint32_t>(floorf(__fmul_rn(static_cast<float>(x), sx))), a.W - 1);
        const size_t src = static_cast<size_t>(yy) * a.W + xx, HW = static_cast<size_t>(a.H) * a.W;
        const bool labelled = n < a.Bl;
        const uint32_t b = labelled ? n : n - a.Bl, G = labelled ? a.Bl : a.Bu;
        const int64_t *lab = labelled ? a.label_l : a.label_u;
        const bool valid = __ldg(lab + b * HW + src) != a.ignore;
        float lo, hi;
        if (labelled) { lo = hi = valid ? 1.0f : 0.0f; }                               // :420-425,434-439
        else {
            const float e = __ldg(a.entropy + b * HW + src);
            lo = (valid && e <= tl) ? 1.0f : 0.0f;                                       // :408-410
            hi = a.negative_high_entropy ? ((valid && e >= th) ? 1.0f : 0.0f) : 1.0f;    // :416-418 / :442-450
        }
        uint32_t m = 0;
        if (b == 0 && valid) {
            for (uint32_t g = 0; g < G; ++g) {
                const int64_t t = __ldg(lab + g * HW + src);
                m |= 1u << static_cast<uint32_t>(t == a.ignore ? 0 : t);
            }
        }
        a.bits[i] = m;
        a.low_mask[i] = lo;
        a.high_mask[i] = hi;
    }
}

// ------------------------------------------------------------------ classify
struct ClassifyArgs {
    const uint32_t *label_bits;     // [P]   bit c = label[:, c] != 0      (labelled images first)
    const float *prob_l, *prob_u;   // [Bl,C,hw], [Bu,C,hw] teacher probabilities
    const float *low_mask, *high_mask;   // [P] fp32
    uint32_t Pl, P, C, hw;
    float thr, nthr;                // current_class_threshold, current_class_negative_threshold
    int low_rank, high_rank;
    uint32_t *bits3;                // [3][P]  0: low-valid, 1: anchor, 2: negative key
    uint32_t *blockcnt;             // [3][C][nb]
    uint32_t nb;
};

__global__ void __launch_bounds__(kBlk)
classify_kernel(ClassifyArgs a)
{
    __shared__ uint32_t s_cnt[3][kMaxC];
    if (threadIdx.x < 3 * kMaxC) (&s_cnt[0][0])[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t pix = blockIdx.x * kBlk + threadIdx.x;
    uint32_t lv = 0, an = 0, ng = 0;
    const uint32_t bits = (pix < a.P) ? __ldg(a.label_bits + pix) : 0u;
    if (bits) {
        const bool labelled = pix < a.Pl;
        const uint32_t q = labelled ? pix : pix - a.Pl;
        const uint32_t n = q / a.hw, p = q - n * a.hw;
        const float *pr = (labelled ? a.prob_l : a.prob_u) + static_cast<size_t>(n) * a.C * a.hw + p;
        const bool lm = __ldg(a.low_mask + pix) != 0.0f;      // low_valid = label * low_mask   (:80)
        const bool hm = __ldg(a.high_mask + pix) != 0.0f;     // high_valid = label * high_mask (:81)
        if (lm) lv = bits;
        uint32_t rest = bits;
        while (rest) {
            const int i = __ffs(rest) - 1;
            rest &= rest - 1;
            const float pi = __ldg(pr + static_cast<size_t>(i) * a.hw);
            if (lm && pi > a.thr) an |= (1u << i);                               // :108-110
            // labelled pixels never yield keys: class_mask_l requires label[:, i] == 0 (:137) while
            // high_valid requires label[:, i] != 0 (:81,105) -- quirk Q2.
            if (!labelled && hm && pi < a.nthr) {                                // :111-113
                int r = 0;                                                       // rank in stable descending order (:94)
                for (uint32_t j = 0; j < a.C; ++j) {
                    const float pj = __ldg(pr + static_cast<size_t>(j) * a.hw);
                    r += (pj > pi || (pj == pi && static_cast<int>(j) < i)) ? 1 : 0;
                }
                if (r >= a.low_rank && r < a.high_rank) ng |= (1u << i);         // :127-129,140
            }
        }
    }
    if (pix < a.P) {
        a.bits3[pix] = lv;
        a.bits3[a.P + pix] = an;
        a.bits3[2 * static_cast<size_t>(a.P) + pix] = ng;
    }
    const uint32_t any = __ballot_sync(0xffffffffu, bits != 0);
    if (any) {
        const int lane = threadIdx.x & 31;
        uint32_t u0 = lv, u1 = an, u2 = ng;                    // warp-wide OR of present classes
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            u0 |= __shfl_xor_sync(0xffffffffu, u0, o);
            u1 |= __shfl_xor_sync(0xffffffffu, u1, o);
            u2 |= __shfl_xor_sync(0xffffffffu, u2, o);
        }
        uint32_t rest = u0 | u1 | u2;
        while (rest) {
            const int c = __ffs(rest) - 1;
            rest &= rest - 1;
            const uint32_t b0 = __ballot_sync(0xffffffffu, (lv >> c) & 1u);
            const uint32_t b1 = __ballot_sync(0xffffffffu, (an >> c) & 1u);
            const uint32_t b2 = __ballot_sync(0xffffffffu, (ng >> c) & 1u);
            if (lane == 0) {
                if (b0) atomicAdd(&s_cnt[0][c], __popc(b0));
                if (b1) atomicAdd(&s_cnt[1][c], __popc(b1));
                if (b2) atomicAdd(&s_cnt[2][c], __popc(b2));
            }
        }
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < 3 * a.C; j += kBlk) {
        const uint32_t t = j / a.C, c = j - t * a.C;
        a.blockcnt[(static_cast<size_t>(t) * a.C + c) * a.nb + blockIdx.x] = s_cnt[t][c];
    }
}

// one block per (type, class): exclusive scan over the nb per-block counts
__global__ void __launch_bounds__(1024)
scan_kernel(const uint32_t *__restrict__ blockcnt, uint32_t *__restrict__ blockoff,
            uint32_t *__restrict__ totals, uint32_t nb)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const uint32_t *in = blockcnt + static_cast<size_t>(blockIdx.x) * nb;
    uint32_t *out = blockoff + static_cast<size_t>(blockIdx.x) * nb;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + tid;
        const uint32_t v = (i < nb) ? in[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < wid; ++w) wbase += warp_tot[w];
        const uint32_t carry = carry_s;
        if (i < nb) out[i] = carry + wbase + inc - v;
        __syncthreads();
        if (tid == 1023) carry_s = carry + wbase + inc;
        __syncthreads();
    }
    if (tid == 0) totals[blockIdx.x] = carry_s;
}

// ------------------------------------------------------------------ shared tile loader
// Loads features of 32 consecutive pixels (tile t) into smem tile[32][D+1]; rows of pixels whose
// bitmask is zero are skipped (left untouched).
__device__ __forceinline__ void load_tile(const float *__restrict__ feat, long long sn, long long sd, long long sp,
                                          uint32_t hw, uint32_t P, uint32_t D, uint32_t tile_first,
                                          const uint32_t *s_bits, float *tile)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (sp == 1) {                  // NCHW: lanes run along pixels (coalesced), warps stride over channels
        const uint32_t pix = tile_first + lane;
        const bool on = pix < P && s_bits[lane] != 0;
        const uint32_t n = pix / hw, p = pix - n * hw;
        const float *src = feat + static_cast<long long>(n) * sn + p;
        for (uint32_t d = wid; d < D; d += nw)
            if (on) tile[lane * (D + 1) + d] = __ldg(src + static_cast<long long>(d) * sd);
    } else {                        // channels-last: lanes run along channels
        for (int px = wid; px < 32; px += nw) {
            const uint32_t pix = tile_first + px;
            if (pix >= P || s_bits[px] == 0) continue;
            const uint32_t n = pix / hw, p = pix - n * hw;
            const float *src = feat + static_cast<long long>(n) * sn + static_cast<long long>(p) * sp;
            for (uint32_t d = lane; d < D; d += 32) tile[px * (D + 1) + d] = __ldg(src + static_cast<long long>(d) * sd);
        }
    }
}

// ------------------------------------------------------------------ prototypes
__global__ void __launch_bounds__(256)
proto_partial_kernel(const float *__restrict__ feat, long long sn, long long sd, long long sp,
                     uint32_t hw, uint32_t P, uint32_t C, uint32_t D,
                     const uint32_t *__restrict__ lv_bits, float *__restrict__ partial)
{
    extern __shared__ float smem[];
    float *tile = smem;                                   // [32][D+1]
    float *acc = smem + 32 * (D + 1);                     // [C][D]
    __shared__ uint32_t s_bits[32];
    __shared__ uint32_t s_any;
    for (uint32_t j = threadIdx.x; j < C * D; j += 256) acc[j] = 0.0f;
    const uint32_t ntiles = (P + 31) / 32;
    // tiles are dealt round-robin: with the reference's label_onehot quirk only image 0 of each half of the
    // batch has members, and contiguous ranges would leave that work to a handful of blocks
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t pix = t * 32 + threadIdx.x;
            const uint32_t b = pix < P ? __ldg(lv_bits + pix) : 0u;
            s_bits[threadIdx.x] = b;
            const uint32_t any = __ballot_sync(0xffffffffu, b != 0);
            if (threadIdx.x == 0) s_any = any;
        }
        __syncthreads();
        if (s_any == 0) continue;
        load_tile(feat, sn, sd, sp, hw, P, D, t * 32, s_bits, tile);
        __syncthreads();
        for (uint32_t d = threadIdx.x; d < D; d += 256) {
            for (int px = 0; px < 32; ++px) {             // fixed order: deterministic sums
                uint32_t b = s_bits[px];
                if (!b) continue;
                const float v = tile[px * (D + 1) + d];
                while (b) {
                    const int c = __ffs(b) - 1;
                    b &= b - 1;
                    acc[c * D + d] += v;
                }
            }
        }
    }
    __syncthreads();
    float *out = partial + static_cast<size_t>(blockIdx.x) * C * D;
    for (uint32_t j = threadIdx.x; j < C * D; j += 256) out[j] = acc[j];
}

// Channels-last features (sd == 1, D == 256): thread d owns channel d and keeps the C class sums in REGISTERS; a pixel's
// 256 features are one coalesced 1 KB (fp32) / 512 B (bf16) row read straight from the network's output tensor, members are
// found from a broadcast word of the class bitmask, and the per-class adds are predicated register operations.  The shared-
// memory read-modify-write form above serialises on LDS->FADD->STS chains (1.6 ms per V16 step with the reference's
// multi-hot slot-0 labels: ~17 classes per member pixel); this one is bound by the member rows it reads.  Same fixed
// summation order per (block, channel): pixels ascending -> deterministic.
template <int C, typename T>
__global__ void __launch_bounds__(256)
proto_partial_cl_kernel(const T *__restrict__ feat, long long sn, long long sp, uint32_t hw, uint32_t P,
                        const uint32_t *__restrict__ lv_bits, float *__restrict__ partial)
{
    constexpr uint32_t D = 256;
    __shared__ uint32_t s_bits[2][32];
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.0f;
    const uint32_t ntiles = (P + 31) / 32;
    const uint32_t d = threadIdx.x;
    uint32_t it = 0;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {       // round-robin (see proto_partial_kernel)
        uint32_t *sb = s_bits[it & 1];
        if (threadIdx.x < 32) {
            const uint32_t pix = t * 32 + threadIdx.x;
            sb[threadIdx.x] = pix < P ? __ldg(lv_bits + pix) : 0u;
        }
        __syncthreads();                                   // double-buffered: one barrier per tile
#pragma unroll 4
        for (int px = 0; px < 32; ++px) {
            const uint32_t b = sb[px];
            if (!b) continue;
            const uint32_t pix = t * 32 + px;
            const uint32_t n = pix / hw, q = pix - n * hw;
            const float v = static_cast<float>(feat[static_cast<long long>(n) * sn + static_cast<long long>(q) * sp + d]);
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] += ((b >> c) & 1u) ? v : 0.0f;
        }
    }
    float *out = partial + static_cast<size_t>(blockIdx.x) * C * D;
#pragma unroll
    for (int c = 0; c < C; ++c) out[c * D + d] = acc[c];
}

__global__ void __launch_bounds__(256)
proto_reduce_kernel(const float *__restrict__ partial, int nparts, uint32_t C, uint32_t D,
                    const uint32_t *__restrict__ lv_totals, float *__restrict__ proto)
{
    const uint32_t j = blockIdx.x * 256u + threadIdx.x;
    if (j >= C * D) return;
    float s = 0.0f;
    for (int k = 0; k < nparts; ++k) s += partial[static_cast<size_t>(k) * C * D + j];
    proto[j] = s / static_cast<float>(lv_totals[j / D]);   // 0/0 -> NaN like torch.mean of an empty set
}

// ------------------------------------------------------------------ pack negative keys
// One block per 256-pixel classify block.  Destination row of a key =
// class_base[c] + blockoff[neg][c][block] + (rank of the pixel among the block's members of c).
__global__ void __launch_bounds__(256)
pack_keys_kernel(const float *__restrict__ feat, long long sn, long long sd, long long sp,
                 uint32_t hw, uint32_t P, uint32_t C, uint32_t D, uint32_t nb,
                 const uint32_t *__restrict__ ng_bits, const uint32_t *__restrict__ blockoff_ng,
                 const uint32_t *__restrict__ class_base, float *__restrict__ packed)
{
    extern __shared__ float smem[];
    float *tile = smem;                                   // [32][D+1]
    __shared__ uint32_t s_bits[32];
    __shared__ uint32_t s_run[kMaxC];
    __shared__ uint32_t s_dst[32 * kMaxC];                // (pixel-in-tile << 27 | unused) .. keep two arrays
    __shared__ uint8_t s_px[32 * kMaxC];
    __shared__ uint32_t s_n, s_any;
    const uint32_t blk = blockIdx.x;
    if (threadIdx.x < C)
        s_run[threadIdx.x] = class_base[threadIdx.x] + blockoff_ng[static_cast<size_t>(threadIdx.x) * nb + blk];
    for (int sub = 0; sub < kBlk / 32; ++sub) {
        __syncthreads();
        const uint32_t first = blk * kBlk + sub * 32;
        if (threadIdx.x < 32) {
            const uint32_t pix = first + threadIdx.x;
            const uint32_t b = pix < P ? __ldg(ng_bits + pix) : 0u;
            s_bits[threadIdx.x] = b;
            uint32_t u = b;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) u |= __shfl_xor_sync(0xffffffffu, u, o);
            uint32_t n = 0;
            uint32_t rest = u;
            while (rest) {                                 // classes in ascending order; order inside a class = pixel order
                const int c = __ffs(rest) - 1;
                rest &= rest - 1;
                const uint32_t bal = __ballot_sync(0xffffffffu, (b >> c) & 1u);
                const uint32_t before = __popc(bal & ((1u << threadIdx.x) - 1u));
                if ((b >> c) & 1u) {
                    s_dst[n + before] = s_run[c] + before;
                    s_px[n + before] = static_cast<uint8_t>(threadIdx.x);
                }
                n += __popc(bal);
                __syncwarp();
                if (threadIdx.x == 0) s_run[c] += __popc(bal);
                __syncwarp();
            }
            if (threadIdx.x == 0) { s_n = n; s_any = u; }
        }
        __syncthreads();
        if (s_any == 0) continue;
        load_tile(feat, sn, sd, sp, hw, P, D, first, s_bits, tile);
        __syncthreads();
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (uint32_t e = wid; e < s_n; e += 8) {
            float *dst = packed + static_cast<size_t>(s_dst[e]) * D;
            const float *src = tile + s_px[e] * (D + 1);
            for (uint32_t d = lane; d < D; d += 32) dst[d] = src[d];
        }
    }
}

// ------------------------------------------------------------------ bank append
// desc[i] = {src_first_row, dst_class_row_base, dst_first (ring position), cap, count}
struct AppendDesc { uint32_t src_first, dst_base, dst_first, cap, count; };

__global__ void __launch_bounds__(256)
bank_append_kernel(const float *__restrict__ src, float *__restrict__ bank, uint32_t D,
                   const AppendDesc *__restrict__ desc, int ndesc)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = blockIdx.y; i < ndesc; i += gridDim.y) {
        const AppendDesc d = desc[i];
        for (uint32_t r = blockIdx.x * 8 + wid; r < d.count; r += gridDim.x * 8) {
            const float4 *s = reinterpret_cast<const float4 *>(src + static_cast<size_t>(d.src_first + r) * D);
            float4 *t = reinterpret_cast<float4 *>(bank + (static_cast<size_t>(d.dst_base) + (d.dst_first + r) % d.cap) * D);
            for (uint32_t v = lane; v < D / 4; v += 32) t[v] = __ldg(s + v);
        }
    }
}

// ------------------------------------------------------------------ InfoNCE
struct InfoNceArgs {
    const float *rep;               // student features, strides below (elements)
    long long sn, sd, sp;
    uint32_t hw, P, D, nb;
    const uint32_t *an_bits;        // [P] anchor membership bits
    const uint32_t *blockoff_an;    // [C][nb]
    const int32_t *act_class;       // [nact] class whose anchors / prototype are used (list position j, quirk Q1)
    const int32_t *a_ord;           // [nact][nq] sampled anchor ordinals (torch.randint, :179-181)
    const int32_t *neg_rows;        // [nact][nq][nneg] sampled bank rows, already mapped to physical rows (:194-197)
    const float *proto;             // [C][D]
    const float *bank;              // [rows][D]
    int nact, nq, nneg;
    float inv_temp, scale;          // 1/temperature, 1/(nq * valid_seg)
    float *loss_q;                  // [nact*nq]  per-query CE
    float *grad_rows;               // [nact*nq][D]  scale * d CE_q / d anchor
    int32_t *anchor_pix;            // [nact*nq]
    const float *const *class_bank; // kPeer only: [nact] base of the bank shard that holds the active class's ring --
                                    // local memory or a peer GPU's (CUDA IPC mapping, read over NVLink); neg_rows are
                                    // rows inside that shard
};

__device__ __forceinline__ float2 warp_sum2(float a, float b)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    return make_float2(a, b);
}

template <bool kPeer>
__global__ void __launch_bounds__(128)
infonce_fwd_kernel(InfoNceArgs a)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= a.nact * a.nq) return;
    const int act = w / a.nq;
    const int cls = a.act_class[act];
    // ---- ordinal -> pixel (k-th anchor pixel of class cls in row-major order)
    const uint32_t k = static_cast<uint32_t>(a.a_ord[w]);
    const uint32_t *off = a.blockoff_an + static_cast<size_t>(cls) * a.nb;
    uint32_t lo = 0, hi = a.nb;                            // last block with off[b] <= k
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= k) lo = mid; else hi = mid;
    }
    uint32_t kk = k - __ldg(off + lo);
    const uint32_t base = lo * kBlk + lane * 8;
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t pix = base + j;
        const uint32_t b = pix < a.P ? __ldg(a.an_bits + pix) : 0u;
        mine |= ((b >> cls) & 1u) << j;
    }
    uint32_t inc = __popc(mine);
    const uint32_t cnt = inc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    const uint32_t excl = inc - cnt;
    int found = -1;
    if (kk >= excl && kk < inc) {
        uint32_t r = kk - excl, m = mine;
        for (uint32_t t = 0; t < r; ++t) m &= m - 1;
        found = static_cast<int>(base + __ffs(m) - 1);
    }
    const uint32_t who = __ballot_sync(0xffffffffu, found >= 0);
    const int pix = __shfl_sync(0xffffffffu, found, who ? __ffs(who) - 1 : 0);
    if (lane == 0) a.anchor_pix[w] = pix;
    if (pix < 0) { if (lane == 0) a.loss_q[w] = __uint_as_float(0x7fc00000u); return; }   // inconsistent ordinal
    // ---- anchor row: lane owns channels [4*lane, 4*lane+4) and [128+4*lane, ...)
    const uint32_t n = static_cast<uint32_t>(pix) / a.hw, p = static_cast<uint32_t>(pix) - n * a.hw;
    const float *ar = a.rep + static_cast<long long>(n) * a.sn + static_cast<long long>(p) * a.sp;
    float av[8];
    const uint32_t d0 = 4 * lane, d1 = 128 + 4 * lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        av[j] = (d0 + j < a.D) ? __ldg(ar + static_cast<long long>(d0 + j) * a.sd) : 0.0f;
        av[4 + j] = (d1 + j < a.D) ? __ldg(ar + static_cast<long long>(d1 + j) * a.sd) : 0.0f;
    }
    float ss = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += av[j] * av[j];
    ss = warp_sum(ss);
    const float eps = 1e-8f;
    const float anorm = sqrtf(ss);
    const float ainv = 1.0f / fmaxf(anorm, eps);           // torch.cosine_similarity: x / max(||x||, eps)
#pragma unroll
    for (int j = 0; j < 8; ++j) av[j] *= ainv;
    // ---- keys: prototype first (:201-207,220-222), then the sampled negatives
    float m = -INFINITY, s = 0.0f, u = 0.0f, l0 = 0.0f;
    float V[8], K0[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { V[j] = 0.0f; K0[j] = 0.0f; }
    const int32_t *rows = a.neg_rows + static_cast<size_t>(w) * a.nneg;
    const float *bank = kPeer ? a.class_bank[act] : a.bank;
    const bool c0 = d0 < a.D, c1 = d1 < a.D;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *kr0 = a.proto + static_cast<size_t>(cls) * a.D;
    float4 n0 = c0 ? __ldg(reinterpret_cast<const float4 *>(kr0 + d0)) : z4;      // row t+1 is in flight while row t is reduced
    float4 n1 = c1 ? __ldg(reinterpret_cast<const float4 *>(kr0 + d1)) : z4;
    for (int t = 0; t <= a.nneg; ++t) {
        const float4 x0 = n0, x1 = n1;
        if (t < a.nneg) {
            const float *kn_ = bank + static_cast<size_t>(__ldg(rows + t)) * a.D;
            n0 = c0 ? __ldg(reinterpret_cast<const float4 *>(kn_ + d0)) : z4;
            n1 = c1 ? __ldg(reinterpret_cast<const float4 *>(kn_ + d1)) : z4;
        }
        float kv[8];
        kv[0] = x0.x; kv[1] = x0.y; kv[2] = x0.z; kv[3] = x0.w;
        kv[4] = x1.x; kv[5] = x1.y; kv[6] = x1.z; kv[7] = x1.w;
        float dot = 0.0f, kn = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { dot += av[j] * kv[j]; kn += kv[j] * kv[j]; }
        const float2 r2 = warp_sum2(dot, kn);
        const float kinv = 1.0f / fmaxf(sqrtf(r2.y), eps);
        const float l = r2.x * kinv;                       // cosine similarity
        const float z = l * a.inv_temp;
        const float mn = fmaxf(m, z);
        const float cs = __expf(m - mn), wgt = __expf(z - mn);
        s = s * cs + wgt;
        u = u * cs + wgt * l;
#pragma unroll
        for (int j = 0; j < 8; ++j) V[j] = V[j] * cs + wgt * (kv[j] * kinv);
        m = mn;
        if (t == 0) {
            l0 = l;
#pragma unroll
            for (int j = 0; j < 8; ++j) K0[j] = kv[j] * kinv;
        }
    }
    const float loss = m + logf(s) - l0 * a.inv_temp;      // CE(logits/temp, 0)  (:228-230)
    if (lane == 0) a.loss_q[w] = loss;
    // d CE / d anchor = (1/(temp*||a||)) [ (sum_k p_k k^ - k^_0) - (sum_k p_k l_k - l_0) a^ ]
    const float sinv = 1.0f / s;
    const float coef = a.scale * a.inv_temp * ainv;
    const float proj = (anorm >= eps) ? (u * sinv - l0) : 0.0f;
    float *g = a.grad_rows + static_cast<size_t>(w) * a.D;
    float gv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) gv[j] = coef * ((V[j] * sinv - K0[j]) - proj * av[j]);
    if (c0) *reinterpret_cast<float4 *>(g + d0) = make_float4(gv[0], gv[1], gv[2], gv[3]);
    if (c1) *reinterpret_cast<float4 *>(g + d1) = make_float4(gv[4], gv[5], gv[6], gv[7]);
}

// Same kernel with kDepth key rows in flight per warp instead of one (the loop is latency-bound at depth 1: ~1 us per
// 1 KB row, 51 rows per query).  Arithmetic and its order are unchanged, so results are bit-identical.  Opt-in
// (U2PL_INFONCE_DEPTH=2|4) until measured.
template <bool kPeer, int kDepth>
__device__ __forceinline__ void infonce_query_pipelined(const InfoNceArgs &a, const int w, const int lane)
{
    const int act = w / a.nq;
    const int cls = a.act_class[act];
    // ---- ordinal -> pixel (k-th anchor pixel of class cls in row-major order)
    const uint32_t k = static_cast<uint32_t>(a.a_ord[w]);
    const uint32_t *off = a.blockoff_an + static_cast<size_t>(cls) * a.nb;
    uint32_t lo = 0, hi = a.nb;                            // last block with off[b] <= k
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= k) lo = mid; else hi = mid;
    }
    uint32_t kk = k - __ldg(off + lo);
    const uint32_t base = lo * kBlk + lane * 8;
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t pix = base + j;
        const uint32_t b = pix < a.P ? __ldg(a.an_bits + pix) : 0u;
        mine |= ((b >> cls) & 1u) << j;
    }
    uint32_t inc = __popc(mine);
    const uint32_t cnt = inc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    const uint32_t excl = inc - cnt;
    int found = -1;
    if (kk >= excl && kk < inc) {
        uint32_t r = kk - excl, m = mine;
        for (uint32_t t = 0; t < r; ++t) m &= m - 1;
        found = static_cast<int>(base + __ffs(m) - 1);
    }
    const uint32_t who = __ballot_sync(0xffffffffu, found >= 0);
    const int pix = __shfl_sync(0xffffffffu, found, who ? __ffs(who) - 1 : 0);
    if (lane == 0) a.anchor_pix[w] = pix;
    if (pix < 0) { if (lane == 0) a.loss_q[w] = __uint_as_float(0x7fc00000u); return; }   // inconsistent ordinal
    // ---- anchor row: lane owns channels [4*lane, 4*lane+4) and [128+4*lane, ...)
    const uint32_t n = static_cast<uint32_t>(pix) / a.hw, p = static_cast<uint32_t>(pix) - n * a.hw;
    const float *ar = a.rep + static_cast<long long>(n) * a.sn + static_cast<long long>(p) * a.sp;
    float av[8];
    const uint32_t d0 = 4 * lane, d1 = 128 + 4 * lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        av[j] = (d0 + j < a.D) ? __ldg(ar + static_cast<long long>(d0 + j) * a.sd) : 0.0f;
        av[4 + j] = (d1 + j < a.D) ? __ldg(ar + static_cast<long long>(d1 + j) * a.sd) : 0.0f;
    }
    float ss = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += av[j] * av[j];
    ss = warp_sum(ss);
    const float eps = 1e-8f;
    const float anorm = sqrtf(ss);
    const float ainv = 1.0f / fmaxf(anorm, eps);           // torch.cosine_similarity: x / max(||x||, eps)
#pragma unroll
    for (int j = 0; j < 8; ++j) av[j] *= ainv;
    // ---- keys: prototype first (:201-207,220-222), then the sampled negatives
    float m = -INFINITY, s = 0.0f, u = 0.0f, l0 = 0.0f;
    float V[8], K0[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { V[j] = 0.0f; K0[j] = 0.0f; }
    const int32_t *rows = a.neg_rows + static_cast<size_t>(w) * a.nneg;
    const float *bank = kPeer ? a.class_bank[act] : a.bank;
    const bool c0 = d0 < a.D, c1 = d1 < a.D;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *kr0 = a.proto + static_cast<size_t>(cls) * a.D;
    // key 0 is the prototype, key k >= 1 is bank row rows[k-1]; q0/q1[i] hold key (t + i) of the current group
    float4 q0[kDepth], q1[kDepth];
    q0[0] = c0 ? __ldg(reinterpret_cast<const float4 *>(kr0 + d0)) : z4;
    q1[0] = c1 ? __ldg(reinterpret_cast<const float4 *>(kr0 + d1)) : z4;
#pragma unroll
    for (int i = 1; i < kDepth; ++i) {
        q0[i] = z4; q1[i] = z4;
        if (i - 1 < a.nneg) {
            const float *kn_ = bank + static_cast<size_t>(__ldg(rows + i - 1)) * a.D;
            q0[i] = c0 ? __ldg(reinterpret_cast<const float4 *>(kn_ + d0)) : z4;
            q1[i] = c1 ? __ldg(reinterpret_cast<const float4 *>(kn_ + d1)) : z4;
        }
    }
    for (int t = 0; t <= a.nneg; t += kDepth) {
#pragma unroll
        for (int i = 0; i < kDepth; ++i) {
            if (t + i > a.nneg) continue;                      // (uniform across the warp; no `break`: keeps q0/q1 in registers)
            const float4 x0 = q0[i], x1 = q1[i];
            const int nxt = t + i + kDepth;                    // key that takes this slot: bank row rows[nxt - 1]
            if (nxt <= a.nneg) {
                const float *kn_ = bank + static_cast<size_t>(__ldg(rows + nxt - 1)) * a.D;
                q0[i] = c0 ? __ldg(reinterpret_cast<const float4 *>(kn_ + d0)) : z4;
                q1[i] = c1 ? __ldg(reinterpret_cast<const float4 *>(kn_ + d1)) : z4;
            }
            float kv[8];
            kv[0] = x0.x; kv[1] = x0.y; kv[2] = x0.z; kv[3] = x0.w;
            kv[4] = x1.x; kv[5] = x1.y; kv[6] = x1.z; kv[7] = x1.w;
            float dot = 0.0f, kn = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { dot += av[j] * kv[j]; kn += kv[j] * kv[j]; }
            const float2 r2 = warp_sum2(dot, kn);
            const float kinv = 1.0f / fmaxf(sqrtf(r2.y), eps);
            const float l = r2.x * kinv;                       // cosine similarity
            const float z = l * a.inv_temp;
            const float mn = fmaxf(m, z);
            const float cs = __expf(m - mn), wgt = __expf(z - mn);
            s = s * cs + wgt;
            u = u * cs + wgt * l;
#pragma unroll
            for (int j = 0; j < 8; ++j) V[j] = V[j] * cs + wgt * (kv[j] * kinv);
            m = mn;
            if (t + i == 0) {
                l0 = l;
#pragma unroll
                for (int j = 0; j < 8; ++j) K0[j] = kv[j] * kinv;
            }
        }
    }
    const float loss = m + logf(s) - l0 * a.inv_temp;      // CE(logits/temp, 0)  (:228-230)
    if (lane == 0) a.loss_q[w] = loss;
    // d CE / d anchor = (1/(temp*||a||)) [ (sum_k p_k k^ - k^_0) - (sum_k p_k l_k - l_0) a^ ]
    const float sinv = 1.0f / s;
    const float coef = a.scale * a.inv_temp * ainv;
    const float proj = (anorm >= eps) ? (u * sinv - l0) : 0.0f;
    float *g = a.grad_rows + static_cast<size_t>(w) * a.D;
    float gv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) gv[j] = coef * ((V[j] * sinv - K0[j]) - proj * av[j]);
    if (c0) *reinterpret_cast<float4 *>(g + d0) = make_float4(gv[0], gv[1], gv[2], gv[3]);
    if (c1) *reinterpret_cast<float4 *>(g + d1) = make_float4(gv[4], gv[5], gv[6], gv[7]);
}

template <bool kPeer, int kDepth>
__global__ void __launch_bounds__(128)
infonce_fwd_pipelined_kernel(InfoNceArgs a)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= a.nact * a.nq) return;
    infonce_query_pipelined<kPeer, kDepth>(a, w, lane);
}

// sum of per-query losses in a fixed order: loss = scale * sum_q CE_q
__global__ void __launch_bounds__(256)
infonce_loss_kernel(const float *__restrict__ loss_q, int n, float scale, float *__restrict__ loss)
{
    __shared__ double sd[256];
    double s = 0.0;
    for (int j = threadIdx.x; j < n; j += 256) s += static_cast<double>(loss_q[j]);
    sd[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sd[threadIdx.x] += sd[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = static_cast<float>(sd[0] * static_cast<double>(scale));
}

__global__ void __launch_bounds__(128)
infonce_bwd_kernel(const float *__restrict__ grad_rows, const int32_t *__restrict__ anchor_pix, int nrows,
                   uint32_t D, uint32_t hw, long long sn, long long sd, long long sp,
                   const float *__restrict__ upstream, float *__restrict__ grad_rep)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= nrows) return;
    const int pix = anchor_pix[w];
    if (pix < 0) return;
    const float up = upstream ? __ldg(upstream) : 1.0f;
    const uint32_t n = static_cast<uint32_t>(pix) / hw, p = static_cast<uint32_t>(pix) - n * hw;
    float *dst = grad_rep + static_cast<long long>(n) * sn + static_cast<long long>(p) * sp;
    const float *g = grad_rows + static_cast<size_t>(w) * D;
    for (uint32_t d = lane; d < D; d += 32) atomicAdd(dst + static_cast<long long>(d) * sd, up * g[d]);
}

}  // namespace u2pl

using namespace u2pl;

extern "C" int u2pl_onehot_to_bits(const int64_t *onehot, int64_t B, int64_t C, int64_t hw, uint32_t *bits, void *stream)
{
    if (B <= 0 || hw <= 0 || C <= 0 || C > kMaxC || B * hw >= (1LL << 31)) return bad_arg("onehot_to_bits: need 0 < C <= 32, B*hw < 2^31");
    const uint32_t P = static_cast<uint32_t>(B * hw);
    const int grid = static_cast<int>(std::min<long long>((P + 255) / 256, 148 * 8));
    onehot_to_bits_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(onehot, static_cast<uint32_t>(B), static_cast<uint32_t>(C),
                                                                             static_cast<uint32_t>(hw), bits);
    return check_launch("onehot_to_bits");
}


extern "C" int u2pl_contra_prep_lowres(const int64_t *label_l, const int64_t *label_u, const float *entropy,
                                       const float *thresh, int lo_idx, int hi_idx,
                                       int64_t Bl, int64_t Bu, int64_t H, int64_t W, int64_t h, int64_t w,
                                       int64_t C, int64_t ignore, int negative_high_entropy,
                                       uint32_t *label_bits, float *low_mask, float *high_mask, void *stream)
{
    if (C <= 0 || C > kMaxC || (Bl + Bu) * h * w >= (1LL << 31) || h <= 0 || w <= 0) return bad_arg("contra_prep_lowres: bad shape");
    PrepArgs a;
    a.label_l = label_l; a.label_u = label_u; a.entropy = entropy; a.thresh = thresh; a.lo_idx = lo_idx; a.hi_idx = hi_idx;
    a.Bl = static_cast<uint32_t>(Bl); a.Bu = static_cast<uint32_t>(Bu); a.H = static_cast<uint32_t>(H); a.W = static_cast<uint32_t>(W);
    a.h = static_cast<uint32_t>(h); a.w = static_cast<uint32_t>(w); a.ignore = ignore; a.negative_high_entropy = negative_high_entropy;
    a.bits = label_bits; a.low_mask = low_mask; a.high_mask = high_mask;
    const long long P = (Bl + Bu) * h * w;
    const int grid = static_cast<int>(std::min<long long>((P + 255) / 256, 148 * 8));
    prep_lowres_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("contra_prep_lowres");
}

extern "C" int64_t u2pl_contra_num_blocks(int64_t P) { return (P + kBlk - 1) / kBlk; }

extern "C" int u2pl_contra_classify(const uint32_t *label_bits, const float *prob_l, const float *prob_u,
                                    const float *low_mask, const float *high_mask,
                                    int64_t Bl, int64_t Bu, int64_t C, int64_t hw,
                                    float thr, float nthr, int low_rank, int high_rank,
                                    uint32_t *bits3, uint32_t *blockcnt, uint32_t *blockoff, uint32_t *totals, void *stream)
{
    const int64_t P = (Bl + Bu) * hw;
    if (C <= 0 || C > kMaxC || P <= 0 || P >= (1LL << 31)) return bad_arg("contra_classify: need 0 < C <= 32, P < 2^31");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ClassifyArgs a;
    a.label_bits = label_bits; a.prob_l = prob_l; a.prob_u = prob_u; a.low_mask = low_mask; a.high_mask = high_mask;
    a.Pl = static_cast<uint32_t>(Bl * hw); a.P = static_cast<uint32_t>(P); a.C = static_cast<uint32_t>(C); a.hw = static_cast<uint32_t>(hw);
    a.thr = thr; a.nthr = nthr; a.low_rank = low_rank; a.high_rank = high_rank;
    a.bits3 = bits3; a.blockcnt = blockcnt; a.nb = static_cast<uint32_t>((P + kBlk - 1) / kBlk);
    classify_kernel<<<a.nb, kBlk, 0, s>>>(a);
    scan_kernel<<<static_cast<int>(3 * C), 1024, 0, s>>>(blockcnt, blockoff, totals, a.nb);
    return check_launch("contra_classify", 2);
}

static size_t tile_smem(int64_t D) { return static_cast<size_t>(32) * (D + 1) * 4; }

extern "C" int64_t u2pl_contra_proto_parts(void) { return 2 * kNumSMs; }

extern "C" int u2pl_contra_proto(const float *rep_teacher, int64_t sn, int64_t sd, int64_t sp,
                                 int64_t P, int64_t C, int64_t D, int64_t hw,
                                 const uint32_t *lv_bits, const uint32_t *lv_totals,
                                 float *partial, float *proto, void *stream)
{
    if (C <= 0 || C > kMaxC || D <= 0 || D > 1024 || P <= 0) return bad_arg("contra_proto: bad shape");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int parts = 2 * kNumSMs;
    const size_t smem = tile_smem(D) + static_cast<size_t>(C) * D * 4;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(proto_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) { set_error("contra_proto: shared memory request too large"); return static_cast<int>(e); }
        configured = smem;
    }
    if (sd == 1 && D == 256 && (C == 19 || C == 21)) {    // channels-last network output: register-accumulating kernel
        if (C == 19) proto_partial_cl_kernel<19, float><<<parts, 256, 0, s>>>(rep_teacher, sn, sp, static_cast<uint32_t>(hw), static_cast<uint32_t>(P), lv_bits, partial);
        else proto_partial_cl_kernel<21, float><<<parts, 256, 0, s>>>(rep_teacher, sn, sp, static_cast<uint32_t>(hw), static_cast<uint32_t>(P), lv_bits, partial);
    } else
    proto_partial_kernel<<<parts, 256, smem, s>>>(rep_teacher, sn, sd, sp, static_cast<uint32_t>(hw), static_cast<uint32_t>(P),
                                                  static_cast<uint32_t>(C), static_cast<uint32_t>(D), lv_bits, partial);
    proto_reduce_kernel<<<static_cast<int>((C * D + 255) / 256), 256, 0, s>>>(partial, parts, static_cast<uint32_t>(C),
                                                                            static_cast<uint32_t>(D), lv_totals, proto);
    return check_launch("contra_proto", 2);
}

extern "C" int u2pl_contra_pack_keys(const float *rep_teacher, int64_t sn, int64_t sd, int64_t sp,
                                     int64_t P, int64_t C, int64_t D, int64_t hw,
                                     const uint32_t *ng_bits, const uint32_t *blockoff_ng, const uint32_t *class_base,
                                     float *packed, void *stream)
{
    if (C <= 0 || C > kMaxC || D <= 0 || D > 1024 || P <= 0) return bad_arg("contra_pack_keys: bad shape");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t nb = static_cast<uint32_t>((P + kBlk - 1) / kBlk);
    const size_t smem = tile_smem(D);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(pack_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) { set_error("contra_pack_keys: shared memory request too large"); return static_cast<int>(e); }
        configured = smem;
    }
    pack_keys_kernel<<<nb, 256, smem, s>>>(rep_teacher, sn, sd, sp, static_cast<uint32_t>(hw), static_cast<uint32_t>(P),
                                           static_cast<uint32_t>(C), static_cast<uint32_t>(D), nb, ng_bits, blockoff_ng, class_base, packed);
    return check_launch("contra_pack_keys");
}

extern "C" int u2pl_bank_append(const float *src_rows, float *bank, int64_t D, const uint32_t *desc, int ndesc,
                                int64_t max_count, void *stream)
{
    if (D <= 0 || D % 4 != 0) return bad_arg("bank_append: D must be a positive multiple of 4");
    if (ndesc <= 0 || max_count <= 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int gx = static_cast<int>(std::min<long long>((max_count + 7) / 8, 64));
    dim3 grid(gx, std::min(ndesc, 1024));
    bank_append_kernel<<<grid, 256, 0, s>>>(src_rows, bank, static_cast<uint32_t>(D), reinterpret_cast<const AppendDesc *>(desc), ndesc);
    return check_launch("bank_append");
}

static int infonce_depth()
{
    static const int d = [] { const char *e = getenv("U2PL_INFONCE_DEPTH"); return e ? atoi(e) : 2; }();   // two rows in flight: 75 vs 80 us (config 4)
    return d;
}

extern "C" int u2pl_infonce_forward(const float *rep, int64_t sn, int64_t sd, int64_t sp,
                                    int64_t P, int64_t D, int64_t hw,
                                    const uint32_t *an_bits, const uint32_t *blockoff_an,
                                    const int32_t *act_class, const int32_t *a_ord, const int32_t *neg_rows,
                                    const float *proto, const float *bank,
                                    int nact, int nq, int nneg, float temperature, int valid_seg,
                                    float *loss_q, float *grad_rows, int32_t *anchor_pix, float *loss, void *stream)
{
    if (D <= 0 || D > kMaxD || D % 4 != 0) return bad_arg("infonce_forward: D must be a multiple of 4, <= 256");
    if (nact <= 0 || nq <= 0 || nneg < 0 || valid_seg <= 0) return bad_arg("infonce_forward: empty problem");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    InfoNceArgs a;
    a.rep = rep; a.sn = sn; a.sd = sd; a.sp = sp;
    a.hw = static_cast<uint32_t>(hw); a.P = static_cast<uint32_t>(P); a.D = static_cast<uint32_t>(D);
    a.nb = static_cast<uint32_t>((P + kBlk - 1) / kBlk);
    a.an_bits = an_bits; a.blockoff_an = blockoff_an; a.act_class = act_class; a.a_ord = a_ord; a.neg_rows = neg_rows;
    a.proto = proto; a.bank = bank; a.class_bank = nullptr; a.nact = nact; a.nq = nq; a.nneg = nneg;
    a.inv_temp = 1.0f / temperature;
    a.scale = 1.0f / (static_cast<float>(nq) * static_cast<float>(valid_seg));
    a.loss_q = loss_q; a.grad_rows = grad_rows; a.anchor_pix = anchor_pix;
    const int warps = nact * nq;
    const int depth = infonce_depth();
    // (a persistent grid of fewer warps walking the queries in class order was tried to keep the rows in flight inside the
    // TLB's reach: 4-12x slower -- each query has ~7 us of dependent look-ups in front of its rows, which only many resident
    // warps hide; profiles/r02_contra_bench.txt)
    if (depth == 4) infonce_fwd_pipelined_kernel<false, 4><<<(warps + 3) / 4, 128, 0, s>>>(a);
    else if (depth == 2) infonce_fwd_pipelined_kernel<false, 2><<<(warps + 3) / 4, 128, 0, s>>>(a);
    else infonce_fwd_kernel<false><<<(warps + 3) / 4, 128, 0, s>>>(a);
    infonce_loss_kernel<<<1, 256, 0, s>>>(loss_q, warps, a.scale, loss);
    return check_launch("infonce_forward", 2);
}

extern "C" int u2pl_infonce_forward_sharded(const float *rep, int64_t sn, int64_t sd, int64_t sp,
                                            int64_t P, int64_t D, int64_t hw,
                                            const uint32_t *an_bits, const uint32_t *blockoff_an,
                                            const int32_t *act_class, const int32_t *a_ord, const int32_t *neg_rows,
                                            const float *proto, const float *const *class_bank,
                                            int nact, int nq, int nneg, float temperature, int valid_seg,
                                            float *loss_q, float *grad_rows, int32_t *anchor_pix, float *loss, void *stream)
{
    if (D <= 0 || D > kMaxD || D % 4 != 0) return bad_arg("infonce_forward_sharded: D must be a multiple of 4, <= 256");
    if (nact <= 0 || nq <= 0 || nneg < 0 || valid_seg <= 0) return bad_arg("infonce_forward_sharded: empty problem");
    if (!class_bank) return bad_arg("infonce_forward_sharded: class_bank is NULL");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    InfoNceArgs a;
    a.rep = rep; a.sn = sn; a.sd = sd; a.sp = sp;
    a.hw = static_cast<uint32_t>(hw); a.P = static_cast<uint32_t>(P); a.D = static_cast<uint32_t>(D);
    a.nb = static_cast<uint32_t>((P + kBlk - 1) / kBlk);
    a.an_bits = an_bits; a.blockoff_an = blockoff_an; a.act_class = act_class; a.a_ord = a_ord; a.neg_rows = neg_rows;
    a.proto = proto; a.bank = nullptr; a.class_bank = class_bank; a.nact = nact; a.nq = nq; a.nneg = nneg;
    a.inv_temp = 1.0f / temperature;
    a.scale = 1.0f / (static_cast<float>(nq) * static_cast<float>(valid_seg));
    a.loss_q = loss_q; a.grad_rows = grad_rows; a.anchor_pix = anchor_pix;
    const int warps = nact * nq;
    const int depth = infonce_depth();
    if (depth == 4) infonce_fwd_pipelined_kernel<true, 4><<<(warps + 3) / 4, 128, 0, s>>>(a);
    else if (depth == 2) infonce_fwd_pipelined_kernel<true, 2><<<(warps + 3) / 4, 128, 0, s>>>(a);
    else infonce_fwd_kernel<true><<<(warps + 3) / 4, 128, 0, s>>>(a);
    infonce_loss_kernel<<<1, 256, 0, s>>>(loss_q, warps, a.scale, loss);
    return check_launch("infonce_forward_sharded", 2);
}

// ------------------------------------------------------------------ peer-mapped bank shards (CUDA IPC)
// One cudaMalloc'd shard per rank, exported with cudaIpcGetMemHandle, opened by every other rank of the same box with
// cudaIpcOpenMemHandle (peer access over NVLink is enabled lazily by the driver).  The handles travel through the
// host-side process group (u2pl_b200/bank.py); nothing here depends on torch or NCCL.
extern "C" int u2pl_shard_alloc(int64_t bytes, void **dptr, unsigned char *handle64)
{
    if (bytes <= 0 || !dptr || !handle64) return bad_arg("shard_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, static_cast<size_t>(bytes));
    if (e == cudaSuccess) e = cudaMemset(p, 0, static_cast<size_t>(bytes));
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { if (p) cudaFree(p); set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
    memcpy(handle64, &h, 64);
    *dptr = p;
    return 0;
}

extern "C" int u2pl_shard_open(const unsigned char *handle64, void **dptr)
{
    if (!handle64 || !dptr) return bad_arg("shard_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
    return 0;
}

extern "C" int u2pl_shard_close(void *dptr, int owned)
{
    if (!dptr) return 0;
    cudaError_t e = owned ? cudaFree(dptr) : cudaIpcCloseMemHandle(dptr);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
    return 0;
}

extern "C" int u2pl_infonce_backward(const float *grad_rows, const int32_t *anchor_pix, int nrows,
                                     int64_t D, int64_t hw, int64_t sn, int64_t sd, int64_t sp,
                                     const float *upstream, float *grad_rep, void *stream)
{
    if (nrows <= 0) return 0;
    infonce_bwd_kernel<<<(nrows + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        grad_rows, anchor_pix, nrows, static_cast<uint32_t>(D), static_cast<uint32_t>(hw), sn, sd, sp, upstream, grad_rep);
    return check_launch("infonce_backward");
}
