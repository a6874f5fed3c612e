// conv_tc.cu -- stride-1 "same" convolution (1x1, or RxS with dilation) of a channels-last bf16 tensor as an
// IMPLICIT GEMM on the tcgen05 tensor cores, with the eval-mode BatchNorm scale/shift, the residual add and
// the ReLU of a ResNet block folded into the epilogue:
//
//     out[n,h,w,co] = act( (sum_{r,s,ci} x[n, h+(r-R/2)d, w+(s-S/2)d, ci] * wgt[co,r,s,ci]) * scale[co] + shift[co]
//                          + residual[n,h,w,co] )
//
// It is the contraction behind every stride-1 convolution of the network (resnet.py:25-41 conv3x3/conv1x1,
// base.py:38-75 ASPP branches d=12/24/36, decoder.py:60-113 heads), used for the teacher's eval-mode
// pseudo-label forward (train_semi.py:318-319).
//
// No im2col buffer exists anywhere: the A operand of tap (r,s), channel block kb, is ONE 4-D TMA box
// {64 channels, TW, TH, 1 image} of x fetched at spatial offset ((r-R/2)d, (s-S/2)d) -- TMA zero-fills the
// part of the box that falls outside the image (negative or >= H/W coordinates), which is exactly the zero
// padding of the convolution -- and lands in shared memory as the same 128-row x 128-byte SWIZZLE_128B K-major
// tile the flat GEMM uses (row = i*TW + j).  The weights are read as the row-major [Cout, R*S*Cin] matrix a
// channels-last weight tensor already is.  All taps and channel blocks accumulate into one TMEM tile.
// A 1x1 convolution uses the same kernel with the tensor viewed as {Cin, N*H*W, 1, 1} and a {64,128,1,1} box
// (no spatial tiling, no tile waste).
//
// CTA layout, pipeline and barriers are those of gemm_bf16_tn_persistent_kernel (gemm_tc.cu): warp 0 TMA
// producer, warp 1 TMEM allocation + single-thread MMA issue, warps 2-5 epilogue, 4-stage smem ring, two
// TMEM accumulator stages, one persistent CTA per SM looping over (pixel tile, channel tile) pairs.
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "conv_tc2.cuh"
#include "conv_tc3.cuh"

namespace u2pl {

namespace convtc {
constexpr int kBM = 128, kBK = 64, kStages = 4;
constexpr int kTileABytes = kBM * kBK * 2;
constexpr int kThreads = 192;
}  // namespace convtc
// The output-channel tile kBN is a template parameter: 128 (layers with Cout <= 128) or 256.  With a 128x128 tile one
// k-block costs 256 tensor-core cycles but needs 32 KB of shared-memory fill -- right at the per-SM L2->smem rate, which
// is why the flat 128x128 GEMM stops at ~0.6 of cuBLAS; a 128x256 tile doubles the MMA work per A byte (48 KB per 512
// cycles) and uses all 512 TMEM columns for the two accumulator stages.

struct ConvParams {
    int Nimg, H, W, Cin, Cout;        // logical NHWC geometry the tensor map was built from (flat 1x1: Nimg=H=1, W=N*H*W)
    int R, S, dil;                    // taps and dilation (R, S odd)
    int log2_tw;                      // pixel tile = TH x TW with TH*TW == 128, TW = 1 << log2_tw
    int tiles_h, tiles_w;             // pixel tiles per image
    const float *scale, *shift;       // [Cout] or null
    const __nv_bfloat16 *residual;    // same layout as D, or null
    int relu;
    __nv_bfloat16 *D;
    float *stat_part;                 // kStats only: [pixel tiles][2][Cout] per-tile sum / sum of squares of the STORED values
    const float *in_scale, *in_shift; // kXform only: [Cin] affine applied to x while it sits in shared memory
    int in_relu;                      //              followed by max(., 0)
};

// kXform: the input tensor is the RAW output of the previous convolution and its BatchNorm (+ReLU) is applied here,
// on the A tile, between the TMA load and the MMA: four extra warps rewrite each stage in place (16-byte chunks, the
// SWIZZLE_128B pattern followed by hand), make the writes visible to the async proxy and hand the stage to the MMA warp
// through `ready`.  Pixels outside the image stay zero (the padding applies to the activated tensor).  The K loop runs
// channel-block-outer / tap-inner in this variant so that a thread's 8 channels -- its 16 scale/shift values -- stay in
// registers for nine consecutive stages.
template <int kBN, bool kStats, bool kXform>
__global__ void __launch_bounds__(kXform ? convtc::kThreads + 128 : convtc::kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, ConvParams p)
{
    using namespace convtc;
    constexpr int kTileBBytes = kBN * kBK * 2, kTmemCols = kBN;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *sA = smem, *sB = smem + kStages * kTileABytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + kStages * kTileBBytes);
    uint64_t *empty = full + kStages;
    uint64_t *tmem_full = empty + kStages;            // [2]
    uint64_t *tmem_empty = tmem_full + 2;             // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *s_par = reinterpret_cast<float *>(tmem_slot + 2);      // [2 stages][scale kBN | shift kBN]
    float *s_tr = s_par + 4 * kBN;                                // kStats: [4 warps][32 rows][33] transpose buffers
    float *s_red = s_tr + (kStats ? 4 * 32 * 33 : 0);             // kStats: [4 warps][kBN][2] per-warp column sums
    uint64_t *ready = reinterpret_cast<uint64_t *>(s_red + (kStats ? 4 * kBN * 2 : 0));   // kXform: [kStages] transform -> MMA

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tw = 1 << p.log2_tw, th = kBM >> p.log2_tw;
    const int tiles_img = p.tiles_h * p.tiles_w;
    const int tiles_m = p.Nimg * tiles_img, tiles_n = (p.Cout + kBN - 1) / kBN;
    const int num_tiles = tiles_m * tiles_n;
    const int kb_per_tap = (p.Cin + kBK - 1) / kBK;
    const int nkb = p.R * p.S * kb_per_tap;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, 4); }
        if (kXform)
            for (int s = 0; s < kStages; ++s) mbar_init(ready + s, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                              // ---------------- TMA producer
            uint32_t it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int tm = t / tiles_n, n0 = (t % tiles_n) * kBN;
                const int img = tm / tiles_img, rem = tm % tiles_img;
                const int h0 = (rem / p.tiles_w) * th, w0 = (rem % p.tiles_w) * tw;
                const int taps = p.R * p.S;
                for (int step = 0; step < nkb; ++step, ++it) {
                    {   // kXform: channel block outer, tap inner; otherwise tap outer, channel block inner
                        const int tap = kXform ? step % taps : step / kb_per_tap;
                        const int kb = kXform ? step / taps : step % kb_per_tap;
                        const int dh = (tap / p.S - p.R / 2) * p.dil, dw = (tap % p.S - p.S / 2) * p.dil;
                        const int s = it % kStages;
                        mbar_wait(empty + s, ((it / kStages) & 1) ^ 1);
                        mbar_expect_tx(full + s, kTileABytes + kTileBBytes);
                        tma_load_4d(sA + s * kTileABytes, &map_x, full + s, kb * kBK, w0 + dw, h0 + dh, img);
                        // weight columns of this tap start at tap*Cin; when Cin % 64 != 0 the last block of a tap also
                        // fetches the first columns of the next tap, which meet zero-filled A channels (product 0)
                        tma_load_2d(sB + s * kTileBBytes, &map_w, full + s, tap * p.Cin + kb * kBK, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                              // ---------------- MMA issuer
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kBN >> 3) << 17) |
                                   (static_cast<uint32_t>(kBM >> 4) << 24);
            uint32_t it = 0, tile_i = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tile_i) {
                const uint32_t acc = tile_i & 1, use = tile_i >> 1;
                mbar_wait(tmem_empty + acc, (use & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * kTmemCols;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(kXform ? ready + s : full + s, (it / kStages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = smem_u32(sA + s * kTileABytes), b0 = smem_u32(sB + s * kTileBBytes);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                        umma_f16(d_tmem, smem_desc_sw128(a0 + 32 * k), smem_desc_sw128(b0 + 32 * k), idesc, (kb | k) ? 1u : 0u);
                    umma_commit(empty + s);
                }
                umma_commit(tmem_full + acc);
            }
        }
    } else if (kXform && warp >= 6) {                 // ---------------- A-tile transform (warps 6..9)
        const int tid = threadIdx.x - kThreads;       // 0..127: chunk = 8 channels, rows rbase + 16 i
        const int chunk = tid & 7, rbase = tid >> 3;
        const int taps = p.R * p.S;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int tm = t / tiles_n;
            const int img = tm / tiles_img, rem = tm % tiles_img;
            const int h0 = (rem / p.tiles_w) * th, w0 = (rem % p.tiles_w) * tw;
            (void)img;
            for (int kb = 0; kb < kb_per_tap; ++kb) {
                float sc[8], sh[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {         // channels beyond Cin (zero-filled by TMA) must stay zero
                    const int c = kb * kBK + chunk * 8 + e;
                    sc[e] = (c < p.Cin && p.in_scale) ? __ldg(p.in_scale + c) : (c < p.Cin ? 1.0f : 0.0f);
                    sh[e] = (c < p.Cin && p.in_shift) ? __ldg(p.in_shift + c) : 0.0f;
                }
                for (int tap = 0; tap < taps; ++tap, ++it) {
                    const int dh = (tap / p.S - p.R / 2) * p.dil, dw = (tap % p.S - p.S / 2) * p.dil;
                    const int s = it % kStages;
                    mbar_wait(full + s, (it / kStages) & 1);
                    uint8_t *tile = sA + s * kTileABytes;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = rbase + 16 * i;
                        const int hh = h0 + dh + (r >> p.log2_tw), ww = w0 + dw + (r & (tw - 1));
                        if (static_cast<unsigned>(hh) < static_cast<unsigned>(p.H) && static_cast<unsigned>(ww) < static_cast<unsigned>(p.W)) {
                            uint4 *cp = reinterpret_cast<uint4 *>(tile + r * 128 + ((chunk ^ (r & 7)) << 4));
                            uint4 v = *cp;
                            __nv_bfloat162 *hv = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float2 f = __bfloat1622float2(hv[e]);
                                f.x = fmaf(f.x, sc[2 * e], sh[2 * e]);
                                f.y = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
                                if (p.in_relu) { f.x = fmaxf(f.x, 0.0f); f.y = fmaxf(f.y, 0.0f); }
                                hv[e] = __floats2bfloat162_rn(f.x, f.y);
                            }
                            *cp = v;
                        }                             // rows outside the image were zero-filled and stay zero (padding)
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ready + s);
                }
            }
        }
    } else if (warp >= 2 && warp < 6) {               // ---------------- epilogue (warps 2..5)
        const int q = warp & 3;
        const bool affine = p.scale != nullptr || p.shift != nullptr;
        const int m = q * 32 + lane, ti = m >> p.log2_tw, tj = m & (tw - 1);
        uint32_t tile_i = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tile_i) {
            const int tm = t / tiles_n, n0 = (t % tiles_n) * kBN;
            const int img = tm / tiles_img, rem = tm % tiles_img;
            const int h = (rem / p.tiles_w) * th + ti, w = (rem % p.tiles_w) * tw + tj;
            const bool live = h < p.H && w < p.W;
            const size_t pix = (static_cast<size_t>(img) * p.H + h) * p.W + w;
            const uint32_t acc = tile_i & 1, use = tile_i >> 1;
            float *s_scale = s_par + acc * 2 * kBN, *s_shift = s_scale + kBN;
            if (affine) {
                for (int e = threadIdx.x - 64; e < kBN; e += 128) {
                    const int c = n0 + e;
                    s_scale[e] = (p.scale && c < p.Cout) ? __ldg(p.scale + c) : 1.0f;
                    s_shift[e] = (p.shift && c < p.Cout) ? __ldg(p.shift + c) : 0.0f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(tmem_full + acc, use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int j = 0; j < kBN / 32; ++j) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + acc * kTmemCols + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(j * 32), r);
                if (j == kBN / 32 - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty + acc);
                }
                const int c0 = n0 + j * 32;
                float *trow = s_tr + (q * 32 + lane) * 33;        // this thread's row of its warp's transpose buffer
                if (kStats && !(live && c0 + 32 <= p.Cout)) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) trow[e] = 0.0f;  // dead rows / channels beyond Cout contribute nothing
                }
                if (live) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int c = c0 + g * 8;
                        if (c < p.Cout) {             // Cout % 8 == 0
                            const size_t off = pix * p.Cout + c;
                            float v[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                v[e] = __uint_as_float(r[g * 8 + e]);
                                if (affine) v[e] = fmaf(v[e], s_scale[j * 32 + g * 8 + e], s_shift[j * 32 + g * 8 + e]);
                            }
                            if (p.residual) {
                                const uint4 rr = __ldg(reinterpret_cast<const uint4 *>(p.residual + off));
                                const __nv_bfloat162 *rh = reinterpret_cast<const __nv_bfloat162 *>(&rr);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 f = __bfloat1622float2(rh[e]);
                                    v[2 * e] += f.x;
                                    v[2 * e + 1] += f.y;
                                }
                            }
                            if (p.relu) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.0f);
                            }
                            uint4 o;
                            __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
                            for (int e = 0; e < 4; ++e) oh[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
                            *reinterpret_cast<uint4 *>(p.D + off) = o;
                            if (kStats) {             // statistics of the values as stored (bf16-rounded), like bn_stats
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 f = __bfloat1622float2(oh[e]);
                                    trow[g * 8 + 2 * e] = f.x;
                                    trow[g * 8 + 2 * e + 1] = f.y;
                                }
                            }
                        }
                    }
                }
                if (kStats) {                         // column sums over this warp's 32 rows: lane c owns column c0 + c
                    __syncwarp();
                    const float *tcol = s_tr + q * 32 * 33 + lane;
                    float cs = 0.0f, cq = 0.0f;
#pragma unroll
                    for (int r2 = 0; r2 < 32; ++r2) {
                        const float xv = tcol[r2 * 33];
                        cs += xv;
                        cq = fmaf(xv, xv, cq);
                    }
                    s_red[(q * kBN + j * 32 + lane) * 2] = cs;
                    s_red[(q * kBN + j * 32 + lane) * 2 + 1] = cq;
                    __syncwarp();
                }
            }
            if (kStats) {                             // 4 warps -> one partial per (pixel tile, channel)
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int e = threadIdx.x - 64; e < kBN; e += 128) {
                    const int c = n0 + e;
                    if (c < p.Cout) {
                        float a = 0.0f, b = 0.0f;
#pragma unroll
                        for (int k2 = 0; k2 < 4; ++k2) { a += s_red[(k2 * kBN + e) * 2]; b += s_red[(k2 * kBN + e) * 2 + 1]; }
                        p.stat_part[(static_cast<size_t>(tm) * 2) * p.Cout + c] = a;
                        p.stat_part[(static_cast<size_t>(tm) * 2 + 1) * p.Cout + c] = b;
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * kTmemCols) : "memory");
}

// x viewed as {C, W, H, N} (innermost first), bf16, dense NHWC; box {64, tw, th, 1}; 128-byte swizzle; zero OOB fill
static bool make_map_nhwc(CUtensorMap *map, const void *base, int64_t n, int64_t h, int64_t w, int64_t c, int th, int tw)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(n)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(c) * 2, static_cast<cuuint64_t>(w) * c * 2, static_cast<cuuint64_t>(h) * w * c * 2};
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(convtc::kBK), static_cast<cuuint32_t>(tw), static_cast<cuuint32_t>(th), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool make_map_weight(CUtensorMap *map, const void *base, int64_t cout, int64_t ktot, int bn)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(cout)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ktot) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(convtc::kBK), static_cast<cuuint32_t>(bn)};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace u2pl

using namespace u2pl;

template <int kBN, bool kStats, bool kXform>
static cudaError_t conv_configure(size_t smem)
{
    return cudaFuncSetAttribute(conv_tc_kernel<kBN, kStats, kXform>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
}

template <int kBN, bool kStats, bool kXform>
static void conv_run(unsigned grid, size_t smem, cudaStream_t st, const CUtensorMap &mx, const CUtensorMap &mw, const ConvParams &p)
{
    conv_tc_kernel<kBN, kStats, kXform><<<grid, kXform ? convtc::kThreads + 128 : convtc::kThreads, smem, st>>>(mx, mw, p);
}

// Kernel selection: U2PL_CONV_V unset / 3 = flat-tile kernel (conv_tc3.cu) for everything but the operand-transform
// variant; 1 = the patch-tiled 1-CTA kernel below for everything; 2 = the CTA-pair kernel where eligible.
static bool flat_kernel(bool xform)
{
    static const int forced = [] { const char *e = getenv("U2PL_CONV_V"); return e ? atoi(e) : 0; }();
    return !xform && forced != 1 && forced != 2;
}

static int conv_launch(const void *x, const void *wgt, void *out, int64_t n, int64_t h, int64_t w, int64_t cin, int64_t cout,
                       int ksize, int dilation, const float *in_scale, const float *in_shift, int in_relu,
                       const float *scale, const float *shift, const void *residual, int relu,
                       float *stat_part, const char *what, void *stream)
{
    using namespace convtc;
    if (n <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0 || (cin % 8) || (cout % 8) || (ksize != 1 && ksize != 3) || dilation < 1)
        return bad_arg("conv_bf16_nhwc: need Cin % 8 == 0, Cout % 8 == 0, ksize in {1,3}, dilation >= 1");
    if (n * h * w >= (1LL << 31) || static_cast<int64_t>(ksize) * ksize * cin >= (1LL << 31))
        return bad_arg("conv_bf16_nhwc: tensor too large for 32-bit TMA coordinates");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wgt) | reinterpret_cast<uintptr_t>(out) |
         reinterpret_cast<uintptr_t>(residual)) & 15)
        return bad_arg("conv_bf16_nhwc: operands must be 16-byte aligned");
    const bool xform_req = in_scale != nullptr || in_shift != nullptr || in_relu != 0;
    if (conv_tc2_eligible(cout, xform_req))           // CTA-pair kernel (conv_tc2.cu), opt-in
        return conv_tc2_launch(x, wgt, out, n, h, w, cin, cout, ksize, dilation, scale, shift, residual, relu, stat_part, what, stream);
    if (flat_kernel(xform_req))                       // default: flat pixel tiles, im2col TMA, TMA-store epilogue (conv_tc3.cu)
        return conv_tc3_launch(x, wgt, out, n, h, w, cin, cout, ksize, dilation, scale, shift, residual, relu, stat_part, what, stream);
    ConvParams p;
    CUtensorMap mx, mw;
    bool ok;
    if (ksize == 1) {                                 // flat: every 128 consecutive pixels are one tile
        p.Nimg = 1; p.H = 1; p.W = static_cast<int>(n * h * w);
        p.log2_tw = 7; p.tiles_h = 1; p.tiles_w = (p.W + kBM - 1) / kBM;
        ok = make_map_nhwc(&mx, x, 1, 1, n * h * w, cin, 1, kBM);
    } else {                                          // 8 x 16 pixel tiles inside each image
        p.Nimg = static_cast<int>(n); p.H = static_cast<int>(h); p.W = static_cast<int>(w);
        p.log2_tw = 4; p.tiles_h = (p.H + 7) / 8; p.tiles_w = (p.W + 15) / 16;
        ok = make_map_nhwc(&mx, x, n, h, w, cin, 8, 16);
    }
    // channel tile: 256 when the layer has more than 128 output channels (U2PL_CONV_BN=128 forces the narrow tile)
    static const int forced_bn = [] { const char *e = getenv("U2PL_CONV_BN"); return e ? atoi(e) : 0; }();
    const int bn = (forced_bn == 128 || forced_bn == 256) ? forced_bn : (cout > 128 ? 256 : 128);
    ok = ok && make_map_weight(&mw, wgt, cout, static_cast<int64_t>(ksize) * ksize * cin, bn);
    if (!ok) { set_error("conv_bf16_nhwc: cuTensorMapEncodeTiled failed"); return U2PL_E_BADARG; }
    p.Cin = static_cast<int>(cin); p.Cout = static_cast<int>(cout);
    p.R = p.S = ksize; p.dil = dilation;
    p.scale = scale; p.shift = shift; p.residual = static_cast<const __nv_bfloat16 *>(residual); p.relu = relu;
    p.D = static_cast<__nv_bfloat16 *>(out);
    p.stat_part = stat_part;
    p.in_scale = in_scale; p.in_shift = in_shift; p.in_relu = in_relu;
    auto smem_for = [](int b, bool stats) {                      // (+64: the kXform `ready` barriers)
        return static_cast<size_t>(kStages) * (kTileABytes + b * kBK * 2) + 1024 + 256 + 64 + 4 * b * sizeof(float) +
               (stats ? (4 * 32 * 33 + 4 * b * 2) * sizeof(float) : 0);
    };
    static bool configured = false;
    if (!configured) {
        cudaError_t e = conv_configure<128, false, false>(smem_for(128, false));
        if (e == cudaSuccess) e = conv_configure<256, false, false>(smem_for(256, false));
        if (e == cudaSuccess) e = conv_configure<128, true, false>(smem_for(128, true));
        if (e == cudaSuccess) e = conv_configure<256, true, false>(smem_for(256, true));
        if (e == cudaSuccess) e = conv_configure<128, false, true>(smem_for(128, false));
        if (e == cudaSuccess) e = conv_configure<256, false, true>(smem_for(256, false));
        if (e == cudaSuccess) e = conv_configure<128, true, true>(smem_for(128, true));
        if (e == cudaSuccess) e = conv_configure<256, true, true>(smem_for(256, true));
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
        configured = true;
    }
    const bool stats = stat_part != nullptr;
    const bool xform = in_scale != nullptr || in_shift != nullptr || in_relu != 0;
    const size_t smem = smem_for(bn, stats);
    const long long tiles = static_cast<long long>(p.Nimg) * p.tiles_h * p.tiles_w * ((cout + bn - 1) / bn);
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned g = static_cast<unsigned>(tiles < sms ? tiles : sms);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int variant = (bn == 256 ? 4 : 0) | (stats ? 2 : 0) | (xform ? 1 : 0);
    switch (variant) {
    case 0: conv_run<128, false, false>(g, smem, st, mx, mw, p); break;
    case 1: conv_run<128, false, true>(g, smem, st, mx, mw, p); break;
    case 2: conv_run<128, true, false>(g, smem, st, mx, mw, p); break;
    case 3: conv_run<128, true, true>(g, smem, st, mx, mw, p); break;
    case 4: conv_run<256, false, false>(g, smem, st, mx, mw, p); break;
    case 5: conv_run<256, false, true>(g, smem, st, mx, mw, p); break;
    case 6: conv_run<256, true, false>(g, smem, st, mx, mw, p); break;
    default: conv_run<256, true, true>(g, smem, st, mx, mw, p); break;
    }
    return check_launch(what);
}

extern "C" int u2pl_conv_bf16_nhwc(const void *x, const void *wgt, void *out, int64_t n, int64_t h, int64_t w, int64_t cin,
                                   int64_t cout, int ksize, int dilation, const float *scale, const float *shift,
                                   const void *residual, int relu, void *stream)
{
    return conv_launch(x, wgt, out, n, h, w, cin, cout, ksize, dilation, nullptr, nullptr, 0, scale, shift, residual, relu, nullptr,
                       "conv_bf16_nhwc", stream);
}

static int64_t stat_parts_1cta(int64_t n, int64_t h, int64_t w, int ksize)
{
    if (ksize == 1) return (n * h * w + convtc::kBM - 1) / convtc::kBM;
    return n * ((h + 7) / 8) * ((w + 15) / 16);
}

// rows of the caller's partial-sum buffer: the largest of the kernels' needs (the pair kernel writes two per pixel tile)
extern "C" int64_t u2pl_conv_stat_parts(int64_t n, int64_t h, int64_t w, int ksize) { return conv_tc2_stat_parts(n, h, w, ksize); }

static int stat_parts_used(int64_t n, int64_t h, int64_t w, int64_t cout, int ksize, bool xform)
{
    if (conv_tc2_eligible(cout, xform)) return static_cast<int>(conv_tc2_stat_parts(n, h, w, ksize));
    if (flat_kernel(xform)) return static_cast<int>(conv_tc3_stat_parts(n, h, w));
    return static_cast<int>(stat_parts_1cta(n, h, w, ksize));
}

extern "C" int u2pl_conv_bf16_nhwc_stats(const void *x, const void *wgt, void *out, int64_t n, int64_t h, int64_t w,
                                         int64_t cin, int64_t cout, int ksize, int dilation, float *stat_part, float *sums,
                                         void *stream)
{
    if (!stat_part || !sums) return bad_arg("conv_bf16_nhwc_stats: stat_part and sums are required");
    int rc = conv_launch(x, wgt, out, n, h, w, cin, cout, ksize, dilation, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, stat_part,
                         "conv_bf16_nhwc_stats", stream);
    if (rc != 0) return rc;
    return bn_reduce_parts(stat_part, stat_parts_used(n, h, w, cout, ksize, false), static_cast<int>(2 * cout), sums, stream);
}

extern "C" int u2pl_conv_bf16_nhwc_ex(const void *x, const void *wgt, void *out, int64_t n, int64_t h, int64_t w, int64_t cin,
                                      int64_t cout, int ksize, int dilation, const float *in_scale, const float *in_shift,
                                      int in_relu, const float *scale, const float *shift, const void *residual, int relu,
                                      float *stat_part, float *sums, void *stream)
{
    if ((stat_part == nullptr) != (sums == nullptr)) return bad_arg("conv_bf16_nhwc_ex: stat_part and sums go together");
    int rc = conv_launch(x, wgt, out, n, h, w, cin, cout, ksize, dilation, in_scale, in_shift, in_relu, scale, shift, residual, relu,
                         stat_part, "conv_bf16_nhwc_ex", stream);
    if (rc != 0 || !stat_part) return rc;
    return bn_reduce_parts(stat_part, stat_parts_used(n, h, w, cout, ksize, in_scale != nullptr || in_shift != nullptr || in_relu != 0),
                           static_cast<int>(2 * cout), sums, stream);
}
