// conv_tc3.cu -- FLAT-TILE implicit-GEMM convolution on tcgen05: the default forward / data-gradient kernel.
//
// Same contract as conv_tc.cu (stride-1 "same" 1x1 / 3x3 dilated convolution of a channels-last bf16 tensor, epilogue
// act(conv * scale[co] + shift[co] + residual), optional per-channel statistics of the stored values) and the same CTA
// anatomy (warp 0 TMA producer, warp 1 TMEM + single-thread tcgen05.mma, warps 2..5 epilogue, smem ring, two TMEM
// accumulator stages, one persistent CTA per SM).  Two things differ, both measured as the gap to cuDNN on B200
// (profiles/r02_conv_selftest_perf_v1.txt: raw MMA rate 1.5 PFLOP/s, useful rate 1.05-1.2):
//
//   * PIXEL TILES ARE FLAT.  A tile is 128 CONSECUTIVE pixels of the flattened N*H*W index, not an 8x16 patch of one
//     image: on the network's 65x65 / 129x129 maps the patch tiling spent 27 % / 12 % of all MMA rows on pixels outside
//     the image.  The A operand of tap (r,s) / channel block kb is ONE TMA load in IM2COL mode
//     (cuTensorMapEncodeIm2col: bounding box = the image shifted by -padding, 64 channels per pixel, 128 pixels per
//     column; the tap's offset {s*dil, r*dil} rides in the instruction): the TMA unit walks W, then H, then N from the
//     tile's first pixel, zero-fills what falls outside the image (= the convolution's zero padding) and beyond the last
//     image (= the M tail), and delivers the same 128-row SWIZZLE_128B K-major tile as before.  No im2col buffer exists.
//   * THE EPILOGUE LEAVES THROUGH SHARED MEMORY AND TMA.  Each epilogue warp owns 32 accumulator rows and a private,
//     double-buffered 4 KB staging slice: tcgen05.ld (64 columns) -> scale/shift (+ the residual rows, which a TMA load
//     brought INTO the slice one chunk ahead) -> ReLU -> bf16 rows in SWIZZLE_128B order -> one cp.async.bulk.tensor
//     store of the {64 ch, 32 px} box.  Global traffic is whole 128-byte lines in both directions (the per-thread
//     16-byte stores of conv_tc.cu ran the memory-bound 256 -> 1024 1x1 layers at 6 us per tile instead of ~3), the
//     M / Cout tails are clipped by the tensor map, and no warp ever waits for another one.
//   * statistics (train mode): column sums / sums of squares of the staged bf16 values, combined across the four warps
//     in shared memory in a fixed order: one partial row per pixel tile.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "conv_tc3.cuh"

namespace u2pl {
namespace convtc3 {

constexpr int kBM = 128, kBK = 64;
constexpr int kEpiWarps = 8;                          // two per TMEM lane quarter, one per column half of the tile
constexpr int kThreads = 64 + 32 * kEpiWarps;         // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kTileABytes = kBM * kBK * 2;            // 16 KB
// kCh = channels per staged store: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B; leaves room for a
// fourth pipeline stage of the 256-wide tile)

struct Params {
    int M;                            // N*H*W pixels
    int H, W, Cin, Cout;
    int taps, S, dil;                 // taps = R*S (1 or 9)
    const float *scale, *shift;       // [Cout] or null
    int has_residual, relu;
    float *stat_part;                 // kStats: [ceil(M/128)][2][Cout]
};

template <int kBN, int kStages, int kCh>
struct Smem {
    static constexpr int tileB = kBN * kBK * 2;
    static constexpr int slice = 32 * kCh * 2;                         // 32 rows of one warp
    static constexpr int a = 0;
    static constexpr int b = kStages * kTileABytes;
    static constexpr int stage = b + kStages * tileB;                  // [8 warps][2 buffers][slice]
    // one auxiliary region: [scale kBN | shift kBN] floats when there is an affine epilogue, else (statistics variant,
    // which has none) [2 halves][4 quarters][kCh][2] floats of per-warp column sums
    static constexpr int aux = stage + kEpiWarps * 2 * slice;
    static constexpr int aux_bytes = (2 * kBN * 4 > 2 * 4 * kCh * 2 * 4) ? 2 * kBN * 4 : 2 * 4 * kCh * 2 * 4;
    static constexpr int bars = aux + aux_bytes;
    static constexpr int total = bars + 256;
    static_assert(total <= 232448, "shared memory budget of one CTA");
};

__device__ __forceinline__ void tma_load_im2col_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c, int w, int h, int n,
                                                   uint16_t off_w, uint16_t off_h)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}

template <int kBN, int kStages, int kCh, bool kStats>
__global__ void __launch_bounds__(kThreads, 1)
conv_flat_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res, Params p)
{
    using L = Smem<kBN, kStages, kCh>;
    constexpr int kTileBBytes = L::tileB, kTmemCols = kBN;
    constexpr int kChunk = kCh, kSliceBytes = L::slice, kRowB = kCh * 2, kGroups = kCh / 8;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem + L::a, *sB = smem + L::b, *sStage = smem + L::stage;
    float *s_par = reinterpret_cast<float *>(smem + L::aux);
    float *s_red = reinterpret_cast<float *>(smem + L::aux);          // (never both: the statistics variant has no affine)
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L::bars);
    uint64_t *empty = full + kStages;
    uint64_t *tmem_full = empty + kStages;            // [2]
    uint64_t *tmem_empty = tmem_full + 2;             // [2]
    uint64_t *res_full = tmem_empty + 2;              // [8 warps][2 buffers]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(res_full + 2 * kEpiWarps);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (p.M + kBM - 1) / kBM, tiles_n = (p.Cout + kBN - 1) / kBN;
    const int num_tiles = tiles_m * tiles_n;
    const int kb_per_tap = (p.Cin + kBK - 1) / kBK;
    const int nkb = p.taps * kb_per_tap;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();         // SWIZZLE_128B tiles need a 1024-byte aligned base (no slack is budgeted)
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
        if (p.has_residual) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_res) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, kEpiWarps); }
        for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(res_full + i, 1);
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
            const int hw = p.H * p.W;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int tm = t / tiles_n, n0 = (t % tiles_n) * kBN;
                const int m0 = tm * kBM;
                const int img = m0 / hw, rem = m0 - img * hw;
                const int h0 = rem / p.W, w0 = rem - h0 * p.W;
                const int pad = (p.taps == 1) ? 0 : p.dil;
                for (int step = 0; step < nkb; ++step, ++it) {
                    const int tap = step / kb_per_tap, kb = step - tap * kb_per_tap;
                    const int s = it % kStages;
                    mbar_wait(empty + s, ((it / kStages) & 1) ^ 1);
                    mbar_expect_tx(full + s, kTileABytes + kTileBBytes);
                    if (p.taps == 1) {
                        tma_load_2d(sA + s * kTileABytes, &map_x, full + s, kb * kBK, m0);
                    } else {
                        const int r = tap / p.S, sx = tap - r * p.S;
                        tma_load_im2col_4d(sA + s * kTileABytes, &map_x, full + s, kb * kBK, w0 - pad, h0 - pad, img,
                                           static_cast<uint16_t>(sx * p.dil), static_cast<uint16_t>(r * p.dil));
                    }
                    // weight columns of this tap start at tap*Cin; when Cin % 64 != 0 the last block of a tap also fetches
                    // the first columns of the next tap, which meet zero-filled A channels (product 0)
                    tma_load_2d(sB + s * kTileBBytes, &map_w, full + s, tap * p.Cin + kb * kBK, n0);
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
                    mbar_wait(full + s, (it / kStages) & 1);
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
    } else {                                          // ---------------- epilogue (warps 2..9)
        const int ew = warp - 2;
        const int q = warp & 3;                       // TMEM lane quarter == 32-row group of the tile
        const int half = ew >> 2;                     // column half of the tile this warp drains
        const int et = threadIdx.x - 64;              // 0..255
        const int eh = (ew & 3) * 32 + lane;          // 0..127 inside the half's group of four warps
        const bool affine = p.scale != nullptr || p.shift != nullptr;
        float *s_scale = s_par, *s_shift = s_par + kBN;
        uint8_t *slice = sStage + ew * 2 * kSliceBytes;
        uint64_t *rbar = res_full + ew * 2;
        constexpr int kHalfN = kBN / 2;
        constexpr int kChunks = kHalfN / kChunk;
        uint32_t tile_i = 0, chunk_i = 0;             // chunk_i: running chunk count of this warp (buffer = chunk_i & 1)
        uint32_t res_phase[2] = {0, 0};
        // residual of the very first chunk
        if (p.has_residual && lane == 0 && blockIdx.x < num_tiles) {
            const int t = blockIdx.x;
            const int tm = t / tiles_n, n0 = (t % tiles_n) * kBN + half * kHalfN;
            if (n0 < p.Cout && tm * kBM + q * 32 < p.M) {
                mbar_expect_tx(rbar, kSliceBytes);
                tma_load_2d(slice, &map_res, rbar, n0, tm * kBM + q * 32);
            }
        }
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tile_i) {
            const int tm = t / tiles_n, n0 = (t % tiles_n) * kBN;
            const int row0 = tm * kBM + q * 32;       // first pixel of this warp's rows
            const uint32_t acc = tile_i & 1, use = tile_i >> 1;
            if (affine) {
                asm volatile("bar.sync 1, 256;" ::: "memory");         // every warp is done with the previous tile's parameters
                for (int e = et; e < kBN; e += 32 * kEpiWarps) {
                    const int c = n0 + e;
                    s_scale[e] = (p.scale && c < p.Cout) ? __ldg(p.scale + c) : 1.0f;
                    s_shift[e] = (p.shift && c < p.Cout) ? __ldg(p.shift + c) : 0.0f;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait(tmem_full + acc, use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int j = 0; j < kChunks; ++j, ++chunk_i) {
                const int cl = half * kHalfN + j * kChunk, c0 = n0 + cl;
                const uint32_t buf = chunk_i & 1;
                uint8_t *stg = slice + buf * kSliceBytes;
                const bool chunk_live = c0 < p.Cout && row0 < p.M;
                uint32_t r[kCh];
                {
                    const uint32_t taddr = tmem_base + acc * kTmemCols + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(cl);
                    uint32_t(&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
                    tmem_ld_32x32(taddr, r0);
                    if (kCh == 64) {
                        uint32_t(&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[kCh - 32]);
                        tmem_ld_32x32(taddr + 32, r1);
                    }
                }
                if (j == kChunks - 1) {               // last read of this accumulator stage by this warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty + acc);
                }
                float v[kCh];
#pragma unroll
                for (int e = 0; e < kCh; ++e) v[e] = __uint_as_float(r[e]);
                if (affine) {                         // parameters as 16-byte shared-memory broadcasts
#pragma unroll
                    for (int e = 0; e < kCh; e += 4) {
                        const float4 sc = *reinterpret_cast<const float4 *>(s_scale + cl + e);
                        const float4 sf = *reinterpret_cast<const float4 *>(s_shift + cl + e);
                        v[e] = fmaf(v[e], sc.x, sf.x); v[e + 1] = fmaf(v[e + 1], sc.y, sf.y);
                        v[e + 2] = fmaf(v[e + 2], sc.z, sf.z); v[e + 3] = fmaf(v[e + 3], sc.w, sf.w);
                    }
                }
                // The OTHER buffer is about to be reused (residual prefetch of the next chunk now, its rows next iteration):
                // the store that read it (previous chunk) must have finished reading shared memory.
                // (without a residual nothing touches the other buffer now: only the store of two chunks ago must be done)
                if (lane == 0) {
                    if (p.has_residual) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                }
                __syncwarp();
                if (p.has_residual && lane == 0) {    // prefetch the residual rows of the NEXT chunk (possibly of the next tile)
                    int nt = t, nj = j + 1;
                    if (nj == kChunks) { nj = 0; nt = t + gridDim.x; }
                    if (nt < num_tiles) {
                        const int ntm = nt / tiles_n, nn0 = (nt % tiles_n) * kBN + half * kHalfN + nj * kChunk;
                        if (nn0 < p.Cout && ntm * kBM + q * 32 < p.M) {
                            mbar_expect_tx(rbar + (buf ^ 1), kSliceBytes);
                            tma_load_2d(slice + (buf ^ 1) * kSliceBytes, &map_res, rbar + (buf ^ 1), nn0, ntm * kBM + q * 32);
                        }
                    }
                }
                if (p.has_residual && chunk_live) {
                    mbar_wait(rbar + buf, res_phase[buf] & 1);
                    ++res_phase[buf];
                }
                uint8_t *row = stg + lane * kRowB;
                const int swz = (kCh == 64) ? (lane & 7) : ((lane >> 1) & 3);      // SWIZZLE_128B / SWIZZLE_64B: 16-byte unit ^= row bits
#pragma unroll
                for (int g = 0; g < kGroups; ++g) {   // 8 channels = one 16-byte unit at its swizzled position
                    uint4 *cp = reinterpret_cast<uint4 *>(row + ((g ^ swz) << 4));
                    float w8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) w8[e] = v[g * 8 + e];
                    if (p.has_residual && chunk_live) {
                        const uint4 rr = *cp;
                        const __nv_bfloat162 *rh = reinterpret_cast<const __nv_bfloat162 *>(&rr);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __bfloat1622float2(rh[e]);
                            w8[2 * e] += f.x;
                            w8[2 * e + 1] += f.y;
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) w8[e] = fmaxf(w8[e], 0.0f);
                    }
                    uint4 o;
                    __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
                    for (int e = 0; e < 4; ++e) oh[e] = __floats2bfloat162_rn(w8[2 * e], w8[2 * e + 1]);
                    *cp = o;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (chunk_live) tma_store_2d(&map_out, stg, c0, row0);      // TMA clips rows >= M and channels >= Cout
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (kStats) {
                    // Column sums over this warp's 32 staged (bf16-rounded) rows: lane owns columns lane (and lane + 32).  Rows
                    // beyond M and channels beyond Cout hold exact zeros (zero-filled operands, no affine in this variant).
                    constexpr int kCols = kCh / 32;
                    float cs[kCols], cq[kCols];
#pragma unroll
                    for (int h2 = 0; h2 < kCols; ++h2) {
                        const int c = lane + 32 * h2;
                        const uint8_t *colp = stg + (c & 7) * 2;
                        cs[h2] = 0.0f; cq[h2] = 0.0f;
#pragma unroll 8
                        for (int rr = 0; rr < 32; ++rr) {
                            const int sw = (kCh == 64) ? (rr & 7) : ((rr >> 1) & 3);
                            const float xv = __bfloat162float(*reinterpret_cast<const __nv_bfloat16 *>(colp + rr * kRowB + (((c >> 3) ^ sw) << 4)));
                            cs[h2] += xv;
                            cq[h2] = fmaf(xv, xv, cq[h2]);
                        }
                    }
                    float *redb = s_red + half * (4 * kChunk * 2);
#pragma unroll
                    for (int h2 = 0; h2 < kCols; ++h2) {
                        redb[((ew & 3) * kChunk + lane + 32 * h2) * 2] = cs[h2];
                        redb[((ew & 3) * kChunk + lane + 32 * h2) * 2 + 1] = cq[h2];
                    }
                    if (half) asm volatile("bar.sync 3, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
                    if (eh < kChunk && c0 + eh < p.Cout) {
                        float a = 0.0f, b = 0.0f;
#pragma unroll
                        for (int k2 = 0; k2 < 4; ++k2) { a += redb[(k2 * kChunk + eh) * 2]; b += redb[(k2 * kChunk + eh) * 2 + 1]; }
                        float *dst = p.stat_part + static_cast<size_t>(tm) * 2 * p.Cout;
                        dst[c0 + eh] = a;
                        dst[p.Cout + c0 + eh] = b;
                    }
                    if (half) asm volatile("bar.sync 3, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // stores complete before the CTA exits
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * kTmemCols) : "memory");
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const int *, const int *, cuuint32_t, cuuint32_t, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeIm2colFn encode_im2col_fn()
{
    static EncodeIm2colFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeIm2colFn>(p);
    }
    return fn;
}

// x as {C, W, H, N}: bounding box = the image moved by -pad in W and H (as many anchors as output pixels), 64 channels per
// pixel, 128 pixels per column, SWIZZLE_128B, zero fill outside
static bool make_map_im2col(CUtensorMap *map, const void *base, int64_t n, int64_t h, int64_t w, int64_t c, int pad)
{
    EncodeIm2colFn fn = encode_im2col_fn();
    if (!fn) return false;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(n)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(c) * 2, static_cast<cuuint64_t>(w) * c * 2, static_cast<cuuint64_t>(h) * w * c * 2};
    const int lower[2] = {-pad, -pad}, upper[2] = {-pad, -pad};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, lower, upper, kBK, kBM, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    // Drivers up to 13.1 set a descriptor bit for tensors smaller than 128 KB that makes the im2col load fault (the
    // work-around CUTLASS applies in make_im2col_tma_copy_desc); only unit-test sized tensors are that small.
    int drv = 0;
    if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && n * h * w * c * 2 < 131072)
        reinterpret_cast<uint64_t *>(map)[1] &= ~(1ull << 21);
    return true;
}

// row-major [rows, cols] bf16 matrix, box {box_cols, box_rows}; swizzle span = the box's row bytes (128 or 64)
static bool make_map_2d(CUtensorMap *map, const void *base, int64_t rows, int64_t cols, int box_rows, int box_cols)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int kBN, int kStages, int kCh, bool kStats>
static int run(const CUtensorMap &mx, const CUtensorMap &mw, const CUtensorMap &mo, const CUtensorMap &mr, const Params &p,
               unsigned grid, cudaStream_t st)
{
    using L = Smem<kBN, kStages, kCh>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_flat_kernel<kBN, kStages, kCh, kStats>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
        configured = true;
    }
    conv_flat_kernel<kBN, kStages, kCh, kStats><<<grid, kThreads, L::total, st>>>(mx, mw, mo, mr, p);
    return 0;
}

}  // namespace convtc3

int64_t conv_tc3_stat_parts(int64_t n, int64_t h, int64_t w) { return (n * h * w + convtc3::kBM - 1) / convtc3::kBM; }

int conv_tc3_launch(const void *x, const void *wgt, void *out, int64_t n, int64_t h, int64_t w, int64_t cin, int64_t cout,
                    int ksize, int dilation, const float *scale, const float *shift, const void *residual, int relu,
                    float *stat_part, const char *what, void *stream)
{
    using namespace convtc3;
    const int64_t M = n * h * w;
    if (ksize == 3 && (dilation > 127 || 2 * dilation > 255))
        return bad_arg("conv_bf16_nhwc: dilation too large for the im2col tensor map (corner offsets are 8-bit)");
    Params p;
    CUtensorMap mx, mw, mo, mr;
    const int bn = cout > 128 ? 256 : 128;
    // 256-wide tile, two configurations (B200, profiles/r02_conv_flat_*): the compute-bound 3x3 layers want the fourth
    // pipeline stage (64-byte staged rows leave room for it); the memory-bound 1x1 layers want 128-byte staged rows
    // (whole lines per TMA store) and do with three stages.  U2PL_CONV_STAGES=3|4 forces one of them.
    static const int forced_stages = [] { const char *e = getenv("U2PL_CONV_STAGES"); return e ? atoi(e) : 0; }();
    const int stages256 = (forced_stages == 3 || forced_stages == 4) ? forced_stages : (ksize == 1 ? 3 : 4);
    const int ch = (bn == 256 && stages256 == 4) ? 32 : 64;
    bool ok = (ksize == 1) ? make_map_2d(&mx, x, M, cin, kBM, kBK) : make_map_im2col(&mx, x, n, h, w, cin, dilation);
    ok = ok && make_map_2d(&mw, wgt, cout, static_cast<int64_t>(ksize) * ksize * cin, bn, kBK);
    ok = ok && make_map_2d(&mo, out, M, cout, 32, ch) && make_map_2d(&mr, residual ? residual : out, M, cout, 32, ch);
    if (!ok) { set_error("conv_bf16_nhwc (flat kernel): tensor map encoding failed"); return U2PL_E_BADARG; }
    p.M = static_cast<int>(M); p.H = static_cast<int>(h); p.W = static_cast<int>(w);
    p.Cin = static_cast<int>(cin); p.Cout = static_cast<int>(cout);
    p.taps = ksize * ksize; p.S = ksize; p.dil = dilation;
    p.scale = scale; p.shift = shift; p.has_residual = residual != nullptr; p.relu = relu;
    p.stat_part = stat_part;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = ((M + kBM - 1) / kBM) * ((cout + bn - 1) / bn);
    const unsigned grid = static_cast<unsigned>(tiles < sms ? tiles : sms);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if (bn == 256) {
        if (stat_part) rc = stages256 == 4 ? run<256, 4, 32, true>(mx, mw, mo, mr, p, grid, st) : run<256, 3, 64, true>(mx, mw, mo, mr, p, grid, st);
        else rc = stages256 == 4 ? run<256, 4, 32, false>(mx, mw, mo, mr, p, grid, st) : run<256, 3, 64, false>(mx, mw, mo, mr, p, grid, st);
    } else {
        rc = stat_part ? run<128, 4, 64, true>(mx, mw, mo, mr, p, grid, st) : run<128, 4, 64, false>(mx, mw, mo, mr, p, grid, st);
    }
    if (rc != 0) return rc;
    return check_launch(what);
}

}  // namespace u2pl
