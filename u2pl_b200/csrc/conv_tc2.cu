// conv_tc2.cu -- the CTA-PAIR (tcgen05 cta_group::2) form of the implicit-GEMM convolution of conv_tc.cu.
//
// Same contract as conv_tc.cu (stride-1 "same" 1x1 / 3x3 dilated convolution of a channels-last bf16 tensor, epilogue
// act(conv * scale[co] + shift[co] + residual), optional per-channel statistics of the stored values), different
// machine mapping -- the one cuDNN's own sm_100 kernels use (256x256x64 2-SM tiles):
//
//   * a cluster of two CTAs (one TPC) computes a 256-pixel x 256-channel tile.  CTA r loads ITS 128 pixels of the A
//     operand (one 4-D TMA box per tap / channel block, zero fill = padding) and HALF of the weight tile (128 of the
//     256 output channels); one thread of the leader CTA issues tcgen05.mma.cta_group::2 (M 256, N 256, K 16), which
//     reads the B halves out of both CTAs' shared memory.  Operand fill drops from 94 B/clk/SM (1-CTA 128x256 tile) to
//     64 B/clk/SM -- the L2 -> SM path (~42 B/clk/SM sustained chip-wide) is what held conv_tc.cu at 1.0-1.2 PFLOP/s.
//   * the epilogue leaves through shared memory and TMA: eight warps (two per TMEM lane quarter, one per 128-column
//     half) read the accumulators with tcgen05.ld, apply scale/shift (+ the residual tile, which TMA brought INTO the
//     staging buffer) + ReLU, write bf16 rows into a SWIZZLE_128B staging tile and one thread issues a
//     cp.async.bulk.tensor store of the {64 ch, tw, th, 1} box.  Stores are full 128-byte lines (the 1-CTA kernel's
//     per-thread 16-byte stores ran the 256 -> 1024 1x1 layers at 0.4 PFLOP/s), and TMA clips the part of the box that
//     lies outside the image, so partial tiles need no predicates.
//   * statistics (train mode): column sums / sums of squares are taken from the staged bf16 tile (the values as
//     stored), fixed order, one partial per (pixel tile, half, channel).
//
// Pair protocol (validated in isolation by tools/cu/umma_2cta_probe.cu on a B200): both CTAs' TMA loads complete on
// the LEADER's `full` barrier (peer producer arrives remotely), tcgen05.commit.multicast frees the stage in both CTAs
// and publishes the accumulator to both epilogues; both epilogues release the accumulator stage on the leader's
// `tmem_empty` barrier.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "conv_tc2.cuh"

namespace u2pl {
namespace convtc2 {

constexpr int kBM = 128;                 // pixels per CTA (256 per pair)
constexpr int kBN = 256;                 // output channels per pair tile
constexpr int kBNHalf = 128;             // weight rows each CTA loads
constexpr int kBK = 64;
constexpr int kStages = 4;
constexpr int kTileBytes = kBM * kBK * 2;            // 16 KB: A tile == B half tile
constexpr int kStageBytes = 2 * kTileBytes;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;        // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kChunk = 64;                           // channels per staged store
constexpr int kStoreBytes = kBM * kChunk * 2;        // 16 KB staging tile
constexpr int kTmemCols = 512;                       // two 256-column accumulator stages

struct Smem {                                        // carved out of dynamic shared memory (1024-aligned base)
    static constexpr int a = 0;
    static constexpr int b = kStages * kTileBytes;
    static constexpr int store = 2 * kStages * kTileBytes;             // [2 halves][2 buffers][16 KB]
    static constexpr int par = store + 4 * kStoreBytes;                // [2 acc stages][scale 256 | shift 256] floats
    static constexpr int bars = par + 2 * 2 * kBN * 4;
    static constexpr int total = bars + 256;
};

__device__ __forceinline__ uint32_t cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t bar_cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void *dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar)
{
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Debug timeline (U2PL_CONV2_TRACE=1): clock64 of cluster 0's producer / MMA threads at the first 128 k-blocks.
__device__ long long g_trace[6][128];

template <bool kStats, bool kTrace = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res, Conv2Params p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *sA = smem + Smem::a, *sB = smem + Smem::b, *sStore = smem + Smem::store;
    float *s_par = reinterpret_cast<float *>(smem + Smem::par);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + Smem::bars);      // [kStages]  (used in the leader only)
    uint64_t *empty = full + kStages;                                      // [kStages]  one copy per CTA
    uint64_t *tmem_full = empty + kStages;                                 // [2]        one copy per CTA
    uint64_t *tmem_empty = tmem_full + 2;                                  // [2]        (leader only)
    uint64_t *res_full = tmem_empty + 2;                                   // [2 halves][2 buffers] residual tile landed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(res_full + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int tw = 1 << p.log2_tw, th = kBM >> p.log2_tw;
    const int tiles_img = p.tiles_h * p.tiles_w;
    const int tiles_m = p.Nimg * tiles_img;                    // 128-pixel tiles
    const int pairs_m = (tiles_m + 1) >> 1;
    const int tiles_n = (p.Cout + kBN - 1) / kBN;
    const int num_tiles = pairs_m * tiles_n;
    const int kb_per_tap = (p.Cin + kBK - 1) / kBK;
    const int nkb = p.R * p.S * kb_per_tap;
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
        if (p.has_residual) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_res) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 2); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, 2 * kEpiWarps); }
        for (int i = 0; i < 4; ++i) mbar_init(res_full + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                              // ---------------- TMA producer (both CTAs)
            uint32_t it = 0;
            for (int t = pair_id; t < num_tiles; t += num_pairs) {
                const int tp = t / tiles_n, n0 = (t % tiles_n) * kBN;
                const int tm = 2 * tp + static_cast<int>(rank);        // tm >= tiles_m (odd tile count): img >= Nimg -> zero fill
                const int img = tm / tiles_img, rem = tm % tiles_img;
                const int h0 = (rem / p.tiles_w) * th, w0 = (rem % p.tiles_w) * tw;
                for (int step = 0; step < nkb; ++step, ++it) {
                    const int tap = step / kb_per_tap, kb = step % kb_per_tap;
                    const int dh = (tap / p.S - p.R / 2) * p.dil, dw = (tap % p.S - p.S / 2) * p.dil;
                    const int s = it % kStages;
                    if (kTrace && pair_id == 0 && it < 128) g_trace[rank * 2][it] = clock64();
                    mbar_wait(empty + s, ((it / kStages) & 1) ^ 1);
                    if (kTrace && pair_id == 0 && it < 128) g_trace[rank * 2 + 1][it] = clock64();
                    const uint32_t leader_full = map_to_cta(smem_u32(full + s), 0);
                    if (rank == 0) mbar_expect_tx(full + s, 2 * kStageBytes);         // both CTAs' A tile and B half
                    else remote_arrive(leader_full);
                    tma_load_4d_pair(sA + s * kTileBytes, &map_x, leader_full, kb * kBK, w0 + dw, h0 + dh, img);
                    tma_load_2d_pair(sB + s * kTileBytes, &map_w, leader_full, tap * p.Cin + kb * kBK, n0 + static_cast<int>(rank) * kBNHalf);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {                 // ---------------- MMA issuer (leader only)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kBN >> 3) << 17) |
                                   (static_cast<uint32_t>((2 * kBM) >> 4) << 24);
            uint32_t it = 0, tile_i = 0;
            for (int t = pair_id; t < num_tiles; t += num_pairs, ++tile_i) {
                const uint32_t acc = tile_i & 1, use = tile_i >> 1;
                mbar_wait(tmem_empty + acc, (use & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * kBN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % kStages;
                    if (kTrace && pair_id == 0 && it < 128) g_trace[4][it] = clock64();
                    mbar_wait(full + s, (it / kStages) & 1);
                    if (kTrace && pair_id == 0 && it < 128) g_trace[5][it] = clock64();
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = smem_u32(sA + s * kTileBytes), b0 = smem_u32(sB + s * kTileBytes);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                        umma_f16_pair(d_tmem, smem_desc_sw128(a0 + 32 * k), smem_desc_sw128(b0 + 32 * k), idesc, (kb | k) ? 1u : 0u);
                    umma_commit_pair(empty + s);      // frees stage s in both CTAs
                }
                umma_commit_pair(tmem_full + acc);    // accumulators complete, both CTAs
            }
        }
    } else {                                          // ---------------- epilogue (warps 2..9, both CTAs)
        const int ew = warp - 2;
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int half = ew >> 2;                     // column half (128 channels) this warp drains
        const int et = (ew & 3) * 32 + lane;          // 0..127 inside the half's group of four warps
        const int m = q * 32 + lane;                  // accumulator row == pixel of the tile
        const int ti = m >> p.log2_tw, tj = m & (tw - 1);
        const bool affine = p.scale != nullptr || p.shift != nullptr;
        uint8_t *stage_base = sStore + half * 2 * kStoreBytes;
        const uint32_t leader_tmem_empty0 = map_to_cta(smem_u32(tmem_empty), 0);
        uint32_t tile_i = 0, chunk_i = 0, res_use0 = 0, res_use1 = 0;   // residual barrier phases, per staging buffer
        for (int t = pair_id; t < num_tiles; t += num_pairs, ++tile_i) {
            const int tp = t / tiles_n, n0 = (t % tiles_n) * kBN;
            const int tm = 2 * tp + static_cast<int>(rank);
            const int img = tm / tiles_img, rem = tm % tiles_img;
            const int h0 = (rem / p.tiles_w) * th, w0 = (rem % p.tiles_w) * tw;
            const bool live = tm < tiles_m && (h0 + ti) < p.H && (w0 + tj) < p.W;
            const uint32_t acc = tile_i & 1, use = tile_i >> 1;
            float *s_scale = s_par + acc * 2 * kBN, *s_shift = s_scale + kBN;
            if (affine) {
                const int e = threadIdx.x - 64;       // 0..255
                const int c = n0 + e;
                s_scale[e] = (p.scale && c < p.Cout) ? __ldg(p.scale + c) : 1.0f;
                s_shift[e] = (p.shift && c < p.Cout) ? __ldg(p.shift + c) : 0.0f;
                named_bar(1, 32 * kEpiWarps);
            }
            mbar_wait(tmem_full + acc, use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int j = 0; j < 2; ++j, ++chunk_i) {
                const int cl = half * kBNHalf + j * kChunk;       // first column of this chunk inside the tile
                const int c0 = n0 + cl;
                const uint32_t buf = chunk_i & 1;
                uint8_t *stg = stage_base + buf * kStoreBytes;
                uint64_t *rbar = res_full + half * 2 + buf;
                // the TMA store that last read this buffer (two chunks ago) must have finished reading shared memory
                if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                named_bar(2 + half, 128);
                const bool chunk_live = c0 < p.Cout && tm < tiles_m;
                if (p.has_residual && chunk_live && et == 0) {    // residual tile -> staging buffer (same box as the store)
                    mbar_expect_tx(rbar, kStoreBytes);
                    tma_load_4d(stg, &map_res, rbar, c0, w0, h0, img);
                }
                uint32_t r[64];
                {
                    const uint32_t taddr = tmem_base + acc * kBN + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(cl);
                    uint32_t(&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
                    uint32_t(&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
                    tmem_ld_32x32(taddr, r0);
                    tmem_ld_32x32(taddr + 32, r1);
                }
                if (j == 1) {                         // last read of this accumulator stage by this warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) remote_arrive(leader_tmem_empty0 + acc * 8);
                }
                if (p.has_residual && chunk_live) {
                    if (buf) { mbar_wait(rbar, res_use1 & 1); ++res_use1; }
                    else { mbar_wait(rbar, res_use0 & 1); ++res_use0; }
                }
                uint8_t *row = stg + m * 128;
#pragma unroll
                for (int g = 0; g < 8; ++g) {         // 8 channels = one 16-byte chunk, SWIZZLE_128B position
                    uint4 *cp = reinterpret_cast<uint4 *>(row + ((g ^ (m & 7)) << 4));
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        v[e] = __uint_as_float(r[g * 8 + e]);
                        if (affine) v[e] = fmaf(v[e], s_scale[cl + g * 8 + e], s_shift[cl + g * 8 + e]);
                    }
                    if (p.has_residual && chunk_live) {
                        const uint4 rr = *cp;
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
                    if (kStats && !(live && c0 + g * 8 < p.Cout)) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = 0.0f;       // dead pixels / channels contribute nothing
                    }
                    uint4 o;
                    __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
                    for (int e = 0; e < 4; ++e) oh[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
                    *cp = o;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                named_bar(2 + half, 128);
                if (et == 0) {
                    if (chunk_live) tma_store_4d(&map_out, stg, c0, w0, h0, img);      // TMA clips rows/channels outside the tensor
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (kStats) {                         // column sums over the staged (bf16-rounded) tile: thread = (channel, row half)
                    const int c = et & 63, rh2 = et >> 6;
                    if (chunk_live && c0 + c < p.Cout) {
                        float cs = 0.0f, cq = 0.0f;
                        const uint8_t *colp = stg + (c & 7) * 2;
#pragma unroll 8
                        for (int rr = rh2 * 64; rr < rh2 * 64 + 64; ++rr) {
                            const float xv = __bfloat162float(*reinterpret_cast<const __nv_bfloat16 *>(colp + rr * 128 + (((c >> 3) ^ (rr & 7)) << 4)));
                            cs += xv;
                            cq = fmaf(xv, xv, cq);
                        }
                        float *dst = p.stat_part + (static_cast<size_t>(tm) * 2 + rh2) * 2 * p.Cout;
                        dst[c0 + c] = cs;
                        dst[p.Cout + c0 + c] = cq;
                    }
                }
            }
        }
        if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // stores complete before the CTA exits
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                               // nobody frees TMEM / exits while the peer still uses it
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
}

// x / out / residual viewed as {C, W, H, N} (innermost first), bf16, dense NHWC; box {64, tw, th, 1}; 128-byte swizzle
static bool make_map_nhwc(CUtensorMap *map, const void *base, int64_t n, int64_t h, int64_t w, int64_t c, int th, int tw)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(n)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(c) * 2, static_cast<cuuint64_t>(w) * c * 2, static_cast<cuuint64_t>(h) * w * c * 2};
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(tw), static_cast<cuuint32_t>(th), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool make_map_weight(CUtensorMap *map, const void *base, int64_t cout, int64_t ktot)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(cout)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ktot) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(kBNHalf)};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace convtc2

bool conv_tc2_eligible(int64_t cout, bool xform)
{
    // Opt-in (U2PL_CONV_V=2).  Measured on B200 (profiles/r02_conv_selftest_perf_*.txt): the pair kernel is correct but
    // runs at ~0.55 PFLOP/s where the 1-CTA kernel reaches 1.05-1.2 PFLOP/s on the same layers, so the 1-CTA kernel is
    // the default until the pair pipeline's stall is found.
    static const int forced = [] { const char *e = getenv("U2PL_CONV_V"); return e ? atoi(e) : 0; }();
    if (forced != 2) return false;
    return !xform && cout > 128;
}

int64_t conv_tc2_stat_parts(int64_t n, int64_t h, int64_t w, int ksize)
{
    const int64_t tiles = (ksize == 1) ? (n * h * w + convtc2::kBM - 1) / convtc2::kBM : n * ((h + 7) / 8) * ((w + 15) / 16);
    return 2 * tiles;
}

int conv_tc2_launch(const void *x, const void *wgt, void *out, int64_t n, int64_t h, int64_t w, int64_t cin, int64_t cout,
                    int ksize, int dilation, const float *scale, const float *shift, const void *residual, int relu,
                    float *stat_part, const char *what, void *stream)
{
    using namespace convtc2;
    Conv2Params p;
    CUtensorMap mx, mw, mo, mr;
    bool ok;
    const void *res_base = residual ? residual : out;
    if (ksize == 1) {                                 // flat: every 128 consecutive pixels are one tile
        p.Nimg = 1; p.H = 1; p.W = static_cast<int>(n * h * w);
        p.log2_tw = 7; p.tiles_h = 1; p.tiles_w = (p.W + kBM - 1) / kBM;
        ok = make_map_nhwc(&mx, x, 1, 1, n * h * w, cin, 1, kBM) && make_map_nhwc(&mo, out, 1, 1, n * h * w, cout, 1, kBM) &&
             make_map_nhwc(&mr, res_base, 1, 1, n * h * w, cout, 1, kBM);
    } else {                                          // 8 x 16 pixel tiles inside each image
        p.Nimg = static_cast<int>(n); p.H = static_cast<int>(h); p.W = static_cast<int>(w);
        p.log2_tw = 4; p.tiles_h = (p.H + 7) / 8; p.tiles_w = (p.W + 15) / 16;
        ok = make_map_nhwc(&mx, x, n, h, w, cin, 8, 16) && make_map_nhwc(&mo, out, n, h, w, cout, 8, 16) &&
             make_map_nhwc(&mr, res_base, n, h, w, cout, 8, 16);
    }
    ok = ok && make_map_weight(&mw, wgt, cout, static_cast<int64_t>(ksize) * ksize * cin);
    if (!ok) { set_error("conv_bf16_nhwc (pair kernel): cuTensorMapEncodeTiled failed"); return U2PL_E_BADARG; }
    p.Cin = static_cast<int>(cin); p.Cout = static_cast<int>(cout);
    p.R = p.S = ksize; p.dil = dilation;
    p.scale = scale; p.shift = shift; p.has_residual = residual != nullptr; p.relu = relu;
    p.stat_part = stat_part;
    const size_t smem = Smem::total + 1024;
    static bool configured = false;
    static int max_clusters = 0;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * kNumSMs); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, conv_tc2_kernel<false>, &cfg) != cudaSuccess || max_clusters < 1) {
            cudaGetLastError();
            int dev = 0, sms = kNumSMs;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            max_clusters = sms / 2;
        }
        if (const char *e = getenv("U2PL_CONV2_CLUSTERS")) {           // A/B override of the persistent grid (clusters)
            const int v = atoi(e);
            if (v > 0) max_clusters = v;
        }
        if (getenv("U2PL_CONV_DEBUG")) fprintf(stderr, "[conv_tc2] clusters resident: %d (smem %zu B/CTA)\n", max_clusters, smem);
        configured = true;
    }
    const long long tiles_m = static_cast<long long>(p.Nimg) * p.tiles_h * p.tiles_w;
    const long long tiles = ((tiles_m + 1) / 2) * ((cout + kBN - 1) / kBN);
    const unsigned pairs = static_cast<unsigned>(tiles < max_clusters ? tiles : max_clusters);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static const bool trace = getenv("U2PL_CONV2_TRACE") != nullptr;
    if (trace && !stat_part) {                        // debug: timeline of cluster 0, printed after a device synchronise
        static bool tconf = false;
        if (!tconf) { cudaFuncSetAttribute(conv_tc2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); tconf = true; }
        conv_tc2_kernel<false, true><<<2 * pairs, kThreads, smem, st>>>(mx, mw, mo, mr, p);
        cudaDeviceSynchronize();
        static long long h[6][128];
        cudaMemcpyFromSymbol(h, g_trace, sizeof(h));
        static int dumps = 0;
        if (dumps++ < 2) {
            fprintf(stderr, "[conv_tc2 trace] cin %lld cout %lld k %d: k-block | leader producer wait-start, wait-end | MMA full wait-start, wait-end | peer producer wait-start, wait-end (leader clocks relative to leader t0; peer relative to peer t0)\n",
                    static_cast<long long>(cin), static_cast<long long>(cout), ksize);
            for (int i = 0; i < 40; ++i)
                fprintf(stderr, "  %3d | %7lld %7lld | %7lld %7lld | %7lld %7lld\n", i, h[0][i] - h[0][0], h[1][i] - h[0][0], h[4][i] - h[0][0],
                        h[5][i] - h[0][0], h[2][i] - h[2][0], h[3][i] - h[2][0]);
        }
        return check_launch(what);
    }
    if (stat_part) conv_tc2_kernel<true><<<2 * pairs, kThreads, smem, st>>>(mx, mw, mo, mr, p);
    else conv_tc2_kernel<false><<<2 * pairs, kThreads, smem, st>>>(mx, mw, mo, mr, p);
    return check_launch(what);
}

}  // namespace u2pl
