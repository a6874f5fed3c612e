// gemm_tc.cu -- bf16 GEMM on the 5th-generation tensor cores (tcgen05) fed by TMA, with a fused
// per-column scale/shift(+ReLU) epilogue.   D[M,N] = act( (A[M,K] . B[N,K]^T) * scale[n] + shift[n] )
//
// This is the contraction behind every 1x1 convolution of the network in channels-last layout
// (resnet.py:39-41: conv1x1; A = activations [N*H*W, Cin], B = weight [Cout, Cin]); with
// scale/shift = folded eval-mode BatchNorm it is conv + BN + ReLU in one pass (teacher pseudo-label
// forward, train_semi.py:318-319).
//
// Structure (one CTA per 128x128 output tile, 192 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor.2d of a 128x64 A tile and a 128x64 B tile per stage
//              (SWIZZLE_128B), completion on the stage's `full` mbarrier
//   warp 1     allocates 128 TMEM columns, then one elected lane issues 4 x tcgen05.mma
//              (M=128, N=128, K=16, kind::f16, fp32 accumulate in TMEM) per stage and
//              tcgen05.commit's the stage's `empty` barrier; after the last K block commits `tmem_full`
//   warps 2-5  epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) -> scale/shift/ReLU ->
//              bf16 -> 16-byte global stores (each lane owns one output row)
// 4-stage shared-memory ring (4 x 32 KB), all synchronisation through mbarriers.
// OOB rows/columns of partial tiles are zero-filled by TMA and masked in the epilogue.
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace u2pl {

constexpr int kBM = 128, kBN = 128, kBK = 64, kStages = 4;
constexpr int kTileABytes = kBM * kBK * 2, kTileBBytes = kBN * kBK * 2;
constexpr int kGemmThreads = 192;
constexpr int kTmemCols = 128;

struct GemmParams {
    int M, N, K;
    const float *scale, *shift;       // may be null
    int relu;
    __nv_bfloat16 *D;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *sA = smem, *sB = smem + kStages * kTileABytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + kStages * kTileBBytes);
    uint64_t *empty = full + kStages;
    uint64_t *tmem_full = empty + kStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);
    float *s_scale = reinterpret_cast<float *>(tmem_slot + 2), *s_shift = s_scale + kBN;   // epilogue parameters of this tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
    const int nkb = (p.K + kBK - 1) / kBK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                  // whole warp: TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                              // ---------------- TMA producer
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(empty + s, ph ^ 1);
                mbar_expect_tx(full + s, kTileABytes + kTileBBytes);
                tma_load_2d(sA + s * kTileABytes, &map_a, full + s, kb * kBK, m0);
                tma_load_2d(sB + s * kTileBBytes, &map_b, full + s, kb * kBK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                              // ---------------- MMA issuer
            // instruction descriptor: D fp32 (bit 4), A bf16 (bit 7), B bf16 (bit 10), both K-major, N>>3 at 17, M>>4 at 24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kBN >> 3) << 17) |
                                   (static_cast<uint32_t>(kBM >> 4) << 24);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(full + s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_u32(sA + s * kTileABytes), b0 = smem_u32(sB + s * kTileBBytes);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k)    // 32 bytes of K per UMMA_K = 16 bf16
                    umma_f16(tmem_base, smem_desc_sw128(a0 + 32 * k), smem_desc_sw128(b0 + 32 * k), idesc, (kb | k) ? 1u : 0u);
                umma_commit(empty + s);               // frees the smem stage once these MMAs have read it
            }
            umma_commit(tmem_full);                   // accumulator complete
        }
    } else {                                          // ---------------- epilogue (warps 2..5)
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        {                                             // stage the tile's 128 scale / shift values once (smem broadcast reads later)
            const int e = threadIdx.x - 64, c = n0 + e;
            s_scale[e] = (p.scale && c < p.N) ? __ldg(p.scale + c) : 1.0f;
            s_shift[e] = (p.shift && c < p.N) ? __ldg(p.shift + c) : 0.0f;
            asm volatile("bar.sync 1, 128;" ::: "memory");       // the four epilogue warps only
        }
        const bool affine = p.scale != nullptr || p.shift != nullptr;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + q * 32 + lane;
#pragma unroll 1
        for (int j = 0; j < kBN / 32; ++j) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(j * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int c0 = n0 + j * 32;
            if (row < p.M) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {         // 8 columns = one 16-byte store
                    const int c = c0 + g * 8;
                    if (c < p.N) {                    // N % 8 == 0: a granule is either fully inside or outside
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float x = __uint_as_float(r[g * 8 + e]);
                            if (affine) x = fmaf(x, s_scale[j * 32 + g * 8 + e], s_shift[j * 32 + g * 8 + e]);
                            if (p.relu) x = fmaxf(x, 0.0f);
                            v[e] = x;
                        }
                        uint4 o;
                        __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
                        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
                        *reinterpret_cast<uint4 *>(p.D + static_cast<size_t>(row) * p.N + c) = o;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
}


// ------------------------------------------------------------------ persistent variant
// One CTA per SM loops over output tiles (n fastest, so CTAs of a wave share A rows in L2).  TMEM holds two
// 128-column accumulator stages: the epilogue warps drain tile i while the MMA warp is already accumulating
// tile i+1, and the TMA ring never drains between tiles.  Extra barriers: tmem_full[2] (MMA -> epilogue,
// tcgen05.commit) and tmem_empty[2] (epilogue -> MMA, one arrival per epilogue warp).
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *sA = smem, *sB = smem + kStages * kTileABytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + kStages * kTileBBytes);
    uint64_t *empty = full + kStages;
    uint64_t *tmem_full = empty + kStages;            // [2]
    uint64_t *tmem_empty = tmem_full + 2;             // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *s_par = reinterpret_cast<float *>(tmem_slot + 2);      // [2 stages][scale 128 | shift 128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (p.N + kBN - 1) / kBN, tiles_m = (p.M + kBM - 1) / kBM;
    const int num_tiles = tiles_m * tiles_n;
    const int nkb = (p.K + kBK - 1) / kBK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, 4); }
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
            uint32_t it = 0;                          // k-blocks issued so far (ring position / phase)
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m0 = (t / tiles_n) * kBM, n0 = (t % tiles_n) * kBN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(empty + s, ((it / kStages) & 1) ^ 1);
                    mbar_expect_tx(full + s, kTileABytes + kTileBBytes);
                    tma_load_2d(sA + s * kTileABytes, &map_a, full + s, kb * kBK, m0);
                    tma_load_2d(sB + s * kTileBBytes, &map_b, full + s, kb * kBK, n0);
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
                mbar_wait(tmem_empty + acc, (use & 1) ^ 1);          // epilogue has drained this accumulator stage
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
    } else {                                          // ---------------- epilogue (warps 2..5)
        const int q = warp & 3;
        const bool affine = p.scale != nullptr || p.shift != nullptr;
        uint32_t tile_i = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tile_i) {
            const int m0 = (t / tiles_n) * kBM, n0 = (t % tiles_n) * kBN;
            const uint32_t acc = tile_i & 1, use = tile_i >> 1;
            float *s_scale = s_par + acc * 2 * kBN, *s_shift = s_scale + kBN;
            if (affine) {                             // stage this tile's parameters (buffer `acc`: the barrier of the next
                const int e = threadIdx.x - 64, c = n0 + e;       // tile orders these writes after every read of tile i-2)
                s_scale[e] = (p.scale && c < p.N) ? __ldg(p.scale + c) : 1.0f;
                s_shift[e] = (p.shift && c < p.N) ? __ldg(p.shift + c) : 0.0f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(tmem_full + acc, use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m0 + q * 32 + lane;
#pragma unroll 1
            for (int j = 0; j < kBN / 32; ++j) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + acc * kTmemCols + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(j * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (j == kBN / 32 - 1) {              // all of this warp's TMEM reads are done: hand the stage back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty + acc);
                }
                const int c0 = n0 + j * 32;
                if (row < p.M) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int c = c0 + g * 8;
                        if (c < p.N) {
                            float v[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                float x = __uint_as_float(r[g * 8 + e]);
                                if (affine) x = fmaf(x, s_scale[j * 32 + g * 8 + e], s_shift[j * 32 + g * 8 + e]);
                                if (p.relu) x = fmaxf(x, 0.0f);
                                v[e] = x;
                            }
                            uint4 o;
                            __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
                            for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
                            *reinterpret_cast<uint4 *>(p.D + static_cast<size_t>(row) * p.N + c) = o;
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * kTmemCols) : "memory");
}

// ------------------------------------------------------------------ host side
// row-major [rows, K] bf16 matrix -> tensor map with a (64 x box_rows) box, 128-byte swizzle, zero OOB fill
static bool make_map(CUtensorMap *map, const void *base, int64_t rows, int64_t K, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace u2pl

using namespace u2pl;

extern "C" int u2pl_gemm_bf16_tn(const void *A, const void *B, void *D, int64_t M, int64_t N, int64_t K,
                                 const float *scale, const float *shift, int relu, void *stream)
{
    if (M <= 0 || N <= 0 || K <= 0 || (K % 8) || (N % 8) || M >= (1LL << 31) || N >= (1LL << 31))
        return bad_arg("gemm_bf16_tn: need K % 8 == 0 and N % 8 == 0");
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(D)) & 15)
        return bad_arg("gemm_bf16_tn: operands must be 16-byte aligned");
    CUtensorMap ma, mb;
    if (!make_map(&ma, A, M, K, kBM) || !make_map(&mb, B, N, K, kBN)) { set_error("gemm_bf16_tn: cuTensorMapEncodeTiled failed"); return U2PL_E_BADARG; }
    const size_t smem = static_cast<size_t>(kStages) * (kTileABytes + kTileBBytes) + 1024 + 256 + 4 * kBN * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gemm_bf16_tn_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
        configured = true;
    }
    GemmParams p;
    p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
    p.scale = scale; p.shift = shift; p.relu = relu; p.D = static_cast<__nv_bfloat16 *>(D);
    // persistent kernel by default (faster or equal on every measured shape); U2PL_GEMM_PERSISTENT=0 selects the
    // one-tile-per-CTA kernel above
    static const bool persistent = [] { const char *e = getenv("U2PL_GEMM_PERSISTENT"); return !(e && e[0] == '0'); }();
    if (persistent) {
        const long long tiles = ((N + kBN - 1) / kBN) * ((M + kBM - 1) / kBM);
        int dev = 0, sms = kNumSMs;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const unsigned g = static_cast<unsigned>(tiles < sms ? tiles : sms);
        gemm_bf16_tn_persistent_kernel<<<g, kGemmThreads, smem, static_cast<cudaStream_t>(stream)>>>(ma, mb, p);
        return check_launch("gemm_bf16_tn(persistent)");
    }
    dim3 grid(static_cast<unsigned>((N + kBN - 1) / kBN), static_cast<unsigned>((M + kBM - 1) / kBM));
    gemm_bf16_tn_kernel<<<grid, kGemmThreads, smem, static_cast<cudaStream_t>(stream)>>>(ma, mb, p);
    return check_launch("gemm_bf16_tn");
}
