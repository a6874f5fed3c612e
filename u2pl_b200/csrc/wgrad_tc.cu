// wgrad_tc.cu -- weight gradient of a stride-1 "same" 3x3 (dilated) convolution of channels-last bf16 tensors on the
// tcgen05 tensor cores, WITHOUT transposing or cropping anything:
//
//     dW[co, r, s, ci] = sum_{n,h,w} gout[n, h, w, co] * x[n, h + (r-1)d, w + (s-1)d, ci]
//
// Per tap this is a GEMM D[co, ci] = A^T B with A = gout viewed as [pixels, Cout] and B = (shifted) x viewed as
// [pixels, Cin]: the contraction index (pixel) is the slow dimension of BOTH channels-last operands, i.e. both are
// "MN-major" in UMMA terms.  One pipeline stage holds 64 pixels (a 4 x 16 patch of one image) of
//   A: two  4-D TMA boxes {64 co, 16 w, 4 h, 1 n} of gout                     -> 128 rows of the M dimension
//   B: four 4-D TMA boxes {64 ci, 16 w, 4 h, 1 n} of x at the tap's offset    -> 256 rows of the N dimension
// (TMA zero-fills what falls outside the image = the convolution's zero padding; outside-image pixels of gout are
// zero-filled too, so partial patches contribute nothing).  Each box is 64 K-rows x 128 bytes, SWIZZLE_128B: the
// MN-major canonical layout with LBO = 8192 (next 64 channels), SBO = 1024 (next 8 pixels), 2048 bytes per K = 16 step
// -- the encoding tools/cu/umma_mn_major_probe.cu checks on the GPU (overridable: U2PL_WGRAD_DESC="lbo,sbo,kstep").
// Work item = (tap, 128-co tile, 256-ci tile, K split); fp32 partial tiles [split][tap][Cout][Cin] are summed by the
// caller.  CTA layout / barriers / persistent loop / two 256-column TMEM stages: as conv_tc.cu.
//
// replaces: the weight-gradient half of torch.ops.aten.convolution_backward (cuDNN) for conv3x3 (resnet.py:25-36),
// the ASPP branches (base.py:38-75) and the decoder convs (decoder.py:60-113).  STATUS: matches a CPU loop on a B200 (tools/cu/tc_selftest.cu,
// profiles/r01_tc_selftest.txt); untimed; opt-in (U2PL_TC_WGRAD=1), not on any default path.
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace u2pl {

namespace wgradtc {
constexpr int kBM = 128, kBN = 256, kBK = 64, kStages = 4;        // M = co, N = ci, K = pixels
constexpr int kBoxBytes = 64 * kBK * 2;                            // one {64 channels, 64 pixels} box = 8 KB
constexpr int kTileABytes = (kBM / 64) * kBoxBytes, kTileBBytes = (kBN / 64) * kBoxBytes;
constexpr int kThreads = 192;
constexpr int kTmemCols = kBN;
constexpr int kPatchH = 4, kPatchW = 16;                           // 64 pixels per K block
}  // namespace wgradtc

struct WgradParams {
    int Nimg, H, W, Cin, Cout, dil;
    int tiles_h, tiles_w;             // 4 x 16 pixel patches per image
    int splits;                       // K splits (each >= 1 K block)
    uint32_t lbo, sbo, kstep;         // MN-major descriptor fields (bytes)
    float *D;                         // [splits][9][Cout][Cin]
};

__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return static_cast<uint64_t>((saddr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(wgradtc::kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_x, WgradParams p)
{
    using namespace wgradtc;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *sA = smem, *sB = smem + kStages * kTileABytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + kStages * kTileBBytes);
    uint64_t *empty = full + kStages;
    uint64_t *tmem_full = empty + kStages;            // [2]
    uint64_t *tmem_empty = tmem_full + 2;             // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (p.Cout + kBM - 1) / kBM, tiles_n = (p.Cin + kBN - 1) / kBN;
    const int patches_img = p.tiles_h * p.tiles_w;
    const int kb_total = p.Nimg * patches_img;
    const int num_items = 9 * tiles_m * tiles_n * p.splits;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
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

    // item -> (split, tap, co tile, ci tile); ci fastest so that CTAs of a wave share the gout patches in L2
    auto decode = [&](int item, int &sp, int &tap, int &co0, int &ci0, int &kb_lo, int &kb_hi) {
        const int nt = item % tiles_n; item /= tiles_n;
        const int mt = item % tiles_m; item /= tiles_m;
        tap = item % 9;
        sp = item / 9;
        co0 = mt * kBM; ci0 = nt * kBN;
        kb_lo = static_cast<int>(static_cast<long long>(kb_total) * sp / p.splits);
        kb_hi = static_cast<int>(static_cast<long long>(kb_total) * (sp + 1) / p.splits);
    };

    if (warp == 0) {
        if (lane == 0) {                              // ---------------- TMA producer
            uint32_t it = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                int sp, tap, co0, ci0, kb_lo, kb_hi;
                decode(item, sp, tap, co0, ci0, kb_lo, kb_hi);
                const int dh = (tap / 3 - 1) * p.dil, dw = (tap % 3 - 1) * p.dil;
                for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
                    const int img = kb / patches_img, rem = kb % patches_img;
                    const int h0 = (rem / p.tiles_w) * kPatchH, w0 = (rem % p.tiles_w) * kPatchW;
                    const int s = it % kStages;
                    mbar_wait(empty + s, ((it / kStages) & 1) ^ 1);
                    mbar_expect_tx(full + s, kTileABytes + kTileBBytes);
#pragma unroll
                    for (int b = 0; b < kBM / 64; ++b)
                        tma_load_4d(sA + s * kTileABytes + b * kBoxBytes, &map_g, full + s, co0 + 64 * b, w0, h0, img);
#pragma unroll
                    for (int b = 0; b < kBN / 64; ++b)
                        tma_load_4d(sB + s * kTileBBytes + b * kBoxBytes, &map_x, full + s, ci0 + 64 * b, w0 + dw, h0 + dh, img);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                              // ---------------- MMA issuer
            // D fp32, A/B bf16, BOTH operands MN-major (transpose bits 15 and 16), N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                   (static_cast<uint32_t>(kBN >> 3) << 17) | (static_cast<uint32_t>(kBM >> 4) << 24);
            uint32_t it = 0, tile_i = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++tile_i) {
                int sp, tap, co0, ci0, kb_lo, kb_hi;
                decode(item, sp, tap, co0, ci0, kb_lo, kb_hi);
                const uint32_t acc = tile_i & 1, use = tile_i >> 1;
                mbar_wait(tmem_empty + acc, (use & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * kTmemCols;
                for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(full + s, (it / kStages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = smem_u32(sA + s * kTileABytes), b0 = smem_u32(sB + s * kTileBBytes);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                        umma_f16(d_tmem, smem_desc_mn_sw128(a0 + k * p.kstep, p.lbo, p.sbo),
                                 smem_desc_mn_sw128(b0 + k * p.kstep, p.lbo, p.sbo), idesc, ((kb - kb_lo) | k) ? 1u : 0u);
                    umma_commit(empty + s);
                }
                umma_commit(tmem_full + acc);
            }
        }
    } else {                                          // ---------------- epilogue (warps 2..5): fp32 tile -> partial buffer
        const int q = warp & 3;
        uint32_t tile_i = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++tile_i) {
            int sp, tap, co0, ci0, kb_lo, kb_hi;
            decode(item, sp, tap, co0, ci0, kb_lo, kb_hi);
            const uint32_t acc = tile_i & 1, use = tile_i >> 1;
            const int co = co0 + q * 32 + lane;
            float *orow = p.D + ((static_cast<size_t>(sp) * 9 + tap) * p.Cout + co) * p.Cin;
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
                if (co < p.Cout) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {     // 4 columns = one 16-byte store (Cin % 8 == 0)
                        const int ci = ci0 + j * 32 + g * 4;
                        if (ci < p.Cin)
                            *reinterpret_cast<float4 *>(orow + ci) = make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]),
                                                                                 __uint_as_float(r[g * 4 + 2]), __uint_as_float(r[g * 4 + 3]));
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

// dense NHWC bf16 tensor viewed as {C, W, H, N}; box {64 channels, 16, 4, 1}: 64 pixels x 128 bytes, SWIZZLE_128B
static bool make_map_patch(CUtensorMap *map, const void *base, int64_t n, int64_t h, int64_t w, int64_t c)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(n)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(c) * 2, static_cast<cuuint64_t>(w) * c * 2, static_cast<cuuint64_t>(h) * w * c * 2};
    const cuuint32_t box[4] = {64, wgradtc::kPatchW, wgradtc::kPatchH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int wgrad_splits(int64_t n, int64_t h, int64_t w, int64_t cin, int64_t cout)
{
    using namespace wgradtc;
    const long long kb_total = n * ((h + kPatchH - 1) / kPatchH) * ((w + kPatchW - 1) / kPatchW);
    const long long items0 = 9LL * ((cout + kBM - 1) / kBM) * ((cin + kBN - 1) / kBN);
    long long s = (2LL * kNumSMs + items0 - 1) / items0;          // aim at about two work items per SM
    if (s > kb_total) s = kb_total;
    if (s > 64) s = 64;
    if (s < 1) s = 1;
    return static_cast<int>(s);
}

}  // namespace u2pl

using namespace u2pl;

extern "C" int u2pl_conv_wgrad_splits(int64_t n, int64_t h, int64_t w, int64_t cin, int64_t cout)
{
    if (n <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0) return 0;
    return wgrad_splits(n, h, w, cin, cout);
}

extern "C" int u2pl_conv_wgrad_bf16_nhwc(const void *x, const void *gout, float *partial, int64_t n, int64_t h, int64_t w,
                                         int64_t cin, int64_t cout, int dilation, void *stream)
{
    using namespace wgradtc;
    if (n <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0 || (cin % 8) || (cout % 8) || dilation < 1)
        return bad_arg("conv_wgrad_bf16_nhwc: need Cin % 8 == 0, Cout % 8 == 0, dilation >= 1");
    if (n * h * w >= (1LL << 31)) return bad_arg("conv_wgrad_bf16_nhwc: tensor too large for 32-bit TMA coordinates");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(partial)) & 15)
        return bad_arg("conv_wgrad_bf16_nhwc: operands must be 16-byte aligned");
    CUtensorMap mg, mx;
    if (!make_map_patch(&mg, gout, n, h, w, cout) || !make_map_patch(&mx, x, n, h, w, cin)) {
        set_error("conv_wgrad_bf16_nhwc: cuTensorMapEncodeTiled failed");
        return U2PL_E_BADARG;
    }
    WgradParams p;
    p.Nimg = static_cast<int>(n); p.H = static_cast<int>(h); p.W = static_cast<int>(w);
    p.Cin = static_cast<int>(cin); p.Cout = static_cast<int>(cout); p.dil = dilation;
    p.tiles_h = (p.H + kPatchH - 1) / kPatchH; p.tiles_w = (p.W + kPatchW - 1) / kPatchW;
    p.splits = wgrad_splits(n, h, w, cin, cout);
    p.lbo = 8192; p.sbo = 1024; p.kstep = 2048;
    if (const char *e = getenv("U2PL_WGRAD_DESC")) {                   // "lbo,sbo,kstep" from the descriptor probe
        unsigned a = 0, b = 0, c = 0;
        if (sscanf(e, "%u,%u,%u", &a, &b, &c) == 3) { p.lbo = a; p.sbo = b; p.kstep = c; }
    }
    p.D = partial;
    const size_t smem = static_cast<size_t>(kStages) * (kTileABytes + kTileBBytes) + 1024 + 256;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return static_cast<int>(e); }
        configured = true;
    }
    const long long items = 9LL * ((cout + kBM - 1) / kBM) * ((cin + kBN - 1) / kBN) * p.splits;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned g = static_cast<unsigned>(items < sms ? items : sms);
    wgrad_tc_kernel<<<g, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(mg, mx, p);
    return check_launch("conv_wgrad_bf16_nhwc");
}
