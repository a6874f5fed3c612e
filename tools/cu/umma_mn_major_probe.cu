// umma_mn_major_probe.cu -- which shared-memory descriptor makes tcgen05.mma read MN-MAJOR operands correctly?
//
// Why: the weight gradient of a convolution is  dW[co, ci] = sum_pix gout[pix, co] * x[pix + shift, ci].  Both operands
// are channels-last tensors, i.e. the contraction index (pixel) is the SLOW dimension of both: "MN-major" operands in
// UMMA terms.  The kernels in u2pl_b200/csrc (gemm_tc.cu, conv_tc.cu) only use K-major SWIZZLE_128B descriptors, which
// were validated on a B200; an MN-major weight-gradient kernel needs the other descriptor flavour (instruction
// descriptor bits 15/16 = transpose A / B, and LBO / SBO with their MN-major meaning).  This probe runs ONE 128x128
// tile, K = 128, with operands TMA-loaded from [K][M] / [K][N] row-major matrices (64-element = 128-byte boxes,
// SWIZZLE_128B), for a descriptor candidate given on the command line, and compares against a CPU product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/umma_probe tools/cu/umma_mn_major_probe.cu
//   /tmp/umma_probe <a_mn 0|1> <b_mn 0|1> <lbo_bytes> <sbo_bytes> <kstep_bytes>
//   tools/umma_probe_sweep.sh   (one process per candidate: a malformed descriptor may kill the context)
//
// Smem image of an MN-major operand stage (what TMA writes for two {64 MN, 64 K} boxes):
//   box b (MN elements 64b .. 64b+63) at byte 8192*b; inside a box K-row r at byte 128*r (16-byte chunks XOR-swizzled
//   with r % 8), so 8 K-rows form one 1024-byte swizzle atom.
// Expected (CUTLASS convention, to be confirmed here): a_mn = b_mn = 1, LBO = 8192 (next 64 MN elements),
// SBO = 1024 (next 8 K rows), K step of 16 = 2048 bytes.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../u2pl_b200/csrc/tc_common.cuh"

using namespace u2pl;

constexpr int kM = 128, kN = 128, kK = 128, kBK = 64;

struct ProbeParams {
    int a_mn, b_mn;
    uint32_t lbo, sbo, kstep;         // bytes
    float *D;                         // [kM][kN]
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return static_cast<uint64_t>((saddr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, ProbeParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    constexpr int kStage = 16384;                                 // one operand, one 64-wide K block
    uint8_t *sA = smem, *sB = smem + (kK / kBK) * kStage;
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + (kK / kBK) * kStage);
    uint64_t *done = full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(full, 1);
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(full, 2 * (kK / kBK) * kStage);
        for (int kb = 0; kb < kK / kBK; ++kb) {
            if (p.a_mn) {                                         // [K][M]: inner coordinate = M element, outer = K row
                tma_load_2d(sA + kb * kStage, &map_a, full, 0, kb * kBK);
                tma_load_2d(sA + kb * kStage + 8192, &map_a, full, 64, kb * kBK);
            } else {                                              // [M][K]: one {64 K, 128 M} box
                tma_load_2d(sA + kb * kStage, &map_a, full, kb * kBK, 0);
            }
            if (p.b_mn) {
                tma_load_2d(sB + kb * kStage, &map_b, full, 0, kb * kBK);
                tma_load_2d(sB + kb * kStage + 8192, &map_b, full, 64, kb * kBK);
            } else {
                tma_load_2d(sB + kb * kStage, &map_b, full, kb * kBK, 0);
            }
        }
        mbar_wait(full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(p.a_mn) << 15) |
                               (static_cast<uint32_t>(p.b_mn) << 16) | (static_cast<uint32_t>(kN >> 3) << 17) |
                               (static_cast<uint32_t>(kM >> 4) << 24);
        for (int kb = 0; kb < kK / kBK; ++kb)
            for (int k = 0; k < kBK / 16; ++k) {
                const uint32_t a0 = smem_u32(sA + kb * kStage), b0 = smem_u32(sB + kb * kStage);
                // K-major operands keep the validated encoding (LBO field 1, SBO 1024, +32 bytes per K step)
                const uint64_t da = p.a_mn ? make_desc(a0 + k * p.kstep, p.lbo, p.sbo) : smem_desc_sw128(a0 + 32 * k);
                const uint64_t db = p.b_mn ? make_desc(b0 + k * p.kstep, p.lbo, p.sbo) : smem_desc_sw128(b0 + 32 * k);
                umma_f16(tmem_base, da, db, idesc, (kb | k) ? 1u : 0u);
            }
        umma_commit(done);
    }
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;
    for (int j = 0; j < kN / 32; ++j) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(j * 32), r);
        for (int e = 0; e < 32; ++e) p.D[row * kN + j * 32 + e] = __uint_as_float(r[e]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128) : "memory");
}

static bool make_map2d(CUtensorMap *map, const void *base, int inner, int outer, int box_inner, int box_outer)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(inner) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: %s a_mn b_mn lbo_bytes sbo_bytes kstep_bytes\n", argv[0]); return 2; }
    ProbeParams p;
    p.a_mn = atoi(argv[1]); p.b_mn = atoi(argv[2]);
    p.lbo = static_cast<uint32_t>(atoi(argv[3])); p.sbo = static_cast<uint32_t>(atoi(argv[4])); p.kstep = static_cast<uint32_t>(atoi(argv[5]));
    // logical operands A[m][k], B[n][k]; small integers so that every partial sum is exact in bf16 x bf16 -> fp32
    std::vector<float> A(kM * kK), B(kN * kK), ref(kM * kN, 0.0f);
    uint32_t seed = 12345u;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return static_cast<float>(static_cast<int>((seed >> 24) % 7) - 3); };
    for (auto &v : A) v = rnd();
    for (auto &v : B) v = rnd();
    for (int m = 0; m < kM; ++m)
        for (int n = 0; n < kN; ++n) {
            float acc = 0.0f;
            for (int k = 0; k < kK; ++k) acc += A[m * kK + k] * B[n * kK + k];
            ref[m * kN + n] = acc;
        }
    // device images: K-major = [rows][K]; MN-major = [K][rows]
    auto image = [&](const std::vector<float> &X, int rows, int mn) {
        std::vector<__nv_bfloat16> h(static_cast<size_t>(rows) * kK);
        for (int r = 0; r < rows; ++r)
            for (int k = 0; k < kK; ++k) h[mn ? (static_cast<size_t>(k) * rows + r) : (static_cast<size_t>(r) * kK + k)] = __float2bfloat16(X[r * kK + k]);
        return h;
    };
    const auto hA = image(A, kM, p.a_mn), hB = image(B, kN, p.b_mn);
    __nv_bfloat16 *dA, *dB;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&p.D, kM * kN * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(p.D, 0xff, kM * kN * 4);
    CUtensorMap ma, mb;
    const bool ok = (p.a_mn ? make_map2d(&ma, dA, kM, kK, 64, kBK) : make_map2d(&ma, dA, kK, kM, kBK, kM)) &&
                    (p.b_mn ? make_map2d(&mb, dB, kN, kK, 64, kBK) : make_map2d(&mb, dB, kK, kN, kBK, kN));
    if (!ok) { fprintf(stderr, "cuTensorMapEncodeTiled failed\n"); return 3; }
    const int smem = 4 * 16384 + 1024 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(ma, mb, p);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("a_mn=%d b_mn=%d lbo=%u sbo=%u kstep=%u  CUDA ERROR %s\n", p.a_mn, p.b_mn, p.lbo, p.sbo, p.kstep, cudaGetErrorString(e)); return 1; }
    std::vector<float> out(kM * kN);
    cudaMemcpy(out.data(), p.D, out.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    float worst = 0.0f;
    for (size_t i = 0; i < out.size(); ++i) {
        const float d = fabsf(out[i] - ref[i]);
        if (!(d <= 1e-3f)) ++bad;
        if (d > worst) worst = d;
    }
    printf("a_mn=%d b_mn=%d lbo=%u sbo=%u kstep=%u  mismatches=%d/%d  max_abs_err=%g  %s\n", p.a_mn, p.b_mn, p.lbo, p.sbo, p.kstep,
           bad, kM * kN, worst, bad == 0 ? "MATCH" : "wrong");
    return bad == 0 ? 0 : 1;
}
