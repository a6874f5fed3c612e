// umma_2cta_probe.cu -- does this cta_group::2 protocol compute a correct 256x256 tile on a CTA pair?
//
// Why: the 1-CTA kernels of u2pl_b200/csrc fill 94-128 bytes of shared memory per tensor-core cycle per SM (DESIGN.md
// 4c/4d), which is what caps them below cuBLAS / cuDNN's 2-SM kernels.  A CTA pair shares the B operand (each CTA holds
// half of its N rows) and halves that to 64 B/clk/SM.  The pair protocol has many moving parts that cannot be checked
// without a GPU; this probe isolates them in ~200 lines with a CPU reference, so that they can be pinned in seconds
// before conv_tc / wgrad_tc / gemm_tc are converted:
//   * cluster launch {2,1,1}, cluster rank, cluster barriers
//   * tcgen05.alloc / dealloc .cta_group::2 issued by one warp of EACH CTA
//   * both CTAs' TMA loads (cp.async.bulk.tensor .cta_group::2) completing on the LEADER's `full` mbarrier
//     (address mapped with mapa), the peer producer's remote arrive on it
//   * tcgen05.mma.cta_group::2 (M = 256, N = 256) issued by the leader only
//   * tcgen05.commit .cta_group::2 .multicast::cluster arriving on `empty` / `done` in both CTAs
// D[256][256] = A[256][K] . B[256][K]^T, K = 256 (4 K blocks through a 2-stage ring).  CTA r holds A rows
// [128r, 128r+128) and B rows [128r, 128r+128); its TMEM holds D rows [128r, 128r+128) x 256 columns.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/cu/umma_2cta_probe.bin tools/cu/umma_2cta_probe.cu
//   timeout 20 ./tools/cu/umma_2cta_probe.bin        (use timeout: a protocol error shows up as a hang)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../u2pl_b200/csrc/tc_common.cuh"

using namespace u2pl;

constexpr int kM = 256, kN = 256, kK = 256, kBK = 64, kStages = 2;
constexpr int kHalf = 128;
constexpr int kTileBytes = kHalf * kBK * 2;                       // 16 KB: one operand half, one K block

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
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank)      // shared::cta address -> shared::cluster address of `rank`
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void remote_arrive(uint32_t bar_cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float *D)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *sA = smem, *sB = smem + kStages * kTileBytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + kStages * kTileBytes);     // used in the leader CTA only
    uint64_t *empty = full + kStages;                                             // one copy per CTA
    uint64_t *done = empty + kStages;                                             // one copy per CTA
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int nkb = kK / kBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 2); mbar_init(empty + s, 1); }    // full: one arrive per CTA's producer
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                            // barrier inits + allocation visible in both CTAs
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {                                  // ---------------- TMA producer (both CTAs)
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kStages;
            mbar_wait(empty + s, ((kb / kStages) & 1) ^ 1);
            const uint32_t leader_full = map_to_cta(smem_u32(full + s), 0);
            if (rank == 0) mbar_expect_tx(full + s, 4 * kTileBytes);            // both CTAs' A and B halves: 64 KB
            else remote_arrive(leader_full);
            tma_load_2d_pair(sA + s * kTileBytes, &map_a, leader_full, kb * kBK, static_cast<int>(rank) * kHalf);
            tma_load_2d_pair(sB + s * kTileBytes, &map_b, leader_full, kb * kBK, static_cast<int>(rank) * kHalf);
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {              // ---------------- MMA issuer (leader only)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kN >> 3) << 17) |
                               (static_cast<uint32_t>(kM >> 4) << 24);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kStages;
            mbar_wait(full + s, (kb / kStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = smem_u32(sA + s * kTileBytes), b0 = smem_u32(sB + s * kTileBytes);
            for (int k = 0; k < kBK / 16; ++k)
                umma_f16_pair(tmem_base, smem_desc_sw128(a0 + 32 * k), smem_desc_sw128(b0 + 32 * k), idesc, (kb | k) ? 1u : 0u);
            umma_commit_pair(empty + s);                            // frees stage s in BOTH CTAs
        }
        umma_commit_pair(done);                                     // accumulators complete, both CTAs
    } else if (warp >= 2) {                                         // ---------------- epilogue: each CTA drains its 128 rows
        const int q = warp & 3;
        mbar_wait(done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = static_cast<int>(rank) * kHalf + q * 32 + lane;
        for (int j = 0; j < kN / 32; ++j) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(j * 32), r);
            for (int e = 0; e < 32; ++e) D[static_cast<size_t>(row) * kN + j * 32 + e] = __uint_as_float(r[e]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                             // nobody frees TMEM / exits while the peer still uses it
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
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

int main()
{
    std::vector<float> A(kM * kK), B(kN * kK), ref(static_cast<size_t>(kM) * kN);
    uint32_t seed = 4242u;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return static_cast<float>(static_cast<int>((seed >> 24) % 7) - 3); };
    for (auto &v : A) v = rnd();
    for (auto &v : B) v = rnd();
    for (int m = 0; m < kM; ++m)
        for (int n = 0; n < kN; ++n) {
            float acc = 0.0f;
            for (int k = 0; k < kK; ++k) acc += A[m * kK + k] * B[n * kK + k];
            ref[static_cast<size_t>(m) * kN + n] = acc;
        }
    std::vector<__nv_bfloat16> hA(A.size()), hB(B.size());
    for (size_t i = 0; i < A.size(); ++i) hA[i] = __float2bfloat16(A[i]);
    for (size_t i = 0; i < B.size(); ++i) hB[i] = __float2bfloat16(B[i]);
    __nv_bfloat16 *dA, *dB;
    float *dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, ref.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, ref.size() * 4);
    CUtensorMap ma, mb;
    if (!make_map2d(&ma, dA, kK, kM, kBK, kHalf) || !make_map2d(&mb, dB, kK, kN, kBK, kHalf)) { fprintf(stderr, "tensor map encode failed\n"); return 3; }
    const int smem = 2 * kStages * kTileBytes + 1024 + 128;
    cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    pair_kernel<<<2, 192, smem>>>(ma, mb, dD);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("2cta probe: CUDA ERROR %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> out(ref.size());
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, bad_top = 0;
    float worst = 0.0f;
    for (size_t i = 0; i < out.size(); ++i) {
        const float d = fabsf(out[i] - ref[i]);
        if (!(d <= 1e-3f)) { ++bad; if (i < out.size() / 2) ++bad_top; }
        if (d > worst) worst = d;
    }
    printf("2cta probe: mismatches=%d/%zu (rows 0-127: %d, rows 128-255: %d) max_abs_err=%g  %s\n", bad, out.size(), bad_top, bad - bad_top,
           worst, bad == 0 ? "MATCH" : "wrong");
    return bad == 0 ? 0 : 1;
}
