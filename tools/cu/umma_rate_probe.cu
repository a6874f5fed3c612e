// umma_rate_probe.cu -- raw issue rate of tcgen05.mma kind::f16 from shared-memory operands, no loads in the loop:
//   cta_group::1  M128 N256 K16   (one CTA per SM, every SM)        expected 128 clk / instruction
//   cta_group::2  M256 N256 K16   (one CTA pair per TPC, leader issues)   expected 128 clk / instruction per pair
// Prints clocks per instruction and the implied dense bf16 TFLOP/s of the chip at the measured SM clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/cu/umma_rate_probe.bin tools/cu/umma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../u2pl_b200/csrc/tc_common.cuh"

using namespace u2pl;

constexpr int kIters = 2048;          // k-blocks of 4 MMAs each
constexpr int kTile = 128 * 64 * 2;   // 16 KB

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int kCG>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *sA = smem, *sB = smem + kTile;                        // A: 128 x 64, B: 256 x 64 (cg1) or this CTA's 128 x 64 half (cg2)
    uint64_t *done = reinterpret_cast<uint64_t *>(smem + 3 * kTile);
    uint32_t *slot = reinterpret_cast<uint32_t *>(done + 1);
    for (int i = threadIdx.x; i < 3 * kTile / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = kCG == 2 ? cluster_rank() : 0;
    if (threadIdx.x == 0) { mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        if (kCG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(256) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(256) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (kCG == 2) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (warp == 0 && lane == 0 && rank == 0) {
        const int M = kCG == 2 ? 256 : 128;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(256 >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        const long long t0 = clock64();
        for (int it = 0; it < kIters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (kCG == 2)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(smem_desc_sw128(a0 + 32 * k)), "l"(smem_desc_sw128(b0 + 32 * k)), "r"(idesc), "r"(1u) : "memory");
                else
                    umma_f16(tmem, smem_desc_sw128(a0 + 32 * k), smem_desc_sw128(b0 + 32 * k), idesc, 1u);
            }
        }
        if (kCG == 2) {
            const uint16_t mask = 3;
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(done)), "h"(mask) : "memory");
        } else {
            umma_commit(done);
        }
        mbar_wait(done, 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    } else if (kCG == 2 && warp == 0 && lane == 0) {
        mbar_wait(done, 0);                                        // peer: keep the CTA (its smem / TMEM) alive until the MMAs finish
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (kCG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        if (kCG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
    }
}

int main()
{
    int sms = 148, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    long long *d, h[512];
    cudaMalloc(&d, sizeof(h));
    const int smem = 3 * kTile + 1024 + 64;
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int cg = 1; cg <= 2; ++cg) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(d, 0, sizeof(h));
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            if (cg == 1) rate_kernel<1><<<sms, 128, smem>>>(d);
            else {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(sms & ~1); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
                cfg.attrs = &at; cfg.numAttrs = 1;
                cudaLaunchKernelEx(&cfg, rate_kernel<2>, d);
            }
            cudaEventRecord(b);
            const cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("cg%d: CUDA ERROR %s\n", cg, cudaGetErrorString(e)); return 1; }
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0; int n = 0;
            for (int i = 0; i < sms; ++i) if (h[i] > 0) { mx = h[i] > mx ? h[i] : mx; ++n; }
            const double clk_per_mma = static_cast<double>(mx) / (kIters * 4.0);
            const double flop = 2.0 * (cg == 2 ? 256.0 : 128.0) * 256.0 * 16.0 * kIters * 4.0 * n;
            printf("cta_group::%d  issuers %d  clocks/MMA %.1f  kernel %.3f ms  -> %.0f TFLOP/s (event-timed, incl. launch)\n", cg, n, clk_per_mma, ms, flop / (ms * 1e-3) / 1e12);
        }
    }
    return 0;
}
