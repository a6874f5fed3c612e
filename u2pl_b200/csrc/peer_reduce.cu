// peer_reduce.cu -- all-reduce (sum) of a SMALL fp32 vector across the GPUs of one NVSwitch box, entirely inside one
// single-CTA kernel over peer-mapped memory (CUDA IPC): no NCCL launch, no proxy thread, no host involvement.
//
// Why: SyncBatchNorm (reference base.py:6-8 -> nn.SyncBatchNorm) exchanges the per-channel [sum | sum of squares]
// ([2,C] floats, <= 16 KB) of every norm layer, forward and backward: ~350 exchanges per training step on the
// ResNet-101 DeepLabv3+ (117 layers x student forward, teacher train-mode forward, student backward).  Each NCCL
// all_reduce of 8 KB costs a kernel launch + ring/tree protocol latency (15-25 us at 8 ranks) on the compute stream --
// 5-8 ms per step that scale-out pays and a single GPU does not (SCALE_r01: 0.913 efficiency at 8).  Here every rank
//   1. PUSHES its vector into slot [parity][rank] of EVERY peer's exchange region (NVLink stores, 16 KB x (W-1)),
//   2. fences (system scope) and raises flag [parity][rank] = seq in every peer's region,
//   3. waits until its OWN region shows seq in all W flags,
//   4. sums the W slots in rank order (the same order on every rank -> bitwise identical results everywhere).
// One-shot, latency = one NVLink store + one flag round (~3-5 us).  Regions are double-buffered by the parity of
// `seq`: a rank can only reach call seq+2 after every peer raised its flag for seq+1, i.e. after every peer finished
// reading the slots of call seq, so slot reuse never races.  All ranks must issue the same sequence of calls (like any
// collective); calls are stream-ordered.
#include "common.cuh"

namespace u2pl {

constexpr int kPeerMaxWorld = 8;
constexpr int kPeerMaxFloats = 4096;                      // 2 x 2048 channels

struct PeerRegions { float *base[kPeerMaxWorld]; };

__host__ __device__ inline size_t peer_region_bytes()
{
    return static_cast<size_t>(2) * kPeerMaxWorld * kPeerMaxFloats * 4 + 2 * kPeerMaxWorld * 4 + 64;
}

__device__ __forceinline__ float *slot_of(float *region, int parity, int r) { return region + (static_cast<size_t>(parity) * kPeerMaxWorld + r) * kPeerMaxFloats; }
__device__ __forceinline__ uint32_t *flags_of(float *region) { return reinterpret_cast<uint32_t *>(region + static_cast<size_t>(2) * kPeerMaxWorld * kPeerMaxFloats); }

__global__ void __launch_bounds__(512)
peer_allreduce_kernel(float *__restrict__ buf, int n, PeerRegions peers, int rank, int world, uint32_t seq)
{
    const int parity = static_cast<int>(seq & 1u), tid = threadIdx.x;
    for (int p = 0; p < world; ++p) {                      // 1. push
        float *dst = slot_of(peers.base[p], parity, rank);
        for (int i = tid; i < n; i += 512) dst[i] = buf[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) {                                     // 2. raise my flag everywhere, 3. wait for everybody's flag here
        uint32_t *f = flags_of(peers.base[tid]) + parity * kPeerMaxWorld + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
        const uint32_t *mine = flags_of(peers.base[rank]) + parity * kPeerMaxWorld + tid;
        uint32_t v;
        do { asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory"); } while (v != seq);
    }
    __syncthreads();
    for (int i = tid; i < n; i += 512) {                   // 4. reduce, rank order
        float s = 0.0f;
        for (int p = 0; p < world; ++p) {
            float v;
            asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(slot_of(peers.base[rank], parity, p) + i) : "memory");
            s += v;
        }
        buf[i] = s;
    }
}

}  // namespace u2pl

using namespace u2pl;

extern "C" int64_t u2pl_peer_region_bytes(void) { return static_cast<int64_t>(peer_region_bytes()); }
extern "C" int64_t u2pl_peer_max_floats(void) { return kPeerMaxFloats; }

extern "C" int u2pl_peer_allreduce_f32(float *buf, int64_t n, void *const *peer_bases, int rank, int world, uint32_t seq, void *stream)
{
    if (!buf || !peer_bases || n <= 0 || n > kPeerMaxFloats || world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world || seq == 0)
        return bad_arg("peer_allreduce_f32: need 0 < n <= 4096, 1 <= world <= 8, seq >= 1");
    PeerRegions pr;
    for (int p = 0; p < kPeerMaxWorld; ++p) pr.base[p] = p < world ? static_cast<float *>(peer_bases[p]) : nullptr;
    peer_allreduce_kernel<<<1, 512, 0, static_cast<cudaStream_t>(stream)>>>(buf, static_cast<int>(n), pr, rank, world, seq);
    return check_launch("peer_allreduce_f32");
}
