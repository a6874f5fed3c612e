// common.cuh -- shared host-side plumbing of libu2pl_b200.so.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/u2pl_b200.h"

namespace u2pl {

void set_error(const char *msg);           // c_abi.cu
void count_launch(int n = 1);              // c_abi.cu

inline int check_launch(const char *what, int n = 1)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        static thread_local char buf[256];
        snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
        set_error(buf);
        return static_cast<int>(e);
    }
    count_launch(n);
    return 0;
}

int bn_reduce_parts(const float *partial, int nparts, int c2, float *sums, void *stream);   // bn.cu: [nparts][c2] -> [c2]

inline int bad_arg(const char *msg) { set_error(msg); return U2PL_E_BADARG; }

constexpr int kNumSMs = 148;               // B200

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace u2pl
