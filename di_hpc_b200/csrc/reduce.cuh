// reduce.cuh -- deterministic loss reductions.
//
// Every mean-type loss of the path is reduced in two fixed-order stages so results are reproducible
// run to run (the reference uses float atomicAdd per block, e.g.
// include/hpc/rll/cuda/rl_utils/td_lambda_kernel.h:35-38, vtrace_kernel.h:197-223, whose summation
// order varies):
//   1. each thread accumulates its terms in fp64, the CTA tree-reduces in fp64 and writes ONE
//      double per loss term to partials[term * nblocks + blockIdx.x]
//   2. finalize_sums<<<1, 256>>> adds the per-CTA partials in a fixed order, scales, writes fp32.
// Replaces reduce.h:13-99 (warpReduce*/blockReduce*) of the reference; re-derived, shuffle based.
#pragma once
#include "common.cuh"

namespace hpcrll {

#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum NT values per thread over the whole CTA (all threads must call; blockDim.x <= 1024).
// Thread 0 returns the totals in v[]; `scratch` needs NT*32 doubles of shared memory.
template <int NT>
__device__ __forceinline__ void block_sum(double (&v)[NT], double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NT; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();  // scratch may be in use by a previous call
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NT; ++k) scratch[k * 32 + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            double x = lane < nwarp ? scratch[k * 32 + lane] : 0.0;
            v[k] = warp_sum(x);
        }
    }
}

// out[k] = scale[k] * sum_{i<n} partials[k*n + i], fixed order.  One block of 256 threads.
template <int NT>
__global__ void __launch_bounds__(256) finalize_sums(const double* __restrict__ partials, int n,
                                                      const double* __restrict__ scale_dev, double s0, double s1,
                                                      double s2, double s3, double s4, float* __restrict__ out) {
    __shared__ double scratch[NT * 32];
    double v[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        double a = 0.0;
        const double* p = partials + static_cast<size_t>(k) * n;
        for (int i0 = threadIdx.x; i0 < n; i0 += 256 * 8) {  // 8 loads in flight, adds in the original order
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = __ldcg(p + min(i0 + u * 256, n - 1));  // clamped index: no predicate
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (i0 + u * 256 < n) a += x[u];
        }
        v[k] = a;
    }
    block_sum<NT>(v, scratch);
    if (threadIdx.x == 0) {
        const double s[5] = {s0, s1, s2, s3, s4};
        (void)scale_dev;
#pragma unroll
        for (int k = 0; k < NT; ++k) out[k] = static_cast<float>(v[k] * s[k]);
    }
}

#endif  // __CUDACC__

int launch_scale_copy(const float* in, const float* scale_dev, float* out, int64_t n, int64_t zero_tail,
                      cudaStream_t stream);

}  // namespace hpcrll
