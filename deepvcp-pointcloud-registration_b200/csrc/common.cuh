// Shared device helpers for the DeepVCP hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dvcp_b200.h"

#define DVCP_NUM_SMS 148
// The tensor-core kernels that fill the GPU (embedding, CPG) are launched as several CTAs per SM slot instead of one
// persistent CTA per SM: in a stream of batches some SMs are held for milliseconds by the sampling CTAs of the next
// batches (one CTA per cloud, the whole register file), and a CTA that waits for such an SM must not hold a fixed
// share of the work. The hardware hands the queued CTAs to whichever SMs are free (measured at K8, embedding with
// 2 / 3 / 4 / 6 / 8 / 16 CTAs per slot: 3.50 / 3.40 / 3.35 / 3.30 / 3.28 / 3.33 ms per pipelined step; alone the
// kernel pays its per-CTA set-up: 0.79 -> 0.83 ms at 8).
#define DVCP_DFE_WAVES 8
#define DVCP_CPG_WAVES 4

#define DVCP_CHECK_LAUNCH()                              \
    do {                                                 \
        cudaError_t e__ = cudaGetLastError();            \
        if (e__ != cudaSuccess) return (int)e__;         \
    } while (0)

#define DVCP_CUDA(call)                                  \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return (int)e__;         \
    } while (0)

namespace dvcp {

struct Cloud {
    const float *p;
    int64_t bs, ps, cs;
    __device__ __forceinline__ float at(int b, int n, int c) const {
        return __ldg(p + (int64_t)b * bs + (int64_t)n * ps + (int64_t)c * cs);
    }
};
struct CloudD {
    const double *p;
    int64_t bs, ps, cs;
    __device__ __forceinline__ double at(int b, int n, int c) const {
        return __ldg(p + (int64_t)b * bs + (int64_t)n * ps + (int64_t)c * cs);
    }
};
static inline Cloud as_cloud(const dvcp_cloud_t &c) {
    return Cloud{(const float *)c.base, c.bstride, c.pstride, c.cstride};
}
static inline CloudD as_cloud_d(const dvcp_cloud_t &c) {
    return CloudD{(const double *)c.base, c.bstride, c.pstride, c.cstride};
}

// ---- exact-arithmetic building blocks (never contracted by the compiler) ----
// pointnet2_utils.py:80  sum((xyz - c)**2, -1): each square rounded, left to right.
__device__ __forceinline__ float sq3_nofma(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
// pointnet2_utils.py:35-40 expanded form with a K=3 SGEMM dot.
__device__ __forceinline__ float sqdist_expanded(float qx, float qy, float qz, float qq, float px,
                                                 float py, float pz, float pp) {
    float dot = __fmaf_rn(qz, pz, __fmaf_rn(qy, py, __fmul_rn(qx, px)));
    return __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), qq), pp);
}
// knn_cuda contract: ssd += d*d with contraction, x then y then z.
__device__ __forceinline__ float sqdist_direct(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// ---- key of the spatial index: position of a 10-bit-per-axis cell on a space-filling curve ----
// The consumers only need "32 consecutive points of the order form a compact box"; the order itself never
// changes a result (members are always decided by the exact arithmetic). The Hilbert curve has no jumps, so
// the boxes of consecutive runs are tighter than along the Z-order (Morton) curve, whose runs straddle
// octant boundaries: fewer buckets pass the box tests of the KNN / ball-query / sampling kernels.
__device__ __forceinline__ unsigned spread_bits10(unsigned v) {   // bit i -> bit 3 i
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ unsigned spatial_key(unsigned x, unsigned y, unsigned z) {
#ifndef DVCP_MORTON_INDEX
    // Skilling's axes -> transposed Hilbert index (3 axes, 10 bits), then interleaved
    unsigned X[3] = {x, y, z};
#pragma unroll
    for (unsigned Q = 512u; Q > 1u; Q >>= 1) {
        const unsigned P = Q - 1u;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) {
                X[0] ^= P;
            } else {
                const unsigned t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    unsigned t = 0u;
#pragma unroll
    for (unsigned Q = 512u; Q > 1u; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1u;
    x = X[0] ^ t; y = X[1] ^ t; z = X[2] ^ t;
#endif
    return (spread_bits10(x) << 2) | (spread_bits10(y) << 1) | spread_bits10(z);
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// argmax over a warp of (value bits, tie -> smaller `lo` payload wins is encoded
// by the caller as a larger `tie` word). Both words are unsigned; returns the
// lexicographic maximum of (hi, lo) over the warp, broadcast to every lane.
__device__ __forceinline__ void warp_max_pair(unsigned &hi, unsigned &lo) {
    unsigned m = __reduce_max_sync(0xffffffffu, hi);
    unsigned l = (hi == m) ? lo : 0u;
    lo = __reduce_max_sync(0xffffffffu, l);
    hi = m;
}

}  // namespace dvcp
