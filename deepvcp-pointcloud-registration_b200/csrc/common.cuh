// Shared device helpers for the DeepVCP hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dvcp_b200.h"

#define DVCP_NUM_SMS 148

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
