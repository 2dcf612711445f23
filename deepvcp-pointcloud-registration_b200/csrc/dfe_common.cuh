// Deep-feature-embedding "mini-PointNet" (deep_feat_embedding.py:23-61):
// three affine maps 35 -> 32 -> 32 -> 32 without activations, then a max over
// the neighbour axis. CUDA-core float32 form: one neighbour row per lane, weights
// broadcast from shared memory as float4.
#pragma once
#include "common.cuh"

namespace dvcp {

constexpr int DFE_IN = 35;
constexpr int DFE_W1_LD = 36;                     // rows padded to a float4 multiple
constexpr int DFE_SMEM_FLOATS = 32 * DFE_W1_LD + 32 * 32 + 32 * 32 + 3 * 32;

struct DfeSmem {
    float *w1, *w2, *w3, *b1, *b2, *b3;
    __device__ explicit DfeSmem(float *base)
        : w1(base), w2(base + 32 * DFE_W1_LD), w3(w2 + 1024), b1(w3 + 1024), b2(b1 + 32), b3(b2 + 32) {}
};

__device__ __forceinline__ void dfe_stage_weights(const dvcp_dfe_params_t &p, float *base) {
    DfeSmem s(base);
    for (int i = threadIdx.x; i < 32 * DFE_W1_LD; i += blockDim.x) {
        const int o = i / DFE_W1_LD, k = i - o * DFE_W1_LD;
        s.w1[i] = k < DFE_IN ? p.W1[o * DFE_IN + k] : 0.f;
    }
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        s.w2[i] = p.W2[i];
        s.w3[i] = p.W3[i];
    }
    for (int i = threadIdx.x; i < 32; i += blockDim.x) {
        s.b1[i] = p.b1[i];
        s.b2[i] = p.b2[i];
        s.b3[i] = p.b3[i];
    }
}

// x[36] (x[35] ignored, must be finite) -> y[32]
__device__ __forceinline__ void dfe_row(const float (&x)[36], const DfeSmem &s, float (&y)[32]) {
    float h[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        float acc = s.b1[o];
        const float4 *w = reinterpret_cast<const float4 *>(s.w1 + o * DFE_W1_LD);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float4 v = w[k];
            acc = fmaf(v.x, x[4 * k], acc);
            acc = fmaf(v.y, x[4 * k + 1], acc);
            acc = fmaf(v.z, x[4 * k + 2], acc);
            acc = fmaf(v.w, x[4 * k + 3], acc);   // k == 8: weight is 0 for column 35
        }
        h[o] = acc;
    }
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        float acc = s.b2[o];
        const float4 *w = reinterpret_cast<const float4 *>(s.w2 + o * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 v = w[k];
            acc = fmaf(v.x, h[4 * k], acc);
            acc = fmaf(v.y, h[4 * k + 1], acc);
            acc = fmaf(v.z, h[4 * k + 2], acc);
            acc = fmaf(v.w, h[4 * k + 3], acc);
        }
        y[o] = acc;
    }
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        float acc = s.b3[o];
        const float4 *w = reinterpret_cast<const float4 *>(s.w3 + o * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 v = w[k];
            acc = fmaf(v.x, y[4 * k], acc);
            acc = fmaf(v.y, y[4 * k + 1], acc);
            acc = fmaf(v.z, y[4 * k + 2], acc);
            acc = fmaf(v.w, y[4 * k + 3], acc);
        }
        h[o] = acc;
    }
#pragma unroll
    for (int o = 0; o < 32; ++o) y[o] = h[o];
}

// max over the 32 lanes of every y[o]; lane o ends up holding channel o's max.
// Butterfly transpose-reduce: 31 shuffles. y is clobbered.
__device__ __forceinline__ float warp_colmax(float (&y)[32]) {
    const unsigned lane = lane_id();
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? y[i] : y[i + s];
            const float keep = up ? y[i + s] : y[i];
            y[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, s));
        }
    }
    return y[0];
}

}  // namespace dvcp
