// Target-side feature gathering fused with the deep-feature embedding, and the
// embedding on a materialised input.
//
// Reference: get_cat_feat_tgt.py:53-96 + deep_feat_embedding.py:23-61
// (A.8). For candidate q with neighbours j = 0..31 (indices idx[q,j], distances
// dist[q,j]) the embedding input row is
//     x_j = [ tgt_xyz[idx_j] - cand_q (3, float32),
//             float32( double(tgt_feat[idx_j, f]) * w_q[f] ), f = 0..31 ]
// with w_q[f] = double(dist[q,f]) / sum_j double(dist[q,j])  -- the weight is
// indexed by FEATURE f, not by neighbour (quirk Q7; the per-neighbour variant is
// selectable). The float64 [B,64,C,32,35] tensor of the reference is never
// materialised: rows are built in registers, one neighbour per lane.
#include "dfe_common.cuh"

namespace dvcp {

constexpr int DFE_WARPS = 8;

__global__ void __launch_bounds__(DFE_WARPS * 32)
dfe_tgt_fused_kernel(const float *__restrict__ cand, Cloud txyz, const float *__restrict__ tfeat,
                     const float *__restrict__ kdist, const int32_t *__restrict__ kidx, int N, int64_t Q,
                     dvcp_dfe_params_t P, int per_feature_weight, float *__restrict__ out) {
    __shared__ __align__(16) float s_w[DFE_SMEM_FLOATS];
    dfe_stage_weights(P, s_w);
    __syncthreads();
    const DfeSmem W(s_w);
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * DFE_WARPS;
    for (int64_t q = (int64_t)blockIdx.x * DFE_WARPS + warp; q < Q; q += stride) {
        const int64_t row = (int64_t)b * Q + q;
        const int id = kidx[row * 32 + lane];
        const double dj = (double)kdist[row * 32 + lane];
        double sum = dj;
#pragma unroll
        for (int s = 16; s; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
        const double wl = dj / sum;   // lane l holds w[l]
        const float cx = __ldg(cand + row * 3), cy = __ldg(cand + row * 3 + 1), cz = __ldg(cand + row * 3 + 2);
        float x[36];
        x[0] = txyz.at(b, id, 0) - cx;
        x[1] = txyz.at(b, id, 1) - cy;
        x[2] = txyz.at(b, id, 2) - cz;
        const float4 *fp = reinterpret_cast<const float4 *>(tfeat + ((int64_t)b * N + id) * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 v = __ldg(fp + k);
            const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ch = 4 * k + e;
                const double w = per_feature_weight ? __shfl_sync(0xffffffffu, wl, ch) : wl;
                x[3 + ch] = (float)((double)f[e] * w);
            }
        }
        x[35] = 0.f;
        float y[32];
        dfe_row(x, W, y);
        const float m = warp_colmax(y);
        out[row * 32 + lane] = m;
    }
}

// X [rows, K, 35] -> out [rows, 32]; one warp per row group, lanes over neighbours
// (K > 32 handled in several passes).
template <typename T>
__global__ void __launch_bounds__(DFE_WARPS * 32)
dfe_dense_kernel(const T *__restrict__ X, int64_t rows, int K, dvcp_dfe_params_t P, float *__restrict__ out) {
    __shared__ __align__(16) float s_w[DFE_SMEM_FLOATS];
    dfe_stage_weights(P, s_w);
    __syncthreads();
    const DfeSmem W(s_w);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * DFE_WARPS;
    for (int64_t r = (int64_t)blockIdx.x * DFE_WARPS + warp; r < rows; r += stride) {
        float best = -INFINITY;
        for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + lane;
            const bool ok = k < K;
            float x[36];
            const T *xp = X + (r * K + (ok ? k : 0)) * DFE_IN;
#pragma unroll
            for (int c = 0; c < DFE_IN; ++c) x[c] = (float)xp[c];
            x[35] = 0.f;
            float y[32];
            dfe_row(x, W, y);
            if (!ok) {
#pragma unroll
                for (int o = 0; o < 32; ++o) y[o] = -INFINITY;
            }
            best = fmaxf(best, warp_colmax(y));
        }
        out[r * 32 + lane] = best;
    }
}

}  // namespace dvcp

using namespace dvcp;

static bool dfe_params_ok(const dvcp_dfe_params_t &p) {
    return p.W1 && p.b1 && p.W2 && p.b2 && p.W3 && p.b3;
}

extern "C" int dvcp_dfe_tgt_fused(const float *cand, dvcp_cloud_t tgt_xyz, const float *tgt_feat,
                                  const float *knn_dist, const int32_t *knn_idx, int B, int N, int64_t Q,
                                  dvcp_dfe_params_t dfe, int quirks, float *out, dvcp_stream_t stream) {
    if (!cand || !tgt_xyz.base || !tgt_feat || !knn_dist || !knn_idx || !out || !dfe_params_ok(dfe) || B <= 0 ||
        N <= 0 || Q <= 0)
        return DVCP_E_ARG;
    if (B > 65535) return DVCP_E_UNSUPPORTED;
    int64_t gx = (Q + DFE_WARPS - 1) / DFE_WARPS;
    const int64_t cap = (int64_t)DVCP_NUM_SMS * 8 / (B < 8 ? B : 8) + 1;
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, B);
    dfe_tgt_fused_kernel<<<grid, DFE_WARPS * 32, 0, (cudaStream_t)stream>>>(
        cand, as_cloud(tgt_xyz), tgt_feat, knn_dist, knn_idx, N, Q, dfe, (quirks >> 1) & 1, out);
    DVCP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dvcp_dfe_dense(const void *X, int dtype, int64_t rows, int K, dvcp_dfe_params_t dfe, float *out,
                              dvcp_stream_t stream) {
    if (!X || !out || !dfe_params_ok(dfe) || rows <= 0 || K <= 0) return DVCP_E_ARG;
    int64_t gx = (rows + DFE_WARPS - 1) / DFE_WARPS;
    if (gx > DVCP_NUM_SMS * 8) gx = DVCP_NUM_SMS * 8;
    if (dtype == 0)
        dfe_dense_kernel<float><<<(unsigned)gx, DFE_WARPS * 32, 0, (cudaStream_t)stream>>>((const float *)X, rows, K, dfe, out);
    else if (dtype == 1)
        dfe_dense_kernel<double><<<(unsigned)gx, DFE_WARPS * 32, 0, (cudaStream_t)stream>>>((const double *)X, rows, K, dfe, out);
    else
        return DVCP_E_ARG;
    DVCP_CHECK_LAUNCH();
    return 0;
}
