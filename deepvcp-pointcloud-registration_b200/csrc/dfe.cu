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


// ---------------------------------------------------------------- backward (training) ----
// Gradient of the fused target-side embedding  out[q,c] = max_j (Wc x_j + bc)[c]  (Wc = W3 W2 W1: the three
// Linear layers have no activation between them, deep_feat_embedding.py:48-50) without the reference's float64
// [B,64,C,32,35] tensor (763 MB per pair at the KITTI shape) or its autograd graph. One warp per candidate:
// the rows x_j are rebuilt and pushed through the SAME three-layer arithmetic as the forward kernel, so the
// arg-max neighbour of every channel is the forward's; then
//     dWc[c,:] += g[c] x_{j*(c)},   dbc[c] += g[c],   dfeat[idx_j, f] += w * sum_{c: j*(c) = j} g[c] Wc[c,3+f]
// (distances, coordinates and candidates carry no gradient in the reference: knn_cuda runs under no_grad,
// the clouds are data). dWc / dbc are accumulated in registers over a warp's candidates and added to the global
// accumulators once; autograd takes them on to W1, W2, W3, b1, b2, b3 through the 32 x 35 product on the host.
constexpr int DFB_WARPS = 8;

__device__ __forceinline__ unsigned orderable(float v) {
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(DFB_WARPS * 32)
dfe_tgt_backward_kernel(const float *__restrict__ cand, Cloud txyz, const float *__restrict__ tfeat,
                        const float *__restrict__ kdist, const int32_t *__restrict__ kidx, int N, int64_t Q,
                        dvcp_dfe_params_t P, const float *__restrict__ Wc, int per_feature_weight,
                        const float *__restrict__ gout, float *__restrict__ dWc, float *__restrict__ dbc,
                        float *__restrict__ dfeat) {
    extern __shared__ __align__(16) float s_all[];
    float *s_w = s_all;                                   // the three layers, as the forward stages them
    float *s_wc = s_w + DFE_SMEM_FLOATS;                  // collapsed map [32][36] (column 35 zero)
    float *s_x = s_wc + 32 * 36 + (threadIdx.x >> 5) * (32 * 36 + 32);   // this warp's rows [32][36] + weights [32]
    dfe_stage_weights(P, s_w);
    for (int i = threadIdx.x; i < 32 * 36; i += blockDim.x) s_wc[i] = (i % 36) < DFE_IN ? Wc[(i / 36) * DFE_IN + (i % 36)] : 0.f;
    __syncthreads();
    const DfeSmem W(s_w);
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *s_wt = s_x + 32 * 36;
    float accW[32], accW2[32];   // dWc[c][lane] and (lanes 0..2) dWc[c][32 + lane]
    float accB = 0.f;            // dbc[lane]
#pragma unroll
    for (int c = 0; c < 32; ++c) accW[c] = accW2[c] = 0.f;
    const int64_t stride = (int64_t)gridDim.x * DFB_WARPS;
    for (int64_t q = (int64_t)blockIdx.x * DFB_WARPS + warp; q < Q; q += stride) {
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
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 9; ++k)
            *reinterpret_cast<float4 *>(s_x + lane * 36 + 4 * k) = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
        s_wt[lane] = (float)wl;
        float y[32];
        dfe_row(x, W, y);
        __syncwarp();
        const float g = gout[row * 32 + lane];   // lane = channel
        accB += g;
        bool won = false;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const unsigned u = orderable(y[c]);
            const unsigned m = __reduce_max_sync(0xffffffffu, u);
            const int j = __ffs(__ballot_sync(0xffffffffu, u == m)) - 1;   // ties: the lowest neighbour
            const float gc = __shfl_sync(0xffffffffu, g, c);
            accW[c] = fmaf(gc, s_x[j * 36 + lane], accW[c]);
            if (lane < 3) accW2[c] = fmaf(gc, s_x[j * 36 + 32 + lane], accW2[c]);
            y[c] = lane == j ? gc : 0.f;          // the gradient this row receives through channel c
            won |= lane == j;
        }
        if (won) {
            float *dst = dfeat + ((int64_t)b * N + id) * 32;
#pragma unroll 4
            for (int f = 0; f < 32; ++f) {
                float dx = 0.f;
#pragma unroll
                for (int c = 0; c < 32; ++c) dx = fmaf(y[c], s_wc[c * 36 + 3 + f], dx);
                atomicAdd(dst + f, dx * (per_feature_weight ? s_wt[f] : s_wt[lane]));
            }
        }
    }
    // flush: lane k holds column k of every channel's row
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        if (accW[c] != 0.f) atomicAdd(dWc + c * DFE_IN + lane, accW[c]);
        if (lane < 3 && accW2[c] != 0.f) atomicAdd(dWc + c * DFE_IN + 32 + lane, accW2[c]);
    }
    if (accB != 0.f) atomicAdd(dbc + lane, accB);
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

extern "C" int dvcp_dfe_tgt_backward(const float *cand, dvcp_cloud_t tgt_xyz, const float *tgt_feat, const float *knn_dist,
                                     const int32_t *knn_idx, int B, int N, int64_t Q, dvcp_dfe_params_t dfe,
                                     const float *w_collapsed, int quirks, const float *grad_out, float *grad_w,
                                     float *grad_b, float *grad_feat, dvcp_stream_t stream) {
    if (!cand || !tgt_xyz.base || !tgt_feat || !knn_dist || !knn_idx || !dfe_params_ok(dfe) || !w_collapsed || !grad_out ||
        !grad_w || !grad_b || !grad_feat || B <= 0 || N <= 0 || Q <= 0)
        return DVCP_E_ARG;
    if (B > 65535) return DVCP_E_UNSUPPORTED;
    const size_t smem = (DFE_SMEM_FLOATS + 32 * 36 + DFB_WARPS * (32 * 36 + 32)) * sizeof(float);
    DVCP_CUDA(cudaFuncSetAttribute(dfe_tgt_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t gx = (Q + DFB_WARPS - 1) / DFB_WARPS;
    const int64_t cap = (int64_t)DVCP_NUM_SMS * 4 / (B < 4 ? B : 4) + 1;
    if (gx > cap) gx = cap;
    dfe_tgt_backward_kernel<<<dim3((unsigned)gx, B), DFB_WARPS * 32, smem, (cudaStream_t)stream>>>(
        cand, as_cloud(tgt_xyz), tgt_feat, knn_dist, knn_idx, N, Q, dfe, w_collapsed, (quirks >> 1) & 1, grad_out, grad_w,
        grad_b, grad_feat);
    DVCP_CHECK_LAUNCH();
    return 0;
}
