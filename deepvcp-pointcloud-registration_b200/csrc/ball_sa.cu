// Ball query, grouping + shared MLP + max (the set-abstraction layer), the
// weighting MLP and the top-K key-point selection.
//
// Reference: pointnet2_utils.py:19-60,87-138,176-202; weighting_layer.py:26-33.
//
// Ball membership: d2 = ((-2 * dot) + |q|^2) + |p|^2 with
// dot = fma(qz,pz, fma(qy,py, qx*px)), |.|^2 = (x*x + y*y) + z*z, all float32
// round-to-nearest; a point is a member unless d2 > float32(radius**2); the first
// `nsample` members in ascending index order are kept and short lists are padded
// with their first entry.
//
// Layout: the cloud of one batch item is staged once per CTA into shared memory
// as x[], y[], z[], |p|^2[] (coalesced float loads, conflict-free LDS); each warp
// owns one query at a time and walks the cloud 32 points per step, so the
// ascending-index order falls out of ballot/popc and no sort is needed.
#include "common.cuh"

namespace dvcp {

constexpr int BQ_TILE = 8192;   // points per shared-memory tile (4 floats each = 128 KB)
constexpr int BQ_WARPS = 16;

__device__ __forceinline__ float norm2_nofma(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

__device__ __forceinline__ void stage_tile(const Cloud &c, int b, int base, int count, float *sx,
                                           float *sy, float *sz, float *sp) {
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const float x = c.at(b, base + i, 0), y = c.at(b, base + i, 1), z = c.at(b, base + i, 2);
        sx[i] = x;
        sy[i] = y;
        sz[i] = z;
        sp[i] = norm2_nofma(x, y, z);
    }
}

// ------------------------------------------------------------- ball query ----
// grid (ceil(S / (BQ_WARPS * QPW)), B); each warp handles QPW consecutive queries.
constexpr int BQ_QPW = 4;

__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(Cloud xyz, Cloud qry, int N, int S, float r2, int nsample, int64_t *__restrict__ out) {
    extern __shared__ float smem[];
    float *sx = smem, *sy = sx + BQ_TILE, *sz = sy + BQ_TILE, *sp = sz + BQ_TILE;
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = (blockIdx.x * BQ_WARPS + warp) * BQ_QPW;
    float qx[BQ_QPW], qy[BQ_QPW], qz[BQ_QPW], qq[BQ_QPW];
    int cnt[BQ_QPW];
    int first[BQ_QPW];
#pragma unroll
    for (int j = 0; j < BQ_QPW; ++j) {
        const int q = min(q0 + j, S - 1);
        qx[j] = qry.at(b, q, 0);
        qy[j] = qry.at(b, q, 1);
        qz[j] = qry.at(b, q, 2);
        qq[j] = norm2_nofma(qx[j], qy[j], qz[j]);
        cnt[j] = (q0 + j < S) ? 0 : nsample;   // out-of-range queries are "full" from the start
        first[j] = N;
    }
    for (int base = 0; base < N; base += BQ_TILE) {
        const int count = min(BQ_TILE, N - base);
        __syncthreads();
        stage_tile(xyz, b, base, count, sx, sy, sz, sp);
        __syncthreads();
        bool all_full = true;
#pragma unroll
        for (int j = 0; j < BQ_QPW; ++j) all_full &= cnt[j] >= nsample;
        if (all_full) continue;
        for (int i = 0; i < count; i += 32) {
            const int n = i + lane;
            const bool ok = n < count;
            const float px = ok ? sx[n] : 0.f, py = ok ? sy[n] : 0.f, pz = ok ? sz[n] : 0.f,
                        pp = ok ? sp[n] : 0.f;
#pragma unroll
            for (int j = 0; j < BQ_QPW; ++j) {
                if (cnt[j] >= nsample) continue;
                const float d2 = sqdist_expanded(qx[j], qy[j], qz[j], qq[j], px, py, pz, pp);
                const bool in = ok && !(d2 > r2);
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (m) {
                    if (first[j] == N) first[j] = base + i + (__ffs(m) - 1);
                    const int slot = cnt[j] + __popc(m & ((1u << lane) - 1u));
                    if (in && slot < nsample)
                        out[((int64_t)b * S + q0 + j) * nsample + slot] = base + n;
                    cnt[j] += __popc(m);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < BQ_QPW; ++j) {
        if (q0 + j >= S) continue;
        for (int s = cnt[j] + lane; s < nsample; s += 32)
            out[((int64_t)b * S + q0 + j) * nsample + s] = first[j];
    }
}

// ------------------------------------------------- fused set abstraction -----
// One warp per centroid: ball query (as above) + for every member the shared MLP
// [3+D] -> C1 -> C2 -> C3 (conv1x1 + folded eval-BN + ReLU) + running max.
// Padding slots of the reference repeat member 0, so they do not change the max.
// Lane l owns output channel l (C3 <= 32... C3 up to 64 uses two per lane).
struct SaParams {
    const float *W[3], *b[3], *alpha[3], *beta[3];
    int cin[3], cout[3];
    int n_layers;
};

constexpr int SA_WARPS = 16;
constexpr int SA_MAXC = 64;       // widest layer supported
constexpr int SA_MAXIN = 3 + 64;  // widest input supported

__global__ void __launch_bounds__(SA_WARPS * 32)
sa_layer_kernel(Cloud xyz, Cloud feats, int D, const int32_t *__restrict__ cidx, int N, int S, float r2,
                int nsample, SaParams P, float *__restrict__ out_feat, float *__restrict__ out_xyz) {
    extern __shared__ float smem[];
    float *sx = smem, *sy = sx + BQ_TILE, *sz = sy + BQ_TILE, *sp = sz + BQ_TILE;
    float *swt = sp + BQ_TILE;   // weights: per layer W[cout][cin], b, alpha, beta
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // stage weights
    int woff[3], boff[3];
    {
        int off = 0;
        for (int l = 0; l < P.n_layers; ++l) {
            woff[l] = off;
            off += P.cin[l] * P.cout[l];
            boff[l] = off;
            off += 3 * P.cout[l];
        }
        for (int l = 0; l < P.n_layers; ++l) {
            for (int i = threadIdx.x; i < P.cin[l] * P.cout[l]; i += blockDim.x) swt[woff[l] + i] = P.W[l][i];
            for (int i = threadIdx.x; i < P.cout[l]; i += blockDim.x) {
                swt[boff[l] + i] = P.b[l][i];
                swt[boff[l] + P.cout[l] + i] = P.alpha[l][i];
                swt[boff[l] + 2 * P.cout[l] + i] = P.beta[l][i];
            }
        }
    }
    const int s = blockIdx.x * SA_WARPS + warp;
    const bool active = s < S;
    const int c = active ? cidx[(int64_t)b * S + s] : 0;
    const float qx = xyz.at(b, c, 0), qy = xyz.at(b, c, 1), qz = xyz.at(b, c, 2);
    const float qq = norm2_nofma(qx, qy, qz);
    if (active && out_xyz && lane < 3) out_xyz[((int64_t)b * S + s) * 3 + lane] = lane == 0 ? qx : (lane == 1 ? qy : qz);
    const int clast = P.cout[P.n_layers - 1];
    float best0 = -INFINITY, best1 = -INFINITY;   // channels lane, lane+32
    int cnt = active ? 0 : nsample;

    for (int base = 0; base < N; base += BQ_TILE) {
        const int count = min(BQ_TILE, N - base);
        __syncthreads();
        stage_tile(xyz, b, base, count, sx, sy, sz, sp);
        __syncthreads();
        if (cnt >= nsample) continue;
        for (int i = 0; i < count && cnt < nsample; i += 32) {
            const int n = i + lane;
            const bool ok = n < count;
            const float px = ok ? sx[n] : 0.f, py = ok ? sy[n] : 0.f, pz = ok ? sz[n] : 0.f,
                        pp = ok ? sp[n] : 0.f;
            const float d2 = sqdist_expanded(qx, qy, qz, qq, px, py, pz, pp);
            const bool in = ok && !(d2 > r2);
            unsigned m = __ballot_sync(0xffffffffu, in);
            if (!m) continue;
            // members of this step, ascending; keep only the first nsample overall
            int take = min(__popc(m), nsample - cnt);
            cnt += __popc(m);
            while (take-- > 0) {
                const int src_lane = __ffs(m) - 1;
                m &= m - 1;
                const int pn = base + i + src_lane;
                // input row: [p - centre (3), feats (D)], every lane holds the full row
                float a_in[SA_MAXIN];
                const float mx = __shfl_sync(0xffffffffu, px, src_lane);
                const float my = __shfl_sync(0xffffffffu, py, src_lane);
                const float mz = __shfl_sync(0xffffffffu, pz, src_lane);
                a_in[0] = mx - qx;
                a_in[1] = my - qy;
                a_in[2] = mz - qz;
                for (int d = 0; d < D; ++d) a_in[3 + d] = feats.at(b, pn, d);
                // generic layers through shared scratch would be slow; layer widths are
                // small, so every lane computes up to two output channels per layer and
                // the row is re-broadcast with shuffles.
                float cur0 = 0.f, cur1 = 0.f;
                int cin = 3 + D;
                for (int l = 0; l < P.n_layers; ++l) {
                    const int co = P.cout[l];
                    const float *W = swt + woff[l];
                    const float *bb = swt + boff[l];
                    float y0 = 0.f, y1 = 0.f;
                    const int o0 = lane, o1 = lane + 32;
                    if (l == 0) {
                        if (o0 < co) {
                            float acc = 0.f;
                            for (int k = 0; k < cin; ++k) acc = fmaf(W[o0 * cin + k], a_in[k], acc);
                            y0 = acc;
                        }
                        if (o1 < co) {
                            float acc = 0.f;
                            for (int k = 0; k < cin; ++k) acc = fmaf(W[o1 * cin + k], a_in[k], acc);
                            y1 = acc;
                        }
                    } else {
                        float acc0 = 0.f, acc1 = 0.f;
                        for (int k = 0; k < cin; ++k) {
                            const float v = (k < 32) ? __shfl_sync(0xffffffffu, cur0, k)
                                                     : __shfl_sync(0xffffffffu, cur1, k - 32);
                            if (o0 < co) acc0 = fmaf(W[o0 * cin + k], v, acc0);
                            if (o1 < co) acc1 = fmaf(W[o1 * cin + k], v, acc1);
                        }
                        y0 = acc0;
                        y1 = acc1;
                    }
                    if (o0 < co) y0 = fmaxf(fmaf(y0 + bb[o0], bb[co + o0], bb[2 * co + o0]), 0.f);
                    if (o1 < co) y1 = fmaxf(fmaf(y1 + bb[o1], bb[co + o1], bb[2 * co + o1]), 0.f);
                    cur0 = y0;
                    cur1 = y1;
                    cin = co;
                }
                best0 = fmaxf(best0, cur0);
                best1 = fmaxf(best1, cur1);
            }
        }
    }
    if (active) {
        if (lane < clast) out_feat[((int64_t)b * S + s) * clast + lane] = best0;
        if (lane + 32 < clast) out_feat[((int64_t)b * S + s) * clast + lane + 32] = best1;
    }
}

// ------------------------------------------------------ square_distance -----
__global__ void square_distance_kernel(Cloud src, Cloud dst, int S, int N, float *__restrict__ out) {
    const int b = blockIdx.z, s = blockIdx.y;
    const float qx = src.at(b, s, 0), qy = src.at(b, s, 1), qz = src.at(b, s, 2);
    const float qq = norm2_nofma(qx, qy, qz);
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float px = dst.at(b, n, 0), py = dst.at(b, n, 1), pz = dst.at(b, n, 2);
        out[((int64_t)b * S + s) * N + n] = sqdist_expanded(qx, qy, qz, qq, px, py, pz, norm2_nofma(px, py, pz));
    }
}

// --------------------------------------------------------- index_points -----
__global__ void index_points_kernel(const float *__restrict__ pts, const int64_t *__restrict__ idx, int N,
                                    int C, int64_t M, float *__restrict__ out) {
    const int b = blockIdx.y;
    const int64_t total = M * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / C;
        const int c = (int)(i - m * C);
        const int64_t n = idx[(int64_t)b * M + m];
        out[(int64_t)b * total + i] = pts[((int64_t)b * N + n) * C + c];
    }
}

// ---------------------------------------------------- weighting MLP ----------
// One thread per point: 32 -> 16 (ReLU) -> 8 (ReLU) -> 1 (Softplus, beta 1,
// threshold 20 as torch.nn.Softplus). Weights staged in shared memory.
__global__ void __launch_bounds__(256)
weighting_kernel(const float *__restrict__ X, int64_t rows, const float *__restrict__ W1,
                 const float *__restrict__ b1, const float *__restrict__ W2, const float *__restrict__ b2,
                 const float *__restrict__ W3, const float *__restrict__ b3, float *__restrict__ scores) {
    __shared__ float w1[16 * 32], c1[16], w2[8 * 16], c2[8], w3[8], c3;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) w1[i] = W1[i];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) w2[i] = W2[i];
    if (threadIdx.x < 16) c1[threadIdx.x] = b1[threadIdx.x];
    if (threadIdx.x < 8) {
        c2[threadIdx.x] = b2[threadIdx.x];
        w3[threadIdx.x] = W3[threadIdx.x];
    }
    if (threadIdx.x == 0) c3 = b3[0];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float x[32];
    const float4 *xp = reinterpret_cast<const float4 *>(X + r * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 v = __ldg(xp + i);
        x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
    }
    float h1[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) acc = fmaf(w1[o * 32 + k], x[k], acc);
        h1[o] = fmaxf(acc + c1[o], 0.f);
    }
    float h2[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = fmaf(w2[o * 16 + k], h1[k], acc);
        h2[o] = fmaxf(acc + c2[o], 0.f);
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fmaf(w3[k], h2[k], acc);
    acc += c3;
    scores[r] = acc > 20.f ? acc : log1pf(expf(acc));
}

// ------------------------------------------------------------ top-K ----------
// One CTA per batch item; K rounds of block arg-max over (score bits, lowest
// index). Scores are softplus outputs (>= 0) so their bit patterns order like
// the values; general floats are mapped to an order-preserving unsigned key.
__device__ __forceinline__ unsigned ordered_key(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(1024, 1)
topk_kernel(const float *__restrict__ scores, int S, int K, int64_t *__restrict__ out) {
    extern __shared__ unsigned skey[];   // S keys; 0 = taken
    __shared__ unsigned s_hi[2][32], s_lo[2][32];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int n = tid; n < S; n += blockDim.x) {
        const float v = scores[(int64_t)b * S + n];
        skey[n] = (v != v) ? 1u : max(ordered_key(v), 1u);   // NaN last; 0 reserved
    }
    __syncthreads();
    for (int i = 0; i < K; ++i) {
        unsigned hi = 0u, lo = 0u;
        for (int n = tid; n < S; n += blockDim.x) {
            const unsigned k = skey[n];
            const unsigned l = 0xffffffffu - (unsigned)n;
            if (k > hi || (k == hi && l > lo)) {
                hi = k;
                lo = l;
            }
        }
        warp_max_pair(hi, lo);
        if (lane == 0) {
            s_hi[i & 1][warp] = hi;
            s_lo[i & 1][warp] = lo;
        }
        __syncthreads();
        hi = s_hi[i & 1][lane];
        lo = s_lo[i & 1][lane];
        warp_max_pair(hi, lo);
        const unsigned win = 0xffffffffu - lo;
        if (tid == 0) {
            out[(int64_t)b * K + i] = win;
            skey[win] = 0u;
        }
        __syncthreads();
    }
}

}  // namespace dvcp

using namespace dvcp;

extern "C" int dvcp_ball_query(dvcp_cloud_t xyz, dvcp_cloud_t new_xyz, int B, int N, int S, float radius2,
                               int nsample, int64_t *out, dvcp_stream_t stream) {
    if (!xyz.base || !new_xyz.base || !out || B <= 0 || N <= 0 || S <= 0 || nsample <= 0) return DVCP_E_ARG;
    const size_t smem = 4 * BQ_TILE * sizeof(float);
    DVCP_CUDA(cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((S + BQ_WARPS * BQ_QPW - 1) / (BQ_WARPS * BQ_QPW), B);
    ball_query_kernel<<<grid, BQ_WARPS * 32, smem, (cudaStream_t)stream>>>(as_cloud(xyz), as_cloud(new_xyz), N, S,
                                                                          radius2, nsample, out);
    DVCP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dvcp_sa_layer(dvcp_cloud_t xyz, dvcp_cloud_t feats, int D, const int32_t *centroid_idx, int B,
                             int N, int S, float radius2, int nsample, const dvcp_mlp_layer_t *layers,
                             int n_layers, float *out_feat, float *out_xyz, dvcp_stream_t stream) {
    if (!xyz.base || !centroid_idx || !layers || !out_feat || B <= 0 || N <= 0 || S <= 0 || nsample <= 0)
        return DVCP_E_ARG;
    if (D < 0 || (D > 0 && !feats.base)) return DVCP_E_ARG;
    if (n_layers < 1 || n_layers > 3 || 3 + D > SA_MAXIN) return DVCP_E_UNSUPPORTED;
    SaParams P;
    P.n_layers = n_layers;
    int cin = 3 + D;
    size_t wfloats = 0;
    for (int l = 0; l < n_layers; ++l) {
        if (layers[l].in_ch != cin || layers[l].out_ch < 1 || layers[l].out_ch > SA_MAXC) return DVCP_E_UNSUPPORTED;
        P.W[l] = layers[l].W; P.b[l] = layers[l].b; P.alpha[l] = layers[l].alpha; P.beta[l] = layers[l].beta;
        if (!P.W[l] || !P.b[l] || !P.alpha[l] || !P.beta[l]) return DVCP_E_ARG;
        P.cin[l] = cin; P.cout[l] = layers[l].out_ch;
        wfloats += (size_t)cin * layers[l].out_ch + 3 * layers[l].out_ch;
        cin = layers[l].out_ch;
    }
    for (int l = n_layers; l < 3; ++l) { P.W[l] = P.b[l] = P.alpha[l] = P.beta[l] = nullptr; P.cin[l] = P.cout[l] = 0; }
    Cloud f = D > 0 ? as_cloud(feats) : Cloud{nullptr, 0, 0, 0};
    const size_t smem = (4 * BQ_TILE + wfloats) * sizeof(float);
    DVCP_CUDA(cudaFuncSetAttribute(sa_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((S + SA_WARPS - 1) / SA_WARPS, B);
    sa_layer_kernel<<<grid, SA_WARPS * 32, smem, (cudaStream_t)stream>>>(as_cloud(xyz), f, D, centroid_idx, N, S,
                                                                        radius2, nsample, P, out_feat, out_xyz);
    DVCP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dvcp_square_distance(dvcp_cloud_t src, dvcp_cloud_t dst, int B, int S, int N, float *out,
                                    dvcp_stream_t stream) {
    if (!src.base || !dst.base || !out || B <= 0 || S <= 0 || N <= 0) return DVCP_E_ARG;
    if (S > 65535 || B > 65535) return DVCP_E_UNSUPPORTED;
    dim3 grid(min((N + 255) / 256, 1024), S, B);
    square_distance_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(as_cloud(src), as_cloud(dst), S, N, out);
    DVCP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dvcp_index_points(const float *points, const int64_t *idx, int B, int N, int C, int64_t M,
                                 float *out, dvcp_stream_t stream) {
    if (!points || !idx || !out || B <= 0 || N <= 0 || C <= 0 || M <= 0) return DVCP_E_ARG;
    if (B > 65535) return DVCP_E_UNSUPPORTED;
    const int64_t total = M * C;
    int64_t gx = (total + 255) / 256; if (gx > 148 * 16) gx = 148 * 16;
    dim3 grid((unsigned)gx, B);
    index_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(points, idx, N, C, M, out);
    DVCP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dvcp_weighting_scores(const float *X, int B, int S, const float *W1, const float *b1,
                                     const float *W2, const float *b2, const float *W3, const float *b3,
                                     float *scores, dvcp_stream_t stream) {
    if (!X || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !scores || B <= 0 || S <= 0) return DVCP_E_ARG;
    const int64_t rows = (int64_t)B * S;
    weighting_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, rows, W1, b1, W2, b2, W3,
                                                                                    b3, scores);
    DVCP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dvcp_topk(const float *scores, int B, int S, int K, int64_t *topk, dvcp_stream_t stream) {
    if (!scores || !topk || B <= 0 || S <= 0 || K <= 0 || K > S) return DVCP_E_ARG;
    const size_t smem = (size_t)S * sizeof(unsigned);
    if (smem > 200 * 1024) return DVCP_E_UNSUPPORTED;
    DVCP_CUDA(cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_kernel<<<B, 1024, smem, (cudaStream_t)stream>>>(scores, S, K, topk);
    DVCP_CHECK_LAUNCH();
    return 0;
}
