// Candidate-grid generation and exact K-nearest-neighbour search.
//
// Reference: voxelize.py:19-83 (candidates) and the third-party knn_cuda.KNN
// call sites get_cat_feat_tgt.py:45,52 / deepVCP_loss.py:70,72.
//
// KNN contract: float32; d2 = fma(dz,dz, fma(dy,dy, dx*dx)) with d = ref - query;
// neighbours ordered by (d2, index) ascending; sqrt_rn(d2) returned; indices
// 0-based. The order is made total by the 64-bit key (bits(d2) << 32 | index), so
// the result does not depend on the order points are visited in.
//
// Brute-force kernel (any N): the reference cloud is staged tile by tile into shared memory as
// x[], y[], z[] (coalesced loads, conflict-free LDS); each warp owns KNN_QPW
// queries and keeps each query's current K best as ONE key per lane, sorted
// across the lanes. A step evaluates 32 points; lanes whose key beats the current
// K-th are inserted with ballot + shuffle-up.
#include <stdlib.h>

#include "common.cuh"

namespace dvcp {

// ------------------------------------------------------------ candidates ----
// value = float32( ((c - r) - s/2) + s * i ), all in float64, no contraction.
__global__ void candidates_kernel(const double *__restrict__ centres, int64_t M, double r, double s, int G,
                                  float *__restrict__ out) {
    const int64_t C = (int64_t)G * G * G;
    const int64_t total = M * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = t / C;
        const int c = (int)(t - m * C);
        const int iz = c % G, iy = (c / G) % G, ix = c / (G * G);
        const double half = s / 2;
        const double sx = __dsub_rn(__dsub_rn(centres[3 * m], r), half);
        const double sy = __dsub_rn(__dsub_rn(centres[3 * m + 1], r), half);
        const double sz = __dsub_rn(__dsub_rn(centres[3 * m + 2], r), half);
        float *o = out + 3 * t;
        o[0] = (float)__dadd_rn(sx, __dmul_rn(s, (double)ix));
        o[1] = (float)__dadd_rn(sy, __dmul_rn(s, (double)iy));
        o[2] = (float)__dadd_rn(sz, __dmul_rn(s, (double)iz));
    }
}

// ------------------------------------------------------------------- KNN ----
constexpr int KNN_TILE = 8192;
constexpr int KNN_WARPS = 16;
constexpr int KNN_QPW = 4;

__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_kernel(Cloud ref, const float *__restrict__ query, int N, int64_t Q, int K, float *__restrict__ dist,
           int64_t *__restrict__ idx64, int32_t *__restrict__ idx32) {
    extern __shared__ float smem[];
    float *sx = smem, *sy = sx + KNN_TILE, *sz = sy + KNN_TILE;
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t q0 = ((int64_t)blockIdx.x * KNN_WARPS + warp) * KNN_QPW;
    const unsigned long long INF = 0xffffffffffffffffull;
    float qx[KNN_QPW], qy[KNN_QPW], qz[KNN_QPW];
    unsigned long long list[KNN_QPW], worst[KNN_QPW];
#pragma unroll
    for (int j = 0; j < KNN_QPW; ++j) {
        const int64_t q = min(q0 + j, Q - 1);
        const float *qp = query + ((int64_t)b * Q + q) * 3;
        qx[j] = __ldg(qp);
        qy[j] = __ldg(qp + 1);
        qz[j] = __ldg(qp + 2);
        list[j] = INF;
        worst[j] = INF;
    }
    for (int base = 0; base < N; base += KNN_TILE) {
        const int count = min(KNN_TILE, N - base);
        __syncthreads();
        for (int i = threadIdx.x; i < count; i += blockDim.x) {
            sx[i] = ref.at(b, base + i, 0);
            sy[i] = ref.at(b, base + i, 1);
            sz[i] = ref.at(b, base + i, 2);
        }
        __syncthreads();
        if (q0 >= Q) continue;
        for (int i = 0; i < count; i += 32) {
            const int n = i + lane;
            const bool ok = n < count;
            const float px = ok ? sx[n] : 0.f, py = ok ? sy[n] : 0.f, pz = ok ? sz[n] : 0.f;
#pragma unroll
            for (int j = 0; j < KNN_QPW; ++j) {
                const float d2 = sqdist_direct(px - qx[j], py - qy[j], pz - qz[j]);
                const unsigned long long key =
                    ok ? (((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)(base + n)) : INF;
                unsigned m = __ballot_sync(0xffffffffu, key < worst[j]);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned long long c = __shfl_sync(0xffffffffu, key, src);
                    if (c < worst[j]) {
                        const int pos = __popc(__ballot_sync(0xffffffffu, list[j] < c));
                        const unsigned long long up = __shfl_up_sync(0xffffffffu, list[j], 1);
                        if (lane < K) list[j] = lane > pos ? up : (lane == pos ? c : list[j]);
                        worst[j] = __shfl_sync(0xffffffffu, list[j], K - 1);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < KNN_QPW; ++j) {
        if (q0 + j >= Q || lane >= K) continue;
        const int64_t o = ((int64_t)b * Q + q0 + j) * K + lane;
        dist[o] = __fsqrt_rn(__uint_as_float((unsigned)(list[j] >> 32)));
        const unsigned id = (unsigned)(list[j] & 0xffffffffu);
        if (idx64) idx64[o] = id;
        if (idx32) idx32[o] = (int32_t)id;
    }
}

}  // namespace dvcp
#include "knn_index.cuh"
namespace dvcp {

template <int T, bool BIG>   // T = buckets / 32
__global__ void __launch_bounds__(KNI_WARPS * 32)
knn_indexed_kernel(dvcp_cloud_index_t index, const float *__restrict__ query, int64_t Q, int K, int chain,
                   int zline, float loose, float *__restrict__ dist, int64_t *__restrict__ idx64,
                   int32_t *__restrict__ idx32) {
    constexpr int TT = T < 32 ? T : 32;
    constexpr int SB = (T + 31) / 32;
    __shared__ unsigned long long s_buf[KNI_WARPS][KNI_BUF];
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long INF = 0xffffffffffffffffull;
    const int cap = index.cap;
    KnnCtx ctx;
    ctx.box = index.bucket_box + (int64_t)b * (cap / 32) * 8;
    ctx.spt = reinterpret_cast<const float4 *>(index.sorted_pt) + (int64_t)b * cap;
    ctx.buf = s_buf[warp];
    ctx.spt_lane = ctx.spt + lane;
    ctx.box_lane = reinterpret_cast<const float4 *>(ctx.box) + lane * 2;
    // keep the two bases in registers: rematerialised from the kernel parameters they cost five address
    // instructions and a constant load per bucket visit instead of one IMAD.WIDE
    asm volatile("" : "+l"(ctx.spt_lane), "+l"(ctx.box_lane));
    ctx.K = K;
    ctx.loose = loose;
    Box6 sb[SB];
#pragma unroll
    for (int s = 0; s < SB; ++s) sb[s] = load_super_box(ctx.box, TT, s * 32 + lane);
    const int64_t nchains = (Q + chain - 1) / chain;
    for (int64_t ch = (int64_t)blockIdx.x * KNI_WARPS + warp; ch < nchains; ch += (int64_t)gridDim.x * KNI_WARPS) {
        const int64_t q0 = ch * chain, q1 = min(q0 + chain, Q);
        unsigned long long list = INF;   // result of the previous query of the chain (lane j: j-th neighbour)
        // boustrophedon over the z-lines of the chain (consecutive queries stay adjacent): line / position kept as
        // counters (a division per query was 3 % of the kernel's instructions)
        int line = 0, k = 0;
        int64_t line0 = q0;                       // first query of the current line
        bool rev = false;                         // odd, complete line: walked backwards
        for (int64_t qi = q0; qi < q1; ++qi) {
            const int64_t q = rev ? line0 + (zline - 1 - k) : line0 + k;
            if (++k == zline) {
                k = 0;
                ++line;
                line0 += zline;
                rev = (line & 1) && line0 + zline <= q1;
            }
            const float *qp = query + ((int64_t)b * Q + q) * 3;
            const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
            const unsigned long long res = knn_query<T, BIG>(ctx, sb, qx, qy, qz, list, lane);
            if (lane < K) {
                const int64_t o = ((int64_t)b * Q + q) * K + lane;
                dist[o] = __fsqrt_rn(__uint_as_float(KnnKey<BIG>::d2bits(res)));
                const unsigned id = KnnKey<BIG>::id(res);
                if (idx64) idx64[o] = id;
                if (idx32) idx32[o] = (int32_t)id;
            }
            list = res;
        }
    }
}

template <int T, bool BIG = false>
static int launch_knn_indexed(dvcp_cloud_index_t index, const float *query, int B, int64_t Q, int K, int chain,
                              int zline, float *dist, int64_t *idx64, int32_t *idx32, cudaStream_t st) {
    const int64_t nchains = (Q + chain - 1) / chain;
    int64_t gx = (nchains + KNI_WARPS - 1) / KNI_WARPS;
    // (capping the grid lower, i.e. fewer resident search warps beside the sampling of the next batch, lengthens the
    //  pipeline's period: 4.21 / 4.30 / 4.45 ms for the full / half / quarter grid)
    const int64_t cap = (int64_t)DVCP_NUM_SMS * 64 / (B < 64 ? B : 64) + 1;
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, B);
    static const float loose = [] {   // DVCP_KNN_LOOSE: development override of the bound-quality switch
        const char *e = getenv("DVCP_KNN_LOOSE");
        return e ? (float)atof(e) : KNI_LOOSE;
    }();
    knn_indexed_kernel<T, BIG><<<grid, KNI_WARPS * 32, 0, st>>>(index, query, Q, K, chain, zline, loose, dist, idx64, idx32);
    DVCP_CHECK_LAUNCH();
    return 0;
}

}  // namespace dvcp

using namespace dvcp;

extern "C" int dvcp_grid_size(double r, double s) {
    if (!(r > 0) || !(s > 0)) return DVCP_E_ARG;
    const double start = (0.0 - r) - s / 2;
    const double g = ceil(((0.0 + r) - start) / s);
    return (g < 1 || g > 64) ? DVCP_E_UNSUPPORTED : (int)g;
}

extern "C" int dvcp_candidates(const double *centres, int64_t M, double r, double s, int G, float *out,
                               dvcp_stream_t stream) {
    if (!centres || !out || M <= 0 || G <= 0 || !(s > 0)) return DVCP_E_ARG;
    const int64_t total = M * G * G * G;
    int64_t blocks = (total + 255) / 256;
    if (blocks > DVCP_NUM_SMS * 8) blocks = DVCP_NUM_SMS * 8;
    candidates_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(centres, M, r, s, G, out);
    DVCP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dvcp_knn_indexed(dvcp_cloud_index_t index, const float *query, int B, int N, int64_t Q, int K,
                                int chain, int zline, float *dist, int64_t *idx64, int32_t *idx32,
                                dvcp_stream_t stream) {
    if (!index.sorted_pt || !index.bucket_box || !query || !dist || (!idx64 && !idx32) ||
        B <= 0 || N <= 0 || Q <= 0 || chain < 1)
        return DVCP_E_ARG;
    if (zline < 1 || zline > chain) zline = chain;
    if (K < 1 || K > 32 || K > N || B > 65535 || index.cap < N) return DVCP_E_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    switch (index.cap / 1024) {
        case 1: return launch_knn_indexed<1>(index, query, B, Q, K, chain, zline, dist, idx64, idx32, st);
        case 2: return launch_knn_indexed<2>(index, query, B, Q, K, chain, zline, dist, idx64, idx32, st);
        case 4: return launch_knn_indexed<4>(index, query, B, Q, K, chain, zline, dist, idx64, idx32, st);
        case 8: return launch_knn_indexed<8>(index, query, B, Q, K, chain, zline, dist, idx64, idx32, st);
        case 16: return launch_knn_indexed<16>(index, query, B, Q, K, chain, zline, dist, idx64, idx32, st);
        case 32: return launch_knn_indexed<32>(index, query, B, Q, K, chain, zline, dist, idx64, idx32, st);
        case 64: return launch_knn_indexed<64>(index, query, B, Q, K, chain, zline, dist, idx64, idx32, st);
        case 128: return launch_knn_indexed<128, true>(index, query, B, Q, K, chain, zline, dist, idx64, idx32, st);
    }
    return DVCP_E_UNSUPPORTED;
}

extern "C" int dvcp_knn(dvcp_cloud_t ref, const float *query, int B, int N, int64_t Q, int K, float *dist,
                        int64_t *idx64, int32_t *idx32, dvcp_stream_t stream) {
    if (!ref.base || !query || !dist || (!idx64 && !idx32) || B <= 0 || N <= 0 || Q <= 0) return DVCP_E_ARG;
    if (K < 1 || K > 32 || K > N) return DVCP_E_UNSUPPORTED;
    if (B > 65535) return DVCP_E_UNSUPPORTED;
    const size_t smem = 3 * KNN_TILE * sizeof(float);
    DVCP_CUDA(cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t per = KNN_WARPS * KNN_QPW;
    dim3 grid((unsigned)((Q + per - 1) / per), B);
    knn_kernel<<<grid, KNN_WARPS * 32, smem, (cudaStream_t)stream>>>(as_cloud(ref), query, N, Q, K, dist, idx64, idx32);
    DVCP_CHECK_LAUNCH();
    return 0;
}
