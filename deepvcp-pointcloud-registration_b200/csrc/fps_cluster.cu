// Farthest point sampling of ONE cloud by a thread-block CLUSTER of 8 CTAs
// (pointnet2_utils.py:63-84; same exact results as fps.cu, see the semantics there).
//
// Why a cluster: with B = 8 pairs there are only 16 clouds for 148 SMs, and FPS is a
// chain of N dependent selections. The batched-round scheme of fps.cu (many exact
// selections per block-wide step) leaves the per-step bucket updates as the bulk of
// the work; this kernel spreads them over 8 SMs and keeps ONE cluster barrier per
// step:
//   * the cloud arrives Morton-sorted with bucket boxes (dvcp_build_index); bucket j
//     belongs to CTA j % 8, so every centroid's neighbourhood is spread over the CTAs;
//   * per step each CTA (1) applies the centroids accepted in the previous step to
//     the buckets they can reach, (2) finds its buckets' best keys above its own
//     largest second-best key, (3) pushes them, with coordinates, into the shared
//     memory of all 8 CTAs (DSMEM stores), (4) cluster barrier, (5) every CTA
//     resolves the same candidate list redundantly (so no second exchange is needed).
// Exactness argument: identical to the batched rounds of fps.cu (candidates = ALL
// points above S = the largest second-best key of any bucket; accepted in key order
// while no accepted centroid can lower them).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dvcp {

constexpr int FC_C = 8;         // CTAs per cluster = per cloud
constexpr int FC_WARPS = 8;
constexpr int FC_THREADS = FC_WARPS * 32;
constexpr int FC_MAXNBL = 64;   // buckets per CTA (16384 points / 32 / 8)

struct FcShared {
    // exchange area, double-buffered by step parity; written by the peers
    unsigned long long r_key[2][FC_C][FC_MAXNBL];
    float4 r_xyz[2][FC_C][FC_MAXNBL];
    unsigned long long r_S[2][FC_C];
    unsigned r_cnt[2][FC_C];
    // local
    unsigned long long l_key[FC_MAXNBL];
    float4 l_xyz[FC_MAXNBL];
    unsigned long long c_key[FC_C * FC_MAXNBL];
    float4 c_xyz[FC_C * FC_MAXNBL];
    unsigned short c_top[32];
    unsigned long long T;
    float4 acc[32];
    unsigned long long best[FC_MAXNBL], sec[FC_MAXNBL];
    unsigned F[FC_MAXNBL];
    float box[6][FC_MAXNBL];
    unsigned l_cnt, n_cand, n_acc;
    unsigned long long l_S;
};

__device__ __forceinline__ void fc_top2(unsigned hi0, unsigned lo0, unsigned long long &best, unsigned long long &sec) {
    unsigned hi = hi0, lo = lo0;
    warp_max_pair(hi, lo);
    const bool mine = hi0 == hi && lo0 == lo;
    unsigned h2 = mine ? 0u : hi0, l2 = mine ? 0u : lo0;
    warp_max_pair(h2, l2);
    best = ((unsigned long long)hi << 32) | lo;
    sec = ((unsigned long long)h2 << 32) | l2;
}

__global__ void __launch_bounds__(FC_THREADS, 1)
fps_cluster_kernel(Cloud xyz, dvcp_cloud_index_t index, int N, int npoint, const int64_t *__restrict__ start,
                   int64_t *__restrict__ out64, int32_t *__restrict__ out32) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / FC_C;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cap = index.cap, NB = cap / 32, NBL = NB / FC_C, H2 = (NBL + 31) / 32;
    FcShared &sh = *reinterpret_cast<FcShared *>(smem_raw);
    float *s_x = reinterpret_cast<float *>(smem_raw + sizeof(FcShared));
    float *s_y = s_x + NBL * 32, *s_z = s_y + NBL * 32, *s_d = s_z + NBL * 32;
    unsigned short *s_id = reinterpret_cast<unsigned short *>(s_d + NBL * 32);

    // ---- load my buckets (bucket jl of this CTA = global bucket jl * 8 + rank) ----
    const float *gx = index.sorted_xyz + (int64_t)b * 3 * cap;
    const int32_t *gi = index.sorted_idx + (int64_t)b * cap;
    const float *gbox = index.bucket_box + (int64_t)b * NB * 8;
    for (int jl = warp; jl < NBL; jl += FC_WARPS) {
        const int jg = jl * FC_C + rank, gp = jg * 32 + lane, p = jl * 32 + lane;
        const int id = __ldg(gi + gp);
        const float x = __ldg(gx + gp), y = __ldg(gx + cap + gp), z = __ldg(gx + 2 * cap + gp);
        s_x[p] = x;
        s_y[p] = y;
        s_z[p] = z;
        const unsigned idu = id < 0 ? 0xffffu : (unsigned)id;
        s_id[p] = (unsigned short)idu;
        const float d0 = id < 0 ? 0.0f : 1e10f;
        s_d[p] = d0;
        unsigned long long bk, sk;
        fc_top2(__float_as_uint(d0), ((0xffffu - idu) << 16) | (unsigned)gp, bk, sk);
        if (lane == 0) {
            sh.best[jl] = bk;
            sh.sec[jl] = sk;
            sh.F[jl] = 0u;
        }
        if (lane < 6) {
            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(gbox + (int64_t)jg * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4 *>(gbox + (int64_t)jg * 8) + 1);
            const float v = lane == 0 ? b0.x : lane == 1 ? b0.y : lane == 2 ? b0.z : lane == 3 ? b0.w : lane == 4 ? b1.x : b1.y;
            // an empty bucket gets an unreachable box
            sh.box[lane][jl] = b1.z > 0.f ? v : (lane < 3 ? INFINITY : -INFINITY);
        }
    }
    const unsigned startidx = (unsigned)start[b];
    if (tid == 0) {
        sh.acc[0] = make_float4(xyz.at(b, (int)startidx, 0), xyz.at(b, (int)startidx, 1), xyz.at(b, (int)startidx, 2), 0.f);
        sh.n_acc = 1u;
        sh.n_cand = 0u;
        if (rank == 0) {
            if (out64) out64[(int64_t)b * npoint] = startidx;
            if (out32) out32[(int64_t)b * npoint] = (int32_t)startidx;
        }
    }
    cluster.sync();   // peers exist and are initialised before anyone writes into them

    int produced = 0;
    for (int step = 0;; ++step) {
        const int buf = step & 1;
        // ---- (1) which of my buckets can each accepted centroid reach? ----
        const int A = (int)sh.n_acc;
        for (int item = warp; item < A * H2; item += FC_WARPS) {
            const int a = item / H2, jl = (item - a * H2) * 32 + lane;
            if (jl < NBL) {
                const float4 c = sh.acc[a];
                const float ex = fmaxf(fmaxf(__fsub_rn(sh.box[0][jl], c.x), __fsub_rn(c.x, sh.box[3][jl])), 0.0f);
                const float ey = fmaxf(fmaxf(__fsub_rn(sh.box[1][jl], c.y), __fsub_rn(c.y, sh.box[4][jl])), 0.0f);
                const float ez = fmaxf(fmaxf(__fsub_rn(sh.box[2][jl], c.z), __fsub_rn(c.z, sh.box[5][jl])), 0.0f);
                const float bestval = __uint_as_float((unsigned)(sh.best[jl] >> 32));
                if (sq3_nofma(ex, ey, ez) < bestval) atomicOr(&sh.F[jl], 1u << a);
            }
        }
        __syncthreads();
        // ---- (2) lower the distances of the reached buckets, refresh their two best keys ----
        for (int jl = warp; jl < NBL; jl += FC_WARPS) {
            unsigned F = sh.F[jl];
            if (F) {
                const int p = jl * 32 + lane;
                const float x = s_x[p], y = s_y[p], z = s_z[p];
                float dk = s_d[p];
                do {
                    const int a = __ffs(F) - 1;
                    F &= F - 1;
                    const float4 c = sh.acc[a];
                    const float d = sq3_nofma(__fsub_rn(x, c.x), __fsub_rn(y, c.y), __fsub_rn(z, c.z));
                    dk = d < dk ? d : dk;
                } while (F);
                s_d[p] = dk;
                unsigned long long bk, sk;
                fc_top2(__float_as_uint(dk), ((0xffffu - (unsigned)s_id[p]) << 16) | (unsigned)((jl * FC_C + rank) * 32 + lane),
                        bk, sk);
                if (lane == 0) {
                    sh.best[jl] = bk;
                    sh.sec[jl] = sk;
                    sh.F[jl] = 0u;
                }
            }
        }
        produced += A;
        if (produced >= npoint) break;
        __syncthreads();
        // ---- (3) my candidates: bucket maxima above my largest second-best key ----
        if (warp == 0) {
            unsigned shi = 0u, slo = 0u;
            for (int h = 0; h < H2; ++h) {
                const int jl = h * 32 + lane;
                const unsigned long long s = jl < NBL ? sh.sec[jl] : 0ull;
                const unsigned hi = (unsigned)(s >> 32), lo = (unsigned)s;
                if (hi > shi || (hi == shi && lo > slo)) {
                    shi = hi;
                    slo = lo;
                }
            }
            warp_max_pair(shi, slo);
            const unsigned long long S = ((unsigned long long)shi << 32) | slo;
            unsigned cnt = 0u;
            for (int h = 0; h < H2; ++h) {
                const int jl = h * 32 + lane;
                const unsigned long long key = jl < NBL ? sh.best[jl] : 0ull;
                const bool isc = (key >> 32) != 0ull && key > S;
                const unsigned m = __ballot_sync(0xffffffffu, isc);
                if (isc) {
                    const unsigned slot = cnt + __popc(m & ((1u << lane) - 1u));
                    const int p = jl * 32 + (int)(key & 31u);
                    sh.l_key[slot] = key;
                    sh.l_xyz[slot] = make_float4(s_x[p], s_y[p], s_z[p], 0.f);
                }
                cnt += __popc(m);
            }
            if (lane == 0) {
                sh.l_cnt = cnt;
                sh.l_S = S;
            }
        }
        __syncthreads();
        // ---- (4) push them into every CTA of the cluster (warp w -> CTA w), one barrier ----
        {
            FcShared *peer = cluster.map_shared_rank(&sh, warp);
            const unsigned cnt = sh.l_cnt;
            for (unsigned e = lane; e < cnt; e += 32) {
                peer->r_key[buf][rank][e] = sh.l_key[e];
                peer->r_xyz[buf][rank][e] = sh.l_xyz[e];
            }
            if (lane == 0) {
                peer->r_cnt[buf][rank] = cnt;
                peer->r_S[buf][rank] = sh.l_S;
            }
        }
        cluster.sync();
        // ---- (5) every CTA resolves the same list: S, candidates above S ----
        unsigned long long S = 0ull;
#pragma unroll
        for (int c = 0; c < FC_C; ++c) {
            const unsigned long long v = sh.r_S[buf][c];
            S = v > S ? v : S;
        }
        {
            const int c = warp;   // one warp per source CTA
            const unsigned cnt = sh.r_cnt[buf][c];
            for (unsigned e = lane; e < cnt; e += 32) {
                const unsigned long long key = sh.r_key[buf][c][e];
                if (key > S) {
                    const unsigned slot = atomicAdd(&sh.n_cand, 1u);
                    sh.c_key[slot] = key;
                    sh.c_xyz[slot] = sh.r_xyz[buf][c][e];
                }
            }
        }
        __syncthreads();
        const int n = (int)sh.n_cand;
        if (n == 0) {
            // every remaining distance is 0: the argmax stays the lowest index (pointnet2_utils.py:83),
            // which is the largest best key anywhere. Every CTA knows only its own: take it from global order.
            // All best keys have hi == 0 here, so the winner is the largest lo = lowest original index overall.
            if (warp == 0) {
                unsigned hi = 0u, lo = 0u;
                for (int h = 0; h < H2; ++h) {
                    const int jl = h * 32 + lane;
                    const unsigned long long k2 = jl < NBL ? sh.best[jl] : 0ull;
                    if ((unsigned)k2 > lo) lo = (unsigned)k2;
                }
                warp_max_pair(hi, lo);
                if (lane == 0) {
                    FcShared *leader = cluster.map_shared_rank(&sh, 0);
                    leader->r_S[buf ^ 1][rank] = (unsigned long long)lo;   // the other buffer is idle now
                }
            }
            cluster.sync();
            if (rank == 0) {
                unsigned long long m = 0ull;
                for (int c = 0; c < FC_C; ++c) m = sh.r_S[buf ^ 1][c] > m ? sh.r_S[buf ^ 1][c] : m;
                const unsigned idx = 0xffffu - ((unsigned)m >> 16);
                for (int i = produced + tid; i < npoint; i += FC_THREADS) {
                    if (out64) out64[(int64_t)b * npoint + i] = idx;
                    if (out32) out32[(int64_t)b * npoint + i] = (int32_t)idx;
                }
            }
            break;
        }
        // more than 32 candidates (rare): keep the 32 largest, T = the 33rd
        if (n > 32) {
            for (int t = tid; t < n; t += FC_THREADS) {
                const unsigned long long key = sh.c_key[t];
                int rk = 0;
                for (int u = 0; u < n; ++u) rk += sh.c_key[u] > key;
                if (rk < 32) sh.c_top[rk] = (unsigned short)t;
                if (rk == 32) sh.T = key;
            }
            __syncthreads();
        }
        if (warp == 0) {
            const int m = n < 32 ? n : 32;
            const unsigned long long T = n > 32 ? sh.T : S;
            const int src = lane < m ? (n > 32 ? (int)sh.c_top[lane] : lane) : 0;
            const unsigned long long key = lane < m ? sh.c_key[src] : 0ull;
            const float4 q = sh.c_xyz[src];
            const unsigned khi = (unsigned)(key >> 32), klo = (unsigned)(key & 0xffffffffu);
            const float dj = __uint_as_float(khi);
            unsigned H = 0u, K = 0u, L = 0u;
            for (int i = 0; i < m; ++i) {
                const float xi = __shfl_sync(0xffffffffu, q.x, i), yi = __shfl_sync(0xffffffffu, q.y, i),
                            zi = __shfl_sync(0xffffffffu, q.z, i);
                const unsigned hi_i = __shfl_sync(0xffffffffu, khi, i), lo_i = __shfl_sync(0xffffffffu, klo, i);
                const bool before = hi_i > khi || (hi_i == khi && lo_i > klo);
                const float d = sq3_nofma(__fsub_rn(q.x, xi), __fsub_rn(q.y, yi), __fsub_rn(q.z, zi));
                const bool kill = before && d < dj;
                const unsigned long long nk = ((unsigned long long)__float_as_uint(d) << 32) | klo;
                H |= (unsigned)before << i;
                K |= (unsigned)kill << i;
                L |= (unsigned)(kill && nk <= T) << i;
            }
            const unsigned validm = m == 32 ? 0xffffffffu : ((1u << m) - 1u);
            unsigned acc = 0u, rej = ~validm;
            int state = lane < m ? 0 : 2;
            while ((acc | rej) != 0xffffffffu) {
                if (state == 0) {
                    if (K & acc) state = 2;
                    else if ((K & ~rej) == 0u) state = 1;
                }
                acc = __ballot_sync(0xffffffffu, state == 1);
                rej = __ballot_sync(0xffffffffu, state == 2);
            }
            const bool stopper = lane < m && state == 2 && (L & acc) == 0u;
            const unsigned stopm = __ballot_sync(0xffffffffu, stopper);
            if (stopm) {
                const int fs = __ffs(__ballot_sync(0xffffffffu, stopper && (H & stopm) == 0u)) - 1;
                acc &= __shfl_sync(0xffffffffu, H, fs);
            }
            const int rk = __popc(acc & H);
            const int rem = npoint - produced;
            if (((acc >> lane) & 1u) && rk < rem) {
                sh.acc[rk] = make_float4(q.x, q.y, q.z, 0.f);
                if (rank == 0) {
                    const unsigned idx = 0xffffu - (klo >> 16);
                    if (out64) out64[(int64_t)b * npoint + produced + rk] = idx;
                    if (out32) out32[(int64_t)b * npoint + produced + rk] = (int32_t)idx;
                }
            }
            if (lane == 0) {
                const int na = __popc(acc);
                sh.n_acc = (unsigned)(na < rem ? na : rem);
                sh.n_cand = 0u;
            }
        }
        __syncthreads();
    }
    cluster.sync();   // nobody leaves while a peer may still address its shared memory
}

}  // namespace dvcp

// Launcher used by dvcp_fps (fps.cu). The index must already hold the cloud (dvcp_build_index).
int dvcp_fps_cluster_launch(dvcp_cloud_t xyz, dvcp_cloud_index_t index, int B, int N, int npoint, const int64_t *start,
                            int64_t *out64, int32_t *out32, cudaStream_t st) {
    using namespace dvcp;
    const int NB = index.cap / 32;
    if (NB % FC_C != 0 || NB / FC_C > FC_MAXNBL || NB / FC_C < 1) return DVCP_E_UNSUPPORTED;
    const int NBL = NB / FC_C;
    const size_t smem = sizeof(FcShared) + (size_t)NBL * 32 * (4 * sizeof(float) + sizeof(unsigned short));
    DVCP_CUDA(cudaFuncSetAttribute(fps_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * FC_C));
    cfg.blockDim = dim3(FC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = FC_C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DVCP_CUDA(cudaLaunchKernelEx(&cfg, fps_cluster_kernel, as_cloud(xyz), index, N, npoint, start, out64, out32));
    return 0;
}
