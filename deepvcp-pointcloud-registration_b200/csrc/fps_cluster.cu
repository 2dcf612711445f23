// Farthest point sampling of ONE cloud by a thread-block CLUSTER of 8 CTAs
// (pointnet2_utils.py:63-84; same exact results as fps.cu, see the semantics there).
//
// Why a cluster: with B = 8 pairs there are only 16 clouds for 148 SMs, and FPS is a
// chain of N dependent selections. The batched-round scheme of fps.cu (many exact
// selections per block-wide step) leaves the per-step bucket updates as the bulk of
// the work; this kernel spreads them over 8 SMs and keeps ONE cluster barrier per
// step:
//   * the cloud arrives Hilbert-sorted with bucket boxes (dvcp_build_index); bucket j
//     belongs to CTA j % 8, so every centroid's neighbourhood is spread over the CTAs;
//   * per step each CTA (1) applies the centroids accepted in the previous step to
//     the buckets they can reach, (2) finds its buckets' best keys above its own
//     largest second-best key, (3) pushes them, with coordinates, into the shared
//     memory of all 8 CTAs (DSMEM stores), (4) cluster barrier, (5) every CTA
//     resolves the same candidate list redundantly (so no second exchange is needed).
//     Which buckets a centroid can reach is found AHEAD of its acceptance: while warp 0
//     resolves the candidate list, the other warps box-test EVERY candidate against the
//     CTA's buckets (the bucket maxima do not change in between); the next step masks
//     the result with the accepted set. The box tests leave the critical path.
// Exactness argument: as for the batched rounds of fps.cu, generalised to FC_E exposed
// keys per bucket: with S = the largest (FC_E+1)-th best key of any bucket, every point
// whose key exceeds S is among the FC_E best of its bucket, so the exposed keys above S
// are ALL the points above S; they are accepted in key order while no centroid accepted
// in the same step can lower them.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dvcp {

#ifdef DVCP_FPS_TIMING
__device__ long long g_fc_time[24];
__device__ int g_fc_smid[1024];
#define FC_TICK(slot)                                                   \
    do {                                                                \
        if (blockIdx.x == 0 && tid == 0) {                              \
            const long long now__ = clock64();                          \
            g_fc_time[slot] += now__ - t_last__;                        \
            t_last__ = now__;                                           \
        }                                                               \
    } while (0)
#else
#define FC_TICK(slot)
#endif

constexpr int FC_C = 8;         // CTAs per cluster = per cloud
// warps per CTA is a template parameter W (>= FC_C: warp w < FC_C pushes to CTA w): 16 is fastest for a kernel
// running alone; 8 halves the CTA's registers / threads so that other kernels keep most of each SM
constexpr int FC_MAXNBL = 64;   // buckets per CTA (16384 points / 32 / 8)
constexpr int FC_E = 3;         // keys a bucket exposes per step (its FC_E best)
constexpr int FC_CAP = 128;     // candidates resolved per step
constexpr int FC_LCAP = 32;     // candidates one CTA pushes per step
constexpr int FC_LBUF = FC_E * FC_MAXNBL;

static_assert(FC_CAP == 128 && FC_E * FC_MAXNBL <= 8 * 32, "thread mapping");

struct FcShared {
    // exchange area, double-buffered by step parity; written by the peers
    unsigned long long r_key[2][FC_C][FC_LCAP];
    float4 r_xyz[2][FC_C][FC_LCAP];
    unsigned long long r_S[2][FC_C];
    unsigned r_cnt[2][FC_C];
    // my exposed keys above my own threshold (and the 32 largest of them if there are more)
    unsigned long long l_key[FC_LBUF];
    float4 l_xyz[FC_LBUF];
    unsigned long long l2_key[FC_LCAP];
    float4 l2_xyz[FC_LCAP];
    // the cluster's candidates: gathered, then in descending key order
    unsigned long long c_key[FC_C * FC_LCAP];
    float4 c_xyz[FC_C * FC_LCAP];
    unsigned long long s_key[FC_CAP];
    float4 s_xyz[FC_CAP];
    unsigned K[FC_CAP][4], L[FC_CAP][4];   // row r, bit i: earlier candidate i lowers r / lowers r to a key <= T
    unsigned long long top[FC_MAXNBL][FC_E + 1];   // per bucket: its FC_E + 1 largest keys, descending
    unsigned F[FC_MAXNBL][4];              // per bucket: candidates (rows of s_xyz) that can reach it
    float box[6][FC_MAXNBL];
    unsigned long long T, l_S;
    unsigned accmask[4];                   // bit r: candidate r (row of s_xyz) was accepted in the last resolution
    unsigned l_cnt, n_cand, n_acc;
};

// the FC_E + 1 largest (dist bits, tie word) keys of the 32 points of one bucket
__device__ __forceinline__ void fc_bucket_top(unsigned hi, unsigned lo, unsigned long long *top, int lane) {
#pragma unroll
    for (int e = 0; e <= FC_E; ++e) {
        unsigned h = hi, l = lo;
        warp_max_pair(h, l);
        if (lane == 0) top[e] = ((unsigned long long)h << 32) | l;
        if (hi == h && lo == l) {
            hi = 0u;
            lo = 0u;
        }
    }
}

template <int FC_WARPS>
__global__ void __launch_bounds__(FC_WARPS * 32, FC_WARPS == 16 ? 2 : 4)
fps_cluster_kernel(Cloud xyz, dvcp_cloud_index_t index, int N, int npoint, const int64_t *__restrict__ start,
                   int64_t *__restrict__ out64, int32_t *__restrict__ out32) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int FC_THREADS = FC_WARPS * 32;
    static_assert(FC_WARPS >= FC_C && FC_THREADS % FC_CAP == 0 && FC_THREADS <= 4 * FC_CAP, "thread mapping");
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
    const float4 *gpt = reinterpret_cast<const float4 *>(index.sorted_pt) + (int64_t)b * cap;
    const float *gbox = index.bucket_box + (int64_t)b * NB * 8;
    for (int jl = warp; jl < NBL; jl += FC_WARPS) {
        const int jg = jl * FC_C + rank, gp = jg * 32 + lane, p = jl * 32 + lane;
        const float4 P = __ldg(gpt + gp);
        const int id = __float_as_int(P.w);
        s_x[p] = id < 0 ? 0.f : P.x;
        s_y[p] = id < 0 ? 0.f : P.y;
        s_z[p] = id < 0 ? 0.f : P.z;
        const unsigned idu = id < 0 ? 0xffffu : (unsigned)id;
        s_id[p] = (unsigned short)idu;
        const float d0 = id < 0 ? 0.0f : 1e10f;
        s_d[p] = d0;
        fc_bucket_top(__float_as_uint(d0), ((0xffffu - idu) << 16) | (unsigned)gp, sh.top[jl], lane);
        if (lane < 4) sh.F[jl][lane] = 0u;
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
        // the start point is "candidate 0, accepted" of a step that happened before the loop
        sh.s_xyz[0] = make_float4(xyz.at(b, (int)startidx, 0), xyz.at(b, (int)startidx, 1), xyz.at(b, (int)startidx, 2), 0.f);
        sh.accmask[0] = 1u;
        sh.accmask[1] = sh.accmask[2] = sh.accmask[3] = 0u;
        sh.n_acc = 1u;
        sh.n_cand = 0u;
        sh.l_cnt = 0u;
        if (rank == 0) {
            if (out64) out64[(int64_t)b * npoint] = startidx;
            if (out32) out32[(int64_t)b * npoint] = (int32_t)startidx;
        }
    }
    cluster.sync();   // peers exist and are initialised before anyone writes into them

    // Box tests: which of my buckets can candidate a (row a of s_xyz) reach? lane = bucket (its box stays in
    // registers while the candidates stream by), the warps first .. first + nw - 1 stride over the candidates.
    auto box_tests = [&](int m_cand, int first, int nw) {
        for (int h = 0; h < H2; ++h) {
            const int jl = h * 32 + lane;
            if (jl < NBL) {
                const float nx = sh.box[0][jl], ny = sh.box[1][jl], nz = sh.box[2][jl];
                const float xx = sh.box[3][jl], xy = sh.box[4][jl], xz = sh.box[5][jl];
                const float bestval = __uint_as_float((unsigned)(sh.top[jl][0] >> 32));
#pragma unroll 4
                for (int a = warp - first; a < m_cand; a += nw) {
                    const float4 c = sh.s_xyz[a];
                    const float ex = fmaxf(fmaxf(__fsub_rn(nx, c.x), __fsub_rn(c.x, xx)), 0.0f);
                    const float ey = fmaxf(fmaxf(__fsub_rn(ny, c.y), __fsub_rn(c.y, xy)), 0.0f);
                    const float ez = fmaxf(fmaxf(__fsub_rn(nz, c.z), __fsub_rn(c.z, xz)), 0.0f);
                    if (sq3_nofma(ex, ey, ez) < bestval) atomicOr(&sh.F[jl][a >> 5], 1u << (a & 31));
                }
            }
        }
    };
    __syncthreads();
    box_tests(1, 0, FC_WARPS);   // the start point

    int produced = 0;
#ifdef DVCP_FPS_TIMING
    long long t_last__ = clock64();
    if (tid == 0 && blockIdx.x < 1024) { unsigned sid; asm volatile("mov.u32 %0, %%smid;" : "=r"(sid)); g_fc_smid[blockIdx.x] = (int)sid; }
    if (blockIdx.x == 0 && tid == 0) for (int i = 0; i < 24; ++i) g_fc_time[i] = 0;
#endif
    for (int step = 0;; ++step) {
        const int buf = step & 1;
        FC_TICK(15);
        const int A = (int)sh.n_acc;
        FC_TICK(0);
        __syncthreads();
        FC_TICK(1);
        // ---- (2) lower the distances of the reached buckets, refresh their largest keys ----
        for (int jl = warp; jl < NBL; jl += FC_WARPS) {
            uint4 Fw = *reinterpret_cast<const uint4 *>(sh.F[jl]);
            const bool any = (Fw.x | Fw.y | Fw.z | Fw.w) != 0u;
            Fw.x &= sh.accmask[0]; Fw.y &= sh.accmask[1]; Fw.z &= sh.accmask[2]; Fw.w &= sh.accmask[3];   // accepted only
            if (any) {
                __syncwarp();
                if (lane < 4) sh.F[jl][lane] = 0u;
            }
            if (Fw.x | Fw.y | Fw.z | Fw.w) {
                const int p = jl * 32 + lane;
                const float x = s_x[p], y = s_y[p], z = s_z[p];
                float dk = s_d[p];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    unsigned F = w == 0 ? Fw.x : w == 1 ? Fw.y : w == 2 ? Fw.z : Fw.w;
                    while (F) {
                        const int a = w * 32 + __ffs(F) - 1;
                        F &= F - 1;
                        const float4 c = sh.s_xyz[a];
                        const float d = sq3_nofma(__fsub_rn(x, c.x), __fsub_rn(y, c.y), __fsub_rn(z, c.z));
                        dk = d < dk ? d : dk;
                    }
                }
                s_d[p] = dk;
                fc_bucket_top(__float_as_uint(dk),
                              ((0xffffu - (unsigned)s_id[p]) << 16) | (unsigned)((jl * FC_C + rank) * 32 + lane), sh.top[jl], lane);
            }
        }
        produced += A;
        if (produced >= npoint) break;
        FC_TICK(2);
        __syncthreads();
        FC_TICK(3);
        // ---- (3) my candidates: exposed keys above S_c = my largest (FC_E+1)-th bucket key ----
        unsigned long long Sc;
        {
            unsigned shi = 0u, slo = 0u;
            for (int h = 0; h < H2; ++h) {
                const int jl = h * 32 + lane;
                const unsigned long long s = jl < NBL ? sh.top[jl][FC_E] : 0ull;
                const unsigned hi = (unsigned)(s >> 32), lo = (unsigned)s;
                if (hi > shi || (hi == shi && lo > slo)) {
                    shi = hi;
                    slo = lo;
                }
            }
            warp_max_pair(shi, slo);   // every warp computes the same value
            Sc = ((unsigned long long)shi << 32) | slo;
        }
        {
            const int t = tid;   // FC_E * NBL <= 192 < FC_THREADS
            const int jl = t % NBL, e = t / NBL;
            const unsigned long long key = e < FC_E ? sh.top[jl][e] : 0ull;
            const bool isc = (key >> 32) != 0ull && key > Sc;
            const unsigned m = __ballot_sync(0xffffffffu, isc);
            if (m) {
                unsigned basev = 0u;
                if (lane == 0) basev = atomicAdd(&sh.l_cnt, (unsigned)__popc(m));
                basev = __shfl_sync(0xffffffffu, basev, 0);
                if (isc) {
                    const unsigned slot = basev + __popc(m & ((1u << lane) - 1u));
                    const int p = jl * 32 + (int)(key & 31u);
                    sh.l_key[slot] = key;
                    sh.l_xyz[slot] = make_float4(s_x[p], s_y[p], s_z[p], 0.f);
                }
            }
        }
        FC_TICK(4);
        __syncthreads();
        FC_TICK(5);
        unsigned lcnt = sh.l_cnt;
        const unsigned long long *pk = sh.l_key;
        const float4 *px = sh.l_xyz;
        if (lcnt > FC_LCAP) {
            // keep my 32 largest; the 33rd becomes my threshold (nothing above it is withheld)
            for (int t = tid; t < (int)lcnt; t += FC_THREADS) {
                const unsigned long long key = sh.l_key[t];
                int rk = 0;
                for (int u = 0; u < (int)lcnt; ++u) rk += sh.l_key[u] > key;
                if (rk < FC_LCAP) {
                    sh.l2_key[rk] = key;
                    sh.l2_xyz[rk] = sh.l_xyz[t];
                }
                if (rk == FC_LCAP) sh.l_S = key;
            }
            __syncthreads();
            Sc = sh.l_S;
            lcnt = FC_LCAP;
            pk = sh.l2_key;
            px = sh.l2_xyz;
        }
        // ---- (4) push them into every CTA of the cluster (warp w -> CTA w), one barrier ----
        if (warp < FC_C) {
            FcShared *peer = cluster.map_shared_rank(&sh, warp);
            if (lane < (int)lcnt) {
                peer->r_key[buf][rank][lane] = pk[lane];
                peer->r_xyz[buf][rank][lane] = px[lane];
            }
            if (lane == 0) {
                peer->r_cnt[buf][rank] = lcnt;
                peer->r_S[buf][rank] = Sc;
            }
        }
        FC_TICK(6);
        cluster.sync();
        FC_TICK(7);
        // ---- (5) every CTA resolves the same list: S, candidates above S ----
        unsigned long long S = 0ull;
#pragma unroll
        for (int c = 0; c < FC_C; ++c) {
            const unsigned long long v = sh.r_S[buf][c];
            S = v > S ? v : S;
        }
        if (warp < FC_C) {
            const int c = warp;   // one warp per source CTA
            const unsigned cnt = sh.r_cnt[buf][c];
            const unsigned long long key = lane < (int)cnt ? sh.r_key[buf][c][lane] : 0ull;
            const bool keep = key > S;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (m) {
                unsigned basev = 0u;
                if (lane == 0) basev = atomicAdd(&sh.n_cand, (unsigned)__popc(m));
                basev = __shfl_sync(0xffffffffu, basev, 0);
                if (keep) {
                    const unsigned slot = basev + __popc(m & ((1u << lane) - 1u));
                    sh.c_key[slot] = key;
                    sh.c_xyz[slot] = sh.r_xyz[buf][c][lane];
                }
            }
        }
        __syncthreads();
        FC_TICK(8);
        const int n = (int)sh.n_cand;
        if (n == 0) {
            // every remaining distance is 0: the argmax stays the lowest index (pointnet2_utils.py:83).
            // All keys have dist bits 0 now, so the winner is the largest tie word anywhere.
            if (warp == 0) {
                unsigned hi = 0u, lo = 0u;
                for (int h = 0; h < H2; ++h) {
                    const int jl = h * 32 + lane;
                    const unsigned long long k2 = jl < NBL ? sh.top[jl][0] : 0ull;
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
                unsigned long long mx = 0ull;
                for (int c = 0; c < FC_C; ++c) mx = sh.r_S[buf ^ 1][c] > mx ? sh.r_S[buf ^ 1][c] : mx;
                const unsigned idx = 0xffffu - ((unsigned)mx >> 16);
                for (int i = produced + tid; i < npoint; i += FC_THREADS) {
                    if (out64) out64[(int64_t)b * npoint + i] = idx;
                    if (out32) out32[(int64_t)b * npoint + i] = (int32_t)idx;
                }
            }
            break;
        }
        // descending key order (rank = number of larger keys; keys are distinct); beyond FC_CAP candidates the
        // (FC_CAP+1)-th key is the threshold T. P adjacent lanes share one key and count a slice of the list each
        // (n <= FC_C * FC_LCAP = 256 <= FC_THREADS / 2): the phase is a serial scan per thread, so its length is n / P
        {
            const int P = n * 4 <= FC_THREADS ? 4 : (n * 2 <= FC_THREADS ? 2 : 1);   // n <= 256 <= FC_THREADS
            const int sh_p = P == 4 ? 2 : (P == 2 ? 1 : 0);
            const int t = tid >> sh_p, part = tid & (P - 1);
            const unsigned long long key = t < n ? sh.c_key[t] : 0ull;
            // my slice of the list: a contiguous run, four independent loads in flight per trip
            const int len = (n + P - 1) >> sh_p, u0 = part * len, u1 = min(u0 + len, n);
            int rk = 0;
            if (t < n) {
                int u = u0;
                for (; u + 4 <= u1; u += 4) {
                    const unsigned long long a = sh.c_key[u], b = sh.c_key[u + 1], c = sh.c_key[u + 2], d = sh.c_key[u + 3];
                    rk += (int)(a > key) + (int)(b > key) + (int)(c > key) + (int)(d > key);
                }
                for (; u < u1; ++u) rk += sh.c_key[u] > key;
            }
            if (P >= 2) rk += __shfl_xor_sync(0xffffffffu, rk, 1);
            if (P == 4) rk += __shfl_xor_sync(0xffffffffu, rk, 2);
            if (t < n && part == 0) {
                if (rk < FC_CAP) {
                    sh.s_key[rk] = key;
                    sh.s_xyz[rk] = sh.c_xyz[t];
                }
                if (rk == FC_CAP) sh.T = key;
            }
        }
        FC_TICK(9);
        __syncthreads();
        FC_TICK(10);
        const int m = n < FC_CAP ? n : FC_CAP;
#ifdef DVCP_FPS_TIMING
        if (blockIdx.x == 0 && tid == 0) { g_fc_time[16] += 1; g_fc_time[17] += n; g_fc_time[18] += A; g_fc_time[19] += sh.l_cnt; g_fc_time[20] += (n > FC_CAP); }
#endif
        const unsigned long long T = n > FC_CAP ? sh.T : S;
        // pair tests: thread (row r, part) covers the earlier candidates i of its share of the 4 mask words
        {
            constexpr int PARTS = FC_THREADS / FC_CAP, WPP = 4 / PARTS;   // words per part
            const int r = tid & (FC_CAP - 1), part = tid / FC_CAP;
            if (r < m) {
                const unsigned long long key = sh.s_key[r];
                const float4 q = sh.s_xyz[r];
                const unsigned klo = (unsigned)key;
                const float dj = __uint_as_float((unsigned)(key >> 32));
#pragma unroll
                for (int ww = 0; ww < WPP; ++ww) {
                    const int w = part * WPP + ww;
                    unsigned kw = 0u, lw = 0u;
                    const int i0 = w * 32, i1 = r < i0 + 32 ? r : i0 + 32;
                    for (int i = i0; i < i1; ++i) {
                        const float4 c = sh.s_xyz[i];
                        const float d = sq3_nofma(__fsub_rn(q.x, c.x), __fsub_rn(q.y, c.y), __fsub_rn(q.z, c.z));
                        const bool kill = d < dj;
                        const unsigned long long nk = ((unsigned long long)__float_as_uint(d) << 32) | klo;
                        kw |= (unsigned)kill << (i & 31);
                        lw |= (unsigned)(kill && nk <= T) << (i & 31);
                    }
                    sh.K[r][w] = kw;
                    sh.L[r][w] = lw;
                }
            }
        }
        FC_TICK(11);
        __syncthreads();
        FC_TICK(12);
        if (warp == 0) {
            // lane l resolves rows l, l+32, l+64, l+96 (row r = word r/32, bit r%32 of the masks)
            unsigned Kq[4][4], Lq[4][4];
            int st[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = q * 32 + lane;
                st[q] = r < m ? 0 : 2;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    Kq[q][w] = (r < m && w <= q) ? sh.K[r][w] : 0u;
                    Lq[q][w] = (r < m && w <= q) ? sh.L[r][w] : 0u;
                }
            }
            unsigned acc[4] = {0u, 0u, 0u, 0u}, rej[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) rej[q] = __ballot_sync(0xffffffffu, st[q] == 2);
            // first-come resolution in key order: accepted unless an ACCEPTED earlier candidate lowers me
            while ((acc[0] | rej[0]) != 0xffffffffu || (acc[1] | rej[1]) != 0xffffffffu ||
                   (acc[2] | rej[2]) != 0xffffffffu || (acc[3] | rej[3]) != 0xffffffffu) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (st[q] == 0) {
                        unsigned kill = 0u, open = 0u;
#pragma unroll
                        for (int w = 0; w <= q; ++w) {
                            kill |= Kq[q][w] & acc[w];
                            open |= Kq[q][w] & ~rej[w];
                        }
                        if (kill) st[q] = 2;
                        else if (!open) st[q] = 1;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc[q] = __ballot_sync(0xffffffffu, st[q] == 1);
                    rej[q] = __ballot_sync(0xffffffffu, st[q] == 2);
                }
            }
            // a candidate lowered but still above T ends the step: keep what comes before the first such one
            bool stopped = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                unsigned low = 0u;
#pragma unroll
                for (int w = 0; w <= q; ++w) low |= Lq[q][w] & acc[w];
                const unsigned sm = __ballot_sync(0xffffffffu, q * 32 + lane < m && st[q] == 2 && low == 0u);
                if (stopped) {
                    acc[q] = 0u;
                } else if (sm) {
                    acc[q] &= (1u << (__ffs(sm) - 1)) - 1u;
                    stopped = true;
                }
            }
            const int rem = npoint - produced;
            int below = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = q * 32 + lane;
                const int rk = below + __popc(acc[q] & ((1u << lane) - 1u));
                const bool take = ((acc[q] >> lane) & 1u) && rk < rem;
                if (take && rank == 0) {
                    const unsigned idx = 0xffffu - ((unsigned)sh.s_key[r] >> 16);
                    if (out64) out64[(int64_t)b * npoint + produced + rk] = idx;
                    if (out32) out32[(int64_t)b * npoint + produced + rk] = (int32_t)idx;
                }
                const unsigned tm = __ballot_sync(0xffffffffu, take);
                if (lane == 0) sh.accmask[q] = tm;    // the next step applies rows tm of s_xyz
                below += __popc(acc[q]);
            }
            if (lane == 0) {
                sh.n_acc = (unsigned)(below < rem ? below : rem);
                sh.n_cand = 0u;
                sh.l_cnt = 0u;
            }
        } else {
            // meanwhile: which buckets can each CANDIDATE reach (the accepted ones are a subset; the bucket maxima the
            // test reads do not change before the next step applies them)
            box_tests(m, 1, FC_WARPS - 1);
        }
        FC_TICK(13);
        __syncthreads();
        FC_TICK(14);
    }
    cluster.sync();   // nobody leaves while a peer may still address its shared memory
}

}  // namespace dvcp

// Launcher used by dvcp_fps (fps.cu). The index must already hold the cloud (dvcp_build_index).
template <int W>
static int fps_cluster_launch_w(dvcp::Cloud xyz, dvcp_cloud_index_t index, int B, int N, int npoint, const int64_t *start,
                                int64_t *out64, int32_t *out32, size_t smem, cudaStream_t st) {
    using namespace dvcp;
    auto k = fps_cluster_kernel<W>;
    DVCP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * FC_C));
    cfg.blockDim = dim3(W * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = FC_C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DVCP_CUDA(cudaLaunchKernelEx(&cfg, k, xyz, index, N, npoint, start, out64, out32));
    return 0;
}

// small_cta != 0: 8 warps per CTA instead of 16 (slower alone, friendlier to concurrent kernels)
int dvcp_fps_cluster_launch(dvcp_cloud_t xyz, dvcp_cloud_index_t index, int B, int N, int npoint, const int64_t *start,
                            int64_t *out64, int32_t *out32, int small_cta, cudaStream_t st) {
    using namespace dvcp;
    const int NB = index.cap / 32;
    if (NB % FC_C != 0 || NB / FC_C > FC_MAXNBL || NB / FC_C < 1) return DVCP_E_UNSUPPORTED;
    const int NBL = NB / FC_C;
    const size_t smem = sizeof(FcShared) + (size_t)NBL * 32 * (4 * sizeof(float) + sizeof(unsigned short));
    if (small_cta) return fps_cluster_launch_w<8>(as_cloud(xyz), index, B, N, npoint, start, out64, out32, smem, st);
    return fps_cluster_launch_w<16>(as_cloud(xyz), index, B, N, npoint, start, out64, out32, smem, st);
}

#ifdef DVCP_FPS_TIMING
extern "C" __attribute__((visibility("default"))) int dvcp_debug_fps_smid(int *host1024) {
    return (int)cudaMemcpyFromSymbol(host1024, dvcp::g_fc_smid, 1024 * sizeof(int));
}
extern "C" __attribute__((visibility("default"))) int dvcp_debug_fps_timing(long long *host16) {
    return (int)cudaMemcpyFromSymbol(host16, dvcp::g_fc_time, 24 * sizeof(long long));
}
#endif
