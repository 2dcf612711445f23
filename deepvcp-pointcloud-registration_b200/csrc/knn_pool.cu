// Exact K-nearest-neighbour search for GROUPS of nearby queries (the G^3 candidate lattice around one
// transformed key-point, voxelize.py:19-83 -> get_cat_feat_tgt.py:45,52): one CTA per group.
//
// Contract: float32, d2 = fma(dz,dz, fma(dy,dy, dx*dx)), neighbours ordered by (d2, index),
// sqrt_rn(d2) returned. Results are bit-identical to knn_kernel / knn_indexed_kernel (knn.cu).
//
//   1. POOL. The eight corners of the group's bounding box are solved exactly through the cloud's spatial
//      index (one warp each): their K-th distances D_c say how far the group has to look in every direction.
//      The pool is the 26-face polytope {p : v.p <= max_c (v.c + |v| D_c)} over the directions v in
//      {-1,0,1}^3 (the support planes of the eight balls B(c, D_c)); every target point inside it is copied
//      from the index into shared memory, binned by counting sort into (x, y) cells of the lattice step.
//   2. QUERIES. A warp walks the z-lines of one x-slab of the lattice in boustrophedon order. The K results
//      of the previous query, re-evaluated for this one from shared memory, bound the K-th distance (thr);
//      only the cells within sqrt(thr) of the query are scanned (runs of consecutive pool entries), points
//      with d2 <= thr are compacted into a per-warp list and the K smallest are selected by a bitonic
//      network over 64-bit (d2, index) keys.
//   3. CERTIFICATE. A point that is NOT in the pool violates one of the 26 planes, so it is farther from the
//      query than the query's distance m to that plane. If the pool's K-th distance is below the smallest of
//      the 26 plane distances (minus a rounding allowance), the pool result is the exact global result, ties
//      included. Otherwise the query is answered by the index search of knn_index.cuh with the pool's K-th
//      distance as its bound. The shape of the pool is therefore a matter of speed only, never of results.
//
// Nothing here depends on the queries really being a lattice: any group of queries gives exact results;
// the lattice only makes the pool small and the chained bounds tight.
#include <stdlib.h>

#include "knn_index.cuh"

namespace dvcp {

constexpr int KP_WARPS = 8;
constexpr int KP_THREADS = KP_WARPS * 32;
constexpr int KP_POOL_MAX = 8192;  // pool capacity limit (points): 13-bit positions in the keys
constexpr int KP_POOL_DEFAULT = 3072;
constexpr int KP_POSBITS = 13;
constexpr int KP_BLIST = 1024;     // buckets the pool may draw from
constexpr int KP_MAXNC = 40;       // cells per axis
constexpr int KP_HALO = 3;         // halo cells on each side (the outermost cell catches everything beyond)
constexpr float KP_HALO_SCALE = 1.05f;
constexpr int KP_BUF = 128;        // qualifying points collected per query

struct KpWork {   // a query left to the index search
    int64_t q;     // b * Q + query
    float thr;     // an upper bound of its K-th squared distance (+inf: none)
};

struct KpShared {   // follows the pool (float4[pool_cap]) in dynamic shared memory
    unsigned long long buf[KP_WARPS][KP_BUF];
    unsigned short start[KP_MAXNC * KP_MAXNC + 2];
    unsigned short cur[KP_MAXNC * KP_MAXNC + 2];
    unsigned short blist[KP_BLIST];
    float red[6][KP_WARPS];
    float qlo[3], qhi[3];     // query box
    float off[32];            // support offsets of the 26 pool planes (direction = kp_dir(l))
    float cD[8];              // corner radii
    float g0[2], inv_cell;    // cell grid origin (x, y), 1 / cell
    int nc;                   // cells per axis
    int nblist, npool, whole; // whole: the pool holds every point of the cloud
    int unit_next;
    int ok;                   // 0: the pool could not be built (too dense): every query goes to the index search
};

__device__ __forceinline__ unsigned long long kp_key(float d2, int id, int pos) {
    return ((unsigned long long)__float_as_uint(d2) << 32) | ((unsigned)id << KP_POSBITS) | (unsigned)pos;
}

// plane l (0..25) has the direction ((id / 9) - 1, (id / 3) % 3 - 1, id % 3 - 1) with id = l < 13 ? l : l + 1
__device__ __forceinline__ void kp_dir(int l, float &vx, float &vy, float &vz) {
    const int id = l < 13 ? l : l + 1;
    vx = (float)(id / 9 - 1);
    vy = (float)((id / 3) % 3 - 1);
    vz = (float)(id % 3 - 1);
}
// inside all 26 planes (sums of +-x +-y +-z in float32; the certificate allows for their rounding)
__device__ __forceinline__ bool kp_inside(const float4 &P, const float *off) {
    bool in = __float_as_int(P.w) >= 0;   // unused slots of the index: id -1
#pragma unroll
    for (int l = 0; l < 26; ++l) {
        const int id = l < 13 ? l : l + 1;
        const int dx = id / 9 - 1, dy = (id / 3) % 3 - 1, dz = id % 3 - 1;
        float sdot = 0.f;
        if (dx) sdot = dx > 0 ? P.x : -P.x;
        if (dy) sdot = (dx ? sdot : 0.f) + (dy > 0 ? P.y : -P.y);
        if (dz) sdot = ((dx || dy) ? sdot : 0.f) + (dz > 0 ? P.z : -P.z);
        in = in && sdot <= off[l];
    }
    return in;
}

__device__ __forceinline__ int kp_cell(float v, float g0, float inv, int nc) {
    const float t = (v - g0) * inv;
    return t < 0.f ? 0 : min((int)fminf(t, 1e6f), nc - 1);
}

__device__ __forceinline__ void kp_flush_stats(unsigned long long *stats, int lane, int tid, int npool, unsigned a,
                                               unsigned b, unsigned c, unsigned d, unsigned e, unsigned f, unsigned g) {
    if (!stats || lane != 0) return;
    atomicAdd(stats + 0, (unsigned long long)a);
    atomicAdd(stats + 1, (unsigned long long)b);
    atomicAdd(stats + 2, (unsigned long long)c);
    atomicAdd(stats + 3, (unsigned long long)d);
    atomicAdd(stats + 4, (unsigned long long)e);
    if (tid == 0) atomicAdd(stats + 5, (unsigned long long)npool);
    atomicAdd(stats + 6, (unsigned long long)f);
    atomicAdd(stats + 7, (unsigned long long)g);
}

template <int T, bool BIG>
__global__ void __launch_bounds__(KP_THREADS, 3)
knn_pool_kernel(dvcp_cloud_index_t index, const float *__restrict__ query, int64_t Q, int K, int gsz, int zline,
                float cell_in, int nvalid, int pool_cap, float halo_scale, float *__restrict__ dist, int64_t *__restrict__ idx64,
                int32_t *__restrict__ idx32, unsigned long long *__restrict__ stats, KpWork *__restrict__ work,
                unsigned int *__restrict__ work_count) {
    constexpr int TT = T < 32 ? T : 32;
    constexpr int SB = (T + 31) / 32;
    constexpr int NBUCKETS = T * 32;
    extern __shared__ __align__(16) unsigned char kp_smem[];
    float4 *pool = reinterpret_cast<float4 *>(kp_smem);
    KpShared &S = *reinterpret_cast<KpShared *>(kp_smem + (size_t)pool_cap * sizeof(float4));
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned long long INF = 0xffffffffffffffffull;
#ifdef DVCP_KNN_GROUP_STATS
    const long long t_begin = clock64();
    long long t_probe = 0, t_pool = 0;
#endif
    const int cap = index.cap;
    const float *box = index.bucket_box + (int64_t)b * (cap / 32) * 8;
    const float4 *spt = reinterpret_cast<const float4 *>(index.sorted_pt) + (int64_t)b * cap;
    const int64_t q_begin = (int64_t)blockIdx.x * gsz;
    const int nq = (int)min((int64_t)gsz, Q - q_begin);
    const float *qg = query + ((int64_t)b * Q + q_begin) * 3;

    // ---- query box ----
    {
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = tid; i < nq; i += KP_THREADS) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = __ldg(qg + 3 * i + c);
                mn[c] = fminf(mn[c], v);
                mx[c] = fmaxf(mx[c], v);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int s = 16; s; s >>= 1) {
                mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], s));
                mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], s));
            }
            if (lane == 0) {
                S.red[c][warp] = mn[c];
                S.red[3 + c][warp] = mx[c];
            }
        }
        __syncthreads();
        if (tid < 3) {
            float a = INFINITY, z = -INFINITY;
            for (int w = 0; w < KP_WARPS; ++w) {
                a = fminf(a, S.red[tid][w]);
                z = fmaxf(z, S.red[3 + tid][w]);
            }
            S.qlo[tid] = a;
            S.qhi[tid] = z;
        }
        if (tid == 0) {
            S.nblist = 0;
            S.unit_next = 0;
        }
        __syncthreads();
    }
    const float qlx = S.qlo[0], qly = S.qlo[1], qlz = S.qlo[2], qhx = S.qhi[0], qhy = S.qhi[1], qhz = S.qhi[2];
    // cell size: the lattice step, enlarged if the query box would need more than KP_MAXNC cells
    float cell = cell_in;
    {
        const float ext = fmaxf(qhx - qlx, qhy - qly);
        const float need = ext / (float)(KP_MAXNC - 2 * KP_HALO - 2);
        if (!(cell > need)) cell = need * 1.0001f + 1e-20f;
    }

    // ---- index search context (halo probes and the fallback of uncertified queries) ----
    KnnCtx ctx;
    ctx.box = box;
    ctx.spt = spt;
    ctx.buf = S.buf[warp];
    ctx.spt_lane = spt + lane;
    ctx.box_lane = reinterpret_cast<const float4 *>(box) + lane * 2;
    ctx.K = K;
    ctx.loose = INFINITY;
    Box6 sb[SB];
#pragma unroll
    for (int s = 0; s < SB; ++s) sb[s] = load_super_box(box, TT, s * 32 + lane);

    // ---- corner probes: exact K-th distances of the corners of the query box (one warp each) ----
    const bool probe = nvalid > pool_cap;   // a cloud that fits the pool whole needs no shape
    if (probe) {
        for (int cn = warp; cn < 8; cn += KP_WARPS) {
            const float cx = (cn & 1) ? qhx : qlx, cy = (cn & 2) ? qhy : qly, cz = (cn & 4) ? qhz : qlz;
            const unsigned long long r = knn_query_thr<T, BIG>(ctx, sb, cx, cy, cz, INFINITY, lane);
            const unsigned long long rk = __shfl_sync(0xffffffffu, r, K - 1);
            if (lane == 0) S.cD[cn] = sqrtf(__uint_as_float(KnnKey<BIG>::d2bits(rk)));
        }
    }
#ifdef DVCP_KNN_GROUP_STATS
    __syncthreads();
    t_probe = clock64();
#endif
    if (tid == 0) {
        // cell grid over x, y: KP_HALO cells around the query box, clamped outside
        const float ext = fmaxf(qhx - qlx, qhy - qly);
        int nc = (int)(ext / cell) + 1 + 2 * KP_HALO;
        if (nc > KP_MAXNC) nc = KP_MAXNC;
        S.nc = nc;
        S.g0[0] = qlx - KP_HALO * cell;
        S.g0[1] = qly - KP_HALO * cell;
        S.inv_cell = 1.0f / cell;
    }
    __syncthreads();
    const int nc = S.nc;
    const float g0x = S.g0[0], g0y = S.g0[1], inv_cell = S.inv_cell;
    bool pool_ok = false;
    float lx = 0.f, ly = 0.f, lz = 0.f, hx = 0.f, hy = 0.f, hz = 0.f;   // bounding box of the pool polytope
    for (int attempt = 0; attempt < 4 && !pool_ok; ++attempt) {
        __syncthreads();
        if (tid < 32) {
            float o = INFINITY;
            if (probe && tid < 26) {
                // the allowance over the corners' K-th distances shrinks until the pool fits: far from the cloud the
                // K-th distance is close to convex over the box and a small allowance still certifies most queries
                const float shrink[4] = {1.0f, 0.2f, 0.04f, 0.008f};
                const float scale = 1.0f + shrink[attempt] * (halo_scale - 1.0f);
                float vx, vy, vz;
                kp_dir(tid, vx, vy, vz);
                const float vn = sqrtf(vx * vx + vy * vy + vz * vz);
                o = -INFINITY;
                for (int cn = 0; cn < 8; ++cn) {
                    const float cx = (cn & 1) ? qhx : qlx, cy = (cn & 2) ? qhy : qly, cz = (cn & 4) ? qhz : qlz;
                    const float D = fmaf(S.cD[cn], scale, 1e-5f * (fabsf(cx) + fabsf(cy) + fabsf(cz)) + 1e-6f);
                    o = fmaxf(o, vx * cx + vy * cy + vz * cz + vn * D);
                }
            }
            S.off[tid] = o;
        }
        __syncthreads();
        lx = -S.off[4]; hx = S.off[21]; ly = -S.off[10]; hy = S.off[15]; lz = -S.off[12]; hz = S.off[13];
        for (int i = tid; i < nc * nc + 2; i += KP_THREADS) S.cur[i] = 0;
        if (tid == 0) {
            S.nblist = 0;
            S.npool = 0;
        }
        __syncthreads();
        // ---- the buckets whose box touches the pool box ----
        for (int j = tid; j < NBUCKETS; j += KP_THREADS) {
            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(box + (int64_t)j * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4 *>(box + (int64_t)j * 8) + 1);
            if (b1.z > 0.f && b0.x <= hx && b0.w >= lx && b0.y <= hy && b1.x >= ly && b0.z <= hz && b1.y >= lz) {
                const int at = atomicAdd(&S.nblist, 1);
                if (at < KP_BLIST) S.blist[at] = (unsigned short)j;
            }
        }
        __syncthreads();
        const int nbl = S.nblist;
        if (nbl <= KP_BLIST) {
            // ---- count the pool points per cell ----
            int mine = 0;
            for (int i = warp; i < nbl; i += KP_WARPS) {
                const float4 P = __ldg(spt + (int)S.blist[i] * 32 + lane);
                const bool in = kp_inside(P, S.off);
                if (in) {
                    const int c = kp_cell(P.x, g0x, inv_cell, nc) * nc + kp_cell(P.y, g0y, inv_cell, nc);
                    atomicAdd(reinterpret_cast<unsigned int *>(S.cur) + (c >> 1), 1u << (16 * (c & 1)));
                }
                mine += __popc(__ballot_sync(0xffffffffu, in));
            }
            if (lane == 0 && mine) atomicAdd(&S.npool, mine);
        }
        __syncthreads();
        pool_ok = nbl <= KP_BLIST && S.npool <= pool_cap;   // (a 16-bit cell counter cannot have wrapped then)
        __syncthreads();
        if (!probe) break;   // (cannot happen: the whole cloud fits) no second shape to try
    }
    if (pool_ok) {
        const int nbl = S.nblist;
        // exclusive prefix sum over the cells (one warp; <= 1600 cells)
        if (warp == 0) {
            int run = 0;
            for (int base = 0; base < nc * nc; base += 32) {
                const int i = base + lane;
                const int v = i < nc * nc ? S.cur[i] : 0;
                int inc = v;
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, inc, s);
                    if (lane >= s) inc += t;
                }
                if (i < nc * nc) S.start[i] = (unsigned short)(run + inc - v);
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) {
                S.start[nc * nc] = (unsigned short)run;
                S.whole = run == nvalid;   // every point of the cloud is in the pool: no certificate needed
            }
        }
        __syncthreads();
        for (int i = tid; i < nc * nc + 2; i += KP_THREADS) S.cur[i] = 0;
        __syncthreads();
        // ---- place ----
        for (int i = warp; i < nbl; i += KP_WARPS) {
            const float4 P = __ldg(spt + (int)S.blist[i] * 32 + lane);
            const bool in = kp_inside(P, S.off);
            if (in) {
                const int c = kp_cell(P.x, g0x, inv_cell, nc) * nc + kp_cell(P.y, g0y, inv_cell, nc);
                const unsigned old = atomicAdd(reinterpret_cast<unsigned int *>(S.cur) + (c >> 1), 1u << (16 * (c & 1)));
                const int off = (old >> (16 * (c & 1))) & 0xffffu;
                pool[S.start[c] + off] = P;
            }
        }
        __syncthreads();
    }
#ifdef DVCP_KNN_GROUP_STATS
    t_pool = clock64();
#endif
    const int npool = pool_ok ? S.npool : 0;
    const bool whole = pool_ok && S.whole != 0;
    unsigned long long *buf = S.buf[warp];
    // this lane's pool plane (lanes >= 26: none) for the certificate
    float pvx = 0.f, pvy = 0.f, pvz = 0.f, pinv = 0.f, poff = INFINITY;
    if (lane < 26) {
        kp_dir(lane, pvx, pvy, pvz);
        pinv = rsqrtf(pvx * pvx + pvy * pvy + pvz * pvz) * 0.99999f;
        poff = S.off[lane];
    }

    // optional profiling counters (per warp, flushed once): 0 certified, 1 uncertified, 2 list overflow,
    // 3 cold starts, 4 queries of groups without a pool, 5 pool points (per group), 6 admitted, 7 scanned
    unsigned st_cert = 0, st_unc = 0, st_ovf = 0, st_cold = 0, st_nopool = 0, st_adm = 0, st_scan = 0;
    // ---- queries: units = x-slabs of zline lines (zline^2 queries), claimed dynamically ----
    const int usz = zline * zline;
    const int nunits = (nq + usz - 1) / usz;
    while (true) {
        int unit = 0;
        if (lane == 0) unit = atomicAdd(&S.unit_next, 1);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= nunits) break;
        const int u0 = unit * usz, u1 = min(u0 + usz, nq);
        unsigned long long list = INF;   // previous result in POOL keys (lane j: j-th neighbour); INF = none
        for (int qi = u0; qi < u1; ++qi) {
            int ql = qi;   // boustrophedon over the z-lines of the unit
            {
                const int i = qi - u0, line = i / zline, k = i - line * zline;
                if ((line & 1) && u0 + (line + 1) * zline <= u1) ql = u0 + line * zline + (zline - 1 - k);
            }
            const float qx = __ldg(qg + 3 * ql), qy = __ldg(qg + 3 * ql + 1), qz = __ldg(qg + 3 * ql + 2);
            unsigned long long res = INF;
            bool done = false;
            float fb_thr = INFINITY;   // bound handed to the index search if the pool cannot certify
            if (pool_ok && npool >= K) {
                // ---- bound from the previous result ----
                float thr = INFINITY;
                if (__any_sync(0xffffffffu, list != INF)) {
                    float d2 = 0.f;
                    if (lane < K) {
                        const float4 P = pool[(unsigned)list & ((1u << KP_POSBITS) - 1u)];
                        d2 = sqdist_direct(P.x - qx, P.y - qy, P.z - qz);
                    }
                    thr = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(d2)));
                }
                int cnt = 0;
                if (thr < INFINITY) {
                    // merge the n collected keys into the running list `res` (K best so far, ascending over the lanes)
                    auto flush = [&](int n) {
                        __syncwarp();
                        for (int g = 0; g < n; g += 32) {
                            unsigned long long k = g + lane < n ? buf[g + lane] : INF;
                            k = bitonic_sort32(k, lane);
                            if (res == INF && __all_sync(0xffffffffu, res == INF)) {
                                res = k;
                            } else {
                                const unsigned long long r = __shfl_sync(0xffffffffu, k, 31 - lane);
                                res = u64min(res, r);
#pragma unroll
                                for (int j = 16; j > 0; j >>= 1) res = cmpx64(res, j, (lane & j) == 0);
                            }
                        }
                        __syncwarp();
                    };
                    // ---- scan the cells within sqrt(thr) of the query ----
                    const float r0 = sqrtf(thr);
                    const float rho = fmaf(r0, 1.00002f, 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + r0) + 1e-30f);
                    const int cx0 = kp_cell(qx - rho, g0x, inv_cell, nc), cx1 = kp_cell(qx + rho, g0x, inv_cell, nc);
                    const int cy0 = kp_cell(qy - rho, g0y, inv_cell, nc), cy1 = kp_cell(qy + rho, g0y, inv_cell, nc);
                    const bool fully = cy0 == 0 && cy1 == nc - 1;   // whole rows: one run
                    for (int cx = cx0; cx <= cx1; ++cx) {
                        int i0 = S.start[cx * nc + cy0], i1 = S.start[cx * nc + cy1 + 1];
                        if (fully) {
                            i1 = S.start[(cx1 + 1) * nc];
                            cx = cx1;
                        }
                        st_scan += i1 - i0;
                        for (int i = i0 + lane; i < i1 + ((i0 - i1) & 31); i += 32) {
                            bool qual = false;
                            float d2 = 0.f;
                            float4 P = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (i < i1) {
                                P = pool[i];
                                d2 = sqdist_direct(P.x - qx, P.y - qy, P.z - qz);
                                qual = d2 <= thr;
                            }
                            const unsigned m = __ballot_sync(0xffffffffu, qual);
                            if (m) {
                                if (cnt + __popc(m) > KP_BUF) {
                                    // list full (a loose bound, or a dense shell just beyond the K-th neighbour): fold it
                                    // into the running K best and go on with THEIR K-th distance as the bound
                                    flush(cnt);
                                    st_adm += cnt;
                                    cnt = 0;
                                    ++st_ovf;
                                    const unsigned long long w = __shfl_sync(0xffffffffu, res, K - 1);
                                    if (w != INF) thr = fminf(thr, __uint_as_float((unsigned)(w >> 32)));
                                    qual = qual && d2 <= thr;
                                }
                                const unsigned m2 = __ballot_sync(0xffffffffu, qual);
                                const int slot = cnt + __popc(m2 & ((1u << lane) - 1u));
                                if (qual) buf[slot] = kp_key(d2, __float_as_int(P.w), i);
                                cnt += __popc(m2);
                            }
                        }
                    }
                    flush(cnt);
                    st_adm += cnt;
                    cnt = K;   // the K re-evaluated points of the previous result are in the pool and within thr
                } else {
                    // ---- cold start: stream the whole pool through the sorted list ----
                    unsigned long long worst = INF;
                    for (int i = lane; i < npool + ((-npool) & 31); i += 32) {
                        unsigned long long key = INF;
                        if (i < npool) {
                            const float4 P = pool[i];
                            key = kp_key(sqdist_direct(P.x - qx, P.y - qy, P.z - qz), __float_as_int(P.w), i);
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, key < worst);
                        if (m) knn_merge(res, worst, key, m, K, lane);
                    }
                    cnt = K;
                    ++st_cold;
                    st_scan += npool;
                }
                if (cnt >= K) {
                    if (lane >= K) res = INF;
                    // ---- certificate ----
                    const unsigned kd = (unsigned)(__shfl_sync(0xffffffffu, res, K - 1) >> 32);
                    const float kth = __uint_as_float(kd);
                    // distance from the query to this lane's plane, rounded DOWN-safe: an excluded point p has
                    // fl(v.p) > off, the float32 sums err by < 1e-6 (|p|_1 + |q|_1), and only points within
                    // sqrt(kth) of the query matter
                    float m = (poff - (pvx * qx + pvy * qy + pvz * qz)) * pinv;
                    m = m - 4e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + fabsf(m));
                    if (!(poff < INFINITY)) m = INFINITY;   // lanes without a plane; a pool that holds the whole cloud
                    m = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(fmaxf(m, 0.f))));   // +inf: no plane
                    done = whole || kth * 1.00001f < m * m;
                    if (!done) fb_thr = kth;
                }
            }
            if (done) {
                if (lane < K) {
                    const int64_t o = ((int64_t)b * Q + q_begin + ql) * K + lane;
                    dist[o] = __fsqrt_rn(__uint_as_float((unsigned)(res >> 32)));
                    const unsigned id = (unsigned)res >> KP_POSBITS;
                    if (idx64) idx64[o] = id;
                    if (idx32) idx32[o] = (int32_t)id;
                }
                list = res;
                ++st_cert;
            } else {
                if (pool_ok) ++st_unc; else ++st_nopool;
                // ---- not certified: the exact index search, DEFERRED. Such queries cluster in a few groups (far
                //      from the cloud, dense shell beyond the K-th neighbour) and each costs tens of thousands of
                //      instructions; a second kernel spreads them over the whole GPU, one warp per query, with
                //      the best bound at hand ----
                if (!pool_ok) {   // no pool: bound from the nearest corner probe (d_K is 1-Lipschitz), if probed
                    if (probe) {
                        float bsum = INFINITY;
                        for (int cn = 0; cn < 8; ++cn) {
                            const float cx = (cn & 1) ? qhx : qlx, cy = (cn & 2) ? qhy : qly, cz = (cn & 4) ? qhz : qlz;
                            const float dq = sqrtf((cx - qx) * (cx - qx) + (cy - qy) * (cy - qy) + (cz - qz) * (cz - qz));
                            bsum = fminf(bsum, S.cD[cn] + dq);
                        }
                        bsum = fmaf(bsum, 1.0001f, 1e-5f * (fabsf(qx) + fabsf(qy) + fabsf(qz)) + 1e-6f);
                        fb_thr = bsum * bsum;
                    }
                }
                if (lane == 0) {
                    const unsigned at = atomicAdd(work_count, 1u);
                    work[at] = KpWork{(int64_t)b * Q + q_begin + ql, fb_thr};
                }
                // the pool's own K best (res) stay a valid seed for the next query if they exist
                list = (pool_ok && npool >= K && res != INF) ? res : INF;
                if (!__all_sync(0xffffffffu, lane >= K || list != INF)) list = INF;
            }
        }
    }
    kp_flush_stats(stats, lane, tid, npool, st_cert, st_unc, st_ovf, st_cold, st_nopool, st_adm, st_scan);
#ifdef DVCP_KNN_GROUP_STATS   // development: per-group cycles / pool size / probe cycles after the 8 counters
    __syncthreads();
    if (stats && tid == 0) {
        const long long g = (long long)b * gridDim.x + blockIdx.x, ng = (long long)gridDim.x * gridDim.y;
        stats[8 + g] = (unsigned long long)(clock64() - t_begin);
        stats[8 + ng + g] = (unsigned long long)npool;
        stats[8 + 2 * ng + g] = (unsigned long long)(t_probe - t_begin);
        stats[8 + 3 * ng + g] = (unsigned long long)(t_pool - t_begin);
    }
#endif
}

// The deferred queries: one warp per query, exact index search (knn_index.cuh) with the given bound.
template <int T, bool BIG>
__global__ void __launch_bounds__(KNI_WARPS * 32)
knn_deferred_kernel(dvcp_cloud_index_t index, const float *__restrict__ query, int64_t Q, int K,
                    const KpWork *__restrict__ work, const unsigned int *__restrict__ work_count,
                    float *__restrict__ dist, int64_t *__restrict__ idx64, int32_t *__restrict__ idx32) {
    constexpr int TT = T < 32 ? T : 32;
    constexpr int SB = (T + 31) / 32;
    __shared__ unsigned long long s_buf[KNI_WARPS][KNI_BUF];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned n = *work_count;
    const int cap = index.cap;
    for (unsigned w = blockIdx.x * KNI_WARPS + warp; w < n; w += gridDim.x * KNI_WARPS) {
        const KpWork it = work[w];
        const int b = (int)(it.q / Q);
        KnnCtx ctx;
        ctx.box = index.bucket_box + (int64_t)b * (cap / 32) * 8;
        ctx.spt = reinterpret_cast<const float4 *>(index.sorted_pt) + (int64_t)b * cap;
        ctx.buf = s_buf[warp];
        ctx.spt_lane = ctx.spt + lane;
        ctx.box_lane = reinterpret_cast<const float4 *>(ctx.box) + lane * 2;
        ctx.K = K;
        ctx.loose = INFINITY;
        Box6 sb[SB];
#pragma unroll
        for (int s = 0; s < SB; ++s) sb[s] = load_super_box(ctx.box, TT, s * 32 + lane);
        const float *qp = query + it.q * 3;
        const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
        const unsigned long long r = knn_query_thr<T, BIG>(ctx, sb, qx, qy, qz, it.thr, lane);
        if (lane < K) {
            const int64_t o = it.q * K + lane;
            dist[o] = __fsqrt_rn(__uint_as_float(KnnKey<BIG>::d2bits(r)));
            const unsigned id = KnnKey<BIG>::id(r);
            if (idx64) idx64[o] = id;
            if (idx32) idx32[o] = (int32_t)id;
        }
        __syncwarp();
    }
}

// =====================================================================================================
// Two-kernel form (the default): the pools live in GLOBAL memory (L2 / L1 resident: ~1 K points per group).
//   knn_pool_build_kernel   one CTA per group: exact probes (centre, then the 8 corners bounded through the
//                           centre: d_K is 1-Lipschitz), 26-plane pool shape, counting sort of the pool points by
//                           (x, y) cell into the group's slice of the workspace, per-group meta data;
//   knn_pool_query_kernel   like knn_indexed_kernel: 4-warp CTAs, one warp per CHAIN of z-lines, but the
//                           candidates of a query come from the cells of its group's pool within sqrt(thr)
//                           (runs of consecutive pool entries, float4 loads) instead of a walk over the cloud's
//                           bucket boxes; certificate per query; what is not certified goes to the work list;
//   knn_deferred_kernel     the work list through the index search, one warp per query.
// The monolithic knn_pool_kernel above keeps a group's pool in shared memory; it loses to this form on
// occupancy and on load balance (a dense group is 5-10 x the work of a sparse one).
struct KpMeta {
    float off[32];            // pool planes
    float cx[8], cy[8], cz[8], cD[8];   // the probed corners and their K-th distances (Lipschitz bounds for cold starts)
    float g0x, g0y, g0z, inv_cell;
    int nc, npool, whole, ok;
};
constexpr int KP3_MAXNC = 20;                                   // cells per axis of the 3-D cell grid
constexpr int KP3_CELLS = KP3_MAXNC * KP3_MAXNC * KP3_MAXNC;
constexpr int KP_STARTS = KP3_CELLS + 8;                        // u16 entries per group

struct KpBuildShared {
    unsigned long long buf[KP_WARPS][KP_BUF];
    unsigned short start[KP3_CELLS + 8];
    unsigned short cur[KP3_CELLS + 8];
    unsigned short blist[KP_BLIST];
    float red[6][KP_WARPS];
    float qlo[3], qhi[3];
    float off[32];
    float cD[8];
    float g0[3], inv_cell, d0;
    int nc, nblist, npool, whole;
};

template <int T, bool BIG>
__global__ void __launch_bounds__(KP_THREADS)
knn_pool_build_kernel(dvcp_cloud_index_t index, const float *__restrict__ query, int64_t Q, int K, int gsz, float cell_in,
                      int nvalid, int pool_cap, float halo_scale, float4 *__restrict__ pools,
                      unsigned short *__restrict__ starts_g, KpMeta *__restrict__ metas) {
    constexpr int TT = T < 32 ? T : 32;
    constexpr int SB = (T + 31) / 32;
    constexpr int NBUCKETS = T * 32;
    __shared__ KpBuildShared S;
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cap = index.cap;
    const float *box = index.bucket_box + (int64_t)b * (cap / 32) * 8;
    const float4 *spt = reinterpret_cast<const float4 *>(index.sorted_pt) + (int64_t)b * cap;
    const int64_t q_begin = (int64_t)blockIdx.x * gsz;
    const int nq = (int)min((int64_t)gsz, Q - q_begin);
    const float *qg = query + ((int64_t)b * Q + q_begin) * 3;
    const int64_t gid = (int64_t)b * gridDim.x + blockIdx.x;
    float4 *pool = pools + gid * pool_cap;
    unsigned short *starts = starts_g + gid * KP_STARTS;
    KpMeta &M = metas[gid];

    // ---- query box ----
    {
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = tid; i < nq; i += KP_THREADS) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = __ldg(qg + 3 * i + c);
                mn[c] = fminf(mn[c], v);
                mx[c] = fmaxf(mx[c], v);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int s = 16; s; s >>= 1) {
                mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], s));
                mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], s));
            }
            if (lane == 0) {
                S.red[c][warp] = mn[c];
                S.red[3 + c][warp] = mx[c];
            }
        }
        __syncthreads();
        if (tid < 3) {
            float a = INFINITY, z = -INFINITY;
            for (int w = 0; w < KP_WARPS; ++w) {
                a = fminf(a, S.red[tid][w]);
                z = fmaxf(z, S.red[3 + tid][w]);
            }
            S.qlo[tid] = a;
            S.qhi[tid] = z;
        }
        __syncthreads();
    }
    const float qlx = S.qlo[0], qly = S.qlo[1], qlz = S.qlo[2], qhx = S.qhi[0], qhy = S.qhi[1], qhz = S.qhi[2];
    float cell = cell_in;
    const float ext3 = fmaxf(fmaxf(qhx - qlx, qhy - qly), qhz - qlz);
    {
        const float need = ext3 / (float)(KP3_MAXNC - 2 * KP_HALO - 2);
        if (!(cell > need)) cell = need * 1.0001f + 1e-20f;
    }
    KnnCtx ctx;
    ctx.box = box;
    ctx.spt = spt;
    ctx.buf = S.buf[warp];
    ctx.spt_lane = spt + lane;
    ctx.box_lane = reinterpret_cast<const float4 *>(box) + lane * 2;
    ctx.K = K;
    ctx.loose = INFINITY;
    Box6 sb[SB];
#pragma unroll
    for (int s = 0; s < SB; ++s) sb[s] = load_super_box(box, TT, s * 32 + lane);

    // ---- probes: the centre cold (best-first), then the corners bounded through it (1-Lipschitz) ----
    const bool probe = nvalid > pool_cap;
    if (probe) {
        const float mx_ = 0.5f * (qlx + qhx), my_ = 0.5f * (qly + qhy), mz_ = 0.5f * (qlz + qhz);
        if (warp == 0) {
            const unsigned long long r = knn_query_thr<T, BIG>(ctx, sb, mx_, my_, mz_, INFINITY, lane);
            const unsigned long long rk = __shfl_sync(0xffffffffu, r, K - 1);
            if (lane == 0) S.d0 = sqrtf(__uint_as_float(KnnKey<BIG>::d2bits(rk)));
        }
        __syncthreads();
        const float d0 = S.d0;
        for (int cn = warp; cn < 8; cn += KP_WARPS) {
            const float cx = (cn & 1) ? qhx : qlx, cy = (cn & 2) ? qhy : qly, cz = (cn & 4) ? qhz : qlz;
            const float hd = sqrtf((cx - mx_) * (cx - mx_) + (cy - my_) * (cy - my_) + (cz - mz_) * (cz - mz_));
            const float bd = fmaf(d0 + hd, 1.0001f, 1e-5f * (fabsf(cx) + fabsf(cy) + fabsf(cz)) + 1e-6f);
            const unsigned long long r = knn_query_thr<T, BIG>(ctx, sb, cx, cy, cz, bd * bd, lane);
            const unsigned long long rk = __shfl_sync(0xffffffffu, r, K - 1);
            if (lane == 0) S.cD[cn] = sqrtf(__uint_as_float(KnnKey<BIG>::d2bits(rk)));
        }
    }
    if (tid == 0) {
        int nc = (int)(ext3 / cell) + 1 + 2 * KP_HALO;
        if (nc > KP3_MAXNC) nc = KP3_MAXNC;
        S.nc = nc;
        S.g0[0] = qlx - KP_HALO * cell;
        S.g0[1] = qly - KP_HALO * cell;
        S.g0[2] = qlz - KP_HALO * cell;
        S.inv_cell = 1.0f / cell;
    }
    __syncthreads();
    const int nc = S.nc, ncell = nc * nc * nc;
    const float g0x = S.g0[0], g0y = S.g0[1], g0z = S.g0[2], inv_cell = S.inv_cell;
    auto cell_of = [&](const float4 &P) {
        return (kp_cell(P.x, g0x, inv_cell, nc) * nc + kp_cell(P.y, g0y, inv_cell, nc)) * nc + kp_cell(P.z, g0z, inv_cell, nc);
    };
    bool pool_ok = false;
    float lx = 0.f, ly = 0.f, lz = 0.f, hx = 0.f, hy = 0.f, hz = 0.f;
    for (int attempt = 0; attempt < 4 && !pool_ok; ++attempt) {
        __syncthreads();
        if (tid < 32) {
            float o = INFINITY;
            if (probe && tid < 26) {
                const float shrink[4] = {1.0f, 0.2f, 0.04f, 0.008f};
                const float scale = 1.0f + shrink[attempt] * (halo_scale - 1.0f);
                float vx, vy, vz;
                kp_dir(tid, vx, vy, vz);
                const float vn = sqrtf(vx * vx + vy * vy + vz * vz);
                o = -INFINITY;
                for (int cn = 0; cn < 8; ++cn) {
                    const float cx = (cn & 1) ? qhx : qlx, cy = (cn & 2) ? qhy : qly, cz = (cn & 4) ? qhz : qlz;
                    const float D = fmaf(S.cD[cn], scale, 1e-5f * (fabsf(cx) + fabsf(cy) + fabsf(cz)) + 1e-6f);
                    o = fmaxf(o, vx * cx + vy * cy + vz * cz + vn * D);
                }
            }
            S.off[tid] = o;
        }
        __syncthreads();
        lx = -S.off[4]; hx = S.off[21]; ly = -S.off[10]; hy = S.off[15]; lz = -S.off[12]; hz = S.off[13];
        for (int i = tid; i < ncell + 2; i += KP_THREADS) S.cur[i] = 0;
        if (tid == 0) {
            S.nblist = 0;
            S.npool = 0;
        }
        __syncthreads();
        for (int j = tid; j < NBUCKETS; j += KP_THREADS) {
            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(box + (int64_t)j * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4 *>(box + (int64_t)j * 8) + 1);
            if (b1.z > 0.f && b0.x <= hx && b0.w >= lx && b0.y <= hy && b1.x >= ly && b0.z <= hz && b1.y >= lz) {
                const int at = atomicAdd(&S.nblist, 1);
                if (at < KP_BLIST) S.blist[at] = (unsigned short)j;
            }
        }
        __syncthreads();
        const int nbl = S.nblist;
        if (nbl <= KP_BLIST) {
            int mine = 0;
            for (int i = warp; i < nbl; i += KP_WARPS) {
                const float4 P = __ldg(spt + (int)S.blist[i] * 32 + lane);
                const bool in = kp_inside(P, S.off);
                if (in) {
                    const int c = cell_of(P);
                    atomicAdd(reinterpret_cast<unsigned int *>(S.cur) + (c >> 1), 1u << (16 * (c & 1)));
                }
                mine += __popc(__ballot_sync(0xffffffffu, in));
            }
            if (lane == 0 && mine) atomicAdd(&S.npool, mine);
        }
        __syncthreads();
        pool_ok = nbl <= KP_BLIST && S.npool <= pool_cap;
        __syncthreads();
        if (!probe) break;
    }
    if (pool_ok) {
        const int nbl = S.nblist;
        if (warp == 0) {
            int run = 0;
            for (int base = 0; base < ncell; base += 32) {
                const int i = base + lane;
                const int v = i < ncell ? S.cur[i] : 0;
                int inc = v;
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, inc, s);
                    if (lane >= s) inc += t;
                }
                if (i < ncell) S.start[i] = (unsigned short)(run + inc - v);
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) {
                S.start[ncell] = (unsigned short)run;
                S.whole = run == nvalid;
            }
        }
        __syncthreads();
        for (int i = tid; i < ncell + 2; i += KP_THREADS) S.cur[i] = 0;
        __syncthreads();
        for (int i = warp; i < nbl; i += KP_WARPS) {
            const float4 P = __ldg(spt + (int)S.blist[i] * 32 + lane);
            if (kp_inside(P, S.off)) {
                const int c = cell_of(P);
                const unsigned old = atomicAdd(reinterpret_cast<unsigned int *>(S.cur) + (c >> 1), 1u << (16 * (c & 1)));
                pool[S.start[c] + ((old >> (16 * (c & 1))) & 0xffffu)] = P;
            }
        }
        for (int i = tid; i < ncell + 1; i += KP_THREADS) starts[i] = S.start[i];
    }
    // ---- meta data ----
    if (tid < 32) M.off[tid] = S.off[tid];
    if (tid < 8) {
        M.cx[tid] = (tid & 1) ? qhx : qlx;
        M.cy[tid] = (tid & 2) ? qhy : qly;
        M.cz[tid] = (tid & 4) ? qhz : qlz;
        M.cD[tid] = probe ? S.cD[tid] : INFINITY;
    }
    if (tid == 0) {
        M.g0x = g0x; M.g0y = g0y; M.g0z = g0z; M.inv_cell = inv_cell;
        M.nc = nc;
        M.npool = pool_ok ? S.npool : 0;
        M.whole = pool_ok && S.whole != 0;
        M.ok = pool_ok ? 1 : 0;
    }
}

constexpr int KQ_WARPS = 4;
constexpr int KQ_STACK = 8;    // qualifying points one lane holds before the warp folds them into its K best

// One warp per chain of z-lines. The candidates of a query are the pool points of the 3-D cells within sqrt(thr):
// the (x, y) columns of that window are dealt to the lanes, a lane walks the consecutive pool entries of its
// column's z-range (the pool is sorted by cell) and pushes the points with d2 <= thr onto its own little stack in
// shared memory; when a stack is full, and at the end, the stacks are compacted (one warp prefix sum), sorted in
// chunks of 32 and merged into the K best so far, whose K-th distance tightens thr for what is still to scan.
template <int T, bool BIG>
__global__ void __launch_bounds__(KQ_WARPS * 32, 8)
knn_pool_query_kernel(const float *__restrict__ query, int64_t Q, int K, int gsz, int zline, int chain_lines,
                      int groups_per_item, int pool_cap, const float4 *__restrict__ pools,
                      const unsigned short *__restrict__ starts_g, const KpMeta *__restrict__ metas,
                      float *__restrict__ dist, int64_t *__restrict__ idx64, int32_t *__restrict__ idx32,
                      unsigned long long *__restrict__ stats, KpWork *__restrict__ work,
                      unsigned int *__restrict__ work_count, int64_t nchains_total, int chains_per_group) {
    __shared__ unsigned long long s_stack[KQ_WARPS][KQ_STACK * 32];
    __shared__ unsigned long long s_comp[KQ_WARPS][KQ_STACK * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long INF = 0xffffffffffffffffull;
    unsigned long long *stack = s_stack[warp], *comp = s_comp[warp];
    unsigned st_cert = 0, st_unc = 0, st_ovf = 0, st_cold = 0, st_adm = 0, st_scan = 0;
    const int csz = chain_lines * zline;   // queries per chain
    for (int64_t ch = (int64_t)blockIdx.x * KQ_WARPS + warp; ch < nchains_total; ch += (int64_t)gridDim.x * KQ_WARPS) {
        const int gid = (int)(ch / chains_per_group);          // group over the whole batch
        const int cig = (int)(ch - (int64_t)gid * chains_per_group);
        const int b = gid / groups_per_item;
        const int64_t q_begin = (int64_t)(gid - b * groups_per_item) * gsz;
        const int nq = (int)min((int64_t)gsz, Q - q_begin);
        const int u0 = cig * csz, u1 = min(u0 + csz, nq);
        if (u0 >= nq) continue;
        const int64_t qbase = (int64_t)b * Q + q_begin;
        const float *qg = query + qbase * 3;
        const KpMeta &M = metas[gid];
        const float4 *pool = pools + (int64_t)gid * pool_cap;
        const unsigned short *starts = starts_g + (int64_t)gid * KP_STARTS;
        const int nc = M.nc;
        const bool usable = M.ok != 0 && M.npool >= K, whole = M.whole != 0;
        const float g0x = M.g0x, g0y = M.g0y, g0z = M.g0z, inv_cell = M.inv_cell;
        // this lane's pool plane (certificate) and corner probe (a bound that always holds)
        float pvx = 0.f, pvy = 0.f, pvz = 0.f, pinv = 0.f, poff = INFINITY;
        if (lane < 26) {
            kp_dir(lane, pvx, pvy, pvz);
            pinv = rsqrtf(pvx * pvx + pvy * pvy + pvz * pvz) * 0.99999f;
            poff = M.off[lane];
        }
        const float ccx = M.cx[lane & 7], ccy = M.cy[lane & 7], ccz = M.cz[lane & 7], ccD = M.cD[lane & 7];
        unsigned long long list = INF;
        int line_k = 0, line_dir = 1;   // position inside the current z-line, walking direction (boustrophedon)
        for (int qi = u0; qi < u1; ++qi) {
            const int line0 = qi - line_k;                    // first query of this line (walk order)
            const int line_len = min(zline, u1 - line0);
            const int ql = line_dir > 0 ? qi : line0 + (line_len - 1 - line_k);
            if (++line_k == line_len) {
                line_k = 0;
                line_dir = -line_dir;
            }
            const float qx = __ldg(qg + 3 * ql), qy = __ldg(qg + 3 * ql + 1), qz = __ldg(qg + 3 * ql + 2);
            float thr;
            {   // nearest probed corner: d_K is 1-Lipschitz (+inf when the pool holds the whole cloud unprobed)
                const float dq = sqrtf((ccx - qx) * (ccx - qx) + (ccy - qy) * (ccy - qy) + (ccz - qz) * (ccz - qz));
                float bsum = fmaf(ccD + dq, 1.0001f, 1e-5f * (fabsf(qx) + fabsf(qy) + fabsf(qz)) + 1e-6f);
                bsum = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(bsum)));
                thr = bsum * bsum;
            }
            unsigned long long res = INF;
            bool done = false;
            if (usable) {
                if (__any_sync(0xffffffffu, list != INF)) {   // the previous query's K neighbours, re-evaluated
                    float d2 = 0.f;
                    if (lane < K) {
                        const float4 P = __ldg(pool + ((unsigned)list & ((1u << KP_POSBITS) - 1u)));
                        d2 = sqdist_direct(P.x - qx, P.y - qy, P.z - qz);
                    }
                    thr = fminf(thr, __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(d2))));
                } else {
                    ++st_cold;
                }
                int cnt = 0;   // entries on my stack
                // fold the stacks into the K best so far; their K-th distance becomes the bound
                auto fold = [&]() {
                    int incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31), off = incl - cnt;
                    for (int j = 0; j < cnt; ++j) comp[off + j] = stack[j * 32 + lane];
                    __syncwarp();
                    for (int g = 0; g < total; g += 32) {
                        unsigned long long k = g + lane < total ? comp[g + lane] : INF;
                        k = bitonic_sort32(k, lane);
                        if (__all_sync(0xffffffffu, res == INF)) {
                            res = k;
                        } else {
                            const unsigned long long r = __shfl_sync(0xffffffffu, k, 31 - lane);
                            res = u64min(res, r);
#pragma unroll
                            for (int j = 16; j > 0; j >>= 1) res = cmpx64(res, j, (lane & j) == 0);
                        }
                    }
                    __syncwarp();
                    st_adm += total;
                    cnt = 0;
                    const unsigned long long w = __shfl_sync(0xffffffffu, res, K - 1);
                    if (w != INF) thr = fminf(thr, __uint_as_float((unsigned)(w >> 32)));
                };
                // window of cells
                int cx0 = 0, cx1 = nc - 1, cy0 = 0, cy1 = nc - 1, cz0 = 0, cz1 = nc - 1;
                if (thr < INFINITY) {
                    const float r0 = sqrtf(thr);
                    const float rho = fmaf(r0, 1.00002f, 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + r0) + 1e-30f);
                    cx0 = kp_cell(qx - rho, g0x, inv_cell, nc); cx1 = kp_cell(qx + rho, g0x, inv_cell, nc);
                    cy0 = kp_cell(qy - rho, g0y, inv_cell, nc); cy1 = kp_cell(qy + rho, g0y, inv_cell, nc);
                    cz0 = kp_cell(qz - rho, g0z, inv_cell, nc); cz1 = kp_cell(qz + rho, g0z, inv_cell, nc);
                }
                const int wy = cy1 - cy0 + 1, ncols = (cx1 - cx0 + 1) * wy;
                const float inv_wy = 1.0f / (float)wy;
                for (int c0 = 0; c0 < ncols; c0 += 32) {
                    const int c = c0 + lane;
                    int i = 0, i1 = 0;
                    if (c < ncols) {
                        const int cxi = (int)(((float)c + 0.5f) * inv_wy), cyi = c - cxi * wy;   // exact for c < 2^20
                        const int base = ((cx0 + cxi) * nc + (cy0 + cyi)) * nc;
                        i = __ldg(starts + base + cz0);
                        i1 = __ldg(starts + base + cz1 + 1);
                    }
                    st_scan += i1 - i;
                    // short runs (the usual case: a few points per column): every lane walks its own
#pragma unroll 1
                    for (int it = 0; it < 3 && __any_sync(0xffffffffu, i < i1); ++it) {
                        if (i < i1) {
                            const float4 P = __ldg(pool + i);
                            const float d2 = sqdist_direct(P.x - qx, P.y - qy, P.z - qz);
                            if (d2 <= thr) {
                                stack[cnt * 32 + lane] = kp_key(d2, __float_as_int(P.w), i);
                                ++cnt;
                            }
                            ++i;
                        }
                        if (__any_sync(0xffffffffu, cnt == KQ_STACK)) {
                            fold();
                            ++st_ovf;
                        }
                    }
                    // what is left of long runs (wide windows far from the cloud): the whole warp walks one run at a time
                    unsigned rem = __ballot_sync(0xffffffffu, i < i1);
                    while (rem) {
                        const int srcl = __ffs(rem) - 1;
                        rem &= rem - 1;
                        const int a = __shfl_sync(0xffffffffu, i, srcl), bnd = __shfl_sync(0xffffffffu, i1, srcl);
                        for (int j = a + lane; j < bnd + ((a - bnd) & 31); j += 32) {
                            if (j < bnd) {
                                const float4 P = __ldg(pool + j);
                                const float d2 = sqdist_direct(P.x - qx, P.y - qy, P.z - qz);
                                if (d2 <= thr) {
                                    stack[cnt * 32 + lane] = kp_key(d2, __float_as_int(P.w), j);
                                    ++cnt;
                                }
                            }
                            if (__any_sync(0xffffffffu, cnt == KQ_STACK)) {
                                fold();
                                ++st_ovf;
                            }
                        }
                    }
                }
                fold();
                if (lane >= K) res = INF;
                const unsigned long long wk = __shfl_sync(0xffffffffu, res, K - 1);
                if (wk != INF) {   // the pool had K points within the bound
                    const float kth = __uint_as_float((unsigned)(wk >> 32));
                    float m = (poff - (pvx * qx + pvy * qy + pvz * qz)) * pinv;
                    m = m - 4e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + fabsf(m));
                    if (!(poff < INFINITY)) m = INFINITY;
                    m = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(fmaxf(m, 0.f))));
                    done = whole || kth * 1.00001f < m * m;
                } else {
                    res = INF;
                }
            }
            if (done) {
                if (lane < K) {
                    const int64_t o = (qbase + ql) * K + lane;
                    dist[o] = __fsqrt_rn(__uint_as_float((unsigned)(res >> 32)));
                    const unsigned id = (unsigned)res >> KP_POSBITS;
                    if (idx64) idx64[o] = id;
                    if (idx32) idx32[o] = (int32_t)id;
                }
                list = res;
                ++st_cert;
            } else {
                ++st_unc;
                if (lane == 0) {   // thr: the best bound at hand (the pool's K-th distance if it had K points)
                    const unsigned at = atomicAdd(work_count, 1u);
                    work[at] = KpWork{qbase + ql, thr};
                }
                list = res;   // the pool's K best (if it had them) still seed the next query
                if (!__all_sync(0xffffffffu, lane >= K || list != INF)) list = INF;
            }
        }
    }
    st_scan = __reduce_add_sync(0xffffffffu, st_scan);
    kp_flush_stats(stats, lane, 1, 0, st_cert, st_unc, st_ovf, st_cold, 0u, st_adm, st_scan);
}

template <int T, bool BIG = false>
static int launch_knn_pool(dvcp_cloud_index_t index, const float *query, int B, int N, int64_t Q, int K, int gsz,
                           int zline, float cell, int pool_cap, float *dist, int64_t *idx64, int32_t *idx32,
                           unsigned long long *stats, void *workspace, cudaStream_t st) {
    static const float halo_scale = [] {   // DVCP_KNN_HALO: development override (allowance over the corners' K-th distance)
        const char *e = getenv("DVCP_KNN_HALO");
        return e ? (float)atof(e) : KP_HALO_SCALE;
    }();
    static const int chain_lines_env = [] {   // DVCP_KNN_CHAIN: z-lines per chain of the query kernel
        const char *e = getenv("DVCP_KNN_CHAIN");
        return e ? atoi(e) : 0;
    }();
    static const int mono = [] {   // DVCP_KNN_MONO=1: the monolithic shared-memory kernel
        const char *e = getenv("DVCP_KNN_MONO");
        return e ? atoi(e) : 0;
    }();
    const int64_t groups = (Q + gsz - 1) / gsz;
    unsigned char *ws = reinterpret_cast<unsigned char *>(workspace);
    unsigned int *work_count = reinterpret_cast<unsigned int *>(ws);
    KpWork *work = reinterpret_cast<KpWork *>(ws + 16);
    DVCP_CUDA(cudaMemsetAsync(work_count, 0, 16, st));
    if (mono) {
        int cap = pool_cap <= 0 ? KP_POOL_DEFAULT : pool_cap;
        cap = cap < 64 ? 64 : (cap > KP_POOL_MAX ? KP_POOL_MAX : (cap & ~31));
        auto k = knn_pool_kernel<T, BIG>;
        const int smem = (int)(sizeof(KpShared) + (size_t)cap * sizeof(float4));
        DVCP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        dim3 grid((unsigned)groups, B);
        k<<<grid, KP_THREADS, smem, st>>>(index, query, Q, K, gsz, zline, cell, N, cap, halo_scale, dist, idx64, idx32, stats,
                                          work, work_count);
        DVCP_CHECK_LAUNCH();
    } else {
        int cap = pool_cap <= 0 ? KP_POOL_MAX : pool_cap;
        cap = cap < 64 ? 64 : (cap > KP_POOL_MAX ? KP_POOL_MAX : (cap & ~31));
        size_t off = 16 + (size_t)B * Q * sizeof(KpWork);
        off = (off + 255) & ~(size_t)255;
        KpMeta *metas = reinterpret_cast<KpMeta *>(ws + off);
        off += (size_t)B * groups * sizeof(KpMeta);
        off = (off + 255) & ~(size_t)255;
        unsigned short *starts = reinterpret_cast<unsigned short *>(ws + off);
        off += (size_t)B * groups * KP_STARTS * sizeof(unsigned short);
        off = (off + 255) & ~(size_t)255;
        float4 *pools = reinterpret_cast<float4 *>(ws + off);
        dim3 bgrid((unsigned)groups, B);
        knn_pool_build_kernel<T, BIG><<<bgrid, KP_THREADS, 0, st>>>(index, query, Q, K, gsz, cell, N, cap, halo_scale, pools,
                                                                    starts, metas);
        DVCP_CHECK_LAUNCH();
        const int lines_per_group = (gsz + zline - 1) / zline;
        int chain_lines = chain_lines_env > 0 ? chain_lines_env : 1;
        if (chain_lines > lines_per_group) chain_lines = lines_per_group;
        const int chains_per_group = (lines_per_group + chain_lines - 1) / chain_lines;
        const int64_t nchains = (int64_t)B * groups * chains_per_group;
        int64_t qgrid = (nchains + KQ_WARPS - 1) / KQ_WARPS;
        if (qgrid > (int64_t)DVCP_NUM_SMS * 64) qgrid = (int64_t)DVCP_NUM_SMS * 64;
        knn_pool_query_kernel<T, BIG><<<(unsigned)qgrid, KQ_WARPS * 32, 0, st>>>(
            query, Q, K, gsz, zline, chain_lines, (int)groups, cap, pools, starts, metas, dist, idx64, idx32, stats, work,
            work_count, nchains, chains_per_group);
        DVCP_CHECK_LAUNCH();
    }
    knn_deferred_kernel<T, BIG><<<DVCP_NUM_SMS * 4, KNI_WARPS * 32, 0, st>>>(index, query, Q, K, work, work_count, dist, idx64, idx32);
    DVCP_CHECK_LAUNCH();
    return 0;
}

}  // namespace dvcp

using namespace dvcp;

extern "C" int64_t dvcp_knn_groups_workspace_bytes(int B, int64_t Q, int group) {
    if (B <= 0 || Q <= 0 || group <= 0) return 0;
    const int64_t groups = (Q + group - 1) / group;
    // work list (worst case: every query deferred) + per group: meta data, cell starts, pool of KP_POOL_MAX points
    return 16 + (int64_t)B * Q * (int64_t)sizeof(KpWork) + 3 * 256 +
           (int64_t)B * groups * ((int64_t)sizeof(KpMeta) + KP_STARTS * 2 + (int64_t)KP_POOL_MAX * 16);
}

extern "C" int dvcp_knn_groups(dvcp_cloud_index_t index, const float *query, int B, int N, int64_t Q, int K,
                               int group, int zline, float cell, int pool_cap, float *dist, int64_t *idx64,
                               int32_t *idx32, uint64_t *stats, void *workspace, int64_t workspace_bytes,
                               dvcp_stream_t stream) {
    if (!index.sorted_pt || !index.bucket_box || !query || !dist || (!idx64 && !idx32) || B <= 0 || N <= 0 ||
        Q <= 0 || group < 1 || zline < 1 || !(cell > 0.f))
        return DVCP_E_ARG;
    if (K < 1 || K > 32 || K > N || B > 65535 || index.cap < N || (Q + group - 1) / group > 0x7fffffff)
        return DVCP_E_UNSUPPORTED;
    if (!workspace) return DVCP_E_ARG;
    if (workspace_bytes < dvcp_knn_groups_workspace_bytes(B, Q, group)) return DVCP_E_WORKSPACE;
    if (zline > group) zline = group;
    cudaStream_t st = (cudaStream_t)stream;
    switch (index.cap / 1024) {
        case 1: return launch_knn_pool<1>(index, query, B, N, Q, K, group, zline, cell, pool_cap, dist, idx64, idx32, (unsigned long long *)stats, workspace, st);
        case 2: return launch_knn_pool<2>(index, query, B, N, Q, K, group, zline, cell, pool_cap, dist, idx64, idx32, (unsigned long long *)stats, workspace, st);
        case 4: return launch_knn_pool<4>(index, query, B, N, Q, K, group, zline, cell, pool_cap, dist, idx64, idx32, (unsigned long long *)stats, workspace, st);
        case 8: return launch_knn_pool<8>(index, query, B, N, Q, K, group, zline, cell, pool_cap, dist, idx64, idx32, (unsigned long long *)stats, workspace, st);
        case 16: return launch_knn_pool<16>(index, query, B, N, Q, K, group, zline, cell, pool_cap, dist, idx64, idx32, (unsigned long long *)stats, workspace, st);
        case 32: return launch_knn_pool<32>(index, query, B, N, Q, K, group, zline, cell, pool_cap, dist, idx64, idx32, (unsigned long long *)stats, workspace, st);
        case 64: return launch_knn_pool<64>(index, query, B, N, Q, K, group, zline, cell, pool_cap, dist, idx64, idx32, (unsigned long long *)stats, workspace, st);
    }
    return DVCP_E_UNSUPPORTED;   // 131072-point clouds: 17-bit indices do not fit the pool keys; use dvcp_knn_indexed
}
