// Device code shared by the indexed KNN kernels (knn.cu: warp-per-query search through the spatial
// index; knn_pool.cu: per-group shared-memory pools with the index search as the exact fallback).
// Contract: SURVEY A.5 / get_cat_feat_tgt.py:45,52 -- float32, d2 = fma(dz,dz, fma(dy,dy, dx*dx)),
// neighbours ordered by (d2, index).
#pragma once
#include "common.cuh"

namespace dvcp {

// ------------------------------------------------------ indexed KNN ---------
// Same contract, but the reference cloud comes with its spatial index (Hilbert
// buckets of 32 points + boxes). Per query:
//   * lb[j] = fma-chain squared distance from the query to bucket j's box,
//     evaluated on the clamped offsets with the SAME rounded arithmetic as the
//     point distances; by monotonicity of rounding no member of the bucket can
//     have a smaller rounded distance. A bucket is visited only if lb[j] <= the
//     current K-th best (ties included, so index tie-breaking is preserved).
//   * a warp walks a CHAIN of consecutive queries (one z-line of the candidate
//     lattice); the K-th distance of the previous query plus the step between the
//     two queries bounds the K-th distance of the next one (triangle inequality,
//     inflated for rounding), which prunes from the first bucket on.
//   * the K best are one 64-bit key per lane, sorted across the lanes; a visited
//     bucket (one point per lane) is merged by serial insertion when few points
//     qualify, else by a bitonic sort + merge.
constexpr int KNI_WARPS = 4;

__device__ __forceinline__ unsigned long long u64min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long u64max(unsigned long long a, unsigned long long b) { return a < b ? b : a; }

// compare-exchange with the partner lane: ONE 64-bit comparison decides (min and max formed separately cost
// two: 11 instructions per stage instead of 8, and the selection network is a quarter of this file's work)
__device__ __forceinline__ unsigned long long cmpx64(unsigned long long key, int j, bool keep_min) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, j);
    const bool lt = key < other;
    return (lt == keep_min) ? key : other;   // equal keys: either
}

__device__ __forceinline__ unsigned long long bitonic_sort32(unsigned long long key, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool up = (lane & k) == 0, lower = (lane & j) == 0;
            key = cmpx64(key, j, lower == up);
        }
    }
    return key;
}

// list: ascending across lanes (lanes >= K hold INF). keys: one candidate per lane (INF = none).
// Only the streaming fallback of the indexed kernel uses this.
__device__ __forceinline__ void knn_merge(unsigned long long &list, unsigned long long &worst,
                                          unsigned long long key, unsigned m, int K, int lane) {
    const unsigned long long INF = 0xffffffffffffffffull;
    if (__popc(m) <= 8) {
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const unsigned long long c = __shfl_sync(0xffffffffu, key, src);
            if (c < worst) {
                const int pos = __popc(__ballot_sync(0xffffffffu, list < c));
                const unsigned long long up = __shfl_up_sync(0xffffffffu, list, 1);
                if (lane < K) list = lane > pos ? up : (lane == pos ? c : list);
                worst = __shfl_sync(0xffffffffu, list, K - 1);
            }
        }
    } else {
        unsigned long long s = bitonic_sort32((m >> lane) & 1u ? key : INF, lane);
        const unsigned long long r = __shfl_sync(0xffffffffu, s, 31 - lane);
        list = u64min(list, r);   // the 32 smallest of the union, bitonic
#pragma unroll
        for (int j = 16; j > 0; j >>= 1) {
            list = cmpx64(list, j, (lane & j) == 0);
        }
        if (lane >= K) list = INF;
        worst = __shfl_sync(0xffffffffu, list, K - 1);
    }
}

// Key of a point for one query: (bits(d2), original index, slot in the sorted table) packed so that
// unsigned 64-bit order is (d2, index) order. d2 >= +0, so its sign bit is free:
//   caps <= 65536:  bits(d2) << 32 | index << 16 | slot          (16 + 16 bits)
//   cap  = 131072:  bits(d2) << 33 | index << 16 | slot >> 1     (17 + 16 bits; the slot is one of
//                   2h, 2h + 1 and is told apart by the index stored with the point)
template <bool BIG>
struct KnnKey {
    static constexpr int DSH = BIG ? 33 : 32;
    __device__ __forceinline__ static unsigned long long make(float d2, int id, int pos) {
        if constexpr (!BIG) {   // the low word is 32-bit arithmetic (one IMAD), the distance bits are the high word as they are
            const unsigned lo = ((unsigned)id << 16) | (unsigned)pos;
            return ((unsigned long long)__float_as_uint(d2) << 32) | lo;
        }
        return ((unsigned long long)__float_as_uint(d2) << DSH) | ((unsigned long long)(unsigned)id << 16) |
               (unsigned)(pos >> 1);
    }
    __device__ __forceinline__ static unsigned d2bits(unsigned long long k) { return (unsigned)(k >> DSH); }
    __device__ __forceinline__ static unsigned id(unsigned long long k) {
        return BIG ? ((unsigned)(k >> 16) & 0x1ffffu) : ((unsigned)(k & 0xffffffffu) >> 16);
    }
};

// Two-level pruning. Level 1 "super-buckets" of TT = min(T, 32) Hilbert-consecutive buckets; lane l owns
// the SB = max(T / 32, 1) super-buckets s * 32 + l (T = buckets / 32 = cap / 1024).
struct Box6 {
    float nx, ny, nz, xx, xy, xz;   // min, max
};
__device__ __forceinline__ Box6 load_super_box(const float *box, int TT, int g) {
    Box6 s{INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int t = 0; t < TT; ++t) {
        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(box + (int64_t)(g * TT + t) * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4 *>(box + (int64_t)(g * TT + t) * 8) + 1);
        if (b1.z > 0.f) {
            s.nx = fminf(s.nx, b0.x); s.ny = fminf(s.ny, b0.y); s.nz = fminf(s.nz, b0.z);
            s.xx = fmaxf(s.xx, b0.w); s.xy = fmaxf(s.xy, b1.x); s.xz = fmaxf(s.xz, b1.y);
        }
    }
    return s;
}
// fma-chain squared distance from q to a box: a lower bound of every member's rounded distance
__device__ __forceinline__ float box_lb(float qx, float qy, float qz, float nx, float ny, float nz, float xx,
                                        float xy, float xz) {
    const float ex = fmaxf(fmaxf(nx - qx, qx - xx), 0.f);
    const float ey = fmaxf(fmaxf(ny - qy, qy - xy), 0.f);
    const float ez = fmaxf(fmaxf(nz - qz, qz - xz), 0.f);
    return sqdist_direct(ex, ey, ez);   // empty box (min=+inf) -> +inf
}

constexpr int KNI_BUF = 128;   // qualifying points collected per query before the sort
constexpr float KNI_LOOSE = 3.0f;   // measured: K8 unchanged for >= 3, ModelNet-shaped batch 1.63 -> 1.39 ms

// What one warp needs to search one cloud.
struct KnnCtx {
    const float *box;          // bucket boxes of the cloud
    const float4 *spt;         // points in Hilbert order
    const float4 *spt_lane;    // spt + lane: slot `lane` of bucket j is spt_lane[j * 32] (one IMAD.WIDE per visit)
    const float4 *box_lane;    // box of bucket (g * TT + lane): box_lane[g * TT * 2]
    unsigned long long *buf;   // this warp's KNI_BUF-entry scratch list in shared memory
    int K;
    float loose;               // a chained bound above loose * (previous K-th squared distance) is not used
};

// key of the point at sorted slot `pos` for query (qx,qy,qz)
template <bool BIG>
__device__ __forceinline__ unsigned long long knn_point_key(const float4 *spt, int pos, float qx, float qy, float qz,
                                                            float &d2) {
    const float4 P = __ldg(spt + pos);   // unused slots hold +inf coordinates and id -1
    const int id = __float_as_int(P.w);
    d2 = sqdist_direct(P.x - qx, P.y - qy, P.z - qz);
    return id >= 0 ? KnnKey<BIG>::make(d2, id, pos) : 0xffffffffffffffffull;
}

// key of a point found for an earlier query (its key `old` tells where it is), re-evaluated for this query
template <bool BIG>
__device__ __forceinline__ unsigned long long knn_rekey(const float4 *spt, unsigned long long old, float qx, float qy,
                                                        float qz, float &d2) {
    if constexpr (!BIG) {
        return knn_point_key<false>(spt, (int)(old & 0xffffu), qx, qy, qz, d2);
    } else {
        const int h = (int)(old & 0xffffu) * 2;
        const float4 P0 = __ldg(spt + h), P1 = __ldg(spt + h + 1);
        const bool first = __float_as_int(P0.w) == (int)KnnKey<true>::id(old);
        const float4 P = first ? P0 : P1;
        d2 = sqdist_direct(P.x - qx, P.y - qy, P.z - qz);
        return KnnKey<true>::make(d2, __float_as_int(P.w), first ? h : h + 1);
    }
}

// One exact query by one warp. `list`: K points already known (the result of a nearby query; all-INF = none):
// re-evaluated for this query they bound its K-th distance. sb: this lane's super-boxes. Returns the K nearest
// keys ascending across the lanes (lanes >= K hold INF).
template <int T, bool BIG>
__device__ __forceinline__ unsigned long long knn_query_thr(const KnnCtx &c, const Box6 (&sb)[(T + 31) / 32], float qx,
                                                            float qy, float qz, float thr, int lane) {
    constexpr int TT = T < 32 ? T : 32;
    constexpr int SB = (T + 31) / 32;
    using Key = KnnKey<BIG>;
    const unsigned long long INF = 0xffffffffffffffffull;
    const float *box = c.box;
    const float4 *spt = c.spt;
    unsigned long long *buf = c.buf;
    const int K = c.K;
    auto worst_d2 = [&](unsigned long long w, float none) -> float {
        return w == INF ? none : __uint_as_float(Key::d2bits(w));
    };
    // lower bound of bucket `lane` of super-bucket g (lanes >= TT: +inf)
    auto bucket_lb = [&](int g) -> float {
        float l = INFINITY;
        if (lane < TT) {
            const float4 *bp = c.box_lane + g * (TT * 2);
            const float4 b0 = __ldg(bp);
            const float4 b1 = __ldg(bp + 1);
            if (b1.z > 0.f) l = box_lb(qx, qy, qz, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y);
        }
        return l;
    };
    float lbs[SB];
#pragma unroll
    for (int s = 0; s < SB; ++s) lbs[s] = box_lb(qx, qy, qz, sb[s].nx, sb[s].ny, sb[s].nz, sb[s].xx, sb[s].xy, sb[s].xz);
    // thr: an upper bound of the K-th squared distance (+inf = none: best-first streaming search)
    unsigned long long res = INF;
    if (!(thr < INFINITY)) {
        // ---- best-first over the super-buckets, streaming merge; exact for any start ----
        unsigned long long worst = INF;
        float rem[SB];   // this lane's super-bucket lower bounds; +inf once processed
#pragma unroll
        for (int s = 0; s < SB; ++s) rem[s] = lbs[s];
        while (true) {
            float rmin = rem[0];
#pragma unroll
            for (int s = 1; s < SB; ++s) rmin = fminf(rmin, rem[s]);
            const unsigned mb = __reduce_min_sync(0xffffffffu, __float_as_uint(rmin));
            if (mb == 0x7f800000u || !(__uint_as_float(mb) <= worst_d2(worst, INFINITY))) break;
            const int sl = __ffs(__ballot_sync(0xffffffffu, __float_as_uint(rmin) == mb)) - 1;
            int ss = 0;
#pragma unroll
            for (int s = SB - 1; s >= 0; --s)
                if (__float_as_uint(rem[s]) == mb) ss = s;
            ss = __shfl_sync(0xffffffffu, ss, sl);
#pragma unroll
            for (int s = 0; s < SB; ++s)
                if (lane == sl && s == ss) rem[s] = INFINITY;
            const int g = ss * 32 + sl;
            float l = bucket_lb(g);
            // the super-bucket's buckets in increasing order of their own bound
            while (true) {
                const unsigned lm = __reduce_min_sync(0xffffffffu, __float_as_uint(l));
                if (lm == 0x7f800000u || !(__uint_as_float(lm) <= worst_d2(worst, INFINITY))) break;
                const int bl = __ffs(__ballot_sync(0xffffffffu, __float_as_uint(l) == lm)) - 1;
                if (lane == bl) l = INFINITY;
                float d2;
                const unsigned long long key = knn_point_key<BIG>(spt, (g * TT + bl) * 32 + lane, qx, qy, qz, d2);
                const unsigned m = __ballot_sync(0xffffffffu, key < worst);
                if (m) knn_merge(res, worst, key, m, K, lane);
            }
        }
    } else {
    // ---- collect every point with d2 <= thr from the buckets whose box allows it ----
    int cnt = 0;
#pragma unroll
    for (int s = 0; s < SB; ++s) {
        unsigned sm = __ballot_sync(0xffffffffu, lbs[s] <= thr);
        while (sm) {
            const int g = s * 32 + __ffs(sm) - 1;
            sm &= sm - 1;
            const float l = bucket_lb(g);
            unsigned bm = __ballot_sync(0xffffffffu, l <= thr);
            while (bm) {
                const int j = g * TT + __ffs(bm) - 1;
                bm &= bm - 1;
                const int pos = j * 32 + lane;
                const float4 P = __ldg(c.spt_lane + j * 32);   // unused slots: +inf coordinates, never within thr
                const float d2 = sqdist_direct(P.x - qx, P.y - qy, P.z - qz);
                const bool qual = d2 <= thr;
                const unsigned m = __ballot_sync(0xffffffffu, qual);
                if (m) {
                    const int slot = cnt + __popc(m & ((1u << lane) - 1u));
                    if (qual && slot < KNI_BUF) buf[slot] = Key::make(d2, __float_as_int(P.w), pos);
                    cnt += __popc(m);
                }
            }
        }
    }
    __syncwarp();
    // ---- the K smallest keys, ascending across the lanes ----
    if (cnt <= KNI_BUF) {
        for (int g = 0; g < cnt; g += 32) {
            unsigned long long k = g + lane < cnt ? buf[g + lane] : INF;
            k = bitonic_sort32(k, lane);
            if (g == 0) {
                res = k;
            } else {
                const unsigned long long r = __shfl_sync(0xffffffffu, k, 31 - lane);
                res = u64min(res, r);
#pragma unroll
                for (int j = 16; j > 0; j >>= 1) {
                    res = cmpx64(res, j, (lane & j) == 0);
                }
            }
        }
    } else {
        // too many qualifying points for the buffer (loose bound): streaming merge over the same buckets
        unsigned long long worst = INF;
#pragma unroll
        for (int s = 0; s < SB; ++s) {
            unsigned sm2 = __ballot_sync(0xffffffffu, lbs[s] <= thr);
            while (sm2) {
                const int g = s * 32 + __ffs(sm2) - 1;
                sm2 &= sm2 - 1;
                for (int t = 0; t < TT; ++t) {
                    const int j = g * TT + t;
                    const float4 b0 = __ldg(reinterpret_cast<const float4 *>(box + (int64_t)j * 8));
                    const float4 b1 = __ldg(reinterpret_cast<const float4 *>(box + (int64_t)j * 8) + 1);
                    if (!(b1.z > 0.f)) continue;
                    const float l = box_lb(qx, qy, qz, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y);
                    const float now = fminf(thr, worst_d2(worst, thr));
                    if (!(l <= now)) continue;
                    float d2;
                    const unsigned long long key = knn_point_key<BIG>(spt, j * 32 + lane, qx, qy, qz, d2);
                    const unsigned m = __ballot_sync(0xffffffffu, key < worst && d2 <= thr);
                    if (m) knn_merge(res, worst, key, m, K, lane);
                }
            }
        }
    }
    }
    __syncwarp();
    if (lane >= K) res = INF;
    return res;
}


// Bound on the K-th squared distance of a query from `list`, the K results of a nearby query (all-INF = none):
// re-evaluated for this query they are K known points. +inf = no usable bound.
template <bool BIG>
__device__ __forceinline__ float knn_chain_bound(const KnnCtx &c, float qx, float qy, float qz,
                                                 unsigned long long list, int lane) {
    using Key = KnnKey<BIG>;
    const unsigned long long INF = 0xffffffffffffffffull;
    const int K = c.K;
    const float4 *spt = c.spt;
    float thr;
    if (__any_sync(0xffffffffu, list != INF)) {
        // previous neighbours re-evaluated for this query
        float d2 = 0.f;
        const unsigned long long k0 = (lane < K && list != INF) ? knn_rekey<BIG>(spt, list, qx, qy, qz, d2) : 0ull;
        thr = __uint_as_float(__reduce_max_sync(0xffffffffu, Key::d2bits(k0)));
        // queries further apart than their neighbourhoods are wide (ModelNet-shaped clouds, the dense core of a
        // scan): the chained bound admits many times K points and the best-first search below is cheaper
        const float prevk = __uint_as_float(Key::d2bits(__shfl_sync(0xffffffffu, list, K - 1)));
        if (thr > c.loose * prevk) thr = INFINITY;
    } else {
        thr = INFINITY;   // first query of a chain: best-first streaming search below
    }
    return thr;
}

// One exact query by one warp, chained to the previous query's result.
template <int T, bool BIG>
__device__ __forceinline__ unsigned long long knn_query(const KnnCtx &c, const Box6 (&sb)[(T + 31) / 32], float qx,
                                                        float qy, float qz, unsigned long long list, int lane) {
    return knn_query_thr<T, BIG>(c, sb, qx, qy, qz, knn_chain_bound<BIG>(c, qx, qy, qz, list, lane), lane);
}

}  // namespace dvcp
