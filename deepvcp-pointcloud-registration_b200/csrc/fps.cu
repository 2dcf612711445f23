// Farthest point sampling -- replaces pointnet2_utils.py:63-84 of the reference.
//
// Semantics: running minimum distance initialised to float32(1e10);
// per round d = ((dx*dx) + (dy*dy)) + (dz*dz) with every operation rounded (no
// FMA); dist[n] = d where d < dist[n]; next = FIRST index of the maximum.
//
// Two kernels:
//  * fps_bucketed_kernel  (float32, N <= 16384): exact spatially-pruned FPS. The
//    cloud is Hilbert-sorted into buckets of 32 points (one warp lane per point),
//    each bucket keeps its bounding box and the (max dist, lowest index) of its
//    members. A round only revisits buckets whose box lower bound -- evaluated
//    with the same rounded arithmetic, hence never above any member's rounded
//    distance -- is below the bucket's current maximum. Everything lives in
//    shared memory / registers of ONE CTA per cloud. The batched rounds accept
//    many exact picks per block-wide step (one exposed key per bucket and 32
//    candidates per step; at 16384 points the WIDE rounds: two exposed keys, 64
//    candidates). With `consume` the cloud's index is read instead of rebuilt: the
//    form dvcp_fps_indexed (concurrent = 2) uses for the throughput pipeline.
//  * fps_generic_kernel   (float32/float64, any N <= 57344): plain O(N * npoint).
#include <cstdlib>

#include <cub/block/block_radix_sort.cuh>

#include "common.cuh"

namespace dvcp {

// ------------------------------------------------------------------ generic --
template <typename T>
struct SqDist;
template <>
struct SqDist<float> {
    __device__ static __forceinline__ float eval(float dx, float dy, float dz) {
        return sq3_nofma(dx, dy, dz);
    }
};
template <>
struct SqDist<double> {
    __device__ static __forceinline__ double eval(double dx, double dy, double dz) {
        return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    }
};

template <typename T>
__global__ void __launch_bounds__(1024, 1)
fps_generic_kernel(const T *__restrict__ base, int64_t bs, int64_t ps, int64_t cs, int N, int npoint,
                   const int64_t *__restrict__ start, int64_t *__restrict__ out64,
                   int32_t *__restrict__ out32) {
    extern __shared__ float s_dist[];
    __shared__ unsigned s_hi[2][32], s_lo[2][32];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const T *p = base + (int64_t)b * bs;
    for (int n = tid; n < N; n += blockDim.x) s_dist[n] = 1e10f;
    unsigned far = (unsigned)start[b];
    __syncthreads();
    for (int i = 0; i < npoint; ++i) {
        if (tid == 0) {
            if (out64) out64[(int64_t)b * npoint + i] = far;
            if (out32) out32[(int64_t)b * npoint + i] = (int32_t)far;
        }
        const T cx = p[(int64_t)far * ps], cy = p[(int64_t)far * ps + cs], cz = p[(int64_t)far * ps + 2 * cs];
        unsigned hi = 0u, lo = 0u;
        for (int n = tid; n < N; n += blockDim.x) {
            const T x = p[(int64_t)n * ps], y = p[(int64_t)n * ps + cs], z = p[(int64_t)n * ps + 2 * cs];
            const T d = SqDist<T>::eval(x - cx, y - cy, z - cz);
            float cur = s_dist[n];
            if (d < (T)cur) {
                cur = (float)d;
                s_dist[n] = cur;
            }
            const unsigned bits = __float_as_uint(cur);
            const unsigned l = 0xffffffffu - (unsigned)n;
            if (bits > hi || (bits == hi && l > lo)) {
                hi = bits;
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
        far = 0xffffffffu - lo;
    }
}

// ----------------------------------------------------------------- bucketed --
__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// WARPS warps; each warp owns BPW buckets of 32 points: capacity WARPS * BPW * 32.
// Bucket j (32 consecutive points of the Hilbert order) belongs to warp j % WARPS,
// slot j / WARPS: spatial neighbours are spread over the warps, so the few
// buckets a round revisits are processed in parallel instead of by one warp.
#ifdef DVCP_FPS_TIMING
__device__ long long g_fb_stat[8];
#endif
template <int WARPS, int BPW, bool BATCHED, bool WIDE = false>
__global__ void __launch_bounds__(WARPS * 32, 1)
fps_bucketed_kernel(const float *__restrict__ base, int64_t bs, int64_t ps, int64_t cs, int N,
                    int npoint, const int64_t *__restrict__ start, int64_t *__restrict__ out64,
                    int32_t *__restrict__ out32, dvcp_cloud_index_t index, bool consume) {
    constexpr int THREADS = WARPS * 32;
    constexpr int CAP = WARPS * BPW * 32;
    constexpr int ITEMS = BPW;
    using Sort = cub::BlockRadixSort<unsigned, THREADS, ITEMS, unsigned>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sx = reinterpret_cast<float *>(smem_raw);
    float *sy = sx + CAP;
    float *sz = sy + CAP;
    unsigned short *sidx = reinterpret_cast<unsigned short *>(sz + CAP);
    __shared__ unsigned s_hi[2][WARPS], s_lo[2][WARPS];
    __shared__ float s_box[6][WARPS];
    __shared__ unsigned s_startpos;

    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float *p = base + (int64_t)b * bs;
    const unsigned startidx = start ? (unsigned)start[b] : 0u;

    if (consume) {
        // the cloud's index is already built (dvcp_build_index): read the Hilbert-ordered points instead of sorting again
        const float4 *ip = reinterpret_cast<const float4 *>(index.sorted_pt) + (int64_t)b * CAP;
        for (int i = tid; i < CAP; i += THREADS) {
            const float4 P = __ldg(ip + i);
            const int id = __float_as_int(P.w);
            const bool used = id >= 0;
            sx[i] = used ? P.x : 0.f;
            sy[i] = used ? P.y : 0.f;
            sz[i] = used ? P.z : 0.f;
            sidx[i] = used ? (unsigned short)id : (unsigned short)0xffffu;
            if (used && (unsigned)id == startidx) s_startpos = (unsigned)i;
        }
        __syncthreads();
    } else {
        // ---- cloud bounding box -> Hilbert keys -> block sort (prologue, once) ----
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int n = tid; n < N; n += THREADS) {
    #pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v = __ldg(p + (int64_t)n * ps + c * cs);
                mn[c] = fminf(mn[c], v);
                mx[c] = fmaxf(mx[c], v);
            }
        }
    #pragma unroll
        for (int c = 0; c < 3; ++c) {
            mn[c] = warp_min_f(mn[c]);
            mx[c] = warp_max_f(mx[c]);
            if (lane == 0) {
                s_box[c][warp] = mn[c];
                s_box[3 + c][warp] = mx[c];
            }
        }
        __syncthreads();
        float ext = 0.f;
    #pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a = INFINITY, z = -INFINITY;
            for (int w = 0; w < WARPS; ++w) {
                a = fminf(a, s_box[c][w]);
                z = fmaxf(z, s_box[3 + c][w]);
            }
            mn[c] = a;
            ext = fmaxf(ext, z - a);
        }
        // one scale for all axes: Hilbert cells are cubes, buckets stay compact
        const float scale = ext > 0.f ? 1023.0f / ext : 0.0f;
        unsigned keys[ITEMS], vals[ITEMS];
    #pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int n = tid * ITEMS + i;
            if (n < N) {
                unsigned q[3];
    #pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float v = (__ldg(p + (int64_t)n * ps + c * cs) - mn[c]) * scale;
                    q[c] = (unsigned)fminf(fmaxf(v, 0.0f), 1023.0f);
                }
                keys[i] = spatial_key(q[0], q[1], q[2]);
                vals[i] = (unsigned)n;
            } else {
                keys[i] = 0xffffffffu;   // sentinels sort last
                vals[i] = 0xffffu;
            }
        }
        __syncthreads();
        Sort(*reinterpret_cast<typename Sort::TempStorage *>(smem_raw)).Sort(keys, vals, 0, 32);
        __syncthreads();   // temp storage aliased with sx/sy/sz/sidx: done with it
        // thread `tid` now holds sorted positions tid*ITEMS .. tid*ITEMS+ITEMS-1
    #pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int pos = tid * ITEMS + i;
            const unsigned n = vals[i];
            float x = 0.f, y = 0.f, z = 0.f;
            if (n != 0xffffu) {
                x = __ldg(p + (int64_t)n * ps);
                y = __ldg(p + (int64_t)n * ps + cs);
                z = __ldg(p + (int64_t)n * ps + 2 * cs);
                if (n == startidx) s_startpos = (unsigned)pos;
            }
            sx[pos] = x;
            sy[pos] = y;
            sz[pos] = z;
            sidx[pos] = (unsigned short)n;
        }
        __syncthreads();
    }
    // optional: publish the Hilbert-ordered cloud as a spatial index (ball query / KNN reuse it)
    if (index.sorted_pt && !consume) {
        float4 *op = reinterpret_cast<float4 *>(index.sorted_pt) + (int64_t)b * CAP;
        for (int i = tid; i < CAP; i += THREADS) {
            const bool used = sidx[i] != 0xffffu;
            op[i] = used ? make_float4(sx[i], sy[i], sz[i], __int_as_float((int)sidx[i]))
                         : make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(-1));
        }
    }

    // ---- per-lane state: dist[k] = point `lane` of bucket (k*WARPS + warp);
    //      lane k additionally owns that bucket's box and (max dist, tie word). ----
    float dist[BPW];
    unsigned idp[(BPW + 1) / 2];   // original index of my point in bucket k, two 16-bit fields per word
    float bminx = 0.f, bminy = 0.f, bminz = 0.f, bmaxx = 0.f, bmaxy = 0.f, bmaxz = 0.f;
    unsigned bval = 0u, blo = 0u;     // best (largest) key of my bucket: dist bits, tie word
    unsigned sval = 0u, slo = 0u;     // second-best key of my bucket (batched kernel only)
    unsigned tval = 0u, tlo = 0u;     // third-best key of my bucket (WIDE only)
#pragma unroll
    for (int k = 0; k < (BPW + 1) / 2; ++k) idp[k] = 0u;
#pragma unroll
    for (int k = 0; k < BPW; ++k) {
        const int pos = (k * WARPS + warp) * 32 + lane;
        const unsigned id = sidx[pos];
        const bool valid = id != 0xffffu;
        idp[k >> 1] |= id << (16 * (k & 1));
        dist[k] = valid ? 1e10f : 0.0f;
        const float x = sx[pos], y = sy[pos], z = sz[pos];
        const float a0 = warp_min_f(valid ? x : INFINITY), a1 = warp_min_f(valid ? y : INFINITY),
                    a2 = warp_min_f(valid ? z : INFINITY);
        const float z0 = warp_max_f(valid ? x : -INFINITY), z1 = warp_max_f(valid ? y : -INFINITY),
                    z2 = warp_max_f(valid ? z : -INFINITY);
        const unsigned hi0 = __float_as_uint(dist[k]);
        const unsigned lo0 = ((0xffffu - id) << 16) | (unsigned)pos;
        unsigned hi = hi0, lo = lo0;
        warp_max_pair(hi, lo);
        const bool mine = hi0 == hi && lo0 == lo;
        unsigned h2 = mine ? 0u : hi0, l2 = mine ? 0u : lo0;
        warp_max_pair(h2, l2);
        unsigned h3 = 0u, l3 = 0u;
        if constexpr (WIDE) {
            const bool mine2 = !mine && hi0 == h2 && lo0 == l2;
            h3 = (mine || mine2) ? 0u : hi0;
            l3 = (mine || mine2) ? 0u : lo0;
            warp_max_pair(h3, l3);
        }
        if (lane == k) {
            bminx = a0; bminy = a1; bminz = a2;
            bmaxx = z0; bmaxy = z1; bmaxz = z2;
            bval = hi; blo = lo;
            sval = h2; slo = l2;
            tval = h3; tlo = l3;
        }
        if (index.bucket_box && !consume) {
            const int cntv = __popc(__ballot_sync(0xffffffffu, valid));
            if (lane == 0) {
                float4 *bb = reinterpret_cast<float4 *>(index.bucket_box + ((int64_t)b * (CAP / 32) + k * WARPS + warp) * 8);
                bb[0] = make_float4(a0, a1, a2, z0);
                bb[1] = make_float4(z1, z2, (float)cntv, 0.f);
            }
        }
    }
    if (npoint <= 0) return;

    if constexpr (!BATCHED) {
        // ---- sequential rounds: one centroid per block-wide argmax ----
        unsigned pos = s_startpos, idx = startidx;
        for (int i = 0; i < npoint; ++i) {
            if (tid == 0) {
                if (out64) out64[(int64_t)b * npoint + i] = idx;
                if (out32) out32[(int64_t)b * npoint + i] = (int32_t)idx;
            }
            const float cx = sx[pos], cy = sy[pos], cz = sz[pos];
            // lower bound of the rounded distance to any member of my bucket
            const float ex = fmaxf(fmaxf(__fsub_rn(bminx, cx), __fsub_rn(cx, bmaxx)), 0.0f);
            const float ey = fmaxf(fmaxf(__fsub_rn(bminy, cy), __fsub_rn(cy, bmaxy)), 0.0f);
            const float ez = fmaxf(fmaxf(__fsub_rn(bminz, cz), __fsub_rn(cz, bmaxz)), 0.0f);
            const float lb = sq3_nofma(ex, ey, ez);
            const unsigned mask = __ballot_sync(0xffffffffu, lane < BPW && lb < __uint_as_float(bval));
            if (mask) {
#pragma unroll
                for (int g = 0; g < (BPW + 7) / 8; ++g) {
                    if (!(mask & (0xffu << (8 * g)))) continue;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const int k = g * 8 + kk;
                        if (k >= BPW) break;
                        if (!(mask & (1u << k))) continue;
                        const int pp = (k * WARPS + warp) * 32 + lane;
                        const float d = sq3_nofma(__fsub_rn(sx[pp], cx), __fsub_rn(sy[pp], cy), __fsub_rn(sz[pp], cz));
                        if (d < dist[k]) dist[k] = d;
                        unsigned hi = __float_as_uint(dist[k]);
                        unsigned lo = ((0xffffu - ((idp[k >> 1] >> (16 * (k & 1))) & 0xffffu)) << 16) | (unsigned)pp;
                        warp_max_pair(hi, lo);
                        if (lane == k) {
                            bval = hi;
                            blo = lo;
                        }
                    }
                }
            }
            unsigned hi = lane < BPW ? bval : 0u, lo = lane < BPW ? blo : 0u;
            warp_max_pair(hi, lo);
            if (WARPS > 1) {
                if (lane == 0) {
                    s_hi[i & 1][warp] = hi;
                    s_lo[i & 1][warp] = lo;
                }
                __syncthreads();
                hi = lane < WARPS ? s_hi[i & 1][lane] : 0u;
                lo = lane < WARPS ? s_lo[i & 1][lane] : 0u;
                warp_max_pair(hi, lo);
            }
            pos = lo & 0xffffu;
            idx = 0xffffu - (lo >> 16);
        }
    } else if constexpr (WIDE) {
        // ---- batched rounds, TWO exposed keys per bucket and up to 64 candidates resolved per step ----
        // As below, generalised (the argument of fps_cluster.cu with E = 2): with S the largest THIRD-best key of any
        // bucket, every point whose key exceeds S is among the two best of its bucket, so the best and second-best
        // keys above S are ALL the points above S. About (6 B^2)^(1/3) ~ 100 candidates per step for B = 512 buckets
        // instead of ~ sqrt(pi B / 2) ~ 28 with one exposed key: half the steps. The pair tests of the candidate
        // list are spread over the whole CTA (thread = (candidate, 8 earlier candidates)); warp 0 only resolves.
        constexpr int WCAP = 64;
        __syncthreads();   // everyone has read sidx: its storage (CAP * 2 bytes >= 16 KB here) is reused
        static_assert(CAP * 2 >= 8192 + WCAP * (8 + 16 + 16 + 8 + 8), "scratch fits the index storage");
        unsigned long long *s_cand = reinterpret_cast<unsigned long long *>(sidx);         // [2 * CAP / 32] <= 1024 keys
        unsigned long long *s_top = s_cand + 1024;                                          // [WCAP] descending
        float4 *s_cxyz = reinterpret_cast<float4 *>(s_top + WCAP);                          // [WCAP] their coordinates
        float4 *s_acc = s_cxyz + WCAP;                                                      // [WCAP] accepted centroids
        unsigned char *s_K = reinterpret_cast<unsigned char *>(s_acc + WCAP);               // [WCAP][8]
        unsigned char *s_L = s_K + WCAP * 8;                                                // [WCAP][8]
        __shared__ unsigned long long s_T;
        __shared__ unsigned s_ncand, s_nacc;
        if (tid == 0) {
            const unsigned sp = s_startpos;
            s_acc[0] = make_float4(sx[sp], sy[sp], sz[sp], 0.f);
            s_nacc = 1u;
            s_ncand = 0u;
            if (out64) out64[(int64_t)b * npoint] = startidx;
            if (out32) out32[(int64_t)b * npoint] = (int32_t)startidx;
        }
        __syncthreads();
        int produced = 0;
#ifdef DVCP_FPS_TIMING
        long long t_last__ = clock64();
#define FB_TICK(slot) do { if (b == 0 && tid == 0) { const long long now__ = clock64(); g_fb_stat[slot] += now__ - t_last__; t_last__ = now__; } } while (0)
#else
#define FB_TICK(slot)
#endif
        while (true) {
            // ---- update: the A <= 64 centroids accepted in the previous step ----
            const int A = (int)s_nacc;
            unsigned long long F = 0ull;   // bit a: centroid a can lower a distance in my bucket
            {
                const float bestval = lane < BPW ? __uint_as_float(bval) : 0.0f;
                for (int a = 0; a < A; ++a) {
                    const float4 c = s_acc[a];
                    const float ex = fmaxf(fmaxf(__fsub_rn(bminx, c.x), __fsub_rn(c.x, bmaxx)), 0.0f);
                    const float ey = fmaxf(fmaxf(__fsub_rn(bminy, c.y), __fsub_rn(c.y, bmaxy)), 0.0f);
                    const float ez = fmaxf(fmaxf(__fsub_rn(bminz, c.z), __fsub_rn(c.z, bmaxz)), 0.0f);
                    F |= (unsigned long long)(sq3_nofma(ex, ey, ez) < bestval) << a;
                }
            }
#pragma unroll
            for (int k = 0; k < BPW; ++k) {
                unsigned long long Fk = __shfl_sync(0xffffffffu, F, k);
                if (Fk) {
                    const int pp = (k * WARPS + warp) * 32 + lane;
                    const float x = sx[pp], y = sy[pp], z = sz[pp];
                    float dk = dist[k];
                    do {
                        const int a = __ffsll((long long)Fk) - 1;
                        Fk &= Fk - 1;
                        const float4 c = s_acc[a];
                        const float d = sq3_nofma(__fsub_rn(x, c.x), __fsub_rn(y, c.y), __fsub_rn(z, c.z));
                        dk = d < dk ? d : dk;
                    } while (Fk);
                    // the box test lets a centroid in that lowers nothing in about every other case: the bucket's
                    // largest keys are then what they were
                    if (!__any_sync(0xffffffffu, dk < dist[k])) continue;
                    dist[k] = dk;
                    const unsigned hi0 = __float_as_uint(dk);
                    const unsigned lo0 = ((0xffffu - ((idp[k >> 1] >> (16 * (k & 1))) & 0xffffu)) << 16) | (unsigned)pp;
                    unsigned hi = hi0, lo = lo0;
                    warp_max_pair(hi, lo);
                    const bool mine = hi0 == hi && lo0 == lo;
                    unsigned h2 = mine ? 0u : hi0, l2 = mine ? 0u : lo0;
                    warp_max_pair(h2, l2);
                    const bool mine2 = !mine && hi0 == h2 && lo0 == l2;
                    unsigned h3 = (mine || mine2) ? 0u : hi0, l3 = (mine || mine2) ? 0u : lo0;
                    warp_max_pair(h3, l3);
                    if (lane == k) {
                        bval = hi; blo = lo;
                        sval = h2; slo = l2;
                        tval = h3; tlo = l3;
                    }
                }
            }
            produced += A;
            if (produced >= npoint) break;
            FB_TICK(4);
            // ---- S = largest third-best key over all buckets ----
            {
                unsigned hi = lane < BPW ? tval : 0u, lo = lane < BPW ? tlo : 0u;
                warp_max_pair(hi, lo);
                if (lane == 0) {
                    s_hi[0][warp] = hi;
                    s_lo[0][warp] = lo;
                }
            }
            __syncthreads();
            unsigned S_hi = lane < WARPS ? s_hi[0][lane] : 0u, S_lo = lane < WARPS ? s_lo[0][lane] : 0u;
            warp_max_pair(S_hi, S_lo);
            const unsigned long long S = ((unsigned long long)S_hi << 32) | S_lo;
            // ---- candidates: best and second-best keys above S (never a point of distance 0) ----
            {
                const unsigned long long key1 = ((unsigned long long)bval << 32) | blo;
                const unsigned long long key2 = ((unsigned long long)sval << 32) | slo;
                const bool c1 = lane < BPW && bval != 0u && key1 > S, c2 = lane < BPW && sval != 0u && key2 > S;
                const unsigned m1 = __ballot_sync(0xffffffffu, c1), m2 = __ballot_sync(0xffffffffu, c2);
                if (m1 | m2) {
                    unsigned basev = 0u;
                    if (lane == 0) basev = atomicAdd(&s_ncand, (unsigned)(__popc(m1) + __popc(m2)));
                    basev = __shfl_sync(0xffffffffu, basev, 0);
                    const unsigned below = (1u << lane) - 1u;
                    if (c1) s_cand[basev + __popc(m1 & below)] = key1;
                    if (c2) s_cand[basev + __popc(m1) + __popc(m2 & below)] = key2;
                }
            }
            __syncthreads();
            const int n = (int)s_ncand;
            if (n == 0) {
                // every remaining distance is 0: the argmax stays the lowest index (pointnet2_utils.py:83)
                unsigned hi = lane < BPW ? bval : 0u, lo = lane < BPW ? blo : 0u;
                warp_max_pair(hi, lo);
                if (lane == 0) {
                    s_hi[1][warp] = hi;
                    s_lo[1][warp] = lo;
                }
                __syncthreads();
                hi = lane < WARPS ? s_hi[1][lane] : 0u;
                lo = lane < WARPS ? s_lo[1][lane] : 0u;
                warp_max_pair(hi, lo);
                const unsigned idx = 0xffffu - (lo >> 16);
                for (int i = produced + tid; i < npoint; i += THREADS) {
                    if (out64) out64[(int64_t)b * npoint + i] = idx;
                    if (out32) out32[(int64_t)b * npoint + i] = (int32_t)idx;
                }
                break;
            }
            FB_TICK(5);
            // ---- rank the candidates (number of larger keys; keys are distinct): the WCAP largest go to s_top in
            //      descending order with their coordinates, the next one is the threshold T. P adjacent lanes share a key ----
            {
                const int P = n * 4 <= THREADS ? 4 : (n * 2 <= THREADS ? 2 : 1);
                const int shp = P == 4 ? 2 : (P == 2 ? 1 : 0);
                const int len = (n + P - 1) >> shp;
                for (int base = 0; base < n; base += THREADS >> shp) {   // warp-uniform trip count (usually one)
                    const int t = base + (tid >> shp), part = tid & (P - 1);
                    const unsigned long long key = t < n ? s_cand[t] : 0ull;
                    const int u0 = part * len, u1 = min(u0 + len, n);
                    int rk = 0;
                    if (t < n) {
                        int u = u0;
                        for (; u + 4 <= u1; u += 4) {
                            const unsigned long long a = s_cand[u], bb = s_cand[u + 1], c = s_cand[u + 2], d = s_cand[u + 3];
                            rk += (int)(a > key) + (int)(bb > key) + (int)(c > key) + (int)(d > key);
                        }
                        for (; u < u1; ++u) rk += s_cand[u] > key;
                    }
                    if (P >= 2) rk += __shfl_xor_sync(0xffffffffu, rk, 1);
                    if (P == 4) rk += __shfl_xor_sync(0xffffffffu, rk, 2);
                    if (t < n && part == 0) {
                        if (rk < WCAP) {
                            const unsigned pp = (unsigned)key & 0xffffu;
                            s_top[rk] = key;
                            s_cxyz[rk] = make_float4(sx[pp], sy[pp], sz[pp], 0.f);
                        }
                        if (rk == WCAP) s_T = key;
                    }
                }
            }
            __syncthreads();
            FB_TICK(6);
            const int m = n < WCAP ? n : WCAP;
            const unsigned long long T = n > WCAP ? s_T : S;
#ifdef DVCP_FPS_TIMING
            if (b == 0 && tid == 0) { g_fb_stat[0] += 1; g_fb_stat[1] += n; g_fb_stat[2] += A; }
#endif
            // ---- pair tests: item (row r, part p) covers the earlier candidates 8p .. 8p+7 of candidate r ----
            for (int it = tid; it < WCAP * 8; it += THREADS) {
                const int r = it & (WCAP - 1), part = it / WCAP;
                if (r < m) {
                    const unsigned long long key = s_top[r];
                    const float4 q = s_cxyz[r];
                    const unsigned klo = (unsigned)key;
                    const float dj = __uint_as_float((unsigned)(key >> 32));
                    unsigned kb = 0u, lb = 0u;
                    const int i0 = part * 8, i1 = r < i0 + 8 ? r : i0 + 8;
                    for (int i = i0; i < i1; ++i) {
                        const float4 c = s_cxyz[i];
                        const float d = sq3_nofma(__fsub_rn(q.x, c.x), __fsub_rn(q.y, c.y), __fsub_rn(q.z, c.z));
                        const bool kill = d < dj;
                        const unsigned long long nk = ((unsigned long long)__float_as_uint(d) << 32) | klo;
                        kb |= (unsigned)kill << (i & 7);
                        lb |= (unsigned)(kill && nk <= T) << (i & 7);
                    }
                    s_K[r * 8 + part] = (unsigned char)kb;
                    s_L[r * 8 + part] = (unsigned char)lb;
                }
            }
            __syncthreads();
            FB_TICK(7);
            if (warp == 0) {
                // lane l resolves rows l and l + 32 (row r = word r / 32, bit r % 32 of the masks)
                unsigned Kq[2][2], Lq[2][2];
                int st[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int r = q * 32 + lane;
                    st[q] = r < m ? 0 : 2;
                    const uint2 kw = *reinterpret_cast<const uint2 *>(s_K + r * 8), lw = *reinterpret_cast<const uint2 *>(s_L + r * 8);
                    Kq[q][0] = r < m ? kw.x : 0u;
                    Kq[q][1] = (r < m && q == 1) ? kw.y : 0u;
                    Lq[q][0] = r < m ? lw.x : 0u;
                    Lq[q][1] = (r < m && q == 1) ? lw.y : 0u;
                }
                unsigned acc[2] = {0u, 0u}, rej[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) rej[q] = __ballot_sync(0xffffffffu, st[q] == 2);
                // first-come resolution in key order: accepted unless an ACCEPTED earlier candidate lowers me
                while ((acc[0] | rej[0]) != 0xffffffffu || (acc[1] | rej[1]) != 0xffffffffu) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
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
                    for (int q = 0; q < 2; ++q) {
                        acc[q] = __ballot_sync(0xffffffffu, st[q] == 1);
                        rej[q] = __ballot_sync(0xffffffffu, st[q] == 2);
                    }
                }
                // a candidate lowered but still above T ends the step: keep what comes before the first such one
                bool stopped = false;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
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
                for (int q = 0; q < 2; ++q) {
                    const int r = q * 32 + lane;
                    const int rk = below + __popc(acc[q] & ((1u << lane) - 1u));
                    if (((acc[q] >> lane) & 1u) && rk < rem) {
                        s_acc[rk] = s_cxyz[r];
                        const unsigned idx = 0xffffu - ((unsigned)s_top[r] >> 16);
                        if (out64) out64[(int64_t)b * npoint + produced + rk] = idx;
                        if (out32) out32[(int64_t)b * npoint + produced + rk] = (int32_t)idx;
                    }
                    below += __popc(acc[q]);
                }
                if (lane == 0) {
                    s_nacc = (unsigned)(below < rem ? below : rem);
                    s_ncand = 0u;
                }
            }
            __syncthreads();
            FB_TICK(3);
        }
    } else {
        // ---- batched rounds (exact): several centroids per block-wide step ----
        // Order all points by the key (dist, lowest index first). Let S be the largest
        // SECOND-best key of any bucket: every point whose key exceeds S is the best of
        // its bucket, so the bucket maxima above S are ALL the points above S. Walking
        // those candidates in key order, a candidate whose distance no centroid accepted
        // in this step can lower (d >= its dist for each of them) is still the global
        // maximum when its turn comes, hence exactly the next FPS pick; a candidate that
        // is lowered to a key <= T (T = S, or the 33rd candidate key) drops behind every
        // remaining candidate and is skipped; a candidate lowered but still above T ends
        // the step. Then all accepted centroids update the buckets they can reach.
        __syncthreads();   // everyone has read sidx: its storage becomes the candidate list
        unsigned long long *s_cand = reinterpret_cast<unsigned long long *>(sidx);   // [CAP / 32]
        __shared__ unsigned long long s_top[32];
        __shared__ unsigned long long s_T;
        __shared__ float4 s_acc[32];
        __shared__ unsigned s_ncand, s_nacc;
        if (tid == 0) {
            const unsigned sp = s_startpos;
            s_acc[0] = make_float4(sx[sp], sy[sp], sz[sp], 0.f);
            s_nacc = 1u;
            s_ncand = 0u;
            if (out64) out64[(int64_t)b * npoint] = startidx;
            if (out32) out32[(int64_t)b * npoint] = (int32_t)startidx;
        }
        __syncthreads();
        int produced = 0;
        while (true) {
            // ---- update: the A centroids accepted in the previous step ----
            const int A = (int)s_nacc;
            unsigned F = 0u;   // bit a: centroid a can lower a distance in my bucket
            {
                const float bestval = lane < BPW ? __uint_as_float(bval) : 0.0f;
                for (int a = 0; a < A; ++a) {
                    const float4 c = s_acc[a];
                    const float ex = fmaxf(fmaxf(__fsub_rn(bminx, c.x), __fsub_rn(c.x, bmaxx)), 0.0f);
                    const float ey = fmaxf(fmaxf(__fsub_rn(bminy, c.y), __fsub_rn(c.y, bmaxy)), 0.0f);
                    const float ez = fmaxf(fmaxf(__fsub_rn(bminz, c.z), __fsub_rn(c.z, bmaxz)), 0.0f);
                    F |= (unsigned)(sq3_nofma(ex, ey, ez) < bestval) << a;
                }
            }
#pragma unroll
            for (int k = 0; k < BPW; ++k) {
                unsigned Fk = __shfl_sync(0xffffffffu, F, k);
                if (Fk) {
                    const int pp = (k * WARPS + warp) * 32 + lane;
                    const float x = sx[pp], y = sy[pp], z = sz[pp];
                    float dk = dist[k];
                    do {
                        const int a = __ffs(Fk) - 1;
                        Fk &= Fk - 1;
                        const float4 c = s_acc[a];
                        const float d = sq3_nofma(__fsub_rn(x, c.x), __fsub_rn(y, c.y), __fsub_rn(z, c.z));
                        dk = d < dk ? d : dk;
                    } while (Fk);
                    dist[k] = dk;
                    const unsigned hi0 = __float_as_uint(dk);
                    const unsigned lo0 = ((0xffffu - ((idp[k >> 1] >> (16 * (k & 1))) & 0xffffu)) << 16) | (unsigned)pp;
                    unsigned hi = hi0, lo = lo0;
                    warp_max_pair(hi, lo);
                    const bool mine = hi0 == hi && lo0 == lo;
                    unsigned h2 = mine ? 0u : hi0, l2 = mine ? 0u : lo0;
                    warp_max_pair(h2, l2);
                    if (lane == k) {
                        bval = hi; blo = lo;
                        sval = h2; slo = l2;
                    }
                }
            }
            produced += A;
            if (produced >= npoint) break;
            // ---- S = largest second-best key over all buckets ----
            {
                unsigned hi = lane < BPW ? sval : 0u, lo = lane < BPW ? slo : 0u;
                warp_max_pair(hi, lo);
                if (lane == 0) {
                    s_hi[0][warp] = hi;
                    s_lo[0][warp] = lo;
                }
            }
            __syncthreads();
            unsigned S_hi = lane < WARPS ? s_hi[0][lane] : 0u, S_lo = lane < WARPS ? s_lo[0][lane] : 0u;
            warp_max_pair(S_hi, S_lo);
            const unsigned long long S = ((unsigned long long)S_hi << 32) | S_lo;
            // ---- candidates: bucket maxima above S (never a point of distance 0) ----
            {
                const unsigned long long key = ((unsigned long long)bval << 32) | blo;
                const bool isc = lane < BPW && bval != 0u && key > S;
                const unsigned m = __ballot_sync(0xffffffffu, isc);
                if (m) {
                    unsigned basev = 0u;
                    if (lane == 0) basev = atomicAdd(&s_ncand, (unsigned)__popc(m));
                    basev = __shfl_sync(0xffffffffu, basev, 0);
                    if (isc) s_cand[basev + __popc(m & ((1u << lane) - 1u))] = key;
                }
            }
            __syncthreads();
            const int n = (int)s_ncand;
            if (n == 0) {
                // every remaining distance is 0: the argmax stays the lowest index (pointnet2_utils.py:83)
                unsigned hi = lane < BPW ? bval : 0u, lo = lane < BPW ? blo : 0u;
                warp_max_pair(hi, lo);
                if (lane == 0) {
                    s_hi[1][warp] = hi;
                    s_lo[1][warp] = lo;
                }
                __syncthreads();
                hi = lane < WARPS ? s_hi[1][lane] : 0u;
                lo = lane < WARPS ? s_lo[1][lane] : 0u;
                warp_max_pair(hi, lo);
                const unsigned idx = 0xffffu - (lo >> 16);
                for (int i = produced + tid; i < npoint; i += THREADS) {
                    if (out64) out64[(int64_t)b * npoint + i] = idx;
                    if (out32) out32[(int64_t)b * npoint + i] = (int32_t)idx;
                }
                break;
            }
            // ---- rank the candidates; the 32 largest go to s_top in descending order ----
            for (int t = tid; t < n; t += THREADS) {
                const unsigned long long key = s_cand[t];
                int rank = 0;
                for (int u = 0; u < n; ++u) rank += s_cand[u] > key;
                if (rank < 32) s_top[rank] = key;
                if (rank == 32) s_T = key;
            }
            __syncthreads();
            if (warp == 0) {
                const int m = n < 32 ? n : 32;
                const unsigned long long T = n > 32 ? s_T : S;
                const unsigned long long key = lane < m ? s_top[lane] : 0ull;
                const unsigned klo = (unsigned)(key & 0xffffffffu);
                const unsigned p = klo & 0xffffu;
                const float x = sx[p], y = sy[p], z = sz[p];
                const float dj = __uint_as_float((unsigned)(key >> 32));
                unsigned K = 0u, L = 0u;   // bit i: earlier candidate i lowers me / lowers me to a key <= T
                for (int i = 0; i < m; ++i) {
                    const float xi = __shfl_sync(0xffffffffu, x, i), yi = __shfl_sync(0xffffffffu, y, i),
                                zi = __shfl_sync(0xffffffffu, z, i);
                    const float d = sq3_nofma(__fsub_rn(x, xi), __fsub_rn(y, yi), __fsub_rn(z, zi));
                    const bool kill = i < lane && d < dj;
                    const unsigned long long nk = ((unsigned long long)__float_as_uint(d) << 32) | klo;
                    K |= (unsigned)kill << i;
                    L |= (unsigned)(kill && nk <= T) << i;
                }
                // first-come resolution in key order: accepted unless an ACCEPTED earlier candidate lowers me
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
                if (stopm) acc &= (1u << (__ffs(stopm) - 1)) - 1u;
                const int rank = __popc(acc & ((1u << lane) - 1u));
                const int rem = npoint - produced;
                if (((acc >> lane) & 1u) && rank < rem) {
                    s_acc[rank] = make_float4(x, y, z, 0.f);
                    const unsigned idx = 0xffffu - (klo >> 16);
                    if (out64) out64[(int64_t)b * npoint + produced + rank] = idx;
                    if (out32) out32[(int64_t)b * npoint + produced + rank] = (int32_t)idx;
                }
                if (lane == 0) {
                    const int na = __popc(acc);
                    s_nacc = (unsigned)(na < rem ? na : rem);
                    s_ncand = 0u;
                }
            }
            __syncthreads();
        }
    }
}

template <int WARPS, int BPW, bool BATCHED, bool WIDE = false>
static int launch_bucketed_impl(const dvcp_cloud_t &c, int B, int N, int npoint, const int64_t *start,
                           int64_t *o64, int32_t *o32, dvcp_cloud_index_t index, cudaStream_t st, bool consume) {
    constexpr int CAP = WARPS * BPW * 32;
    if (index.sorted_pt && (index.cap != CAP || !index.bucket_box)) return DVCP_E_ARG;
    using Sort = cub::BlockRadixSort<unsigned, WARPS * 32, BPW, unsigned>;
    size_t data = (size_t)CAP * (3 * sizeof(float) + sizeof(unsigned short));
    size_t smem = data > sizeof(typename Sort::TempStorage) ? data : sizeof(typename Sort::TempStorage);
    auto k = fps_bucketed_kernel<WARPS, BPW, BATCHED, WIDE>;
    DVCP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<B, WARPS * 32, smem, st>>>((const float *)c.base, c.bstride, c.pstride, c.cstride, N, npoint,
                                   start, o64, o32, index, consume);
    DVCP_CHECK_LAUNCH();
    return 0;
}

// DVCP_FPS_SEQUENTIAL=1 selects the one-centroid-per-round loop (same results; for timing comparisons).
static bool fps_sequential_mode() {
    static const bool v = [] {
        const char *e = getenv("DVCP_FPS_SEQUENTIAL");
        return e && e[0] == '1';
    }();
    return v;
}

template <int WARPS, int BPW>
static int launch_bucketed(const dvcp_cloud_t &c, int B, int N, int npoint, const int64_t *start,
                           int64_t *o64, int32_t *o32, dvcp_cloud_index_t index, cudaStream_t st, bool consume) {
    if (fps_sequential_mode()) return launch_bucketed_impl<WARPS, BPW, false>(c, B, N, npoint, start, o64, o32, index, st, consume);
    if constexpr (WARPS == 16 && BPW == 32) {   // 16384 points, 512 buckets: two exposed keys per bucket, 64 candidates per step
        return launch_bucketed_impl<WARPS, BPW, true, true>(c, B, N, npoint, start, o64, o32, index, st, consume);
    }
    else
        return launch_bucketed_impl<WARPS, BPW, true>(c, B, N, npoint, start, o64, o32, index, st, consume);
}

static int dispatch_bucketed(const dvcp_cloud_t &xyz, int B, int N, int npoint, const int64_t *start,
                             int64_t *out64, int32_t *out32, dvcp_cloud_index_t index, cudaStream_t st,
                             bool consume = false) {   // consume: `index` is already built and is only read
    if (N <= 1024) return launch_bucketed<4, 8>(xyz, B, N, npoint, start, out64, out32, index, st, consume);
    if (N <= 2048) return launch_bucketed<8, 8>(xyz, B, N, npoint, start, out64, out32, index, st, consume);
    if (N <= 4096) return launch_bucketed<16, 8>(xyz, B, N, npoint, start, out64, out32, index, st, consume);
    if (N <= 8192) return launch_bucketed<16, 16>(xyz, B, N, npoint, start, out64, out32, index, st, consume);
    return launch_bucketed<16, 32>(xyz, B, N, npoint, start, out64, out32, index, st, consume);
}

}  // namespace dvcp

extern "C" int dvcp_index_capacity(int N) {
    if (N < 64 || N > 16384) return 0;
    if (N <= 1024) return 1024;
    if (N <= 2048) return 2048;
    if (N <= 4096) return 4096;
    if (N <= 8192) return 8192;
    return 16384;
}

extern "C" int dvcp_build_index(dvcp_cloud_t xyz, int B, int N, dvcp_cloud_index_t index, dvcp_stream_t stream) {
    using namespace dvcp;
    if (!xyz.base || !index.sorted_pt || !index.bucket_box || B <= 0) return DVCP_E_ARG;
    if (dvcp_index_capacity(N) == 0) return DVCP_E_UNSUPPORTED;
    return dispatch_bucketed(xyz, B, N, 0, nullptr, nullptr, nullptr, index, (cudaStream_t)stream);
}

// fps_cluster.cu: one cloud per cluster of 8 CTAs, consuming an index that is already built.
int dvcp_fps_cluster_launch(dvcp_cloud_t xyz, dvcp_cloud_index_t index, int B, int N, int npoint, const int64_t *start,
                            int64_t *out64, int32_t *out32, int small_cta, cudaStream_t st);

// Few large clouds (the K8 batch: 16 clouds for 148 SMs): spread each over a cluster. Many clouds fill
// the GPU one CTA each, which does less total work. DVCP_FPS_CLUSTER=0/1 forces the choice.
static bool fps_use_cluster(int B, int N) {
    static const int forced = [] {
        const char *e = getenv("DVCP_FPS_CLUSTER");
        return e ? (e[0] == '1' ? 1 : 0) : -1;
    }();
    if (N <= 2048) return false;
    if (forced >= 0) return forced == 1;
    return B * 8 <= 2 * DVCP_NUM_SMS;
}

extern "C" int dvcp_fps(dvcp_cloud_t xyz, int dtype, int B, int N, int npoint, const int64_t *start,
                        int64_t *out64, int32_t *out32, dvcp_cloud_index_t index, dvcp_stream_t stream) {
    using namespace dvcp;
    if (!xyz.base || !start || (!out64 && !out32) || B <= 0 || N <= 0 || npoint <= 0) return DVCP_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == 0 && N <= 16384 && N >= 64 && index.sorted_pt && !fps_sequential_mode() && fps_use_cluster(B, N)) {
        const int rc = dispatch_bucketed(xyz, B, N, 0, nullptr, nullptr, nullptr, index, st);   // index only
        if (rc != 0) return rc;
        return dvcp_fps_cluster_launch(xyz, index, B, N, npoint, start, out64, out32, 0, st);
    }
    if (dtype == 0 && N <= 16384 && N >= 64) return dispatch_bucketed(xyz, B, N, npoint, start, out64, out32, index, st);
    if (index.sorted_pt) return DVCP_E_UNSUPPORTED;
    if (N > 57344) return DVCP_E_UNSUPPORTED;
    size_t smem = (size_t)N * sizeof(float);
    if (dtype == 0) {
        auto k = fps_generic_kernel<float>;
        DVCP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<B, 1024, smem, st>>>((const float *)xyz.base, xyz.bstride, xyz.pstride, xyz.cstride, N, npoint,
                                 start, out64, out32);
    } else if (dtype == 1) {
        auto k = fps_generic_kernel<double>;
        DVCP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<B, 1024, smem, st>>>((const double *)xyz.base, xyz.bstride, xyz.pstride, xyz.cstride, N, npoint,
                                 start, out64, out32);
    } else {
        return DVCP_E_ARG;
    }
    DVCP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dvcp_fps_indexed(dvcp_cloud_t xyz, int B, int N, int npoint, const int64_t *start, int64_t *out64,
                                int32_t *out32, dvcp_cloud_index_t index, int concurrent, dvcp_stream_t stream) {
    using namespace dvcp;
    if (!xyz.base || !start || (!out64 && !out32) || B <= 0 || N <= 0 || npoint <= 0 || !index.sorted_pt ||
        !index.bucket_box)
        return DVCP_E_ARG;
    if (index.cap != dvcp_index_capacity(N) || index.cap == 0) return DVCP_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (!fps_sequential_mode() && concurrent != 2 && fps_use_cluster(B, N))
        return dvcp_fps_cluster_launch(xyz, index, B, N, npoint, start, out64, out32, concurrent, st);
    // one CTA per cloud, reading the index that is already built (never rewriting it: others may be reading it)
    return dispatch_bucketed(xyz, B, N, npoint, start, out64, out32, index, st, true);
}

// Test hook: force the plain kernel for float32 clouds (parity of the two paths).
extern "C" int dvcp_fps_plain(dvcp_cloud_t xyz, int B, int N, int npoint, const int64_t *start,
                              int64_t *out64, dvcp_stream_t stream) {
    using namespace dvcp;
    if (!xyz.base || !start || !out64 || B <= 0 || N <= 0 || npoint <= 0) return DVCP_E_ARG;
    if (N > 57344) return DVCP_E_UNSUPPORTED;
    size_t smem = (size_t)N * sizeof(float);
    auto k = fps_generic_kernel<float>;
    DVCP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<B, 1024, smem, (cudaStream_t)stream>>>((const float *)xyz.base, xyz.bstride, xyz.pstride,
                                               xyz.cstride, N, npoint, start, out64, nullptr);
    DVCP_CHECK_LAUNCH();
    return 0;
}

#ifdef DVCP_FPS_TIMING
extern "C" __attribute__((visibility("default"))) int dvcp_debug_fps_bucketed_stat(long long *host8, int reset) {
    long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int rc = (int)cudaMemcpyFromSymbol(host8, dvcp::g_fb_stat, 8 * sizeof(long long));
    if (reset) rc |= (int)cudaMemcpyToSymbol(dvcp::g_fb_stat, z, 8 * sizeof(long long));
    return rc;
}
#endif
