/*
 * oracle/dvcp_oracle.c -- CPU restatement of the index-producing stages of the
 * DeepVCP registration hot path.
 *
 * THIS IS TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it. The product
 * path (deepvcp-pointcloud-registration_b200/) never does.
 *
 * Every function cites the reference lines whose arithmetic it restates
 * (paths relative to /root/reference). Floating-point contraction is OFF for
 * the whole file (-ffp-contract=off); fused multiply-adds appear only where
 * the reference's own arithmetic has them and are spelled fmaf().
 *
 * Build (oracle/Makefile):
 *   gcc -O2 -fopenmp -mfma -ffp-contract=off -shared -fPIC dvcp_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------
 * Farthest point sampling -- pointnet2_utils.py:63-84.
 *   distance[] starts at float32(1e10)                               (:74)
 *   d = ((dx*dx) + (dy*dy)) + (dz*dz), every op rounded, no FMA       (:80)
 *   distance[n] = d where d < distance[n]                             (:81-82)
 *   next = first index of the maximum of distance[]                   (:83)
 * xyz is one cloud [N,3]; start is the random first index the caller drew.
 * ------------------------------------------------------------------------- */
void orc_fps_f32(const float *xyz, int64_t N, int64_t npoint, int64_t start,
                 int64_t *out)
{
    float *dist = (float *)malloc(sizeof(float) * (size_t)N);
    for (int64_t n = 0; n < N; ++n) dist[n] = 1e10f;
    int64_t far = start;
    for (int64_t i = 0; i < npoint; ++i) {
        out[i] = far;
        const float cx = xyz[3 * far], cy = xyz[3 * far + 1], cz = xyz[3 * far + 2];
        float best = -INFINITY;
        int64_t besti = 0;
        for (int64_t n = 0; n < N; ++n) {
            const float dx = xyz[3 * n] - cx;
            const float dy = xyz[3 * n + 1] - cy;
            const float dz = xyz[3 * n + 2] - cz;
            const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
            const float s1 = xx + yy;
            const float d = s1 + zz;
            if (d < dist[n]) dist[n] = d;
            if (dist[n] > best) { best = dist[n]; besti = n; }
        }
        far = besti;
    }
    free(dist);
}

/* float64 clouds: the squared distance is formed in double, compared with the
 * float32 running minimum and stored rounded to float32
 * (pointnet2_utils.py:80-82, `dist[mask].float()`). */
void orc_fps_f64(const double *xyz, int64_t N, int64_t npoint, int64_t start,
                 int64_t *out)
{
    float *dist = (float *)malloc(sizeof(float) * (size_t)N);
    for (int64_t n = 0; n < N; ++n) dist[n] = 1e10f;
    int64_t far = start;
    for (int64_t i = 0; i < npoint; ++i) {
        out[i] = far;
        const double cx = xyz[3 * far], cy = xyz[3 * far + 1], cz = xyz[3 * far + 2];
        float best = -INFINITY;
        int64_t besti = 0;
        for (int64_t n = 0; n < N; ++n) {
            const double dx = xyz[3 * n] - cx;
            const double dy = xyz[3 * n + 1] - cy;
            const double dz = xyz[3 * n + 2] - cz;
            const double xx = dx * dx, yy = dy * dy, zz = dz * dz;
            const double s1 = xx + yy;
            const double d = s1 + zz;
            if (d < (double)dist[n]) dist[n] = (float)d;
            if (dist[n] > best) { best = dist[n]; besti = n; }
        }
        far = besti;
    }
    free(dist);
}

/* ---------------------------------------------------------------------------
 * square_distance -- pointnet2_utils.py:35-40, one (query, point)
 * entry: dot is a K=3 SGEMM accumulation (x first, FMA for y and z); the two
 * squared norms are sums of individually rounded squares; the three terms are
 * combined as ((-2*dot) + |q|^2) + |p|^2.
 * ------------------------------------------------------------------------- */
static inline float orc_norm2(const float *p)
{
    const float xx = p[0] * p[0], yy = p[1] * p[1], zz = p[2] * p[2];
    const float s1 = xx + yy;
    return s1 + zz;
}

static inline float orc_sqdist_expanded(const float *q, float qq, const float *p, float pp)
{
    const float m0 = q[0] * p[0];
    const float m1 = fmaf(q[1], p[1], m0);
    const float dot = fmaf(q[2], p[2], m1);
    const float a = -2.0f * dot;
    const float b = a + qq;
    return b + pp;
}

void orc_square_distance_f32(const float *src, int64_t S, const float *dst, int64_t N,
                             float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s) {
        const float qq = orc_norm2(src + 3 * s);
        for (int64_t n = 0; n < N; ++n)
            out[s * N + n] = orc_sqdist_expanded(src + 3 * s, qq, dst + 3 * n, orc_norm2(dst + 3 * n));
    }
}

/* ---------------------------------------------------------------------------
 * Ball query -- pointnet2_utils.py:87-107.
 * For each query the first `nsample` point indices (ascending) whose expanded
 * squared distance is NOT greater than float32(radius**2); short lists are
 * padded with their first entry; an empty list is all N (the reference then
 * fails in index_points).
 * ------------------------------------------------------------------------- */
void orc_ball_query_f32(const float *xyz, int64_t N, const float *new_xyz, int64_t S,
                        float radius2, int64_t nsample, int64_t *out)
{
    float *pp = (float *)malloc(sizeof(float) * (size_t)N);
    for (int64_t n = 0; n < N; ++n) pp[n] = orc_norm2(xyz + 3 * n);
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s) {
        const float *q = new_xyz + 3 * s;
        const float qq = orc_norm2(q);
        int64_t cnt = 0;
        int64_t *row = out + s * nsample;
        for (int64_t n = 0; n < N && cnt < nsample; ++n) {
            const float d = orc_sqdist_expanded(q, qq, xyz + 3 * n, pp[n]);
            if (!(d > radius2)) row[cnt++] = n;
        }
        const int64_t first = cnt ? row[0] : N;
        for (int64_t j = cnt; j < nsample; ++j) row[j] = first;
    }
    free(pp);
}

/* float64 clouds (the reference's loaders hand over float64: ModelNet40Dataset.py:38, KITTIDataset.py:84):
 * the same expressions evaluated in double -- torch's K=3 DGEMM accumulates like the SGEMM (x first, FMA for
 * y and z; probed against the reference on this host: 100 % bit-match) -- and the comparison is against the
 * Python double radius**2 (pointnet2_utils.py:102; no float32 cast happens for a float64 tensor). */
static inline double orc_norm2_d(const double *p)
{
    const double xx = p[0] * p[0], yy = p[1] * p[1], zz = p[2] * p[2];
    const double s1 = xx + yy;
    return s1 + zz;
}

static inline double orc_sqdist_expanded_d(const double *q, double qq, const double *p, double pp)
{
    const double m0 = q[0] * p[0];
    const double m1 = fma(q[1], p[1], m0);
    const double dot = fma(q[2], p[2], m1);
    const double a = -2.0 * dot;
    const double b = a + qq;
    return b + pp;
}

void orc_square_distance_f64(const double *src, int64_t S, const double *dst, int64_t N, double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s) {
        const double qq = orc_norm2_d(src + 3 * s);
        for (int64_t n = 0; n < N; ++n)
            out[s * N + n] = orc_sqdist_expanded_d(src + 3 * s, qq, dst + 3 * n, orc_norm2_d(dst + 3 * n));
    }
}

void orc_ball_query_f64(const double *xyz, int64_t N, const double *new_xyz, int64_t S,
                        double radius2, int64_t nsample, int64_t *out)
{
    double *pp = (double *)malloc(sizeof(double) * (size_t)N);
    for (int64_t n = 0; n < N; ++n) pp[n] = orc_norm2_d(xyz + 3 * n);
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; ++s) {
        const double *q = new_xyz + 3 * s;
        const double qq = orc_norm2_d(q);
        int64_t cnt = 0;
        int64_t *row = out + s * nsample;
        for (int64_t n = 0; n < N && cnt < nsample; ++n) {
            const double d = orc_sqdist_expanded_d(q, qq, xyz + 3 * n, pp[n]);
            if (!(d > radius2)) row[cnt++] = n;
        }
        if (cnt == 0) {
            for (int64_t j = 0; j < nsample; ++j) row[j] = N;
        } else {
            for (int64_t j = cnt; j < nsample; ++j) row[j] = row[0];
        }
    }
    free(pp);
}

/* ---------------------------------------------------------------------------
 * K nearest neighbours -- the contract of the third-party `knn_cuda` extension
 * at its call sites get_cat_feat_tgt.py:45,52 and deepVCP_loss.py:70,72
 * (the extension's source is not part of /root/reference, so this
 * restates its published algorithm: direct-form float32 squared distance with
 * contracted accumulation `ssd += d*d`, stable ascending order so the lower
 * index wins ties, square root of the kept distances, 0-based int64 indices).
 * ref [N,3], query [Q,3] -> dist [Q,K], idx [Q,K].
 * ------------------------------------------------------------------------- */
void orc_knn_f32(const float *ref, int64_t N, const float *query, int64_t Q, int64_t K,
                 float *dist, int64_t *idx)
{
#pragma omp parallel
    {
        float *bd = (float *)malloc(sizeof(float) * (size_t)K);
        int64_t *bi = (int64_t *)malloc(sizeof(int64_t) * (size_t)K);
#pragma omp for schedule(static)
        for (int64_t q = 0; q < Q; ++q) {
            const float qx = query[3 * q], qy = query[3 * q + 1], qz = query[3 * q + 2];
            int64_t cnt = 0;
            for (int64_t n = 0; n < N; ++n) {
                const float dx = ref[3 * n] - qx;
                const float dy = ref[3 * n + 1] - qy;
                const float dz = ref[3 * n + 2] - qz;
                float d = dx * dx;
                d = fmaf(dy, dy, d);
                d = fmaf(dz, dz, d);
                if (cnt == K && !(d < bd[K - 1])) continue; /* strict <: earlier index stays */
                int64_t j = cnt < K ? cnt : K - 1;
                while (j > 0 && d < bd[j - 1]) {
                    bd[j] = bd[j - 1];
                    bi[j] = bi[j - 1];
                    --j;
                }
                bd[j] = d;
                bi[j] = n;
                if (cnt < K) ++cnt;
            }
            for (int64_t j = 0; j < K; ++j) {
                dist[q * K + j] = j < cnt ? sqrtf(bd[j]) : INFINITY;
                idx[q * K + j] = j < cnt ? bi[j] : -1;
            }
        }
        free(bd);
        free(bi);
    }
}

/* ---------------------------------------------------------------------------
 * Candidate grid -- voxelize.py:19-29,44-83.
 * Per axis the lattice starts at (c - r) - s/2 in float64 and steps by s;
 * value i is float32(start + s*i) evaluated in float64; G points per axis;
 * candidate order (ix*G + iy)*G + iz; no sphere rejection (:75-77).
 * centres [M,3] float64 -> out [M, G^3, 3] float32.
 * ------------------------------------------------------------------------- */
void orc_candidates(const double *centres, int64_t M, double r, double s, int64_t G,
                    float *out)
{
    const double half = s / 2;
    for (int64_t m = 0; m < M; ++m) {
        double start[3];
        for (int a = 0; a < 3; ++a) {
            const double lo = centres[3 * m + a] - r;
            start[a] = lo - half;
        }
        float *o = out + m * G * G * G * 3;
        for (int64_t ix = 0; ix < G; ++ix)
            for (int64_t iy = 0; iy < G; ++iy)
                for (int64_t iz = 0; iz < G; ++iz) {
                    const double sx = s * (double)ix, sy = s * (double)iy, sz = s * (double)iz;
                    *o++ = (float)(start[0] + sx);
                    *o++ = (float)(start[1] + sy);
                    *o++ = (float)(start[2] + sz);
                }
    }
}

/* Number of lattice points per axis that torch.arange(lo - s/2, hi, s) yields
 * for lo = c - r, hi = c + r (voxelize.py:62-64): ceil((hi - start)/s) in f64. */
int64_t orc_grid_size(double c, double r, double s)
{
    const double lo = c - r, hi = c + r;
    const double start = lo - s / 2;
    return (int64_t)ceil((hi - start) / s);
}

/* ---------------------------------------------------------------------------
 * Voxel-grid filter (no counterpart in the reference's loader, which only has the
 * random down-sample of KITTIDataset.py:11-16): cell = floor((p - o) / cell) per axis in float32, one
 * output row per occupied cell in ascending (ix, iy, iz) order, centroid (mode 0: float64 sums in
 * ascending point index) or first point (mode 1). pts [M, stride] float32, the first `channels` are
 * reduced. Returns the number of cells; out [>= cells, channels], cnt [>= cells] (nullable).
 * ------------------------------------------------------------------------- */
typedef struct { unsigned long long key; int64_t idx; } orc_vk_t;
static int orc_vk_cmp(const void *a, const void *b)
{
    const orc_vk_t *x = (const orc_vk_t *)a, *y = (const orc_vk_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}
int64_t orc_voxel_filter(const float *pts, int64_t M, int64_t stride, int64_t channels, float ox, float oy, float oz,
                         float cell, int64_t mode, float *out, int32_t *cnt)
{
    orc_vk_t *v = (orc_vk_t *)malloc(sizeof(orc_vk_t) * (size_t)M);
    const float o[3] = {ox, oy, oz};
    for (int64_t i = 0; i < M; ++i) {
        unsigned long long k = 0;
        int ok = 1;
        for (int a = 0; a < 3; ++a) {
            const float d = pts[i * stride + a] - o[a];
            const float q = d / cell;
            const float f = floorf(q);
            if (!(fabsf(f) <= 1048575.0f)) ok = 0;
            else k = (k << 21) | (unsigned long long)((long long)f + 1048576ll);
        }
        v[i].key = ok ? k : 0xffffffffffffffffull;
        v[i].idx = i;
    }
    qsort(v, (size_t)M, sizeof(orc_vk_t), orc_vk_cmp);
    int64_t n = 0;
    for (int64_t i = 0; i < M;) {
        if (v[i].key == 0xffffffffffffffffull) break;
        int64_t j = i;
        double acc[4] = {0, 0, 0, 0};
        while (j < M && v[j].key == v[i].key) {
            if (mode == 0 || j == i)
                for (int64_t c = 0; c < channels; ++c) acc[c] += (double)pts[v[j].idx * stride + c];
            ++j;
        }
        const double div = mode == 0 ? (double)(j - i) : 1.0;
        for (int64_t c = 0; c < channels; ++c) out[n * channels + c] = (float)(acc[c] / div);
        if (cnt) cnt[n] = (int32_t)(j - i);
        ++n;
        i = j;
    }
    free(v);
    return n;
}
