// Data ingest for KITTI-shaped scans: the step immediately before the hot path.
//
// Reference: KITTIDataset.py:11-16 (downsample: N rows of the raw [M, 4] scan, chosen by the caller's
// np.random.choice), :44-46 (split into xyz and reflectance), :67-84 (target = R @ src + t in float64) and
// the [B, 3, N] channel-major layout the model takes (deepVCP.py:24). One launch for B scans: raw rows are
// gathered straight into the model's layout (coalesced writes), the target cloud is produced in the same
// pass, nothing goes through host memory.
#include <cub/cub.cuh>

#include "common.cuh"

namespace dvcp {

__global__ void __launch_bounds__(256)
ingest_kitti_kernel(const float4 *__restrict__ raw, const int64_t *__restrict__ scan_offset,
                    const int64_t *__restrict__ idx, const double *__restrict__ R, const double *__restrict__ t, int N,
                    float *__restrict__ src, float *__restrict__ tgt, float *__restrict__ refl) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int64_t lo = scan_offset[b], rows = scan_offset[b + 1] - lo;
    const int64_t i = idx ? idx[(int64_t)b * N + n] : n;
    float4 p = make_float4(NAN, NAN, NAN, NAN);   // an index outside the scan poisons the point instead of reading wild
    if (i >= 0 && i < rows) p = __ldg(raw + lo + i);
    float *s = src + (int64_t)b * 3 * N;
    s[n] = p.x;
    s[N + n] = p.y;
    s[2 * N + n] = p.z;
    if (refl) refl[(int64_t)b * N + n] = p.w;
    if (tgt) {
        const double *Rb = R + (int64_t)b * 9, *tb = t + (int64_t)b * 3;
        float *o = tgt + (int64_t)b * 3 * N;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double v = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(Rb[3 * r], (double)p.x), __dmul_rn(Rb[3 * r + 1], (double)p.y)),
                                                 __dmul_rn(Rb[3 * r + 2], (double)p.z)), tb[r]);
            o[r * N + n] = (float)v;
        }
    }
}

// point-major float4 copy (x, y, z, 0) of a cloud given in any layout: one 16-byte load per gathered
// neighbour in the embedding kernel
__global__ void __launch_bounds__(256)
pack_xyz4_kernel(Cloud c, int N, float4 *__restrict__ out) {
    const int b = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) out[(int64_t)b * N + n] = make_float4(c.at(b, n, 0), c.at(b, n, 1), c.at(b, n, 2), 0.f);
}


// ModelNet-shaped pair (ModelNet40Dataset.py:38-41,62-92): rows (x, y, z, nx, ny, nz) in float64 as np.loadtxt
// returns them -> source [B,6,N] and target [B,6,N] = (R xyz + t, R normals), channel-major, in float64 (what
// the reference's loader yields) or float32 (what the fast kernels take). Products and sums in float64,
// left to right, not contracted.
template <typename OUT>
__global__ void __launch_bounds__(256)
ingest_modelnet_kernel(const double *__restrict__ raw, const double *__restrict__ R, const double *__restrict__ t, int M,
                       int N, OUT *__restrict__ src, OUT *__restrict__ tgt) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double *row = raw + ((int64_t)b * M + n) * 6;
    double v[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) v[c] = row[c];
    OUT *s = src + (int64_t)b * 6 * N;
#pragma unroll
    for (int c = 0; c < 6; ++c) s[(int64_t)c * N + n] = (OUT)v[c];
    if (tgt) {
        const double *Rb = R + (int64_t)b * 9, *tb = t + (int64_t)b * 3;
        OUT *o = tgt + (int64_t)b * 6 * N;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double p = __dadd_rn(__dadd_rn(__dmul_rn(Rb[3 * r], v[0]), __dmul_rn(Rb[3 * r + 1], v[1])), __dmul_rn(Rb[3 * r + 2], v[2]));
            const double q = __dadd_rn(__dadd_rn(__dmul_rn(Rb[3 * r], v[3]), __dmul_rn(Rb[3 * r + 1], v[4])), __dmul_rn(Rb[3 * r + 2], v[5]));
            o[(int64_t)r * N + n] = (OUT)__dadd_rn(p, tb[r]);
            o[(int64_t)(3 + r) * N + n] = (OUT)q;
        }
    }
}

// ---- voxel-grid filter: one output point per occupied cell of a cubic lattice ----
// key of point i: cell index (floor((p - origin) / cell), float32 arithmetic) packed 21 bits per axis.
__global__ void __launch_bounds__(256)
voxel_keys_kernel(const float *__restrict__ pts, int stride, int64_t M, float ox, float oy, float oz, float cell,
                  unsigned long long *__restrict__ keys, unsigned *__restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const float *p = pts + i * stride;
    const float fx = floorf(__fdiv_rn(__fsub_rn(p[0], ox), cell)), fy = floorf(__fdiv_rn(__fsub_rn(p[1], oy), cell)),
                fz = floorf(__fdiv_rn(__fsub_rn(p[2], oz), cell));
    const float lim = 1048575.0f;   // 2^20 - 1: cells beyond +-2^20 of the origin (or NaN) share the last key
    const bool ok = fabsf(fx) <= lim && fabsf(fy) <= lim && fabsf(fz) <= lim;
    unsigned long long k = 0xffffffffffffffffull;
    if (ok) {
        const unsigned long long ix = (unsigned long long)((long long)fx + 1048576ll), iy = (unsigned long long)((long long)fy + 1048576ll),
                                 iz = (unsigned long long)((long long)fz + 1048576ll);
        k = (ix << 42) | (iy << 21) | iz;
    }
    keys[i] = k;
    vals[i] = (unsigned)i;
}
__global__ void __launch_bounds__(256)
voxel_heads_kernel(const unsigned long long *__restrict__ keys, int64_t M, unsigned *__restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const unsigned long long k = keys[i];
    head[i] = (k != 0xffffffffffffffffull && (i == 0 || keys[i - 1] != k)) ? 1u : 0u;
}
// thread per segment head: the points of a cell in ascending original index (the sort is stable), summed
// sequentially in float64 -> centroid (mode 0) or the first point (mode 1); remaining channels likewise
__global__ void __launch_bounds__(256)
voxel_reduce_kernel(const float *__restrict__ pts, int stride, int channels, int64_t M,
                    const unsigned long long *__restrict__ keys, const unsigned *__restrict__ vals,
                    const unsigned *__restrict__ head, const unsigned *__restrict__ slot, int mode, int64_t cap,
                    float *__restrict__ out, int32_t *__restrict__ out_count, int64_t *__restrict__ n_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        // cells found = inclusive scan at the last element
        *n_out = (int64_t)slot[M - 1] + (int64_t)head[M - 1];
    }
    if (i >= M || !head[i]) return;
    const int64_t o = slot[i];
    if (o >= cap) return;
    const unsigned long long k = keys[i];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int cnt = 0;
    for (int64_t j = i; j < M && keys[j] == k; ++j) {
        const float *p = pts + (int64_t)vals[j] * stride;
        if (mode == 0 || cnt == 0)
            for (int c = 0; c < channels; ++c) acc[c] = __dadd_rn(acc[c], (double)p[c]);
        ++cnt;
    }
    const double div = mode == 0 ? (double)cnt : 1.0;
    for (int c = 0; c < channels; ++c) out[o * channels + c] = (float)__ddiv_rn(acc[c], div);
    if (out_count) out_count[o] = cnt;
}
}  // namespace dvcp

using namespace dvcp;

extern "C" int dvcp_pack_xyz4(dvcp_cloud_t xyz, int B, int N, float *out, dvcp_stream_t stream) {
    if (!xyz.base || !out || B <= 0 || N <= 0) return DVCP_E_ARG;
    if (B > 65535 || ((uintptr_t)out & 15)) return DVCP_E_UNSUPPORTED;
    pack_xyz4_kernel<<<dim3((N + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(as_cloud(xyz), N, reinterpret_cast<float4 *>(out));
    DVCP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dvcp_ingest_kitti(const float *raw, const int64_t *scan_offset, const int64_t *idx, const double *R,
                                 const double *t, int B, int N, float *src, float *tgt, float *reflectance,
                                 dvcp_stream_t stream) {
    if (!raw || !scan_offset || !src || B <= 0 || N <= 0) return DVCP_E_ARG;
    if (tgt && (!R || !t)) return DVCP_E_ARG;
    if (B > 65535 || ((uintptr_t)raw & 15)) return DVCP_E_UNSUPPORTED;
    ingest_kitti_kernel<<<dim3((N + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(raw), scan_offset, idx, R, t, N, src, tgt, reflectance);
    DVCP_CHECK_LAUNCH();
    return 0;
}


extern "C" int dvcp_ingest_modelnet(const double *raw, const double *R, const double *t, int B, int M, int N, int out_f64,
                                    void *src, void *tgt, dvcp_stream_t stream) {
    if (!raw || !src || B <= 0 || N <= 0 || M < N) return DVCP_E_ARG;
    if (tgt && (!R || !t)) return DVCP_E_ARG;
    if (B > 65535) return DVCP_E_UNSUPPORTED;
    const dim3 grid((N + 255) / 256, B);
    if (out_f64)
        ingest_modelnet_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(raw, R, t, M, N, (double *)src, (double *)tgt);
    else
        ingest_modelnet_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(raw, R, t, M, N, (float *)src, (float *)tgt);
    DVCP_CHECK_LAUNCH();
    return 0;
}

// workspace layout: keys_in | keys_out (8 M each) | vals_in | vals_out | head | slot (4 M each) | cub temp
static size_t voxel_cub_bytes(int64_t M) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const unsigned *)nullptr, (unsigned *)nullptr, (int)M, 0, 63);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const unsigned *)nullptr, (unsigned *)nullptr, (int)M);
    return a > b ? a : b;
}
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int64_t dvcp_voxel_filter_workspace_bytes(int64_t M) {
    if (M <= 0 || M >= (1ll << 31)) return DVCP_E_ARG;
    return (int64_t)(2 * align256(8 * (size_t)M) + 4 * align256(4 * (size_t)M) + align256(voxel_cub_bytes(M)));
}

extern "C" int dvcp_voxel_grid_filter(const float *pts, int stride, int channels, int64_t M, float ox, float oy, float oz,
                                      float cell, int mode, void *workspace, int64_t capacity, float *out,
                                      int32_t *out_count, int64_t *n_out, dvcp_stream_t stream) {
    if (!pts || !workspace || !out || !n_out || M <= 0 || M >= (1ll << 31) || stride < 3 || channels < 3 || channels > 4 ||
        channels > stride || !(cell > 0.f) || capacity <= 0 || (mode != 0 && mode != 1))
        return DVCP_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *w = (unsigned char *)workspace;
    unsigned long long *k0 = (unsigned long long *)w, *k1 = (unsigned long long *)(w + align256(8 * (size_t)M));
    w += 2 * align256(8 * (size_t)M);
    unsigned *v0 = (unsigned *)w, *v1 = (unsigned *)(w + align256(4 * (size_t)M)), *head = (unsigned *)(w + 2 * align256(4 * (size_t)M)),
             *slot = (unsigned *)(w + 3 * align256(4 * (size_t)M));
    void *tmp = w + 4 * align256(4 * (size_t)M);
    size_t tmp_bytes = voxel_cub_bytes(M);
    const unsigned blocks = (unsigned)((M + 255) / 256);
    voxel_keys_kernel<<<blocks, 256, 0, st>>>(pts, stride, M, ox, oy, oz, cell, k0, v0);
    DVCP_CHECK_LAUNCH();
    DVCP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, (int)M, 0, 63, st));   // stable: ties keep index order
    voxel_heads_kernel<<<blocks, 256, 0, st>>>(k1, M, head);
    DVCP_CHECK_LAUNCH();
    DVCP_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, head, slot, (int)M, st));
    voxel_reduce_kernel<<<blocks, 256, 0, st>>>(pts, stride, channels, M, k1, v1, head, slot, mode, capacity, out, out_count, n_out);
    DVCP_CHECK_LAUNCH();
    return 0;
}
