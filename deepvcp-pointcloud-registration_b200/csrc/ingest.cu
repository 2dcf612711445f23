// Data ingest for KITTI-shaped scans: the step immediately before the hot path.
//
// Reference: KITTIDataset.py:11-16 (downsample: N rows of the raw [M, 4] scan, chosen by the caller's
// np.random.choice), :44-46 (split into xyz and reflectance), :67-84 (target = R @ src + t in float64) and
// the [B, 3, N] channel-major layout the model takes (deepVCP.py:24). One launch for B scans: raw rows are
// gathered straight into the model's layout (coalesced writes), the target cloud is produced in the same
// pass, nothing goes through host memory.
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
