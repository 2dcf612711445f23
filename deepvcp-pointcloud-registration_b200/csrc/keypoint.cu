// Key-point stage of DeepVCP.forward -- deepVCP.py:44-67,86-91,101.
//
// One CTA per pair does everything between the weighting layer's top-K and the
// candidate grid:
//   1. gather the key-points from src_pts; in reference mode the [C,K] gather
//      result is re-read row-major as [K,C] (deepVCP.py:46, quirk Q3);
//   2. sample_and_group(npoint=K, radius=1, nsample=32) AMONG the key-points
//      (deepVCP.py:54-56): FPS over the K key-points from the caller-drawn start,
//      then a ball query of the permuted key-points against the key-points;
//   3. index the full-cloud feature table with those 0..K-1 indices (Q5,
//      deepVCP.py:62);
//   4. Get_Cat_Feat_Src (get_cat_feat_src.py:37-53);
//   5. feat_embedding_layer(src=True) (deep_feat_embedding.py:30-44);
//   6. centres = R_init @ key-point in float64 (deepVCP.py:86-91; t_init is not
//      applied, Q6).
#include "dfe_common.cuh"

namespace dvcp {

constexpr int KP_MAX = 128;     // key-points per pair supported
constexpr int KP_MAXC = 8;      // input channels supported
constexpr int KP_THREADS = 256;

// arithmetic of the cloud's dtype T (float32, or float64 as the reference's loaders produce): every expression
// below is evaluated in T exactly where torch evaluates it in the tensor's dtype; the cast to float32 happens
// where the reference's `.float()` does (deep_feat_embedding.py:29)
template <class T> struct KpArith;
template <> struct KpArith<float> {
    static __device__ __forceinline__ float sq3(float x, float y, float z) { return sq3_nofma(x, y, z); }
    static __device__ __forceinline__ float sqd(float qx, float qy, float qz, float qq, float px, float py, float pz, float pp) {
        return sqdist_expanded(qx, qy, qz, qq, px, py, pz, pp);
    }
    static __device__ __forceinline__ float root(float v) { return sqrtf(v); }
};
template <> struct KpArith<double> {
    static __device__ __forceinline__ double sq3(double x, double y, double z) {
        return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
    }
    static __device__ __forceinline__ double sqd(double qx, double qy, double qz, double qq, double px, double py, double pz, double pp) {
        const double dot = __fma_rn(qz, pz, __fma_rn(qy, py, __dmul_rn(qx, px)));
        return __dadd_rn(__dadd_rn(__dmul_rn(-2.0, dot), qq), pp);
    }
    static __device__ __forceinline__ double root(double v) { return sqrt(v); }
};

template <class T>
__global__ void __launch_bounds__(KP_THREADS)
keypoint_stage_kernel(const T *__restrict__ src_pts, int C_in, int N, const int64_t *__restrict__ topk,
                      int Kp, const int64_t *__restrict__ kp_start, const float *__restrict__ src_feat, int S,
                      const double *__restrict__ R_init, const double *__restrict__ t_init, int64_t t_bstride,
                      T radius2, int nsample, dvcp_dfe_params_t P, int ref_layout, T *__restrict__ keypts, int64_t *__restrict__ picked,
                      float *__restrict__ src_cat, float *__restrict__ src_dfe, double *__restrict__ centres) {
    using A = KpArith<T>;
    __shared__ __align__(16) float s_w[DFE_SMEM_FLOATS];
    __shared__ T s_kp[KP_MAX * KP_MAXC];
    __shared__ T s_x[KP_MAX], s_y[KP_MAX], s_z[KP_MAX], s_pp[KP_MAX];
    __shared__ int s_fps[KP_MAX];
    __shared__ int s_pick[KP_MAX][32];
    // grid = (KP_SPLIT, B): every CTA of a pair repeats the cheap serial part (gather, FPS and ball query among
    // the K key-points) and takes every KP_SPLIT-th key-point row of the embedding; CTA 0 writes the shared outputs
    const int b = blockIdx.y, part = blockIdx.x, nparts = gridDim.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = KP_THREADS / 32;
    const bool lead = part == 0;
    dfe_stage_weights(P, s_w);
    const T *pts = src_pts + (int64_t)b * C_in * N;
    const int64_t *tk = topk + (int64_t)b * Kp;

    // 1. gather
    for (int f = tid; f < Kp * C_in; f += KP_THREADS) {
        int k, c;   // source (channel c, key-point k) of flat element f of keypts[Kp][C_in]
        if (ref_layout) {
            c = f / Kp;
            k = f - c * Kp;
        } else {
            k = f / C_in;
            c = f - k * C_in;
        }
        const T v = __ldg(pts + (int64_t)c * N + tk[k]);
        s_kp[f] = v;
        if (keypts && lead) keypts[(int64_t)b * Kp * C_in + f] = v;
    }
    __syncthreads();
    for (int k = tid; k < Kp; k += KP_THREADS) {
        const T x = s_kp[k * C_in], y = s_kp[k * C_in + 1], z = s_kp[k * C_in + 2];
        s_x[k] = x;
        s_y[k] = y;
        s_z[k] = z;
        s_pp[k] = A::sq3(x, y, z);
        if (centres && lead) {
            const double *R = R_init + (int64_t)b * 9;
            double *o = centres + ((int64_t)b * Kp + k) * 3;
#pragma unroll
            for (int r = 0; r < 3; ++r)
            {
                o[r] = __dadd_rn(__dadd_rn(__dmul_rn(R[3 * r], (double)x), __dmul_rn(R[3 * r + 1], (double)y)),
                                 __dmul_rn(R[3 * r + 2], (double)z));
                if (t_init) o[r] = __dadd_rn(o[r], t_init[(int64_t)b * t_bstride + r]);   // intended mode only (Q6)
            }
        }
    }
    __syncthreads();

    // 2a. FPS over the key-points (warp 0; lane holds points lane, lane+32, ...)
    if (warp == 0) {
        float dmin[KP_MAX / 32];
#pragma unroll
        for (int t = 0; t < KP_MAX / 32; ++t) dmin[t] = (lane + 32 * t < Kp) ? 1e10f : -1.0f;
        unsigned far = (unsigned)kp_start[b];
        for (int i = 0; i < Kp; ++i) {
            if (lane == 0) s_fps[i] = (int)far;
            const T cx = s_x[far], cy = s_y[far], cz = s_z[far];
            unsigned hi = 0u, lo = 0u;
#pragma unroll
            for (int t = 0; t < KP_MAX / 32; ++t) {
                const int n = lane + 32 * t;
                if (n < Kp) {
                    // float64 clouds: formed in double, compared with the float32 running minimum, stored
                    // rounded to float32 (pointnet2_utils.py:80-82)
                    const T d = A::sq3(s_x[n] - cx, s_y[n] - cy, s_z[n] - cz);
                    if (d < (T)dmin[t]) dmin[t] = (float)d;
                    const unsigned bits = __float_as_uint(dmin[t]);
                    const unsigned l = 0xffffffffu - (unsigned)n;
                    if (bits > hi || (bits == hi && l > lo)) {
                        hi = bits;
                        lo = l;
                    }
                }
            }
            warp_max_pair(hi, lo);
            far = 0xffffffffu - lo;
        }
    }
    __syncthreads();

    // 2b. ball query of the permuted key-points among the key-points
    for (int i = warp; i < Kp; i += nwarps) {
        const int c = s_fps[i];
        const T qx = s_x[c], qy = s_y[c], qz = s_z[c], qq = s_pp[c];
        int cnt = 0, first = Kp;
        for (int base = 0; base < Kp && cnt < nsample; base += 32) {
            const int n = base + lane;
            const bool ok = n < Kp;
            const bool in = ok && !(A::sqd(qx, qy, qz, qq, s_x[ok ? n : 0], s_y[ok ? n : 0], s_z[ok ? n : 0], s_pp[ok ? n : 0]) > radius2);
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (m) {
                if (first == Kp) first = base + __ffs(m) - 1;
                const int slot = cnt + __popc(m & ((1u << lane) - 1u));
                if (in && slot < nsample) s_pick[i][slot] = n;
                cnt += __popc(m);
            }
        }
        __syncwarp();
        if (lane >= cnt && lane < nsample) s_pick[i][lane] = first;
        __syncwarp();
        if (picked && lead && lane < nsample) picked[((int64_t)b * Kp + i) * nsample + lane] = s_pick[i][lane];
    }
    __syncthreads();

    // 3-5. per key-point row i: lanes over the nsample (= 32) group members
    const DfeSmem W(s_w);
    for (int i = part * nwarps + warp; i < Kp; i += nparts * nwarps) {
        const int c = s_fps[i];
        const int pj = s_pick[i][lane];
        // grouped_xyz_norm (pointnet2_utils.py:128): member - permuted key-point
        const T gx = s_x[pj] - s_x[c], gy = s_y[pj] - s_y[c], gz = s_z[pj] - s_z[c];
        // get_cat_feat_src.py:37-45: || kp_i - grouped + 1e-6 ||_2 with the UN-permuted key-point i
        const T kx = s_x[i], ky = s_y[i], kz = s_z[i];
        const T ex = (kx - gx) + (T)1e-6, ey = (ky - gy) + (T)1e-6, ez = (kz - gz) + (T)1e-6;
        const T dist = A::root(A::sq3(ex, ey, ez));
        T sum = dist;
#pragma unroll
        for (int s = 16; s; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
        const T wn = dist / sum;
        float x[36];   // the concatenated row after the reference's X.float()
        x[0] = (float)(gx - kx);
        x[1] = (float)(gy - ky);
        x[2] = (float)(gz - kz);
        const float4 *fp = reinterpret_cast<const float4 *>(src_feat + ((int64_t)b * S + pj) * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 v = __ldg(fp + k);
            x[3 + 4 * k] = (float)((T)v.x * wn);
            x[4 + 4 * k] = (float)((T)v.y * wn);
            x[5 + 4 * k] = (float)((T)v.z * wn);
            x[6 + 4 * k] = (float)((T)v.w * wn);
        }
        x[35] = 0.f;
        if (src_cat) {
            float *o = src_cat + (((int64_t)b * Kp + i) * 32 + lane) * DFE_IN;
#pragma unroll
            for (int k = 0; k < DFE_IN; ++k) o[k] = x[k];
        }
        if (src_dfe) {
            float y[32];
            dfe_row(x, W, y);
            src_dfe[((int64_t)b * Kp + i) * 32 + lane] = warp_colmax(y);
        }
    }
}

}  // namespace dvcp

using namespace dvcp;

template <class T>
static int keypoint_stage_launch(const T *src_pts, int C_in, int B, int N, const int64_t *topk, int Kp,
                                 const int64_t *kp_start, const float *src_feat, int S, const double *R_init,
                                 const double *t_init, int64_t t_bstride, T radius2, int nsample,
                                 dvcp_dfe_params_t dfe, int quirks, T *keypts, int64_t *picked, float *src_cat,
                                 float *src_dfe, double *centres, dvcp_stream_t stream) {
    if (!src_pts || !topk || !kp_start || !src_feat || B <= 0 || N <= 0 || S <= 0) return DVCP_E_ARG;
    if (centres && !R_init) return DVCP_E_ARG;
    if (!dfe.W1 || !dfe.b1 || !dfe.W2 || !dfe.b2 || !dfe.W3 || !dfe.b3) return DVCP_E_ARG;
    if (C_in < 3 || C_in > KP_MAXC || Kp < 1 || Kp > KP_MAX || nsample != 32 || Kp > S || B > 65535) return DVCP_E_UNSUPPORTED;
    const int split = (Kp + KP_THREADS / 32 - 1) / (KP_THREADS / 32);   // one key-point row per warp
    keypoint_stage_kernel<T><<<dim3(split, B), KP_THREADS, 0, (cudaStream_t)stream>>>(
        src_pts, C_in, N, topk, Kp, kp_start, src_feat, S, R_init, (quirks & DVCP_QUIRK_IGNORE_T_INIT) ? nullptr : t_init,
        t_bstride, radius2, nsample, dfe, quirks & 1, keypts, picked, src_cat, src_dfe, centres);
    DVCP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dvcp_keypoint_stage(const float *src_pts, int C_in, int B, int N, const int64_t *topk, int Kp,
                                   const int64_t *kp_start, const float *src_feat, int S, const double *R_init,
                                   const double *t_init, int64_t t_bstride, float radius2, int nsample,
                                   dvcp_dfe_params_t dfe, int quirks, float *keypts,
                                   int64_t *picked, float *src_cat, float *src_dfe, double *centres,
                                   dvcp_stream_t stream) {
    return keypoint_stage_launch<float>(src_pts, C_in, B, N, topk, Kp, kp_start, src_feat, S, R_init, t_init, t_bstride,
                                        radius2, nsample, dfe, quirks, keypts, picked, src_cat, src_dfe, centres, stream);
}

extern "C" int dvcp_keypoint_stage_f64(const double *src_pts, int C_in, int B, int N, const int64_t *topk, int Kp,
                                       const int64_t *kp_start, const float *src_feat, int S, const double *R_init,
                                       const double *t_init, int64_t t_bstride, double radius2, int nsample,
                                       dvcp_dfe_params_t dfe, int quirks, double *keypts,
                                       int64_t *picked, float *src_cat, float *src_dfe, double *centres,
                                       dvcp_stream_t stream) {
    return keypoint_stage_launch<double>(src_pts, C_in, B, N, topk, Kp, kp_start, src_feat, S, R_init, t_init, t_bstride,
                                         radius2, nsample, dfe, quirks, keypts, picked, src_cat, src_dfe, centres, stream);
}
