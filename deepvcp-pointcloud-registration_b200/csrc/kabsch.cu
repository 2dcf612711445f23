// Batched Kabsch pose solve -- deepVCP_loss.py:13-44 (get_rigid_transform) and
// :57-90 (svd_optimization: solve, 1-NN outlier rejection, solve again).
//
// One warp per problem, float64 throughout (the model path casts to double,
// deepVCP_loss.py:106-107). Points are read lane-strided (coalesced), the 3x3
// covariance is reduced with shuffles, and the SVD is a closed-loop one-sided
// Jacobi on the 3x3 matrix held in registers of every lane (no shared memory, no
// library call). R = V U^T is the orthogonal polar factor of H^T; it does not
// depend on the sign/order conventions of the SVD. Reference mode applies no
// reflection correction (quirk Q10: the reference builds Z but never uses it);
// with DVCP_QUIRK_NO_REFLECTION_FIX clear the term of the smallest singular value
// is flipped when det(V U^T) < 0 (R = V diag(1, 1, -1) U^T). Optional
// per-correspondence weights give the weighted Kabsch solve.
#include "common.cuh"

namespace dvcp {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// H (row-major 3x3) -> R = V U^T, where H = U S V^T.
__device__ void polar_from_svd(const double (&H)[9], double (&R)[9], bool fix_reflection) {
    double A[3][3], V[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            A[i][j] = H[3 * i + j];
            V[i][j] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            double al = 0, be = 0, ga = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                al += A[i][p] * A[i][p];
                be += A[i][q] * A[i][q];
                ga += A[i][p] * A[i][q];
            }
            const double lim = 1e-300 + 1e-32 * al * be;
            if (ga * ga > lim) {
                off += ga * ga / (al * be + 1e-300);
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double ap = A[i][p], aq = A[i][q];
                    A[i][p] = c * ap - s * aq;
                    A[i][q] = s * ap + c * aq;
                    const double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - s * vq;
                    V[i][q] = s * vp + c * vq;
                }
            }
        }
        if (off < 1e-30) break;
    }
    // columns of A are sigma_k u_k
    double sig[3], U[3][3];
    double smax = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        sig[k] = sqrt(A[0][k] * A[0][k] + A[1][k] * A[1][k] + A[2][k] * A[2][k]);
        smax = fmax(smax, sig[k]);
    }
    int nbad = 0, bad = -1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (sig[k] > 1e-13 * smax && sig[k] > 0) {
            const double inv = 1.0 / sig[k];
            U[0][k] = A[0][k] * inv;
            U[1][k] = A[1][k] * inv;
            U[2][k] = A[2][k] * inv;
        } else {
            ++nbad;
            bad = k;
        }
    }
    if (nbad == 1) {
        // rank 2: the third left vector is only defined up to sign; take the one
        // that makes R a proper rotation.
        const int a = (bad + 1) % 3, b = (bad + 2) % 3;
        double u[3] = {U[1][a] * U[2][b] - U[2][a] * U[1][b], U[2][a] * U[0][b] - U[0][a] * U[2][b],
                       U[0][a] * U[1][b] - U[1][a] * U[0][b]};
        const double detV = V[0][0] * (V[1][1] * V[2][2] - V[1][2] * V[2][1]) -
                            V[0][1] * (V[1][0] * V[2][2] - V[1][2] * V[2][0]) +
                            V[0][2] * (V[1][0] * V[2][1] - V[1][1] * V[2][0]);
        // det(U) with u in column `bad` equals +1 for (a,b,bad) cyclic
        const double sgn = detV >= 0 ? 1.0 : -1.0;
        U[0][bad] = sgn * u[0];
        U[1][bad] = sgn * u[1];
        U[2][bad] = sgn * u[2];
    } else if (nbad > 1) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) U[i][k] = V[i][k];   // H ~ 0: R = I
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            R[3 * i + j] = V[i][0] * U[j][0] + V[i][1] * U[j][1] + V[i][2] * U[j][2];
    if (fix_reflection) {
        const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) +
                           R[2] * (R[3] * R[7] - R[4] * R[6]);
        if (det < 0) {
            int km = 0;
            if (sig[1] < sig[km]) km = 1;
            if (sig[2] < sig[km]) km = 2;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    double vi = km == 0 ? V[i][0] : (km == 1 ? V[i][1] : V[i][2]);
                    double uj = km == 0 ? U[j][0] : (km == 1 ? U[j][1] : U[j][2]);
                    R[3 * i + j] -= 2.0 * vi * uj;
                }
        }
    }
}

// Kabsch over the points whose weight is non-zero (wt(i): 1 / 0 for the plain and
// the inlier-masked solve, any non-negative weight for the weighted one). lx(c, i)
// returns coordinate c of point i. Every lane returns the same R, t.
template <typename LX, typename LY, typename WT>
__device__ void kabsch_warp(int n, LX lx, LY ly, WT wt, bool fix_reflection, double (&R)[9], double (&t)[3]) {
    const int lane = lane_id();
    double sx[3] = {0, 0, 0}, sy[3] = {0, 0, 0}, cnt = 0;
    for (int i = lane; i < n; i += 32) {
        const double w = wt(i);
        if (w != 0.0) {
            cnt += w;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                sx[c] += w * lx(c, i);
                sy[c] += w * ly(c, i);
            }
        }
    }
    cnt = warp_sum_d(cnt);
    double mx[3], my[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        mx[c] = warp_sum_d(sx[c]) / cnt;
        my[c] = warp_sum_d(sy[c]) / cnt;
    }
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = lane; i < n; i += 32) {
        const double w = wt(i);
        if (w != 0.0) {
            double dx[3], dy[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                dx[c] = lx(c, i) - mx[c];
                dy[c] = w * (ly(c, i) - my[c]);
            }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) H[3 * r + c] += dx[r] * dy[c];
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) H[k] = warp_sum_d(H[k]);
    polar_from_svd(H, R, fix_reflection);
#pragma unroll
    for (int r = 0; r < 3; ++r) t[r] = my[r] - (R[3 * r] * mx[0] + R[3 * r + 1] * mx[1] + R[3 * r + 2] * mx[2]);
}

constexpr int KB_WARPS = 4;

template <typename T>
__global__ void __launch_bounds__(KB_WARPS * 32)
kabsch_kernel(const T *__restrict__ x, const T *__restrict__ y, const double *__restrict__ wgt, int B, int n,
              bool fix_reflection, double *__restrict__ Ro, double *__restrict__ to) {
    const int b = blockIdx.x * KB_WARPS + (threadIdx.x >> 5);
    if (b >= B) return;
    const T *xb = x + (int64_t)b * 3 * n, *yb = y + (int64_t)b * 3 * n;
    double R[9], t[3];
    kabsch_warp(
        n, [&](int c, int i) { return (double)xb[c * n + i]; }, [&](int c, int i) { return (double)yb[c * n + i]; },
        [&](int i) { return wgt ? wgt[(int64_t)b * n + i] : 1.0; }, fix_reflection, R, t);
    const int lane = lane_id();
    if (lane < 9) Ro[(int64_t)b * 9 + lane] = R[lane];
    if (lane < 3) to[(int64_t)b * 3 + lane] = t[lane];
}

constexpr int KR_MAXN = 1024;

// x, yp: element (b, c, i) at base[b * 3 * n + c * cs + i * ps] -- the reference's [B,3,n] float64 arguments
// (cs = n, ps = 1) or the forward's own [B,n,3] float32 outputs (cs = 1, ps = 3; cast like .double())
template <typename T>
__global__ void __launch_bounds__(32)
kabsch_refine_kernel(const T *__restrict__ x, const T *__restrict__ yp, int cs, int ps, const double *__restrict__ Rt,
                     const double *__restrict__ tt, int n, int keepn, bool fix_reflection, double *__restrict__ R2o,
                     double *__restrict__ t2o, double *__restrict__ R1o, double *__restrict__ t1o,
                     int64_t *__restrict__ inliers) {
    __shared__ float s_ref[3][KR_MAXN];   // float32(y_pred1)
    __shared__ float s_d[KR_MAXN];        // 1-NN distance of y_true[i]
    __shared__ unsigned char s_keep[KR_MAXN];
    const int b = blockIdx.x, lane = lane_id();
    const T *xb = x + (int64_t)b * 3 * n, *yb = yp + (int64_t)b * 3 * n;
    const double *Rg = Rt + (int64_t)b * 9, *tg = tt + (int64_t)b * 3;
    double R1[9], t1[3];
    auto lx = [&](int c, int i) { return (double)xb[c * cs + i * ps]; };
    kabsch_warp(n, lx, [&](int c, int i) { return (double)yb[c * cs + i * ps]; }, [](int) { return 1.0; }, fix_reflection, R1, t1);
    if (R1o && lane < 9) R1o[(int64_t)b * 9 + lane] = R1[lane];
    if (t1o && lane < 3) t1o[(int64_t)b * 3 + lane] = t1[lane];
    // y_pred1 = R1 x + t1 (float64), cast to float32 inside the KNN (knn_cuda casts)
    for (int i = lane; i < n; i += 32) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
            s_ref[r][i] = (float)(R1[3 * r] * lx(0, i) + R1[3 * r + 1] * lx(1, i) + R1[3 * r + 2] * lx(2, i) + t1[r]);
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        float q[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            q[r] = (float)(Rg[3 * r] * lx(0, i) + Rg[3 * r + 1] * lx(1, i) + Rg[3 * r + 2] * lx(2, i) + tg[r]);
        float best = INFINITY;
        for (int j = 0; j < n; ++j)
            best = fminf(best, sqdist_direct(s_ref[0][j] - q[0], s_ref[1][j] - q[1], s_ref[2][j] - q[2]));
        s_d[i] = __fsqrt_rn(best);
    }
    __syncwarp();
    // keep the `keepn` smallest (distance, index)
    for (int i = lane; i < n; i += 32) {
        const float di = s_d[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const float dj = s_d[j];
            rank += (dj < di) || (dj == di && j < i);
        }
        s_keep[i] = rank < keepn;
        // the inlier list in the order of torch.topk(largest=False, sorted=True) (deepVCP_loss.py:77)
        if (inliers && rank < keepn) inliers[(int64_t)b * keepn + rank] = i;
    }
    __syncwarp();
    double R2[9], t2[3];
    kabsch_warp(
        n, lx, [&](int c, int i) { return R1[3 * c] * lx(0, i) + R1[3 * c + 1] * lx(1, i) + R1[3 * c + 2] * lx(2, i) + t1[c]; },
        [&](int i) { return s_keep[i] != 0 ? 1.0 : 0.0; }, fix_reflection, R2, t2);
    if (lane < 9) R2o[(int64_t)b * 9 + lane] = R2[lane];
    if (lane < 3) t2o[(int64_t)b * 3 + lane] = t2[lane];
}

}  // namespace dvcp

using namespace dvcp;

extern "C" int dvcp_kabsch(const void *x, const void *y, int dtype, const double *weights, int B, int n, int quirks,
                           double *R, double *t, dvcp_stream_t stream) {
    if (!x || !y || !R || !t || B <= 0 || n <= 0) return DVCP_E_ARG;
    const unsigned grid = (unsigned)((B + KB_WARPS - 1) / KB_WARPS);
    const bool fix = !(quirks & DVCP_QUIRK_NO_REFLECTION_FIX);
    if (dtype == 0)
        kabsch_kernel<float><<<grid, KB_WARPS * 32, 0, (cudaStream_t)stream>>>((const float *)x, (const float *)y, weights, B, n, fix, R, t);
    else if (dtype == 1)
        kabsch_kernel<double><<<grid, KB_WARPS * 32, 0, (cudaStream_t)stream>>>((const double *)x, (const double *)y, weights, B, n, fix, R, t);
    else
        return DVCP_E_ARG;
    DVCP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dvcp_kabsch_refine(const double *x, const double *y_pred, const double *R_true,
                                  const double *t_true, int B, int n, int keep, int quirks, double *R2, double *t2,
                                  double *R1, double *t1, int64_t *inliers, dvcp_stream_t stream) {
    if (!x || !y_pred || !R_true || !t_true || !R2 || !t2 || B <= 0 || n <= 0 || keep <= 0 || keep > n)
        return DVCP_E_ARG;
    if (n > KR_MAXN) return DVCP_E_UNSUPPORTED;
    kabsch_refine_kernel<double><<<B, 32, 0, (cudaStream_t)stream>>>(x, y_pred, n, 1, R_true, t_true, n, keep,
                                                                     !(quirks & DVCP_QUIRK_NO_REFLECTION_FIX), R2, t2, R1, t1,
                                                                     inliers);
    DVCP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dvcp_pose_from_forward(const float *src_keypts, const float *tgt_vcp, const double *R_true,
                                      const double *t_true, int B, int n, int keep, int quirks, double *R2, double *t2,
                                      dvcp_stream_t stream) {
    if (!src_keypts || !tgt_vcp || !R_true || !t_true || !R2 || !t2 || B <= 0 || n <= 0 || keep <= 0 || keep > n)
        return DVCP_E_ARG;
    if (n > KR_MAXN) return DVCP_E_UNSUPPORTED;
    kabsch_refine_kernel<float><<<B, 32, 0, (cudaStream_t)stream>>>(src_keypts, tgt_vcp, 1, 3, R_true, t_true, n, keep,
                                                                    !(quirks & DVCP_QUIRK_NO_REFLECTION_FIX), R2, t2,
                                                                    nullptr, nullptr, nullptr);
    DVCP_CHECK_LAUNCH();
    return 0;
}
