// Corresponding-point generation -- cpg.py:27-60.
//
//   cost[c', f'] = (src[f'] - T'[c', f'])^2, where T' re-reads the LOGICAL
//   row-major order of the [32, C] target tensor as [C, 32] (quirk Q4);
//   volume [32, G, G, G] -> Conv3d 32->16 -> 16->4 -> 4->1 (k=3, pad=1, bias, no
//   activation) -> softmax over the C voxels -> vcp = sum w_c cand_c / sum w_c.
//
// Two paths: up to 11^3 voxels the whole chain runs in ONE kernel with the volume in
// shared memory (cpg_fused_kernel below); larger grids use one kernel per layer (one
// thread per output voxel; activations travel through an L2-resident workspace laid
// out [M][channel][voxel]; weights re-ordered in shared memory to [tap][cin][cout] so
// one LDS.128 feeds four FMAs).
#include "common.cuh"

namespace dvcp {

__global__ void __launch_bounds__(256)
cpg_cost_kernel(const float *__restrict__ src, const float *__restrict__ tgt, int layout, int C,
                float *__restrict__ vol) {
    const int64_t m = blockIdx.y;
    const float *t = tgt + m * 32 * (int64_t)C;
    const float *s = src + m * 32;
    float *v = vol + m * 32 * (int64_t)C;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 32 * C; e += gridDim.x * blockDim.x) {
        // e enumerates (f', c') with c' fastest so that the write is coalesced
        const int f = e / C, c = e - f * C;
        const int L = c * 32 + f;   // flat position in the logical [32, C] tensor
        const float tv = layout == 0 ? __ldg(t + L) : __ldg(t + (int64_t)(L % C) * 32 + (L / C));
        const float d = __ldg(s + f) - tv;
        v[e] = d * d;
    }
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(128)
cpg_conv_kernel(const float *__restrict__ in, const float *__restrict__ w, const float *__restrict__ bias, int G,
                float *__restrict__ out) {
    extern __shared__ __align__(16) float sw[];   // [27][CIN][COUT]
    const int C = G * G * G;
    for (int i = threadIdx.x; i < 27 * CIN * COUT; i += blockDim.x) {
        const int co = i % COUT, ci = (i / COUT) % CIN, tap = i / (COUT * CIN);
        sw[i] = w[(co * CIN + ci) * 27 + tap];
    }
    __syncthreads();
    const int64_t m = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int iz = c % G, iy = (c / G) % G, ix = c / (G * G);
    const float *src = in + m * CIN * (int64_t)C;
    float acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = __ldg(bias + o);
    for (int dx = -1; dx <= 1; ++dx) {
        const int x = ix + dx;
        if (x < 0 || x >= G) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = iy + dy;
            if (y < 0 || y >= G) continue;
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz) {
                const int z = iz + dz;
                if (z < 0 || z >= G) continue;
                const int tap = ((dx + 1) * 3 + (dy + 1)) * 3 + (dz + 1);
                const float *p = src + (x * G + y) * G + z;
                const float *wt = sw + tap * CIN * COUT;
#pragma unroll 4
                for (int ci = 0; ci < CIN; ++ci) {
                    const float v = __ldg(p + (int64_t)ci * C);
                    if (COUT % 4 == 0) {
                        const float4 *w4 = reinterpret_cast<const float4 *>(wt + ci * COUT);
#pragma unroll
                        for (int o = 0; o < COUT / 4; ++o) {
                            const float4 ww = w4[o];
                            acc[4 * o] = fmaf(ww.x, v, acc[4 * o]);
                            acc[4 * o + 1] = fmaf(ww.y, v, acc[4 * o + 1]);
                            acc[4 * o + 2] = fmaf(ww.z, v, acc[4 * o + 2]);
                            acc[4 * o + 3] = fmaf(ww.w, v, acc[4 * o + 3]);
                        }
                    } else {
#pragma unroll
                        for (int o = 0; o < COUT; ++o) acc[o] = fmaf(wt[ci * COUT + o], v, acc[o]);
                    }
                }
            }
        }
    }
    float *dst = out + m * COUT * (int64_t)C + c;
#pragma unroll
    for (int o = 0; o < COUT; ++o) dst[(int64_t)o * C] = acc[o];
}

// softmax over the C voxels of one key-point + weighted candidate sum.
__global__ void __launch_bounds__(256)
cpg_softmax_vcp_kernel(const float *__restrict__ logits, const float *__restrict__ cand, int C,
                       float *__restrict__ vcp) {
    __shared__ float red[4][8];
    const int64_t m = blockIdx.x;
    const float *l = logits + m * C;
    const float *cp = cand + m * C * 3;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float mx = -INFINITY;
    for (int c = tid; c < C; c += blockDim.x) mx = fmaxf(mx, l[c]);
#pragma unroll
    for (int s = 16; s; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    if (lane == 0) red[0][warp] = mx;
    __syncthreads();
    mx = red[0][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[0][w]);
    float z = 0.f;
    for (int c = tid; c < C; c += blockDim.x) z += expf(l[c] - mx);
#pragma unroll
    for (int s = 16; s; s >>= 1) z += __shfl_xor_sync(0xffffffffu, z, s);
    __syncthreads();
    if (lane == 0) red[0][warp] = z;
    __syncthreads();
    z = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) z += red[0][w];
    float a[4] = {0.f, 0.f, 0.f, 0.f};   // sum w*x, w*y, w*z, sum w
    for (int c = tid; c < C; c += blockDim.x) {
        const float w = expf(l[c] - mx) / z;
        a[0] = fmaf(w, cp[3 * c], a[0]);
        a[1] = fmaf(w, cp[3 * c + 1], a[1]);
        a[2] = fmaf(w, cp[3 * c + 2], a[2]);
        a[3] += w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int s = 16; s; s >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], s);
        if (lane == 0) red[k][warp] = a[k];
    }
    __syncthreads();
    if (tid < 3) {
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            num += red[tid][w];
            den += red[3][w];
        }
        vcp[m * 3 + tid] = num / den;
    }
}

// ------------------------------------------------------------ fused (G <= 11) --
// One CTA per key-point volume; nothing but the inputs and the 3 output floats
// touches global memory. The cost volume lives in shared memory half of the input
// channels at a time ([16][C] = 85 KB at G = 11), conv1's accumulators stay in
// registers across the two halves (thread = 4 consecutive z voxels x 16 output
// channels: per (dx, dy, cin) 6 input LDS + 12 broadcast LDS.128 of weights feed 192
// FMAs), its output overwrites the volume, conv2 / conv3 / softmax follow in place.
constexpr int CF_THREADS = 768;   // two threads per (line, z-group): 8 of conv1's 16 output channels each
constexpr int CF_ITEMS = 384;     // >= G * G * ceil(G / 4) for G <= 11
constexpr int CF_MAXG = 11;

__device__ __forceinline__ float cf_block_sum(float v, float *red, int tid) {
#pragma unroll
    for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < CF_THREADS / 32; ++w) t += red[w];
    return t;
}

constexpr int CF_MAXVPC = 8;   // volumes one CTA works on at a time (small grids)

__global__ void __launch_bounds__(CF_THREADS, 1)
cpg_fused_kernel(const float *__restrict__ src, const float *__restrict__ tgt, int layout,
                 const float *__restrict__ cand, int64_t M, int G, int VPC, dvcp_cpg_params_t p,
                 float *__restrict__ vcp, float *__restrict__ logits_out) {
    extern __shared__ __align__(16) float sm[];
    __shared__ float red[CF_THREADS / 32];
    __shared__ float s_src[CF_MAXVPC][32];
    const int C = G * G * G, Cp = (C + 3) & ~3, VB = 21 * Cp;
    float *W1 = sm;                   // [27][32][16]
    float *W2 = W1 + 27 * 32 * 16;    // [27][16][4]
    float *W3 = W2 + 27 * 16 * 4;     // [27][4]  (108 floats, padded to 112)
    float *VOL = W3 + 112;            // per volume: cost half / conv1 out [16][Cp], conv2 out [4][Cp], logits [Cp]
    const int tid = threadIdx.x;
    for (int i = tid; i < 27 * 32 * 16; i += CF_THREADS) {
        const int co = i & 15, ci = (i >> 4) & 31, tap = i >> 9;
        W1[i] = __ldg(p.w1 + (co * 32 + ci) * 27 + tap);
    }
    for (int i = tid; i < 27 * 16 * 4; i += CF_THREADS) {
        const int co = i & 3, ci = (i >> 2) & 15, tap = i >> 6;
        W2[i] = __ldg(p.w2 + (co * 16 + ci) * 27 + tap);
    }
    for (int i = tid; i < 27 * 4; i += CF_THREADS) W3[i] = __ldg(p.w3 + (i & 3) * 27 + (i >> 2));

    // thread = (local volume, line (x, y), group of 4 z, half of conv1's output channels)
    const int NG = (G + 3) >> 2, items = G * G * NG;
    const int it = tid % CF_ITEMS, ch = tid / CF_ITEMS;   // ch is uniform per warp
    const int vl = it / items, item = it - vl * items;
    const int line = item / NG, zg = item - line * NG;
    const int x = line / G, y = line - x * G, z0 = zg * 4;
    float *A = VOL + (vl < VPC ? vl : 0) * VB, *O2 = A + 16 * Cp;
    const float b3 = __ldg(p.b3);

    for (int64_t m0 = (int64_t)blockIdx.x * VPC; m0 < M; m0 += (int64_t)gridDim.x * VPC) {
        const int nv = (int)(M - m0 < VPC ? M - m0 : VPC);   // volumes of this round
        const bool active = vl < nv;
        __syncthreads();   // weights staged / the previous round is finished with the shared volumes
        if (tid < 32 * nv) s_src[tid >> 5][tid & 31] = __ldg(src + (m0 + (tid >> 5)) * 32 + (tid & 31));
        // conv1 accumulators as float pairs: the packed FFMA2 of sm_100 (two IEEE FMAs per instruction, same
        // results) halves the issue slots of the kernel's dominant instruction
        float2 acc[4][4];
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int o = 0; o < 4; ++o) acc[v][o] = make_float2(__ldg(p.b1 + ch * 8 + 2 * o), __ldg(p.b1 + ch * 8 + 2 * o + 1));

        for (int half = 0; half < 2; ++half) {
            __syncthreads();   // s_src staged; the previous half's reads of the cost volumes are finished
            // cost[c', f'] = (src[f'] - T'[c', f'])^2 with c' * 32 + f' = the element's position in the
            // LOGICAL row-major [32, C] order (quirk Q4). Memory is walked in its own order (coalesced).
            for (int v = 0; v < nv; ++v) {
                const float *t = tgt + (m0 + v) * 32 * (int64_t)C;
                float *Av = VOL + v * VB;
                for (int e0 = tid; e0 < 32 * C; e0 += CF_THREADS * 8) {   // 8 loads in flight per thread
                    float tv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int e = e0 + u * CF_THREADS;
                        tv[u] = e < 32 * C ? __ldg(t + e) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int e = e0 + u * CF_THREADS;
                        const int L = layout == 0 ? e : (e & 31) * C + (e >> 5);
                        const int f = L & 31, c = L >> 5;
                        if (e < 32 * C && (f >> 4) == half) {
                            const float d = s_src[v][f] - tv[u];
                            Av[(f & 15) * Cp + c] = d * d;
                        }
                    }
                }
            }
            __syncthreads();
            if (active) {
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = x + dx;
                    if (xx < 0 || xx >= G) continue;
                    for (int dy = -1; dy <= 1; ++dy) {
                        const int yy = y + dy;
                        if (yy < 0 || yy >= G) continue;
                        const int tap0 = ((dx + 1) * 3 + (dy + 1)) * 3;
                        const float *col = A + (xx * G + yy) * G + z0 - 1;
                        const float *wt = W1 + (tap0 * 32 + half * 16) * 16 + ch * 8;
#pragma unroll 2
                        for (int ci = 0; ci < 16; ++ci) {
                            float in[6];
#pragma unroll
                            for (int k = 0; k < 6; ++k) {
                                const int z = z0 - 1 + k;
                                in[k] = (z >= 0 && z < G) ? col[ci * Cp + k] : 0.f;
                            }
#pragma unroll
                            for (int dz = 0; dz < 3; ++dz) {
                                const float4 *w4 = reinterpret_cast<const float4 *>(wt + (dz * 32 + ci) * 16);
#pragma unroll
                                for (int o4 = 0; o4 < 2; ++o4) {
                                    const float4 w = w4[o4];
                                    const float2 wa = make_float2(w.x, w.y), wb = make_float2(w.z, w.w);
#pragma unroll
                                    for (int v = 0; v < 4; ++v) {
                                        const float2 x2 = make_float2(in[v + dz], in[v + dz]);
                                        acc[v][2 * o4] = __ffma2_rn(wa, x2, acc[v][2 * o4]);
                                        acc[v][2 * o4 + 1] = __ffma2_rn(wb, x2, acc[v][2 * o4 + 1]);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();   // every read of the cost volumes is done: they become conv1's output [16][Cp]
        if (active) {
            const int c0 = (x * G + y) * G + z0;
#pragma unroll
            for (int v = 0; v < 4; ++v)
                if (z0 + v < G) {
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        A[(ch * 8 + 2 * o) * Cp + c0 + v] = acc[v][o].x;
                        A[(ch * 8 + 2 * o + 1) * Cp + c0 + v] = acc[v][o].y;
                    }
                }
        }
        __syncthreads();
        // ---- conv2 16 -> 4 ----
        if (active && ch == 0) {
            float2 a2[4][2];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                a2[v][0] = make_float2(__ldg(p.b2), __ldg(p.b2 + 1));
                a2[v][1] = make_float2(__ldg(p.b2 + 2), __ldg(p.b2 + 3));
            }
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= G) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = y + dy;
                    if (yy < 0 || yy >= G) continue;
                    const int tap0 = ((dx + 1) * 3 + (dy + 1)) * 3;
                    const float *col = A + (xx * G + yy) * G + z0 - 1;
#pragma unroll 4
                    for (int ci = 0; ci < 16; ++ci) {
                        float in[6];
#pragma unroll
                        for (int k = 0; k < 6; ++k) {
                            const int z = z0 - 1 + k;
                            in[k] = (z >= 0 && z < G) ? col[ci * Cp + k] : 0.f;
                        }
#pragma unroll
                        for (int dz = 0; dz < 3; ++dz) {
                            const float4 w = *reinterpret_cast<const float4 *>(W2 + ((tap0 + dz) * 16 + ci) * 4);
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                const float2 x2 = make_float2(in[v + dz], in[v + dz]);
                                a2[v][0] = __ffma2_rn(make_float2(w.x, w.y), x2, a2[v][0]);
                                a2[v][1] = __ffma2_rn(make_float2(w.z, w.w), x2, a2[v][1]);
                            }
                        }
                    }
                }
            }
            const int c0 = (x * G + y) * G + z0;
#pragma unroll
            for (int v = 0; v < 4; ++v)
                if (z0 + v < G) {
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        O2[(2 * o) * Cp + c0 + v] = a2[v][o].x;
                        O2[(2 * o + 1) * Cp + c0 + v] = a2[v][o].y;
                    }
                }
        }
        __syncthreads();
        // ---- conv3 4 -> 1: thread per voxel of every volume of the round ----
        for (int i = tid; i < nv * C; i += CF_THREADS) {
            const int v = i / C, c = i - v * C;
            const float *O2v = VOL + v * VB + 16 * Cp;
            const int iz = c % G, iy = (c / G) % G, ix = c / (G * G);
            float a3 = b3;
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = ix + dx;
                if (xx < 0 || xx >= G) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = iy + dy;
                    if (yy < 0 || yy >= G) continue;
#pragma unroll
                    for (int dz = -1; dz <= 1; ++dz) {
                        const int zz = iz + dz;
                        if (zz < 0 || zz >= G) continue;
                        const int tap = ((dx + 1) * 3 + (dy + 1)) * 3 + (dz + 1);
                        const float4 w = *reinterpret_cast<const float4 *>(W3 + tap * 4);
                        const int cc = (xx * G + yy) * G + zz;
                        a3 = fmaf(w.x, O2v[cc], a3);
                        a3 = fmaf(w.y, O2v[Cp + cc], a3);
                        a3 = fmaf(w.z, O2v[2 * Cp + cc], a3);
                        a3 = fmaf(w.w, O2v[3 * Cp + cc], a3);
                    }
                }
            }
            VOL[v * VB + 20 * Cp + c] = a3;
            if (logits_out) logits_out[(m0 + v) * C + c] = a3;
        }
        __syncthreads();
        // ---- softmax over the C voxels + weighted candidate sum, volume after volume ----
        for (int v = 0; v < nv; ++v) {
            const float *LG = VOL + v * VB + 20 * Cp;
            float mx = -INFINITY;
            for (int c = tid; c < C; c += CF_THREADS) mx = fmaxf(mx, LG[c]);
#pragma unroll
            for (int s = 16; s; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
            __syncthreads();
            if ((tid & 31) == 0) red[tid >> 5] = mx;
            __syncthreads();
            mx = red[0];
#pragma unroll
            for (int w = 1; w < CF_THREADS / 32; ++w) mx = fmaxf(mx, red[w]);
            float zp = 0.f;
            for (int c = tid; c < C; c += CF_THREADS) zp += expf(LG[c] - mx);
            const float Z = cf_block_sum(zp, red, tid);
            const float *cp = cand + (m0 + v) * C * 3;
            float a[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c = tid; c < C; c += CF_THREADS) {
                const float w = expf(LG[c] - mx) / Z;
                a[0] = fmaf(w, __ldg(cp + 3 * c), a[0]);
                a[1] = fmaf(w, __ldg(cp + 3 * c + 1), a[1]);
                a[2] = fmaf(w, __ldg(cp + 3 * c + 2), a[2]);
                a[3] += w;
            }
            const float sx = cf_block_sum(a[0], red, tid), sy = cf_block_sum(a[1], red, tid),
                        sz = cf_block_sum(a[2], red, tid), sw = cf_block_sum(a[3], red, tid);
            if (tid == 0) {
                vcp[(m0 + v) * 3] = sx / sw;
                vcp[(m0 + v) * 3 + 1] = sy / sw;
                vcp[(m0 + v) * 3 + 2] = sz / sw;
            }
        }
    }
}

}  // namespace dvcp

namespace dvcp {
int cpg_tc_launch(const float *src_dfe, const float *tgt_dfe, const float *cand, int64_t M, int G, dvcp_cpg_params_t p,
                  float *vcp, float *logits, float *image, cudaStream_t st);   // cpg_tc.cu
int cpg_tcz_launch(const float *src_dfe, const float *tgt_dfe, const float *cand, int64_t M, int G, dvcp_cpg_params_t p,
                   float *vcp, float *logits, float *image, cudaStream_t st);  // cpg_tc.cu
}
extern "C" int64_t dvcp_cpg_tc_image_bytes(void);

using namespace dvcp;

extern "C" int64_t dvcp_cpg_workspace_bytes(int64_t M, int G) {
    if (M <= 0 || G <= 0) return 0;
    const int64_t layered = M * (int64_t)G * G * G * (32 + 16 + 4 + 1) * (int64_t)sizeof(float);
    const int64_t tc = dvcp_cpg_tc_image_bytes();   // the tensor-core kernel's split weight image
    return layered > tc ? layered : tc;
}

extern "C" int dvcp_cpg_path(const float *src_dfe, const float *tgt_dfe, int layout, const float *cand, int64_t M,
                             int G, dvcp_cpg_params_t p, float *vcp, float *logits, void *workspace,
                             int64_t workspace_bytes, int path, dvcp_stream_t stream) {
    if (!src_dfe || !tgt_dfe || !cand || !vcp || !workspace || M <= 0 || G <= 0) return DVCP_E_ARG;
    if (!p.w1 || !p.b1 || !p.w2 || !p.b2 || !p.w3 || !p.b3 || (layout != 0 && layout != 1)) return DVCP_E_ARG;
    if (M > 65535 || G > 64) return DVCP_E_UNSUPPORTED;
    if (workspace_bytes < dvcp_cpg_workspace_bytes(M, G)) return DVCP_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int C = G * G * G;
    if (path < DVCP_CPG_AUTO || path > DVCP_CPG_TCZ) return DVCP_E_ARG;
    if (path == DVCP_CPG_FUSED && G > CF_MAXG) return DVCP_E_UNSUPPORTED;
    if ((path == DVCP_CPG_TC || path == DVCP_CPG_TCZ) && (G > CF_MAXG || G < 2 || layout != 0)) return DVCP_E_UNSUPPORTED;
    if (path == DVCP_CPG_TCZ || (path == DVCP_CPG_AUTO && layout == 0 && G >= 2 && G <= CF_MAXG))
        return cpg_tcz_launch(src_dfe, tgt_dfe, cand, M, G, p, vcp, logits, (float *)workspace, st);
    if (path == DVCP_CPG_TC)
        return cpg_tc_launch(src_dfe, tgt_dfe, cand, M, G, p, vcp, logits, (float *)workspace, st);
    if (G <= CF_MAXG && path != DVCP_CPG_LAYERED) {
        const int Cp = (C + 3) & ~3;
        const int items = G * G * ((G + 3) / 4);
        int VPC = CF_ITEMS / items;          // volumes per CTA round: all 384 item threads busy on small grids
        if (VPC > CF_MAXVPC) VPC = CF_MAXVPC;
        if ((int64_t)VPC * DVCP_NUM_SMS > M) VPC = (int)(M / DVCP_NUM_SMS);   // few volumes: one CTA each first
        if (VPC < 1) VPC = 1;
        const int smem = (27 * 32 * 16 + 27 * 16 * 4 + 112 + VPC * 21 * Cp) * (int)sizeof(float);
        DVCP_CUDA(cudaFuncSetAttribute(cpg_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int64_t grid = (M + VPC - 1) / VPC;
        if (VPC > 1 && grid > DVCP_NUM_SMS) grid = DVCP_NUM_SMS;   // persistent over rounds: weights staged once
        cpg_fused_kernel<<<(unsigned)grid, CF_THREADS, smem, st>>>(src_dfe, tgt_dfe, layout, cand, M, G, VPC, p, vcp, logits);
        DVCP_CHECK_LAUNCH();
        return 0;
    }
    float *v0 = (float *)workspace;
    float *v1 = v0 + M * 32 * (int64_t)C;
    float *v2 = v1 + M * 16 * (int64_t)C;
    float *v3 = v2 + M * 4 * (int64_t)C;
    {
        dim3 grid((32 * C + 255) / 256, (unsigned)M);
        cpg_cost_kernel<<<grid, 256, 0, st>>>(src_dfe, tgt_dfe, layout, C, v0);
        DVCP_CHECK_LAUNCH();
    }
    dim3 cgrid((C + 127) / 128, (unsigned)M);
    {
        auto k = cpg_conv_kernel<32, 16>;
        const int smem = 27 * 32 * 16 * sizeof(float);
        DVCP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k<<<cgrid, 128, smem, st>>>(v0, p.w1, p.b1, G, v1);
        DVCP_CHECK_LAUNCH();
    }
    cpg_conv_kernel<16, 4><<<cgrid, 128, 27 * 16 * 4 * sizeof(float), st>>>(v1, p.w2, p.b2, G, v2);
    DVCP_CHECK_LAUNCH();
    cpg_conv_kernel<4, 1><<<cgrid, 128, 27 * 4 * 1 * sizeof(float), st>>>(v2, p.w3, p.b3, G, v3);
    DVCP_CHECK_LAUNCH();
    cpg_softmax_vcp_kernel<<<(unsigned)M, 256, 0, st>>>(v3, cand, C, vcp);
    DVCP_CHECK_LAUNCH();
    if (logits) DVCP_CUDA(cudaMemcpyAsync(logits, v3, M * (int64_t)C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int dvcp_cpg(const float *src_dfe, const float *tgt_dfe, int layout, const float *cand, int64_t M,
                        int G, dvcp_cpg_params_t p, float *vcp, float *logits, void *workspace,
                        int64_t workspace_bytes, dvcp_stream_t stream) {
    return dvcp_cpg_path(src_dfe, tgt_dfe, layout, cand, M, G, p, vcp, logits, workspace, workspace_bytes,
                         DVCP_CPG_AUTO, stream);
}
