// Corresponding-point generation with conv1 on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Reference: cpg.py:27-60. conv1 (Conv3d 32 -> 16, 3x3x3, zero padding) is 88 % of the
// layer's arithmetic; here it is an IMPLICIT GEMM
//     D[voxel, cout] = sum over (tap, cin) of  V[voxel + shift(tap), cin] * W1[cout, cin, tap]
// with M = voxels, N = 16, K = 27 taps x 32 channels, evaluated as tcgen05.mma kind::tf32 with the 3xTF32
// split (V = Vh + Vl, W = Wh + Wl; Vh*Wh + Vl*Wh + Vh*Wl, FP32 accumulate in TMEM: FP32-level accuracy).
//
// The cost volume V of one key-point lives in shared memory as a ZERO-PADDED volume (G+2)^3 in the
// K-major no-swizzle operand layout with the rows laid out contiguously:
//     element (row r, channel k) of a channel octet at  plane(k / 4) + r * 16 B + (k % 4) * 4 B
// (core matrices of 8 rows x 16 B follow each other every 128 B = SBO, the two K halves of an MMA are one
// plane = LBO apart). A row is a voxel of the padded volume in linear order, so the rows a tap needs are
// the SAME rows shifted by a constant: the A descriptor of tap (dx, dy, dz) is the base descriptor plus
// ((dx * Gp + dy) * Gp + dz) rows, zero padding comes from the halo rows, and no im2col copy exists.
// M tiles of 128 consecutive padded rows cover the interior (15 tiles for G = 11; the results of halo
// rows are ignored). Eight input channels (one MMA K step) are resident at a time: 4 passes per volume,
// 27 taps x 15 tiles x 2 MMAs each (Vh [Wh | Wl], N = 32; Vl Wh, N = 16), accumulating into 15 x 32 TMEM columns.
//
// conv2 / conv3 / softmax / weighted sum follow in the same CTA on the CUDA cores (they are 12 % of the
// arithmetic) with the volume in shared memory, as in cpg_fused_kernel (cpg.cu).
//
// Two kernels share this file: cpg_tc_kernel (the form above: one tap per MMA pair, DVCP_CPG_TC) and
// cpg_tcz_kernel (further down: the three z taps of a column as the N dimension, a third of the operand
// fetches; DVCP_CPG_TCZ, what dvcp_cpg picks). tools/cpg_timing.py prints the per-phase cycle counts of both
// (library built with -DDVCP_CPG_TIMING).
//
// The target embedding is read as flat[c' * 32 + f'] (the LOGICAL [32, C] order, `layout` 0 of dvcp_cpg):
// 32 contiguous floats per voxel. DeepVCP.match has the embedding kernel write that order directly.
#include "common.cuh"

namespace dvcp {

#ifdef DVCP_CPG_TIMING
__device__ long long g_ct_time[16];
#define CT_TICK(slot)                                                   \
    do {                                                                \
        if (blockIdx.x == 0 && tid == 0) {                              \
            const long long now__ = clock64();                          \
            g_ct_time[slot] += now__ - t_last__;                        \
            t_last__ = now__;                                           \
        }                                                               \
    } while (0)
#else
#define CT_TICK(slot)
#endif

constexpr int CT_MAXG = 11;
constexpr int CT_B_FLOATS_Q = 27 * 2 * 32 * 4;   // per channel octet: [tap][k half][hi cout 0..15 | lo cout 0..15][4]
constexpr int CT_B_FLOATS = 4 * CT_B_FLOATS_Q;

__device__ __forceinline__ uint32_t ct_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: LBO = distance of the two 16-byte K halves, SBO = distance of 8-row groups
__device__ __forceinline__ uint64_t ct_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = 16 / 32
constexpr uint32_t CT_IDESC16 = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t CT_IDESC32 = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void ct_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void ct_mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "CT_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra CT_WAIT_DONE;\n\t"
        "bra CT_WAIT_LOOP;\n\t"
        "CT_WAIT_DONE:\n\t}" ::"r"(ct_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool ct_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float ct_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// conv1 weights [16][32][27] -> the N-side operand image, split hi / lo; the 32 rows of a tap are
// [hi of cout 0..15 | lo of cout 0..15], so ONE MMA forms Vh * [Wh | Wl] (the operand Vh is fetched from shared
// memory once for both terms: the MMAs of this kernel are bound by their operand reads, N being only 16):
//   image[q][tap][kh][h * 16 + n][e] = part_h( w1[n][8 q + 4 kh + e][tap] )
__global__ void cpg_tc_prepare_kernel(const float *__restrict__ w1, float *__restrict__ image) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CT_B_FLOATS) return;
    const int e = i & 3, n = (i >> 2) & 15, h = (i >> 6) & 1, kh = (i >> 7) & 1, tap = (i >> 8) % 27, q = (i >> 8) / 27;
    const float w = __ldg(w1 + (n * 32 + 8 * q + 4 * kh + e) * 27 + tap);
    const float hi = ct_hi(w);
    image[i] = h == 0 ? hi : w - hi;
}

template <int TH>
__device__ __forceinline__ float ct_block_sum(float v, float *red, int tid) {
#pragma unroll
    for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < TH / 32; ++w) t += red[w];
    return t;
}

// conv2 (16 -> 4) of one volume from shared memory. ZV = z values per item (their input window stays in registers),
// CIS = lanes that share one item's input channels (their partial sums are combined by shuffles).
template <int TH, int ZV, int CIS>
__device__ __forceinline__ void ct_conv2(const float *A1, float *O2, const float *W2, const float *__restrict__ b2g, int G,
                                         int Cp, int tid) {
    const int NG = (G + ZV - 1) / ZV, items = G * G * NG * CIS;
    const float4 bias = make_float4(__ldg(b2g), __ldg(b2g + 1), __ldg(b2g + 2), __ldg(b2g + 3));
    for (int it0 = 0; it0 < items; it0 += TH) {   // uniform trip count: the lanes of a warp shuffle together
        const int it = it0 + tid;
        const bool act = it < items;
        const int part = it % CIS, vi = act ? it / CIS : 0;
        const int line = vi / NG, zg = vi - line * NG;
        const int x = line / G, y = line - x * G, z0 = zg * ZV;
        float2 a2[ZV][2];
#pragma unroll
        for (int v = 0; v < ZV; ++v) {
            a2[v][0] = part == 0 ? make_float2(bias.x, bias.y) : make_float2(0.f, 0.f);
            a2[v][1] = part == 0 ? make_float2(bias.z, bias.w) : make_float2(0.f, 0.f);
        }
        if (act) {
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= G) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = y + dy;
                    if (yy < 0 || yy >= G) continue;
                    const int tap0 = ((dx + 1) * 3 + (dy + 1)) * 3;
                    const float *col = A1 + (xx * G + yy) * G + z0 - 1;
#pragma unroll 4
                    for (int cj = 0; cj < 16 / CIS; ++cj) {
                        const int ci = part * (16 / CIS) + cj;
                        float in[ZV + 2];
#pragma unroll
                        for (int k = 0; k < ZV + 2; ++k) {
                            const int z = z0 - 1 + k;
                            in[k] = (z >= 0 && z < G) ? col[ci * Cp + k] : 0.f;
                        }
#pragma unroll
                        for (int dz = 0; dz < 3; ++dz) {
                            const float4 w = *reinterpret_cast<const float4 *>(W2 + ((tap0 + dz) * 16 + ci) * 4);
#pragma unroll
                            for (int v = 0; v < ZV; ++v) {
                                const float2 x2 = make_float2(in[v + dz], in[v + dz]);
                                a2[v][0] = __ffma2_rn(make_float2(w.x, w.y), x2, a2[v][0]);
                                a2[v][1] = __ffma2_rn(make_float2(w.z, w.w), x2, a2[v][1]);
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int sft = 1; sft < CIS; sft <<= 1) {
#pragma unroll
            for (int v = 0; v < ZV; ++v) {
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    a2[v][o].x += __shfl_xor_sync(0xffffffffu, a2[v][o].x, sft);
                    a2[v][o].y += __shfl_xor_sync(0xffffffffu, a2[v][o].y, sft);
                }
            }
        }
        if (act && part == 0) {
            const int c0 = (x * G + y) * G + z0;
#pragma unroll
            for (int v = 0; v < ZV; ++v)
                if (z0 + v < G) {
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        O2[(2 * o) * Cp + c0 + v] = a2[v][o].x;
                        O2[(2 * o + 1) * Cp + c0 + v] = a2[v][o].y;
                    }
                }
        }
    }
}

// conv2 (16 -> 4), conv3 (4 -> 1), softmax over the voxels and the weighted candidate sum of ONE volume, on the
// CUDA cores with the volume in shared memory: A1 = conv1 out [16][Cp] (bias added), O2 = conv2 out [4][Cp], LG = logits.
#ifdef DVCP_CPG_TIMING
#define CT_TAIL_TIME_PARAM , long long &t_last__
#define CT_TAIL_TIME_ARG , t_last__
#else
#define CT_TAIL_TIME_PARAM
#define CT_TAIL_TIME_ARG
#endif
template <int TH>
__device__ __forceinline__ void ct_tail(const float *A1, float *O2, float *LG, const float *W2, const float *W3,
                                        const dvcp_cpg_params_t &p, float b3, int G, int Cp, int C, int64_t m,
                                        const float *__restrict__ cand, float *__restrict__ vcp,
                                        float *__restrict__ logits_out, float *red, int tid CT_TAIL_TIME_PARAM) {
    const int warp = tid >> 5, lane = tid & 31;
    // ---- conv2 16 -> 4, packed FFMA2. Large grids: thread = (line (x, y), group of 3 z) with the z window in registers
    //      (484 of the 512 threads at G = 11). Small grids (G <= 8) have too few such items for the CTA: thread =
    //      (voxel, half of the input channels) at G <= 5, = voxel otherwise ----
    if (G > 8)
        ct_conv2<TH, 3, 1>(A1, O2, W2, p.b2, G, Cp, tid);
    else if (G > 5)
        ct_conv2<TH, 1, 1>(A1, O2, W2, p.b2, G, Cp, tid);
    else
        ct_conv2<TH, 1, 2>(A1, O2, W2, p.b2, G, Cp, tid);
    __syncthreads();
    CT_TICK(4);
    // ---- conv3 4 -> 1: thread per voxel, one accumulator per input channel (four independent FMA chains) ----
    for (int c = tid; c < C; c += TH) {
        const int iz = c % G, iy = (c / G) % G, ix = c / (G * G);
        float a3[4] = {b3, 0.f, 0.f, 0.f};
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = ix + dx;
            if (xx < 0 || xx >= G) continue;
#pragma unroll
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
                    a3[0] = fmaf(w.x, O2[cc], a3[0]);
                    a3[1] = fmaf(w.y, O2[Cp + cc], a3[1]);
                    a3[2] = fmaf(w.z, O2[2 * Cp + cc], a3[2]);
                    a3[3] = fmaf(w.w, O2[3 * Cp + cc], a3[3]);
                }
            }
        }
        const float lg = (a3[0] + a3[1]) + (a3[2] + a3[3]);
        LG[c] = lg;
        if (logits_out) logits_out[m * C + c] = lg;
    }
    __syncthreads();
    CT_TICK(5);
    // ---- softmax over the C voxels + weighted candidate sum ----
    {
        float mx = -INFINITY;
        for (int c = tid; c < C; c += TH) mx = fmaxf(mx, LG[c]);
#pragma unroll
        for (int s = 16; s; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
        __syncthreads();
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        mx = red[0];
#pragma unroll
        for (int w = 1; w < TH / 32; ++w) mx = fmaxf(mx, red[w]);
        float zp = 0.f;
        for (int c = tid; c < C; c += TH) zp += expf(LG[c] - mx);
        const float Z = ct_block_sum<TH>(zp, red, tid);
        const float *cp = cand + m * C * 3;
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = tid; c < C; c += TH) {
            const float w = expf(LG[c] - mx) / Z;
            a[0] = fmaf(w, __ldg(cp + 3 * c), a[0]);
            a[1] = fmaf(w, __ldg(cp + 3 * c + 1), a[1]);
            a[2] = fmaf(w, __ldg(cp + 3 * c + 2), a[2]);
            a[3] += w;
        }
        const float sx = ct_block_sum<TH>(a[0], red, tid), sy = ct_block_sum<TH>(a[1], red, tid),
                    sz = ct_block_sum<TH>(a[2], red, tid), sw = ct_block_sum<TH>(a[3], red, tid);
        if (tid == 0) {
            vcp[m * 3] = sx / sw;
            vcp[m * 3 + 1] = sy / sw;
            vcp[m * 3 + 2] = sz / sw;
        }
    }
}

// TH threads per CTA: 512 with one CTA per SM for the large volumes; 256 with TWO CTAs per SM for small grids (G <= 6),
// whose volumes are a chain of short phases separated by barriers -- two co-resident CTAs fill each other's gaps.
template <int TH>
__global__ void __launch_bounds__(TH, TH == 512 ? 1 : 2)
cpg_tc_kernel(const float *__restrict__ src, const float *__restrict__ tgt, const float *__restrict__ cand, int64_t M,
              int G, int R, unsigned tmem_cols, const float *__restrict__ bimage, dvcp_cpg_params_t p,
              float *__restrict__ vcp, float *__restrict__ logits_out) {
    extern __shared__ __align__(128) unsigned char ct_smem[];
    __shared__ float red[TH / 32];
    __shared__ __align__(16) float s_src[32];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Gp = G + 2, Gp2 = Gp * Gp, C = G * G * G, Cp = (C + 3) & ~3;
    const int m_lo = Gp2 + Gp + 1;                             // first interior row of the padded volume
    const int ntiles = ((G - 1) * m_lo + 1 + 127) >> 7;        // M tiles of 128 padded rows over the interior span
    const int nissue = ntiles < TH / 32 ? ntiles : TH / 32;   // warps that issue MMAs (one commit each)
    // shared memory: A planes [hi k0 | hi k1 | lo k0 | lo k1][R rows][16 B]; B image of the current octet; W2; W3
    float4 *sA = reinterpret_cast<float4 *>(ct_smem);
    float *sB = reinterpret_cast<float *>(ct_smem + (size_t)4 * R * 16);
    float *W2 = sB + 2 * CT_B_FLOATS_Q;   // [27][16][4] (two B buffers before it)
    float *W3 = W2 + 27 * 16 * 4;      // [27][4] (padded to 112)
    // after conv1 the A region is dead and holds conv1 out [16][Cp], conv2 out [4][Cp], logits [Cp]
    float *A1 = reinterpret_cast<float *>(ct_smem), *O2 = A1 + 16 * Cp, *LG = O2 + 4 * Cp;

    for (int i = tid; i < 27 * 16 * 4; i += TH) {
        const int co = i & 3, ci = (i >> 2) & 15, tap = i >> 6;
        W2[i] = __ldg(p.w2 + (co * 16 + ci) * 27 + tap);
    }
    for (int i = tid; i < 27 * 4; i += TH) W3[i] = __ldg(p.w3 + (i & 3) * 27 + (i >> 2));
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ct_smem_u32(&s_bar)), "r"((unsigned)nissue));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ct_smem_u32(&s_tmem)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    const uint32_t a_base = ct_smem_u32(sA), b_base = ct_smem_u32(sB);
    const uint32_t plane = (uint32_t)R * 16u;   // bytes of one A plane = LBO of the A operand
    const float b3 = __ldg(p.b3);
    unsigned phase = 0;
#ifdef DVCP_CPG_TIMING
    long long t_last__ = clock64();
    if (blockIdx.x == 0 && tid == 0) for (int i = 0; i < 16; ++i) g_ct_time[i] = 0;
#endif

    for (int64_t m = blockIdx.x; m < M; m += gridDim.x) {
        const float *t = tgt + m * 32 * (int64_t)C;
        __syncthreads();   // the previous volume is finished with the shared volumes
        CT_TICK(7);
        // zero the A planes (halo rows must read as zero; the conv stages of the previous volume overwrote them)
        for (int i = tid; i < 4 * R; i += TH) sA[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < 32) s_src[tid] = __ldg(src + m * 32 + tid);
        __syncthreads();
        // software pipeline over the four channel octets: the weight image of octet q+1 (cp.async into the other
        // B buffer) and its target values (registers) are fetched while the tensor core works on octet q
        constexpr int CT_VPT = TH == 512 ? 3 : 1;   // voxels per thread: ceil(11^3 / 512); 6^3 <= 256
        float4 pt0[CT_VPT], pt1[CT_VPT];
        auto fetch_tgt = [&](int q) {
#pragma unroll
            for (int u = 0; u < CT_VPT; ++u) {
                const int c = tid + u * TH;
                if (c < C) {
                    pt0[u] = __ldg(reinterpret_cast<const float4 *>(t + (int64_t)c * 32 + 8 * q));
                    pt1[u] = __ldg(reinterpret_cast<const float4 *>(t + (int64_t)c * 32 + 8 * q + 4));
                }
            }
        };
        auto fetch_b = [&](int q) {
            const float4 *bi = reinterpret_cast<const float4 *>(bimage + (size_t)q * CT_B_FLOATS_Q);
            const uint32_t dst = ct_smem_u32(sB) + (uint32_t)(q & 1) * (CT_B_FLOATS_Q * 4);
            for (int i = tid; i < CT_B_FLOATS_Q / 4; i += TH)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)i * 16u), "l"(bi + i) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        fetch_b(0);
        fetch_tgt(0);
        CT_TICK(0);
        for (int q = 0; q < 4; ++q) {
            // ---- cost volume of channels 8q .. 8q+7: cost[c', f'] = (src[f'] - flat[c' * 32 + f'])^2 ----
            const float4 s0 = *reinterpret_cast<const float4 *>(s_src + 8 * q), s1 = *reinterpret_cast<const float4 *>(s_src + 8 * q + 4);
#pragma unroll
            for (int u = 0; u < CT_VPT; ++u) {
                const int c = tid + u * TH;
                if (c >= C) continue;
                const float4 t0 = pt0[u], t1 = pt1[u];
                const int z = c % G, y = (c / G) % G, x = c / (G * G);
                const int r = ((x + 1) * Gp + (y + 1)) * Gp + (z + 1);
                float4 v0, v1, h0, h1;
                v0.x = (s0.x - t0.x) * (s0.x - t0.x); v0.y = (s0.y - t0.y) * (s0.y - t0.y);
                v0.z = (s0.z - t0.z) * (s0.z - t0.z); v0.w = (s0.w - t0.w) * (s0.w - t0.w);
                v1.x = (s1.x - t1.x) * (s1.x - t1.x); v1.y = (s1.y - t1.y) * (s1.y - t1.y);
                v1.z = (s1.z - t1.z) * (s1.z - t1.z); v1.w = (s1.w - t1.w) * (s1.w - t1.w);
                h0 = make_float4(ct_hi(v0.x), ct_hi(v0.y), ct_hi(v0.z), ct_hi(v0.w));
                h1 = make_float4(ct_hi(v1.x), ct_hi(v1.y), ct_hi(v1.z), ct_hi(v1.w));
                sA[r] = h0;
                sA[R + r] = h1;
                sA[2 * R + r] = make_float4(v0.x - h0.x, v0.y - h0.y, v0.z - h0.z, v0.w - h0.w);
                sA[3 * R + r] = make_float4(v1.x - h1.x, v1.y - h1.y, v1.z - h1.z, v1.w - h1.w);
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");   // this octet's weight image has landed
            if (q < 3) {
                fetch_b(q + 1);     // the other B buffer: its last readers (octet q-1's MMAs) were waited for
                fetch_tgt(q + 1);
            }
            // operands were written with ordinary stores: make them visible to the tensor core's (async) proxy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            CT_TICK(1);
            // ---- MMA issue: warp w owns the accumulator tiles w, w + 16, ... (independent accumulators, so the
            //      issue work -- one thread can only feed the tensor core every few cycles -- is spread over the
            //      warps); the loop is warp-uniform so the descriptors live in uniform registers, one elected
            //      lane issues ----
            if (warp < nissue) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const bool leader = ct_elect_one();
                for (int tile = warp; tile < ntiles; tile += TH / 32) {
                    const uint32_t d = tmem_base + (uint32_t)tile * 32u;   // columns 0..15: Vh Wh + Vl Wh, 16..31: Vh Wl
                    const uint32_t row0 = (uint32_t)(m_lo + tile * 128);
                    const uint64_t a0 = ct_desc(a_base + row0 * 16u, plane, 128u);
                    const uint64_t w0 = ct_desc(b_base + (uint32_t)(q & 1) * (CT_B_FLOATS_Q * 4), 512u, 128u);
#pragma unroll
                    for (int tap = 0; tap < 27; ++tap) {
                        const int dx = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dz = tap % 3 - 1;
                        // the address field counts 16-byte units = rows: a tap shifts the descriptor by a constant
                        const uint64_t a_hi = a0 + (uint64_t)(int64_t)(dx * Gp2 + dy * Gp + dz);
                        const uint64_t a_lo = a_hi + (uint64_t)(2u * (plane >> 4));
                        const uint64_t w = w0 + (uint64_t)(tap * 64);   // 1024 B per tap
                        if (leader) {
                            ct_mma(d, a_hi, w, CT_IDESC32, (q | tap) != 0);   // Vh * [Wh | Wl]
                            ct_mma(d, a_lo, w, CT_IDESC16, 1u);               // Vl * Wh (the first 16 rows)
                        }
                    }
                }
                if (leader)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ct_smem_u32(&s_bar))
                                 : "memory");
                __syncwarp();
            }
            // everybody waits for the MMAs of this octet: the operands are rewritten next
            ct_mbar_wait(&s_bar, phase);
            phase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            CT_TICK(2);
        }
        __syncthreads();   // every thread has passed the wait: the A region may be reused
        // ---- conv1 accumulators: TMEM -> + bias -> shared [16][Cp] ----
        {
            const int lq = warp & 3;   // a warp reads the TMEM lanes 32 (warp % 4) .. + 31
            for (int tile = warp >> 2; tile < ntiles; tile += TH / 128) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)tile * 32u;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                      "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                      "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                      "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int r = m_lo + tile * 128 + lq * 32 + lane;
                const int zp = r % Gp, yp = (r / Gp) % Gp, xp = r / Gp2;
                if (xp >= 1 && xp <= G && yp >= 1 && yp <= G && zp >= 1 && zp <= G) {
                    const int c = ((xp - 1) * G + (yp - 1)) * G + (zp - 1);
#pragma unroll
                    for (int o = 0; o < 16; ++o)
                        A1[o * Cp + c] = (__uint_as_float(v[o]) + __uint_as_float(v[16 + o])) + __ldg(p.b1 + o);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        __syncthreads();
        CT_TICK(3);
        ct_tail<TH>(A1, O2, LG, W2, W3, p, b3, G, Cp, C, m, cand, vcp, logits_out, red, tid CT_TAIL_TIME_ARG);
        CT_TICK(6);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols)
                     : "memory");
    }
}

// rows of one A plane: the last tile's last row plus the largest tap shift, rounded up to 8
static inline int ct_rows(int G) {
    const int Gp = G + 2, m_lo = Gp * Gp + Gp + 1;
    const int ntiles = ((G - 1) * m_lo + 1 + 127) / 128;
    return ((m_lo + ntiles * 128 + m_lo) + 7) & ~7;
}

int cpg_tc_launch(const float *src_dfe, const float *tgt_dfe, const float *cand, int64_t M, int G, dvcp_cpg_params_t p,
                  float *vcp, float *logits, float *image, cudaStream_t st) {
    if (G < 2 || G > CT_MAXG) return DVCP_E_UNSUPPORTED;
    cpg_tc_prepare_kernel<<<(CT_B_FLOATS + 255) / 256, 256, 0, st>>>(p.w1, image);
    DVCP_CHECK_LAUNCH();
    const int R = ct_rows(G);
    const int smem = 4 * R * 16 + (2 * CT_B_FLOATS_Q + 27 * 16 * 4 + 112) * (int)sizeof(float);
    const int Gp = G + 2, m_lo = Gp * Gp + Gp + 1, ntiles = ((G - 1) * m_lo + 1 + 127) / 128;
    unsigned tmem_cols = 32;
    while ((int)tmem_cols < ntiles * 32) tmem_cols *= 2;   // allocation: a power of two >= 32
    if (G <= 6) {
        // small volumes (6^3 is the reference's own grid): 256-thread CTAs, two per SM (2 x smem <= 227 KB, 2 x TMEM <= 512)
        DVCP_CUDA(cudaFuncSetAttribute(cpg_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int64_t grid = M < 2 * DVCP_NUM_SMS ? M : 2 * DVCP_NUM_SMS;
        cpg_tc_kernel<256><<<(unsigned)grid, 256, smem, st>>>(src_dfe, tgt_dfe, cand, M, G, R, tmem_cols, image, p, vcp, logits);
    } else {
        DVCP_CUDA(cudaFuncSetAttribute(cpg_tc_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int64_t grid = M < DVCP_NUM_SMS ? M : DVCP_NUM_SMS;
        cpg_tc_kernel<512><<<(unsigned)grid, 512, smem, st>>>(src_dfe, tgt_dfe, cand, M, G, R, tmem_cols, image, p, vcp, logits);
    }
    DVCP_CHECK_LAUNCH();
    return 0;
}


// =====================================================================================================================
// Second form of conv1 ("z taps in N"): the MMAs of cpg_tc_kernel are bound by their operand reads from shared memory
// (N is only 16 output channels, and every one of the 27 taps fetches the 128 x 8 operand rows again: 9.5 KB per tap,
// tile and channel octet; the MMA phase is 61 % of the kernel, tools/cpg_timing.py). Here the three taps of a z column
// share one fetch of the volume: with z the fastest coordinate of the padded volume, tap (dx, dy, dz) of output row r
// reads row r + (dx Gp + dy) Gp + dz, so
//     Y_dz[r'] = sum over (dx, dy, cin) of V[r' + (dx Gp + dy) Gp, cin] * W1[., cin, (dx, dy, dz)]      (one GEMM, N = 3 x 16)
//     out[r]   = Y_-1[r - 1] + Y_0[r] + Y_+1[r + 1]                                                     (epilogue)
// i.e. 9 descriptor shifts instead of 27, each MMA pair forming Vh * [Wh(dz = -1, 0, +1) | Wl(dz = -1, 0, +1)]
// (N = 96) and Vl * Wh (N = 48): 12.5 KB of operand reads per THREE taps instead of 28.5 KB. The accumulators take
// 96 TMEM columns per tile of 128 rows, so the tiles of a volume are processed in groups of <= 5 (480 columns): per
// group the rows it needs (its tiles + the (dx, dy) halo on both sides) are staged octet after octet, and the
// epilogue adds the three z-shifted column groups into the conv1 output in shared memory (three passes, one per dz,
// so that no two threads ever update the same element at the same time).
// =====================================================================================================================
constexpr uint32_t CT_IDESC96 = (1u << 4) | (2u << 7) | (2u << 10) | ((96u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t CT_IDESC48 = (1u << 4) | (2u << 7) | (2u << 10) | ((48u >> 3) << 17) | ((128u >> 4) << 24);

// weight image of the z-taps-in-N form: image[q][sh = (dx, dy)][kh][n = h * 48 + dz * 16 + cout][e]
//   = part_h( w1[cout][8 q + 4 kh + e][tap = sh * 3 + dz] ); the first 48 rows of a (sh, kh) block are the hi parts
__global__ void cpg_tcz_prepare_kernel(const float *__restrict__ w1, float *__restrict__ image) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CT_B_FLOATS) return;
    const int e = i & 3, n = (i >> 2) % 96, kh = (i / 384) & 1, sh = (i / 768) % 9, q = i / 6912;
    const int h = n / 48, dz = (n % 48) >> 4, co = n & 15;
    const float w = __ldg(w1 + (co * 32 + 8 * q + 4 * kh + e) * 27 + sh * 3 + dz);
    const float hi = ct_hi(w);
    image[i] = h == 0 ? hi : w - hi;
}

// interior voxels (x, y, z in 1..G of the padded volume) whose padded row index is below rho
__device__ __forceinline__ int ct_voxels_below(int rho, int G, int Gp, int Gp2) {
    if (rho <= 0) return 0;
    const int x = rho / Gp2, rem = rho - x * Gp2, y = rem / Gp, z = rem - y * Gp;
    int n = min(max(x - 1, 0), G) * G * G;
    if (x >= 1 && x <= G) {
        n += min(max(y - 1, 0), G) * G;
        if (y >= 1 && y <= G) n += min(max(z - 1, 0), G);
    }
    return n;
}

__device__ __forceinline__ void ct_tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

__device__ __forceinline__ void ct_tmem_ld8(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

// TG = tiles per group (<= 5: 96 TMEM columns each), Rg = rows of one A plane (TG * 128 + 2 * halo, rounded up to 8).
// Warp roles inside the step loop: the LAST warp issues the MMAs (tcgen05.mma queues a few instructions and then
// blocks its issuer: a warp that also staged operands would hold every barrier up), the other warps stage the
// operands; the hand-over is by mbarriers in both directions, no block-wide barrier inside a group.
template <int TH>
__global__ void __launch_bounds__(TH, TH == 512 ? 1 : 2)
cpg_tcz_kernel(const float *__restrict__ src, const float *__restrict__ tgt, const float *__restrict__ cand, int64_t M,
               int G, int Rg, int TG, unsigned tmem_cols, const float *__restrict__ bimage, dvcp_cpg_params_t p,
               float *__restrict__ vcp, float *__restrict__ logits_out) {
    extern __shared__ __align__(128) unsigned char ct_smem[];
    constexpr int NW = TH - 32;   // staging threads (warps 0 .. TH / 32 - 2)
    __shared__ float red[TH / 32];
    __shared__ __align__(16) float s_src[32];
    __shared__ float s_b1[16];
    // s_done_h / s_done_l: the Vh / Vl MMAs of a step have completed (their planes may be rewritten);
    // s_full_h / s_full_l: the hi / lo planes of a step are staged (one arrival per staging warp)
    __shared__ __align__(8) uint64_t s_done_h, s_done_l, s_full_h, s_full_l;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool mma_warp = warp == TH / 32 - 1;
    const int Gp = G + 2, Gp2 = Gp * Gp, C = G * G * G, Cp = (C + 3) & ~3;
    const int m_lo = Gp2 + Gp + 1;                             // first interior row of the padded volume
    const int H = Gp2 + Gp;                                    // largest (dx, dy) shift = halo rows on either side
    const int ntiles = ((G - 1) * m_lo + 1 + 127) >> 7;        // M tiles of 128 padded rows over the interior span
    const int ngroups = (ntiles + TG - 1) / TG;
    // shared memory: A planes [hi k0 | hi k1 | lo k0 | lo k1][Rg rows][16 B]; two B images; W2; W3; conv1 out [16][Cp]
    float4 *sA = reinterpret_cast<float4 *>(ct_smem);
    float *sB = reinterpret_cast<float *>(ct_smem + (size_t)4 * Rg * 16);
    float *W2 = sB + 2 * CT_B_FLOATS_Q;   // [27][16][4]
    float *W3 = W2 + 27 * 16 * 4;         // [27][4] (padded to 112)
    float *A1 = W3 + 112;                 // conv1 out [16][Cp]
    // after conv1 the A region is dead and holds conv2 out [4][Cp], logits [Cp]
    float *O2 = reinterpret_cast<float *>(ct_smem), *LG = O2 + 4 * Cp;

    for (int i = tid; i < 27 * 16 * 4; i += TH) {
        const int co = i & 3, ci = (i >> 2) & 15, tap = i >> 6;
        W2[i] = __ldg(p.w2 + (co * 16 + ci) * 27 + tap);
    }
    for (int i = tid; i < 27 * 4; i += TH) W3[i] = __ldg(p.w3 + (i & 3) * 27 + (i >> 2));
    if (tid < 16) s_b1[tid] = __ldg(p.b1 + tid);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ct_smem_u32(&s_done_h)), "r"(1u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ct_smem_u32(&s_done_l)), "r"(1u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ct_smem_u32(&s_full_h)), "r"((unsigned)(NW / 32)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ct_smem_u32(&s_full_l)), "r"((unsigned)(NW / 32)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ct_smem_u32(&s_tmem)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    const uint32_t a_base = ct_smem_u32(sA), b_base = ct_smem_u32(sB);
    const uint32_t plane = (uint32_t)Rg * 16u;   // bytes of one A plane = LBO of the A operand
    const float b3 = __ldg(p.b3);
    unsigned ph_done_h = 0, ph_done_l = 0, ph_full = 0;   // every thread waits for every completion of the done barriers
#ifdef DVCP_CPG_TIMING
    long long t_last__ = clock64();
    if (blockIdx.x == 0 && tid == 0) for (int i = 0; i < 16; ++i) g_ct_time[i] = 0;
#endif

    for (int64_t m = blockIdx.x; m < M; m += gridDim.x) {
        const float *t = tgt + m * 32 * (int64_t)C;
        __syncthreads();   // the previous volume is finished with the shared volumes
        CT_TICK(7);
        if (tid < 32) s_src[tid] = __ldg(src + m * 32 + tid);
        for (int o = 0; o < 16; ++o) {   // the epilogues add into the bias
            const float bo = s_b1[o];
            for (int i = tid; i < Cp / 4; i += TH) reinterpret_cast<float4 *>(A1 + o * Cp)[i] = make_float4(bo, bo, bo, bo);
        }
        // Software pipeline over the steps (group g, channel octet q), in two halves: the Vh MMAs of a step read the
        // hi planes only and the Vl MMAs the lo planes only, so the hi planes of the NEXT octet are written as soon as
        // this step's Vh MMAs have completed (while its Vl MMAs run), and the lo planes when the Vl MMAs have completed
        // (while the next Vh MMAs run): the tensor core always has a half step queued while the CUDA cores stage.
        // The weight image of the next step (cp.async into the other B buffer) and its target values (registers) are
        // fetched a step ahead.
        constexpr int CT_VPT = TH == 512 ? 2 : 1;   // voxels per staging thread and step (checked by the launcher)
        float4 pt0[CT_VPT], pt1[CT_VPT], l0[CT_VPT], l1[CT_VPT];
        int rl[CT_VPT];           // plane row of my voxels (-1: none)
        int c_lo = 0, c_hi = 0;   // the interior voxels whose rows the current group stages
        auto group_voxels = [&](int g, int &lo, int &hi) {
            const int rowbase = 128 * g * TG + 1;   // = m_lo + 128 g TG - H: padded row of the planes' local row 0
            const int nt = min(TG, ntiles - g * TG);
            lo = ct_voxels_below(rowbase, G, Gp, Gp2);
            hi = ct_voxels_below(rowbase + nt * 128 + 2 * H, G, Gp, Gp2);
        };
        auto fetch_tgt = [&](int q, int lo, int hi) {
#pragma unroll
            for (int u = 0; u < CT_VPT; ++u) {
                const int c = lo + tid + u * NW;
                if (c < hi) {
                    pt0[u] = __ldg(reinterpret_cast<const float4 *>(t + (int64_t)c * 32 + 8 * q));
                    pt1[u] = __ldg(reinterpret_cast<const float4 *>(t + (int64_t)c * 32 + 8 * q + 4));
                }
            }
        };
        auto fetch_b = [&](int step) {   // staging threads only
            const float4 *bi = reinterpret_cast<const float4 *>(bimage + (size_t)(step & 3) * CT_B_FLOATS_Q);
            const uint32_t dst = ct_smem_u32(sB) + (uint32_t)(step & 1) * (CT_B_FLOATS_Q * 4);
            for (int i = tid; i < CT_B_FLOATS_Q / 4; i += NW)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)i * 16u), "l"(bi + i) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (!mma_warp) {
            fetch_b(0);
            group_voxels(0, c_lo, c_hi);
            fetch_tgt(0, c_lo, c_hi);
        }
        CT_TICK(0);
        for (int g = 0; g < ngroups; ++g) {
            const int rowbase = 128 * g * TG + 1;
            const int nt = min(TG, ntiles - g * TG);
            // zero the A planes: halo rows must read as zero (the interior rows are rewritten by every octet)
            for (int i = tid; i < 4 * Rg; i += TH) sA[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!mma_warp) {
#pragma unroll
                for (int u = 0; u < CT_VPT; ++u) {
                    const int c = c_lo + tid + u * NW;
                    const int z = c % G, y = (c / G) % G, x = c / (G * G);
                    rl[u] = c < c_hi ? ((x + 1) * Gp + (y + 1)) * Gp + (z + 1) - rowbase : -1;
                }
            }
            __syncthreads();   // planes zeroed; the previous group's (or volume's) readers of the shared volumes are done
            CT_TICK(8);
            if (mma_warp) {
                // ================= MMA issue: one elected lane, all tiles of the group =================
                const bool leader = ct_elect_one();
                for (int q = 0; q < 4; ++q) {
                    const int step = g * 4 + q;
                    const uint64_t w0 = ct_desc(b_base + (uint32_t)(step & 1) * (CT_B_FLOATS_Q * 4), 1536u, 128u);
                    if (q > 0) {   // keeps this warp's phase bookkeeping in step with the staging warps'
                        ct_mbar_wait(&s_done_h, ph_done_h);
                        ph_done_h ^= 1;
                    }
                    ct_mbar_wait(&s_full_h, ph_full);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (leader) {
                        for (int tile = 0; tile < nt; ++tile) {
                            const uint32_t d = tmem_base + (uint32_t)tile * 96u;   // columns 0..47: (Vh + Vl) Wh, 48..95: Vh Wl
                            const uint64_t a0 = ct_desc(a_base + (uint32_t)(H + tile * 128) * 16u, plane, 128u);
#pragma unroll
                            for (int sh = 0; sh < 9; ++sh) {
                                const int dx = sh / 3 - 1, dy = sh % 3 - 1;
                                // the address field counts 16-byte units = rows: a shift moves the descriptor by a constant
                                ct_mma(d, a0 + (uint64_t)(int64_t)(dx * Gp2 + dy * Gp), w0 + (uint64_t)(sh * 192), CT_IDESC96,
                                       (q | sh) != 0);   // Vh * [Wh | Wl]
                            }
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ct_smem_u32(&s_done_h))
                                     : "memory");
                    }
                    __syncwarp();
                    if (q > 0) {
                        ct_mbar_wait(&s_done_l, ph_done_l);
                        ph_done_l ^= 1;
                    }
                    ct_mbar_wait(&s_full_l, ph_full);
                    ph_full ^= 1;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (leader) {
                        for (int tile = 0; tile < nt; ++tile) {
                            const uint32_t d = tmem_base + (uint32_t)tile * 96u;
                            const uint64_t a0 = ct_desc(a_base + (uint32_t)(H + tile * 128) * 16u, plane, 128u) + (uint64_t)(2u * (plane >> 4));
#pragma unroll
                            for (int sh = 0; sh < 9; ++sh) {
                                const int dx = sh / 3 - 1, dy = sh % 3 - 1;
                                ct_mma(d, a0 + (uint64_t)(int64_t)(dx * Gp2 + dy * Gp), w0 + (uint64_t)(sh * 192), CT_IDESC48, 1u);   // Vl * Wh
                            }
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ct_smem_u32(&s_done_l))
                                     : "memory");
                    }
                    __syncwarp();
                }
            } else {
                // ================= operand staging =================
                for (int q = 0; q < 4; ++q) {
                    const int step = g * 4 + q;
                    if (q > 0) {   // the Vh MMAs of the previous octet have read the hi planes
                        ct_mbar_wait(&s_done_h, ph_done_h);
                        ph_done_h ^= 1;
                    }
                    // ---- cost volume of channels 8q .. 8q+7: cost[c', f'] = (src[f'] - flat[c' * 32 + f'])^2 ----
                    {
                        const float4 s0 = *reinterpret_cast<const float4 *>(s_src + 8 * q), s1 = *reinterpret_cast<const float4 *>(s_src + 8 * q + 4);
#pragma unroll
                        for (int u = 0; u < CT_VPT; ++u) {
                            if (rl[u] < 0) continue;
                            const float4 t0 = pt0[u], t1 = pt1[u];
                            float4 v0, v1, h0, h1;
                            v0.x = (s0.x - t0.x) * (s0.x - t0.x); v0.y = (s0.y - t0.y) * (s0.y - t0.y);
                            v0.z = (s0.z - t0.z) * (s0.z - t0.z); v0.w = (s0.w - t0.w) * (s0.w - t0.w);
                            v1.x = (s1.x - t1.x) * (s1.x - t1.x); v1.y = (s1.y - t1.y) * (s1.y - t1.y);
                            v1.z = (s1.z - t1.z) * (s1.z - t1.z); v1.w = (s1.w - t1.w) * (s1.w - t1.w);
                            h0 = make_float4(ct_hi(v0.x), ct_hi(v0.y), ct_hi(v0.z), ct_hi(v0.w));
                            h1 = make_float4(ct_hi(v1.x), ct_hi(v1.y), ct_hi(v1.z), ct_hi(v1.w));
                            sA[rl[u]] = h0;
                            sA[Rg + rl[u]] = h1;
                            l0[u] = make_float4(v0.x - h0.x, v0.y - h0.y, v0.z - h0.z, v0.w - h0.w);
                            l1[u] = make_float4(v1.x - h1.x, v1.y - h1.y, v1.z - h1.z, v1.w - h1.w);
                        }
                    }
                    asm volatile("cp.async.wait_group 0;" ::: "memory");   // this step's weight image has landed
                    if (q < 3) fetch_tgt(q + 1, c_lo, c_hi);             // the next octet of the same voxels
                    // operands were written with ordinary stores: make them visible to the tensor core's (async) proxy
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ct_smem_u32(&s_full_h)) : "memory");
                    CT_TICK(1);
                    if (q > 0) {   // the Vl MMAs of the previous octet have read the lo planes (and the other B buffer)
                        ct_mbar_wait(&s_done_l, ph_done_l);
                        ph_done_l ^= 1;
                    }
#pragma unroll
                    for (int u = 0; u < CT_VPT; ++u) {
                        if (rl[u] < 0) continue;
                        sA[2 * Rg + rl[u]] = l0[u];
                        sA[3 * Rg + rl[u]] = l1[u];
                    }
                    // the other B buffer: last read by the previous step (waited for above; at q == 0 by the previous
                    // group's final step, drained before its epilogue)
                    if (step + 1 < 4 * ngroups) fetch_b(step + 1);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ct_smem_u32(&s_full_l)) : "memory");
                    CT_TICK(9);
                }
                // next group's voxels and their first octet: in flight during the drain and the epilogue
                if (g + 1 < ngroups) {
                    group_voxels(g + 1, c_lo, c_hi);
                    fetch_tgt(0, c_lo, c_hi);
                }
            }
            ct_mbar_wait(&s_done_h, ph_done_h);
            ph_done_h ^= 1;
            ct_mbar_wait(&s_done_l, ph_done_l);
            ph_done_l ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            CT_TICK(10);
            // ---- the group's accumulators: out[r] = Y_-1[r - 1] + Y_0[r] + Y_+1[r + 1]. A warp reads the TMEM lanes
            //      32 (warp % 4) .. + 31 = 32 consecutive rows; the neighbours' terms come by shuffle, the two terms
            //      that cross a 32-row boundary are added in a second pass after a barrier (pass B), so that no two
            //      threads update the same element at the same time. A work item is (tile, half of the 16 channels). ----
            const int lq = warp & 3;
            constexpr int CT_EI = 3;   // items per warp: 2 * TG / (TH / 128) rounded up (TG <= 5 at TH = 512, <= 2 at 256)
            float keep[CT_EI][8];      // lane 0: Y_+1 of my row (feeds row - 1); lane 31: Y_-1 of my row (feeds row + 1)
            int keep_c[CT_EI];         // voxel they feed (-1: none)
#pragma unroll
            for (int e = 0; e < CT_EI; ++e) {
                const int item = (warp >> 2) + e * (TH / 128);
                keep_c[e] = -1;
                if (item < 2 * nt) {
                    const int tile = item >> 1, o0 = (item & 1) * 8;
                    const int r = m_lo + (g * TG + tile) * 128 + lq * 32 + lane;
                    const int zp = r % Gp, yp = (r / Gp) % Gp, xp = r / Gp2;
                    uint32_t v[6][8];   // [dz][hi / lo]
                    const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(tile * 96 + o0);
#pragma unroll
                    for (int dz = 0; dz < 3; ++dz) {
                        ct_tmem_ld8(taddr + (uint32_t)(dz * 16), v[2 * dz]);
                        ct_tmem_ld8(taddr + (uint32_t)(48 + dz * 16), v[2 * dz + 1]);
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const bool inter = xp >= 1 && xp <= G && yp >= 1 && yp <= G && zp >= 1 && zp <= G;
                    const int c = ((xp - 1) * G + (yp - 1)) * G + (zp - 1);
                    float acc[8];
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        // rows outside the interior carry zeros in every column they could feed (their operand rows are halo)
                        const float ym = inter ? __uint_as_float(v[0][o]) + __uint_as_float(v[1][o]) : 0.f;   // Y_-1[r]: feeds r + 1
                        const float y0 = __uint_as_float(v[2][o]) + __uint_as_float(v[3][o]);
                        const float yp1 = inter ? __uint_as_float(v[4][o]) + __uint_as_float(v[5][o]) : 0.f;  // Y_+1[r]: feeds r - 1
                        const float from_up = __shfl_down_sync(0xffffffffu, yp1, 1);   // Y_+1[r + 1]
                        const float from_dn = __shfl_up_sync(0xffffffffu, ym, 1);      // Y_-1[r - 1]
                        acc[o] = y0 + (lane < 31 ? from_up : 0.f) + (lane > 0 ? from_dn : 0.f);
                        keep[e][o] = lane == 0 ? yp1 : ym;
                    }
                    if (inter) {
                        float *dst = A1 + o0 * Cp + c;
                        float cur[8];
#pragma unroll
                        for (int o = 0; o < 8; ++o) cur[o] = dst[o * Cp];
#pragma unroll
                        for (int o = 0; o < 8; ++o) dst[o * Cp] = cur[o] + acc[o];
                        // the boundary terms: lane 0 feeds the row below (z - 1), lane 31 the row above (z + 1), same (x, y) line
                        if (lane == 0 && zp >= 2) keep_c[e] = o0 * Cp + c - 1;
                        if (lane == 31 && zp <= G - 1) keep_c[e] = o0 * Cp + c + 1;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
#pragma unroll
            for (int e = 0; e < CT_EI; ++e) {
                if (keep_c[e] >= 0) {
                    float *dst = A1 + keep_c[e];
#pragma unroll
                    for (int o = 0; o < 8; ++o) dst[o * Cp] += keep[e][o];
                }
            }
            CT_TICK(3);
        }
        __syncthreads();
        ct_tail<TH>(A1, O2, LG, W2, W3, p, b3, G, Cp, C, m, cand, vcp, logits_out, red, tid CT_TAIL_TIME_ARG);
        CT_TICK(6);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols)
                     : "memory");
    }
}

// host side of ct_voxels_below (the launcher checks that a step's voxels fit the per-thread registers)
static inline int ct_voxels_below_host(int rho, int G) {
    const int Gp = G + 2, Gp2 = Gp * Gp;
    if (rho <= 0) return 0;
    const int x = rho / Gp2, rem = rho - x * Gp2, y = rem / Gp, z = rem - y * Gp;
    auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
    int n = clampi(x - 1, 0, G) * G * G;
    if (x >= 1 && x <= G) {
        n += clampi(y - 1, 0, G) * G;
        if (y >= 1 && y <= G) n += clampi(z - 1, 0, G);
    }
    return n;
}

int cpg_tcz_launch(const float *src_dfe, const float *tgt_dfe, const float *cand, int64_t M, int G, dvcp_cpg_params_t p,
                   float *vcp, float *logits, float *image, cudaStream_t st) {
    if (G < 2 || G > CT_MAXG) return DVCP_E_UNSUPPORTED;
    const int Gp = G + 2, m_lo = Gp * Gp + Gp + 1, H = Gp * Gp + Gp, ntiles = ((G - 1) * m_lo + 1 + 127) / 128;
    const bool small = G <= 6;               // 256-thread CTAs, two per SM: 256 TMEM columns each
    const int tgmax = small ? 2 : 5;         // 96 columns per tile
    const int ngroups = (ntiles + tgmax - 1) / tgmax;
    const int TG = (ntiles + ngroups - 1) / ngroups;   // balanced groups
    const int Rg = (TG * 128 + 2 * H + 7) & ~7;
    const int C = G * G * G, Cp = (C + 3) & ~3;
    int maxvox = 0;
    for (int g = 0; g < ngroups; ++g) {
        const int rowbase = 128 * g * TG + 1, nt = ntiles - g * TG < TG ? ntiles - g * TG : TG;
        const int n = ct_voxels_below_host(rowbase + nt * 128 + 2 * H, G) - ct_voxels_below_host(rowbase, G);
        maxvox = n > maxvox ? n : maxvox;
    }
    if (maxvox > (small ? 224 : 960)) return DVCP_E_UNSUPPORTED;   // staging threads x voxels per thread
    cpg_tcz_prepare_kernel<<<(CT_B_FLOATS + 255) / 256, 256, 0, st>>>(p.w1, image);
    DVCP_CHECK_LAUNCH();
    const int smem = 4 * Rg * 16 + (2 * CT_B_FLOATS_Q + 27 * 16 * 4 + 112 + 16 * Cp) * (int)sizeof(float);
    unsigned tmem_cols = 32;
    while ((int)tmem_cols < TG * 96) tmem_cols *= 2;   // allocation: a power of two >= 32
    if (small) {
        DVCP_CUDA(cudaFuncSetAttribute(cpg_tcz_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int64_t grid = M < 2 * DVCP_NUM_SMS ? M : 2 * DVCP_NUM_SMS;   // small grids go with small clouds: persistent
        cpg_tcz_kernel<256><<<(unsigned)grid, 256, smem, st>>>(src_dfe, tgt_dfe, cand, M, G, Rg, TG, tmem_cols, image, p, vcp, logits);
    } else {
        DVCP_CUDA(cudaFuncSetAttribute(cpg_tcz_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int64_t grid = M < DVCP_NUM_SMS * DVCP_CPG_WAVES ? M : DVCP_NUM_SMS * DVCP_CPG_WAVES;
        cpg_tcz_kernel<512><<<(unsigned)grid, 512, smem, st>>>(src_dfe, tgt_dfe, cand, M, G, Rg, TG, tmem_cols, image, p, vcp, logits);
    }
    DVCP_CHECK_LAUNCH();
    return 0;
}

}  // namespace dvcp

extern "C" int64_t dvcp_cpg_tc_image_bytes(void) { return (int64_t)dvcp::CT_B_FLOATS * (int64_t)sizeof(float); }

#ifdef DVCP_CPG_TIMING
extern "C" __attribute__((visibility("default"))) int dvcp_debug_cpg_timing(long long *host16) {
    return (int)cudaMemcpyFromSymbol(host16, dvcp::g_ct_time, 16 * sizeof(long long));
}
#endif
