// Spatial index of clouds larger than one CTA can sort (16384 < N <= 131072).
//
// Same product as the single-CTA builder in fps.cu (dvcp_cloud_index_t: points in Hilbert order as
// float4 (x, y, z, index bits), buckets of 32 with boxes), built by several CTAs per cloud:
//   1. bounding box: block reductions + 6 atomics per CTA on order-preserving integer images;
//   2. runs: each CTA sorts 16384 (key = 30-bit Hilbert code, value = point index) pairs in shared
//      memory (CUB block radix sort, the same primitive the single-CTA builder uses);
//   3. merges: log2(cap / 16384) passes; every element finds its output slot by its own offset plus
//      its rank in the partner run (binary search; lower bound for the left run, upper bound for
//      the right one, so equal keys give a permutation). No shared memory, all reads hit L2;
//   4. publish: a warp per bucket writes the float4 points and the bucket's box.
// The order among equal Hilbert codes is irrelevant to every consumer: pruning by boxes never changes
// results (members are always decided by the exact arithmetic).
//
// Consumer: dvcp_knn_indexed (knn.cu) -- SURVEY 8(d) SWEEP row, KNN at N = 32k..128k. The sampling /
// SA kernels keep the single-CTA index (N <= 16384).
#include <cub/block/block_radix_sort.cuh>

#include "common.cuh"

namespace dvcp {

constexpr int IB_RUN = 16384;   // elements per sorted run (one CTA)
constexpr int IB_THREADS = 512;
constexpr int IB_ITEMS = IB_RUN / IB_THREADS;

// order-preserving float -> unsigned (so integer atomicMin / atomicMax order floats)
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ unsigned expand10(unsigned v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// bbox[b][0..2] = ord(min), bbox[b][3..5] = ~ord(max); both start at 0xffffffff and shrink by atomicMin
__global__ void __launch_bounds__(256)
ib_bbox_kernel(Cloud c, int N, unsigned *__restrict__ bbox) {
    const int b = blockIdx.y;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = c.at(b, n, k);
            mn[k] = fminf(mn[k], v);
            mx[k] = fmaxf(mx[k], v);
        }
    }
    __shared__ unsigned s[6];
    if (threadIdx.x < 6) s[threadIdx.x] = 0xffffffffu;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        unsigned a = f2ord(mn[k]), z = ~f2ord(mx[k]);
        a = __reduce_min_sync(0xffffffffu, a);
        z = __reduce_min_sync(0xffffffffu, z);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&s[k], a);
            atomicMin(&s[3 + k], z);
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) atomicMin(&bbox[b * 6 + threadIdx.x], s[threadIdx.x]);
}

// one CTA per run of IB_RUN slots of one cloud: Hilbert keys -> sorted (key, value) run
__global__ void __launch_bounds__(IB_THREADS, 1)
ib_run_sort_kernel(Cloud c, int N, int cap, const unsigned *__restrict__ bbox, unsigned *__restrict__ keys,
                   unsigned *__restrict__ vals) {
    using Sort = cub::BlockRadixSort<unsigned, IB_THREADS, IB_ITEMS, unsigned>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.y, run = blockIdx.x, tid = threadIdx.x;
    float mn[3], ext = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        mn[k] = ord2f(bbox[b * 6 + k]);
        ext = fmaxf(ext, ord2f(~bbox[b * 6 + 3 + k]) - mn[k]);
    }
    const float scale = ext > 0.f ? 1023.0f / ext : 0.0f;   // one scale for all axes: cubic cells
    unsigned k[IB_ITEMS], v[IB_ITEMS];
#pragma unroll
    for (int i = 0; i < IB_ITEMS; ++i) {
        const int n = run * IB_RUN + tid * IB_ITEMS + i;
        if (n < N) {
            unsigned q[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float t = (c.at(b, n, a) - mn[a]) * scale;
                q[a] = (unsigned)fminf(fmaxf(t, 0.0f), 1023.0f);
            }
            k[i] = spatial_key(q[0], q[1], q[2]);
            v[i] = (unsigned)n;
        } else {
            k[i] = 0xffffffffu;   // unused slots sort behind every point
            v[i] = 0xffffffffu;
        }
    }
    Sort(*reinterpret_cast<typename Sort::TempStorage *>(smem_raw)).Sort(k, v, 0, 32);
    const int64_t o = (int64_t)b * cap + (int64_t)run * IB_RUN + tid * IB_ITEMS;
#pragma unroll
    for (int i = 0; i < IB_ITEMS; i += 4) {
        *reinterpret_cast<uint4 *>(keys + o + i) = make_uint4(k[i], k[i + 1], k[i + 2], k[i + 3]);
        *reinterpret_cast<uint4 *>(vals + o + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
}

// merges runs of length L pairwise: element at offset o of the left run goes to o + #(right < key),
// element at offset o of the right run to o + #(left <= key)
__global__ void __launch_bounds__(256)
ib_merge_kernel(const unsigned *__restrict__ kin, const unsigned *__restrict__ vin, unsigned *__restrict__ kout,
                unsigned *__restrict__ vout, int cap, int L) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const unsigned *kb = kin + (int64_t)b * cap;
    const unsigned key = kb[i];
    const int run = i / L, o = i - run * L;
    const bool right = run & 1;
    const unsigned *other = kb + (int64_t)(run ^ 1) * L;
    int lo = 0, hi = L;   // first position of `other` whose key is >= key (left) or > key (right)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const unsigned m = __ldg(other + mid);
        const bool before = right ? (m <= key) : (m < key);
        if (before) lo = mid + 1; else hi = mid;
    }
    const int64_t dst = (int64_t)b * cap + (int64_t)(run >> 1) * 2 * L + o + lo;
    kout[dst] = key;
    vout[dst] = vin[(int64_t)b * cap + i];
}

// warp per bucket: gather the points, write float4 (x, y, z, index bits) and the bucket box
__global__ void __launch_bounds__(256)
ib_publish_kernel(Cloud c, int cap, const unsigned *__restrict__ vals, dvcp_cloud_index_t index) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int bucket = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (bucket >= cap / 32) return;
    const int pos = bucket * 32 + lane;
    const unsigned n = vals[(int64_t)b * cap + pos];
    const bool valid = n != 0xffffffffu;
    float x = INFINITY, y = INFINITY, z = INFINITY;
    if (valid) {
        x = c.at(b, (int)n, 0);
        y = c.at(b, (int)n, 1);
        z = c.at(b, (int)n, 2);
    }
    reinterpret_cast<float4 *>(index.sorted_pt)[(int64_t)b * cap + pos] =
        make_float4(x, y, z, __int_as_float(valid ? (int)n : -1));
    float a0 = x, a1 = y, a2 = z;
    float z0 = valid ? x : -INFINITY, z1 = valid ? y : -INFINITY, z2 = valid ? z : -INFINITY;
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        a0 = fminf(a0, __shfl_xor_sync(0xffffffffu, a0, s));
        a1 = fminf(a1, __shfl_xor_sync(0xffffffffu, a1, s));
        a2 = fminf(a2, __shfl_xor_sync(0xffffffffu, a2, s));
        z0 = fmaxf(z0, __shfl_xor_sync(0xffffffffu, z0, s));
        z1 = fmaxf(z1, __shfl_xor_sync(0xffffffffu, z1, s));
        z2 = fmaxf(z2, __shfl_xor_sync(0xffffffffu, z2, s));
    }
    const int cnt = __popc(__ballot_sync(0xffffffffu, valid));
    if (lane == 0) {
        float4 *bb = reinterpret_cast<float4 *>(index.bucket_box + ((int64_t)b * (cap / 32) + bucket) * 8);
        bb[0] = make_float4(a0, a1, a2, z0);
        bb[1] = make_float4(z1, z2, (float)cnt, 0.f);
    }
}

}  // namespace dvcp

using namespace dvcp;

extern "C" int dvcp_index_capacity_any(int N) {
    if (N < 64 || N > 131072) return 0;
    if (N <= 16384) return dvcp_index_capacity(N);
    if (N <= 32768) return 32768;
    if (N <= 65536) return 65536;
    return 131072;
}

extern "C" int64_t dvcp_build_index_workspace_bytes(int B, int N) {
    const int cap = dvcp_index_capacity_any(N);
    if (cap <= 16384 || B <= 0) return 0;
    return (int64_t)B * cap * 4 * sizeof(unsigned) + (int64_t)B * 6 * sizeof(unsigned);
}

extern "C" int dvcp_build_index_ws(dvcp_cloud_t xyz, int B, int N, dvcp_cloud_index_t index, void *workspace,
                                   int64_t workspace_bytes, dvcp_stream_t stream) {
    if (!xyz.base || !index.sorted_pt || !index.bucket_box || B <= 0) return DVCP_E_ARG;
    const int cap = dvcp_index_capacity_any(N);
    if (cap == 0 || B > 65535) return DVCP_E_UNSUPPORTED;
    if (index.cap != cap) return DVCP_E_ARG;
    if (cap <= 16384) return dvcp_build_index(xyz, B, N, index, stream);
    if (!workspace || workspace_bytes < dvcp_build_index_workspace_bytes(B, N)) return DVCP_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned *k0 = (unsigned *)workspace, *v0 = k0 + (int64_t)B * cap, *k1 = v0 + (int64_t)B * cap,
             *v1 = k1 + (int64_t)B * cap, *bbox = v1 + (int64_t)B * cap;
    const Cloud c = as_cloud(xyz);
    DVCP_CUDA(cudaMemsetAsync(bbox, 0xff, (size_t)B * 6 * sizeof(unsigned), st));
    ib_bbox_kernel<<<dim3(32, B), 256, 0, st>>>(c, N, bbox);
    DVCP_CHECK_LAUNCH();
    using Sort = cub::BlockRadixSort<unsigned, IB_THREADS, IB_ITEMS, unsigned>;
    const size_t smem = sizeof(typename Sort::TempStorage);
    DVCP_CUDA(cudaFuncSetAttribute(ib_run_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ib_run_sort_kernel<<<dim3(cap / IB_RUN, B), IB_THREADS, smem, st>>>(c, N, cap, bbox, k0, v0);
    DVCP_CHECK_LAUNCH();
    for (int L = IB_RUN; L < cap; L *= 2) {
        ib_merge_kernel<<<dim3(cap / 256, B), 256, 0, st>>>(k0, v0, k1, v1, cap, L);
        DVCP_CHECK_LAUNCH();
        unsigned *t = k0; k0 = k1; k1 = t;
        t = v0; v0 = v1; v1 = t;
    }
    ib_publish_kernel<<<dim3(cap / 32 / 8, B), 256, 0, st>>>(c, cap, v0, index);
    DVCP_CHECK_LAUNCH();
    return 0;
}
