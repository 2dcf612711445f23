// Microbenchmark: latency / throughput of redux.sync (CREDUX) and shuffles on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_redux_dep(unsigned *out, long long *cyc, int iters) {
    unsigned v = threadIdx.x * 2654435761u;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        unsigned m = __reduce_max_sync(0xffffffffu, v);
        v = (v ^ m) + 1u;   // dependent
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
__global__ void k_redux_indep(unsigned *out, long long *cyc, int iters) {
    unsigned v0 = threadIdx.x * 2654435761u, v1 = v0 + 7, v2 = v0 + 13, v3 = v0 + 29;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        unsigned m0 = __reduce_max_sync(0xffffffffu, v0);
        unsigned m1 = __reduce_max_sync(0xffffffffu, v1);
        unsigned m2 = __reduce_max_sync(0xffffffffu, v2);
        unsigned m3 = __reduce_max_sync(0xffffffffu, v3);
        v0 = (v0 ^ m0) + 1u; v1 = (v1 ^ m1) + 1u; v2 = (v2 ^ m2) + 1u; v3 = (v3 ^ m3) + 1u;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = v0 + v1 + v2 + v3;
}
__global__ void k_shfl_dep(unsigned *out, long long *cyc, int iters) {
    unsigned v = threadIdx.x * 2654435761u;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        unsigned m = __shfl_xor_sync(0xffffffffu, v, 1);
        v = (v ^ m) + 1u;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
__global__ void k_bar(unsigned *out, long long *cyc, int iters) {
    __shared__ unsigned s[32];
    unsigned v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v + i;
        __syncthreads();
        v += s[(threadIdx.x >> 5) ^ 1];
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
__global__ void k_lds_dep(unsigned *out, long long *cyc, int iters) {
    __shared__ unsigned s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 37 + 11) & 1023;
    __syncthreads();
    unsigned v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) v = s[v];
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}
int main() {
    unsigned *out; long long *cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    long long h[4]; const int it = 10000;
    for (int threads : {32, 128, 256, 512}) {
        k_redux_dep<<<1, threads>>>(out, cyc, it); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("redux dependent      threads %4d: %.1f cyc/iter\n", threads, (double)h[0] / it);
        k_redux_indep<<<1, threads>>>(out, cyc, it); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("redux 4 independent  threads %4d: %.1f cyc/iter (4 redux)\n", threads, (double)h[0] / it);
        k_shfl_dep<<<1, threads>>>(out, cyc, it); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("shfl dependent       threads %4d: %.1f cyc/iter\n", threads, (double)h[0] / it);
        k_bar<<<1, threads>>>(out, cyc, it); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("2x syncthreads+smem  threads %4d: %.1f cyc/iter\n", threads, (double)h[0] / it);
        k_lds_dep<<<1, threads>>>(out, cyc, it); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("lds dependent        threads %4d: %.1f cyc/iter\n", threads, (double)h[0] / it);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
