// Throughput of the legacy warp-level mma.sync.m16n8k8 TF32 path on sm_100a (is it worth using for the
// CPG conv1 implicit GEMM?). Prints MACs per clock per SM for several warps-per-SM settings.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(float *out, int iters) {
    unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, threadIdx.x * 5};
    float c[4][4] = {};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
    for (int j = 0; j < 4; ++j) for (int e = 0; e < 4; ++e) s += c[j][e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 1024 * 4 * sizeof(float));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        k<<<148, warps * 32>>>(out, 100);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        k<<<148, warps * 32>>>(out, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double mmas = (double)148 * warps * iters * 4;
        const double macs = mmas * 16 * 8 * 8;
        printf("warps/SM %2d: %.3f ms, %.1f TFLOP/s tf32 (mma.sync), %.0f MAC/clk/SM at 1.965 GHz\n", warps, ms,
               2 * macs / ms / 1e9, macs / 148 / (ms * 1e-3 * 1.965e9));
    }
    return 0;
}
