#include <cuda_runtime.h>
#include <cstdio>
__global__ void __launch_bounds__(256) dmma884(double* out, int iters, double a, double b) {
    double c0[8][2];
    for (int i = 0; i < 8; ++i) { c0[i][0] = threadIdx.x + i; c0[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i][0]), "+d"(c0[i][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += c0[i][0] + c0[i][1];
    if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) dmma1688(double* out, int iters, double a, double b) {
    double c0[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c0[i][j] = threadIdx.x + i + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
              : "+d"(c0[i][0]), "+d"(c0[i][1]), "+d"(c0[i][2]), "+d"(c0[i][3]) : "d"(a), "d"(a), "d"(a), "d"(a), "d"(b), "d"(b));
    }
    double s = 0; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c0[i][j];
    if (s == 12345.678) out[0] = s;
}
int main() {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 1 << 14, blocks = 148 * 8;
    for (int which = 0; which < 2; ++which) for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) dmma884<<<blocks, 256>>>(d, iters, 0.999, 1e-9); else dmma1688<<<blocks, 256>>>(d, iters, 0.999, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = which == 0 ? 2.0 * 8 * 8 * 4 * 8 : 2.0 * 16 * 8 * 8 * 4;
        flops *= double(iters) * (256 / 32) * blocks;
        printf("%s rep %d: %.3f ms %.2f TFLOP/s (%s)\n", which == 0 ? "m8n8k4" : "m16n8k8", rep, ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
