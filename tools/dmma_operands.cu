// DMMA.8x8x4 throughput against operand variety (one B200): does the 37 TFLOP/s of tools/dmma_peak.cu (every DMMA reads the
// SAME A and B registers) survive when every DMMA reads different A/B registers, as a real mat-vec does?
#include <cuda_runtime.h>
#include <cstdio>
#define DMMA(c0, c1, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b))
template <int MODE>
__global__ void __launch_bounds__(256) kern(double* out, int iters, double seed) {
    double c[4][2], a[16], b[8];
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 16; ++i) { a[i] = seed * (i + 1) + threadIdx.x * 1e-6; asm volatile("" : "+d"(a[i])); }
    for (int i = 0; i < 8; ++i) { b[i] = seed * (i + 3) * 1e-3; asm volatile("" : "+d"(b[i])); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
            for (int rb = 0; rb < 4; ++rb) {
                if (MODE == 0) DMMA(c[rb][0], c[rb][1], a[0], b[0]);                       // same A, same B
                if (MODE == 1) DMMA(c[rb][0], c[rb][1], a[0], b[s]);                       // same A, B per k-step
                if (MODE == 2) DMMA(c[rb][0], c[rb][1], a[(rb * 4 + s) & 15], b[s]);       // different A every time
                if (MODE == 3) { double na = (s & 1) ? -a[(rb * 4 + s) & 15] : a[(rb * 4 + s) & 15]; DMMA(c[rb][0], c[rb][1], na, b[s]); }
                if (MODE == 4) DMMA(c[rb][0], c[rb][1], a[(rb * 4 + s) & 15], b[0]);       // different A, same B
            }
    }
    double s = 0; for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
int main() {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 12;
    for (int warps_per_sm : {8, 16, 32, 64}) for (int mode = 0; mode < 5; ++mode) {
        const int blocks = 148 * warps_per_sm / 8;
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) kern<0><<<blocks, 256>>>(d, iters, 1e-3);
            if (mode == 1) kern<1><<<blocks, 256>>>(d, iters, 1e-3);
            if (mode == 2) kern<2><<<blocks, 256>>>(d, iters, 1e-3);
            if (mode == 3) kern<3><<<blocks, 256>>>(d, iters, 1e-3);
            if (mode == 4) kern<4><<<blocks, 256>>>(d, iters, 1e-3);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double flops = 2.0 * 256 * 32 * double(iters) * 8 * blocks;
        printf("{\"warps_per_sm\": %d, \"mode\": %d, \"ms\": %.3f, \"TFLOPs\": %.2f, \"err\": \"%s\"}\n", warps_per_sm, mode, best, flops / best / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
