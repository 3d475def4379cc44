// Sustained FP64 throughput, SM clock and board power of the vector pipe (DFMA) against the tensor pipe (DMMA m8n8k4) on one
// B200: is an FP64 flop cheaper in energy on the tensor path?  (Decides whether a DMMA form of the k = 4 pass could run at a
// higher clock under the 1 kW cap.)  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_power fp64_power.cu
// Run next to `nvidia-smi --query-gpu=timestamp,power.draw,clocks.sm --format=csv -lms 100`; the program prints the wall-clock
// window of each phase.  Not part of the product path.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <ctime>

__global__ void __launch_bounds__(256) dfma_chains(double* out, int iters, double a, double b) {
    double c[16];
    for (int i = 0; i < 16; ++i) c[i] = 1.0 + 1e-3 * (threadIdx.x * 16 + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(a, c[i], b * (i + 1));
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) dmma884(double* out, int iters, double a, double b) {
    double c0[8][2];
    for (int i = 0; i < 8; ++i) {
        c0[i][0] = 1.0 + 1e-3 * (threadIdx.x + i);
        c0[i][1] = 1.0 - 1e-3 * i;
    }
    const double av = a * (1.0 + 1e-6 * (threadIdx.x & 31)), bv = ((threadIdx.x & 1) ? -b : b) * (1.0 + 1e-6 * (threadIdx.x & 7));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i][0]), "+d"(c0[i][1])
                         : "d"(av), "d"(bv));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c0[i][0] + c0[i][1];
    if (s == 12345.678) out[0] = s;
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
}

int main() {
    double* d;
    cudaMalloc(&d, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = 148 * 8;
    for (int which = 0; which < 2; ++which) {
        const int iters = which == 0 ? 1 << 15 : 1 << 14;
        const double flops_per_launch =
            which == 0 ? 2.0 * 16 * double(iters) * 256 * blocks : 2.0 * 8 * 8 * 4 * 8 * double(iters) * (256 / 32) * blocks;
        const double t_begin = now_s();
        double total_ms = 0;
        int launches = 0;
        while (now_s() - t_begin < 6.0) {
            cudaEventRecord(e0);
            if (which == 0)
                dfma_chains<<<blocks, 256>>>(d, iters, -0.999999, 0.5);
            else
                dmma884<<<blocks, 256>>>(d, iters, 0.37, 0.73);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            total_ms += ms;
            ++launches;
        }
        const double t_end = now_s();
        printf("{\"phase\": \"%s\", \"t_begin\": %.3f, \"t_end\": %.3f, \"launches\": %d, \"TFLOPs\": %.2f, \"err\": \"%s\"}\n",
               which == 0 ? "DFMA" : "DMMA.884", t_begin, t_end, launches, flops_per_launch * launches / total_ms / 1e9,
               cudaGetErrorString(cudaGetLastError()));
        fflush(stdout);
        // idle gap so the two windows are easy to tell apart in the power log
        struct timespec ts = {2, 0};
        nanosleep(&ts, nullptr);
    }
    return 0;
}
