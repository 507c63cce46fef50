// Micro-benchmark: FP64 pipe throughput on B200 (DFMA / DADD / DMUL / MUFU.RCP64H per clk per SM).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(double* out, double seed, int iters) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1 + i * 0.37) + i;
    const double c = seed * 1.0000001, d = 0.999999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) a[i] = fma(a[i], d, c);
                if (OP == 1) a[i] = a[i] + c;
                if (OP == 2) a[i] = a[i] * d;
                if (OP == 3) a[i] = fmaf((float)a[i], 0.99f, 1.0f);  // mixed: F2F + FFMA + F2F
                if (OP == 4) a[i] = fmax(a[i], c) + d;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name) {
    double* out;
    const int blocks = 148 * 8, threads = 256, iters = 500;
    cudaMalloc(&out, blocks * threads * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 1.2345, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 1.2345, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double inner = (double)blocks * threads * iters * 64.0;
    printf("%-28s %8.3f ms  %7.1f stmts/clk/SM (at %d MHz nominal)\n", name, ms, inner / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1000);
    cudaFree(out);
}

int main() {
    run<0>("DFMA");
    run<1>("DADD");
    run<2>("DMUL");
    run<3>("F2F+FFMA+F2F");
    run<4>("DSETP/FSEL max + DADD");
    return 0;
}
