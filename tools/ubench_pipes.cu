// Micro-benchmark: per-SM throughput of the integer instructions the SGM recurrence is made of (B200).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench_pipes.cu && /tmp/ubench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(uint32_t* out, uint32_t seed, int iters) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1 + i * 977u) + i;
    uint32_t c = seed | 0x01010101u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) a[i] = __popc(a[i]) + c;                       // POPC (+ IADD)
                if (OP == 1) a[i] = __viaddmin_u16x2(a[i], c, a[(i + 1) & 7]);  // VIADDMNMX
                if (OP == 2) a[i] = __vminu2(a[i], a[(i + 1) & 7]) + c;      // VIMNMX.U16x2 (+ IADD)
                if (OP == 3) a[i] = __byte_perm(a[i], a[(i + 1) & 7], c);   // PRMT
                if (OP == 4) a[i] = a[i] ^ (a[(i + 1) & 7] & c);            // LOP3
                if (OP == 5) a[i] = a[i] * 65536u + a[(i + 1) & 7];         // IMAD
                if (OP == 6) a[i] = a[i] + c;                               // IADD
                if (OP == 7) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);  // SHFL
                if (OP == 8) a[i] = __vimin3_u16x2(a[i], a[(i + 1) & 7], c);  // VIMNMX3
                if (OP == 9) a[i] = __popc(a[i] ^ c) + (a[(i+1)&7] & 0xff);   // LOP3 + POPC + LOP3 + IADD mix
                if (OP == 10) a[i] = __reduce_min_sync(0xffffffffu, a[i]) + c;  // REDUX full warp
                if (OP == 11) a[i] = __reduce_min_sync(0xffu << (threadIdx.x & 24), a[i]) + c;  // REDUX, 8-lane groups
                if (OP == 12) {  // 8-lane group min by 3 shuffles
                    uint32_t v = a[i];
                    v = min(v, __shfl_xor_sync(0xffffffffu, v, 1));
                    v = min(v, __shfl_xor_sync(0xffffffffu, v, 2));
                    v = min(v, __shfl_xor_sync(0xffffffffu, v, 4));
                    a[i] = v + c;
                }
                if (OP == 13) a[i] = __dp4a(a[i], 0x01010101u, a[(i + 1) & 7]);  // IDP4A
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void check_redux(uint32_t* out) {
    const uint32_t v = (threadIdx.x * 2654435761u) >> 7;
    out[threadIdx.x] = __reduce_min_sync(0xffu << (threadIdx.x & 24), v);
    uint32_t w = v;
    w = min(w, __shfl_xor_sync(0xffffffffu, w, 1));
    w = min(w, __shfl_xor_sync(0xffffffffu, w, 2));
    w = min(w, __shfl_xor_sync(0xffffffffu, w, 4));
    out[32 + threadIdx.x] = w;
}

template <int OP>
void run(const char* name, int opsPerInner) {
    uint32_t* out;
    const int blocks = 148 * 8, threads = 256, iters = 2000;
    cudaMalloc(&out, blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 12345u, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 12345u, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double inner = (double)blocks * threads * iters * 64.0;  // executions of the marked statement
    const double perSmPerClk = inner / (ms * 1e-3) / 148.0 / (clk * 1e3);
    printf("%-28s %8.3f ms  %7.1f stmts/clk/SM (at %d MHz nominal)  [%d instr per stmt]\n", name, ms, perSmPerClk, clk / 1000, opsPerInner);
    cudaFree(out);
}

int main() {
    run<0>("POPC+IADD", 2);
    run<1>("VIADDMNMX.U16x2", 1);
    run<2>("VIMNMX.U16x2+IADD", 2);
    run<3>("PRMT", 1);
    run<4>("LOP3", 1);
    run<5>("IMAD", 1);
    run<6>("IADD", 1);
    run<7>("SHFL", 1);
    run<8>("VIMNMX3.U16x2", 1);
    run<9>("LOP3+POPC+LOP3+IADD", 4);
    run<10>("REDUX.MIN full warp (+IADD)", 2);
    run<11>("REDUX.MIN 8-lane masks (+IADD)", 2);
    run<12>("3xSHFL+3xMIN group min (+IADD)", 7);
    run<13>("IDP4A", 1);
    // correctness of the group-masked REDUX against the shuffle tree
    {
        uint32_t* d;
        cudaMalloc(&d, 64 * 4);
        check_redux<<<1, 32>>>(d);
        uint32_t h[64];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 32; ++i) bad += h[i] != h[32 + i];
        printf("REDUX 8-lane-mask vs shuffle tree: %s\n", bad ? "MISMATCH" : "identical");
    }
    return 0;
}
