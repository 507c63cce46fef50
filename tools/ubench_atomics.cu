// Micro-benchmark: L2 reduction (RED) throughput on B200 for f64 / u64 / u32 adds, scattered over a 1.5 MB region
// (statistics records of 16 slots x 3329 labels x 128 B would be 6.8 MB; both stay in L2).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <typename T>
__global__ void k(T* data, int nWords, int perThread, int lanesActive) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    if ((int)(threadIdx.x & 31) >= lanesActive) return;
    unsigned h = tid * 2654435761u + 12345u;
    for (int i = 0; i < perThread; ++i) {
        h = h * 1664525u + 1013904223u;
        atomicAdd(data + (h >> 8) % nWords, (T)1);
    }
}

template <typename T>
void run(const char* name, int lanesActive) {
    const int nWords = 1536 * 1024 / 8;
    T* d;
    cudaMalloc(&d, nWords * sizeof(T) * 2);
    cudaMemset(d, 0, nWords * sizeof(T) * 2);
    const int blocks = 148 * 8, threads = 256, per = 30;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<T><<<blocks, threads>>>(d, nWords, per, lanesActive);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<T><<<blocks, threads>>>(d, nWords, per, lanesActive);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads / 32 * lanesActive * per;
    printf("%-10s lanes/warp %2d: %8.3f ms  %7.2f G atomics/s\n", name, lanesActive, ms, ops / (ms * 1e-3) / 1e9);
    cudaFree(d);
}

int main() {
    for (int lanes : {32, 8}) {
        run<double>("f64", lanes);
        run<unsigned long long>("u64", lanes);
        run<unsigned int>("u32", lanes);
        run<float>("f32", lanes);
    }
    return 0;
}
