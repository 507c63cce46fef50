// TEST INFRASTRUCTURE ONLY.  Minimal stand-ins for the OpenCV / project types the reference's kernels
// mention, so that the kernels can be compiled VERBATIM from /root/reference.
// Nothing here is reference code.
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/memcpy_async.h>
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <utility>
#include <vector>

namespace cv {
struct Point3f {
    float x, y, z;
};
namespace cuda {
template <typename T>
struct PtrStepSz {
    T* data;
    size_t step;
    int cols, rows;
    __host__ __device__ operator T*() const { return data; }
    __host__ __device__ T* ptr(int y = 0) const { return (T*)((char*)data + (size_t)y * step); }
    __host__ __device__ T& operator()(int y, int x) const { return ptr(y)[x]; }
};
}  // namespace cuda
}  // namespace cv

namespace cart {
typedef int16_t disparity_t;
typedef int16_t derivative_t;
typedef int16_t optical_flow_t;
__device__ inline void assignColor(float, float, uint8_t*) {}
namespace contour {
typedef uint16_t label_t;
}
}  // namespace cart
#define CARTSLAM_DISPARITY_INVALID (-32768)
namespace cg = cooperative_groups;

// ---- harness helpers (ours) -------------------------------------------------------------------
namespace refh {
#define REF_CUDA(x)                                                                         \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            fprintf(stderr, "ref harness: %s -> %s\n", #x, cudaGetErrorString(e_));        \
            return -1;                                                                      \
        }                                                                                   \
    } while (0)

// pitched device image that mimics a cv::cuda::GpuMat allocation (cudaMallocPitch)
constexpr int kSlackRows = 16;
template <typename T>
struct DevMat {
    T* data = nullptr;
    size_t step = 0;
    int cols = 0, rows = 0, ch = 1;
    int create(int r, int c, int channels = 1) {
        rows = r;
        cols = c;
        ch = channels;
        // kSlackRows zero-filled rows follow the image: the reference reads up to yPad rows past the last
        // row; a real GpuMat allocation usually has mapped memory there, a tight one faults.
        cudaError_t e = cudaMallocPitch((void**)&data, &step, (size_t)c * channels * sizeof(T), r + kSlackRows);
        if (e == cudaSuccess) e = cudaMemset(data, 0, step * (size_t)(r + kSlackRows));
        if (e != cudaSuccess) fprintf(stderr, "ref harness: cudaMallocPitch(%d x %d x %d) -> %s\n", r, c, channels, cudaGetErrorString(e));
        return e == cudaSuccess ? 0 : -1;
    }
    int upload(const T* host) {
        cudaError_t e = cudaMemcpy2D(data, step, host, (size_t)cols * ch * sizeof(T), (size_t)cols * ch * sizeof(T), rows,
                                     cudaMemcpyHostToDevice);
        if (e != cudaSuccess) fprintf(stderr, "ref harness: upload -> %s\n", cudaGetErrorString(e));
        return e == cudaSuccess ? 0 : -1;
    }
    int download(T* host) const {
        cudaError_t e = cudaMemcpy2D(host, (size_t)cols * ch * sizeof(T), data, step, (size_t)cols * ch * sizeof(T), rows,
                                     cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) fprintf(stderr, "ref harness: download -> %s\n", cudaGetErrorString(e));
        return e == cudaSuccess ? 0 : -1;
    }
    int zero() { return cudaMemset2D(data, step, 0, (size_t)cols * ch * sizeof(T), rows) == cudaSuccess ? 0 : -1; }
    cv::cuda::PtrStepSz<T> view() const { return cv::cuda::PtrStepSz<T>{data, step, cols, rows}; }
    ~DevMat() {
        if (data) cudaFree(data);
    }
};

struct Timer {
    cudaEvent_t a, b;
    Timer() {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
    }
    ~Timer() {
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
    void start() { cudaEventRecord(a); }
    float stop() {
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        return ms;
    }
};
}  // namespace refh
