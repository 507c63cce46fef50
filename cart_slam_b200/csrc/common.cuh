// Shared declarations for the cartb200 CUDA sources (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/cartb200.h"

namespace cb {

constexpr int16_t kInvalid = -32768;  // CARTSLAM_DISPARITY_INVALID
constexpr int kNumSMs = 148;

template <typename T>
struct Img {  // pitched device image view
    T* data;
    size_t pitch;  // bytes
    __host__ __device__ __forceinline__ T* row(int y) const { return (T*)((char*)data + (size_t)y * pitch); }
    __host__ __device__ __forceinline__ T& at(int x, int y) const { return row(y)[x]; }
};

template <typename T>
struct ImgBatch {  // n frames with a constant byte stride; optionally a two-level layout (inner x outer)
    T* data;
    size_t pitch, frameStride;  // bytes
    int inner = 0;              // > 0: frame f lives at (f % inner) * frameStride + (f / inner) * outerStride
    size_t outerStride = 0;     //      (chunks x consecutive steps of a sequence processed in one launch)
    __host__ __device__ __forceinline__ Img<T> frame(int f) const {
        const size_t off = inner > 0 ? (size_t)(f % inner) * frameStride + (size_t)(f / inner) * outerStride
                                     : (size_t)f * frameStride;
        return Img<T>{(T*)((char*)data + off), pitch};
    }
    __host__ __device__ __forceinline__ ImgBatch<T> from(int f0) const {  // sub-batch starting at simple-layout frame f0
        return ImgBatch<T>{(T*)((char*)data + (size_t)f0 * frameStride), pitch, frameStride, inner, outerStride};
    }
};

inline int ceilDiv(int a, int b) { return (a + b - 1) / b; }
inline size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct PlaneRanges {
    int hS, hE, vS, vE;
};

// previous frames of the temporal smoothing vote, passed to the kernels by value
constexpr int kMaxTemporal = CARTB200_MAX_TEMPORAL_DISTANCE;
struct TemporalRefs {
    const uint8_t* planes[kMaxTemporal];  // planes_unsmoothed of frame id-(k+1)
    const int16_t* flow[kMaxTemporal];    // optflow (CV_16SC2, S10.5) of frame id-k
    size_t planesPitch[kMaxTemporal], flowPitch[kMaxTemporal];  // bytes
    int count;
};

// candidate planes (a, b, c, d) of the region-distance kernel, passed by value in sets of kMaxPlaneSet
constexpr int kMaxPlaneSet = 16;
struct PlaneSet {
    double abcd[kMaxPlaneSet][4];
    int count;
};

}  // namespace cb

// The context. Owns every scratch buffer, sized at create time for max_batch frames.
struct cartb200_ctx {
    cartb200_config cfg;
    std::string err;
    long long launches = 0;
    size_t scratchBytes = 0;
    int W = 0, H = 0, D = 0, B = 0, P = 0;
    int device = 0;  // the CUDA device the context was created on; every entry point checks it is current
    // SGM scratch
    uint8_t* grayL = nullptr;   // [B][H][grayPitch]
    uint8_t* grayR = nullptr;
    size_t grayPitch = 0;
    uint32_t* censusL = nullptr;  // [B][H][censusPitch/4]
    uint32_t* censusR = nullptr;
    size_t censusPitch = 0;      // bytes per census row = 4 * cenRowWords
    size_t cenRowWords = 0;      // [cenMargin zeros][W words][zeros]
    int cenMargin = 0;
    uint8_t* volumes = nullptr;  // [P][B][H][W][D]
    size_t volFrameStride = 0, volPathStride = 0;
    uint16_t* wtaL = nullptr;  // [B][H][dispPitch/2]
    uint32_t* wtaR = nullptr;  // [B][H][rkPitch] right-image keys (aggregated cost << 16 | disparity)
    size_t rkPitch = 0;        // elements
    uint16_t* medL = nullptr;
    uint16_t* medR = nullptr;
    size_t dispPitch = 0;
    // planeseg scratch
    int32_t* paramsDev = nullptr;  // [B][4]
    uint32_t* votes = nullptr;     // [B][maxLabels][4]
    // superpixels
    int maxLabels = 0, spBlocksPerRow = 0;
    uint16_t* spLabels = nullptr;  // [B][2][H][spLabelPitch/2] persistent labels in plane 0, plane 1 = ping-pong partner
    size_t spLabelPitch = 0;
    uint8_t* spYcc = nullptr;  // [B][H][W][4] Y,Cr,Cb,border-flag scratch
    double* spStats = nullptr;     // [B]{[labels][16] records, [labels][8] stored costs, [labels][16] deltas}
    int* spTileMap = nullptr;      // [tilesY][tilesX] index into spTileTab, -1 = interior tile
    uint32_t* spTileTab = nullptr; // [edge tiles][66*66] source pixel (y << 16 | x) of the reference's label tile
    // the path kernels of one aggregation run on up to three streams (horizontal pair / vertical pair / diagonals) so
    // that the last, partially filled wave of one launch is filled by the next (lazy)
    cudaStream_t aggStream[2] = {nullptr, nullptr};
    cudaEvent_t aggFork = nullptr, aggJoin[2] = {nullptr, nullptr};
    // sliced SGM (disparity_batch): the winner-takes-all pass of slice i runs on this stream beside the aggregation of
    // slice i + 1 (lazy)
    cudaStream_t wtaStream = nullptr;
    cudaEvent_t evSliceAgg = nullptr, evSliceWta = nullptr;
    // sequence runner scratch (lazy)
    void* seq = nullptr;
};

#define CB_CHECK_CUDA(ctx, expr)                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                        \
            return CARTB200_E_CUDA;                                                                  \
        }                                                                                            \
    } while (0)

#define CB_LAUNCH_CHECK(ctx)                                                                         \
    do {                                                                                             \
        (ctx)->launches++;                                                                           \
        cudaError_t e__ = cudaPeekAtLastError();                                                     \
        if (e__ != cudaSuccess) {                                                                    \
            (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + std::to_string(__LINE__); \
            return CARTB200_E_CUDA;                                                                  \
        }                                                                                            \
    } while (0)

// stage launchers (defined in the per-stage .cu files); all return CARTB200_* codes
namespace cb {
int launch_gray_census(cartb200_ctx* c, int n, ImgBatch<const uint8_t> left, ImgBatch<const uint8_t> right, cudaStream_t s);
int disparity_batch(cartb200_ctx* c, int n, ImgBatch<const uint8_t> left, ImgBatch<const uint8_t> right, ImgBatch<int16_t> disp, cudaStream_t s);
int launch_aggregate(cartb200_ctx* c, int n, cudaStream_t s);
int launch_aggregate_range(cartb200_ctx* c, int n, int p0, int p1, cudaStream_t s);
int launch_wta(cartb200_ctx* c, int n, cudaStream_t s);
int launch_sgm_post(cartb200_ctx* c, int n, ImgBatch<int16_t> disp, cudaStream_t s);
int launch_interpolate(cartb200_ctx* c, int n, ImgBatch<int16_t> disp, int radius, int iterations, int minD, int maxD, cudaStream_t s);
int launch_interpolate_from(cartb200_ctx* c, int n, ImgBatch<const int16_t> src, ImgBatch<int16_t> dst, int radius, int iterations, int minD, int maxD, cudaStream_t s);
int launch_derivative(cartb200_ctx* c, int n, ImgBatch<const int16_t> disp, ImgBatch<int16_t> deriv, int32_t* hist, cudaStream_t s);
int launch_naive_derivative(cartb200_ctx* c, int n, ImgBatch<const int16_t> disp, ImgBatch<int16_t> deriv, int32_t* hist, cudaStream_t s);
int launch_classify(cartb200_ctx* c, int n, ImgBatch<const int16_t> deriv, int channels, int channel, const int32_t* paramsDev, ImgBatch<uint8_t> planes, cudaStream_t s);
int launch_sp_planeseg(cartb200_ctx* c, int n, ImgBatch<const int16_t> deriv, ImgBatch<const uint16_t> labels, int maxLabel, const int32_t* paramsDev, ImgBatch<uint8_t> unsm, ImgBatch<uint8_t> planes, cudaStream_t s);
int launch_sp_reset(cartb200_ctx* c, int n, const int* slotsDev, cudaStream_t s);
int launch_sp_relax(cartb200_ctx* c, int n, const int* slotsDev, int iterations, ImgBatch<const uint8_t> left, ImgBatch<const int16_t> deriv, bool hasDeriv, ImgBatch<uint16_t> out, cudaStream_t s, int scratchBase = 0);
int launch_classify_temporal(cartb200_ctx* c, Img<const int16_t> deriv, int channels, int channel, PlaneRanges pr, const TemporalRefs& refs, Img<uint8_t> unsm, Img<uint8_t> smoothed, cudaStream_t s);
int launch_sp_planeseg_temporal(cartb200_ctx* c, Img<const int16_t> deriv, Img<const uint16_t> labels, int maxLabel, PlaneRanges pr, const TemporalRefs& refs, Img<uint8_t> unsm, Img<uint8_t> planes, cudaStream_t s);
int launch_label_statistics(cartb200_ctx* c, Img<const uint16_t> labels, Img<const float> xyz, int nLabels, uint32_t* count, uint32_t* invalid, cudaStream_t s);
int launch_region_inliers(cartb200_ctx* c, Img<const uint16_t> labels, Img<const float> xyz, int nLabels, const double* planesHost, int nPlanes, double threshold, uint32_t* inliers, cudaStream_t s);
int launch_overlay_planes(cartb200_ctx* c, Img<const uint8_t> bgr, Img<const uint8_t> planes, Img<uint8_t> out, cudaStream_t s);
int launch_overlay_boundaries(cartb200_ctx* c, Img<const uint8_t> bgr, Img<const uint16_t> labels, Img<uint8_t> out, cudaStream_t s);
int launch_depth(cartb200_ctx* c, int n, ImgBatch<const int16_t> disp, ImgBatch<float> xyz, const float* q16Host, cudaStream_t s);
int launch_border_map(cartb200_ctx* c, Img<const uint16_t> labels, Img<uint8_t> border, cudaStream_t s);
int launch_resize_bgr8(const uint8_t* src, size_t srcPitch, int sw, int sh, uint8_t* dst, size_t dstPitch, int dw, int dh, cudaStream_t s);
// per-device kernel attributes (dynamic shared memory limits), applied by cartb200_create on the context's device
cudaError_t sgm_set_kernel_attributes();
cudaError_t post_set_kernel_attributes();
cudaError_t sp_set_kernel_attributes();
void build_sp_tile_tables(int W, int H, std::vector<int>& tileMap, std::vector<uint32_t>& tab);
}  // namespace cb
