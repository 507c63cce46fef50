/* TEST INFRASTRUCTURE ONLY: CPU restatement ("oracle") of the reference's disparity -> planeseg path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker or the CPU baseline - never the product path.
 * All images are tightly packed row-major.  See sgm.cpp / stages.cpp / superpixels.cpp for the
 * reference file:line each function follows.  SGM stage: PARITY UNPINNED (third-party, see sgm.cpp). */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int orc_gray(const uint8_t* bgr, int W, int H, uint8_t* gray);
int orc_ycrcb(const uint8_t* bgr, int W, int H, uint8_t* out);
int orc_census(const uint8_t* gray, int W, int H, uint32_t* out);
int orc_sgm_dirs(int paths, int* out);
int orc_sgm_path(const uint32_t* cl, const uint32_t* cr, int W, int H, int D, int minDisp, int P1, int P2, int dx,
                 int dy, uint8_t* L);
int orc_sgm_wta(const uint8_t* const* Ls, int P, int W, int H, int D, int uniquenessRatio, uint16_t* left,
                uint16_t* right);
int orc_median3(const uint16_t* in, int W, int H, uint16_t* out);
int orc_lr_check_range(const uint16_t* left, const uint16_t* right, const uint8_t* grayLeft, int W, int H, int minDisp,
                       int16_t* out);
int orc_sgm_compute(const uint8_t* leftBgr, const uint8_t* rightBgr, int W, int H, int D, int minDisp, int P1, int P2,
                    int uniquenessRatio, int paths, int16_t* disp, uint32_t* censusL, uint32_t* censusR,
                    uint8_t* volumes, uint16_t* leftRaw, uint16_t* rightRaw);
int orc_interpolate(int16_t* disp, int W, int H, int radius, int iterations, int minDisparity, int maxDisparity,
                    uint8_t* defined);
int orc_derivative(const int16_t* disp, int W, int H, int16_t* deriv, int32_t* hist, uint8_t* defined);
int orc_naive_derivative(const int16_t* disp, int W, int H, int16_t* deriv, int32_t* hist, uint8_t* defined);
int orc_classify(const int16_t* deriv, long n, int stride, int hS, int hE, int vS, int vE, uint8_t* planes);
int orc_sp_planeseg(const int16_t* deriv2, const uint16_t* labels, int W, int H, int maxLabel, int hS, int hE, int vS,
                    int vE, uint8_t* planesUnsmoothed, uint8_t* planes);
/* temporal smoothing vote: planeseg.cu:199-240 / sp_planeseg.cu:79-117 */
int orc_classify_temporal(const int16_t* deriv, int W, int H, int stride, int hS, int hE, int vS, int vE, int count,
                          const uint8_t* const* prevPlanes, const int16_t* const* prevFlow, uint8_t* planesUnsmoothed,
                          uint8_t* planesSmoothed);
int orc_sp_planeseg_temporal(const int16_t* deriv2, const uint16_t* labels, int W, int H, int maxLabel, int hS, int hE,
                             int vS, int vE, int count, const uint8_t* const* prevPlanes, const int16_t* const* prevFlow,
                             uint8_t* planesUnsmoothed, uint8_t* planes);
/* superpixel consumers of the plane fit: countPixels / calculateRegionDistance, planefit.cu:38-138 */
int orc_label_statistics(const uint16_t* labels, const float* xyz, int W, int H, int nLabels, uint32_t* pixelCount,
                         uint32_t* pixelCountInvalid);
int orc_region_inliers(const uint16_t* labels, const float* xyz, int W, int H, int nLabels, const double* planes,
                       int nPlanes, double threshold, uint32_t* inliers);
/* overlay kernels: overlayPlanes (planeseg_vis.cu:28-56), overlayBoundaryVisualization (superpixels/visualization.cu:9-42) */
int orc_overlay_planes(const uint8_t* bgr, const uint8_t* planes, int W, int H, uint8_t* out);
int orc_overlay_boundaries(const uint8_t* bgr, const uint16_t* labels, int W, int H, uint8_t* out);
/* depth: DepthModule (src/modules/depth.cpp:9-25): convertTo(CV_32F, 1/16) + cv::cuda::reprojectImageTo3D(Q), 3 channels */
int orc_depth(const int16_t* disp, int W, int H, const float* Q16, float* xyz);
int orc_find_peaks(const int32_t* hist, int n, int* out, int maxPeaks);
int orc_histogram_peak_update(const int32_t* hist, int* params);
int orc_block_init(int W, int H, int bw, int bh, uint16_t* labels);
int orc_border_map(const uint16_t* labels, int W, int H, uint8_t* border, uint8_t* defined);
int orc_sp_relax(uint16_t* labels, int W, int H, int maxLabel, const uint8_t* ycrcb, const int16_t* deriv2,
                 int iterations, double directCost, double diagCost, double wCompact, double progressive, double wDisp,
                 double wImage, int32_t* borderCounts, int32_t* moved);
int orc_resize_bgr8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
/* copyToShared simulation exposed for the tile-loader unit test (sanity_check.cu:58-65 idea):
 * fills out[(tileH+2*yPad) * (tileW+2*xPad)] int32 values and def flags for one block. */
int orc_tile_i32(const int32_t* img, int W, int H, int bx, int by, int bdx, int bdy, int XB, int YB, int yPad, int xPad,
                 int interp, long allocElems, int32_t undef, int32_t* out, uint8_t* def);
#ifdef __cplusplus
}
#endif
