/* cartb200 - C ABI of the B200-native dense stereo front end (disparity -> plane segmentation).
 *
 * This is the drop-in boundary for CART-SLAM's module layer: every entry point replaces the device work
 * of one reference module (or one third-party call inside it) and is what a reference-side FFI for the
 * path would bind.  Plain pointers and sizes only; no C++ / torch types.
 *
 * Conventions
 *  - All image pointers are DEVICE pointers unless the name ends in `_host`.  Pitches are in BYTES.
 *  - Every call is asynchronous on the caller-supplied `stream` (a cudaStream_t passed as void*; NULL =
 *    the legacy default stream), except the `_host` convenience calls, which synchronise before returning.
 *  - Batch calls take `n` frames laid out with a constant byte stride between frames.
 *  - Return value: 0 on success, < 0 = CARTB200_E_*.  Nothing throws, nothing calls exit().
 *    cartb200_last_error(ctx) returns a description of the last failure on that context.
 *  - A context is not thread-safe; use one context per in-flight frame slot (the reference runs up to
 *    12 frames concurrently, /root/reference/include/cartslam.hpp:3-5).
 */
#ifndef CARTB200_H
#define CARTB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CARTB200_OK 0
#define CARTB200_E_ARG (-1)
#define CARTB200_E_SHAPE (-2)
#define CARTB200_E_CUDA (-3)
#define CARTB200_E_NOMEM (-4)
#define CARTB200_E_UNSUPPORTED (-5)

#define CARTB200_DISPARITY_INVALID (-32768) /* /root/reference/include/modules/disparity.hpp:17 */
#define CARTB200_PLANE_HORIZONTAL 0         /* /root/reference/include/modules/planeseg.hpp:37-41 */
#define CARTB200_PLANE_VERTICAL 1
#define CARTB200_PLANE_UNKNOWN 2

typedef struct cartb200_ctx cartb200_ctx;

/* Mirrors the constructor arguments of the reference modules on the path.
 * disparity:   ImageDisparityModule(imageRes, minDisparity, numDisparities, blockSize, smoothingRadius,
 *              smoothingIterations)  /root/reference/include/modules/disparity.hpp:26-34, JSON defaults
 *              /root/reference/src/cartconfig.cpp:144-152.  p1/p2/uniqueness_ratio/paths are the values
 *              the reference gets implicitly from cv::cuda::createStereoSGM (10, 120, 12, 4 = MODE_HH4).
 * superpixels: SuperPixelModule(...) /root/reference/include/modules/superpixels.hpp:17-28, JSON defaults
 *              /root/reference/src/cartconfig.cpp:121-134. */
typedef struct cartb200_config {
    int width, height;
    int max_batch; /* frames per batched call (scratch is sized for this) */
    /* disparity (enable_sgm = 0 skips allocating the SGM scratch: census, path volumes, WTA images) */
    int enable_sgm;
    int min_disparity;    /* 4 */
    int num_disparities;  /* 64, 128 or 256 */
    int p1, p2;           /* 10, 120 */
    int uniqueness_ratio; /* 12 */
    int paths;            /* 4 (MODE_HH4) or 8 (MODE_HH) */
    int smoothing_radius; /* <= 0: no interpolation */
    int smoothing_iterations;
    /* superpixels: 0 = no scratch; 1 = everything; 2 = only the per-label vote table cartb200_sp_planeseg* need, sized
     * for the label count sp_block_size implies (a module that consumes labels produced elsewhere) */
    int enable_superpixels;
    int sp_block_size; /* 12 */
    double sp_direct_clique_cost, sp_diagonal_clique_cost;
    double sp_compactness_weight, sp_progressive_compactness_cost, sp_image_weight, sp_disparity_weight;
    /* Kept for ABI compatibility, ignored.  The label costs are always evaluated in the reference's operation order
     * (gaussian.cu:30-43, compactness.cu:28-35, contourrelaxation.cu:102-144) with a fully specified logarithm: the
     * labels are bit-identical to the scalar oracle, also over arbitrarily long warm-started chains.  (A faster
     * approximate mode existed while the exact evaluation was slower than it; it is gone.) */
    int sp_exact;
} cartb200_config;

/* Fills `cfg` with the reference's JSON defaults (cartconfig.cpp:121-152) for a WxH stream. */
void cartb200_default_config(cartb200_config* cfg, int width, int height);

int cartb200_create(const cartb200_config* cfg, cartb200_ctx** out);
/* Description of the last failed cartb200_create on the calling thread (a failed create leaves no context to ask). */
const char* cartb200_last_create_error(void);
void cartb200_destroy(cartb200_ctx* ctx);
const char* cartb200_last_error(const cartb200_ctx* ctx);
const char* cartb200_version(void);
/* Number of kernels this context has launched so far (bench.py's `gpu_launches`). */
long long cartb200_launch_count(const cartb200_ctx* ctx);
/* Bytes of device scratch the context owns. */
size_t cartb200_scratch_bytes(const cartb200_ctx* ctx);

/* ---- disparity stage ------------------------------------------------------------------------------
 * Replaces ImageDisparityModule::runInternal's device work
 * (/root/reference/src/modules/disparity/disparity.cu:49-80): cvtColor BGR2GRAY x2 (:66-67),
 * cv::cuda::StereoSGM::compute (:71, third party) and disparity::interpolate (:73-75,
 * /root/reference/src/modules/disparity/interpolation.cu:85-99).
 * left/right: n frames of CV_8UC3 BGR.  disparity: n frames of CV_16SC1 (x16 fixed point). */
int cartb200_disparity(cartb200_ctx* ctx, int n, const uint8_t* left_bgr, const uint8_t* right_bgr, size_t bgr_pitch,
                       size_t bgr_frame_stride, int16_t* disparity, size_t disp_pitch, size_t disp_frame_stride,
                       void* stream);

/* The three sub-stages of cartb200_disparity, exposed for parity tests and profiling.
 * gray_census: BGR -> gray (kept inside the context for the L/R-check mask) + 9x7 census, both images.
 * aggregate:   P path volumes (u8 [n][y][x][d]) inside the context.
 * wta_post:    WTA + uniqueness + sub-pixel, right disparity, 3x3 medians, L/R check, range correction,
 *              then (smoothing_radius > 0) the interpolation pass. */
int cartb200_sgm_gray_census(cartb200_ctx* ctx, int n, const uint8_t* left_bgr, const uint8_t* right_bgr,
                             size_t bgr_pitch, size_t bgr_frame_stride, void* stream);
int cartb200_sgm_aggregate(cartb200_ctx* ctx, int n, void* stream);
/* One path only (0 L->R, 1 R->L, 2 T->B, 3 B->T, 4..7 diagonals) - for per-kernel timing and parity tests. */
int cartb200_sgm_aggregate_path(cartb200_ctx* ctx, int n, int path, void* stream);
int cartb200_sgm_wta_post(cartb200_ctx* ctx, int n, int16_t* disparity, size_t disp_pitch, size_t disp_frame_stride,
                          void* stream);
/* In-place smoothing of an existing CV_16SC1 image (interpolation.cu:85-99). min_disparity is the
 * already-scaled lower bound (module passes minDisparity*16), max_disparity the upper bound (module passes
 * the image width, unscaled - reproduced as is). */
int cartb200_interpolate(cartb200_ctx* ctx, int n, int16_t* disparity, size_t disp_pitch, size_t disp_frame_stride,
                         int radius, int iterations, int min_disparity, int max_disparity, void* stream);

/* Debug/parity access to the context's intermediates of the last SGM call (device pointers, tightly
 * packed with the returned pitches).  which: 0 census L (u32), 1 census R (u32), 2 gray L (u8),
 * 3 left raw WTA (u16), 4 right raw WTA (u16), 10+p aggregated volume of path p (u8 [n][H][W][D]). */
int cartb200_sgm_intermediate(cartb200_ctx* ctx, int which, const void** ptr, size_t* pitch, size_t* frame_stride);

/* ---- derivative stage ----------------------------------------------------------------------------
 * Replaces ImageDisparityDerivativeModule::runInternal
 * (/root/reference/src/modules/disparity/derivative.cu:151-184: calculateDirectionalDerivatives :27-97 +
 * mergeDerivativeHistograms :99-116).  derivative: CV_16SC2 (ch0 vertical, ch1 horizontal);
 * histogram: n x 256 x 2 int32 (bin = value + 128; ch0 vertical) - "disparity_derivative_histogram". */
int cartb200_derivative(cartb200_ctx* ctx, int n, const int16_t* disparity, size_t disp_pitch,
                        size_t disp_frame_stride, int16_t* derivative, size_t deriv_pitch, size_t deriv_frame_stride,
                        int32_t* histogram, void* stream);

/* ---- naive plane segmentation --------------------------------------------------------------------
 * Replaces DisparityPlaneSegmentationModule::runInternal's kernels
 * (/root/reference/src/modules/planeseg/planeseg.cu: calculateDerivatives :31-142, mergeHistogram :144-158,
 * classifyPlanes :160-243 without the temporal vote).
 * naive_derivative: derivative CV_16SC1 + this frame's 256-bin histogram (n x 256 int32); the running
 * total across frames is host state (the module layer adds frames in id order).
 * classify: range rule on channel `channel` of a derivative image with `channels` int16 per pixel. */
int cartb200_naive_derivative(cartb200_ctx* ctx, int n, const int16_t* disparity, size_t disp_pitch,
                              size_t disp_frame_stride, int16_t* derivative, size_t deriv_pitch,
                              size_t deriv_frame_stride, int32_t* histogram, void* stream);
/* params: n x 4 int32 on the HOST: {horizontalStart, horizontalEnd, verticalStart, verticalEnd} per frame
 * (PlaneParameters is passed by value to the reference kernels, planeseg.cu:349-350). */
int cartb200_classify(cartb200_ctx* ctx, int n, const int16_t* derivative, size_t deriv_pitch,
                      size_t deriv_frame_stride, int channels, int channel, const int32_t* params_host, uint8_t* planes,
                      size_t planes_pitch, size_t planes_frame_stride, void* stream);

/* ---- superpixels ---------------------------------------------------------------------------------
 * Replaces SuperPixelModule::runInternal (/root/reference/src/modules/superpixels.cu:71-121) and
 * ContourRelaxation::relax (/root/reference/src/modules/superpixels/contourrelaxation/contourrelaxation.cu:349-447).
 * The context owns `max_batch` persistent label images ("slots", one per independent sequence chunk).
 * reset: createBlockInitialization (initialization.cu:39-59) on the given slots; *max_label_out (host)
 *        receives the label COUNT ("superpixels_max_label").
 * relax: for slot s (0..n-1): BGR->YCrCb of left image s, statistics init, `iterations` relaxation
 *        iterations against derivative image s (CV_16SC2; may be NULL when the disparity weight is <= 0),
 *        then copies the slot's labels to labels_out (CV_16UC1). slot_ids (host, n ints) selects slots;
 *        NULL = 0..n-1. */
int cartb200_superpixels_reset(cartb200_ctx* ctx, int n, const int* slot_ids_host, int* max_label_out, void* stream);
int cartb200_superpixels_relax(cartb200_ctx* ctx, int n, const int* slot_ids_host, int iterations,
                               const uint8_t* left_bgr, size_t bgr_pitch, size_t bgr_frame_stride,
                               const int16_t* derivative, size_t deriv_pitch, size_t deriv_frame_stride,
                               uint16_t* labels_out, size_t labels_pitch, size_t labels_frame_stride, void* stream);
/* Overwrite / read a slot's persistent label image (ContourRelaxation::setLabelImage, contourrelaxation.cu:335-338). */
int cartb200_superpixels_set_labels(cartb200_ctx* ctx, int slot, const uint16_t* labels, size_t labels_pitch,
                                    void* stream);
/* Border membership map of findBorderPixels (contourrelaxation.cu:146-219) for parity tests: border u8. */
int cartb200_superpixels_border_map(cartb200_ctx* ctx, const uint16_t* labels, size_t labels_pitch, uint8_t* border,
                                    size_t border_pitch, void* stream);

/* ---- superpixel plane segmentation ---------------------------------------------------------------
 * Replaces SuperPixelDisparityPlaneSegmentationModule::runInternal's kernels
 * (/root/reference/src/modules/planeseg/sp_planeseg.cu: performSuperPixelClassifications :25-134 and
 * classifyPlanes :136-184, no temporal vote).  planes_unsmoothed = per-pixel class from the vertical
 * derivative; planes = per-superpixel majority.  max_label = label COUNT.
 * Fails with CARTB200_E_UNSUPPORTED when (max_label+1)*6 > 32768 like the reference (:327-331). */
int cartb200_sp_planeseg(cartb200_ctx* ctx, int n, const int16_t* derivative, size_t deriv_pitch,
                         size_t deriv_frame_stride, const uint16_t* labels, size_t labels_pitch,
                         size_t labels_frame_stride, int max_label, const int32_t* params_host,
                         uint8_t* planes_unsmoothed, uint8_t* planes, size_t planes_pitch, size_t planes_frame_stride,
                         void* stream);

/* ---- temporal smoothing vote -----------------------------------------------
 * Replaces the previousPlanesCount > 0 branch of classifyPlanes
 * (/root/reference/src/modules/planeseg/planeseg.cu:160-243, vote :199-240) and of performSuperPixelClassifications
 * (/root/reference/src/modules/planeseg/sp_planeseg.cu:25-134, vote :79-117).  The reference hands its kernels two
 * device arrays of cv_mat_ptr_t {data, step} (/root/reference/include/utils/cuda.cuh:47-51) filled by the module
 * (planeseg.cu:300-343, sp_planeseg.cu:256-316); here the same list is a HOST array, entry k =
 *   planes_unsmoothed of frame id-(k+1)  ("planes_unsmoothed", CV_8UC1, values 0..2)  and
 *   optflow           of frame id-k      ("optflow", CV_16SC2, S10.5 fixed point; pitch a multiple of 4).
 * The optical flow itself is produced outside this library (the reference uses the NVOFA engine through OpenCV,
 * /root/reference/src/modules/optflow.cpp:58-70).  previous_count == 0 (frame id 1): smoothed = unsmoothed, as the
 * module returns one image under both keys (planeseg.cu:361-368).  One frame per call: the module surface is
 * per-frame; all images of a call share the context's W x H. */
#define CARTB200_MAX_TEMPORAL_DISTANCE 8
typedef struct cartb200_temporal_ref {
    const uint8_t* planes_unsmoothed;
    size_t planes_pitch;
    const int16_t* optflow;
    size_t optflow_pitch;
} cartb200_temporal_ref;
int cartb200_classify_temporal(cartb200_ctx* ctx, const int16_t* derivative, size_t deriv_pitch, int channels, int channel,
                               const int32_t* params_host /* 4 */, int previous_count,
                               const cartb200_temporal_ref* previous_host, uint8_t* planes_unsmoothed,
                               uint8_t* planes_smoothed, size_t planes_pitch, void* stream);
int cartb200_sp_planeseg_temporal(cartb200_ctx* ctx, const int16_t* derivative, size_t deriv_pitch, const uint16_t* labels,
                                  size_t labels_pitch, int max_label, const int32_t* params_host /* 4 */,
                                  int previous_count, const cartb200_temporal_ref* previous_host,
                                  uint8_t* planes_unsmoothed, uint8_t* planes, size_t planes_pitch, void* stream);

/* ---- superpixel consumers of the plane fit ------------------------------------
 * label_statistics replaces countPixels (/root/reference/src/modules/planefit.cu:38-83, called from
 * generateLabelStatistics :182-209): pixel_count[l], pixel_count_invalid[l] for l < n_labels (= maxLabelId + 1),
 * DEVICE uint32 arrays (the reference's label_statistics_t holds two uint16, /root/reference/include/modules/planefit.hpp:20-23,
 * updated by a 16-bit atomic that is not carry-safe; counts here are exact).
 * region_inliers replaces calculateRegionDistance (planefit.cu:85-138, called from attemptAssignment :211-275):
 * inliers[p * n_labels + l] = valid-depth pixels of label l closer than `threshold` to plane p.
 * planes_host: n_planes x {a, b, c, d} doubles on the HOST (plane_t, planefit.cu:27-32).
 * depth: CV_32FC3 "depth" image (X, Y, Z), pitch a multiple of 4. */
int cartb200_label_statistics(cartb200_ctx* ctx, const uint16_t* labels, size_t labels_pitch, const float* depth,
                              size_t depth_pitch, int n_labels, uint32_t* pixel_count, uint32_t* pixel_count_invalid,
                              void* stream);
int cartb200_region_inliers(cartb200_ctx* ctx, const uint16_t* labels, size_t labels_pitch, const float* depth,
                            size_t depth_pitch, int n_labels, const double* planes_host, int n_planes, double threshold,
                            uint32_t* inliers, void* stream);

/* Overlay kernels of the same row (visual QA).  overlay_planes replaces overlayPlanes
 * (/root/reference/src/modules/planeseg/planeseg_vis.cu:28-56): out = image / 2 + PlaneColor[plane] / 2 (BGR, CV_8UC3).
 * overlay_superpixel_boundaries replaces overlayBoundaryVisualization
 * (/root/reference/src/modules/superpixels/visualization.cu:9-42, called from computeBoundaryOverlay :46-66): pixels whose
 * right or lower neighbour has another label turn red; the last row and column of `out` are not written, as in the
 * reference. */
int cartb200_overlay_planes(cartb200_ctx* ctx, const uint8_t* image_bgr, size_t image_pitch, const uint8_t* planes,
                            size_t planes_pitch, uint8_t* out_bgr, size_t out_pitch, void* stream);
int cartb200_overlay_superpixel_boundaries(cartb200_ctx* ctx, const uint8_t* image_bgr, size_t image_pitch,
                                           const uint16_t* labels, size_t labels_pitch, uint8_t* out_bgr, size_t out_pitch,
                                           void* stream);

/* ---- depth (the stage right after disparity) ------------------------------
 * Replaces DepthModule::runInternal (/root/reference/src/modules/depth.cpp:9-25): convertTo(CV_32F, 1/16) +
 * cv::cuda::reprojectImageTo3D(disparityFloat, depth, Q, 3).  q16_host: the 4x4 reprojection matrix Q
 * (CameraIntrinsics::Q, /root/reference/include/datasource.hpp:11-18; built at src/sources/kitti.cpp:141-148),
 * row-major floats on the HOST.  depth: n frames of CV_32FC3 (X, Y, Z), "depth" key.  Invalid disparities are
 * not treated specially (neither does the reference's GPU path). */
int cartb200_depth(cartb200_ctx* ctx, int n, const int16_t* disparity, size_t disp_pitch, size_t disp_frame_stride,
                   const float* q16_host, float* depth, size_t depth_pitch, size_t depth_frame_stride, void* stream);

/* ---- host-side parameter estimation --------------------------------------------------------------
 * HistogramPeakPlaneParameterProvider::updatePlaneParameters (/root/reference/src/modules/planeseg/planeseg.cu:405-458)
 * + util::findPeaks (/root/reference/src/utils/peaks.cpp:12-72).  hist256: HOST 256 x int32.
 * params: HOST {horizontalCenter, verticalCenter, hStart, hEnd, vStart, vEnd}, updated in place.
 * Returns 1 if the ranges were updated, 0 if the reference's early returns fired, < 0 on error. */
int cartb200_histogram_peak_update(const int32_t* hist256_host, int32_t* params_host);

/* ---- whole-path convenience with HOST buffers (the end-to-end call timed by bench.py `e2e`) -------
 * Runs disparity -> derivative -> [superpixels] -> planeseg for n_frames consecutive frames of ONE
 * sequence given in pinned or pageable host memory, reproducing the reference's per-sequence state
 * (superpixel reset schedule, running histograms, parameter update cadence; frame ids start_id..).
 * pipeline: 0 = naive (config/modules/kitti-naive-segmentation.json: disparity -> disparity_planeseg),
 *           1 = superpixel (config/modules/kitti-planeseg.json minus optflow/depth/vis/temporal smoothing).
 * provider: 0 = static (static_params = {hS,hE,vS,vE}), 1 = histogram_peak.
 * Outputs (host, tightly packed): planes n x H x W u8; optional (may be NULL) disparity n x H x W s16.
 * Host<->device copies are issued inside this call, overlapped with compute. */
typedef struct cartb200_sequence_opts {
    int pipeline;
    int provider;
    int static_params[4];
    int update_interval;   /* 30 */
    int reset_interval;    /* 10 */
    int sp_initial_iterations, sp_iterations, sp_reset_iterations; /* 18, 6, 64 */
    int start_id;          /* 1 */
} cartb200_sequence_opts;
void cartb200_default_sequence_opts(cartb200_sequence_opts* o);
int cartb200_run_sequence_host(cartb200_ctx* ctx, const cartb200_sequence_opts* opts, int n_frames,
                               const uint8_t* left_bgr_host, const uint8_t* right_bgr_host, uint8_t* planes_host,
                               int16_t* disparity_host);
/* Same pipeline with inputs already resident on the device (bench.py `value`): device buffers tightly
 * packed; planes_dev n x H x W u8. */
int cartb200_run_sequence_device(cartb200_ctx* ctx, const cartb200_sequence_opts* opts, int n_frames,
                                 const uint8_t* left_bgr_dev, const uint8_t* right_bgr_dev, uint8_t* planes_dev,
                                 int16_t* disparity_dev, void* stream);

/* ---- source-side resize ------------------------------------------------------------------------------------------
 * Replaces cv::cuda::resize(image, image, imageSize, 0, 0, cv::INTER_LINEAR, stream) on the CV_8UC3 frames of
 * KITTIDataSource::getNextInternal (/root/reference/src/sources/kitti.cpp:166-169).  Device buffers, pitches in bytes,
 * no context needed.  Bit-identical to oracle/stages.cpp orc_resize_bgr8; parity against OpenCV's kernel is unpinned
 * (third-party cudawarping source, restated from the published algorithm). */
int cartb200_resize_bgr8(const uint8_t* src_bgr, size_t src_pitch, int src_width, int src_height, uint8_t* dst_bgr,
                         size_t dst_pitch, int dst_width, int dst_height, void* stream);

/* ---- one sequence sharded over several GPUs (BASELINE.json configs[4]) ------------------
 * The superpixel chain is cut at every reset frame, so a shard that starts at id 1 or at a multiple of
 * sp_reset_iterations (opts->start_id) reproduces the unsharded labels.  The plane parameters of the histogram_peak
 * providers, however, come from a RUNNING histogram over all earlier frames (planeseg.cu:379-403,
 * sp_planeseg.cu:352-388), which crosses shard boundaries.  The two-pass scheme:
 *   phase 1 (per shard)    everything up to the per-frame histograms; hist_host = n x 256 int32 (HOST; naive pipeline:
 *                          the frame's derivative histogram, superpixel pipeline: its vertical channel).  Derivative and
 *                          label images of the n frames stay inside the context.
 *   exchange               the ranks all-gather their histograms (a few hundred KB) - the only data-path collective
 *                          besides the final result gather;
 *   cartb200_sequence_parameters (CPU, any rank)   the reference's bookkeeping over the WHOLE sequence in id order:
 *                          hist = n_total x 256, opts->start_id = id of the first row; params = n_total x 4
 *                          {hStart, hEnd, vStart, vEnd} per frame (static provider: the static ranges);
 *   phase 2 (per shard)    params_host = the shard's n x 4 slice -> classify / vote + assign -> planes_dev n x H x W.
 * phase 2 must follow a phase 1 of the same context, pipeline and n.  Running phase 1, cartb200_sequence_parameters and
 * phase 2 on one whole sequence equals cartb200_run_sequence_device bit for bit. */
int cartb200_run_sequence_phase1_device(cartb200_ctx* ctx, const cartb200_sequence_opts* opts, int n_frames,
                                        const uint8_t* left_bgr_dev, const uint8_t* right_bgr_dev, int32_t* hist_host,
                                        int16_t* disparity_dev, void* stream);
int cartb200_run_sequence_phase2_device(cartb200_ctx* ctx, const cartb200_sequence_opts* opts, int n_frames,
                                        const int32_t* params_host, uint8_t* planes_dev, void* stream);
/* the same two phases with HOST image / plane buffers (uploads and downloads overlapped with compute inside the calls) */
int cartb200_run_sequence_phase1_host(cartb200_ctx* ctx, const cartb200_sequence_opts* opts, int n_frames,
                                      const uint8_t* left_bgr_host, const uint8_t* right_bgr_host, int32_t* hist_host,
                                      int16_t* disparity_host);
int cartb200_run_sequence_phase2_host(cartb200_ctx* ctx, const cartb200_sequence_opts* opts, int n_frames,
                                      const int32_t* params_host, uint8_t* planes_host);
int cartb200_sequence_parameters(const cartb200_sequence_opts* opts, int n_frames, const int32_t* hist_host,
                                 int32_t* params_host);

/* CPU evaluation of the closed-form reference-tile mapping used by the kernels (no GPU needed); lets the
 * CPU test-suite check it against the oracle's literal copyToShared simulation.  Returns the value the
 * reference's shared tile holds at local (lx, ly) of block (bx, by) for an int32 image. */
int32_t cartb200_debug_ref_tile_i32(const int32_t* img_host, int W, int H, int bx, int by, int bdx, int bdy, int XB,
                                    int YB, int y_pad, int x_pad, int interp, long alloc_elems, int32_t undef, int lx,
                                    int ly);

#ifdef __cplusplus
}
#endif
#endif /* CARTB200_H */
