// The hot-path modules of CART-SLAM, same names / constructor parameters / data keys as the reference,
// each one a thin host shell around the cartb200 C ABI (include/cartb200.h):
//   ImageDisparityModule                          /root/reference/include/modules/disparity.hpp:24-44
//   ImageDisparityDerivativeModule                /root/reference/include/modules/disparity.hpp:70-79
//   DepthModule                                   /root/reference/include/modules/depth.hpp:12-20
//   SuperPixelModule                              /root/reference/include/modules/superpixels.hpp:15-41
//   PlaneParameters / providers                   /root/reference/include/modules/planeseg.hpp:25-113
//   DisparityPlaneSegmentationModule              /root/reference/include/modules/planeseg.hpp:115-162
//   SuperPixelDisparityPlaneSegmentationModule    /root/reference/include/modules/planeseg.hpp:164-186
// Temporal smoothing is supported when some module provides the "optflow" key (CV_16SC2, S10.5):
// the reference's producer is the NVOFA hardware engine behind OpenCV (out of scope); ExternalOpticalFlowModule below is
// the hook an integrator replaces with a real flow source.
#pragma once
#include <cmath>
#include <functional>

#include "cart/core.hpp"

struct cartb200_ctx;

#define CARTSLAM_KEY_DISPARITY "disparity"
#define CARTSLAM_KEY_DISPARITY_DERIVATIVE "disparity_derivative"
#define CARTSLAM_KEY_DISPARITY_DERIVATIVE_HISTOGRAM "disparity_derivative_histogram"
#define CARTSLAM_KEY_DEPTH "depth"
#define CARTSLAM_KEY_SUPERPIXELS "superpixels"
#define CARTSLAM_KEY_SUPERPIXELS_MAX_LABEL "superpixels_max_label"
#define CARTSLAM_KEY_PLANES "planes"
#define CARTSLAM_KEY_PLANES_UNSMOOTHED "planes_unsmoothed"
#define CARTSLAM_KEY_PLANE_PARAMETERS "plane_parameters"
#define CARTSLAM_KEY_OPTFLOW "optflow" /* /root/reference/include/modules/optflow.hpp:13 */
#define CARTSLAM_KEY_DISPARITY_DERIVATIVE_HIST "disp_derivative_histogram"
#define CARTSLAM_DISPARITY_INVALID (-32768)
#define CARTSLAM_PLANE_COUNT 3
#define CARTSLAM_PLANE_TEMPORAL_DISTANCE_DEFAULT 3

namespace cart {

typedef int16_t disparity_t;
typedef int16_t derivative_t;
typedef int16_t optical_flow_t;  // S10.5 fixed point, two channels (/root/reference/include/modules/optflow.hpp:17)
namespace contour {
typedef uint16_t label_t;
}

// RAII handle on a cartb200 context; throws std::runtime_error with cartb200_last_error on failure.
class Kernels {
   public:
    // superpixels: 0 = none, 1 = full relaxation scratch, 2 = only the vote table of the superpixel plane segmentation
    Kernels(Size size, bool sgm, int superpixels, int minDisparity = 4, int numDisparities = 256, int smoothingRadius = -1,
            int smoothingIterations = 5, int spBlockSize = 12, double direct = 0.5, double diagonal = 0.5 / std::sqrt(2.0),
            double wCompact = 0.1, double progressive = 0.0, double wImage = 1.5, double wDisparity = 1.0);
    ~Kernels();
    Kernels(const Kernels&) = delete;
    cartb200_ctx* get() const { return ctx; }
    void check(int rc, const char* what) const;
    std::mutex mutex;  // a context is not thread-safe; modules serialise their frames on it

   private:
    cartb200_ctx* ctx = nullptr;
};

class ImageDisparityModule : public SyncWrapperSystemModule {
   public:
    ImageDisparityModule(const Size imageRes, int minDisparity = 4, int numDisparities = 256, int blockSize = 3,
                         int smoothingRadius = -1, int smoothingIterations = 5);
    system_data_t runInternal(System& system, SystemRunData& data) override;

   private:
    std::unique_ptr<Kernels> kernels;
};

class ImageDisparityDerivativeModule : public SyncWrapperSystemModule {
   public:
    ImageDisparityDerivativeModule();
    system_data_t runInternal(System& system, SystemRunData& data) override;

   private:
    std::unique_ptr<Kernels> kernels;  // created lazily: the image size is only known from the data
};

// DepthModule (/root/reference/include/modules/depth.hpp:12-20, src/modules/depth.cpp:9-25): "disparity" ->
// "depth" (CV_32FC3 X, Y, Z) with the data source's reprojection matrix Q.
class DepthModule : public SyncWrapperSystemModule {
   public:
    DepthModule();
    system_data_t runInternal(System& system, SystemRunData& data) override;

   private:
    std::unique_ptr<Kernels> kernels;  // created lazily: the image size is only known from the data
};

class SuperPixelModule : public SyncWrapperSystemModule {
   public:
    SuperPixelModule(const Size imageRes, const unsigned int initialIterations = 18, const unsigned int iterations = 6,
                     const unsigned int blockSize = 12, const unsigned int resetIterations = 64,
                     const double directCliqueCost = 0.5, const double diagonalCliqueCost = 0.5 / std::sqrt(2.0),
                     const double compactnessWeight = 0.05, const double progressiveCompactnessCost = 0.0,
                     const double imageWeight = 1.0, const double disparityWeight = 1.25);
    system_data_t runInternal(System& system, SystemRunData& data) override;

   private:
    std::unique_ptr<Kernels> kernels;
    const unsigned int initialIterations, iterations, resetIterations, blockSize;
    const bool requiresDisparityDerivative;
    contour::label_t maxLabelId = 0;
};

struct PlaneParameters {
    PlaneParameters(int horizontalCenter, int verticalCenter, std::pair<int, int> horizontalRange, std::pair<int, int> verticalRange)
        : horizontalRange(horizontalRange), verticalRange(verticalRange), horizontalCenter(horizontalCenter), verticalCenter(verticalCenter) {}
    const std::pair<int, int> horizontalRange, verticalRange;
    const int horizontalCenter, verticalCenter;
};

enum Plane { HORIZONTAL = 0, VERTICAL = 1, UNKNOWN = 2 };

class PlaneParameterProvider {
   public:
    virtual ~PlaneParameterProvider() = default;
    PlaneParameters getPlaneParameters() const { return PlaneParameters(horizontalCenter, verticalCenter, horizontalRange, verticalRange); }
    // histogram: 256 bins (bin = derivative + 128)
    virtual void updatePlaneParameters(System& system, SystemRunData& data, const std::vector<int32_t>& histogram) = 0;

   protected:
    PlaneParameterProvider(int horizontalCenter = 0, int verticalCenter = 0, std::pair<int, int> horizontalRange = {0, 0},
                           std::pair<int, int> verticalRange = {0, 0})
        : horizontalRange(horizontalRange), verticalRange(verticalRange), horizontalCenter(horizontalCenter), verticalCenter(verticalCenter) {}
    std::pair<int, int> horizontalRange, verticalRange;
    int horizontalCenter, verticalCenter;
};

class HistogramPeakPlaneParameterProvider : public PlaneParameterProvider {
   public:
    void updatePlaneParameters(System& system, SystemRunData& data, const std::vector<int32_t>& histogram) override;
};

class StaticPlaneParameterProvider : public PlaneParameterProvider {
   public:
    StaticPlaneParameterProvider(int horizontalCenter, int verticalCenter, std::pair<int, int> horizontalRange, std::pair<int, int> verticalRange)
        : PlaneParameterProvider(horizontalCenter, verticalCenter, horizontalRange, verticalRange) {}
    void updatePlaneParameters(System&, SystemRunData&, const std::vector<int32_t>&) override {}
};

// Stand-in for ImageOpticalFlowModule (/root/reference/src/modules/optflow.cpp:52-134): provides "optflow" for every
// frame after the first (a null entry for frame 1, like the reference, :119-121) from a user callback that fills a
// host CV_16SC2 image (rows x cols x 2 int16, S10.5).  The default callback writes a constant flow.
class ExternalOpticalFlowModule : public SyncWrapperSystemModule {
   public:
    typedef std::function<void(uint32_t id, int rows, int cols, optical_flow_t* flowXY)> flow_fn_t;
    explicit ExternalOpticalFlowModule(flow_fn_t fn);
    ExternalOpticalFlowModule(double flowX, double flowY);  // constant flow in pixels
    system_data_t runInternal(System& system, SystemRunData& data) override;

   private:
    flow_fn_t fn;
};

// previousPlanes / previousOpticalFlow lists of the temporal vote (planeseg.cu:300-343, sp_planeseg.cu:256-316);
// keeps the images alive until the kernel that reads them has finished.
struct TemporalHistory {
    std::vector<std::shared_ptr<image_t>> planes, flows;
    int count() const { return (int)planes.size(); }
};
TemporalHistory collectTemporalHistory(SystemRunData& data, unsigned int distance);

class DisparityPlaneSegmentationModule : public SyncWrapperSystemModule {
   public:
    DisparityPlaneSegmentationModule(std::shared_ptr<PlaneParameterProvider> planeParameterProvider, const int updateInterval = 30,
                                     const int resetInterval = 10, const bool useTemporalSmoothing = false,
                                     const unsigned int temporalSmoothingDistance = CARTSLAM_PLANE_TEMPORAL_DISTANCE_DEFAULT);
    system_data_t runInternal(System& system, SystemRunData& data) override;

   private:
    void updatePlaneParameters(System& system, SystemRunData& data);
    const int updateInterval, resetInterval;
    const bool useTemporalSmoothing;
    const unsigned int temporalSmoothingDistance;
    std::shared_ptr<PlaneParameterProvider> planeParameterProvider;
    std::unique_ptr<Kernels> kernels;
    std::mutex derivativeHistogramMutex;
    std::vector<int64_t> derivativeHistogram;  // running total (the reference keeps it on the GPU)
};

class SuperPixelDisparityPlaneSegmentationModule : public SyncWrapperSystemModule {
   public:
    SuperPixelDisparityPlaneSegmentationModule(std::shared_ptr<PlaneParameterProvider> planeParameterProvider,
                                               const int updateInterval = 30, const int resetInterval = 10,
                                               const bool useTemporalSmoothing = false,
                                               const unsigned int temporalSmoothingDistance = CARTSLAM_PLANE_TEMPORAL_DISTANCE_DEFAULT);
    system_data_t runInternal(System& system, SystemRunData& data) override;

   private:
    void updatePlaneParameters(System& system, SystemRunData& data);
    const int updateInterval, resetInterval;
    const bool useTemporalSmoothing;
    const unsigned int temporalSmoothingDistance;
    std::shared_ptr<PlaneParameterProvider> planeParameterProvider;
    std::unique_ptr<Kernels> kernels;
    std::mutex derivativeHistogramMutex;
    bool histogramCreated = false;
    std::vector<int64_t> derivativeHistogram;
};

// ---- configuration (same JSON keys and defaults as /root/reference/src/cartconfig.cpp:56-228) --------
namespace config {
// Module types outside the hot-path scope (visualisation, optflow, depth, features, planefit, ...) are
// rejected with std::runtime_error unless skipOutOfScope is set, in which case they are skipped with a warning
// (and `use_temporal_smoothing` is forced off when no module provides "optflow").  Extra type of this build:
// {"type": "external_optflow", "flow_x": px, "flow_y": px} -> ExternalOpticalFlowModule with a constant flow.
void applyModuleConfigText(const std::string& jsonText, std::shared_ptr<System> system, bool skipOutOfScope = false);
void readModuleConfig(const std::string& path, std::shared_ptr<System> system, bool skipOutOfScope = false);
}  // namespace config

}  // namespace cart
