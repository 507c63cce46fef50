// Hot-path modules: host shells around the cartb200 C ABI.  Each runInternal mirrors the data flow of
// the reference module it replaces (cited per function); all device work happens in libcartb200.
#include "cart/modules.hpp"

#include <cuda_runtime.h>

#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../../include/cartb200.h"

namespace cart {

namespace {
// The reference creates and synchronises one stream per module call (e.g. derivative.cu:171-179); here the streams are
// recycled through a small pool instead of being created and destroyed for every frame and module.
// (the event is kept for callers that prefer a blocking wait; StreamGuard::sync polls)
struct PooledStream {
    cudaStream_t s = nullptr;
    cudaEvent_t done = nullptr;
};
struct StreamPool {
    std::mutex m;
    std::vector<PooledStream> free;
    PooledStream take() {
        {
            std::lock_guard<std::mutex> lock(m);
            if (!free.empty()) {
                PooledStream s = free.back();
                free.pop_back();
                return s;
            }
        }
        PooledStream s;
        if (cudaStreamCreateWithFlags(&s.s, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.done, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess)
            throw std::runtime_error("cudaStreamCreate failed");
        return s;
    }
    void give(PooledStream s) {
        std::lock_guard<std::mutex> lock(m);
        free.push_back(s);
    }
};
StreamPool& streamPool() {
    static StreamPool* pool = new StreamPool();
    return *pool;
}
struct StreamGuard {
    PooledStream ps;
    cudaStream_t s = nullptr;
    StreamGuard() : ps(streamPool().take()), s(ps.s) {}
    ~StreamGuard() { streamPool().give(ps); }  // every module synchronises the stream before it returns
    // Short waits (one frame in flight: a module's kernels take 50-300 us) are polled - a blocking-sync event costs ~100 us
    // of wake-up latency per module call.  With a dozen frames in flight the waits are long and dozens of threads polling
    // cudaStreamQuery fight over the driver's context lock with the threads that launch kernels, so after a bounded
    // number of polls the thread blocks on the stream's event instead.
    void sync() {
        cudaError_t e = cudaErrorNotReady;
        for (int i = 0; i < 48 && (e = cudaStreamQuery(s)) == cudaErrorNotReady; ++i) std::this_thread::yield();
        if (e == cudaErrorNotReady) {
            e = cudaEventRecord(ps.done, s);
            if (e == cudaSuccess) e = cudaEventSynchronize(ps.done);
        }
        if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e));
    }
};
}  // namespace

Kernels::Kernels(Size size, bool sgm, int superpixels, int minDisparity, int numDisparities, int smoothingRadius,
                 int smoothingIterations, int spBlockSize, double direct, double diagonal, double wCompact, double progressive,
                 double wImage, double wDisparity) {
    cartb200_config cfg;
    cartb200_default_config(&cfg, size.width, size.height);
    cfg.max_batch = 1;
    cfg.enable_sgm = sgm;
    cfg.enable_superpixels = superpixels;
    cfg.min_disparity = minDisparity;
    cfg.num_disparities = numDisparities;
    cfg.smoothing_radius = smoothingRadius;
    cfg.smoothing_iterations = smoothingIterations;
    cfg.sp_block_size = spBlockSize;
    cfg.sp_direct_clique_cost = direct;
    cfg.sp_diagonal_clique_cost = diagonal;
    cfg.sp_compactness_weight = wCompact;
    cfg.sp_progressive_compactness_cost = progressive;
    cfg.sp_image_weight = wImage;
    cfg.sp_disparity_weight = wDisparity;
    const int rc = cartb200_create(&cfg, &ctx);
    if (rc != CARTB200_OK) throw std::runtime_error("cartb200_create failed with code " + std::to_string(rc));
}

Kernels::~Kernels() { cartb200_destroy(ctx); }

void Kernels::check(int rc, const char* what) const {
    if (rc != CARTB200_OK) throw std::runtime_error(std::string(what) + ": " + cartb200_last_error(ctx));
}

// ---- ImageDisparityModule (disparity.hpp:26-34, disparity.cu:49-80) ----------------------------------
ImageDisparityModule::ImageDisparityModule(const Size imageRes, int minDisparity, int numDisparities, int blockSize,
                                           int smoothingRadius, int smoothingIterations)
    : SyncWrapperSystemModule("ImageDisparity") {
    (void)blockSize;  // setBlockSize is a no-op for cv::cuda::StereoSGM
    providesData.push_back(CARTSLAM_KEY_DISPARITY);
    kernels.reset(new Kernels(imageRes, true, 0, minDisparity, numDisparities, smoothingRadius, smoothingIterations));
}

system_data_t ImageDisparityModule::runInternal(System&, SystemRunData& data) {
    if (data.dataElement->type != DataElementType::STEREO) throw std::runtime_error("ImageDisparityModule requires StereoDataElement");
    auto stereo = std::static_pointer_cast<StereoDataElement>(data.dataElement);
    image_t disparity(stereo->left.rows, stereo->left.cols, IMG_16SC1);
    if (stereo->left.pitch != stereo->right.pitch) throw std::runtime_error("left/right pitch mismatch");
    StreamGuard st;
    {
        std::lock_guard<std::mutex> lock(kernels->mutex);
        kernels->check(cartb200_disparity(kernels->get(), 1, stereo->left.as<uint8_t>(), stereo->right.as<uint8_t>(), stereo->left.pitch,
                                          0, disparity.as<int16_t>(), disparity.pitch, 0, st.s),
                       "ImageDisparityModule");
        st.sync();
    }
    return MODULE_RETURN(CARTSLAM_KEY_DISPARITY, std::shared_ptr<void>(std::make_shared<image_t>(disparity)));
}

// ---- ImageDisparityDerivativeModule (derivative.cu:151-184) -------------------------------------------
ImageDisparityDerivativeModule::ImageDisparityDerivativeModule() : SyncWrapperSystemModule("ImageDisparityDerivative") {
    requiresData.push_back(module_dependency_t(CARTSLAM_KEY_DISPARITY));
    providesData.push_back(CARTSLAM_KEY_DISPARITY_DERIVATIVE);
    providesData.push_back(CARTSLAM_KEY_DISPARITY_DERIVATIVE_HISTOGRAM);
}

system_data_t ImageDisparityDerivativeModule::runInternal(System&, SystemRunData& data) {
    auto disparity = data.getData<image_t>(CARTSLAM_KEY_DISPARITY);
    if (disparity->empty() || disparity->type != IMG_16SC1) throw std::runtime_error("Disparity must be of type CV_16SC1");
    image_t derivatives(disparity->rows, disparity->cols, IMG_16SC2);
    image_t histogram(1, 256, IMG_32SC2);
    StreamGuard st;
    {
        static std::mutex createMutex;
        std::lock_guard<std::mutex> lock(createMutex);
        if (!kernels) kernels.reset(new Kernels(disparity->size(), false, 0));
    }
    {
        std::lock_guard<std::mutex> lock(kernels->mutex);
        kernels->check(cartb200_derivative(kernels->get(), 1, disparity->as<int16_t>(), disparity->pitch, 0, derivatives.as<int16_t>(),
                                           derivatives.pitch, 0, histogram.as<int32_t>(), st.s),
                       "ImageDisparityDerivativeModule");
        st.sync();
    }
    return MODULE_RETURN_ALL(MODULE_MAKE_PAIR(CARTSLAM_KEY_DISPARITY_DERIVATIVE, image_t, derivatives),
                             MODULE_MAKE_PAIR(CARTSLAM_KEY_DISPARITY_DERIVATIVE_HISTOGRAM, image_t, histogram));
}

// ---- DepthModule (depth.cpp:9-25) -----------------------------------------------------------------------
DepthModule::DepthModule() : SyncWrapperSystemModule("Depth") {
    requiresData.push_back(module_dependency_t(CARTSLAM_KEY_DISPARITY));
    providesData.push_back(CARTSLAM_KEY_DEPTH);
}

system_data_t DepthModule::runInternal(System& system, SystemRunData& data) {
    auto disparity = data.getData<image_t>(CARTSLAM_KEY_DISPARITY);
    if (disparity->empty() || disparity->type != IMG_16SC1) throw std::runtime_error("Disparity must be of type CV_16SC1");
    const CameraIntrinsics in = system.getDataSource()->getCameraIntrinsics();
    image_t depth(disparity->rows, disparity->cols, IMG_32FC3);
    StreamGuard st;
    {
        static std::mutex createMutex;
        std::lock_guard<std::mutex> lock(createMutex);
        if (!kernels) kernels.reset(new Kernels(disparity->size(), false, 0));
    }
    {
        std::lock_guard<std::mutex> lock(kernels->mutex);
        kernels->check(cartb200_depth(kernels->get(), 1, disparity->as<int16_t>(), disparity->pitch, 0, in.Q, depth.as<float>(),
                                      depth.pitch, 0, st.s),
                       "DepthModule");
        st.sync();
    }
    return MODULE_RETURN_SHARED(CARTSLAM_KEY_DEPTH, image_t, depth);
}

// ---- SuperPixelModule (superpixels.cu:19-121) -----------------------------------------------------------
SuperPixelModule::SuperPixelModule(const Size imageRes, const unsigned int initialIterations, const unsigned int iterations,
                                   const unsigned int blockSize, const unsigned int resetIterations, const double directCliqueCost,
                                   const double diagonalCliqueCost, const double compactnessWeight,
                                   const double progressiveCompactnessCost, const double imageWeight, const double disparityWeight)
    : SyncWrapperSystemModule("SuperPixelDetect"),
      initialIterations(initialIterations),
      iterations(iterations),
      resetIterations(resetIterations),
      blockSize(blockSize),
      requiresDisparityDerivative(disparityWeight > 0) {
    if (blockSize < 1) throw std::invalid_argument("blockSize must be more than 1");
    if (directCliqueCost < 0) throw std::invalid_argument("directCliqueCost must be non-negative");
    if (compactnessWeight < 0 || imageWeight < 0 || disparityWeight < 0) throw std::invalid_argument("weight must be non-negative");
    if (resetIterations < 1) throw std::invalid_argument("resetIterations must be positive");
    if (disparityWeight > 0) requiresData.push_back(module_dependency_t(CARTSLAM_KEY_DISPARITY_DERIVATIVE));
    providesData.push_back(CARTSLAM_KEY_SUPERPIXELS);
    providesData.push_back(CARTSLAM_KEY_SUPERPIXELS_MAX_LABEL);
    // context creation performs createBlockInitialization on its label slot (superpixels.cu:56-58)
    kernels.reset(new Kernels(imageRes, false, 1, 4, 256, -1, 5, (int)blockSize, directCliqueCost, diagonalCliqueCost, compactnessWeight,
                              progressiveCompactnessCost, imageWeight, disparityWeight));
    int ml = 0;
    kernels->check(cartb200_superpixels_reset(kernels->get(), 1, nullptr, &ml, nullptr), "SuperPixelModule");
    cudaDeviceSynchronize();
    maxLabelId = (contour::label_t)ml;
}

system_data_t SuperPixelModule::runInternal(System&, SystemRunData& data) {
    image_t image = getReferenceImage(data.dataElement);
    std::shared_ptr<image_t> derivative;
    if (requiresDisparityDerivative) derivative = data.getData<image_t>(CARTSLAM_KEY_DISPARITY_DERIVATIVE);
    const unsigned int numIterations = (data.id == 1 || data.id % resetIterations == 0) ? initialIterations : iterations;
    image_t relaxed(image.rows, image.cols, IMG_16UC1);
    StreamGuard st;
    {
        // the label image is persistent across frames: frames are serialised here (superpixels.cu:99)
        std::lock_guard<std::mutex> lock(kernels->mutex);
        if (data.id % resetIterations == 0) {
            int ml = 0;
            kernels->check(cartb200_superpixels_reset(kernels->get(), 1, nullptr, &ml, st.s), "SuperPixelModule reset");
            maxLabelId = (contour::label_t)ml;
        }
        kernels->check(cartb200_superpixels_relax(kernels->get(), 1, nullptr, (int)numIterations, image.as<uint8_t>(), image.pitch, 0,
                                                  derivative ? derivative->as<int16_t>() : nullptr, derivative ? derivative->pitch : 0, 0,
                                                  relaxed.as<uint16_t>(), relaxed.pitch, 0, st.s),
                       "SuperPixelModule");
        st.sync();
    }
    return MODULE_RETURN_ALL(MODULE_MAKE_PAIR(CARTSLAM_KEY_SUPERPIXELS, image_t, relaxed),
                             MODULE_MAKE_PAIR(CARTSLAM_KEY_SUPERPIXELS_MAX_LABEL, contour::label_t, maxLabelId));
}

// ---- plane parameter providers (planeseg.cu:405-458) ----------------------------------------------------
void HistogramPeakPlaneParameterProvider::updatePlaneParameters(System&, SystemRunData&, const std::vector<int32_t>& histogram) {
    int32_t p[6] = {horizontalCenter, verticalCenter, horizontalRange.first, horizontalRange.second, verticalRange.first, verticalRange.second};
    const int rc = cartb200_histogram_peak_update(histogram.data(), p);
    if (rc < 0) throw std::runtime_error("histogram_peak_update failed");
    if (rc == 0) CART_LOG_WARN("PlaneParameters", "Histogram peak provider: ranges not updated");
    horizontalCenter = p[0];
    verticalCenter = p[1];
    horizontalRange = {p[2], p[3]};
    verticalRange = {p[4], p[5]};
}

// ---- temporal smoothing support --------------------------------------------------------
static void checkTemporalDistance(bool use, unsigned int distance) {
    if (use && (distance < 1 || distance > CARTB200_MAX_TEMPORAL_DISTANCE))
        throw std::runtime_error("temporal_smoothing_distance must be 1.." + std::to_string(CARTB200_MAX_TEMPORAL_DISTANCE));
}

// the dependency list of planeseg.hpp:128-137 / sp_planeseg.cu:200-209
static void addTemporalDependencies(std::vector<module_dependency_t>& requiresData, unsigned int distance) {
    requiresData.push_back(module_dependency_t(CARTSLAM_KEY_OPTFLOW));
    for (unsigned int i = 1; i <= distance; i++) {
        requiresData.push_back(module_dependency_t(CARTSLAM_KEY_PLANES_UNSMOOTHED, -(int)i));
        if ((i + 1) <= distance) requiresData.push_back(module_dependency_t(CARTSLAM_KEY_OPTFLOW, -(int)i));
    }
}

// planeseg.cu:300-337 / sp_planeseg.cu:256-310: entry k = planes_unsmoothed of run id-(k+1) and optflow of run id-k
TemporalHistory collectTemporalHistory(SystemRunData& data, unsigned int distance) {
    TemporalHistory h;
    if (data.id <= 1) return h;
    std::shared_ptr<image_t> flow = data.getData<image_t>(CARTSLAM_KEY_OPTFLOW);
    for (int i = 1; i <= (int)distance; i++) {
        if ((int)data.id - i <= 0) break;
        auto relativeRun = data.getRelativeRun((int8_t)-i);
        auto prev = relativeRun->getData<image_t>(CARTSLAM_KEY_PLANES_UNSMOOTHED);  // blocks until available
        if (!flow || flow->empty() || flow->type != IMG_16SC2) throw std::runtime_error("optflow must be a CV_16SC2 image");
        h.planes.push_back(prev);
        h.flows.push_back(flow);
        flow.reset();
        if (relativeRun->id > 1 && h.count() < (int)distance) flow = relativeRun->getData<image_t>(CARTSLAM_KEY_OPTFLOW);
    }
    return h;
}

static std::vector<cartb200_temporal_ref> temporalRefs(const TemporalHistory& h) {
    std::vector<cartb200_temporal_ref> refs(h.count());
    for (int k = 0; k < h.count(); ++k)
        refs[k] = {h.planes[k]->as<uint8_t>(), h.planes[k]->pitch, h.flows[k]->as<int16_t>(), h.flows[k]->pitch};
    return refs;
}

ExternalOpticalFlowModule::ExternalOpticalFlowModule(flow_fn_t fn) : SyncWrapperSystemModule("ExternalOpticalFlow"), fn(std::move(fn)) {
    providesData.push_back(CARTSLAM_KEY_OPTFLOW);
}

ExternalOpticalFlowModule::ExternalOpticalFlowModule(double flowX, double flowY)
    : ExternalOpticalFlowModule([flowX, flowY](uint32_t, int rows, int cols, optical_flow_t* out) {
          const optical_flow_t fx = (optical_flow_t)std::lrint(flowX * 32.0), fy = (optical_flow_t)std::lrint(flowY * 32.0);
          for (size_t i = 0; i < (size_t)rows * cols; ++i) {
              out[2 * i] = fx;
              out[2 * i + 1] = fy;
          }
      }) {}

system_data_t ExternalOpticalFlowModule::runInternal(System&, SystemRunData& data) {
    if (data.id <= 1) return MODULE_RETURN(CARTSLAM_KEY_OPTFLOW, std::shared_ptr<void>());  // optflow.cpp:119-121
    const image_t reference = getReferenceImage(data.dataElement);
    std::vector<optical_flow_t> host((size_t)reference.rows * reference.cols * 2);
    fn(data.id, reference.rows, reference.cols, host.data());
    image_t flow(reference.rows, reference.cols, IMG_16SC2);
    flow.upload(host.data(), (size_t)reference.cols * 2 * sizeof(optical_flow_t));
    syncStream(nullptr);
    return MODULE_RETURN_SHARED(CARTSLAM_KEY_OPTFLOW, image_t, flow);
}

// ---- DisparityPlaneSegmentationModule (planeseg.cu:246-403) ----------------------------------------------
DisparityPlaneSegmentationModule::DisparityPlaneSegmentationModule(std::shared_ptr<PlaneParameterProvider> provider, const int updateInterval,
                                                                   const int resetInterval, const bool useTemporalSmoothing,
                                                                   const unsigned int temporalSmoothingDistance)
    : SyncWrapperSystemModule("PlaneSegmentation"),
      updateInterval(updateInterval),
      resetInterval(resetInterval),
      useTemporalSmoothing(useTemporalSmoothing),
      temporalSmoothingDistance(temporalSmoothingDistance),
      planeParameterProvider(provider) {
    checkTemporalDistance(useTemporalSmoothing, temporalSmoothingDistance);
    requiresData.push_back(module_dependency_t(CARTSLAM_KEY_DISPARITY));
    if (useTemporalSmoothing) addTemporalDependencies(requiresData, temporalSmoothingDistance);  // planeseg.hpp:128-137
    providesData.push_back(CARTSLAM_KEY_PLANES);
    if (useTemporalSmoothing) providesData.push_back(CARTSLAM_KEY_PLANES_UNSMOOTHED);  // planeseg.hpp:141-143
}

system_data_t DisparityPlaneSegmentationModule::runInternal(System& system, SystemRunData& data) {
    auto disparity = data.getData<image_t>(CARTSLAM_KEY_DISPARITY);
    if (disparity->empty()) return MODULE_NO_RETURN_VALUE;
    if (disparity->type != IMG_16SC1) throw std::runtime_error("Disparity must be of type CV_16SC1");
    image_t derivatives(disparity->rows, disparity->cols, IMG_16SC1), planes(disparity->rows, disparity->cols, IMG_8UC1);
    image_t frameHist(1, 256, IMG_32SC1);
    std::vector<int32_t> hist(256);
    StreamGuard st;
    {
        static std::mutex createMutex;
        std::lock_guard<std::mutex> lock(createMutex);
        if (!kernels) kernels.reset(new Kernels(disparity->size(), false, 0));
    }
    std::lock_guard<std::mutex> klock(kernels->mutex);  // also orders the running total by arrival
    kernels->check(cartb200_naive_derivative(kernels->get(), 1, disparity->as<int16_t>(), disparity->pitch, 0, derivatives.as<int16_t>(),
                                             derivatives.pitch, 0, frameHist.as<int32_t>(), st.s),
                   "DisparityPlaneSegmentationModule");
    frameHist.download(hist.data(), 256 * sizeof(int32_t), st.s);
    {
        std::lock_guard<std::mutex> lock(derivativeHistogramMutex);
        if (derivativeHistogram.empty()) derivativeHistogram.assign(256, 0);
        for (int i = 0; i < 256; ++i) derivativeHistogram[i] += hist[i];  // mergeHistogram (planeseg.cu:144-158)
    }
    updatePlaneParameters(system, data);
    const PlaneParameters p = planeParameterProvider->getPlaneParameters();
    const int32_t params[4] = {p.horizontalRange.first, p.horizontalRange.second, p.verticalRange.first, p.verticalRange.second};
    if (useTemporalSmoothing) {  // planeseg.cu:300-376
        const TemporalHistory history = collectTemporalHistory(data, temporalSmoothingDistance);
        const std::vector<cartb200_temporal_ref> refs = temporalRefs(history);
        image_t smoothed(disparity->rows, disparity->cols, IMG_8UC1);
        if (smoothed.pitch != planes.pitch) throw std::runtime_error("plane image pitch mismatch");
        kernels->check(cartb200_classify_temporal(kernels->get(), derivatives.as<int16_t>(), derivatives.pitch, 1, 0, params, history.count(),
                                                  refs.data(), planes.as<uint8_t>(), smoothed.as<uint8_t>(), planes.pitch, st.s),
                       "DisparityPlaneSegmentationModule classify (temporal)");
        st.sync();
        if (data.id == 1) {  // both keys name the same image on the first run (planeseg.cu:361-368)
            auto ptr = std::shared_ptr<void>(std::make_shared<image_t>(planes));
            return MODULE_RETURN_ALL(std::make_pair(std::string(CARTSLAM_KEY_PLANES), ptr),
                                     std::make_pair(std::string(CARTSLAM_KEY_PLANES_UNSMOOTHED), ptr));
        }
        return MODULE_RETURN_ALL(MODULE_MAKE_PAIR(CARTSLAM_KEY_PLANES, image_t, smoothed),
                                 MODULE_MAKE_PAIR(CARTSLAM_KEY_PLANES_UNSMOOTHED, image_t, planes));
    }
    kernels->check(cartb200_classify(kernels->get(), 1, derivatives.as<int16_t>(), derivatives.pitch, 0, 1, 0, params, planes.as<uint8_t>(),
                                     planes.pitch, 0, st.s),
                   "DisparityPlaneSegmentationModule classify");
    st.sync();
    return MODULE_RETURN_SHARED(CARTSLAM_KEY_PLANES, image_t, planes);
}

void DisparityPlaneSegmentationModule::updatePlaneParameters(System& system, SystemRunData& data) {
    if (data.id % updateInterval != 1) return;  // planeseg.cu:381
    std::vector<int32_t> histogram(256);
    {
        std::lock_guard<std::mutex> lock(derivativeHistogramMutex);
        for (int i = 0; i < 256; ++i) histogram[i] = (int32_t)derivativeHistogram[i];
        if (data.id % (updateInterval * resetInterval) == 1)
            std::fill(derivativeHistogram.begin(), derivativeHistogram.end(), 0);  // planeseg.cu:391-394
    }
    planeParameterProvider->updatePlaneParameters(system, data, histogram);
    system.insertGlobalData(CARTSLAM_KEY_PLANE_PARAMETERS, std::make_shared<PlaneParameters>(planeParameterProvider->getPlaneParameters()));
    system.insertGlobalData(CARTSLAM_KEY_DISPARITY_DERIVATIVE_HIST, std::make_shared<std::vector<int32_t>>(histogram));
}

// ---- SuperPixelDisparityPlaneSegmentationModule (sp_planeseg.cu:188-388) ---------------------------------
SuperPixelDisparityPlaneSegmentationModule::SuperPixelDisparityPlaneSegmentationModule(std::shared_ptr<PlaneParameterProvider> provider,
                                                                                       const int updateInterval, const int resetInterval,
                                                                                       const bool useTemporalSmoothing,
                                                                                       const unsigned int temporalSmoothingDistance)
    : SyncWrapperSystemModule("SPPlaneSegmentation"),
      updateInterval(updateInterval),
      resetInterval(resetInterval),
      useTemporalSmoothing(useTemporalSmoothing),
      temporalSmoothingDistance(temporalSmoothingDistance),
      planeParameterProvider(provider) {
    checkTemporalDistance(useTemporalSmoothing, temporalSmoothingDistance);
    requiresData.push_back(module_dependency_t(CARTSLAM_KEY_SUPERPIXELS));
    requiresData.push_back(module_dependency_t(CARTSLAM_KEY_SUPERPIXELS_MAX_LABEL));
    requiresData.push_back(module_dependency_t(CARTSLAM_KEY_DISPARITY_DERIVATIVE));
    requiresData.push_back(module_dependency_t(CARTSLAM_KEY_DISPARITY_DERIVATIVE_HISTOGRAM));
    if (useTemporalSmoothing) addTemporalDependencies(requiresData, temporalSmoothingDistance);  // sp_planeseg.cu:200-209
    providesData.push_back(CARTSLAM_KEY_PLANES);
    providesData.push_back(CARTSLAM_KEY_PLANES_UNSMOOTHED);
}

system_data_t SuperPixelDisparityPlaneSegmentationModule::runInternal(System& system, SystemRunData& data) {
    auto derivatives = data.getData<image_t>(CARTSLAM_KEY_DISPARITY_DERIVATIVE);
    if (derivatives->type != IMG_16SC2) throw std::runtime_error("Disparity derivative must be of type CV_16SC2");
    updatePlaneParameters(system, data);
    image_t planes(derivatives->rows, derivatives->cols, IMG_8UC1), smoothed(derivatives->rows, derivatives->cols, IMG_8UC1);
    auto labels = data.getData<image_t>(CARTSLAM_KEY_SUPERPIXELS);
    const int maxLabel = *data.getData<contour::label_t>(CARTSLAM_KEY_SUPERPIXELS_MAX_LABEL);
    const PlaneParameters p = planeParameterProvider->getPlaneParameters();
    const int32_t params[4] = {p.horizontalRange.first, p.horizontalRange.second, p.verticalRange.first, p.verticalRange.second};
    if (planes.pitch != smoothed.pitch) throw std::runtime_error("plane image pitch mismatch");
    StreamGuard st;
    {
        static std::mutex createMutex;
        std::lock_guard<std::mutex> lock(createMutex);
        // sp_block_size 1 sizes the vote table for any label count the reference accepts (<= 5461, sp_planeseg.cu:327-331)
        if (!kernels) {
            int bs = 1;
            while ((long)((derivatives->cols + bs - 1) / bs) * ((derivatives->rows + bs - 1) / bs) > 16383) ++bs;
            kernels.reset(new Kernels(derivatives->size(), false, 2, 4, 256, -1, 5, bs));  // vote table only
        }
    }
    if (useTemporalSmoothing) {  // sp_planeseg.cu:256-345
        const TemporalHistory history = collectTemporalHistory(data, temporalSmoothingDistance);
        const std::vector<cartb200_temporal_ref> refs = temporalRefs(history);
        std::lock_guard<std::mutex> lock(kernels->mutex);
        kernels->check(cartb200_sp_planeseg_temporal(kernels->get(), derivatives->as<int16_t>(), derivatives->pitch, labels->as<uint16_t>(),
                                                     labels->pitch, maxLabel, params, history.count(), refs.data(), planes.as<uint8_t>(),
                                                     smoothed.as<uint8_t>(), planes.pitch, st.s),
                       "SuperPixelDisparityPlaneSegmentationModule (temporal)");
        st.sync();
    } else {
        std::lock_guard<std::mutex> lock(kernels->mutex);
        kernels->check(cartb200_sp_planeseg(kernels->get(), 1, derivatives->as<int16_t>(), derivatives->pitch, 0, labels->as<uint16_t>(),
                                            labels->pitch, 0, maxLabel, params, planes.as<uint8_t>(), smoothed.as<uint8_t>(), planes.pitch, 0,
                                            st.s),
                       "SuperPixelDisparityPlaneSegmentationModule");
        st.sync();
    }
    return MODULE_RETURN_ALL(MODULE_MAKE_PAIR(CARTSLAM_KEY_PLANES, image_t, smoothed),
                             MODULE_MAKE_PAIR(CARTSLAM_KEY_PLANES_UNSMOOTHED, image_t, planes));
}

void SuperPixelDisparityPlaneSegmentationModule::updatePlaneParameters(System& system, SystemRunData& data) {
    auto histImage = data.getData<image_t>(CARTSLAM_KEY_DISPARITY_DERIVATIVE_HISTOGRAM);
    std::vector<int32_t> both(512), histogram(256);
    histImage->download(both.data(), 512 * sizeof(int32_t));
    for (int i = 0; i < 256; ++i) histogram[i] = both[2 * i];  // channel 0 = vertical (sp_planeseg.cu:358-359)
    {
        std::lock_guard<std::mutex> lock(derivativeHistogramMutex);
        if (!histogramCreated) {  // created as zeros; the first frame is not added (sp_planeseg.cu:364-366)
            histogramCreated = true;
            derivativeHistogram.assign(256, 0);
        } else {
            for (int i = 0; i < 256; ++i) {
                derivativeHistogram[i] += histogram[i];
                histogram[i] = (int32_t)derivativeHistogram[i];
            }
        }
        if (data.id % (updateInterval * resetInterval) == 1) std::fill(derivativeHistogram.begin(), derivativeHistogram.end(), 0);
    }
    if (data.id % updateInterval != 1) return;
    planeParameterProvider->updatePlaneParameters(system, data, histogram);
    system.insertGlobalData(CARTSLAM_KEY_PLANE_PARAMETERS, std::make_shared<PlaneParameters>(planeParameterProvider->getPlaneParameters()));
    system.insertGlobalData(CARTSLAM_KEY_DISPARITY_DERIVATIVE_HIST, std::make_shared<std::vector<int32_t>>(histogram));
}

}  // namespace cart
