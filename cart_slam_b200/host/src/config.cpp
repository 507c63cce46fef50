// JSON module configuration with the reference's keys and defaults
// (/root/reference/src/cartconfig.cpp:56-80 parameter providers, :106-228 module switch).
#include <cmath>
#include <fstream>
#include <set>
#include <sstream>

#include "../../../third_party/nlohmann/json.hpp"
#include "cart/modules.hpp"
#include "cart/sources.hpp"

namespace cart::config {

namespace {
using nlohmann::json;

template <typename T>
T get(const json& data, const std::string& key, const T& def) {
    auto it = data.find(key);
    return it == data.end() ? def : it->get<T>();
}
template <typename T>
T need(const json& data, const std::string& key) {
    auto it = data.find(key);
    if (it == data.end()) throw std::runtime_error("Key " + key + " not found.");
    return it->get<T>();
}

std::shared_ptr<PlaneParameterProvider> readParameterProvider(const json& data) {
    if (data.find("type") == data.end()) throw std::runtime_error("Parameter provider type not found.");
    const std::string type = data["type"].get<std::string>();
    if (type == "static") {
        auto h = std::make_pair(need<int>(data, "horizontal_range_min"), need<int>(data, "horizontal_range_max"));
        auto v = std::make_pair(need<int>(data, "vertical_range_min"), need<int>(data, "vertical_range_max"));
        return std::make_shared<StaticPlaneParameterProvider>((h.first + h.second) / 2, (v.first + v.second) / 2, h, v);
    }
    if (type == "histogram_peak") return std::make_shared<HistogramPeakPlaneParameterProvider>();
    throw std::runtime_error("Unknown parameter provider type.");
}

// module types of the reference that are outside the hot-path scope
const std::set<std::string> kOutOfScope = {
    "superpixels_visualization", "depth_visualization", "zed_disparity", "disparity_visualization",
    "disparity_derivative_visualization", "features", "features_visualization", "optflow", "optflow_visualization",
    "planefit", "planecluster", "planefit_visualization", "disparity_planeseg_visualization", "bev_planeseg_visualization"};
}  // namespace

void applyModuleConfigText(const std::string& text, std::shared_ptr<System> system, bool skipOutOfScope) {
    const json modules = json::parse(text);
    if (!modules.is_array()) throw std::runtime_error("Modules configuration is not an array.");
    const Size size = system->getDataSource()->getImageSize();
    bool haveFlow = false;
    for (const auto& m : modules)
        if (m.is_object() && m.value("type", std::string()) == "external_optflow") haveFlow = true;
    for (const auto& m : modules) {
        if (!m.is_object()) throw std::runtime_error("Module configuration is not an object.");
        const std::string type = m.at("type").get<std::string>();
        if (type == "superpixels") {
            const double direct = get(m, "direct_clique_cost", 0.5);
            system->addModule<SuperPixelModule>(size, (unsigned)get(m, "initial_iterations", 18), (unsigned)get(m, "iterations", 6),
                                                (unsigned)get(m, "block_size", 12), (unsigned)get(m, "reset_iterations", 64), direct,
                                                get(m, "diagonal_clique_cost", direct / std::sqrt(2.0)), get(m, "compactness_weight", 0.1),
                                                get(m, "progressive_compactness_cost", 0.0), get(m, "image_weight", 1.5),
                                                get(m, "disparity_weight", 1.0));
        } else if (type == "disparity") {
            system->addModule<ImageDisparityModule>(size, get(m, "min_disparity", 4), get(m, "num_disparities", 256), get(m, "block_size", 3),
                                                    get(m, "smoothing_radius", -1), get(m, "smoothing_iterations", 5));
        } else if (type == "disparity_derivative") {
            system->addModule<ImageDisparityDerivativeModule>();
        } else if (type == "depth") {
            system->addModule<DepthModule>();
        } else if (type == "external_optflow") {  // this build's stand-in for the NVOFA "optflow" module
            system->addModule<ExternalOpticalFlowModule>(get(m, "flow_x", 0.0), get(m, "flow_y", 0.0));
            haveFlow = true;
        } else if (type == "disparity_planeseg" || type == "superpixel_disparity_planeseg") {
            auto provider = readParameterProvider(m.at("parameter_provider"));
            bool temporal = get(m, "use_temporal_smoothing", false);
            if (temporal && skipOutOfScope && !haveFlow) {
                CART_LOG_WARN("config", type + ": use_temporal_smoothing needs a module that provides optflow - disabled");
                temporal = false;
            }
            const int update = get(m, "update_interval", 30), reset = get(m, "reset_interval", 10);
            const unsigned dist = (unsigned)get(m, "temporal_smoothing_distance", CARTSLAM_PLANE_TEMPORAL_DISTANCE_DEFAULT);
            if (type == "disparity_planeseg")
                system->addModule<DisparityPlaneSegmentationModule>(provider, update, reset, temporal, dist);
            else
                system->addModule<SuperPixelDisparityPlaneSegmentationModule>(provider, update, reset, temporal, dist);
        } else if (kOutOfScope.count(type)) {
            if (!skipOutOfScope)
                throw std::runtime_error("Module type " + type + " is outside the scope of the B200 disparity->planeseg path.");
            CART_LOG_WARN("config", "skipping out-of-scope module " + type);
        } else {
            throw std::runtime_error("Unknown module type " + type + ".");
        }
    }
}

void readModuleConfig(const std::string& path, std::shared_ptr<System> system, bool skipOutOfScope) {
    std::ifstream file(path);
    if (!file.is_open()) throw std::runtime_error("Could not open file " + path);
    std::stringstream ss;
    ss << file.rdbuf();
    applyModuleConfigText(ss.str(), system, skipOutOfScope);
}

// createDataSource, /root/reference/src/cartconfig.cpp:82-104
std::shared_ptr<DataSource> createDataSourceFromText(const std::string& text) {
    const json cfg = json::parse(text);
    if (!cfg.is_object()) throw std::runtime_error("Data source configuration is not an object.");
    const std::string sourcePath = cfg.at("path").get<std::string>();
    const std::string type = cfg.at("type").get<std::string>();
    // "width" / "height" (extension): the imageSize constructor argument of KITTIDataSource (kitti.hpp:12), which the
    // reference's JSON does not expose; frames are resized on the device when it differs from the files' size
    if (type == "kitti")
        return std::make_shared<sources::KITTIDataSource>(sourcePath, get(cfg, "sequence", 0),
                                                          Size(get(cfg, "width", 0), get(cfg, "height", 0)));
    if (type == "zed") throw std::runtime_error("Data source type zed needs the proprietary ZED SDK (outside the scope of this build).");
    throw std::runtime_error("Unknown data source type.");
}

std::shared_ptr<DataSource> readDataSourceConfig(const std::string& path) {
    std::ifstream file(path);
    if (!file.is_open()) throw std::runtime_error("Could not open file " + path);
    std::stringstream ss;
    ss << file.rdbuf();
    return createDataSourceFromText(ss.str());
}

}  // namespace cart::config
