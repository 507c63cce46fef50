// C entry point of the module layer, used by the tests (ctypes) and by integrators who want the whole
// reference-shaped pipeline behind one call: builds a System from a JSON module list (the reference's
// config/modules/*.json schema), feeds it `n` frames from host memory in id order, returns the planes.
#include <cstdlib>
#include <cstring>
#include <deque>
#include <exception>
#include <string>

#include "cart/modules.hpp"
#include "cart/sources.hpp"
#include "cart/inflate.hpp"

namespace {
std::string g_error;
void describe(const std::exception& e, std::string& out, int depth = 0) {
    out += std::string(depth ? " <- " : "") + e.what();
    try {
        std::rethrow_if_nested(e);
    } catch (const std::exception& n) {
        describe(n, out, depth + 1);
    } catch (...) {
    }
}
}  // namespace

extern "C" {

const char* cartb200_host_last_error() { return g_error.c_str(); }

// sequential != 0: wait for each frame before starting the next (the canonical in-order schedule);
// 0: up to CARTSLAM_CONCURRENT_RUN_LIMIT frames in flight like the reference's main loop.
// planes_out: n x H x W u8 (key "planes"); labels_out (nullable): n x H x W u16 (key "superpixels");
// disparity_out (nullable): n x H x W s16.
// q16 (nullable): row-major 4x4 reprojection matrix of the source (CameraIntrinsics::Q); depth_out (nullable):
// n x H x W x 3 float (key "depth").
int cartb200_host_run_config_ex(const char* modules_json, int skip_out_of_scope, int width, int height, int n,
                                const uint8_t* left_bgr, const uint8_t* right_bgr, int sequential, const float* q16,
                                uint8_t* planes_out, uint16_t* labels_out, int16_t* disparity_out, float* depth_out);

int cartb200_host_run_config(const char* modules_json, int skip_out_of_scope, int width, int height, int n,
                             const uint8_t* left_bgr, const uint8_t* right_bgr, int sequential, uint8_t* planes_out,
                             uint16_t* labels_out, int16_t* disparity_out) {
    return cartb200_host_run_config_ex(modules_json, skip_out_of_scope, width, height, n, left_bgr, right_bgr, sequential,
                                       nullptr, planes_out, labels_out, disparity_out, nullptr);
}

int cartb200_host_run_config_ex(const char* modules_json, int skip_out_of_scope, int width, int height, int n,
                                const uint8_t* left_bgr, const uint8_t* right_bgr, int sequential, const float* q16,
                                uint8_t* planes_out, uint16_t* labels_out, int16_t* disparity_out, float* depth_out) {
    using namespace cart;
    try {
        auto source = std::make_shared<MemoryDataSource>(Size(width, height), n, left_bgr, right_bgr);
        if (q16) {
            CameraIntrinsics in;
            std::memcpy(in.Q, q16, sizeof(in.Q));
            source->setCameraIntrinsics(in);
        }
        // CARTB200_HOST_WORKERS (testing aid): size of the System's thread pool (default CARTSLAM_WORKER_THREADS)
        size_t workers = CARTSLAM_WORKER_THREADS;
        if (const char* w = std::getenv("CARTB200_HOST_WORKERS"))
            if (std::atoi(w) > 0) workers = (size_t)std::atoi(w);
        auto system = std::make_shared<System>(source, workers);
        config::applyModuleConfigText(modules_json, system, skip_out_of_scope != 0);
        const size_t px = (size_t)width * height;
        std::deque<std::pair<uint32_t, std::future<void>>> inflight;
        auto collect = [&]() {
            const uint32_t fid = inflight.front().first;
            inflight.front().second.get();
            inflight.pop_front();
            auto run = system->getRunById(fid);  // still retained: we lag by < CARTSLAM_RUN_RETENTION frames
            if (planes_out && run->hasData(CARTSLAM_KEY_PLANES))
                run->getData<image_t>(CARTSLAM_KEY_PLANES)->download(planes_out + px * (fid - 1), (size_t)width);
            if (labels_out && run->hasData(CARTSLAM_KEY_SUPERPIXELS))
                run->getData<image_t>(CARTSLAM_KEY_SUPERPIXELS)->download(labels_out + px * (fid - 1), (size_t)width * 2);
            if (disparity_out && run->hasData(CARTSLAM_KEY_DISPARITY))
                run->getData<image_t>(CARTSLAM_KEY_DISPARITY)->download(disparity_out + px * (fid - 1), (size_t)width * 2);
            if (depth_out && run->hasData(CARTSLAM_KEY_DEPTH))
                run->getData<image_t>(CARTSLAM_KEY_DEPTH)->download(depth_out + px * 3 * (fid - 1), (size_t)width * 12);
        };
        uint32_t id = 0;
        while (!source->isFinished()) {
            auto f = system->run();
            inflight.emplace_back(++id, std::move(f));
            if (sequential || inflight.size() >= CARTSLAM_CONCURRENT_RUN_LIMIT) collect();
        }
        while (!inflight.empty()) collect();
        return 0;
    } catch (const std::exception& e) {
        g_error.clear();
        describe(e, g_error);
        return -1;
    }
}

// Decodes a PNG to BGR with the source layer's reader (no GPU needed).  out may be NULL to query the size.
int cartb200_host_decode_png(const char* path, uint8_t* out, size_t out_capacity, int* width, int* height) {
    try {
        std::vector<uint8_t> bgr;
        int w = 0, h = 0;
        cart::util::readPngBgr(path, bgr, w, h);
        if (width) *width = w;
        if (height) *height = h;
        if (out) {
            if (out_capacity < bgr.size()) throw std::runtime_error("output buffer too small");
            std::memcpy(out, bgr.data(), bgr.size());
        }
        return 0;
    } catch (const std::exception& e) {
        g_error.clear();
        describe(e, g_error);
        return -1;
    }
}

// The PNG reader's zlib-stream decoder on a caller's buffers (tests, tools/inflate_bench.py).  Returns 0 when `in` decodes
// to exactly out_size bytes with a matching Adler-32, -1 otherwise.  repeat > 1 decodes that many times (timing).
int cartb200_host_inflate(const uint8_t* in, size_t in_size, uint8_t* out, size_t out_size, int repeat) {
    if (!in || (!out && out_size)) return -1;
    try {
        std::vector<uint8_t> src(in_size + cart::png::kInflatePad, 0), dst(out_size + cart::png::kInflatePad);
        std::memcpy(src.data(), in, in_size);
        bool ok = true;
        for (int r = 0; r < (repeat > 1 ? repeat : 1) && ok; ++r) ok = cart::png::inflateZlib(src.data(), in_size, dst.data(), out_size);
        if (!ok) return -1;
        if (out_size) std::memcpy(out, dst.data(), out_size);
        return 0;
    } catch (const std::exception&) {
        return -1;
    }
}

// Builds the data source from a reference-style source config (config/sources/*.json), reports its image size and
// reprojection matrix, and - when modules_json is given - runs the module list over at most max_frames frames.
// Returns the number of frames processed (>= 0) or -1.  Output arrays hold max_frames frames.
int cartb200_host_run_source(const char* source_json, const char* modules_json, int skip_out_of_scope, int max_frames,
                             int* width, int* height, float* q16_out, uint8_t* planes_out, int16_t* disparity_out,
                             float* depth_out) {
    using namespace cart;
    try {
        auto source = config::createDataSourceFromText(source_json);
        const Size size = source->getImageSize();
        if (width) *width = size.width;
        if (height) *height = size.height;
        if (q16_out) std::memcpy(q16_out, source->getCameraIntrinsics().Q, 16 * sizeof(float));
        if (!modules_json) return 0;
        // CARTB200_HOST_WORKERS (testing aid): size of the System's thread pool (default CARTSLAM_WORKER_THREADS)
        size_t workers = CARTSLAM_WORKER_THREADS;
        if (const char* w = std::getenv("CARTB200_HOST_WORKERS"))
            if (std::atoi(w) > 0) workers = (size_t)std::atoi(w);
        auto system = std::make_shared<System>(source, workers);
        config::applyModuleConfigText(modules_json, system, skip_out_of_scope != 0);
        const size_t px = (size_t)size.width * size.height;
        // up to CARTSLAM_CONCURRENT_RUN_LIMIT frames in flight like the reference's main loop; the modules keep id order
        int done = 0, started = 0;
        std::deque<std::pair<uint32_t, std::future<void>>> inflight;
        auto collect = [&]() {
            const uint32_t fid = inflight.front().first;
            inflight.front().second.get();
            inflight.pop_front();
            auto run = system->getRunById(fid);
            if (planes_out && run->hasData(CARTSLAM_KEY_PLANES))
                run->getData<image_t>(CARTSLAM_KEY_PLANES)->download(planes_out + px * (fid - 1), (size_t)size.width);
            if (disparity_out && run->hasData(CARTSLAM_KEY_DISPARITY))
                run->getData<image_t>(CARTSLAM_KEY_DISPARITY)->download(disparity_out + px * (fid - 1), (size_t)size.width * 2);
            if (depth_out && run->hasData(CARTSLAM_KEY_DEPTH))
                run->getData<image_t>(CARTSLAM_KEY_DEPTH)->download(depth_out + px * 3 * (fid - 1), (size_t)size.width * 12);
            ++done;
        };
        while (started < max_frames && !source->isFinished()) {
            auto f = system->run();
            inflight.emplace_back((uint32_t)++started, std::move(f));
            if (inflight.size() >= CARTSLAM_CONCURRENT_RUN_LIMIT) collect();
        }
        while (!inflight.empty()) collect();
        return done;
    } catch (const std::exception& e) {
        g_error.clear();
        describe(e, g_error);
        return -1;
    }
}

}  // extern "C"
