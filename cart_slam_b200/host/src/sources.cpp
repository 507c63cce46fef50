// KITTI odometry data source + PNG reader (see cart/sources.hpp for the reference files this follows).
#include "cart/sources.hpp"
#include "cart/inflate.hpp"

#include <cuda_runtime.h>

#include "../../../include/cartb200.h"
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace cart {
namespace util {

std::string resolvePath(const std::string& path) {
    if (!path.empty() && path[0] == '~') {
        const char* home = getenv("HOME");
        if (home) return std::string(home) + path.substr(1);
    }
    return path;
}

namespace {
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

void readPngBgr(const std::string& path, std::vector<uint8_t>& bgr, int& width, int& height) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f.is_open()) throw std::runtime_error("Failed to open image " + path + ": " + strerror(errno));
    const std::streamsize fileSize = f.tellg();
    // the three working buffers (file, compressed stream, filtered scanlines) are kept per decode thread: a fresh 1.4 MB
    // vector per image costs its zero fill and its page faults again for every frame
    static thread_local std::vector<uint8_t> file, idat, raw;
    file.resize(fileSize > 0 ? (size_t)fileSize : 0);
    f.seekg(0);
    if (!file.empty() && !f.read(reinterpret_cast<char*>(file.data()), fileSize)) throw std::runtime_error("Failed to read image " + path);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 + 25 || std::memcmp(file.data(), sig, 8) != 0) throw std::runtime_error(path + ": not a PNG file");
    int bitDepth = 0, colorType = 0, interlace = 0;
    width = height = 0;
    std::vector<uint8_t> palette;
    idat.clear();
    size_t pos = 8;
    bool end = false;
    while (!end && pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
        const uint8_t* data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) throw std::runtime_error(path + ": truncated PNG chunk");
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) throw std::runtime_error(path + ": bad IHDR");
            width = (int)be32(data);
            height = (int)be32(data + 4);
            bitDepth = data[8];
            colorType = data[9];
            interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            palette.assign(data, data + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            end = true;
        }
        pos += 12 + (size_t)len;
    }
    if (width <= 0 || height <= 0 || width > 65535 || height > 65535) throw std::runtime_error(path + ": bad PNG size");
    if (bitDepth != 8 || interlace != 0) throw std::runtime_error(path + ": only 8-bit non-interlaced PNG files are supported");
    int ch;
    switch (colorType) {
        case 0: ch = 1; break;  // gray
        case 2: ch = 3; break;  // RGB
        case 3: ch = 1; break;  // palette
        case 4: ch = 2; break;  // gray + alpha
        case 6: ch = 4; break;  // RGBA
        default: throw std::runtime_error(path + ": unsupported PNG colour type");
    }
    if (colorType == 3 && palette.size() < 3) throw std::runtime_error(path + ": palette image without PLTE");
    const size_t stride = (size_t)width * ch;
    const size_t rawSize = (stride + 1) * (size_t)height;
    if (raw.size() < rawSize + png::kInflatePad) raw.resize(rawSize + png::kInflatePad);
    // own zlib-stream decoder (cart/inflate.hpp: whole-buffer, known output size); CARTB200_PNG_ZLIB=1 uses zlib's
    const size_t idatSize = idat.size();
    idat.resize(idatSize + png::kInflatePad, 0);
    static const bool useZlib = getenv("CARTB200_PNG_ZLIB") && atoi(getenv("CARTB200_PNG_ZLIB")) != 0;
    if (useZlib) {
        uLongf rawLen = (uLongf)rawSize;
        if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idatSize) != Z_OK || rawLen != rawSize)
            throw std::runtime_error(path + ": PNG inflate failed");
    } else if (!png::inflateZlib(idat.data(), idatSize, raw.data(), rawSize)) {
        throw std::runtime_error(path + ": PNG inflate failed");
    }
    // undo the scanline filters in place, one specialised loop per filter type (the per-byte dispatch of a generic
    // loop costs more than the inflate)
    bgr.resize((size_t)width * height * 3);
    const uint8_t* prev = nullptr;  // previous (unfiltered) row; nullptr = the all-zero row above the image
    for (int y = 0; y < height; ++y) {
        uint8_t* row = &raw[(stride + 1) * (size_t)y];
        const int filter = row[0];
        uint8_t* px = row + 1;
        const size_t bpp = (size_t)ch;
        switch (filter) {
            case 0: break;
            case 1:
                for (size_t i = bpp; i < stride; ++i) px[i] = (uint8_t)(px[i] + px[i - bpp]);
                break;
            case 2:
                if (prev)
                    for (size_t i = 0; i < stride; ++i) px[i] = (uint8_t)(px[i] + prev[i]);
                break;
            case 3:
                for (size_t i = 0; i < bpp && i < stride; ++i) px[i] = (uint8_t)(px[i] + ((prev ? prev[i] : 0) >> 1));
                for (size_t i = bpp; i < stride; ++i) px[i] = (uint8_t)(px[i] + ((px[i - bpp] + (prev ? prev[i] : 0)) >> 1));
                break;
            case 4:
                for (size_t i = 0; i < bpp && i < stride; ++i) px[i] = (uint8_t)(px[i] + (prev ? prev[i] : 0));  // paeth(0, b, 0) = b
                if (prev) {
                    for (size_t i = bpp; i < stride; ++i) px[i] = (uint8_t)(px[i] + paeth(px[i - bpp], prev[i], prev[i - bpp]));
                } else {
                    for (size_t i = bpp; i < stride; ++i) px[i] = (uint8_t)(px[i] + px[i - bpp]);  // paeth(a, 0, 0) = a
                }
                break;
            default: throw std::runtime_error(path + ": bad PNG filter");
        }
        prev = px;
        uint8_t* out = &bgr[(size_t)y * width * 3];
        if (colorType == 2) {  // RGB -> BGR (cv::imread(IMREAD_COLOR) returns BGR)
            for (int x = 0; x < width; ++x) {
                out[3 * x] = px[3 * x + 2];
                out[3 * x + 1] = px[3 * x + 1];
                out[3 * x + 2] = px[3 * x];
            }
        } else if (colorType == 6) {  // RGBA: alpha dropped
            for (int x = 0; x < width; ++x) {
                out[3 * x] = px[4 * x + 2];
                out[3 * x + 1] = px[4 * x + 1];
                out[3 * x + 2] = px[4 * x];
            }
        } else if (colorType == 3) {
            for (int x = 0; x < width; ++x) {
                const size_t k = (size_t)px[x] * 3;
                if (k + 2 >= palette.size()) throw std::runtime_error(path + ": palette index out of range");
                out[3 * x] = palette[k + 2];
                out[3 * x + 1] = palette[k + 1];
                out[3 * x + 2] = palette[k];
            }
        } else {  // gray (+ alpha): replicated
            for (int x = 0; x < width; ++x) out[3 * x] = out[3 * x + 1] = out[3 * x + 2] = px[(size_t)x * ch];
        }
    }
}
}  // namespace util

namespace sources {
namespace {
constexpr int kLeftCam = 2, kRightCam = 3;  // kitti.cpp:11-12

std::string addLeadingZeros(int number, size_t length) {  // kitti.cpp:18-21
    const std::string s = std::to_string(number);
    return std::string(length - std::min(length, s.length()), '0') + s;
}

struct Calibration {
    int cameraId = -1;
    float fx = 0, fy = 0, cx = 0, cy = 0, baseline = 0;
};

// One line of calib.txt, "P2: 12 floats" (kitti.cpp:32-85).  Like the reference, values are split at single spaces and
// the line is accepted only when exactly 11 separators were seen (the 12th value is never needed).
bool readLine(std::string line, Calibration& out) {
    size_t pos = line.find(": ");
    if (pos == std::string::npos) return false;
    std::string token = line.substr(0, pos);
    line.erase(0, pos + 2);
    if (token.empty() || token[0] != 'P') return false;
    Calibration local;
    try {
        local.cameraId = std::stoi(token.substr(1));
        float fubx = 0;
        int i = 0;
        while ((pos = line.find(' ')) != std::string::npos) {
            token = line.substr(0, pos);
            line.erase(0, pos + 1);
            switch (i) {
                case 0: local.fx = std::stof(token); break;
                case 5: local.fy = std::stof(token); break;
                case 3: fubx = std::stof(token); break;
                case 2: local.cx = std::stof(token); break;
                case 6: local.cy = std::stof(token); break;
            }
            i++;
        }
        if (i != 11) return false;
        local.baseline = -fubx / local.fx;
    } catch (const std::exception&) {
        return false;
    }
    out = local;
    return true;
}
}  // namespace

KITTIDataSource::KITTIDataSource(std::string basePath, int sequence, Size imageSize)
    : DataSource(imageSize), path(util::resolvePath(basePath + "/sequences/" + addLeadingZeros(sequence, 2))) {
    init();
}

KITTIDataSource::KITTIDataSource(std::string path_, Size imageSize) : DataSource(imageSize), path(util::resolvePath(path_)) { init(); }

std::string KITTIDataSource::framePath(int cam, int frame) const {
    return path + "/image_" + std::to_string(cam) + "/" + addLeadingZeros(frame, 6) + ".png";
}

void KITTIDataSource::init() {
    const std::string calibPath = path + "/calib.txt";
    std::ifstream in(calibPath);
    if (!in.is_open()) throw std::runtime_error("Failed to open calibration file at " + calibPath + ": " + strerror(errno));
    Calibration left, right;
    bool haveLeft = false, haveRight = false;
    std::string line;
    while (std::getline(in, line)) {
        Calibration c;
        if (!readLine(line, c)) continue;
        if (c.cameraId == kLeftCam) {
            left = c;
            haveLeft = true;
        } else if (c.cameraId == kRightCam) {
            right = c;
            haveRight = true;
        }
    }
    if (!haveLeft || !haveRight) throw std::runtime_error("Failed to read calibration file");
    // the first image gives the native size (kitti.cpp:129-135)
    int w = 0, h = 0;
    util::readPngBgr(framePath(kLeftCam, 0), bufL, w, h);
    if (imageSize.width == 0 || imageSize.height == 0) imageSize = Size(w, h);
    fileSize = Size(w, h);  // frames of another size than imageSize are resized on the device (kitti.cpp:166-169)
    const float scaleWidth = (float)imageSize.width / (float)w, scaleHeight = (float)imageSize.height / (float)h;
    // reprojection matrix, kitti.cpp:141-148
    CameraIntrinsics k;
    k.Q[0 * 4 + 3] = -left.cx * scaleWidth;
    k.Q[1 * 4 + 3] = -left.cy * scaleHeight;
    k.Q[2 * 4 + 2] = 0;
    k.Q[2 * 4 + 3] = left.fx * scaleWidth;
    k.Q[3 * 4 + 2] = (float)(-1.0 / left.baseline);
    k.Q[3 * 4 + 3] = (left.cx - right.cx) * scaleWidth / left.baseline;
    intrinsics = k;
}

bool KITTIDataSource::isNextReady() {
    struct stat st;
    return stat(framePath(kLeftCam, currentFrame).c_str(), &st) == 0;
}

bool KITTIDataSource::isFinished() { return !isNextReady(); }

KITTIDataSource::~KITTIDataSource() {
    {
        std::lock_guard<std::mutex> lock(ringMutex);
        stopPrefetch = true;
    }
    ringCv.notify_all();
    for (auto& w : workers) w.join();
    for (auto& s : ring) {
        cudaFreeHost(s.left);
        cudaFreeHost(s.right);
    }
    if (copyStream) cudaStreamDestroy((cudaStream_t)copyStream);
}

void KITTIDataSource::startPrefetch() {
    unsigned threads = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("CARTB200_KITTI_DECODE_THREADS"))
        if (atoi(e) > 0) threads = (unsigned)atoi(e);
    size_t depth = 2 * (size_t)threads;
    if (const char* e = getenv("CARTB200_KITTI_RING"))
        if (atoi(e) > 0) depth = (size_t)atoi(e);
    const size_t bytes = (size_t)fileSize.width * fileSize.height * 3;
    ring.resize(depth);
    for (auto& s : ring)
        if (cudaMallocHost((void**)&s.left, bytes) != cudaSuccess || cudaMallocHost((void**)&s.right, bytes) != cudaSuccess)
            throw std::runtime_error("KITTIDataSource: cudaMallocHost of the frame ring failed");
    for (unsigned i = 0; i < threads; ++i) workers.emplace_back([this] { prefetchWorker(); });
}

// Each worker claims the next frame whose slot is free, decodes both images into the slot and marks it ready.
void KITTIDataSource::prefetchWorker() {
    std::vector<uint8_t> tmp;
    const size_t bytes = (size_t)fileSize.width * fileSize.height * 3;
    for (;;) {
        int frame;
        Slot* slot;
        {
            std::unique_lock<std::mutex> lock(ringMutex);
            ringCv.wait(lock, [this] { return stopPrefetch || (!sawEnd && ring[(size_t)nextToLoad % ring.size()].state == Slot::FREE); });
            if (stopPrefetch) return;
            frame = nextToLoad++;
            slot = &ring[(size_t)frame % ring.size()];
            slot->state = Slot::LOADING;
            slot->frame = frame;
        }
        Slot::State result = Slot::READY;
        std::string error;
        struct stat st;
        if (stat(framePath(kLeftCam, frame).c_str(), &st) != 0) {
            result = Slot::END;
        } else {
            try {
                int w = 0, h = 0;
                for (int cam = 0; cam < 2; ++cam) {
                    util::readPngBgr(framePath(cam ? kRightCam : kLeftCam, frame), tmp, w, h);
                    if (w != fileSize.width || h != fileSize.height)
                        throw std::runtime_error("KITTIDataSource: frame " + std::to_string(frame) + " has a different size");
                    std::memcpy(cam ? slot->right : slot->left, tmp.data(), bytes);
                }
            } catch (const std::exception& e) {
                result = Slot::FAILED;
                error = e.what();
            }
        }
        {
            std::lock_guard<std::mutex> lock(ringMutex);
            slot->state = result;
            slot->error = error;
            if (result == Slot::END) sawEnd = true;  // frames are consecutive: nothing beyond the first gap is read
        }
        ringCv.notify_all();
    }
}

std::shared_ptr<DataElement> KITTIDataSource::getNextInternal(void* stream) {
    if (ring.empty()) startPrefetch();
    Slot* slot = &ring[(size_t)currentFrame % ring.size()];
    {
        std::unique_lock<std::mutex> lock(ringMutex);
        ringCv.wait(lock, [&] { return slot->frame == currentFrame && slot->state != Slot::LOADING && slot->state != Slot::FREE; });
        if (slot->state == Slot::END) throw std::runtime_error("KITTIDataSource: frame " + std::to_string(currentFrame) + " does not exist");
        if (slot->state == Slot::FAILED) throw std::runtime_error(slot->error);
    }
    image_t l(imageSize.height, imageSize.width, IMG_8UC3), r(imageSize.height, imageSize.width, IMG_8UC3);
    if (!copyStream) {
        cudaStream_t s = nullptr;
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) throw std::runtime_error("cudaStreamCreate failed");
        copyStream = s;
    }
    (void)stream;  // the caller's stream is the legacy default stream in System::startNewRun: use a private one
    if (fileSize.width == imageSize.width && fileSize.height == imageSize.height) {
        l.upload(slot->left, (size_t)imageSize.width * 3, copyStream);  // pinned source: truly asynchronous copies
        r.upload(slot->right, (size_t)imageSize.width * 3, copyStream);
    } else {  // cv::cuda::resize(..., INTER_LINEAR) of both images, kitti.cpp:166-169
        image_t nl(fileSize.height, fileSize.width, IMG_8UC3), nr(fileSize.height, fileSize.width, IMG_8UC3);
        nl.upload(slot->left, (size_t)fileSize.width * 3, copyStream);
        nr.upload(slot->right, (size_t)fileSize.width * 3, copyStream);
        if (cartb200_resize_bgr8(nl.as<uint8_t>(), nl.pitch, fileSize.width, fileSize.height, l.as<uint8_t>(), l.pitch,
                                 imageSize.width, imageSize.height, copyStream) != 0 ||
            cartb200_resize_bgr8(nr.as<uint8_t>(), nr.pitch, fileSize.width, fileSize.height, r.as<uint8_t>(), r.pitch,
                                 imageSize.width, imageSize.height, copyStream) != 0)
            throw std::runtime_error("KITTIDataSource: resize failed");
        syncStream(copyStream);  // nl / nr go back to the image pool when they leave this scope
    }
    syncStream(copyStream);  // the slot goes back to the decoders
    {
        std::lock_guard<std::mutex> lock(ringMutex);
        slot->state = Slot::FREE;
        slot->frame = -1;
    }
    ringCv.notify_all();
    ++currentFrame;
    return std::make_shared<StereoDataElement>(l, r);
}
}  // namespace sources
}  // namespace cart
