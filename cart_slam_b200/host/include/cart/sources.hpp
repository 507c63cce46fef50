// On-disk data sources of the reference, re-created without OpenCV:
//   KITTIDataSource   /root/reference/include/sources/kitti.hpp:9-24, src/sources/kitti.cpp:32-186
//                     (image_2/%06d.png + image_3/%06d.png + calib.txt -> reprojection matrix Q)
//   createDataSource  /root/reference/src/cartconfig.cpp:82-104 ({"type": "kitti", "path", "sequence"})
// cv::imread is replaced by a small PNG reader (8-bit gray / RGB / RGBA / palette, non-interlaced; zlib inflate),
// which returns BGR like IMREAD_COLOR does.  The ZED source needs the proprietary ZED SDK: out of scope.
#pragma once
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "cart/core.hpp"

namespace cart {
namespace util {
// /root/reference/include/utils/path.hpp:5-15
std::string resolvePath(const std::string& path);
// Decodes a PNG file to tightly packed 8-bit BGR.  Throws std::runtime_error on malformed / unsupported files.
void readPngBgr(const std::string& path, std::vector<uint8_t>& bgr, int& width, int& height);
}  // namespace util

namespace sources {
// Frames are decoded ahead of the consumer by a small pool of threads (PNG inflate costs ~10 ms per KITTI image and
// core) into a ring of PINNED host slots; getNext() uploads the ready slot with two asynchronous copies on the caller's
// stream.  CARTB200_KITTI_DECODE_THREADS / CARTB200_KITTI_RING override the pool size (default min(cores, 16)) and the
// ring depth (default 2 x threads).  The reference reads and uploads one frame at a time inside getNextInternal
// (kitti.cpp:155-186).
class KITTIDataSource : public DataSource {
   public:
    KITTIDataSource(std::string basePath, int sequence, Size imageSize = Size(0, 0));
    explicit KITTIDataSource(std::string path, Size imageSize = Size(0, 0));
    ~KITTIDataSource() override;
    bool isNextReady() override;
    bool isFinished() override;
    DataElementType getProvidedType() override { return STEREO; }

   protected:
    std::shared_ptr<DataElement> getNextInternal(void* stream) override;

   private:
    void init();
    std::string framePath(int cam, int frame) const;
    void startPrefetch();
    void prefetchWorker();
    std::string path;
    Size fileSize;  // size of the PNG files; imageSize (DataSource) is what the modules get
    int currentFrame = 0;
    std::vector<uint8_t> bufL, bufR;
    // prefetch ring: frame f lives in slot f % ring.size()
    struct Slot {
        uint8_t* left = nullptr;   // pinned, W*H*3 each
        uint8_t* right = nullptr;
        int frame = -1;
        enum State { FREE, LOADING, READY, FAILED, END } state = FREE;
        std::string error;
    };
    std::vector<Slot> ring;
    std::vector<std::thread> workers;
    std::mutex ringMutex;
    std::condition_variable ringCv;
    int nextToLoad = 0;
    bool stopPrefetch = false, sawEnd = false;
    void* copyStream = nullptr;
};
}  // namespace sources

namespace config {
// Same JSON schema as the reference's config/sources/*.json
std::shared_ptr<DataSource> createDataSourceFromText(const std::string& jsonText);
std::shared_ptr<DataSource> readDataSourceConfig(const std::string& path);
}  // namespace config
}  // namespace cart
