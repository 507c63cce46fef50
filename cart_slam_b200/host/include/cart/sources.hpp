// On-disk data sources of the reference, re-created without OpenCV:
//   KITTIDataSource   /root/reference/include/sources/kitti.hpp:9-24, src/sources/kitti.cpp:32-186
//                     (image_2/%06d.png + image_3/%06d.png + calib.txt -> reprojection matrix Q)
//   createDataSource  /root/reference/src/cartconfig.cpp:82-104 ({"type": "kitti", "path", "sequence"})
// cv::imread is replaced by a small PNG reader (8-bit gray / RGB / RGBA / palette, non-interlaced; zlib inflate),
// which returns BGR like IMREAD_COLOR does.  The ZED source needs the proprietary ZED SDK: out of scope.
#pragma once
#include <string>
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
class KITTIDataSource : public DataSource {
   public:
    KITTIDataSource(std::string basePath, int sequence, Size imageSize = Size(0, 0));
    explicit KITTIDataSource(std::string path, Size imageSize = Size(0, 0));
    bool isNextReady() override;
    bool isFinished() override;
    DataElementType getProvidedType() override { return STEREO; }

   protected:
    std::shared_ptr<DataElement> getNextInternal(void* stream) override;

   private:
    void init();
    std::string framePath(int cam, int frame) const;
    std::string path;
    int currentFrame = 0;
    std::vector<uint8_t> bufL, bufR;
};
}  // namespace sources

namespace config {
// Same JSON schema as the reference's config/sources/*.json
std::shared_ptr<DataSource> createDataSourceFromText(const std::string& jsonText);
std::shared_ptr<DataSource> readDataSourceConfig(const std::string& path);
}  // namespace config
}  // namespace cart
