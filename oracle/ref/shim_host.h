// TEST INFRASTRUCTURE ONLY.  Minimal stand-ins for the OpenCV / log4cxx / project types that the reference's HOST code
// for the plane-parameter estimation mentions (util::findPeaks, /root/reference/src/utils/peaks.cpp, and
// HistogramPeakPlaneParameterProvider::updatePlaneParameters, /root/reference/src/modules/planeseg/planeseg.cu:404-458),
// so that both can be compiled VERBATIM with g++.  Nothing here is reference code.
#pragma once
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <utility>
#include <vector>

namespace cv {
struct Mat {  // a 1 x cols row of int32 (CV_32SC1), as the reference downloads the derivative histogram
    int rows, cols;
    int* p;
    template <typename T>
    T& at(int i) { return reinterpret_cast<T*>(p)[i]; }
    template <typename T>
    const T& at(int i) const { return reinterpret_cast<const T*>(p)[i]; }
};
}  // namespace cv

namespace log4cxx {
typedef void* LoggerPtr;
}
#define LOG4CXX_WARN(logger, msg) do { } while (0)
#define LOG4CXX_DEBUG(logger, msg) do { } while (0)
using std::max;  // the reference's translation unit is CUDA C++, where min / max are global
using std::min;

namespace cart {
class System {};
class SystemRunData {};
namespace util {
class Peak;
}
// the members HistogramPeakPlaneParameterProvider::updatePlaneParameters writes (/root/reference/include/modules/planeseg.hpp:76-104)
class HistogramPeakPlaneParameterProvider {
   public:
    void updatePlaneParameters(log4cxx::LoggerPtr logger, System& system, SystemRunData& data, cv::Mat& histogram);
    std::pair<int, int> horizontalRange{0, 0}, verticalRange{0, 0};
    int horizontalCenter = 0, verticalCenter = 0;
};
}  // namespace cart
