// Host-side module layer of the B200 stereo front end: a drop-in for the shape of CART-SLAM's plugin
// surface on std:: primitives (the reference uses Boost futures / asio, absent here).
//   SystemModule / SyncWrapperSystemModule / module_dependency_t   /root/reference/include/modules/module.hpp:14-56
//   system_data_t, DataContainer (20 s wait timeout)               /root/reference/include/utils/data.hpp:11-77
//   MODULE_RETURN* macros                                           /root/reference/include/utils/modules.hpp:5-9
//   DataElement / StereoDataElement / DataSource                    /root/reference/include/datasource.hpp:11-83
//   SystemRunData / System                                          /root/reference/include/cartslam.hpp:27-113
// Payloads handed between modules are std::shared_ptr<void>, by convention a DeviceImage (the
// reference's convention is cv::cuda::GpuMat).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#define CARTSLAM_RUN_RETENTION 32
#define CARTSLAM_CONCURRENT_RUN_LIMIT 12
#define CARTSLAM_WORKER_THREADS (16 * CARTSLAM_CONCURRENT_RUN_LIMIT)
#define CARTSLAM_WAIT_FOR_DATA_TIMEOUT 20

namespace cart {

void logMessage(const char* level, const std::string& who, const std::string& what);
#define CART_LOG_INFO(who, what) ::cart::logMessage("INFO", (who), (what))
#define CART_LOG_WARN(who, what) ::cart::logMessage("WARN", (who), (what))
#define CART_LOG_ERROR(who, what) ::cart::logMessage("ERROR", (who), (what))

// ---- device images -------------------------------------------------------------------------------
enum ImageType { IMG_8UC1, IMG_8UC3, IMG_16SC1, IMG_16SC2, IMG_16UC1, IMG_32SC1, IMG_32SC2, IMG_32FC3 };
size_t imageElemBytes(ImageType t);

struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
};

// Ref-counted pitched device buffer (what cv::cuda::GpuMat is to the reference).
class DeviceImage {
   public:
    DeviceImage() = default;
    DeviceImage(int rows, int cols, ImageType type);
    bool empty() const { return !buf; }
    Size size() const { return Size(cols, rows); }
    void* ptr() const { return buf ? buf->p : nullptr; }
    template <typename T>
    T* as() const { return static_cast<T*>(ptr()); }
    void upload(const void* host, size_t hostPitch, void* stream = nullptr);   // H2D
    void download(void* host, size_t hostPitch, void* stream = nullptr) const;  // D2H (synchronises)
    int rows = 0, cols = 0;
    size_t pitch = 0;  // bytes
    ImageType type = IMG_8UC1;

   private:
    struct Buf {
        void* p = nullptr;
        size_t widthBytes = 0, rows = 0, pitch = 0;  // the pool key and what goes back to it
        ~Buf();
    };
    std::shared_ptr<Buf> buf;
};
typedef DeviceImage image_t;
void syncStream(void* stream);  // cudaStreamSynchronize; throws std::runtime_error on failure

// ---- data hand-off -------------------------------------------------------------------------------
typedef std::pair<std::string, std::shared_ptr<void>> system_data_pair_t;
typedef std::vector<system_data_pair_t> system_data_t;

#define MODULE_NO_RETURN_VALUE (std::vector<cart::system_data_pair_t>{})
#define MODULE_RETURN(key, value) (std::vector<cart::system_data_pair_t>{std::make_pair(key, value)})
#define MODULE_RETURN_ALL(...) (std::vector<cart::system_data_pair_t>{__VA_ARGS__})
#define MODULE_MAKE_PAIR(key, valueType, ...) std::make_pair(std::string(key), std::shared_ptr<void>(std::make_shared<valueType>(__VA_ARGS__)))
#define MODULE_RETURN_SHARED(key, valueType, ...) (std::vector<cart::system_data_pair_t>{MODULE_MAKE_PAIR(key, valueType, __VA_ARGS__)})

class DataNotAvailableException : public std::exception {
   public:
    explicit DataNotAvailableException(const std::string& key) : key("Missing key \"" + key + "\"") {}
    const char* what() const noexcept override { return key.c_str(); }

   private:
    const std::string key;
};

class ThreadPool {
   public:
    explicit ThreadPool(size_t n);
    ~ThreadPool();
    void post(std::function<void()> fn);
    void join();  // waits until the queue is drained and all workers are idle
    static bool onWorkerThread();  // true on a thread owned by (any) ThreadPool

   private:
    void work();
    std::vector<std::thread> workers;
    std::deque<std::function<void()>> queue;
    std::mutex m;
    std::condition_variable cv, idleCv;
    size_t busy = 0;
    bool stop = false;
};

class DataContainer {
   public:
    virtual ~DataContainer() = default;
    bool hasData(const std::string& key);
    std::vector<std::string> getDataKeys();
    template <typename T>
    std::shared_ptr<T> getData(const std::string& key) {
        std::unique_lock<std::mutex> lock(dataMutex);
        auto it = data.find(key);
        if (it == data.end()) throw std::invalid_argument("Could not find key \"" + key + "\"");
        return std::static_pointer_cast<T>(it->second);
    }
    // Blocks until every key is present; DataNotAvailableException after CARTSLAM_WAIT_FOR_DATA_TIMEOUT s.
    void waitForData(const std::vector<std::string>& keys);
    void insertData(system_data_pair_t entry);

   private:
    std::map<std::string, std::shared_ptr<void>> data;
    std::mutex dataMutex;
    std::condition_variable dataCondition;
};

// ---- data sources ----------------------------------------------------------------------------------
enum DataElementType { STEREO };

class DataElement {
   public:
    explicit DataElement(DataElementType type) : type(type) {}
    virtual ~DataElement() = default;
    const DataElementType type;
};

class StereoDataElement : public DataElement {
   public:
    StereoDataElement() : DataElement(STEREO) {}
    StereoDataElement(image_t left, image_t right) : DataElement(STEREO), left(left), right(right) {}
    image_t left, right;  // CV_8UC3 BGR on the device (datasource.cpp:6-16)
};

image_t getReferenceImage(std::shared_ptr<DataElement> element);

// CameraIntrinsics (/root/reference/include/datasource.hpp:11-18): the 4x4 matrix OpenCV uses to reproject disparity
// images into 3-D space, row-major floats.  Default: identity (the KITTI reader builds it at kitti.cpp:141-148).
struct CameraIntrinsics {
    float Q[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
};

class DataSource {
   public:
    explicit DataSource(Size imageSize) : imageSize(imageSize) {}
    const CameraIntrinsics getCameraIntrinsics() const { return intrinsics; }
    void setCameraIntrinsics(const CameraIntrinsics& in) { intrinsics = in; }
    virtual ~DataSource() = default;
    std::shared_ptr<DataElement> getNext(void* stream);
    virtual bool isNextReady() = 0;
    virtual bool isFinished() = 0;
    virtual DataElementType getProvidedType() = 0;
    Size getImageSize() const { return imageSize; }

   protected:
    virtual std::shared_ptr<DataElement> getNextInternal(void* stream) = 0;
    Size imageSize;
    CameraIntrinsics intrinsics;
};

// Frames handed over in host memory (tightly packed BGR); replaces the KITTI/ZED readers for the bar.
class MemoryDataSource : public DataSource {
   public:
    MemoryDataSource(Size size, int nFrames, const uint8_t* leftBgr, const uint8_t* rightBgr)
        : DataSource(size), n(nFrames), left(leftBgr), right(rightBgr) {}
    ~MemoryDataSource() override;
    bool isNextReady() override { return next < n; }
    bool isFinished() override { return next >= n; }
    DataElementType getProvidedType() override { return STEREO; }

   protected:
    std::shared_ptr<DataElement> getNextInternal(void* stream) override;

   private:
    int n, next = 0;
    const uint8_t *left, *right;
    // frames are staged through a pinned buffer and uploaded on a private non-blocking stream: a copy from pageable
    // memory on the legacy default stream serialises against everything else that is in flight
    uint8_t* staging = nullptr;
    void* copyStream = nullptr;
};

// ---- modules ---------------------------------------------------------------------------------------
class System;
class SystemRunData;

struct module_dependency_t {
    std::string name;
    int8_t runOffset;
    bool optional;
    module_dependency_t(const std::string& name, int runOffset, bool optional) : name(name), runOffset((int8_t)runOffset), optional(optional) {}
    module_dependency_t(const std::string& name, int runOffset) : module_dependency_t(name, runOffset, false) {}
    module_dependency_t(const std::string& name) : module_dependency_t(name, 0, false) {}
    module_dependency_t() : module_dependency_t("", 0, false) {}
};

class SystemModule {
   public:
    explicit SystemModule(const std::string& name) : name(name) {}
    virtual ~SystemModule() = default;
    virtual std::future<system_data_t> run(System& system, SystemRunData& data) = 0;
    const std::vector<module_dependency_t> getRequiredData() const { return requiresData; }
    const std::vector<std::string> getProvidedData() const { return providesData; }
    const std::string name;

   protected:
    std::vector<module_dependency_t> requiresData;
    std::vector<std::string> providesData;
};

class SyncWrapperSystemModule : public SystemModule {
   public:
    explicit SyncWrapperSystemModule(const std::string& name) : SystemModule(name) {}
    std::future<system_data_t> run(System& system, SystemRunData& data) override;  // runInternal on a pool thread
    virtual system_data_t runInternal(System& system, SystemRunData& data) = 0;

   private:
    // Runs enter runInternal in run-id order.  The reference serialises the stateful modules with a mutex only, so with
    // several frames in flight their warm-started state (superpixel labels, running histograms) depends on thread
    // timing; here the order is fixed, which makes the module layer reproduce the sequence runner for any worker count.
    system_data_t runOrdered(System& system, SystemRunData& data);
    std::mutex orderMutex;
    std::condition_variable orderCv;
    uint32_t lastRun = 0;  // id of the last run that went through this module
};

// ---- system ----------------------------------------------------------------------------------------
class SystemRunData : public DataContainer {
   public:
    SystemRunData(uint32_t id, System* system, std::shared_ptr<DataElement> dataElement)
        : dataElement(dataElement), id(id), system(system) {}
    void markAsComplete() { complete = true; }
    bool isComplete() { return complete; }
    std::shared_ptr<SystemRunData> getRelativeRun(int8_t offset);
    std::shared_ptr<DataElement> dataElement;
    const uint32_t id;

   private:
    std::atomic<bool> complete{false};
    System* system;
};

class System : public DataContainer {
   public:
    explicit System(std::shared_ptr<DataSource> dataSource, size_t workerThreads = CARTSLAM_WORKER_THREADS,
                    size_t runRetention = CARTSLAM_RUN_RETENTION, size_t concurrentRunLimit = CARTSLAM_CONCURRENT_RUN_LIMIT);
    ~System();
    std::future<void> run();  // one frame through every module, in dependency order

    template <typename T, typename... Args>
    void addModule(Args... args) {
        addModule(std::make_shared<T>(args...));
    }
    void addModule(std::shared_ptr<SystemModule> module);
    template <typename T>
    std::shared_ptr<T> getModule() {
        for (const auto& m : modules)
            if (auto c = std::dynamic_pointer_cast<T>(m)) return c;
        throw std::invalid_argument("Could not find module");
    }
    std::shared_ptr<SystemRunData> startNewRun(void* stream);
    std::shared_ptr<SystemRunData> getRunById(uint32_t id);
    uint8_t getActiveRunCount();
    ThreadPool& getThreadPool() { return threadPool; }
    void insertGlobalData(const std::string& key, std::shared_ptr<void> data) { insertData(std::make_pair(key, data)); }
    void insertGlobalData(system_data_pair_t data) { insertData(data); }
    std::shared_ptr<DataSource> getDataSource() const { return dataSource; }

   private:
    void verifyDependencies();
    void waitForDependencies(const std::vector<module_dependency_t>& deps, std::shared_ptr<SystemRunData> data);
    bool verifiedDependencies = false;
    std::vector<std::shared_ptr<SystemModule>> runOrder;  // modules in dependency order (same-run dependencies first)
    const size_t runRetention, concurrentRunLimit;
    ThreadPool threadPool;
    uint32_t runId = 0;
    std::shared_ptr<DataSource> dataSource;
    std::vector<std::shared_ptr<SystemModule>> modules;
    std::map<std::string, std::shared_ptr<SystemModule>> dataProvidedBy;
    std::vector<std::shared_ptr<SystemRunData>> runs;
    std::mutex runMutex;
    std::condition_variable runCondition;
};

}  // namespace cart
