// Host runtime of the module layer: device images, data containers, thread pool, data sources and the
// System scheduler (dependency wait -> module run -> insert results), restating
// /root/reference/src/cartslam.cpp:65-333, /root/reference/src/utils/data.cpp:7-56,
// /root/reference/src/modules/module.cpp:7-27 and /root/reference/src/datasource.cpp:19-56 on std:: primitives.
#include "cart/core.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

namespace cart {

void logMessage(const char* level, const std::string& who, const std::string& what) {
    static std::mutex m;
    static const bool quiet = std::getenv("CARTB200_QUIET") != nullptr;
    if (quiet && level[0] == 'I') return;
    std::lock_guard<std::mutex> lock(m);
    std::fprintf(stderr, "[%s] %s: %s\n", level, who.c_str(), what.c_str());
}

static void cudaCheck(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

size_t imageElemBytes(ImageType t) {
    switch (t) {
        case IMG_8UC1: return 1;
        case IMG_8UC3: return 3;
        case IMG_16SC1: return 2;
        case IMG_16SC2: return 4;
        case IMG_16UC1: return 2;
        case IMG_32SC1: return 4;
        case IMG_32SC2: return 8;
        case IMG_32FC3: return 12;
    }
    return 1;
}

// Device images are recycled through a size-keyed pool: a frame allocates about ten of them (inputs, disparity,
// derivatives, labels, planes ...), and cudaFree synchronises the whole device - with cudaMallocPitch / cudaFree per
// image (what cv::cuda::GpuMat does without a BufferPool) the module layer spent ~10 ms of host time per frame on them.
// Buffers return to the pool when the last DeviceImage that shares them goes away; the pool is capped
// (CARTB200_IMAGE_POOL_MB, default 4096) and anything beyond the cap is really freed.
namespace {
struct ImagePool {
    std::mutex m;
    std::map<std::pair<size_t, size_t>, std::vector<std::pair<void*, size_t>>> free;  // (width bytes, rows) -> (ptr, pitch)
    size_t bytes = 0, cap = 4096ull << 20;
    ImagePool() {
        if (const char* e = std::getenv("CARTB200_IMAGE_POOL_MB")) cap = (size_t)std::max(0, std::atoi(e)) << 20;
    }
    ~ImagePool() {  // process exit: the driver may already be gone, ignore errors
        for (auto& kv : free)
            for (auto& b : kv.second) cudaFree(b.first);
    }
    bool take(size_t widthBytes, size_t rows, void** p, size_t* pitch) {
        std::lock_guard<std::mutex> lock(m);
        auto it = free.find({widthBytes, rows});
        if (it == free.end() || it->second.empty()) return false;
        *p = it->second.back().first;
        *pitch = it->second.back().second;
        it->second.pop_back();
        bytes -= *pitch * rows;
        return true;
    }
    void give(size_t widthBytes, size_t rows, void* p, size_t pitch) {
        {
            std::lock_guard<std::mutex> lock(m);
            if (bytes + pitch * rows <= cap) {
                free[{widthBytes, rows}].emplace_back(p, pitch);
                bytes += pitch * rows;
                return;
            }
        }
        cudaFree(p);
    }
};
ImagePool& imagePool() {
    static ImagePool* pool = new ImagePool();  // leaked on purpose: images may outlive static destruction order
    return *pool;
}
}  // namespace

DeviceImage::Buf::~Buf() {
    if (p) imagePool().give(widthBytes, rows, p, pitch);
}

DeviceImage::DeviceImage(int rows_, int cols_, ImageType type_) : rows(rows_), cols(cols_), type(type_) {
    buf = std::make_shared<Buf>();
    buf->widthBytes = (size_t)cols * imageElemBytes(type);
    buf->rows = (size_t)rows;
    if (!imagePool().take(buf->widthBytes, buf->rows, &buf->p, &pitch))
        cudaCheck(cudaMallocPitch(&buf->p, &pitch, buf->widthBytes, rows), "cudaMallocPitch");
    buf->pitch = pitch;
}

void DeviceImage::upload(const void* host, size_t hostPitch, void* stream) {
    cudaCheck(cudaMemcpy2DAsync(ptr(), pitch, host, hostPitch, (size_t)cols * imageElemBytes(type), rows,
                                cudaMemcpyHostToDevice, (cudaStream_t)stream),
              "upload");
}

void DeviceImage::download(void* host, size_t hostPitch, void* stream) const {
    cudaCheck(cudaMemcpy2DAsync(host, hostPitch, ptr(), pitch, (size_t)cols * imageElemBytes(type), rows,
                                cudaMemcpyDeviceToHost, (cudaStream_t)stream),
              "download");
    cudaCheck(cudaStreamSynchronize((cudaStream_t)stream), "download sync");
}

// ---- thread pool -----------------------------------------------------------------------------------
ThreadPool::ThreadPool(size_t n) {
    for (size_t i = 0; i < std::max<size_t>(1, n); ++i) workers.emplace_back([this] { work(); });
}

ThreadPool::~ThreadPool() {
    {
        std::lock_guard<std::mutex> lock(m);
        stop = true;
    }
    cv.notify_all();
    for (auto& w : workers) w.join();
}

void ThreadPool::post(std::function<void()> fn) {
    {
        std::lock_guard<std::mutex> lock(m);
        queue.push_back(std::move(fn));
    }
    cv.notify_one();
}

namespace {
thread_local bool tlsPoolWorker = false;
}
bool ThreadPool::onWorkerThread() { return tlsPoolWorker; }

void ThreadPool::work() {
    tlsPoolWorker = true;
    for (;;) {
        std::function<void()> fn;
        {
            std::unique_lock<std::mutex> lock(m);
            cv.wait(lock, [this] { return stop || !queue.empty(); });
            if (stop && queue.empty()) return;
            fn = std::move(queue.front());
            queue.pop_front();
            ++busy;
        }
        fn();
        {
            std::lock_guard<std::mutex> lock(m);
            --busy;
        }
        idleCv.notify_all();
    }
}

void ThreadPool::join() {
    std::unique_lock<std::mutex> lock(m);
    idleCv.wait(lock, [this] { return queue.empty() && busy == 0; });
}

// ---- data container --------------------------------------------------------------------------------
bool DataContainer::hasData(const std::string& key) {
    std::lock_guard<std::mutex> lock(dataMutex);
    return data.count(key) != 0;
}

std::vector<std::string> DataContainer::getDataKeys() {
    std::lock_guard<std::mutex> lock(dataMutex);
    std::vector<std::string> keys;
    for (const auto& p : data) keys.push_back(p.first);
    return keys;
}

void DataContainer::insertData(system_data_pair_t entry) {
    {
        std::lock_guard<std::mutex> lock(dataMutex);
        data[entry.first] = entry.second;
    }
    dataCondition.notify_all();
}

void DataContainer::waitForData(const std::vector<std::string>& keys) {
    if (keys.empty()) throw std::invalid_argument("No keys provided to wait for");
    for (const auto& key : keys) {
        std::unique_lock<std::mutex> lock(dataMutex);
        // a wait of more than a few seconds means the producing run has failed (data.cpp:41-49)
        if (!dataCondition.wait_for(lock, std::chrono::seconds(CARTSLAM_WAIT_FOR_DATA_TIMEOUT),
                                    [this, &key] { return data.count(key) != 0; }))
            throw DataNotAvailableException(key);
    }
}

// ---- data sources ----------------------------------------------------------------------------------
image_t getReferenceImage(std::shared_ptr<DataElement> element) {
    switch (element->type) {
        case STEREO: return std::static_pointer_cast<StereoDataElement>(element)->left;
        default: throw std::runtime_error("Unknown data element type");
    }
}

void syncStream(void* stream) { cudaCheck(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize"); }

std::shared_ptr<DataElement> DataSource::getNext(void* stream) {
    if (!isNextReady()) throw std::runtime_error("Next element is not ready!");
    auto element = getNextInternal(stream);
    if (element->type == STEREO) {
        auto st = std::static_pointer_cast<StereoDataElement>(element);
        if (st->left.type != IMG_8UC3 || st->right.type != IMG_8UC3)  // datasource.cpp:6-16 forces CV_8UC3
            throw std::runtime_error("stereo images must be CV_8UC3 BGR");
    }
    return element;
}

MemoryDataSource::~MemoryDataSource() {
    if (staging) cudaFreeHost(staging);
    if (copyStream) cudaStreamDestroy((cudaStream_t)copyStream);
}

std::shared_ptr<DataElement> MemoryDataSource::getNextInternal(void*) {
    const size_t frame = (size_t)imageSize.width * imageSize.height * 3;
    if (!staging) {
        cudaCheck(cudaMallocHost((void**)&staging, 2 * frame), "cudaMallocHost");
        cudaStream_t s = nullptr;
        cudaCheck(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
        copyStream = s;
    }
    image_t l(imageSize.height, imageSize.width, IMG_8UC3), r(imageSize.height, imageSize.width, IMG_8UC3);
    std::memcpy(staging, left + frame * next, frame);
    std::memcpy(staging + frame, right + frame * next, frame);
    l.upload(staging, (size_t)imageSize.width * 3, copyStream);
    r.upload(staging + frame, (size_t)imageSize.width * 3, copyStream);
    cudaCheck(cudaStreamSynchronize((cudaStream_t)copyStream), "frame upload");
    ++next;
    return std::make_shared<StereoDataElement>(l, r);
}

// ---- modules ---------------------------------------------------------------------------------------
// The reference posts runInternal to the pool and chains non-blocking continuations (module.cpp:7-19).  Here the caller
// of run() inside System::run already owns a pool worker and blocks on the future, so a second post would need a second
// worker for the same module: with few workers every one of them would be such a waiter and nothing could run.  On a
// pool thread the work is therefore done inline and a ready future is returned; other callers get the posted task.
system_data_t SyncWrapperSystemModule::runOrdered(System& system, SystemRunData& data) {
    {
        std::unique_lock<std::mutex> lock(orderMutex);
        // ids below the expected one (a second System, a restarted source) are let through unordered
        orderCv.wait(lock, [&] { return data.id <= lastRun + 1; });
    }
    struct Advance {
        SyncWrapperSystemModule* m;
        uint32_t id;
        ~Advance() {
            {
                std::lock_guard<std::mutex> lock(m->orderMutex);
                if (id > m->lastRun) m->lastRun = id;
            }
            m->orderCv.notify_all();
        }
    } advance{this, data.id};  // also on an exception: later runs must not wait for a failed one
    return runInternal(system, data);
}

std::future<system_data_t> SyncWrapperSystemModule::run(System& system, SystemRunData& data) {
    auto task = std::make_shared<std::packaged_task<system_data_t()>>([this, &system, &data] { return runOrdered(system, data); });
    auto future = task->get_future();
    if (ThreadPool::onWorkerThread())
        (*task)();
    else
        system.getThreadPool().post([task] { (*task)(); });
    return future;
}

// ---- system ----------------------------------------------------------------------------------------
std::shared_ptr<SystemRunData> SystemRunData::getRelativeRun(int8_t offset) {
    if ((int)id + offset <= 0) throw std::invalid_argument("Offset " + std::to_string(offset) + " out of range");
    return system->getRunById(id + offset);
}

System::System(std::shared_ptr<DataSource> source, size_t workerThreads, size_t runRetention_, size_t concurrentRunLimit_)
    : runRetention(runRetention_), concurrentRunLimit(concurrentRunLimit_), threadPool(workerThreads), dataSource(source) {}

System::~System() { threadPool.join(); }

void System::addModule(std::shared_ptr<SystemModule> module) {
    modules.push_back(module);
    CART_LOG_INFO("System", "Added module " + module->name);
    for (const auto& provides : module->getProvidedData()) dataProvidedBy[provides] = module;
}

void System::verifyDependencies() {
    if (verifiedDependencies) return;
    for (const auto& module : modules)
        for (const auto& dep : module->getRequiredData())
            if (dataProvidedBy.find(dep.name) == dataProvidedBy.end())
                throw std::invalid_argument("Module " + module->name + " requires data " + dep.name +
                                            " which is not provided by any module");
    // Post order = dependency order over the same-run dependencies (stable: configuration order among independent
    // modules).  The pool is FIFO and a task only ever waits for tasks posted before it (same run: earlier in this
    // order; earlier runs: posted earlier), so the oldest unfinished task is never blocked - no deadlock for any
    // worker count >= 1, which the reference gets from its non-blocking future continuations (cartslam.cpp:255-302).
    runOrder.clear();
    std::vector<bool> placed(modules.size(), false);
    for (size_t done = 0; done < modules.size();) {
        bool progress = false;
        for (size_t i = 0; i < modules.size(); ++i) {
            if (placed[i]) continue;
            bool ready = true;
            for (const auto& dep : modules[i]->getRequiredData()) {
                if (dep.runOffset != 0) continue;
                auto provider = dataProvidedBy[dep.name];
                if (provider == modules[i]) continue;
                for (size_t j = 0; j < modules.size(); ++j)
                    if (modules[j] == provider && !placed[j]) ready = false;
            }
            if (!ready) continue;
            placed[i] = true;
            runOrder.push_back(modules[i]);
            ++done;
            progress = true;
        }
        if (!progress) throw std::invalid_argument("The modules' same-run dependencies form a cycle");
    }
    verifiedDependencies = true;
}

// Groups the keys by run offset and waits for each group on the run that owns it (cartslam.cpp:96-167).
void System::waitForDependencies(const std::vector<module_dependency_t>& deps, std::shared_ptr<SystemRunData> data) {
    std::map<int, std::vector<std::string>, std::greater<int>> byOffset;
    for (const auto& d : deps) {
        if ((int)data->id + d.runOffset <= 0) continue;  // the run does not exist (first frames)
        byOffset[d.runOffset].push_back(d.name);
    }
    for (auto& group : byOffset) {
        std::shared_ptr<SystemRunData> run = group.first < 0 ? data->getRelativeRun((int8_t)group.first) : data;
        try {
            run->waitForData(group.second);
        } catch (const std::exception&) {
            std::throw_with_nested(std::runtime_error("Error waiting for dependencies for run ID " + std::to_string(data->id) +
                                                      " from run ID " + std::to_string(run->id)));
        }
    }
}

uint8_t System::getActiveRunCount() {
    for (size_t i = 0; i < runs.size(); ++i)
        if (!runs[i]->isComplete()) return (uint8_t)(runs.size() - i);
    return 0;
}

std::shared_ptr<SystemRunData> System::startNewRun(void* stream) {
    verifyDependencies();
    std::unique_lock<std::mutex> lock(runMutex);
    auto element = dataSource->getNext(stream);
    auto data = std::make_shared<SystemRunData>(++runId, this, element);
    runCondition.wait(lock, [this] { return getActiveRunCount() < concurrentRunLimit; });
    runs.push_back(data);
    if (runs.size() > runRetention) runs.erase(runs.begin());
    return data;
}

std::shared_ptr<SystemRunData> System::getRunById(uint32_t id) {
    std::lock_guard<std::mutex> lock(runMutex);
    if (id > runId) throw std::invalid_argument("Index " + std::to_string(id) + " out of range (too new)");
    if (runs.empty() || id < runs[0]->id) throw std::invalid_argument("Index " + std::to_string(id) + " out of range (too old)");
    return runs[id - runs[0]->id];
}

std::future<void> System::run() {
    if (modules.empty()) throw std::invalid_argument("No modules have been added to the system");
    std::shared_ptr<SystemRunData> runData = startNewRun(nullptr);
    struct Pending {
        std::promise<void> done;
        std::atomic<size_t> left;
        std::mutex m;
        std::exception_ptr error;
    };
    auto pending = std::make_shared<Pending>();
    pending->left = modules.size();
    auto future = pending->done.get_future();
    for (auto module : runOrder) {
        // one task per module and frame: wait for its inputs, run it, publish its outputs (cartslam.cpp:255-302)
        threadPool.post([this, module, runData, pending] {
            try {
                waitForDependencies(module->getRequiredData(), runData);
                system_data_t out = module->run(*this, *runData).get();
                for (auto& entry : out) runData->insertData(entry);
            } catch (...) {
                try {
                    std::throw_with_nested(std::runtime_error("Error running module \"" + module->name + "\" for run ID " +
                                                              std::to_string(runData->id)));
                } catch (...) {
                    std::lock_guard<std::mutex> lock(pending->m);
                    if (!pending->error) pending->error = std::current_exception();
                }
            }
            if (--pending->left == 0) {
                runData->markAsComplete();
                {
                    std::lock_guard<std::mutex> lock(runMutex);
                }
                runCondition.notify_all();
                if (pending->error)
                    pending->done.set_exception(pending->error);
                else
                    pending->done.set_value();
            }
        });
    }
    return future;
}

}  // namespace cart
