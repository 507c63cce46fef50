// C ABI of libcartb200 (include/cartb200.h): context management, argument validation, stage entry
// points and the whole-sequence runner that reproduces the reference's per-sequence state.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <climits>
#include <new>
#include <thread>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "tile_ref.cuh"

namespace cb {
uint32_t debug_median9_host(const uint16_t* v9);
int launch_right_u16(cartb200_ctx* c, int n, uint16_t* out, cudaStream_t s);
int histogram_peak_update(const int32_t* hist, int32_t* params);  // plane_params.cpp
}  // namespace cb

using namespace cb;

namespace {

struct SeqScratch {
    int nFrames = 0, pipeline = -1;
    int phase1Frames = 0, phase1Pipeline = -1;  // what a completed phase 1 left in deriv / labels
    uint8_t* inL = nullptr;     // staged inputs (host variant) [n][H][W*3]
    uint8_t* inR = nullptr;
    int16_t* disp = nullptr;    // [B] or [n] frames, tight pitch
    int16_t* deriv = nullptr;   // pipeline 0: [B] s16; pipeline 1: [n] s16x2
    uint16_t* labels = nullptr; // pipeline 1: [n]
    uint8_t* planes = nullptr;  // [n]
    uint8_t* unsm = nullptr;    // [B]
    int32_t* hist = nullptr;    // [n][512]
    int32_t* histHost = nullptr;  // pinned
    int32_t* paramsHost = nullptr;  // pinned [n][4]
    cudaStream_t copyStream = nullptr;
    cudaStream_t spStream = nullptr;       // superpixel relaxation runs beside the SGM stages of the next batch
    static constexpr int kSpGroups = 8;    // the lock-stepped chunks advance as up to kSpGroups groups on streams of their own
    cudaStream_t spStreamX[kSpGroups - 1] = {};  // groups 1.. (group 0 runs on spStream)
    cudaEvent_t evSpFork = nullptr, evSpJoin[kSpGroups - 1] = {};
    std::vector<cudaEvent_t> evBatch;      // batch k: disparity + derivative done (recorded on the main stream)
    cudaEvent_t evStart = nullptr, evSpDone = nullptr;
    cudaEvent_t evIn[2] = {nullptr, nullptr};
    size_t inCap = 0, inRCap = 0, dispCap = 0, derivCap = 0, labelsCap = 0, planesCap = 0, histCap = 0, unsmCap = 0;
};

template <typename T>
int devAlloc(cartb200_ctx* c, T** p, size_t bytes) {
    if (cudaMalloc((void**)p, bytes) != cudaSuccess) {
        cudaGetLastError();
        c->err = "cudaMalloc of " + std::to_string(bytes) + " bytes failed";
        return CARTB200_E_NOMEM;
    }
    c->scratchBytes += bytes;
    return CARTB200_OK;
}

template <typename T>
int ensureCap(cartb200_ctx* c, T** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return CARTB200_OK;
    if (*p) {
        cudaFree(*p);
        c->scratchBytes -= *cap;
        *p = nullptr;
        *cap = 0;
    }
    int rc = devAlloc(c, p, bytes);
    if (rc == CARTB200_OK) *cap = bytes;
    return rc;
}

}  // namespace

namespace cb {
// WTA + medians / L-R check / range correction + (smoothing_radius > 0) the interpolation pass
int wta_post_batch(cartb200_ctx* c, int n, ImgBatch<int16_t> out, cudaStream_t s) {
    int rc;
    if ((rc = launch_wta(c, n, s))) return rc;
    if (c->cfg.smoothing_radius <= 0) return launch_sgm_post(c, n, out, s);
    // the raw disparity goes to the staging image; the smoothing pass reads it out of place
    ImgBatch<int16_t> tmp{(int16_t*)c->medL, c->dispPitch, c->dispPitch * (size_t)c->H};
    if ((rc = launch_sgm_post(c, n, tmp, s))) return rc;
    // ImageDisparityModule passes minDisparity*16 and the image width (disparity.hpp:27-28, disparity.cu:74)
    return launch_interpolate_from(c, n, ImgBatch<const int16_t>{tmp.data, tmp.pitch, tmp.frameStride}, out,
                                   c->cfg.smoothing_radius, c->cfg.smoothing_iterations, c->cfg.min_disparity * 16, c->W, s);
}

// The SGM scratch of a context is indexed by frame; a FrameWindow makes the stage launchers work on frames [f0, ...)
// by advancing the base pointers for its lifetime (launches capture the pointers by value).
struct FrameWindow {
    cartb200_ctx* c;
    uint8_t *grayL, *grayR, *volumes;
    uint32_t *censusL, *censusR, *wtaR;
    uint16_t *wtaL, *medL, *medR;
    FrameWindow(cartb200_ctx* ctx, int f0)
        : c(ctx), grayL(ctx->grayL), grayR(ctx->grayR), volumes(ctx->volumes), censusL(ctx->censusL), censusR(ctx->censusR),
          wtaR(ctx->wtaR), wtaL(ctx->wtaL), medL(ctx->medL), medR(ctx->medR) {
        const size_t f = (size_t)f0, H = (size_t)c->H;
        c->grayL += f * H * c->grayPitch;
        if (c->grayR) c->grayR += f * H * c->grayPitch;
        c->volumes += f * c->volFrameStride;
        c->censusL += f * H * c->cenRowWords;
        c->censusR += f * H * c->cenRowWords;
        c->wtaR += f * H * c->rkPitch;
        c->wtaL += f * H * (c->dispPitch / 2);
        c->medL += f * H * (c->dispPitch / 2);
        if (c->medR) c->medR += f * H * (c->dispPitch / 2);
    }
    ~FrameWindow() {
        c->grayL = grayL, c->grayR = grayR, c->volumes = volumes, c->censusL = censusL, c->censusR = censusR;
        c->wtaR = wtaR, c->wtaL = wtaL, c->medL = medL, c->medR = medR;
    }
};

// Frames per slice of a batch (0 = unsliced).  Aggregation is bound by the XU / ALU pipes and uses 55 % of the HBM
// bandwidth, the winner-takes-all pass is bound by HBM reads: with the batch cut into slices the WTA pass of slice i
// shares the SMs with the aggregation of slice i + 1.  CARTB200_SGM_SLICE (tuning aid) overrides the default.
static int sgmSliceFrames() {
    const char* e = getenv("CARTB200_SGM_SLICE");  // read per batch: a sweep can change it inside one process
    return e ? atoi(e) : 0;
}

int disparity_batch(cartb200_ctx* c, int n, ImgBatch<const uint8_t> left, ImgBatch<const uint8_t> right, ImgBatch<int16_t> disp,
                    cudaStream_t s) {
    int rc;
    if ((rc = launch_gray_census(c, n, left, right, s))) return rc;
    // a slice is a whole number of inner groups of a two-level batch, so that it is a batch of the same kind
    const int unit = disp.inner > 0 ? disp.inner : 1;
    int slice = sgmSliceFrames();
    slice = slice > 0 ? std::max(unit, slice / unit * unit) : 0;
    if (slice <= 0 || slice >= n) {
        if ((rc = launch_aggregate(c, n, s))) return rc;
        return wta_post_batch(c, n, disp, s);
    }
    if (!c->wtaStream) {
        int lo = 0, hi = 0;
        CB_CHECK_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        // high priority: the CTAs of a WTA pass take the slots that free up first, so that slice i is finished while the
        // aggregation of slice i + 1 is still running
        CB_CHECK_CUDA(c, cudaStreamCreateWithPriority(&c->wtaStream, cudaStreamNonBlocking, hi));
        CB_CHECK_CUDA(c, cudaEventCreateWithFlags(&c->evSliceAgg, cudaEventDisableTiming));
        CB_CHECK_CUDA(c, cudaEventCreateWithFlags(&c->evSliceWta, cudaEventDisableTiming));
    }
    auto slices = [&]() -> int {
        for (int f0 = 0; f0 < n; f0 += slice) {
            const int m = std::min(slice, n - f0);
            FrameWindow w(c, f0);
            if ((rc = launch_aggregate(c, m, s))) return rc;
            CB_CHECK_CUDA(c, cudaEventRecord(c->evSliceAgg, s));
            CB_CHECK_CUDA(c, cudaStreamWaitEvent(c->wtaStream, c->evSliceAgg, 0));
            ImgBatch<int16_t> d = disp;
            d.data = (int16_t*)((char*)disp.data + (disp.inner > 0 ? (size_t)(f0 / disp.inner) * disp.outerStride
                                                                   : (size_t)f0 * disp.frameStride));
            if ((rc = wta_post_batch(c, m, d, c->wtaStream))) return rc;
        }
        return CARTB200_OK;
    };
    rc = slices();
    // join the second stream on every path out of here: nothing may still write `disp` after the caller's stream is done
    if (cudaEventRecord(c->evSliceWta, c->wtaStream) != cudaSuccess || cudaStreamWaitEvent(s, c->evSliceWta, 0) != cudaSuccess) {
        if (rc == CARTB200_OK) {
            c->err = "disparity_batch: joining the WTA stream failed";
            rc = CARTB200_E_CUDA;
        }
    }
    return rc;
}
}  // namespace cb

namespace {

int* slotIota(cartb200_ctx* c) { return reinterpret_cast<int*>(c->paramsDev + 4 * (size_t)c->B); }
int* slotScratch(cartb200_ctx* c) { return slotIota(c) + c->B; }

int checkSgm(cartb200_ctx* c) {
    if (!c->volumes) {
        c->err = "context created with enable_sgm = 0";
        return CARTB200_E_ARG;
    }
    return CARTB200_OK;
}

// every buffer and kernel attribute of a context belongs to the device it was created on
int checkDevice(cartb200_ctx* c) {
    if (!c) return CARTB200_E_ARG;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev != c->device) {
        c->err = "the context was created on CUDA device " + std::to_string(c->device) + " but device " + std::to_string(dev) +
                 " is current: call cudaSetDevice first (one context per device)";
        return CARTB200_E_ARG;
    }
    return CARTB200_OK;
}

int checkBatch(cartb200_ctx* c, int n) {
    if (!c) return CARTB200_E_ARG;
    if (int rc = checkDevice(c)) return rc;
    if (n < 1 || n > c->B) {
        c->err = "batch size " + std::to_string(n) + " outside 1.." + std::to_string(c->B);
        return CARTB200_E_ARG;
    }
    return CARTB200_OK;
}

}  // namespace

namespace {
thread_local std::string g_createError;
}

extern "C" {

const char* cartb200_version(void) { return "cartb200 0.1 (sm_100a)"; }
const char* cartb200_last_create_error(void) { return g_createError.c_str(); }

void cartb200_default_config(cartb200_config* cfg, int width, int height) {
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->width = width;
    cfg->height = height;
    cfg->max_batch = 1;
    cfg->enable_sgm = 1;
    cfg->min_disparity = 4;       // cartconfig.cpp:147
    cfg->num_disparities = 256;   // cartconfig.cpp:148
    cfg->p1 = 10;                 // cv::cuda::createStereoSGM defaults
    cfg->p2 = 120;
    cfg->uniqueness_ratio = 12;   // disparity.hpp:32
    cfg->paths = 4;               // MODE_HH4
    cfg->smoothing_radius = -1;   // cartconfig.cpp:150
    cfg->smoothing_iterations = 5;
    cfg->enable_superpixels = 1;
    cfg->sp_block_size = 12;                 // cartconfig.cpp:126
    cfg->sp_direct_clique_cost = 0.5;        // :128
    cfg->sp_diagonal_clique_cost = 0.5 / 1.4142135623730951;  // :129
    cfg->sp_compactness_weight = 0.1;        // :130
    cfg->sp_progressive_compactness_cost = 0.0;
    cfg->sp_image_weight = 1.5;              // :132
    cfg->sp_disparity_weight = 1.0;          // :133
    cfg->sp_exact = 1;
}

void cartb200_default_sequence_opts(cartb200_sequence_opts* o) {
    std::memset(o, 0, sizeof(*o));
    o->pipeline = 0;
    o->provider = 1;
    o->update_interval = 30;  // cartconfig.cpp:202
    o->reset_interval = 10;   // :203
    o->sp_initial_iterations = 18;
    o->sp_iterations = 6;
    o->sp_reset_iterations = 64;
    o->start_id = 1;
}

int cartb200_create(const cartb200_config* cfg, cartb200_ctx** out) {
    if (!cfg || !out) return CARTB200_E_ARG;
    *out = nullptr;
    cartb200_ctx* c = new (std::nothrow) cartb200_ctx();
    if (!c) return CARTB200_E_NOMEM;
    c->cfg = *cfg;
    auto fail = [&](int rc) {
        // keep the message reachable: the caller cannot read it from a destroyed context, so print it
        fprintf(stderr, "cartb200_create: %s\n", c->err.c_str());
        g_createError = c->err;
        cartb200_destroy(c);
        return rc;
    };
    if (cfg->width < 16 || cfg->height < 8 || cfg->max_batch < 1) {
        c->err = "invalid size / batch";
        return fail(CARTB200_E_ARG);
    }
    if (cfg->width > 8192 || cfg->height > 8192) {
        // coordinates travel in 13 + 13 bits and 32 warp-merged x^2 / y^2 terms are summed in 32-bit integers
        c->err = "images larger than 8192 x 8192 are not supported";
        return fail(CARTB200_E_UNSUPPORTED);
    }
    if (cfg->num_disparities != 64 && cfg->num_disparities != 128 && cfg->num_disparities != 256) {
        c->err = "num_disparities must be 64, 128 or 256 (cv::cuda::StereoSGM restriction)";
        return fail(CARTB200_E_UNSUPPORTED);
    }
    if (cfg->paths != 4 && cfg->paths != 8) {
        c->err = "paths must be 4 (MODE_HH4) or 8 (MODE_HH)";
        return fail(CARTB200_E_UNSUPPORTED);
    }
    if (cfg->min_disparity < 0 || cfg->min_disparity > 1024) {
        c->err = "min_disparity must be in 0..1024";
        return fail(CARTB200_E_UNSUPPORTED);
    }
    if (cfg->p1 < 0 || cfg->p2 < cfg->p1 || 31 + cfg->p2 > 255) {
        c->err = "need 0 <= p1 <= p2 and 31 + p2 <= 255 (u8 path volumes)";
        return fail(CARTB200_E_UNSUPPORTED);
    }
    int dev = 0, ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        c->err = "no CUDA device: the cartb200 path has no CPU fallback";
        return fail(CARTB200_E_CUDA);
    }
    cudaGetDevice(&dev);
    c->device = dev;
    if (sgm_set_kernel_attributes() != cudaSuccess || post_set_kernel_attributes() != cudaSuccess ||
        sp_set_kernel_attributes() != cudaSuccess) {
        c->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(cudaGetLastError());
        return fail(CARTB200_E_CUDA);
    }
    c->W = cfg->width;
    c->H = cfg->height;
    c->D = cfg->num_disparities;
    c->B = cfg->max_batch;
    c->P = cfg->paths;
    const size_t W = c->W, H = c->H, B = c->B;
    int rc;
    c->grayPitch = alignUp(W, 128);
    // census rows: D + 32 zero words in front (right-census reads reach D + 16 words left of column 0),
    // min_disparity + 32 behind (16-word chunk loads run past the last column; the right census is stored
    // shifted by min_disparity)
    c->cenMargin = c->D + 32;
    c->cenRowWords = alignUp((size_t)c->cenMargin + W + (size_t)std::max(0, cfg->min_disparity) + 32, 32);
    c->censusPitch = c->cenRowWords * 4;
    c->dispPitch = alignUp(W * 2, 128);
    c->volFrameStride = H * W * (size_t)c->D;
    c->volPathStride = c->volFrameStride * B;
    if (cfg->enable_sgm) {
        if ((rc = devAlloc(c, &c->grayL, B * H * c->grayPitch))) return fail(rc);
        if ((rc = devAlloc(c, &c->censusL, B * H * c->censusPitch))) return fail(rc);
        if ((rc = devAlloc(c, &c->censusR, B * H * c->censusPitch))) return fail(rc);
        if (cudaMemset(c->censusL, 0, B * H * c->censusPitch) != cudaSuccess || cudaMemset(c->censusR, 0, B * H * c->censusPitch) != cudaSuccess) {
            c->err = "cudaMemset failed";
            return fail(CARTB200_E_CUDA);
        }
        // + 64 KiB: the WTA kernel reads (never uses) up to a few pixels past the end of a row segment
        if ((rc = devAlloc(c, &c->volumes, c->volPathStride * c->P + (64 << 10)))) return fail(rc);
        if ((rc = devAlloc(c, &c->wtaL, B * H * c->dispPitch))) return fail(rc);
        c->rkPitch = alignUp(W, 32);
        if ((rc = devAlloc(c, &c->wtaR, B * H * c->rkPitch * sizeof(uint32_t)))) return fail(rc);
    }
    if ((rc = devAlloc(c, &c->medL, B * H * c->dispPitch))) return fail(rc);  // also the interpolation staging image
    // paramsDev [B][4] + slot iota [B] + slot scratch [B]
    if ((rc = devAlloc(c, &c->paramsDev, (4 * B + 2 * B) * sizeof(int32_t)))) return fail(rc);
    {
        std::vector<int> iota(B);
        for (size_t i = 0; i < B; ++i) iota[i] = (int)i;
        if (cudaMemcpy(slotIota(c), iota.data(), B * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
            c->err = "cudaMemcpy failed";
            return fail(CARTB200_E_CUDA);
        }
    }
    if (cfg->enable_superpixels) {
        if (cfg->sp_block_size < 1) {
            c->err = "blockSize must be more than 1";  // superpixels.cu:36-38
            return fail(CARTB200_E_ARG);
        }
        if (cfg->sp_direct_clique_cost < 0 || cfg->sp_compactness_weight < 0 || cfg->sp_image_weight < 0 ||
            cfg->sp_disparity_weight < 0) {
            c->err = "clique cost / weights must be non-negative";  // superpixels.cu:40-46
            return fail(CARTB200_E_ARG);
        }
        c->spBlocksPerRow = ceilDiv(c->W, cfg->sp_block_size);
        c->maxLabels = c->spBlocksPerRow * ceilDiv(c->H, cfg->sp_block_size);
        if (c->maxLabels >= (1 << 14)) {
            c->err = "superpixel count must stay below 16384 (OUT_OF_BOUNDS marker, contourrelaxation.cu:21)";
            return fail(CARTB200_E_UNSUPPORTED);
        }
        if ((rc = devAlloc(c, &c->votes, B * (size_t)c->maxLabels * 4 * sizeof(uint32_t)))) return fail(rc);
    }
    if (cfg->enable_superpixels == 1) {  // 2 = only the vote table of the superpixel plane segmentation
        c->spLabelPitch = alignUp(W * 2, 128);
        if ((rc = devAlloc(c, &c->spLabels, B * 2 * H * c->spLabelPitch))) return fail(rc);  // two planes per slot
        if ((rc = devAlloc(c, &c->spYcc, B * H * W * 4))) return fail(rc);
        if ((rc = devAlloc(c, &c->spStats, B * (size_t)(c->maxLabels + 1) * 40 * sizeof(double)))) return fail(rc);
        {
            std::vector<int> tileMap;
            std::vector<uint32_t> tab;
            build_sp_tile_tables(c->W, c->H, tileMap, tab);
            if ((rc = devAlloc(c, &c->spTileMap, tileMap.size() * sizeof(int)))) return fail(rc);
            if ((rc = devAlloc(c, &c->spTileTab, std::max<size_t>(tab.size(), 1) * sizeof(uint32_t)))) return fail(rc);
            if (cudaMemcpy(c->spTileMap, tileMap.data(), tileMap.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
                (!tab.empty() && cudaMemcpy(c->spTileTab, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess)) {
                c->err = "cudaMemcpy failed";
                return fail(CARTB200_E_CUDA);
            }
        }
        if ((rc = launch_sp_reset(c, c->B, nullptr, nullptr))) return fail(rc);
        if (cudaDeviceSynchronize() != cudaSuccess) {
            c->err = "superpixel initialisation failed";
            return fail(CARTB200_E_CUDA);
        }
    }
    *out = c;
    return CARTB200_OK;
}

void cartb200_destroy(cartb200_ctx* c) {
    if (!c) return;
    for (int i = 0; i < 2; ++i) {
        if (c->aggStream[i]) cudaStreamDestroy(c->aggStream[i]);
        if (c->aggJoin[i]) cudaEventDestroy(c->aggJoin[i]);
    }
    if (c->aggFork) cudaEventDestroy(c->aggFork);
    if (c->wtaStream) cudaStreamDestroy(c->wtaStream);
    if (c->evSliceAgg) cudaEventDestroy(c->evSliceAgg);
    if (c->evSliceWta) cudaEventDestroy(c->evSliceWta);
    cudaFree(c->grayL);
    cudaFree(c->grayR);
    cudaFree(c->censusL);
    cudaFree(c->censusR);
    cudaFree(c->volumes);
    cudaFree(c->wtaL);
    cudaFree(c->wtaR);
    cudaFree(c->medL);
    cudaFree(c->medR);
    cudaFree(c->paramsDev);
    cudaFree(c->votes);
    cudaFree(c->spLabels);
    cudaFree(c->spYcc);
    cudaFree(c->spStats);
    cudaFree(c->spTileMap);
    cudaFree(c->spTileTab);
    if (c->seq) {
        SeqScratch* q = static_cast<SeqScratch*>(c->seq);
        cudaFree(q->inL);
        cudaFree(q->inR);
        cudaFree(q->disp);
        cudaFree(q->deriv);
        cudaFree(q->labels);
        cudaFree(q->planes);
        cudaFree(q->unsm);
        cudaFree(q->hist);
        cudaFreeHost(q->histHost);
        cudaFreeHost(q->paramsHost);
        if (q->copyStream) cudaStreamDestroy(q->copyStream);
        if (q->spStream) cudaStreamDestroy(q->spStream);
        for (auto& st : q->spStreamX)
            if (st) cudaStreamDestroy(st);
        if (q->evSpFork) cudaEventDestroy(q->evSpFork);
        for (auto& e : q->evSpJoin)
            if (e) cudaEventDestroy(e);
        for (auto& e : q->evBatch) cudaEventDestroy(e);
        if (q->evStart) cudaEventDestroy(q->evStart);
        if (q->evSpDone) cudaEventDestroy(q->evSpDone);
        for (auto& e : q->evIn)
            if (e) cudaEventDestroy(e);
        delete q;
    }
    cudaGetLastError();
    delete c;
}

const char* cartb200_last_error(const cartb200_ctx* c) { return c ? c->err.c_str() : "null context"; }
long long cartb200_launch_count(const cartb200_ctx* c) { return c ? c->launches : 0; }
size_t cartb200_scratch_bytes(const cartb200_ctx* c) { return c ? c->scratchBytes : 0; }

// ---- disparity ---------------------------------------------------------------------------------
int cartb200_sgm_gray_census(cartb200_ctx* c, int n, const uint8_t* l, const uint8_t* r, size_t pitch, size_t fstride,
                             void* stream) {
    int rc = checkBatch(c, n);
    if (rc) return rc;
    if ((rc = checkSgm(c))) return rc;
    if (!l || !r || pitch < (size_t)c->W * 3) {
        c->err = "gray_census: bad image arguments";
        return CARTB200_E_ARG;
    }
    return launch_gray_census(c, n, ImgBatch<const uint8_t>{l, pitch, fstride}, ImgBatch<const uint8_t>{r, pitch, fstride},
                              (cudaStream_t)stream);
}

int cartb200_sgm_aggregate(cartb200_ctx* c, int n, void* stream) {
    int rc = checkBatch(c, n);
    if (rc) return rc;
    if ((rc = checkSgm(c))) return rc;
    return launch_aggregate(c, n, (cudaStream_t)stream);
}

int cartb200_sgm_aggregate_path(cartb200_ctx* c, int n, int path, void* stream) {
    int rc = checkBatch(c, n);
    if (rc) return rc;
    if ((rc = checkSgm(c))) return rc;
    if (path < 0 || path >= c->P) {
        c->err = "aggregate_path: path index out of range";
        return CARTB200_E_ARG;
    }
    return launch_aggregate_range(c, n, path, path + 1, (cudaStream_t)stream);
}

int cartb200_interpolate(cartb200_ctx* c, int n, int16_t* d, size_t pitch, size_t fstride, int radius, int iterations,
                         int minD, int maxD, void* stream) {
    int rc = checkBatch(c, n);
    if (rc) return rc;
    if (!d || pitch < (size_t)c->W * 2) {
        c->err = "interpolate: bad image arguments";
        return CARTB200_E_ARG;
    }
    return launch_interpolate(c, n, ImgBatch<int16_t>{d, pitch, fstride}, radius, iterations, minD, maxD,
                              (cudaStream_t)stream);
}

int cartb200_sgm_wta_post(cartb200_ctx* c, int n, int16_t* d, size_t pitch, size_t fstride, void* stream) {
    int rc = checkBatch(c, n);
    if (rc) return rc;
    if ((rc = checkSgm(c))) return rc;
    if (!d || pitch < (size_t)c->W * 2) {
        c->err = "wta_post: bad image arguments";
        return CARTB200_E_ARG;
    }
    return cb::wta_post_batch(c, n, ImgBatch<int16_t>{d, pitch, fstride}, (cudaStream_t)stream);
}

int cartb200_disparity(cartb200_ctx* c, int n, const uint8_t* l, const uint8_t* r, size_t pitch, size_t fstride,
                       int16_t* d, size_t dpitch, size_t dfstride, void* stream) {
    int rc = checkBatch(c, n);
    if (rc) return rc;
    if ((rc = checkSgm(c))) return rc;
    if (!l || !r || pitch < (size_t)c->W * 3) {
        c->err = "gray_census: bad image arguments";
        return CARTB200_E_ARG;
    }
    if (!d || dpitch < (size_t)c->W * 2) {
        c->err = "wta_post: bad image arguments";
        return CARTB200_E_ARG;
    }
    return cb::disparity_batch(c, n, ImgBatch<const uint8_t>{l, pitch, fstride}, ImgBatch<const uint8_t>{r, pitch, fstride},
                               ImgBatch<int16_t>{d, dpitch, dfstride}, (cudaStream_t)stream);
}

int cartb200_sgm_intermediate(cartb200_ctx* c, int which, const void** ptr, size_t* pitch, size_t* fstride) {
    if (!c || !ptr || !pitch || !fstride) return CARTB200_E_ARG;
    const size_t H = c->H;
    switch (which) {
        case 0: *ptr = c->censusL + c->cenMargin; *pitch = c->censusPitch; *fstride = c->censusPitch * H; return CARTB200_OK;
        case 1: *ptr = c->censusR + c->cenMargin + c->cfg.min_disparity; *pitch = c->censusPitch; *fstride = c->censusPitch * H; return CARTB200_OK;
        case 2: *ptr = c->grayL; *pitch = c->grayPitch; *fstride = c->grayPitch * H; return CARTB200_OK;
        case 3: *ptr = c->wtaL; *pitch = c->dispPitch; *fstride = c->dispPitch * H; return CARTB200_OK;
        case 4: {  // the right image is kept as u32 keys; hand out a u16 copy (synchronous, debug only)
            if (!c->medR && devAlloc(c, &c->medR, (size_t)c->B * H * c->dispPitch)) return CARTB200_E_NOMEM;
            int rc = launch_right_u16(c, c->B, c->medR, nullptr);
            if (rc) return rc;
            CB_CHECK_CUDA(c, cudaDeviceSynchronize());
            *ptr = c->medR; *pitch = c->dispPitch; *fstride = c->dispPitch * H; return CARTB200_OK;
        }
        default:
            if (which >= 10 && which < 10 + c->P) {
                *ptr = c->volumes + (size_t)(which - 10) * c->volPathStride;
                *pitch = (size_t)c->W * c->D;
                *fstride = c->volFrameStride;
                return CARTB200_OK;
            }
    }
    c->err = "sgm_intermediate: unknown selector";
    return CARTB200_E_ARG;
}

// ---- derivative / planeseg ---------------------------------------------------------------------
int cartb200_derivative(cartb200_ctx* c, int n, const int16_t* d, size_t dp, size_t dfs, int16_t* o, size_t op,
                        size_t ofs, int32_t* hist, void* stream) {
    int rc = checkBatch(c, n);
    if (rc) return rc;
    if (!d || !o || !hist || dp < (size_t)c->W * 2 || op < (size_t)c->W * 4 || (op & 3)) {
        c->err = "derivative: bad image arguments (derivative pitch must be a multiple of 4)";
        return CARTB200_E_ARG;
    }
    return launch_derivative(c, n, ImgBatch<const int16_t>{d, dp, dfs}, ImgBatch<int16_t>{o, op, ofs}, hist,
                             (cudaStream_t)stream);
}

int cartb200_naive_derivative(cartb200_ctx* c, int n, const int16_t* d, size_t dp, size_t dfs, int16_t* o, size_t op,
                              size_t ofs, int32_t* hist, void* stream) {
    int rc = checkBatch(c, n);
    if (rc) return rc;
    if (!d || !o || !hist || dp < (size_t)c->W * 2 || op < (size_t)c->W * 2) {
        c->err = "naive_derivative: bad image arguments";
        return CARTB200_E_ARG;
    }
    return launch_naive_derivative(c, n, ImgBatch<const int16_t>{d, dp, dfs}, ImgBatch<int16_t>{o, op, ofs}, hist,
                                   (cudaStream_t)stream);
}

static int uploadParams(cartb200_ctx* c, int n, const int32_t* paramsHost, cudaStream_t s) {
    if (!paramsHost) {
        c->err = "plane parameters missing";
        return CARTB200_E_ARG;
    }
    CB_CHECK_CUDA(c, cudaMemcpyAsync(c->paramsDev, paramsHost, (size_t)n * 4 * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    return CARTB200_OK;
}

int cartb200_classify(cartb200_ctx* c, int n, const int16_t* d, size_t dp, size_t dfs, int channels, int channel,
                      const int32_t* paramsHost, uint8_t* planes, size_t pp, size_t pfs, void* stream) {
    int rc = checkBatch(c, n);
    if (rc) return rc;
    if (!d || !planes || channels < 1 || channel < 0 || channel >= channels || dp < (size_t)c->W * 2 * channels ||
        pp < (size_t)c->W) {
        c->err = "classify: bad arguments";
        return CARTB200_E_ARG;
    }
    if ((rc = uploadParams(c, n, paramsHost, (cudaStream_t)stream))) return rc;
    return launch_classify(c, n, ImgBatch<const int16_t>{d, dp, dfs}, channels, channel, c->paramsDev,
                           ImgBatch<uint8_t>{planes, pp, pfs}, (cudaStream_t)stream);
}

int cartb200_sp_planeseg(cartb200_ctx* c, int n, const int16_t* d, size_t dp, size_t dfs, const uint16_t* labels,
                         size_t lp, size_t lfs, int maxLabel, const int32_t* paramsHost, uint8_t* unsm, uint8_t* planes,
                         size_t pp, size_t pfs, void* stream) {
    int rc = checkBatch(c, n);
    if (rc) return rc;
    if (!c->votes) {
        c->err = "sp_planeseg: context created without superpixels";
        return CARTB200_E_ARG;
    }
    if (!d || !labels || !unsm || !planes || dp < (size_t)c->W * 4 || lp < (size_t)c->W * 2 || pp < (size_t)c->W ||
        maxLabel < 1 || maxLabel > c->maxLabels) {
        c->err = "sp_planeseg: bad arguments";
        return CARTB200_E_ARG;
    }
    if ((size_t)(maxLabel + 1) * 3 * sizeof(uint16_t) > 32768) {  // sp_planeseg.cu:327-331
        c->err = "Shared memory size exceeds maximum. Reduce image size or increase block size.";
        return CARTB200_E_UNSUPPORTED;
    }
    if ((rc = uploadParams(c, n, paramsHost, (cudaStream_t)stream))) return rc;
    return launch_sp_planeseg(c, n, ImgBatch<const int16_t>{d, dp, dfs}, ImgBatch<const uint16_t>{labels, lp, lfs},
                              maxLabel, c->paramsDev, ImgBatch<uint8_t>{unsm, pp, pfs}, ImgBatch<uint8_t>{planes, pp, pfs},
                              (cudaStream_t)stream);
}

static int makeTemporalRefs(cartb200_ctx* c, int count, const cartb200_temporal_ref* prev, cb::TemporalRefs& r) {
    if (count < 0 || count > CARTB200_MAX_TEMPORAL_DISTANCE || (count > 0 && !prev)) {
        c->err = "temporal vote: previous_count must be 0.." + std::to_string(CARTB200_MAX_TEMPORAL_DISTANCE);
        return CARTB200_E_ARG;
    }
    r = cb::TemporalRefs{};
    r.count = count;
    for (int k = 0; k < count; ++k) {
        if (!prev[k].planes_unsmoothed || !prev[k].optflow || prev[k].planes_pitch < (size_t)c->W ||
            prev[k].optflow_pitch < (size_t)c->W * 4 || (prev[k].optflow_pitch & 3) || ((uintptr_t)prev[k].optflow & 3)) {
            c->err = "temporal vote: bad previous frame " + std::to_string(k) + " (optflow is CV_16SC2 with a 4-byte aligned pitch)";
            return CARTB200_E_ARG;
        }
        r.planes[k] = prev[k].planes_unsmoothed;
        r.planesPitch[k] = prev[k].planes_pitch;
        r.flow[k] = prev[k].optflow;
        r.flowPitch[k] = prev[k].optflow_pitch;
    }
    return CARTB200_OK;
}

int cartb200_classify_temporal(cartb200_ctx* c, const int16_t* d, size_t dp, int channels, int channel, const int32_t* paramsHost,
                               int previousCount, const cartb200_temporal_ref* prev, uint8_t* unsm, uint8_t* smoothed, size_t pp,
                               void* stream) {
    if (!c) return CARTB200_E_ARG;
    if (!d || !unsm || !smoothed || !paramsHost || channels < 1 || channel < 0 || channel >= channels ||
        dp < (size_t)c->W * 2 * channels || pp < (size_t)c->W) {
        c->err = "classify_temporal: bad arguments";
        return CARTB200_E_ARG;
    }
    cb::TemporalRefs refs;
    int rc = makeTemporalRefs(c, previousCount, prev, refs);
    if (rc) return rc;
    const cb::PlaneRanges pr{paramsHost[0], paramsHost[1], paramsHost[2], paramsHost[3]};
    return launch_classify_temporal(c, Img<const int16_t>{d, dp}, channels, channel, pr, refs, Img<uint8_t>{unsm, pp},
                                    Img<uint8_t>{smoothed, pp}, (cudaStream_t)stream);
}

int cartb200_sp_planeseg_temporal(cartb200_ctx* c, const int16_t* d, size_t dp, const uint16_t* labels, size_t lp, int maxLabel,
                                  const int32_t* paramsHost, int previousCount, const cartb200_temporal_ref* prev, uint8_t* unsm,
                                  uint8_t* planes, size_t pp, void* stream) {
    if (!c) return CARTB200_E_ARG;
    if (!c->votes) {
        c->err = "sp_planeseg_temporal: context created without superpixels";
        return CARTB200_E_ARG;
    }
    if (!d || !labels || !unsm || !planes || !paramsHost || dp < (size_t)c->W * 4 || lp < (size_t)c->W * 2 || pp < (size_t)c->W ||
        maxLabel < 1 || maxLabel > c->maxLabels) {
        c->err = "sp_planeseg_temporal: bad arguments";
        return CARTB200_E_ARG;
    }
    if ((size_t)(maxLabel + 1) * 3 * sizeof(uint16_t) > 32768) {  // sp_planeseg.cu:327-331
        c->err = "Shared memory size exceeds maximum. Reduce image size or increase block size.";
        return CARTB200_E_UNSUPPORTED;
    }
    cb::TemporalRefs refs;
    int rc = makeTemporalRefs(c, previousCount, prev, refs);
    if (rc) return rc;
    const cb::PlaneRanges pr{paramsHost[0], paramsHost[1], paramsHost[2], paramsHost[3]};
    return launch_sp_planeseg_temporal(c, Img<const int16_t>{d, dp}, Img<const uint16_t>{labels, lp}, maxLabel, pr, refs,
                                       Img<uint8_t>{unsm, pp}, Img<uint8_t>{planes, pp}, (cudaStream_t)stream);
}

static int checkLabelDepth(cartb200_ctx* c, const char* what, const uint16_t* labels, size_t lp, const float* xyz, size_t xp,
                           int nLabels) {
    if (!c) return CARTB200_E_ARG;
    if (!labels || !xyz || lp < (size_t)c->W * 2 || xp < (size_t)c->W * 12 || (xp & 3) || nLabels < 1 || nLabels > 65536) {
        c->err = std::string(what) + ": bad arguments (depth is CV_32FC3 with a pitch that is a multiple of 4)";
        return CARTB200_E_ARG;
    }
    if ((size_t)nLabels * 4 > 32768) {  // planefit.cu:196-199 / :248-251 (the larger of the two shared tables)
        c->err = std::string(what) + ": Shared memory size exceeds maximum. Was " + std::to_string((size_t)nLabels * 4) +
                 " but maximum is 32768.";
        return CARTB200_E_UNSUPPORTED;
    }
    return CARTB200_OK;
}

int cartb200_label_statistics(cartb200_ctx* c, const uint16_t* labels, size_t lp, const float* xyz, size_t xp, int nLabels,
                              uint32_t* count, uint32_t* invalid, void* stream) {
    int rc = checkLabelDepth(c, "label_statistics", labels, lp, xyz, xp, nLabels);
    if (rc) return rc;
    if (!count || !invalid) {
        c->err = "label_statistics: null output";
        return CARTB200_E_ARG;
    }
    return launch_label_statistics(c, Img<const uint16_t>{labels, lp}, Img<const float>{xyz, xp}, nLabels, count, invalid,
                                   (cudaStream_t)stream);
}

int cartb200_region_inliers(cartb200_ctx* c, const uint16_t* labels, size_t lp, const float* xyz, size_t xp, int nLabels,
                            const double* planesHost, int nPlanes, double threshold, uint32_t* inliers, void* stream) {
    int rc = checkLabelDepth(c, "region_inliers", labels, lp, xyz, xp, nLabels);
    if (rc) return rc;
    if (!planesHost || nPlanes < 1 || !inliers) {
        c->err = "region_inliers: bad arguments";
        return CARTB200_E_ARG;
    }
    return launch_region_inliers(c, Img<const uint16_t>{labels, lp}, Img<const float>{xyz, xp}, nLabels, planesHost, nPlanes,
                                 threshold, inliers, (cudaStream_t)stream);
}

int cartb200_overlay_planes(cartb200_ctx* c, const uint8_t* bgr, size_t bp, const uint8_t* planes, size_t pp, uint8_t* out, size_t op,
                            void* stream) {
    if (!c) return CARTB200_E_ARG;
    if (!bgr || !planes || !out || bp < (size_t)c->W * 3 || pp < (size_t)c->W || op < (size_t)c->W * 3) {
        c->err = "overlay_planes: bad arguments";
        return CARTB200_E_ARG;
    }
    return launch_overlay_planes(c, Img<const uint8_t>{bgr, bp}, Img<const uint8_t>{planes, pp}, Img<uint8_t>{out, op}, (cudaStream_t)stream);
}

int cartb200_overlay_superpixel_boundaries(cartb200_ctx* c, const uint8_t* bgr, size_t bp, const uint16_t* labels, size_t lp, uint8_t* out,
                                           size_t op, void* stream) {
    if (!c) return CARTB200_E_ARG;
    if (!bgr || !labels || !out || bp < (size_t)c->W * 3 || lp < (size_t)c->W * 2 || op < (size_t)c->W * 3) {
        c->err = "overlay_superpixel_boundaries: bad arguments";
        return CARTB200_E_ARG;
    }
    return launch_overlay_boundaries(c, Img<const uint8_t>{bgr, bp}, Img<const uint16_t>{labels, lp}, Img<uint8_t>{out, op},
                                     (cudaStream_t)stream);
}

int cartb200_depth(cartb200_ctx* c, int n, const int16_t* d, size_t dp, size_t dfs, const float* q16Host, float* xyz, size_t xp,
                   size_t xfs, void* stream) {
    int rc = checkBatch(c, n);
    if (rc) return rc;
    if (!d || !xyz || !q16Host || dp < (size_t)c->W * 2 || xp < (size_t)c->W * 12 || (xp & 3)) {
        c->err = "depth: bad arguments (depth pitch must be a multiple of 4 and hold 3 floats per pixel)";
        return CARTB200_E_ARG;
    }
    return launch_depth(c, n, ImgBatch<const int16_t>{d, dp, dfs}, ImgBatch<float>{xyz, xp, xfs}, q16Host, (cudaStream_t)stream);
}

int cartb200_resize_bgr8(const uint8_t* src, size_t sp, int sw, int sh, uint8_t* dst, size_t dp, int dw, int dh, void* stream) {
    if (!src || !dst || sw < 1 || sh < 1 || dw < 1 || dh < 1 || sp < (size_t)sw * 3 || dp < (size_t)dw * 3) return CARTB200_E_ARG;
    return launch_resize_bgr8(src, sp, sw, sh, dst, dp, dw, dh, (cudaStream_t)stream);
}

int cartb200_histogram_peak_update(const int32_t* hist, int32_t* params) {
    if (!hist || !params) return CARTB200_E_ARG;
    return cb::histogram_peak_update(hist, params);
}

// ---- superpixels -------------------------------------------------------------------------------
static int resolveSlots(cartb200_ctx* c, int n, const int* slotsHost, const int** dev, cudaStream_t s) {
    if (!c->spLabels) {
        c->err = "context created without superpixels";
        return CARTB200_E_ARG;
    }
    if (!slotsHost) {
        *dev = slotIota(c);
        return CARTB200_OK;
    }
    for (int i = 0; i < n; ++i)
        if (slotsHost[i] < 0 || slotsHost[i] >= c->B) {
            c->err = "slot id out of range";
            return CARTB200_E_ARG;
        }
    CB_CHECK_CUDA(c, cudaMemcpyAsync(slotScratch(c), slotsHost, n * sizeof(int), cudaMemcpyHostToDevice, s));
    *dev = slotScratch(c);
    return CARTB200_OK;
}

int cartb200_superpixels_reset(cartb200_ctx* c, int n, const int* slotsHost, int* maxLabelOut, void* stream) {
    int rc = checkBatch(c, n);
    if (rc) return rc;
    const int* dev;
    if ((rc = resolveSlots(c, n, slotsHost, &dev, (cudaStream_t)stream))) return rc;
    if (maxLabelOut) *maxLabelOut = c->maxLabels;
    return launch_sp_reset(c, n, dev, (cudaStream_t)stream);
}

int cartb200_superpixels_relax(cartb200_ctx* c, int n, const int* slotsHost, int iterations, const uint8_t* bgr,
                               size_t bp, size_t bfs, const int16_t* deriv, size_t dp, size_t dfs, uint16_t* out,
                               size_t op, size_t ofs, void* stream) {
    int rc = checkBatch(c, n);
    if (rc) return rc;
    const int* dev;
    if ((rc = resolveSlots(c, n, slotsHost, &dev, (cudaStream_t)stream))) return rc;
    if (!bgr || bp < (size_t)c->W * 3 || iterations < 0 || (deriv && (dp < (size_t)c->W * 4)) ||
        (out && op < (size_t)c->W * 2)) {
        c->err = "superpixels_relax: bad arguments";
        return CARTB200_E_ARG;
    }
    return launch_sp_relax(c, n, dev, iterations, ImgBatch<const uint8_t>{bgr, bp, bfs},
                           ImgBatch<const int16_t>{deriv, dp, dfs}, deriv != nullptr, ImgBatch<uint16_t>{out, op, ofs},
                           (cudaStream_t)stream);
}

int cartb200_superpixels_set_labels(cartb200_ctx* c, int slot, const uint16_t* labels, size_t lp, void* stream) {
    if (!c || !c->spLabels || slot < 0 || slot >= c->B || !labels || lp < (size_t)c->W * 2) {
        if (c) c->err = "superpixels_set_labels: bad arguments";
        return CARTB200_E_ARG;
    }
    CB_CHECK_CUDA(c, cudaMemcpy2DAsync((char*)c->spLabels + (size_t)slot * 2 * c->spLabelPitch * c->H, c->spLabelPitch, labels,
                                       lp, (size_t)c->W * 2, c->H, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return CARTB200_OK;
}

int cartb200_superpixels_border_map(cartb200_ctx* c, const uint16_t* labels, size_t lp, uint8_t* border, size_t bp,
                                    void* stream) {
    if (!c || !labels || !border || lp < (size_t)c->W * 2 || bp < (size_t)c->W) {
        if (c) c->err = "border_map: bad arguments";
        return CARTB200_E_ARG;
    }
    return launch_border_map(c, Img<const uint16_t>{labels, lp}, Img<uint8_t>{border, bp}, (cudaStream_t)stream);
}

// ---- debug helpers (CPU, no GPU needed) --------------------------------------------------------
struct HostI32 {
    const int32_t* p;
    int W;
    int32_t operator()(int x, int y) const { return p[(size_t)y * W + x]; }
};

int32_t cartb200_debug_ref_tile_i32(const int32_t* img, int W, int H, int bx, int by, int bdx, int bdy, int XB, int YB,
                                    int yPad, int xPad, int interp, long alloc, int32_t undef, int lx, int ly) {
    TileGeom g{W, H, bdx * XB, bdy * YB, xPad, yPad, XB, YB, (int)alloc};
    HostI32 acc{img, W};
    TileEval<int32_t, HostI32> te(acc, g, bx, by, undef);
    return interp ? te.value<true>(lx, ly) : te.value<false>(lx, ly);
}

uint32_t cartb200_debug_median9(const uint16_t* v9) { return cb::debug_median9_host(v9); }

}  // extern "C"

// =================================================================================================
// Whole-sequence runner.  Reproduces, for frames processed in id order, the state the reference's
// modules carry across frames:
//   * SuperPixelModule: label image warm-started from the previous frame, re-blocked when
//     id % reset == 0, `initial` iterations on id == 1 or id % reset == 0 (superpixels.cu:93-113);
//   * DisparityPlaneSegmentationModule: running histogram, parameter update when id % update == 1,
//     zeroed after the download when id % (update*reset) == 1 (planeseg.cu:379-403);
//   * SuperPixelDisparityPlaneSegmentationModule: first call creates a zero running histogram without
//     adding the frame, then accumulate / maybe reset / update (sp_planeseg.cu:352-388).
// The superpixel chain is the only true frame-to-frame dependency, and it is cut at every reset, so the
// sequence is processed as independent chunks [k*reset, (k+1)*reset) advanced in lock step: step s
// handles frame k*reset + s of every chunk k in one batched launch.
namespace {

struct HistState {
    bool created = false;
    std::vector<long long> running = std::vector<long long>(256, 0);
    int32_t params[6] = {0, 0, 0, 0, 0, 0};  // hC, vC, hS, hE, vS, vE
};

// naive module bookkeeping for frame `id` whose own histogram is `h`; writes the ranges to use.  `peak(snapshot)` is what
// happens at an update frame: by default the provider's update on the running parameters
template <class Peak>
void naiveUpdateT(HistState& st, const cartb200_sequence_opts& o, int id, const int32_t* h, int32_t* outParams, Peak&& peak) {
    for (int i = 0; i < 256; ++i) st.running[i] += h[i];  // mergeHistogram (planeseg.cu:144-158)
    if (o.provider == 1 && id % o.update_interval == 1) {
        int32_t snap[256];
        for (int i = 0; i < 256; ++i) snap[i] = (int32_t)st.running[i];
        if (id % (o.update_interval * o.reset_interval) == 1) std::fill(st.running.begin(), st.running.end(), 0);
        peak(snap);
    }
    std::memcpy(outParams, st.params + 2, 4 * sizeof(int32_t));
}
void naiveUpdate(HistState& st, const cartb200_sequence_opts& o, int id, const int32_t* h, int32_t* outParams) {
    naiveUpdateT(st, o, id, h, outParams, [&](const int32_t* snap) { cb::histogram_peak_update(snap, st.params); });
}

// SP module bookkeeping; h = vertical channel of the frame's derivative histogram
template <class Peak>
void spUpdateT(HistState& st, const cartb200_sequence_opts& o, int id, const int32_t* hv, int32_t* outParams, Peak&& peak) {
    int32_t hist[256];
    if (!st.created) {
        st.created = true;  // zeros; the first frame is NOT added (sp_planeseg.cu:364-366)
        for (int i = 0; i < 256; ++i) hist[i] = hv[i];
    } else {
        for (int i = 0; i < 256; ++i) {
            st.running[i] += hv[i];
            hist[i] = (int32_t)st.running[i];
        }
    }
    if (id % (o.update_interval * o.reset_interval) == 1) std::fill(st.running.begin(), st.running.end(), 0);
    if (o.provider == 1 && id % o.update_interval == 1) peak(hist);
    std::memcpy(outParams, st.params + 2, 4 * sizeof(int32_t));
}
void spUpdate(HistState& st, const cartb200_sequence_opts& o, int id, const int32_t* hv, int32_t* outParams) {
    spUpdateT(st, o, id, hv, outParams, [&](const int32_t* hist) { cb::histogram_peak_update(hist, st.params); });
}

// phase 0: the whole pipeline.  phase 1: everything up to the per-frame histograms (copied to histOut, HOST, n x 256, the
// channel the pipeline's parameter provider consumes), derivative / label images stay in the context.  phase 2: plane
// parameters given per frame (paramsIn, HOST, n x 4) -> vote / classify on what phase 1 left behind.  Two phases are
// what lets one sequence be sharded over several GPUs with the histogram_peak provider: the
// running histogram crosses the shard boundaries, the per-frame histograms do not.
int runSequence(cartb200_ctx* c, const cartb200_sequence_opts* o, int n, const uint8_t* inL, const uint8_t* inR,
                bool inputsOnHost, uint8_t* planesOut, int16_t* dispOut, bool outputsOnHost, cudaStream_t userStream,
                int phase = 0, int32_t* histOut = nullptr, const int32_t* paramsIn = nullptr) {
    if (!c || !o || n < 1 || (phase != 2 && (!inL || !inR)) || (phase != 1 && !planesOut) || (phase == 1 && !histOut) ||
        (phase == 2 && !paramsIn)) {
        if (c) c->err = "run_sequence: bad arguments";
        return CARTB200_E_ARG;
    }
    if (int rcDev = checkDevice(c)) return rcDev;
    if (o->pipeline != 0 && o->pipeline != 1) {
        c->err = "run_sequence: unknown pipeline";
        return CARTB200_E_ARG;
    }
    if (o->update_interval < 1 || o->reset_interval < 1) {
        c->err = "run_sequence: intervals must be >= 1";
        return CARTB200_E_ARG;
    }
    const int R = o->sp_reset_iterations;
    if (o->pipeline == 1) {
        if (!c->spLabels) {
            c->err = "run_sequence: superpixel pipeline needs enable_superpixels";
            return CARTB200_E_ARG;
        }
        if (R < 1 || !(o->start_id == 1 || o->start_id % R == 0)) {
            c->err = "run_sequence: the superpixel chain can only start at id 1 or at a reset frame";
            return CARTB200_E_UNSUPPORTED;
        }
    }
    if (!c->seq) c->seq = new SeqScratch();
    SeqScratch* q = static_cast<SeqScratch*>(c->seq);
    const size_t W = c->W, H = c->H, B = c->B;
    const size_t bgrFrame = W * H * 3, pxFrame = W * H;
    int rc;
    cudaStream_t s = userStream;
    if (phase == 2 && (q->phase1Frames != n || q->phase1Pipeline != o->pipeline)) {
        c->err = "run_sequence phase 2: no matching phase 1 (same context, pipeline and frame count) precedes it";
        return CARTB200_E_ARG;
    }
    if (phase != 2) q->phase1Frames = 0;
    if (!q->copyStream) {
        CB_CHECK_CUDA(c, cudaStreamCreateWithFlags(&q->copyStream, cudaStreamNonBlocking));
        for (auto& e : q->evIn) CB_CHECK_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        {   // the latency-bound superpixel kernels get the higher priority: their blocks take free SM slots first,
            // the pipe-bound SGM kernels of the next batch fill the rest
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            if (getenv("CARTB200_SP_PRIORITY") && atoi(getenv("CARTB200_SP_PRIORITY")) == 0) hi = lo;  // tuning aid: no priority
            if (getenv("CARTB200_SP_PRIORITY") && atoi(getenv("CARTB200_SP_PRIORITY")) == 2) hi = (lo + hi) / 2;
            CB_CHECK_CUDA(c, cudaStreamCreateWithPriority(&q->spStream, cudaStreamNonBlocking, hi));
            for (auto& st : q->spStreamX) CB_CHECK_CUDA(c, cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, hi));
            CB_CHECK_CUDA(c, cudaEventCreateWithFlags(&q->evSpFork, cudaEventDisableTiming));
            for (auto& e : q->evSpJoin) CB_CHECK_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        CB_CHECK_CUDA(c, cudaEventCreateWithFlags(&q->evStart, cudaEventDisableTiming));
        CB_CHECK_CUDA(c, cudaEventCreateWithFlags(&q->evSpDone, cudaEventDisableTiming));
    }
    const bool peak = o->provider == 1;
    const bool sp = o->pipeline == 1;
    // Whatever path leaves this function early (an error code from a launch or a copy), work that was already queued on
    // the internal streams must not outlive the call: the caller may free or reuse its buffers as soon as it sees the
    // error.  On success the streams have been joined back into `s` (or synchronised) by the code below.
    struct JoinOnError {
        SeqScratch* q;
        bool ok = false;
        ~JoinOnError() {
            if (ok) return;
            if (q->spStream) cudaStreamSynchronize(q->spStream);
            for (auto& st : q->spStreamX)
                if (st) cudaStreamSynchronize(st);
            if (q->copyStream) cudaStreamSynchronize(q->copyStream);
        }
    } joinGuard{q};
    // storage
    if (inputsOnHost) {
        if ((rc = ensureCap(c, &q->inL, &q->inCap, bgrFrame * n))) return rc;
        if ((rc = ensureCap(c, &q->inR, &q->inRCap, bgrFrame * n))) return rc;
    }
    // superpixel pipeline: derivative and label images are kept per frame (the relaxation stream lags behind the
    // SGM stream, and the histogram-peak provider needs them again in the second phase)
    const size_t nStore = (sp || phase != 0) ? (size_t)n : B;
    if ((rc = ensureCap(c, &q->disp, &q->dispCap, pxFrame * 2 * (dispOut ? (size_t)n : B)))) return rc;
    if ((rc = ensureCap(c, &q->deriv, &q->derivCap, pxFrame * (sp ? 4 : 2) * nStore))) return rc;
    if (sp && (rc = ensureCap(c, &q->labels, &q->labelsCap, pxFrame * 2 * nStore))) return rc;
    if ((rc = ensureCap(c, &q->planes, &q->planesCap, pxFrame * (size_t)n))) return rc;
    if (sp && (rc = ensureCap(c, &q->unsm, &q->unsmCap, pxFrame * B))) return rc;
    if ((rc = ensureCap(c, &q->hist, &q->histCap, (size_t)n * 512 * sizeof(int32_t)))) return rc;
    if (q->nFrames < n) {
        cudaFreeHost(q->histHost);
        cudaFreeHost(q->paramsHost);
        q->histHost = q->paramsHost = nullptr;  // a failed allocation below must not leave stale pointers behind
        q->nFrames = 0;
        const size_t cap = std::max<size_t>((size_t)n, B);
        CB_CHECK_CUDA(c, cudaMallocHost((void**)&q->histHost, cap * 512 * sizeof(int32_t)));
        CB_CHECK_CUDA(c, cudaMallocHost((void**)&q->paramsHost, cap * 4 * sizeof(int32_t)));
        q->nFrames = n;
    }
    uint8_t* planesDev = outputsOnHost ? q->planes : planesOut;
    int16_t* dispDev = dispOut ? (outputsOnHost ? q->disp : dispOut) : q->disp;
    const uint8_t* devL = inputsOnHost ? q->inL : inL;
    const uint8_t* devR = inputsOnHost ? q->inR : inR;

    HistState hs;
    if (!peak) {
        hs.params[2] = o->static_params[0];
        hs.params[3] = o->static_params[1];
        hs.params[4] = o->static_params[2];
        hs.params[5] = o->static_params[3];
    }

    if (!sp) {
        // ---------------- naive pipeline: disparity -> naive derivative/hist -> params -> classify ----
        // inputs are uploaded batch by batch on the copy stream, one batch ahead of the compute
        auto upload = [&](int b0) -> int {
            if (!inputsOnHost || b0 >= n) return CARTB200_OK;
            const int nb = std::min<int>((int)B, n - b0);
            CB_CHECK_CUDA(c, cudaMemcpyAsync(q->inL + bgrFrame * b0, inL + bgrFrame * b0, bgrFrame * nb,
                                             cudaMemcpyHostToDevice, q->copyStream));
            CB_CHECK_CUDA(c, cudaMemcpyAsync(q->inR + bgrFrame * b0, inR + bgrFrame * b0, bgrFrame * nb,
                                             cudaMemcpyHostToDevice, q->copyStream));
            CB_CHECK_CUDA(c, cudaEventRecord(q->evIn[(b0 / B) & 1], q->copyStream));
            return CARTB200_OK;
        };
        if (phase != 2 && (rc = upload(0))) return rc;
        for (int b0 = 0; b0 < n; b0 += (int)B) {
            const int nb = std::min<int>((int)B, n - b0);
            int16_t* dB = dispOut ? dispDev + pxFrame * b0 : dispDev;
            // derivative images: one batch (whole pipeline) or all frames (phase 1 keeps them for phase 2)
            int16_t* vB = phase == 0 ? q->deriv : q->deriv + pxFrame * b0;
            if (phase != 2) {
                if (inputsOnHost) CB_CHECK_CUDA(c, cudaStreamWaitEvent(s, q->evIn[(b0 / B) & 1], 0));
                if ((rc = upload(b0 + (int)B))) return rc;
                if ((rc = cartb200_disparity(c, nb, devL + bgrFrame * b0, devR + bgrFrame * b0, W * 3, bgrFrame, dB, W * 2,
                                             pxFrame * 2, s)))
                    return rc;
                if ((rc = launch_naive_derivative(c, nb, ImgBatch<const int16_t>{dB, W * 2, pxFrame * 2},
                                                  ImgBatch<int16_t>{vB, W * 2, pxFrame * 2}, q->hist + 256 * (size_t)b0, s)))
                    return rc;
                if (outputsOnHost && dispOut)
                    CB_CHECK_CUDA(c, cudaMemcpyAsync(dispOut + pxFrame * b0, dB, pxFrame * 2 * nb, cudaMemcpyDeviceToHost, s));
            }
            if (phase == 1) continue;
            const int32_t* ph = paramsIn ? paramsIn + 4 * (size_t)b0 : q->paramsHost + 4 * (size_t)b0;
            if (phase == 0) {
                int32_t* pw = q->paramsHost + 4 * (size_t)b0;
                if (peak) {
                    CB_CHECK_CUDA(c, cudaMemcpyAsync(q->histHost + 256 * (size_t)b0, q->hist + 256 * (size_t)b0,
                                                     (size_t)nb * 256 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
                    CB_CHECK_CUDA(c, cudaStreamSynchronize(s));
                    for (int i = 0; i < nb; ++i)
                        naiveUpdate(hs, *o, o->start_id + b0 + i, q->histHost + 256 * (size_t)(b0 + i), pw + 4 * i);
                } else {
                    for (int i = 0; i < nb; ++i) std::memcpy(pw + 4 * i, hs.params + 2, 4 * sizeof(int32_t));
                }
            }
            if ((rc = uploadParams(c, nb, ph, s))) return rc;
            if ((rc = launch_classify(c, nb, ImgBatch<const int16_t>{vB, W * 2, pxFrame * 2}, 1, 0, c->paramsDev,
                                      ImgBatch<uint8_t>{planesDev + pxFrame * b0, W, pxFrame}, s)))
                return rc;
            if (outputsOnHost)
                CB_CHECK_CUDA(c, cudaMemcpyAsync(planesOut + pxFrame * b0, planesDev + pxFrame * b0, pxFrame * nb,
                                                 cudaMemcpyDeviceToHost, s));
        }
        if (phase == 1) {
            CB_CHECK_CUDA(c, cudaMemcpyAsync(histOut, q->hist, (size_t)n * 256 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            CB_CHECK_CUDA(c, cudaStreamSynchronize(s));
            q->phase1Frames = n;
            q->phase1Pipeline = 0;
            return joinGuard.ok = true, CARTB200_OK;
        }
        if (outputsOnHost || inputsOnHost) CB_CHECK_CUDA(c, cudaStreamSynchronize(s));
        return joinGuard.ok = true, CARTB200_OK;
    }

    // ---------------- superpixel pipeline ------------------------------------------------------------
    // virtual ids: chunk k covers ids [k*R, k*R + R); frame index (0-based) = id - start_id.
    const int firstId = o->start_id, lastId = o->start_id + n - 1;
    const int k0 = firstId / R, k1 = lastId / R;  // chunk range
    const int nChunks = k1 - k0 + 1;
    // vote + assign for all frames from the derivative / label images kept per frame, `params` = HOST n x 4
    bool planesCopied = false;
    auto votePhase = [&](const int32_t* params) -> int {
        for (int b0 = 0, bi = 0; b0 < n; b0 += (int)B, ++bi) {
            const int nb = std::min<int>((int)B, n - b0);
            if ((rc = uploadParams(c, nb, params + 4 * (size_t)b0, s))) return rc;
            if ((rc = launch_sp_planeseg(c, nb, ImgBatch<const int16_t>{q->deriv + pxFrame * 2 * b0, W * 4, pxFrame * 4},
                                         ImgBatch<const uint16_t>{q->labels + pxFrame * b0, W * 2, pxFrame * 2}, c->maxLabels,
                                         c->paramsDev, ImgBatch<uint8_t>{q->unsm, W, pxFrame},
                                         ImgBatch<uint8_t>{planesDev + pxFrame * b0, W, pxFrame}, s)))
                return rc;
            if (outputsOnHost) {  // download this batch while the next one is voted
                CB_CHECK_CUDA(c, cudaEventRecord(q->evIn[bi & 1], s));
                CB_CHECK_CUDA(c, cudaStreamWaitEvent(q->copyStream, q->evIn[bi & 1], 0));
                CB_CHECK_CUDA(c, cudaMemcpyAsync(planesOut + pxFrame * b0, planesDev + pxFrame * b0, pxFrame * nb,
                                                 cudaMemcpyDeviceToHost, q->copyStream));
            }
        }
        if (outputsOnHost) {
            CB_CHECK_CUDA(c, cudaEventRecord(q->evStart, q->copyStream));
            CB_CHECK_CUDA(c, cudaStreamWaitEvent(s, q->evStart, 0));
            planesCopied = true;
        }
        return CARTB200_OK;
    };
    if (phase == 2) {
        if ((rc = votePhase(paramsIn))) return rc;
        if (outputsOnHost) CB_CHECK_CUDA(c, cudaStreamSynchronize(s));
        return joinGuard.ok = true, CARTB200_OK;
    }
    const bool deferVote = peak || phase == 1;  // parameters are only known once every frame's histogram is
    if (!deferVote) {  // static ranges are the same for every frame: upload them once
        for (size_t j = 0; j < B; ++j) std::memcpy(q->paramsHost + 4 * j, hs.params + 2, 4 * sizeof(int32_t));
        if ((rc = uploadParams(c, (int)B, q->paramsHost, s))) return rc;
    }
    // Chunks are grouped so that every chunk of a group owns one superpixel slot.  The SGM/derivative stages
    // have no frame-to-frame state, so they run on up to B frames at once: `nb` chunks x `k` consecutive
    // steps (two-level ImgBatch); only the superpixel relaxation advances step by step.
    struct Batch {
        int ca, nb, st, k, idA;
        size_t fiA;  // frame index of (first chunk, first step)
    };
    std::vector<Batch> batches;
    const int maxSlots = (int)B;
    for (int g0 = 0; g0 < nChunks; g0 += maxSlots) {
        const int gN = std::min<int>(maxSlots, nChunks - g0);
        const int K = std::max<int>(1, (int)B / gN);
        auto range = [&](int st, int& ca, int& cb) {  // chunks of this group that own a frame at step st
            ca = cb = -1;
            for (int j = 0; j < gN; ++j) {
                const int id = (k0 + g0 + j) * R + st;
                if (id >= firstId && id <= lastId && id >= 1) {
                    if (ca < 0) ca = j;
                    cb = j + 1;
                }
            }
        };
        for (int st = 0; st < R;) {
            int ca, cb;
            range(st, ca, cb);
            if (ca < 0) {
                ++st;
                continue;
            }
            int k = 1;
            for (; st + k < R && k < K; ++k) {
                int ca2, cb2;
                range(st + k, ca2, cb2);
                if (ca2 != ca || cb2 != cb) break;
            }
            const int idA = (k0 + g0 + ca) * R + st;
            batches.push_back(Batch{ca, cb - ca, st, k, idA, (size_t)(idA - firstId)});
            st += k;
        }
    }
    const size_t fStride = (size_t)R;  // frames between consecutive chunks
    // inputs in host memory: batch b + 1 is uploaded on the copy stream (step-major: k consecutive frames of
    // every chunk) while batch b computes
    auto upload = [&](size_t bi) -> int {
        if (!inputsOnHost || bi >= batches.size()) return CARTB200_OK;
        const Batch& bt = batches[bi];
        for (int j = 0; j < bt.nb; ++j) {
            const size_t off = bgrFrame * (bt.fiA + j * fStride);
            CB_CHECK_CUDA(c, cudaMemcpyAsync(q->inL + off, inL + off, bgrFrame * bt.k, cudaMemcpyHostToDevice, q->copyStream));
            CB_CHECK_CUDA(c, cudaMemcpyAsync(q->inR + off, inR + off, bgrFrame * bt.k, cudaMemcpyHostToDevice, q->copyStream));
        }
        CB_CHECK_CUDA(c, cudaEventRecord(q->evIn[bi & 1], q->copyStream));
        return CARTB200_OK;
    };
    // Two streams: `s` runs gray/census/aggregation/WTA/derivative of batch b + 1 while `spS` relaxes the
    // superpixels of batch b.
    cudaStream_t spS = q->spStream;
    CB_CHECK_CUDA(c, cudaEventRecord(q->evStart, s));
    CB_CHECK_CUDA(c, cudaStreamWaitEvent(spS, q->evStart, 0));
    if (inputsOnHost) CB_CHECK_CUDA(c, cudaStreamWaitEvent(q->copyStream, q->evStart, 0));
    // a sequence that starts at id 1 starts from the constructor's block initialisation (superpixels.cu:56-58)
    if (firstId == 1 && (rc = launch_sp_reset(c, 1, slotIota(c), spS))) return rc;
    std::vector<size_t> histPos((size_t)n, 0);  // frame -> row of the batch-ordered histogram staging buffer
    size_t histRows = 0;
    if ((rc = upload(0))) return rc;
    for (size_t bi = 0; bi < batches.size(); ++bi) {
        const Batch& bt = batches[bi];
        const int nb = bt.nb, k = bt.k, total = nb * k;
        const size_t fiA = bt.fiA;
        const int* slots = slotIota(c) + bt.ca;
        if (inputsOnHost) CB_CHECK_CUDA(c, cudaStreamWaitEvent(s, q->evIn[bi & 1], 0));
        if ((rc = upload(bi + 1))) return rc;
        auto seqBatch = [&](auto* base, size_t frameBytes) {  // frames in sequence order, two-level
            using T = std::remove_pointer_t<decltype(base)>;
            return ImgBatch<T>{(T*)((char*)base + frameBytes * fiA), 0, frameBytes * fStride, nb, frameBytes};
        };
        ImgBatch<const uint8_t> bl = seqBatch(devL, bgrFrame), br = seqBatch(devR, bgrFrame);
        bl.pitch = br.pitch = W * 3;
        // disparity (all `total` frames)
        ImgBatch<int16_t> dB = dispOut ? seqBatch(dispDev, pxFrame * 2) : ImgBatch<int16_t>{dispDev, 0, pxFrame * 2};
        dB.pitch = W * 2;
        if ((rc = disparity_batch(c, total, bl, br, dB, s))) return rc;
        // derivative + per-frame histograms (batch-ordered staging rows, one download at the end)
        ImgBatch<int16_t> vB = seqBatch(q->deriv, pxFrame * 4);
        vB.pitch = W * 4;
        const ImgBatch<const int16_t> dBc{dB.data, dB.pitch, dB.frameStride, dB.inner, dB.outerStride};
        if ((rc = launch_derivative(c, total, dBc, vB, q->hist + 512 * histRows, s))) return rc;
        for (int kk = 0; kk < k; ++kk)
            for (int j = 0; j < nb; ++j) histPos[fiA + j * fStride + kk] = histRows + (size_t)(kk * nb + j);
        histRows += (size_t)total;
        ImgBatch<uint16_t> lB = seqBatch(q->labels, pxFrame * 2);
        lB.pitch = W * 2;
        if (q->evBatch.size() <= bi) {
            cudaEvent_t e;
            CB_CHECK_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            q->evBatch.push_back(e);
        }
        CB_CHECK_CUDA(c, cudaEventRecord(q->evBatch[bi], s));
        CB_CHECK_CUDA(c, cudaStreamWaitEvent(spS, q->evBatch[bi], 0));
        ImgBatch<uint8_t> pB = seqBatch(planesDev, pxFrame);
        pB.pitch = W;
        // superpixels, one step at a time: reset + iteration schedule (superpixels.cu:93-113).  The chunks of a batch are
        // independent chains: they are advanced as G groups on G streams, so that the partially filled last wave of one
        // group's relaxation launch (16 frames = 1920 tiles on 740 resident CTAs: 2.6 waves) is filled by the launches
        // of the other groups instead of idling.  CARTB200_SP_SPLIT = G (1 keeps one stream; measured: profiles/r02w_*).
        // (read per batch, not cached: the tests compare the group counts inside one process)
        const int spGroups = std::max(1, std::min<int>(SeqScratch::kSpGroups, getenv("CARTB200_SP_SPLIT") ? atoi(getenv("CARTB200_SP_SPLIT")) : 2));
        // (the one step that gives frame id 1 its own iteration count keeps the single stream)
        const bool hasIdOne = bt.idA <= 1 && 1 < bt.idA + k && !(bt.st == 0 && bt.idA == 1);
        const int G = hasIdOne ? 1 : std::max(1, std::min(spGroups, nb / 2));
        int gOff[SeqScratch::kSpGroups + 1];
        for (int g = 0; g <= G; ++g) gOff[g] = (int)((long)nb * g / G);
        auto gStream = [&](int g) { return g == 0 ? spS : q->spStreamX[g - 1]; };
        const bool split = G > 1;
        if (split) {
            CB_CHECK_CUDA(c, cudaEventRecord(q->evSpFork, spS));
            for (int g = 1; g < G; ++g) CB_CHECK_CUDA(c, cudaStreamWaitEvent(gStream(g), q->evSpFork, 0));
        }
        for (int kk = 0; kk < k; ++kk) {
            auto step = [&](auto batch) {  // the nb frames of step st + kk as a simple strided batch
                using BT = decltype(batch);
                return BT{(decltype(batch.data))((char*)batch.data + batch.outerStride * kk), batch.pitch, batch.frameStride};
            };
            const bool resetStep = bt.st + kk == 0;  // id % R == 0
            const ImgBatch<const uint8_t> sl = step(bl);
            const ImgBatch<int16_t> svm = step(vB);
            const ImgBatch<const int16_t> sv{svm.data, svm.pitch, svm.frameStride};
            const ImgBatch<uint16_t> so = step(lB);
            if (split) {
                const int its = resetStep ? o->sp_initial_iterations : o->sp_iterations;
                for (int g = 0; g < G; ++g) {
                    const int f0 = gOff[g], m = gOff[g + 1] - gOff[g];
                    if (resetStep && (rc = launch_sp_reset(c, m, slots + f0, gStream(g)))) return rc;
                    if ((rc = launch_sp_relax(c, m, slots + f0, its, sl.from(f0), sv.from(f0), true, so.from(f0), gStream(g), f0)))
                        return rc;
                }
                continue;
            }
            if (resetStep && (rc = launch_sp_reset(c, nb, slots, spS))) return rc;
            // id == 1 also gets the initial iteration count (only chunk 0 at step 1 can be id 1)
            if (!resetStep && bt.idA + kk == 1) {
                if ((rc = launch_sp_relax(c, 1, slots, o->sp_initial_iterations, sl, sv, true, so, spS))) return rc;
                if (nb > 1 && (rc = launch_sp_relax(c, nb - 1, slots + 1, o->sp_iterations, sl.from(1), sv.from(1), true,
                                                    so.from(1), spS, 1)))
                    return rc;
            } else {
                const int its = resetStep ? o->sp_initial_iterations : o->sp_iterations;
                if ((rc = launch_sp_relax(c, nb, slots, its, sl, sv, true, so, spS))) return rc;
            }
        }
        if (split) {
            for (int g = 1; g < G; ++g) {
                CB_CHECK_CUDA(c, cudaEventRecord(q->evSpJoin[g - 1], gStream(g)));
                CB_CHECK_CUDA(c, cudaStreamWaitEvent(spS, q->evSpJoin[g - 1], 0));
            }
        }
        if (!deferVote) {
            // static ranges: vote + assign right away (the same ranges for every frame of the batch)
            const ImgBatch<const int16_t> vBc{vB.data, vB.pitch, vB.frameStride, vB.inner, vB.outerStride};
            const ImgBatch<const uint16_t> lBc{lB.data, lB.pitch, lB.frameStride, lB.inner, lB.outerStride};
            if ((rc = launch_sp_planeseg(c, total, vBc, lBc, c->maxLabels, c->paramsDev,
                                         ImgBatch<uint8_t>{q->unsm, W, pxFrame}, pB, spS)))
                return rc;
        }
    }
    if (deferVote) CB_CHECK_CUDA(c, cudaMemcpyAsync(q->histHost, q->hist, histRows * 512 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CB_CHECK_CUDA(c, cudaEventRecord(q->evSpDone, spS));
    CB_CHECK_CUDA(c, cudaStreamWaitEvent(s, q->evSpDone, 0));
    if (deferVote) {
        // second phase: parameters in id order, then vote + assign for every frame
        CB_CHECK_CUDA(c, cudaStreamSynchronize(s));
        std::vector<int32_t> hv(256);
        for (int i = 0; i < n; ++i) {
            const int32_t* h = q->histHost + 512 * histPos[(size_t)i];
            for (int b = 0; b < 256; ++b) hv[b] = h[2 * b];  // channel 0 = vertical (sp_planeseg.cu:358-359)
            if (phase == 1)
                std::memcpy(histOut + 256 * (size_t)i, hv.data(), 256 * sizeof(int32_t));
            else
                spUpdate(hs, *o, firstId + i, hv.data(), q->paramsHost + 4 * (size_t)i);
        }
        if (phase == 1) {
            if (outputsOnHost && dispOut) {
                CB_CHECK_CUDA(c, cudaMemcpyAsync(dispOut, dispDev, pxFrame * 2 * n, cudaMemcpyDeviceToHost, s));
                CB_CHECK_CUDA(c, cudaStreamSynchronize(s));
            }
            q->phase1Frames = n;
            q->phase1Pipeline = 1;
            return joinGuard.ok = true, CARTB200_OK;
        }
        if ((rc = votePhase(q->paramsHost))) return rc;
    }
    if (outputsOnHost) {
        if (!planesCopied) CB_CHECK_CUDA(c, cudaMemcpyAsync(planesOut, planesDev, pxFrame * n, cudaMemcpyDeviceToHost, s));
        if (dispOut) CB_CHECK_CUDA(c, cudaMemcpyAsync(dispOut, dispDev, pxFrame * 2 * n, cudaMemcpyDeviceToHost, s));
    }
    if (outputsOnHost || inputsOnHost) CB_CHECK_CUDA(c, cudaStreamSynchronize(s));
    return joinGuard.ok = true, CARTB200_OK;
}

}  // namespace

extern "C" {

int cartb200_run_sequence_host(cartb200_ctx* c, const cartb200_sequence_opts* o, int n, const uint8_t* l,
                               const uint8_t* r, uint8_t* planes, int16_t* disp) {
    return runSequence(c, o, n, l, r, true, planes, disp, true, nullptr);
}

int cartb200_run_sequence_device(cartb200_ctx* c, const cartb200_sequence_opts* o, int n, const uint8_t* l,
                                 const uint8_t* r, uint8_t* planes, int16_t* disp, void* stream) {
    return runSequence(c, o, n, l, r, false, planes, disp, false, (cudaStream_t)stream);
}

int cartb200_run_sequence_phase1_device(cartb200_ctx* c, const cartb200_sequence_opts* o, int n, const uint8_t* l,
                                        const uint8_t* r, int32_t* hist_host, int16_t* disp, void* stream) {
    return runSequence(c, o, n, l, r, false, nullptr, disp, false, (cudaStream_t)stream, 1, hist_host, nullptr);
}

int cartb200_run_sequence_phase2_device(cartb200_ctx* c, const cartb200_sequence_opts* o, int n, const int32_t* params_host,
                                        uint8_t* planes, void* stream) {
    return runSequence(c, o, n, nullptr, nullptr, false, planes, nullptr, false, (cudaStream_t)stream, 2, nullptr, params_host);
}

int cartb200_run_sequence_phase1_host(cartb200_ctx* c, const cartb200_sequence_opts* o, int n, const uint8_t* l,
                                      const uint8_t* r, int32_t* hist_host, int16_t* disp) {
    return runSequence(c, o, n, l, r, true, nullptr, disp, true, nullptr, 1, hist_host, nullptr);
}

int cartb200_run_sequence_phase2_host(cartb200_ctx* c, const cartb200_sequence_opts* o, int n, const int32_t* params_host,
                                      uint8_t* planes) {
    return runSequence(c, o, n, nullptr, nullptr, false, planes, nullptr, true, nullptr, 2, nullptr, params_host);
}

int cartb200_sequence_parameters(const cartb200_sequence_opts* o, int n, const int32_t* hist, int32_t* params) {
    if (!o || n < 0 || (n > 0 && (!hist || !params)) || (o->pipeline != 0 && o->pipeline != 1) || o->update_interval < 1 ||
        o->reset_interval < 1)
        return CARTB200_E_ARG;
    HistState hs;
    if (o->provider != 1) {
        hs.params[2] = o->static_params[0];
        hs.params[3] = o->static_params[1];
        hs.params[4] = o->static_params[2];
        hs.params[5] = o->static_params[3];
    }
    // The running-histogram bookkeeping is sequential and cheap (0.14 us per frame); the peak detection at the update
    // frames (24 us each: two std::sort calls) is a pure function of the snapshot, which writes a subset of the six
    // parameters and reads none.  For long sequences - one 10,000-frame sequence sharded over N GPUs runs this table on
    // every rank between its two phases - the snapshots are taken first, the peak updates run on a few threads, and the
    // results are applied in frame order.  Same numbers as the frame-by-frame loop (tests/test_sharded_sequence.py).
    const bool threaded = o->provider == 1 && n >= 2048;
    if (!threaded) {
        for (int i = 0; i < n; ++i) {
            if (o->pipeline == 0)
                naiveUpdate(hs, *o, o->start_id + i, hist + 256 * (size_t)i, params + 4 * (size_t)i);
            else
                spUpdate(hs, *o, o->start_id + i, hist + 256 * (size_t)i, params + 4 * (size_t)i);
        }
        return CARTB200_OK;
    }
    std::vector<int32_t> snaps;           // [updates][256]
    std::vector<int> updateOf((size_t)n, -1);
    for (int i = 0; i < n; ++i) {
        auto take = [&](const int32_t* snap) {
            updateOf[(size_t)i] = (int)(snaps.size() / 256);
            snaps.insert(snaps.end(), snap, snap + 256);
        };
        if (o->pipeline == 0)
            naiveUpdateT(hs, *o, o->start_id + i, hist + 256 * (size_t)i, params + 4 * (size_t)i, take);
        else
            spUpdateT(hs, *o, o->start_id + i, hist + 256 * (size_t)i, params + 4 * (size_t)i, take);
    }
    const int nUpd = (int)(snaps.size() / 256);
    constexpr int32_t kUntouched = INT32_MIN;  // no parameter can take this value (bins are 0..255)
    std::vector<int32_t> cand((size_t)nUpd * 6, kUntouched);
    const int nThreads = std::max(1, std::min<int>({8, (int)std::thread::hardware_concurrency(), nUpd / 8}));
    std::atomic<int> next{0};
    auto work = [&]() {  // updates are handed out one at a time: any number of workers, this thread included, finishes the list
        for (int u; (u = next.fetch_add(1)) < nUpd;) cb::histogram_peak_update(snaps.data() + 256 * (size_t)u, cand.data() + 6 * (size_t)u);
    };
    std::vector<std::thread> pool;
    try {
        for (int t = 1; t < nThreads; ++t) pool.emplace_back(work);
    } catch (...) {  // no more threads to be had: no exception may cross the C ABI, the calling thread does the rest
    }
    work();
    for (auto& th : pool) th.join();
    int32_t cur[6];
    std::memcpy(cur, hs.params, sizeof(cur));  // still the start values: the snapshot pass never updated them
    for (int i = 0; i < n; ++i) {  // replay in frame order
        const int u = updateOf[(size_t)i];
        if (u >= 0)
            for (int f = 0; f < 6; ++f)
                if (cand[6 * (size_t)u + f] != kUntouched) cur[f] = cand[6 * (size_t)u + f];
        std::memcpy(params + 4 * (size_t)i, cur + 2, 4 * sizeof(int32_t));
    }
    return CARTB200_OK;
}

}  // extern "C"
