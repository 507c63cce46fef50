// Post-SGM stages: disparity smoothing, directional derivatives + histograms, naive low-pass
// derivative + histogram, plane classification and the superpixel vote.  Each kernel restates one
// reference kernel (cited below) on a B200-sized grid: frames are batched in blockIdx.z, tiles are
// smaller than the reference's 128x128 so a KITTI frame yields hundreds of CTAs instead of 30, the
// reference's tile-loader semantics come from tile_ref.cuh, histograms are privatised in shared memory
// and merged with one atomic per non-empty bin.
#include "common.cuh"
#include "tile_ref.cuh"

namespace cb {

struct DispAccessor {
    Img<const int16_t> im;
    __device__ __forceinline__ int16_t operator()(int x, int y) const { return __ldg(im.row(y) + x); }
};

// ---------------------------------------------------------------------------------------------
// interpolateKernel, /root/reference/src/modules/disparity/interpolation.cu:17-82.
// One CTA per reference 64x64 tile (the tile is the unit of the reference's semantics: halo values
// stay fixed over the iterations).  Canonical Jacobi schedule: two shared buffers.
// src and dst are different images (the reference updates in place while neighbouring blocks still
// read their halos - an inter-block race; canonical = every tile reads the original image).
__global__ void __launch_bounds__(256) interpolate_kernel(ImgBatch<const int16_t> src, ImgBatch<int16_t> dst, int W,
                                                          int H, int radius, int iterations, int minD, int maxD) {
    extern __shared__ int16_t sm[];
    __shared__ unsigned recipM[128];  // ceil(2^32 / count) for the window counts 1 .. (2 radius - 1)^2
    const int pad = radius - 1, S = 64 + 2 * pad, N = S * S;
    if (threadIdx.x >= 1 && threadIdx.x < 128) recipM[threadIdx.x] = 0xFFFFFFFFu / threadIdx.x + 1u;
    int16_t* cur = sm;
    int16_t* nxt = sm + N;
    const int bx = blockIdx.x, by = blockIdx.y, f = blockIdx.z;
    TileGeom g{W, H, 64, 64, pad, pad, 4, 4, N};
    DispAccessor acc{src.frame(f)};
    TileEval<int16_t, DispAccessor> te(acc, g, bx, by, kInvalid);
    // A tile whose padded extent lies inside the image gets no halo phase from the reference's loader: its shared tile is
    // the body copy alone, i.e. the image read pad rows further down (TileEval::body with pXS = startX - pad,
    // pYS = startY - pad reduces to image(startX + lx, startY + ly + pad)).  The general closed form (divisions by the
    // run-time tile stride, the four halo rules) is only evaluated for the tiles on the image border.
    const bool interior = bx * 64 - pad >= 0 && by * 64 - pad >= 0 && bx * 64 + 64 + pad <= W && by * 64 + 64 + pad <= H;
    if (interior) {
        const Img<const int16_t> im = src.frame(f);
        for (int r = threadIdx.x / 32; r < S; r += blockDim.x / 32) {  // one warp per tile row: coalesced, no division
            const int y = by * 64 + r;  // = startY + (r - pad) + pad
            for (int cidx = threadIdx.x % 32; cidx < S; cidx += 32) {
                const int x = bx * 64 + cidx - pad;
                const int16_t v = y < H ? __ldg(im.row(y) + x) : kInvalid;
                cur[r * S + cidx] = v;
                nxt[r * S + cidx] = v;
            }
        }
    } else {
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const int ly = i / S - pad, lx = i % S - pad;
            const int16_t v = te.template value<true>(lx, ly);
            cur[i] = v;
            nxt[i] = v;
        }
    }
    __syncthreads();
    const unsigned minCount = (unsigned)(radius * radius + 1);
    for (int it = 0; it < iterations; ++it) {
        if (radius == 2) {
            // 3x3 window (the shipped configurations): four adjacent pixels per thread share their 3 x 6 window values,
            // read as 32-bit pairs (the tile stride 66 is even and the group starts at an even element)
            for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) {
                const int ly = i >> 4, lx = (i & 15) * 4;
                if (bx * 64 + lx >= W || by * 64 + ly >= H) continue;
                int sum[4] = {0, 0, 0, 0};
                unsigned cnt[4] = {0, 0, 0, 0};
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    const uint32_t* rp = reinterpret_cast<const uint32_t*>(cur + (ly + l) * S + lx);  // columns lx-1 .. lx+4
                    int v[6];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t w2 = rp[k];
                        v[2 * k] = (int)(int16_t)(w2 & 0xFFFFu);
                        v[2 * k + 1] = (int)(int16_t)(w2 >> 16);
                    }
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const bool ok = v[k] > minD && v[k] < maxD;
                        const int m = ok ? v[k] : 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (k >= q && k <= q + 2) {
                                sum[q] += m;
                                cnt[q] += ok ? 1u : 0u;
                            }
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (bx * 64 + lx + q >= W) break;
                    int16_t res = kInvalid;
                    if (cnt[q] > minCount) res = (int16_t)__umulhi((unsigned)sum[q], recipM[cnt[q]]);
                    nxt[(ly + 1) * S + lx + q + 1] = res;
                }
            }
        } else
        for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
            const int ly = i >> 6, lx = i & 63;
            if (bx * 64 + lx >= W || by * 64 + ly >= H) continue;
            int sum = 0;
            unsigned count = 0;
            for (int l = -pad; l <= pad; ++l) {
                const int16_t* rowp = cur + (ly + l + pad) * S + lx + pad;
                for (int k = -pad; k <= pad; ++k) {
                    const int v = rowp[k];
                    if (v > minD && v < maxD) {
                        sum += v;
                        count++;
                    }
                }
            }
            // sum / count for 0 <= sum < 2^28 and count <= 81 by a rounded-up reciprocal (exact: sum * (M count - 2^32) < 2^32)
            int16_t res = kInvalid;
            if (count > minCount) res = (int16_t)__umulhi((unsigned)sum, recipM[count]);
            nxt[(ly + pad) * S + lx + pad] = res;
        }
        __syncthreads();
        int16_t* t = cur;
        cur = nxt;
        nxt = t;
        // every in-image tile cell of `nxt` is rewritten by the next pass and the other cells (halo,
        // out-of-image) are identical in both buffers, so no copy-back is needed
    }
    Img<int16_t> out = dst.frame(f);
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int ly = i >> 6, lx = i & 63;
        const int x = bx * 64 + lx, y = by * 64 + ly;
        if (x < W && y < H) out.at(x, y) = cur[(ly + pad) * S + lx + pad];
    }
}

cudaError_t post_set_kernel_attributes() {  // per device, from cartb200_create
    return cudaFuncSetAttribute(interpolate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

int launch_interpolate_from(cartb200_ctx* c, int n, ImgBatch<const int16_t> src, ImgBatch<int16_t> dst, int radius,
                            int iterations, int minD, int maxD, cudaStream_t s) {
    const int pad = radius - 1, S = 64 + 2 * pad;
    const size_t smem = (size_t)2 * S * S * sizeof(int16_t);
    if (smem > 200 * 1024 || (2 * radius - 1) * (2 * radius - 1) > 127) {
        c->err = "interpolate: smoothing radius too large for shared memory";
        return CARTB200_E_UNSUPPORTED;
    }
    dim3 grid(ceilDiv(c->W, 64), ceilDiv(c->H, 64), n);
    interpolate_kernel<<<grid, 256, smem, s>>>(src, dst, c->W, c->H, radius, iterations, minD, maxD);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

int launch_interpolate(cartb200_ctx* c, int n, ImgBatch<int16_t> disp, int radius, int iterations, int minD, int maxD,
                       cudaStream_t s) {
    if (radius <= 0) return CARTB200_OK;
    // stage the original image (canonical out-of-place read), medL is free at this point
    ImgBatch<int16_t> tmp{(int16_t*)c->medL, c->dispPitch, c->dispPitch * (size_t)c->H};
    for (int f = 0; f < n; ++f) {
        Img<int16_t> a = disp.frame(f), b = tmp.frame(f);
        CB_CHECK_CUDA(c, cudaMemcpy2DAsync(b.data, b.pitch, a.data, a.pitch, (size_t)c->W * 2, c->H,
                                           cudaMemcpyDeviceToDevice, s));
    }
    return launch_interpolate_from(c, n, ImgBatch<const int16_t>{tmp.data, tmp.pitch, tmp.frameStride}, disp, radius,
                                   iterations, minD, maxD, s);
}

// ---------------------------------------------------------------------------------------------
// cv::cuda::resize(src, dst, size, 0, 0, INTER_LINEAR) on CV_8UC3 - the KITTI source's optional resize
// (/root/reference/src/sources/kitti.cpp:166-169).  Normative behaviour: oracle/stages.cpp orc_resize_bgr8 (the same
// float operations in the same order, no FMA contraction: bit-identical to it; parity against OpenCV itself is unpinned).
__global__ void __launch_bounds__(256) resize_bgr8_kernel(const uint8_t* __restrict__ src, size_t srcPitch, int sw, int sh,
                                                          uint8_t* __restrict__ dst, size_t dstPitch, int dw, int dh, float fx,
                                                          float fy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    const float sx = (float)x * fx, sy = (float)y * fy;
    const int x1 = __float2int_rd(sx), y1 = __float2int_rd(sy);
    const int x2 = x1 + 1, y2 = y1 + 1;
    const int x1r = min(x1, sw - 1), y1r = min(y1, sh - 1), x2r = min(x2, sw - 1), y2r = min(y2, sh - 1);
    const float w11 = ((float)x2 - sx) * ((float)y2 - sy), w12 = (sx - (float)x1) * ((float)y2 - sy);
    const float w21 = ((float)x2 - sx) * (sy - (float)y1), w22 = (sx - (float)x1) * (sy - (float)y1);
    const uint8_t *r1 = src + (size_t)y1r * srcPitch, *r2 = src + (size_t)y2r * srcPitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float out = 0.0f;
        out = out + (float)__ldg(r1 + 3 * x1r + c) * w11;
        out = out + (float)__ldg(r1 + 3 * x2r + c) * w12;
        out = out + (float)__ldg(r2 + 3 * x1r + c) * w21;
        out = out + (float)__ldg(r2 + 3 * x2r + c) * w22;
        dst[(size_t)y * dstPitch + 3 * x + c] = (uint8_t)min(255, max(0, __float2int_rn(out)));
    }
}

int launch_resize_bgr8(const uint8_t* src, size_t srcPitch, int sw, int sh, uint8_t* dst, size_t dstPitch, int dw, int dh,
                       cudaStream_t s) {
    const float fx = (float)(1.0 / ((double)dw / (double)sw)), fy = (float)(1.0 / ((double)dh / (double)sh));
    dim3 grid(ceilDiv(dw, 256), dh);
    resize_bgr8_kernel<<<grid, 256, 0, s>>>(src, srcPitch, sw, sh, dst, dstPitch, dw, dh, fx, fy);
    return cudaPeekAtLastError() == cudaSuccess ? CARTB200_OK : CARTB200_E_CUDA;
}

// ---------------------------------------------------------------------------------------------
// calculateDirectionalDerivatives + mergeDerivativeHistograms,
// /root/reference/src/modules/disparity/derivative.cu:27-116.
// CTA = one 128-column reference tile x a band of kDerivRows local rows.  The band's rows -2 .. +2 of the reference's
// (bug-compatible) tile are staged in shared memory once - one closed-form evaluation per staged element, 1.3 per pixel,
// where reading the four taps of every pixel straight through the closed form cost 4 - then thread = one column.
constexpr int kDerivRows = 16;
__global__ void __launch_bounds__(128) derivative_kernel(ImgBatch<const int16_t> disp, ImgBatch<int16_t> deriv,
                                                         int32_t* __restrict__ hist, int W, int H) {
    __shared__ int sh[512];
    __shared__ int16_t tile[kDerivRows + 4][132];  // local rows r0 - 2 .. r0 + kDerivRows + 1, columns -2 .. 129
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sh[i] = 0;
    const int f = blockIdx.z, bx = blockIdx.x;
    constexpr int bandsPerTile = 128 / kDerivRows;
    const int by = blockIdx.y / bandsPerTile, r0 = (blockIdx.y % bandsPerTile) * kDerivRows;
    if (by * 128 + r0 < H) {  // uniform per CTA
        TileGeom g{W, H, 128, 128, 2, 2, 4, 4, 132 * 132};
        DispAccessor acc{disp.frame(f)};
        TileEval<int16_t, DispAccessor> te(acc, g, bx, by, kInvalid);
        for (int i = threadIdx.x; i < (kDerivRows + 4) * 132; i += blockDim.x) {
            const int k = i / 132, cidx = i - k * 132;
            tile[k][cidx] = te.template value<true>(cidx - 2, r0 - 2 + k);
        }
    }
    __syncthreads();
    const int lx = threadIdx.x, x = bx * 128 + lx;
    if (x < W) {
        Img<int16_t> out = deriv.frame(f);
        for (int r = 0; r < kDerivRows; ++r) {
            const int y = by * 128 + r0 + r;
            if (y >= H) break;
            const int16_t up = tile[r][lx + 2], dn = tile[r + 4][lx + 2];
            const int16_t lf = tile[r + 2][lx], rt = tile[r + 2][lx + 4];
            const int16_t dv = (int16_t)(dn - up), dh = (int16_t)(rt - lf);
            const bool vv = up != kInvalid && dn != kInvalid, hv = lf != kInvalid && rt != kInvalid;
            short2 o;
            o.x = vv ? dv : kInvalid;
            o.y = hv ? dh : kInvalid;
            *reinterpret_cast<short2*>(out.row(y) + 2 * x) = o;
            if (vv && dv >= -128 && dv <= 127) atomicAdd(&sh[2 * (dv + 128)], 1);
            if (hv && dh >= -128 && dh <= 127) atomicAdd(&sh[2 * (dh + 128) + 1], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[(size_t)f * 512 + i], sh[i]);
}

int launch_derivative(cartb200_ctx* c, int n, ImgBatch<const int16_t> disp, ImgBatch<int16_t> deriv, int32_t* hist,
                      cudaStream_t s) {
    CB_CHECK_CUDA(c, cudaMemsetAsync(hist, 0, (size_t)n * 512 * sizeof(int32_t), s));
    dim3 grid(ceilDiv(c->W, 128), ceilDiv(c->H, 128) * (128 / kDerivRows), n);
    derivative_kernel<<<grid, 128, 0, s>>>(disp, deriv, hist, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

// ---------------------------------------------------------------------------------------------
// calculateDerivatives (naive), /root/reference/src/modules/planeseg/planeseg.cu:31-142.
// CTA = one 128-column reference tile x a band of kNaiveRows local rows.  Raw rows (with the
// reference's halo semantics) are staged in shared memory, the 5-tap valid-mean is out-of-place
// (canonical schedule), halo rows stay unfiltered (Q10b).
constexpr int kNaiveRows = 16;
__global__ void __launch_bounds__(128) naive_derivative_kernel(ImgBatch<const int16_t> disp, ImgBatch<int16_t> deriv,
                                                               int32_t* __restrict__ hist, int W, int H) {
    __shared__ int sh[256];
    __shared__ int16_t raw[kNaiveRows + 6][128];  // local rows r0-3 .. r0+kNaiveRows+2
    __shared__ int16_t fil[kNaiveRows + 2][128];  // F rows r0-1 .. r0+kNaiveRows
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sh[i] = 0;
    const int f = blockIdx.z, bx = blockIdx.x;
    const int bandsPerTile = 128 / kNaiveRows;
    const int by = blockIdx.y / bandsPerTile, r0 = (blockIdx.y % bandsPerTile) * kNaiveRows;
    const int lx = threadIdx.x, x = bx * 128 + lx;
    TileGeom g{W, H, 128, 128, 0, 2, 4, 4, 128 * 144};
    DispAccessor acc{disp.frame(f)};
    TileEval<int16_t, DispAccessor> te(acc, g, bx, by, kInvalid);
    for (int k = 0; k < kNaiveRows + 6; ++k) {
        const int ly = r0 - 3 + k;
        raw[k][lx] = (ly >= -2 && ly < 130) ? te.template value<true>(lx, ly) : kInvalid;
    }
    // each thread only touches its own column: no barrier needed between the phases
    for (int k = 0; k < kNaiveRows + 2; ++k) {
        const int ly = r0 - 1 + k;  // local row of F
        int16_t v;
        if (ly < 0 || ly >= 128) {
            v = raw[k + 2][lx];  // unfiltered halo row
        } else {
            int16_t sum = 0;
            int count = 0;
#pragma unroll
            for (int t = 0; t < 5; ++t) {
                const int16_t d = raw[k + t][lx];
                if (d != kInvalid) {
                    sum = (int16_t)(sum + d);
                    count++;
                }
            }
            v = count == 0 ? kInvalid : (int16_t)(sum / count);
        }
        fil[k][lx] = v;
    }
    __syncthreads();  // histogram zeroing visible
    if (x < W) {
        Img<int16_t> out = deriv.frame(f);
        for (int k = 0; k < kNaiveRows; ++k) {
            const int y = by * 128 + r0 + k;
            if (y >= H) break;
            const int16_t p = fil[k][lx], cval = fil[k + 1][lx], nx = fil[k + 2][lx];
            const int16_t dv = (int16_t)(nx - p);
            const bool valid = cval != kInvalid && nx != kInvalid && p != kInvalid;
            out.at(x, y) = valid ? dv : kInvalid;
            if (valid && dv >= -128 && dv <= 127) atomicAdd(&sh[dv + 128], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[(size_t)f * 256 + i], sh[i]);
}

int launch_naive_derivative(cartb200_ctx* c, int n, ImgBatch<const int16_t> disp, ImgBatch<int16_t> deriv,
                            int32_t* hist, cudaStream_t s) {
    CB_CHECK_CUDA(c, cudaMemsetAsync(hist, 0, (size_t)n * 256 * sizeof(int32_t), s));
    const int tilesY = ceilDiv(c->H, 128);
    dim3 grid(ceilDiv(c->W, 128), tilesY * (128 / kNaiveRows), n);
    naive_derivative_kernel<<<grid, 128, 0, s>>>(disp, deriv, hist, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

// ---------------------------------------------------------------------------------------------
// classifyPlanes (naive) range rule, /root/reference/src/modules/planeseg/planeseg.cu:188-197.
__device__ __forceinline__ uint8_t classify_one(int16_t d, int hS, int hE, int vS, int vE) {
    if (d != kInvalid && d >= hS && d < hE) return CARTB200_PLANE_HORIZONTAL;
    if (d != kInvalid && d >= vS && d < vE) return CARTB200_PLANE_VERTICAL;
    return CARTB200_PLANE_UNKNOWN;
}

__global__ void __launch_bounds__(256) classify_kernel(ImgBatch<const int16_t> deriv, int channels, int channel,
                                                       const int32_t* __restrict__ params, ImgBatch<uint8_t> planes,
                                                       int W, int H) {
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int hS = params[4 * f], hE = params[4 * f + 1], vS = params[4 * f + 2], vE = params[4 * f + 3];
    const int16_t d = __ldg(deriv.frame(f).row(y) + (size_t)x * channels + channel);
    planes.frame(f).at(x, y) = classify_one(d, hS, hE, vS, vE);
}

int launch_classify(cartb200_ctx* c, int n, ImgBatch<const int16_t> deriv, int channels, int channel,
                    const int32_t* paramsDev, ImgBatch<uint8_t> planes, cudaStream_t s) {
    dim3 grid(ceilDiv(c->W, 256), c->H, n);
    classify_kernel<<<grid, 256, 0, s>>>(deriv, channels, channel, paramsDev, planes, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

// ---------------------------------------------------------------------------------------------
// performSuperPixelClassifications + classifyPlanes (SP),
// /root/reference/src/modules/planeseg/sp_planeseg.cu:25-184 (previousPlanesCount == 0).
// Votes are 32-bit global counters [frame][label][4]; lanes of a warp that hit the same
// (label, plane) counter are merged with __match_any_sync before the atomic.
__global__ void __launch_bounds__(256) sp_vote_kernel(ImgBatch<const int16_t> deriv, ImgBatch<const uint16_t> labels,
                                                      int maxLabel, const int32_t* __restrict__ params,
                                                      ImgBatch<uint8_t> unsm, uint32_t* __restrict__ votes,
                                                      int voteStride, int W, int H) {
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const bool in = x < W;
    unsigned key = 0xFFFFFFFFu;
    if (in) {
        const int hS = params[4 * f], hE = params[4 * f + 1], vS = params[4 * f + 2], vE = params[4 * f + 3];
        const int16_t d = __ldg(deriv.frame(f).row(y) + 2 * (size_t)x);
        const uint8_t p = classify_one(d, hS, hE, vS, vE);
        unsm.frame(f).at(x, y) = p;
        const unsigned l = __ldg(labels.frame(f).row(y) + x);
        if ((int)l < maxLabel) key = l * 4 + p;
    }
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, key);
    if (key != 0xFFFFFFFFu && (threadIdx.x & 31) == (__ffs(peers) - 1))
        atomicAdd(&votes[(size_t)f * voteStride + key], (unsigned)__popc(peers));
}

__global__ void __launch_bounds__(256) sp_assign_kernel(ImgBatch<const uint16_t> labels, int maxLabel,
                                                        const uint32_t* __restrict__ votes, int voteStride,
                                                        ImgBatch<uint8_t> planes, int W, int H) {
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const unsigned l = __ldg(labels.frame(f).row(y) + x);
    uint8_t best = CARTB200_PLANE_UNKNOWN;
    if ((int)l < maxLabel) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(votes + (size_t)f * voteStride + 4 * l));
        int maxVotes = (int)v.z;  // UNKNOWN
        if ((int)v.y > maxVotes) {
            maxVotes = (int)v.y;
            best = CARTB200_PLANE_VERTICAL;
        }
        if ((int)v.x > maxVotes) best = CARTB200_PLANE_HORIZONTAL;
    }
    planes.frame(f).at(x, y) = best;
}

int launch_sp_planeseg(cartb200_ctx* c, int n, ImgBatch<const int16_t> deriv, ImgBatch<const uint16_t> labels,
                       int maxLabel, const int32_t* paramsDev, ImgBatch<uint8_t> unsm, ImgBatch<uint8_t> planes,
                       cudaStream_t s) {
    const int voteStride = c->maxLabels * 4;
    CB_CHECK_CUDA(c, cudaMemsetAsync(c->votes, 0, (size_t)n * voteStride * sizeof(uint32_t), s));
    dim3 grid(ceilDiv(c->W, 256), c->H, n);
    sp_vote_kernel<<<grid, 256, 0, s>>>(deriv, labels, maxLabel, paramsDev, unsm, c->votes, voteStride, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    sp_assign_kernel<<<grid, 256, 0, s>>>(labels, maxLabel, c->votes, voteStride, planes, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

// ---------------------------------------------------------------------------------------------
// Temporal smoothing vote: the per-pixel loop of classifyPlanes
// (/root/reference/src/modules/planeseg/planeseg.cu:199-240, MODE 0) and of performSuperPixelClassifications
// (/root/reference/src/modules/planeseg/sp_planeseg.cu:79-117, MODE 1).  refs.planes[k] = "planes_unsmoothed" of
// frame id-(k+1), refs.flow[k] = "optflow" (CV_16SC2, S10.5) of frame id-k.  Every flow image is read at the CURRENT
// pixel (one coalesced 32-bit load per reference frame), the position walks back by the accumulated integer flow,
// out-of-image positions are skipped but stay accumulated.  The three vote counters live in the bytes of one
// register.  Streaming work: 4 B (flow) + 1 B (gathered plane) per reference frame and pixel -> HBM/latency-bound.
template <int MODE>
__device__ __forceinline__ uint8_t temporal_vote(const TemporalRefs& r, uint8_t plane, int px, int py, int W, int H) {
    unsigned votes = (MODE ? 2u : 1u) << (8 * plane);
    int x = px, y = py;
#pragma unroll 1
    for (int k = 0; k < r.count; ++k) {
        const int fl = __ldg(reinterpret_cast<const int*>(reinterpret_cast<const char*>(r.flow[k]) + (size_t)py * r.flowPitch[k]) + px);
        x -= (int)(int16_t)(fl & 0xFFFF) >> 5;
        y -= (fl >> 16) >> 5;
        if (x < 0 || y < 0 || x >= W || y >= H) continue;
        const unsigned v = __ldg(r.planes[k] + (size_t)y * r.planesPitch[k] + x);
        if (v <= 2u) votes += 1u << (8 * v);  // the reference indexes votes[] with the stored value; > 2 is out of contract
    }
    const unsigned vh = votes & 0xFFu, vv = (votes >> 8) & 0xFFu, vu = (votes >> 16) & 0xFFu;
    const unsigned best = vh > vv ? (unsigned)CARTB200_PLANE_HORIZONTAL : (unsigned)CARTB200_PLANE_VERTICAL;
    const unsigned vb = vh > vv ? vh : vv;
    if (MODE == 0) return vb == 0 ? CARTB200_PLANE_UNKNOWN : (uint8_t)best;
    return vb < vu ? CARTB200_PLANE_UNKNOWN : (uint8_t)best;
}

__global__ void __launch_bounds__(256) classify_temporal_kernel(Img<const int16_t> deriv, int channels, int channel, PlaneRanges pr,
                                                                TemporalRefs refs, Img<uint8_t> unsm, Img<uint8_t> smoothed,
                                                                int W, int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int16_t d = __ldg(deriv.row(y) + (size_t)x * channels + channel);
    const uint8_t p = classify_one(d, pr.hS, pr.hE, pr.vS, pr.vE);
    unsm.at(x, y) = p;
    smoothed.at(x, y) = refs.count > 0 ? temporal_vote<0>(refs, p, x, y, W, H) : p;
}

int launch_classify_temporal(cartb200_ctx* c, Img<const int16_t> deriv, int channels, int channel, PlaneRanges pr,
                             const TemporalRefs& refs, Img<uint8_t> unsm, Img<uint8_t> smoothed, cudaStream_t s) {
    dim3 grid(ceilDiv(c->W, 256), c->H, 1);
    classify_temporal_kernel<<<grid, 256, 0, s>>>(deriv, channels, channel, pr, refs, unsm, smoothed, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

// performSuperPixelClassifications with previousPlanesCount > 0: the voted plane (not the unsmoothed class) feeds
// the per-superpixel counters; "planes_unsmoothed" still holds the range-rule class (sp_planeseg.cu:77).
__global__ void __launch_bounds__(256) sp_vote_temporal_kernel(Img<const int16_t> deriv, Img<const uint16_t> labels, int maxLabel,
                                                               PlaneRanges pr, TemporalRefs refs, Img<uint8_t> unsm,
                                                               uint32_t* __restrict__ votes, int W, int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const bool in = x < W;
    unsigned key = 0xFFFFFFFFu;
    if (in) {
        const int16_t d = __ldg(deriv.row(y) + 2 * (size_t)x);
        uint8_t p = classify_one(d, pr.hS, pr.hE, pr.vS, pr.vE);
        unsm.at(x, y) = p;
        if (refs.count > 0) p = temporal_vote<1>(refs, p, x, y, W, H);
        const unsigned l = __ldg(labels.row(y) + x);
        if ((int)l < maxLabel) key = l * 4 + p;
    }
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, key);
    if (key != 0xFFFFFFFFu && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&votes[key], (unsigned)__popc(peers));
}

int launch_sp_planeseg_temporal(cartb200_ctx* c, Img<const int16_t> deriv, Img<const uint16_t> labels, int maxLabel, PlaneRanges pr,
                                const TemporalRefs& refs, Img<uint8_t> unsm, Img<uint8_t> planes, cudaStream_t s) {
    const int voteStride = c->maxLabels * 4;
    CB_CHECK_CUDA(c, cudaMemsetAsync(c->votes, 0, (size_t)voteStride * sizeof(uint32_t), s));
    dim3 grid(ceilDiv(c->W, 256), c->H, 1);
    sp_vote_temporal_kernel<<<grid, 256, 0, s>>>(deriv, labels, maxLabel, pr, refs, unsm, c->votes, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    sp_assign_kernel<<<grid, 256, 0, s>>>(ImgBatch<const uint16_t>{labels.data, labels.pitch, 0}, maxLabel, c->votes, voteStride,
                                          ImgBatch<uint8_t>{planes.data, planes.pitch, 0}, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

// ---------------------------------------------------------------------------------------------
// Superpixel consumers of the plane fit.
// countPixels, /root/reference/src/modules/planefit.cu:38-83: per label the pixel count and the number of pixels with
// an invalid depth Z (IS_VALID_DEPTH :19).  calculateRegionDistance, :85-138 with calculateDistanceFromPlane :34-36:
// per (plane, label) the number of valid pixels closer than `threshold` to the plane, in double precision with the
// reference's operation order (this library is compiled without FMA contraction, like the oracle).
// One thread per pixel; lanes of a warp that carry the same label are merged (__match_any_sync once per pixel, then
// one ballot per counter) so a 12x12-block superpixel costs ~3 atomics per warp row instead of 32.
__device__ __forceinline__ bool valid_depth(float z) { return isfinite(z) && (double)z <= 40.0 && (double)z > 0.0; }

__global__ void __launch_bounds__(256) label_statistics_kernel(Img<const uint16_t> labels, Img<const float> xyz, int nLabels,
                                                               uint32_t* __restrict__ count, uint32_t* __restrict__ invalid, int W,
                                                               int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    unsigned key = 0xFFFFFFFFu;
    bool inv = false;
    if (x < W) {
        const unsigned l = __ldg(labels.row(y) + x);
        if ((int)l < nLabels) {
            key = l;
            inv = !valid_depth(__ldg(xyz.row(y) + 3 * (size_t)x + 2));
        }
    }
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, key);
    const unsigned invMask = __ballot_sync(active, inv) & peers;
    if (key != 0xFFFFFFFFu && (threadIdx.x & 31) == (__ffs(peers) - 1)) {
        atomicAdd(&count[key], (unsigned)__popc(peers));
        if (invMask) atomicAdd(&invalid[key], (unsigned)__popc(invMask));
    }
}

int launch_label_statistics(cartb200_ctx* c, Img<const uint16_t> labels, Img<const float> xyz, int nLabels, uint32_t* count,
                            uint32_t* invalid, cudaStream_t s) {
    CB_CHECK_CUDA(c, cudaMemsetAsync(count, 0, (size_t)nLabels * sizeof(uint32_t), s));
    CB_CHECK_CUDA(c, cudaMemsetAsync(invalid, 0, (size_t)nLabels * sizeof(uint32_t), s));
    dim3 grid(ceilDiv(c->W, 256), c->H, 1);
    label_statistics_kernel<<<grid, 256, 0, s>>>(labels, xyz, nLabels, count, invalid, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

__global__ void __launch_bounds__(256) region_inliers_kernel(Img<const uint16_t> labels, Img<const float> xyz, int nLabels,
                                                             PlaneSet ps, double threshold, uint32_t* __restrict__ inliers, int W,
                                                             int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    unsigned key = 0xFFFFFFFFu;
    double X = 0.0, Y = 0.0, Z = 0.0;
    bool ok = false;
    if (x < W) {
        const unsigned l = __ldg(labels.row(y) + x);
        if ((int)l < nLabels) {
            key = l;
            const float* q = xyz.row(y) + 3 * (size_t)x;
            const float fz = __ldg(q + 2);
            ok = valid_depth(fz);
            X = (double)__ldg(q);
            Y = (double)__ldg(q + 1);
            Z = (double)fz;
        }
    }
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, key);
    const bool leader = key != 0xFFFFFFFFu && (threadIdx.x & 31) == (__ffs(peers) - 1);
#pragma unroll 1
    for (int p = 0; p < ps.count; ++p) {
        const double a = ps.abcd[p][0], b = ps.abcd[p][1], cc = ps.abcd[p][2], d = ps.abcd[p][3];
        const double dist = fabs(a * X + b * Y + cc * Z + d) / sqrt(a * a + b * b + cc * cc);
        const unsigned hit = __ballot_sync(active, ok && dist < threshold) & peers;
        if (leader && hit) atomicAdd(&inliers[(size_t)p * nLabels + key], (unsigned)__popc(hit));
    }
}

int launch_region_inliers(cartb200_ctx* c, Img<const uint16_t> labels, Img<const float> xyz, int nLabels, const double* planesHost,
                          int nPlanes, double threshold, uint32_t* inliers, cudaStream_t s) {
    CB_CHECK_CUDA(c, cudaMemsetAsync(inliers, 0, (size_t)nPlanes * nLabels * sizeof(uint32_t), s));
    dim3 grid(ceilDiv(c->W, 256), c->H, 1);
    for (int p0 = 0; p0 < nPlanes; p0 += kMaxPlaneSet) {
        PlaneSet ps;
        ps.count = nPlanes - p0 < kMaxPlaneSet ? nPlanes - p0 : kMaxPlaneSet;
        for (int i = 0; i < ps.count; ++i)
            for (int j = 0; j < 4; ++j) ps.abcd[i][j] = planesHost[4 * (p0 + i) + j];
        region_inliers_kernel<<<grid, 256, 0, s>>>(labels, xyz, nLabels, ps, threshold, inliers + (size_t)p0 * nLabels, c->W, c->H);
        CB_LAUNCH_CHECK(c);
    }
    return CARTB200_OK;
}

// ---------------------------------------------------------------------------------------------
// Overlay kernels (visual QA).  overlayPlanes, /root/reference/src/modules/planeseg/planeseg_vis.cu:28-56
// (colour table :21-26 = PlaneColor / 2, /root/reference/include/modules/planeseg.hpp:44-66) and
// overlayBoundaryVisualization, /root/reference/src/modules/superpixels/visualization.cu:9-42 (last row / column not
// written).  One thread per pixel, 7 B of traffic per pixel: HBM-bound, launch-latency sized at one frame.
__global__ void __launch_bounds__(256) overlay_planes_kernel(Img<const uint8_t> bgr, Img<const uint8_t> planes, Img<uint8_t> out, int W,
                                                             int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const unsigned p = __ldg(planes.row(y) + x);
    const uint8_t* in = bgr.row(y) + 3 * (size_t)x;
    uint8_t* o = out.row(y) + 3 * (size_t)x;
    o[0] = (uint8_t)(__ldg(in) / 2 + (p == CARTB200_PLANE_HORIZONTAL ? 127 : 0));
    o[1] = (uint8_t)(__ldg(in + 1) / 2 + (p == CARTB200_PLANE_VERTICAL ? 127 : 0));
    o[2] = (uint8_t)(__ldg(in + 2) / 2 + (p == CARTB200_PLANE_UNKNOWN ? 127 : 0));
}

__global__ void __launch_bounds__(256) overlay_boundaries_kernel(Img<const uint8_t> bgr, Img<const uint16_t> labels, Img<uint8_t> out,
                                                                 int W, int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W - 1 || y >= H - 1) return;
    const unsigned l = __ldg(labels.row(y) + x);
    const bool edge = l != __ldg(labels.row(y) + x + 1) || l != __ldg(labels.row(y + 1) + x);
    const uint8_t* in = bgr.row(y) + 3 * (size_t)x;
    uint8_t* o = out.row(y) + 3 * (size_t)x;
    o[0] = edge ? 0 : __ldg(in);
    o[1] = edge ? 0 : __ldg(in + 1);
    o[2] = edge ? 255 : __ldg(in + 2);
}

int launch_overlay_planes(cartb200_ctx* c, Img<const uint8_t> bgr, Img<const uint8_t> planes, Img<uint8_t> out, cudaStream_t s) {
    dim3 grid(ceilDiv(c->W, 256), c->H, 1);
    overlay_planes_kernel<<<grid, 256, 0, s>>>(bgr, planes, out, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

int launch_overlay_boundaries(cartb200_ctx* c, Img<const uint8_t> bgr, Img<const uint16_t> labels, Img<uint8_t> out, cudaStream_t s) {
    dim3 grid(ceilDiv(c->W, 256), c->H, 1);
    overlay_boundaries_kernel<<<grid, 256, 0, s>>>(bgr, labels, out, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

// ---------------------------------------------------------------------------------------------
// DepthModule::runInternal, /root/reference/src/modules/depth.cpp:9-25: disparity * (1/16) as float, then the
// third-party cv::cuda::reprojectImageTo3D(Q, 3 channels).  Normative arithmetic: oracle/stages.cpp orc_depth
// (single precision, same operation order; compiled without FMA contraction).  12 B written per pixel: HBM-bound.
struct QMat {
    float q[16];
};

__global__ void __launch_bounds__(256) depth_kernel(ImgBatch<const int16_t> disp, ImgBatch<float> xyz, QMat Q, int W, int H) {
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const float d = (float)__ldg(disp.frame(f).row(y) + x) * (1.0f / 16.0f);
    const float fx = (float)x, fy = (float)y;
    const float qx = fx * Q.q[0] + fy * Q.q[1] + Q.q[3], qy = fx * Q.q[4] + fy * Q.q[5] + Q.q[7];
    const float qz = fx * Q.q[8] + fy * Q.q[9] + Q.q[11], qw = fx * Q.q[12] + fy * Q.q[13] + Q.q[15];
    const float iW = 1.0f / (qw + Q.q[14] * d);
    float* o = xyz.frame(f).row(y) + 3 * (size_t)x;
    o[0] = (qx + Q.q[2] * d) * iW;
    o[1] = (qy + Q.q[6] * d) * iW;
    o[2] = (qz + Q.q[10] * d) * iW;
}

int launch_depth(cartb200_ctx* c, int n, ImgBatch<const int16_t> disp, ImgBatch<float> xyz, const float* q16Host, cudaStream_t s) {
    QMat Q;
    for (int i = 0; i < 16; ++i) Q.q[i] = q16Host[i];
    dim3 grid(ceilDiv(c->W, 256), c->H, n);
    depth_kernel<<<grid, 256, 0, s>>>(disp, xyz, Q, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

}  // namespace cb
