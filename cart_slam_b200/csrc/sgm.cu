// Semi-global matching stage, written from scratch for sm_100a.  It replaces the third-party call
// cv::cuda::StereoSGM::compute (/root/reference/src/modules/disparity/disparity.cu:71) and the two
// cvtColor calls in front of it (:66-67).  Normative behaviour: oracle/sgm.cpp decisions D1-D9.
//
// Layout in HBM (per context, sized for max_batch frames):
//   gray   u8  [B][H][grayPitch]          census u32 [B][H][censusPitch/4]
//   volume u8  [P][B][H][W][D]   (D contiguous: one pixel's disparity vector is one or two 128-B lines)
//   wta    u16 [B][H][dispPitch/2] x {left raw, right raw}
// Arithmetic: path costs are <= 31 + P2 (u8 in memory); inside the kernels two disparities share a
// register as u16x2 and the recurrence runs on the DPX/video integer instructions of sm_90+/sm_100
// (VIADDMNMX.U16x2, VIMNMX.U16x2, VIMNMX3) - no tensor cores: nothing here is a dense contraction.
#include "common.cuh"

namespace cb {

// ---------------------------------------------------------------------------------------------
// gray + census.  D1: Y = (1868 B + 9617 G + 4899 R + 8192) >> 14.  D2: 9x7 centre-symmetric census,
// 31 bits, zero where the window leaves the image.
constexpr int kCenTW = 120, kCenTH = 32;  // output tile; staged tile is (120+8) x (32+6)
__global__ void __launch_bounds__(256) gray_census_kernel(ImgBatch<const uint8_t> left, ImgBatch<const uint8_t> right,
                                                          uint8_t* __restrict__ grayL, size_t grayPitch,
                                                          uint32_t* __restrict__ cenL, uint32_t* __restrict__ cenR,
                                                          size_t cenPitch, int W, int H) {
    __shared__ uint8_t g[kCenTH + 6][kCenTW + 8];
    const int f = blockIdx.z >> 1, side = blockIdx.z & 1;
    Img<const uint8_t> src = side ? right.frame(f) : left.frame(f);
    const int x0 = blockIdx.x * kCenTW - 4, y0 = blockIdx.y * kCenTH - 3;
    for (int i = threadIdx.x; i < (kCenTH + 6) * (kCenTW + 8); i += blockDim.x) {
        const int ty = i / (kCenTW + 8), tx = i % (kCenTW + 8);
        const int x = x0 + tx, y = y0 + ty;
        uint8_t v = 0;
        if (x >= 0 && x < W && y >= 0 && y < H) {
            const uint8_t* p = src.row(y) + 3 * (size_t)x;
            const unsigned b = __ldg(p), gg = __ldg(p + 1), r = __ldg(p + 2);
            v = (uint8_t)((b * 1868u + gg * 9617u + r * 4899u + 8192u) >> 14);
            if (side == 0 && tx >= 4 && tx < kCenTW + 4 && ty >= 3 && ty < kCenTH + 3)
                grayL[((size_t)f * H + y) * grayPitch + x] = v;
        }
        g[ty][tx] = v;
    }
    __syncthreads();
    uint32_t* out = (side ? cenR : cenL) + (size_t)f * H * (cenPitch / 4);
    for (int i = threadIdx.x; i < kCenTH * kCenTW; i += blockDim.x) {
        const int ty = i / kCenTW, tx = i % kCenTW;
        const int x = x0 + 4 + tx, y = y0 + 3 + ty;
        if (x >= W || y >= H) continue;
        uint32_t c = 0;
        if (x >= 4 && x < W - 4 && y >= 3 && y < H - 3) {
            const int cx = tx + 4, cy = ty + 3;
#pragma unroll
            for (int dy = -3; dy < 0; ++dy)
#pragma unroll
                for (int dx = -4; dx <= 4; ++dx) c = (c << 1) | (uint32_t)(g[cy + dy][cx + dx] > g[cy - dy][cx - dx]);
#pragma unroll
            for (int dx = -4; dx < 0; ++dx) c = (c << 1) | (uint32_t)(g[cy][cx + dx] > g[cy][cx - dx]);
        }
        out[(size_t)y * (cenPitch / 4) + x] = c;
    }
}

int launch_gray_census(cartb200_ctx* c, int n, ImgBatch<const uint8_t> left, ImgBatch<const uint8_t> right,
                       cudaStream_t s) {
    dim3 grid(ceilDiv(c->W, kCenTW), ceilDiv(c->H, kCenTH), 2 * n);
    gray_census_kernel<<<grid, 256, 0, s>>>(left, right, c->grayL, c->grayPitch, c->censusL, c->censusR,
                                            c->censusPitch, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

// ---------------------------------------------------------------------------------------------
// Path aggregation (D3 + D4).  A group of LPP = D/16 lanes owns one path line; each lane carries 16
// disparities as 8 x u16x2.  Per step: matching cost from the two census words, the recurrence
//   L(d) = C(d) + min(L'(d) - m, L'(d-1) - m + P1, L'(d+1) - m + P1, P2)
// on packed halves, a min-reduction over the group by warp shuffles, one 16-byte store per lane
// (a group writes the pixel's whole D-vector: full 128-B lines).
struct PathArgs {
    const uint32_t* cenL;
    const uint32_t* cenR;
    size_t cenStride;       // words per row
    size_t cenFrameStride;  // words per frame
    uint8_t* vol;           // this path's volume [B][H][W][D]
    size_t volFrameStride;
    int W, H, minDisp, P1, P2;
    int dx, dy;
};

__device__ __forceinline__ uint32_t pack16(uint32_t lo, uint32_t hi) { return lo | (hi << 16); }

// One DP step for a lane's 16 disparities. dp[8] holds L' (previous pixel) on entry, L on exit.
// prevHi = q of disparity (16*lane - 1), nextLo = q of disparity (16*lane + 16); both already minus m, or
// the 0x7FFF sentinel at the ends of the disparity range.  cost[8] packed.  Returns the lane-local minimum.
__device__ __forceinline__ uint32_t dp_step(uint32_t (&dp)[8], const uint32_t (&cost)[8], uint32_t M, uint32_t prevHi,
                                            uint32_t nextLo, uint32_t P1v, uint32_t P2v) {
    uint32_t q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = dp[i] - M;  // halves are >= m: no borrow between them
    uint32_t s[9];                                  // s[i] = (q[2i-1], q[2i])
    s[0] = __byte_perm(prevHi, q[0], 0x5410);       // (prevHi.lo, q0.lo)
#pragma unroll
    for (int i = 1; i < 8; ++i) s[i] = __byte_perm(q[i - 1], q[i], 0x5432);
    s[8] = __byte_perm(q[7], nextLo, 0x5432);
    uint32_t mn = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        // s[i]   = (q[2i-1], q[2i])   : the lower neighbour of each half
        // s[i+1] = (q[2i+1], q[2i+2]) : the upper neighbour of each half
        const uint32_t a = __viaddmin_u16x2(s[i], P1v, q[i]);     // min(q[d-1] + P1, q[d])
        const uint32_t b = __viaddmin_u16x2(s[i + 1], P1v, P2v);  // min(q[d+1] + P1, P2)
        const uint32_t r = __vminu2(a, b);
        dp[i] = r + cost[i];
        mn = __vminu2(mn, dp[i]);
    }
    return min(mn & 0xFFFFu, mn >> 16);
}

template <int LPP>
__device__ __forceinline__ uint32_t group_min(uint32_t v) {
#pragma unroll
    for (int o = 1; o < LPP; o <<= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

__device__ __forceinline__ void store_dp(uint8_t* dst, const uint32_t (&dp)[8]) {
    uint4 o;
    o.x = __byte_perm(dp[0], dp[1], 0x6420);
    o.y = __byte_perm(dp[2], dp[3], 0x6420);
    o.z = __byte_perm(dp[4], dp[5], 0x6420);
    o.w = __byte_perm(dp[6], dp[7], 0x6420);
    *reinterpret_cast<uint4*>(dst) = o;
}

// Generic kernel: any of the 8 directions.  Horizontal lines keep the 16 right-census words in a
// register window that slides by one word per step; other directions fetch them per step (L1-resident).
template <int D, int DXT>  // DXT: +1 / -1 horizontal specialisation, 0 = generic
__global__ void __launch_bounds__(128) aggregate_path_kernel(PathArgs a) {
    constexpr int LPP = D / 16;
    constexpr int GPB = 128 / LPP;  // groups per block
    const int lane = threadIdx.x % LPP;
    const int group = threadIdx.x / LPP;
    const int f = blockIdx.y;
    const int W = a.W, H = a.H;
    const int dx = DXT != 0 ? DXT : a.dx, dy = DXT != 0 ? 0 : a.dy;
    const int line = blockIdx.x * GPB + group;
    int nLines;
    if (dy == 0)
        nLines = H;
    else if (dx == 0)
        nLines = W;
    else
        nLines = W + H - 1;
    // all lanes of a warp must stay in the shuffles: inactive groups walk zero steps but still sync
    int x, y, len;
    if (line >= nLines) {
        x = y = 0;
        len = 0;
    } else if (dy == 0) {
        y = line;
        x = dx > 0 ? 0 : W - 1;
        len = W;
    } else if (dx == 0) {
        x = line;
        y = dy > 0 ? 0 : H - 1;
        len = H;
    } else {
        const int ys = dy > 0 ? 0 : H - 1, xs = dx > 0 ? 0 : W - 1;
        if (line < W) {
            x = line;
            y = ys;
        } else {
            x = xs;
            y = ys + dy * (line - W + 1);
        }
        const int lx = dx > 0 ? W - x : x + 1, ly = dy > 0 ? H - y : y + 1;
        len = min(lx, ly);
    }
    // steps must be warp-uniform for the shuffles
    int maxLen = len;
#pragma unroll
    for (int o = LPP; o < 32; o <<= 1) maxLen = max(maxLen, __shfl_xor_sync(0xFFFFFFFFu, maxLen, o));

    const uint32_t* cl = a.cenL + (size_t)f * a.cenFrameStride;
    const uint32_t* cr = a.cenR + (size_t)f * a.cenFrameStride;
    uint8_t* vol = a.vol + (size_t)f * a.volFrameStride;
    const uint32_t P1v = pack16(a.P1, a.P1), P2v = pack16(a.P2, a.P2);
    const int dbase = 16 * lane + a.minDisp;

    uint32_t dp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dp[i] = 0;
    uint32_t m = 0;
    uint32_t w[16];
    if (DXT != 0) {
        // window for the position BEFORE the first step (so the first step's shift brings it in place)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int xr = (x - dx) - dbase - j;
            w[j] = (len > 0 && xr >= 0 && xr < W) ? __ldg(cr + (size_t)y * a.cenStride + xr) : 0u;
        }
    }
    for (int step = 0; step < maxLen; ++step) {
        const bool act = step < len;
        uint32_t cost[8];
        if (act) {
            const uint32_t* crow = cr + (size_t)y * a.cenStride;
            const uint32_t l = __ldg(cl + (size_t)y * a.cenStride + x);
            if (DXT > 0) {
#pragma unroll
                for (int j = 15; j > 0; --j) w[j] = w[j - 1];
                const int xr = x - dbase;
                w[0] = (xr >= 0 && xr < W) ? __ldg(crow + xr) : 0u;
            } else if (DXT < 0) {
#pragma unroll
                for (int j = 0; j < 15; ++j) w[j] = w[j + 1];
                const int xr = x - dbase - 15;
                w[15] = (xr >= 0 && xr < W) ? __ldg(crow + xr) : 0u;
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int xr = x - dbase - j;
                    w[j] = (xr >= 0 && xr < W) ? __ldg(crow + xr) : 0u;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) cost[i] = pack16(__popc(l ^ w[2 * i]), __popc(l ^ w[2 * i + 1]));
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) cost[i] = 0;
        }
        const uint32_t M = pack16(m, m);
        // neighbours across lanes (q = dp - m), sentinel at the ends of the disparity range
        uint32_t up = __shfl_up_sync(0xFFFFFFFFu, dp[7], 1);
        uint32_t dn = __shfl_down_sync(0xFFFFFFFFu, dp[0], 1);
        const uint32_t prevHi = lane == 0 ? 0x7FFFu : ((up - M) >> 16);
        const uint32_t nextLo = lane == LPP - 1 ? 0x7FFFu : ((dn - M) & 0xFFFFu);
        uint32_t lm = dp_step(dp, cost, M, prevHi, nextLo, P1v, P2v);
        lm = group_min<LPP>(lm);
        if (act) {
            m = lm;
            store_dp(vol + ((size_t)y * W + x) * D + 16 * lane, dp);
            x += dx;
            y += dy;
        }
    }
}

template <int D>
static void launch_path_D(const PathArgs& a, int n, cudaStream_t s) {
    constexpr int GPB = 128 / (D / 16);
    int nLines = a.dy == 0 ? a.H : (a.dx == 0 ? a.W : a.W + a.H - 1);
    dim3 grid(ceilDiv(nLines, GPB), n);
    if (a.dy == 0 && a.dx > 0)
        aggregate_path_kernel<D, 1><<<grid, 128, 0, s>>>(a);
    else if (a.dy == 0 && a.dx < 0)
        aggregate_path_kernel<D, -1><<<grid, 128, 0, s>>>(a);
    else
        aggregate_path_kernel<D, 0><<<grid, 128, 0, s>>>(a);
}

static const int kDirs[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {-1, 1}, {1, -1}, {-1, -1}};

int launch_aggregate(cartb200_ctx* c, int n, cudaStream_t s) {
    for (int p = 0; p < c->P; ++p) {
        PathArgs a;
        a.cenL = c->censusL;
        a.cenR = c->censusR;
        a.cenStride = c->censusPitch / 4;
        a.cenFrameStride = a.cenStride * c->H;
        a.vol = c->volumes + (size_t)p * c->volPathStride;
        a.volFrameStride = c->volFrameStride;
        a.W = c->W;
        a.H = c->H;
        a.minDisp = c->cfg.min_disparity;
        a.P1 = c->cfg.p1;
        a.P2 = c->cfg.p2;
        a.dx = kDirs[p][0];
        a.dy = kDirs[p][1];
        switch (c->D) {
            case 64: launch_path_D<64>(a, n, s); break;
            case 128: launch_path_D<128>(a, n, s); break;
            case 256: launch_path_D<256>(a, n, s); break;
            default: c->err = "num_disparities must be 64, 128 or 256"; return CARTB200_E_UNSUPPORTED;
        }
        CB_LAUNCH_CHECK(c);
    }
    return CARTB200_OK;
}

// ---------------------------------------------------------------------------------------------
// Winner-takes-all (D5 + D6).  One CTA per image row; the row is processed in chunks of CH pixels.
// LPP lanes per pixel, 16 disparities per lane: the P path volumes are read once with 16-byte loads,
// summed to u16, kept in a shared-memory ring of (CH + D) pixels for the right-image minimum
// dR(x) = argmin_d S(x+d, d), which lags the left pass by D-1 pixels.
struct WtaArgs {
    const uint8_t* vol;
    size_t volPathStride, volFrameStride;
    int P, W, H;
    uint16_t* left;
    uint16_t* right;
    size_t pitch;  // elements
    float uniq;
};

template <int D, int CH>
__global__ void __launch_bounds__(256) wta_kernel(WtaArgs a) {
    constexpr int LPP = D / 16;
    constexpr int PPI = 256 / LPP;  // pixels per CTA iteration
    constexpr int R = CH + D;       // ring size in pixels
    extern __shared__ uint16_t ring[];  // [R][D]
    const int y = blockIdx.x, f = blockIdx.y;
    const int W = a.W;
    const int lane = threadIdx.x % LPP, grp = threadIdx.x / LPP;
    const uint8_t* vbase = a.vol + (size_t)f * a.volFrameStride + (size_t)y * W * D + 16 * lane;
    uint16_t* outL = a.left + ((size_t)f * a.H + y) * a.pitch;
    uint16_t* outR = a.right + ((size_t)f * a.H + y) * a.pitch;
    const unsigned gmask = 0xFFFFFFFFu;
    const int nChunks = (W + CH - 1) / CH;
    int rightNext = 0;  // next right pixel to finalise
    for (int ck = 0; ck <= nChunks; ++ck) {
        // ---- left pass over chunk ck -------------------------------------------------------
        if (ck < nChunks) {
            for (int px = grp; px < CH; px += PPI) {
                const int x = ck * CH + px;
                const bool in = x < W;
                uint32_t S[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) S[i] = 0;
                if (in) {
                    for (int p = 0; p < a.P; ++p) {
                        const uint4 v = __ldg(reinterpret_cast<const uint4*>(vbase + (size_t)p * a.volPathStride + (size_t)x * D));
                        S[0] += __byte_perm(v.x, 0, 0x4140);
                        S[1] += __byte_perm(v.x, 0, 0x4342);
                        S[2] += __byte_perm(v.y, 0, 0x4140);
                        S[3] += __byte_perm(v.y, 0, 0x4342);
                        S[4] += __byte_perm(v.z, 0, 0x4140);
                        S[5] += __byte_perm(v.z, 0, 0x4342);
                        S[6] += __byte_perm(v.w, 0, 0x4140);
                        S[7] += __byte_perm(v.w, 0, 0x4342);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(ring + (size_t)(x % R) * D + 16 * lane);
                    dst[0] = make_uint4(S[0], S[1], S[2], S[3]);
                    dst[1] = make_uint4(S[4], S[5], S[6], S[7]);
                }
                // top-2 over packed (S << 16 | d)
                uint32_t b1 = 0xFFFFFFFFu, b2 = 0xFFFFFFFFu;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t d0 = 16 * lane + 2 * i;
                    const uint32_t p0 = (S[i] << 16) | d0, p1 = (S[i] & 0xFFFF0000u) | (d0 + 1);
                    b2 = min(b2, max(b1, p0));
                    b1 = min(b1, p0);
                    b2 = min(b2, max(b1, p1));
                    b1 = min(b1, p1);
                }
#pragma unroll
                for (int o = 1; o < LPP; o <<= 1) {
                    const uint32_t o1 = __shfl_xor_sync(gmask, b1, o), o2 = __shfl_xor_sync(gmask, b2, o);
                    b2 = min(min(b2, o2), max(b1, o1));
                    b1 = min(b1, o1);
                }
                __syncwarp();
                if (in && lane == 0) {
                    const int c1 = (int)(b1 >> 16), d1 = (int)(b1 & 0xFFFF), c2 = (int)(b2 >> 16), d2 = (int)(b2 & 0xFFFF);
                    const bool reject = (__fmul_rn((float)c2, a.uniq) < (float)c1) && (abs(d1 - d2) > 1);
                    uint16_t v = 0xFFFF;
                    if (!reject) {
                        int subp = d1 << 4;
                        if (d1 > 0 && d1 < D - 1) {
                            const uint16_t* Sx = ring + (size_t)(x % R) * D;
                            const int l = Sx[d1 - 1], r = Sx[d1 + 1];
                            const int numer = l - r, denom = l - 2 * c1 + r;
                            if (denom != 0) subp += ((numer << 4) + denom) / (2 * denom);
                        }
                        v = (uint16_t)subp;
                    }
                    outL[x] = v;
                }
            }
        }
        __syncthreads();
        // ---- right pass: pixels whose window [x', x'+D) is complete (or truncated by the row end) ----
        const int avail = min(W, (ck + 1) * CH);  // left pixels < avail are in the ring
        const int rightEnd = (ck >= nChunks - 1) ? W : max(0, avail - D + 1);
        for (int xb = rightNext; xb < rightEnd; xb += PPI) {  // warp-uniform trip count (shuffles inside)
            const int xr = xb + grp;
            uint32_t best = 0xFFFFFFFFu;
            if (xr < rightEnd) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int d = 16 * lane + j;
                    if (xr + d < W) {
                        const uint32_t sv = ring[(size_t)((xr + d) % R) * D + d];
                        best = min(best, (sv << 16) | (uint32_t)d);
                    }
                }
            }
#pragma unroll
            for (int o = 1; o < LPP; o <<= 1) best = min(best, __shfl_xor_sync(gmask, best, o));
            if (lane == 0 && xr < rightEnd) outR[xr] = (uint16_t)(best & 0xFFFF);
        }
        rightNext = max(rightNext, rightEnd);
        __syncthreads();
        if (ck >= nChunks - 1) break;
    }
}

template <int D, int CH>
static int launch_wta_D(cartb200_ctx* c, const WtaArgs& a, int n, cudaStream_t s) {
    const size_t smem = (size_t)(CH + D) * D * sizeof(uint16_t);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(wta_kernel<D, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    dim3 grid(c->H, n);
    wta_kernel<D, CH><<<grid, 256, smem, s>>>(a);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

int launch_wta(cartb200_ctx* c, int n, cudaStream_t s) {
    WtaArgs a;
    a.vol = c->volumes;
    a.volPathStride = c->volPathStride;
    a.volFrameStride = c->volFrameStride;
    a.P = c->P;
    a.W = c->W;
    a.H = c->H;
    a.left = c->wtaL;
    a.right = c->wtaR;
    a.pitch = c->dispPitch / 2;
    a.uniq = (float)(100 - c->cfg.uniqueness_ratio) / 100.0f;
    switch (c->D) {
        case 64: return launch_wta_D<64, 64>(c, a, n, s);
        case 128: return launch_wta_D<128, 64>(c, a, n, s);
        case 256: return launch_wta_D<256, 32>(c, a, n, s);
    }
    c->err = "num_disparities must be 64, 128 or 256";
    return CARTB200_E_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------
// 3x3 medians (D7), left/right consistency (D8) and range correction (D9) in one pass.
__host__ __device__ __forceinline__ void cswap(uint32_t& a, uint32_t& b) {
    const uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
    a = lo;
    b = hi;
}
// 19-exchange median-of-9 network (Paeth / Devillard opt_med9)
__host__ __device__ __forceinline__ uint32_t median9(uint32_t* v) {
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[3]); cswap(v[5], v[8]); cswap(v[4], v[7]);
    cswap(v[3], v[6]); cswap(v[1], v[4]); cswap(v[2], v[5]);
    cswap(v[4], v[7]); cswap(v[4], v[2]); cswap(v[6], v[4]);
    cswap(v[4], v[2]);
    return v[4];
}
__device__ __forceinline__ uint32_t median_at(const uint16_t* img, size_t pitch, int W, int H, int x, int y) {
    if (x < 1 || y < 1 || x >= W - 1 || y >= H - 1) return __ldg(img + (size_t)y * pitch + x);
    uint32_t v[9];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) v[j * 3 + i] = __ldg(img + (size_t)(y + j - 1) * pitch + x + i - 1);
    return median9(v);
}

__global__ void __launch_bounds__(256) sgm_post_kernel(const uint16_t* __restrict__ wl, const uint16_t* __restrict__ wr,
                                                       size_t pitch, const uint8_t* __restrict__ grayL, size_t grayPitch,
                                                       ImgBatch<int16_t> out, int W, int H, int minDisp) {
    const int f = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const uint16_t* L = wl + (size_t)f * H * pitch;
    const uint16_t* Rr = wr + (size_t)f * H * pitch;
    const uint32_t org = median_at(L, pitch, W, H, x, y);
    const int d = (int)org >> 4;
    const int k = x - d;
    bool invalid = grayL[((size_t)f * H + y) * grayPitch + x] == 0 || org == 0xFFFFu;
    if (!invalid && k >= 0 && k < W) invalid = abs((int)median_at(Rr, pitch, W, H, k, y) - d) > 1;
    out.frame(f).at(x, y) = invalid ? (int16_t)((minDisp - 1) * 16) : (int16_t)(uint16_t)(org + minDisp * 16);
}

uint32_t debug_median9_host(const uint16_t* v9) {
    uint32_t v[9];
    for (int i = 0; i < 9; ++i) v[i] = v9[i];
    return median9(v);
}

int launch_sgm_post(cartb200_ctx* c, int n, ImgBatch<int16_t> disp, cudaStream_t s) {
    dim3 grid(ceilDiv(c->W, 256), c->H, n);
    sgm_post_kernel<<<grid, 256, 0, s>>>(c->wtaL, c->wtaR, c->dispPitch / 2, c->grayL, c->grayPitch, disp, c->W, c->H,
                                         c->cfg.min_disparity);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

}  // namespace cb
