// Semi-global matching stage, written from scratch for sm_100a.  It replaces the third-party call
// cv::cuda::StereoSGM::compute (/root/reference/src/modules/disparity/disparity.cu:71) and the two
// cvtColor calls in front of it (:66-67).  Normative behaviour: oracle/sgm.cpp decisions D1-D9.
//
// Layout in HBM (per context, sized for max_batch frames):
//   gray   u8  [B][H][grayPitch]          census u32 [B][H][censusPitch/4]
//   volume u8  [P][B][H][W][D]   (D contiguous: one pixel's disparity vector is one or two 128-B lines)
//   wta    left raw u16 [B][H][dispPitch/2]; right raw as u32 keys (cost, cost high byte, disparity) [B][H][rkPitch]
// Arithmetic: path costs are <= 31 + P2 (u8 in memory); inside the kernels two disparities share a
// register as u16x2 and the recurrence runs on the DPX/video integer instructions of sm_90+/sm_100
// (VIADDMNMX.U16x2, VIMNMX.U16x2, VIMNMX3) - no tensor cores: nothing here is a dense contraction.
#include <cuda_pipeline.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace cb {

// ---------------------------------------------------------------------------------------------
// gray + census.  D1: Y = (1868 B + 9617 G + 4899 R + 8192) >> 14.  D2: 9x7 centre-symmetric census,
// 31 bits, zero where the window leaves the image.
constexpr int kCenTW = 120, kCenTH = 32;  // output tile; staged tile is (120+8) x (32+6)
__global__ void __launch_bounds__(256) gray_census_kernel(ImgBatch<const uint8_t> left, ImgBatch<const uint8_t> right,
                                                          uint8_t* __restrict__ grayL, size_t grayPitch,
                                                          uint32_t* __restrict__ cenL, uint32_t* __restrict__ cenR,
                                                          size_t cenRowWords, int cenMargin, int minDisp, int W, int H) {
    __shared__ __align__(16) uint8_t g[kCenTH + 6][kCenTW + 8];
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
    // census rows carry zero margins (never written after create) so that the aggregation kernels can read
    // out-of-image columns as the 0 the specification asks for; the RIGHT census is stored shifted by
    // min_disparity: element (margin + i) holds cR[i - minDisp], i.e. the word for (x, d) is at x - d.
    uint32_t* out = (side ? cenR : cenL) + (size_t)f * H * cenRowWords + cenMargin + (side ? minDisp : 0);
    // Four horizontally adjacent pixels per thread: the 31 comparisons act on packed bytes (a > b per byte in four
    // logic / add operations on the whole word), the gray rows come as aligned 32-bit words (21 shared loads instead of
    // 248 byte loads), and the bits are collected bytewise - accumulator k gathers bits 8k .. 8k+7 of all four pixels -
    // and transposed into the four census words at the end.  Comparison j (the specification's order) is bit 30 - j.
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(&g[0][0]);
    constexpr int kRowWords = (kCenTW + 8) / 4;  // 32
    auto gt4 = [](uint32_t a, uint32_t b) {      // bit 7 of every byte: a > b (unsigned bytes)
        const uint32_t t = (b & 0x7F7F7F7Fu) + (~a & 0x7F7F7F7Fu);  // low 7 bits of (b + ~a): carry into bit 7 iff b7 > a7 there
        // a > b  <=>  not (b >= a).  Per byte: ge = carry of b + ~a + 1; composed from the high bits and t's bit 7
        // (lop3 0xb2 in nvcc's own expansion of __vcmpgtu4): majority(b, ~a, t) gives the carry out, i.e. b > a ... we
        // need a > b, so the operands are swapped at the call sites below
        return ((b & ~a) | ((b | ~a) & t));  // bit 7 per byte = carry out of b + ~a = (b > a)
    };
    for (int i = threadIdx.x; i < kCenTH * (kCenTW / 4); i += blockDim.x) {
        const int ty = i / (kCenTW / 4), q = i % (kCenTW / 4);
        const int tx = 4 * q;
        const int x = x0 + 4 + tx, y = y0 + 3 + ty;
        if (x >= W || y >= H) continue;
        const int cy = ty + 3;
        // words of the 7 rows: columns tx .. tx + 11 of the staged tile = image columns x - 4 .. x + 7
        uint32_t w[7][3];
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int k = 0; k < 3; ++k) w[r][k] = gw[(cy - 3 + r) * kRowWords + q + k];
        auto bytes4 = [&](int r, int dx) {  // gray values of the four pixels at row offset r - 3, column offset dx
            const int o = dx + 4;           // 0 .. 8
            const uint32_t lo = w[r][o >> 2], hi = w[r][(o >> 2) + ((o & 3) ? 1 : 0)];
            switch (o & 3) {
                case 0: return lo;
                case 1: return __byte_perm(lo, hi, 0x4321);
                case 2: return __byte_perm(lo, hi, 0x5432);
                default: return __byte_perm(lo, hi, 0x6543);
            }
        };
        uint32_t T[4] = {0, 0, 0, 0};
        // comparison j: j < 27: dy = -3 + j / 9, dx = -4 + j % 9 (rows 3 + dy against 3 - dy); else dy = 0, dx = -4 + (j - 27)
#pragma unroll
        for (int p = 0; p < 31; ++p) {  // ascending bit position p: within a byte the first processed bit ends at its bit 0
            const int j = 30 - p;
            const int dy = j < 27 ? -3 + j / 9 : 0, dx = j < 27 ? -4 + j % 9 : -4 + (j - 27);
            const uint32_t a = bytes4(3 + dy, dx), b = bytes4(3 - dy, -dx);
            const uint32_t r = gt4(b, a);  // bit 7 per byte: a > b
            T[p >> 3] = (T[p >> 3] >> 1) | (r & 0x80808080u);
        }
        T[3] >>= 1;  // the top byte only received 7 bits
        // transpose: census word of pixel k = byte k of T[0], T[1], T[2], T[3]
        const uint32_t lo01 = __byte_perm(T[0], T[1], 0x5140), hi01 = __byte_perm(T[0], T[1], 0x7362);  // (T0.0 T1.0 T0.1 T1.1), (T0.2 T1.2 T0.3 T1.3)
        const uint32_t lo23 = __byte_perm(T[2], T[3], 0x5140), hi23 = __byte_perm(T[2], T[3], 0x7362);
        const uint32_t c[4] = {__byte_perm(lo01, lo23, 0x5410), __byte_perm(lo01, lo23, 0x7632), __byte_perm(hi01, hi23, 0x5410),
                               __byte_perm(hi01, hi23, 0x7632)};
        const bool rowOk = y >= 3 && y < H - 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xk = x + k;
            if (xk >= W) break;
            out[(size_t)y * cenRowWords + xk] = (rowOk && xk >= 4 && xk < W - 4) ? c[k] : 0u;
        }
    }
}

int launch_gray_census(cartb200_ctx* c, int n, ImgBatch<const uint8_t> left, ImgBatch<const uint8_t> right,
                       cudaStream_t s) {
    dim3 grid(ceilDiv(c->W, kCenTW), ceilDiv(c->H, kCenTH), 2 * n);
    gray_census_kernel<<<grid, 256, 0, s>>>(left, right, c->grayL, c->grayPitch, c->censusL, c->censusR,
                                            c->cenRowWords, c->cenMargin, c->cfg.min_disparity, c->W, c->H);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

// ---------------------------------------------------------------------------------------------
// Path aggregation (D3 + D4).  A group of LPP = D/16 lanes owns one pixel's disparity vector; each
// lane carries 16 disparities as 8 x u16x2.  Per pixel: matching cost from the two census words
// (XOR + POPC - the POPC pipe, 16 lanes/clk/SM on B200, is the binding unit), the recurrence
//   L(d) = C(d) + min(L'(d) - m, L'(d-1) - m + P1, L'(d+1) - m + P1, P2)
// on packed halves (VIMNMX3.U16x2), a min-reduction over the group by warp shuffles and one 16-byte
// store per lane (a group writes the pixel's whole D-vector: full 128-B lines).
//   * horizontal paths: one group per image row; the right-census words a lane needs slide by one per
//     step, so a (16 + 8)-word register window is refilled with two 16-byte loads every 8 steps (the next
//     refill is prefetched);
//   * vertical paths: one group per FOUR adjacent columns (register blocking in x: 20 census words serve
//     64 cells), five 16-byte loads per row;
//   * diagonal paths (MODE_HH only): the vertical sweep over skewed columns (aggregate_diagonal_kernel); the
//     first, generic one-group-per-path-line kernel is kept for cross-checks.
struct PathArgs {
    const uint32_t* cenL;   // row layout: [margin zeros][W census words][zeros]; points at element 0 of row 0
    const uint32_t* cenR;   // right census stored shifted by min_disparity (word for (x, d) at index x - d)
    size_t cenStride;       // words per row
    size_t cenFrameStride;  // words per frame
    int cenMargin;          // words in front of column 0
    uint8_t* vol;           // this path's volume [B][H][W][D]
    uint8_t* vol2;          // volume of the opposite direction when both are fused in one launch (blockIdx.z == 1)
    size_t volFrameStride;
    int W, H, P1, P2;
    int dx, dy;
    uint32_t P1v, P2v, negP1v;  // (P1, P1), (P2, P2), (-P1, -P1) mod 2^16 as u16x2: read straight from the constant bank
    int pfPixels, pfRows;       // L2 prefetch distances of the horizontal (pixels) and vertical (rows) kernels; 0 = off
};

__device__ __forceinline__ uint32_t pack16(uint32_t lo, uint32_t hi) { return hi * 65536u + lo; }

// One DP step for a lane's 16 disparities. dp[8] holds L' (previous pixel) on entry, L on exit; m = min_k L'(k).
// With q = L' - m:  L(d) = C(d) + min(q(d), q(d-1) + P1, q(d+1) + P1, P2).
//   qp = dp + (P1 - m)                     one 32-bit add per register (halves stay independent: q + P1 < 2^16)
//   qc = min(dp - m, P2)                   VIADDMNMX.U16x2
//   neighbours across registers            PRMT; across lanes: the packed qp of the adjacent lane by shuffle
//   L  = min3(lower, upper, qc) + C        VIMNMX3.U16x2 + add
// Returns the lane-local minimum of the new L.
constexpr uint32_t kSentinel2 = 0x7FFF7FFFu;  // "no neighbour" at the ends of the disparity range

template <int LPP>
__device__ __forceinline__ uint32_t dp_step(uint32_t (&dp)[8], const uint32_t (&cost)[16], uint32_t m, int lane, uint32_t P1v,
                                            uint32_t P2v) {  // cost[2i], cost[2i+1]: matching costs of disparities 2i, 2i+1
    const uint32_t M = m * 0x10001u;
    const uint32_t K = P1v - M;                           // packed (P1 - m) as one 32-bit offset
    const uint32_t negM = ((0x10000u - m) & 0xFFFFu) * 0x10001u;  // per-half -m (mod 2^16)
    uint32_t qp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qp[i] = dp[i] + K;
    uint32_t up = __shfl_up_sync(0xFFFFFFFFu, qp[7], 1);
    uint32_t dn = __shfl_down_sync(0xFFFFFFFFu, qp[0], 1);
    if (lane == 0) up = kSentinel2;
    if (lane == LPP - 1) dn = kSentinel2;
    uint32_t sp[9];  // sp[i] = (q[2i-1] + P1, q[2i] + P1): lower neighbours of reg i, upper neighbours of reg i-1
    sp[0] = __byte_perm(up, qp[0], 0x5432);
#pragma unroll
    for (int i = 1; i < 8; ++i) sp[i] = __byte_perm(qp[i - 1], qp[i], 0x5432);
    sp[8] = __byte_perm(qp[7], dn, 0x5432);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t qc = __viaddmin_u16x2(dp[i], negM, P2v);
        dp[i] = __vimin3_u16x2(sp[i], sp[i + 1], qc) + cost[2 * i] + (cost[2 * i + 1] << 16);
    }
    uint32_t mn = __vimin3_u16x2(dp[0], dp[1], dp[2]);
    mn = __vimin3_u16x2(mn, dp[3], dp[4]);
    mn = __vimin3_u16x2(mn, dp[5], dp[6]);
    mn = __vminu2(mn, dp[7]);
    return min(mn & 0xFFFFu, mn >> 16);
}

// Leaner formulation used by the axis-aligned kernels.
//  * Register i of a lane holds the pair (d_i, d_{i+8}) of its 16 disparities, so the d-1 / d+1 neighbours of register
//    i are simply registers i-1 / i+1; only the two ends need a PRMT with the adjacent lane's value (2 instead of 9).
//  * The state carried between pixels is K = (P1 - m, P1 - m) (one true 32-bit value, halves independent): qp = dp + K
//    is the P1-penalised neighbour value and, through qc = min(qp - P1, P2) with the constant (-P1, -P1), the centre
//    term - no per-pixel negation of m.
//  * The end-of-range sentinels are OR-masks computed once per thread; the lane minimum is returned replicated in both
//    halves so the group reduction stays packed and its result is (m, m), the next K's subtrahend.
constexpr uint32_t kSentinelHi = 0x7FFF0000u, kSentinelLo = 0x00007FFFu;

template <int LPP>
__device__ __forceinline__ uint32_t dp_step2(uint32_t (&dp)[8], const uint32_t (&cost)[16], uint32_t K, uint32_t maskUp,
                                             uint32_t maskDn, uint32_t negP1v, uint32_t P2v) {
    uint32_t qp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qp[i] = dp[i] + K;
    // previous lane's (d7, d15): its d15 is this lane's d-1 neighbour of d0; next lane's (d0, d8): its d0 follows d15
    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, qp[7], 1) | maskUp;    // maskUp = kSentinelHi on the first lane
    const uint32_t dn = __shfl_down_sync(0xFFFFFFFFu, qp[0], 1) | maskDn;  // maskDn = kSentinelLo on the last lane
    const uint32_t lower0 = __byte_perm(up, qp[7], 0x5432);  // (d-1, d7): lower neighbours of register 0 = (d0, d8)
    const uint32_t upper7 = __byte_perm(qp[0], dn, 0x5432);  // (d8, d16): upper neighbours of register 7 = (d7, d15)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t qc = __viaddmin_u16x2(qp[i], negP1v, P2v);
        dp[i] = __vimin3_u16x2(i == 0 ? lower0 : qp[i - 1], i == 7 ? upper7 : qp[i + 1], qc) + cost[i] + (cost[i + 8] << 16);
    }
    uint32_t mn = __vimin3_u16x2(dp[0], dp[1], dp[2]);
    mn = __vimin3_u16x2(mn, dp[3], dp[4]);
    mn = __vimin3_u16x2(mn, dp[5], dp[6]);
    mn = __vminu2(mn, dp[7]);
    return __vminu2(mn, __byte_perm(mn, 0, 0x1032));  // (min, min)
}

// registers (d_i, d_{i+8}) -> 16 bytes d0 .. d15
__device__ __forceinline__ void store_dp2(uint8_t* dst, const uint32_t (&dp)[8]) {
    const uint32_t A = __byte_perm(dp[0], dp[1], 0x6420);  // d0 d8 d1 d9
    const uint32_t B = __byte_perm(dp[2], dp[3], 0x6420);  // d2 d10 d3 d11
    const uint32_t C = __byte_perm(dp[4], dp[5], 0x6420);
    const uint32_t E = __byte_perm(dp[6], dp[7], 0x6420);
    uint4 o;
    o.x = __byte_perm(A, B, 0x6420);
    o.y = __byte_perm(C, E, 0x6420);
    o.z = __byte_perm(A, B, 0x7531);
    o.w = __byte_perm(C, E, 0x7531);
    __stcs(reinterpret_cast<uint4*>(dst), o);  // streaming store: the volume is not re-read before the WTA pass
}

// The same step with registers in natural order (d_2i, d_2i+1): 9 neighbour PRMTs, 4 for the store.  The vertical kernel
// keeps this one (measured: the interleaved step raises it from 80 to 96 registers and costs it a resident CTA).
template <int LPP>
__device__ __forceinline__ uint32_t dp_step2n(uint32_t (&dp)[8], const uint32_t (&cost)[16], uint32_t K, uint32_t maskUp,
                                              uint32_t maskDn, uint32_t negP1v, uint32_t P2v) {
    uint32_t qp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qp[i] = dp[i] + K;
    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, qp[7], 1) | maskUp;    // maskUp / maskDn = kSentinel2 at the range ends
    const uint32_t dn = __shfl_down_sync(0xFFFFFFFFu, qp[0], 1) | maskDn;
    uint32_t sp[9];
    sp[0] = __byte_perm(up, qp[0], 0x5432);
#pragma unroll
    for (int i = 1; i < 8; ++i) sp[i] = __byte_perm(qp[i - 1], qp[i], 0x5432);
    sp[8] = __byte_perm(qp[7], dn, 0x5432);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t qc = __viaddmin_u16x2(qp[i], negP1v, P2v);
        dp[i] = __vimin3_u16x2(sp[i], sp[i + 1], qc) + cost[2 * i] + (cost[2 * i + 1] << 16);
    }
    uint32_t mn = __vimin3_u16x2(dp[0], dp[1], dp[2]);
    mn = __vimin3_u16x2(mn, dp[3], dp[4]);
    mn = __vimin3_u16x2(mn, dp[5], dp[6]);
    mn = __vminu2(mn, dp[7]);
    return __vminu2(mn, __byte_perm(mn, 0, 0x1032));  // (min, min)
}

template <int LPP>
__device__ __forceinline__ uint32_t group_min2(uint32_t v) {  // packed u16x2 minimum over the group's lanes
    // (four full-warp REDUX.MIN with per-group masks instead of the shuffle tree were measured: 5 % slower)
#pragma unroll
    for (int o = 1; o < LPP; o <<= 1) v = __vminu2(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
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
    __stcs(reinterpret_cast<uint4*>(dst), o);  // streaming store: the volume is not re-read before the WTA pass
}

template <int N>
__device__ __forceinline__ void load_words(uint32_t* dst, const uint32_t* src) {  // N consecutive words, 16-byte aligned
#pragma unroll
    for (int k = 0; k < N / 4; ++k) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + k);
        dst[4 * k] = v.x;
        dst[4 * k + 1] = v.y;
        dst[4 * k + 2] = v.z;
        dst[4 * k + 3] = v.w;
    }
}

// ---- horizontal -----------------------------------------------------------------------------------
// U = pixels per register-window refill (census rows are consumed in 16-byte pieces): the window holds the
// 16 + U right-census words a lane needs for U consecutive pixels.
constexpr int kHorizThreads = 128;
constexpr int kHorizU = 8;
template <int D, int DX, int U, bool PF>
__device__ __forceinline__ void horizontal_body(const PathArgs& a, uint8_t* __restrict__ volBase) {
    constexpr int LPP = D / 16;
    constexpr int GPB = kHorizThreads / LPP;
    const int lane = threadIdx.x % LPP;
    const int group = threadIdx.x / LPP;
    const int f = blockIdx.y;
    const int W = a.W;
    const int line = blockIdx.x * GPB + group;
    const bool valid = line < a.H;
    const int y = valid ? line : a.H - 1;
    const uint32_t* cl = a.cenL + (size_t)f * a.cenFrameStride + (size_t)y * a.cenStride + a.cenMargin;
    const uint32_t* cr = a.cenR + (size_t)f * a.cenFrameStride + (size_t)y * a.cenStride + a.cenMargin - 16 * lane - 16;
    uint8_t* vrow = volBase + (size_t)f * a.volFrameStride + (size_t)y * W * D + 16 * lane;
    const uint32_t maskUp = lane == 0 ? kSentinelHi : 0u, maskDn = lane == LPP - 1 ? kSentinelLo : 0u;
    const int nChunks = (W + U - 1) / U;
    // keep the three packed constants in registers (the compiler otherwise re-reads them from the constant bank per pixel)
    uint32_t P1v, P2v, negP1v;
    asm volatile("mov.u32 %0, %1;" : "=r"(P1v) : "r"(a.P1v));
    asm volatile("mov.u32 %0, %1;" : "=r"(P2v) : "r"(a.P2v));
    asm volatile("mov.u32 %0, %1;" : "=r"(negP1v) : "r"(a.negP1v));

    uint32_t dp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dp[i] = 0;
    uint32_t K = P1v;  // m = 0
    // S[k] = shifted right census word (x0 - 16*lane - 16 + k) of the current U-pixel chunk starting at x0
    uint32_t S[16 + U];
    uint32_t Lw[U], Ln[U], Sn[U];  // current chunk's left words; next chunk's left / right words (prefetched)
    {
        const int x0 = DX > 0 ? 0 : U * (nChunks - 1);
        load_words<16>(DX > 0 ? S : S + U, cr + x0 + (DX > 0 ? 0 : U));
        if (PF) {
            load_words<U>(Ln, cl + x0);
            load_words<U>(Sn, cr + x0 + (DX > 0 ? 16 : 0));
        }
    }
    for (int c = 0; c < nChunks; ++c) {
        const int x0 = DX > 0 ? U * c : U * (nChunks - 1 - c);
        if (PF) {
#pragma unroll
            for (int k = 0; k < U; ++k) {
                Lw[k] = Ln[k];
                (DX > 0 ? S + 16 : S)[k] = Sn[k];
            }
            // prefetch the next chunk (the zero margins make one chunk past either end readable)
            const int xn = DX > 0 ? x0 + U : x0 - U;
            load_words<U>(Ln, cl + xn);
            load_words<U>(Sn, cr + xn + (DX > 0 ? 16 : 0));
            if (a.pfPixels > 0) {  // and pull a chunk further ahead towards L2: the register prefetch alone leaves
                                   // long-scoreboard stalls (15 % of the stall samples without it)
                const int xp = DX > 0 ? x0 + a.pfPixels : x0 - a.pfPixels;
                if (DX > 0 ? xp < W : xp >= 0) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(cl + xp));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(cr + xp + (DX > 0 ? 16 : 0)));
                }
            }
        } else {
            load_words<U>(Lw, cl + x0);
            load_words<U>(DX > 0 ? S + 16 : S, cr + x0 + (DX > 0 ? 16 : 0));
        }
        uint8_t* vchunk = vrow + (size_t)x0 * D;  // stores of the chunk use immediate offsets from here
        auto step = [&](int sidx) {
            uint32_t cost[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) cost[k] = __popc(Lw[sidx] ^ S[16 + sidx - k]);
            K = P1v - group_min2<LPP>(dp_step2<LPP>(dp, cost, K, maskUp, maskDn, negP1v, P2v));
            // groups past the last image row recompute row H-1 and store the identical bytes (no predicate in the loop)
            store_dp2(vchunk + sidx * D, dp);
        };
        if (x0 + U <= W) {  // full chunk (all but the one that holds the right image border): no per-pixel bound test
#pragma unroll
            for (int t = 0; t < U; ++t) step(DX > 0 ? t : U - 1 - t);
        } else {
#pragma unroll
            for (int t = 0; t < U; ++t) {
                const int sidx = DX > 0 ? t : U - 1 - t;
                if (x0 + sidx < W) step(sidx);  // warp-uniform: every group of the warp is at the same column
            }
        }
        if (DX > 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k) S[k] = S[k + U];
        } else {
#pragma unroll
            for (int k = 15; k >= 0; --k) S[k + U] = S[k];
        }
    }
}

template <int D, int U, bool PF, int MINB>
__global__ void __launch_bounds__(kHorizThreads, MINB) aggregate_horizontal_kernel(PathArgs a, int dirFirst, int both) {
    // blockIdx.z selects the direction when both are fused in one launch
    const int dir = both ? (blockIdx.z == 0 ? 1 : -1) : dirFirst;
    uint8_t* vol = (both && blockIdx.z == 1) ? a.vol2 : a.vol;
    if (dir > 0)
        horizontal_body<D, 1, U, PF>(a, vol);
    else
        horizontal_body<D, -1, U, PF>(a, vol);
}

// ---- horizontal, 32 disparities per lane -------------------------------------------------------------
// Experiment (CARTB200_HORIZ_VARIANT=6): a lane owns 32 disparities as 16 x u16x2 in the interleaved order (d_i, d_i+16),
// a pixel is spread over D/32 lanes: 2 neighbour + log2(D/32) group-minimum shuffles per 32 cells instead of 2 + log2(D/16)
// per 16 (D = 128: 2 instead of 5 per 16 cells) - POPC and SHFL share the MIO queue.  64-thread CTAs.  Measured on B200
// (64-frame batch, bit-identical volumes): 1.32 ms per path at 80 registers / 24 warps per SM, 1.37-1.43 ms at 96-128
// registers / 16-20 warps, against 1.17 ms for the 16-per-lane kernel at 64 registers / 32 warps: the resident warps the
// wider lane costs matter more than the shuffles it saves.  Kept for reproduction, not used.
template <int NR, int LPP>
__device__ __forceinline__ uint32_t dp_step_w(uint32_t (&dp)[NR], const uint32_t (&cost)[2 * NR], uint32_t K, uint32_t maskUp,
                                              uint32_t maskDn, uint32_t negP1v, uint32_t P2v) {
    uint32_t qp[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) qp[i] = dp[i] + K;
    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, qp[NR - 1], 1) | maskUp;
    const uint32_t dn = __shfl_down_sync(0xFFFFFFFFu, qp[0], 1) | maskDn;
    const uint32_t lower0 = __byte_perm(up, qp[NR - 1], 0x5432);
    const uint32_t upperL = __byte_perm(qp[0], dn, 0x5432);
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        const uint32_t qc = __viaddmin_u16x2(qp[i], negP1v, P2v);
        dp[i] = __vimin3_u16x2(i == 0 ? lower0 : qp[i - 1], i == NR - 1 ? upperL : qp[i + 1], qc) + cost[i] + (cost[i + NR] << 16);
    }
    uint32_t mn = dp[0];
#pragma unroll
    for (int i = 1; i + 1 < NR; i += 2) mn = __vimin3_u16x2(mn, dp[i], dp[i + 1]);
    if ((NR & 1) == 0) mn = __vminu2(mn, dp[NR - 1]);
    return __vminu2(mn, __byte_perm(mn, 0, 0x1032));  // (min, min)
}

constexpr int kHoriz32Threads = 64;
template <int D, int DX, int U>
__device__ __forceinline__ void horizontal_body32(const PathArgs& a, uint8_t* __restrict__ volBase) {
    constexpr int DL = 32, NR = 16;
    constexpr int LPP = D / DL;
    constexpr int GPB = kHoriz32Threads / LPP;
    const int lane = threadIdx.x % LPP;
    const int group = threadIdx.x / LPP;
    const int f = blockIdx.y;
    const int W = a.W;
    const int line = blockIdx.x * GPB + group;
    const int y = line < a.H ? line : a.H - 1;
    const uint32_t* cl = a.cenL + (size_t)f * a.cenFrameStride + (size_t)y * a.cenStride + a.cenMargin;
    const uint32_t* cr = a.cenR + (size_t)f * a.cenFrameStride + (size_t)y * a.cenStride + a.cenMargin - DL * lane - DL;
    uint8_t* vrow = volBase + (size_t)f * a.volFrameStride + (size_t)y * W * D + DL * lane;
    const uint32_t maskUp = lane == 0 ? kSentinelHi : 0u, maskDn = lane == LPP - 1 ? kSentinelLo : 0u;
    const int nChunks = (W + U - 1) / U;
    uint32_t P1v, P2v, negP1v;
    asm volatile("mov.u32 %0, %1;" : "=r"(P1v) : "r"(a.P1v));
    asm volatile("mov.u32 %0, %1;" : "=r"(P2v) : "r"(a.P2v));
    asm volatile("mov.u32 %0, %1;" : "=r"(negP1v) : "r"(a.negP1v));
    uint32_t dp[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) dp[i] = 0;
    uint32_t K = P1v;
    uint32_t S[DL + U];  // S[k] = shifted right census word (x0 - DL*lane - DL + k)
    uint32_t Lw[U];
    {
        const int x0 = DX > 0 ? 0 : U * (nChunks - 1);
        load_words<DL>(DX > 0 ? S : S + U, cr + x0 + (DX > 0 ? 0 : U));
    }
    for (int c = 0; c < nChunks; ++c) {
        const int x0 = DX > 0 ? U * c : U * (nChunks - 1 - c);
        load_words<U>(Lw, cl + x0);
        load_words<U>(DX > 0 ? S + DL : S, cr + x0 + (DX > 0 ? DL : 0));
        if (a.pfPixels > 0) {
            const int xp = DX > 0 ? x0 + a.pfPixels : x0 - a.pfPixels;
            if (DX > 0 ? xp < W : xp >= 0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(cl + xp));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(cr + xp + (DX > 0 ? DL : 0)));
            }
        }
        uint8_t* vchunk = vrow + (size_t)x0 * D;
        auto step = [&](int sidx) {
            uint32_t cost[DL];
#pragma unroll
            for (int k = 0; k < DL; ++k) cost[k] = __popc(Lw[sidx] ^ S[DL + sidx - k]);
            K = P1v - group_min2<LPP>(dp_step_w<NR, LPP>(dp, cost, K, maskUp, maskDn, negP1v, P2v));
            // registers (d_i, d_i+16) -> 32 bytes d0 .. d31 (two 16-byte stores)
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t A = __byte_perm(dp[4 * g], dp[4 * g + 1], 0x6420);      // d d+16 d+1 d+17
                const uint32_t B = __byte_perm(dp[4 * g + 2], dp[4 * g + 3], 0x6420);  // d+2 d+18 d+3 d+19
                lo[g] = __byte_perm(A, B, 0x6420);
                hi[g] = __byte_perm(A, B, 0x7531);
            }
            uint8_t* dst = vchunk + sidx * D;
            __stcs(reinterpret_cast<uint4*>(dst), make_uint4(lo[0], lo[1], lo[2], lo[3]));
            __stcs(reinterpret_cast<uint4*>(dst + 16), make_uint4(hi[0], hi[1], hi[2], hi[3]));
        };
        if (x0 + U <= W) {
#pragma unroll
            for (int t = 0; t < U; ++t) step(DX > 0 ? t : U - 1 - t);
        } else {
#pragma unroll
            for (int t = 0; t < U; ++t) {
                const int sidx = DX > 0 ? t : U - 1 - t;
                if (x0 + sidx < W) step(sidx);
            }
        }
        if (DX > 0) {
#pragma unroll
            for (int k = 0; k < DL; ++k) S[k] = S[k + U];
        } else {
#pragma unroll
            for (int k = DL - 1; k >= 0; --k) S[k + U] = S[k];
        }
    }
}

template <int D, int U, int MINB>
__global__ void __launch_bounds__(kHoriz32Threads, MINB) aggregate_horizontal32_kernel(PathArgs a, int dirFirst, int both) {
    const int dir = both ? (blockIdx.z == 0 ? 1 : -1) : dirFirst;
    uint8_t* vol = (both && blockIdx.z == 1) ? a.vol2 : a.vol;
    if (dir > 0)
        horizontal_body32<D, 1, U>(a, vol);
    else
        horizontal_body32<D, -1, U>(a, vol);
}

// ---- vertical -------------------------------------------------------------------------------------
template <int D, int MINB>
__global__ void __launch_bounds__(128, MINB) aggregate_vertical_kernel(PathArgs a, int dirFirst, int both) {
    constexpr int LPP = D / 16;
    constexpr int GPB = 128 / LPP;
    const int dir = both ? (blockIdx.z == 0 ? 1 : -1) : dirFirst;
    uint8_t* volBase = (both && blockIdx.z == 1) ? a.vol2 : a.vol;
    const int lane = threadIdx.x % LPP;
    const int group = threadIdx.x / LPP;
    const int f = blockIdx.y;
    const int W = a.W, H = a.H;
    const int nQuads = (W + 3) >> 2;
    int quad = blockIdx.x * GPB + group;
    const bool valid = quad < nQuads;
    if (!valid) quad = nQuads - 1;
    const int x0 = 4 * quad;
    const uint32_t* clBase = a.cenL + (size_t)f * a.cenFrameStride + a.cenMargin + x0;
    const uint32_t* crBase = a.cenR + (size_t)f * a.cenFrameStride + a.cenMargin + x0 - 16 * lane - 16;
    uint8_t* vbase = volBase + (size_t)f * a.volFrameStride + (size_t)x0 * D + 16 * lane;
    const uint32_t maskUp = lane == 0 ? kSentinel2 : 0u, maskDn = lane == LPP - 1 ? kSentinel2 : 0u;
    bool st[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) st[c] = valid && x0 + c < W;

    uint32_t dp[4][8];
    uint32_t K[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        K[c] = a.P1v;  // m = 0
#pragma unroll
        for (int i = 0; i < 8; ++i) dp[c][i] = 0;
    }
    uint8_t* vp = vbase + (dir > 0 ? (size_t)0 : (size_t)(H - 1) * W * D);  // row pointer, advanced by one image row per step
    const ptrdiff_t vstep = dir > 0 ? (ptrdiff_t)W * D : -(ptrdiff_t)W * D;
    for (int step = 0; step < H; ++step, vp += vstep) {
        const int y = dir > 0 ? step : H - 1 - step;
        if (a.pfRows > 0) {  // pull the census rows a few steps ahead towards L2 (the loads below are consumed immediately)
            const int yp = dir > 0 ? y + a.pfRows : y - a.pfRows;
            if (yp >= 0 && yp < H) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(clBase + (size_t)yp * a.cenStride));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(crBase + (size_t)yp * a.cenStride));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(crBase + (size_t)yp * a.cenStride + 19));
            }
        }
        const uint4 lw = __ldg(reinterpret_cast<const uint4*>(clBase + (size_t)y * a.cenStride));
        const uint32_t Lw[4] = {lw.x, lw.y, lw.z, lw.w};
        // S[k] = shifted right census word (x0 - 16*lane - 16 + k), k = 0..19; cell (x0+c, j) uses S[16 + c - j]
        // (prefetching the next row in registers was measured: 119 registers, 20 % slower - four independent
        // columns per lane already provide the instruction-level parallelism, occupancy matters more here)
        uint32_t S[20];
        {
            const uint4* src = reinterpret_cast<const uint4*>(crBase + (size_t)y * a.cenStride);
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const uint4 v = __ldg(src + k);
                S[4 * k] = v.x;
                S[4 * k + 1] = v.y;
                S[4 * k + 2] = v.z;
                S[4 * k + 3] = v.w;
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t cost[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) cost[k] = __popc(Lw[c] ^ S[16 + c - k]);
            K[c] = a.P1v - group_min2<LPP>(dp_step2n<LPP>(dp[c], cost, K[c], maskUp, maskDn, a.negP1v, a.P2v));
            if (st[c]) store_dp(vp + c * D, dp[c]);
        }
    }
}

// ---- generic (diagonals) --------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128) aggregate_path_kernel(PathArgs a) {
    constexpr int LPP = D / 16;
    constexpr int GPB = 128 / LPP;  // groups per block
    const int lane = threadIdx.x % LPP;
    const int group = threadIdx.x / LPP;
    const int f = blockIdx.y;
    const int W = a.W, H = a.H;
    const int dx = a.dx, dy = a.dy;
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
    int maxLen = len;  // steps must be warp-uniform for the shuffles
#pragma unroll
    for (int o = LPP; o < 32; o <<= 1) maxLen = max(maxLen, __shfl_xor_sync(0xFFFFFFFFu, maxLen, o));

    const uint32_t* cl = a.cenL + (size_t)f * a.cenFrameStride + a.cenMargin;
    const uint32_t* cr = a.cenR + (size_t)f * a.cenFrameStride + a.cenMargin - 16 * lane;
    uint8_t* vol = a.vol + (size_t)f * a.volFrameStride;
    const uint32_t P1v = pack16(a.P1, a.P1), P2v = pack16(a.P2, a.P2);
    uint32_t dp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dp[i] = 0;
    uint32_t m = 0;
    for (int step = 0; step < maxLen; ++step) {
        const bool act = step < len;
        uint32_t cost[16];
        if (act) {
            const uint32_t l = __ldg(cl + (size_t)y * a.cenStride + x);
            const uint32_t* crow = cr + (size_t)y * a.cenStride + x;  // margins make every index valid
#pragma unroll
            for (int k = 0; k < 16; ++k) cost[k] = __popc(l ^ __ldg(crow - k));
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) cost[k] = 0;
        }
        const uint32_t lm = group_min<LPP>(dp_step<LPP>(dp, cost, m, lane, P1v, P2v));
        if (act) {
            m = lm;
            store_dp(vol + ((size_t)y * W + x) * D + 16 * lane, dp);
            x += dx;
            y += dy;
        }
    }
}

// ---- diagonals (MODE_HH only) -------------------------------------------------------------------------
// A diagonal path (dx, dy = +-1) is vertical in the skewed coordinate u = x - dx * step (step = rows walked
// from the starting edge): the kernel is the 4-columns-per-group vertical sweep over the W + H - 1 skewed
// columns.  u0 and the row loop are multiples of 4, so inside a 4-fold unrolled row loop the position of the
// group's first pixel inside a 16-byte aligned census window is static: census rows are still read with
// aligned 16-byte loads (2 + 6 per row instead of 1 + 5).  Pixels outside the image get cost 0 through the
// third input of the XOR's LOP3, which keeps the path state at zero until the path enters the image
// (oracle/sgm.cpp D4: diagonal paths start where they enter the image); only the stores are predicated.
template <int D, int DX>
__device__ __forceinline__ void diagonal_body(const PathArgs& a, int dy, uint8_t* __restrict__ volBase) {
    constexpr int LPP = D / 16;
    constexpr int GPB = 128 / LPP;
    const int lane = threadIdx.x % LPP;
    const int group = threadIdx.x / LPP;
    const int f = blockIdx.y;
    const int W = a.W, H = a.H;
    const int off4 = DX > 0 ? ((H - 1 + 3) & ~3) : 0;  // u = x - DX * step spans [-(H-1), W-1] (DX > 0) or [0, W+H-2]
    const int nQuads = (W + off4 + (DX < 0 ? H - 1 : 0) + 3) >> 2;
    int quad = blockIdx.x * GPB + group;
    const bool valid = quad < nQuads;
    if (!valid) quad = nQuads - 1;
    const int u0 = 4 * quad - off4;
    const uint32_t* clF = a.cenL + (size_t)f * a.cenFrameStride + a.cenMargin;
    const uint32_t* crF = a.cenR + (size_t)f * a.cenFrameStride + a.cenMargin - 16 * lane - 16;
    uint8_t* vF = volBase + (size_t)f * a.volFrameStride + 16 * lane;
    const uint32_t maskUp = lane == 0 ? kSentinel2 : 0u, maskDn = lane == LPP - 1 ? kSentinel2 : 0u;

    uint32_t dp[4][8];
    uint32_t K[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        K[c] = a.P1v;  // m = 0
#pragma unroll
        for (int i = 0; i < 8; ++i) dp[c][i] = 0;
    }
    // Steps at which at least one of the warp's columns is inside the image: x = u0 + c + DX * step in [0, W).  Before a
    // path enters the image its state is zero (costs are masked) and after it has left nothing is stored, so the steps
    // outside this range are skipped: the sweep costs W * H pixel steps instead of (W + H) * H (the corner triangles).
    int lo = DX > 0 ? -u0 - 3 : u0 - (W - 1), hi = DX > 0 ? W - 1 - u0 : u0 + 3;
    lo = __reduce_min_sync(0xFFFFFFFFu, max(lo, 0));
    hi = __reduce_max_sync(0xFFFFFFFFu, min(hi, H - 1));
    for (int s4 = lo & ~3; s4 <= hi; s4 += 4) {
        const int base4 = u0 + DX * s4;  // multiple of 4
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int step = s4 + r;
            if (step >= H) break;  // uniform
            const int y = dy > 0 ? step : H - 1 - step;
            const int x0 = base4 + DX * r;                                   // pixel of column 0
            const int o = DX > 0 ? r : (r == 0 ? 0 : 4 - r);                 // its position in the aligned window (static)
            int wb = DX > 0 ? base4 : (r == 0 ? base4 : base4 - 4);          // aligned window start
            const bool any = x0 + 3 >= 0 && x0 < W;
            // groups that are completely outside read the zero margin in front of the row (costs are masked anyway)
            if (!any) wb = -8;
            const uint32_t* clRow = clF + (size_t)y * a.cenStride + wb;
            const uint32_t* crRow = crF + (size_t)y * a.cenStride + wb;
            uint32_t Lq[8], Sq[24];
            load_words<8>(Lq, clRow);
            load_words<24>(Sq, crRow);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int x = x0 + c;
                const bool act = valid && x >= 0 && x < W;
                const uint32_t mask = act ? 0xFFFFFFFFu : 0u;
                uint32_t cost[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) cost[k] = __popc((Lq[o + c] ^ Sq[o + 16 + c - k]) & mask);
                K[c] = a.P1v - group_min2<LPP>(dp_step2n<LPP>(dp[c], cost, K[c], maskUp, maskDn, a.negP1v, a.P2v));
                if (act) store_dp(vF + ((size_t)y * W + x) * D, dp[c]);
            }
        }
    }
}

// blockIdx.z selects one of up to four diagonal paths (paths 4..7: (1,1), (-1,1), (1,-1), (-1,-1)) of one launch
template <int D>
__global__ void __launch_bounds__(128) aggregate_diagonal_kernel(PathArgs a, int firstPath, size_t volPathStride) {
    const int p = firstPath + blockIdx.z;  // 4..7
    const int dx = (p & 1) ? -1 : 1, dy = (p & 2) ? -1 : 1;
    uint8_t* vol = a.vol + (size_t)blockIdx.z * volPathStride;
    if (dx > 0)
        diagonal_body<D, 1>(a, dy, vol);
    else
        diagonal_body<D, -1>(a, dy, vol);
}

static const int kDirs[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {-1, 1}, {1, -1}, {-1, -1}};

// st[0]: horizontal paths, st[1]: vertical paths, st[2]: diagonals
template <int D>
static void launch_paths_D(cartb200_ctx* c, PathArgs a, int n, int p0, int p1, cudaStream_t const (&st)[3]) {
    constexpr int GPB = 128 / (D / 16);
    for (int p = p0; p < p1; ++p) {
        cudaStream_t s = kDirs[p][1] == 0 ? st[0] : (kDirs[p][0] == 0 ? st[1] : st[2]);
        a.vol = c->volumes + (size_t)p * c->volPathStride;
        a.dx = kDirs[p][0];
        a.dy = kDirs[p][1];
        // opposite directions of an axis-aligned pair share one launch (twice the resident warps)
        const bool pair = (p == 0 || p == 2) && p + 1 < p1;
        a.vol2 = pair ? c->volumes + (size_t)(p + 1) * c->volPathStride : nullptr;
        if (a.dy == 0) {
            dim3 grid(ceilDiv(a.H, kHorizThreads / (D / 16)), n, pair ? 2 : 1);
            // measured on B200 (64-frame batch, profiles/r01n_aggregate_variant_sweep.txt): with the census rows
            // prefetched into L2, the register prefetch of the next chunk is not needed any more and 64 registers /
            // 8 CTAs per SM beat the 96-register prefetching variant by 2.7 %.  CARTB200_HORIZ_VARIANT (tuning aid)
            // selects the other measured shapes (profiles/r02v_aggregate_variant_sweep.json): 1, 2 = the older ones,
            // 6 = 32 disparities per lane (80 registers, 24 warps/SM: 12 % slower), 7 = 4-pixel window refills at
            // 56 registers / 9 CTAs (3 % slower).
            const int variant = getenv("CARTB200_HORIZ_VARIANT") ? atoi(getenv("CARTB200_HORIZ_VARIANT")) : 0;
            if (variant == 6) {  // 32 disparities per lane (measured and rejected, see horizontal_body32)
                dim3 g32(ceilDiv(a.H, kHoriz32Threads / (D / 32)), n, pair ? 2 : 1);
                aggregate_horizontal32_kernel<D, 4, 12><<<g32, kHoriz32Threads, 0, s>>>(a, a.dx, pair ? 1 : 0);
            } else
            switch (variant) {  // (pixels per window refill, register prefetch, minimum CTAs per SM)
                case 1: aggregate_horizontal_kernel<D, 8, true, 5><<<grid, kHorizThreads, 0, s>>>(a, a.dx, pair ? 1 : 0); break;
                case 2: aggregate_horizontal_kernel<D, 8, false, 6><<<grid, kHorizThreads, 0, s>>>(a, a.dx, pair ? 1 : 0); break;
                case 7: aggregate_horizontal_kernel<D, 4, false, 9><<<grid, kHorizThreads, 0, s>>>(a, a.dx, pair ? 1 : 0); break;
                default: aggregate_horizontal_kernel<D, kHorizU, false, 8><<<grid, kHorizThreads, 0, s>>>(a, a.dx, pair ? 1 : 0);
            }
        } else if (a.dx == 0) {
            dim3 grid(ceilDiv(ceilDiv(a.W, 4), GPB), n, pair ? 2 : 1);
            aggregate_vertical_kernel<D, 6><<<grid, 128, 0, s>>>(a, a.dy, pair ? 1 : 0);  // 7 CTAs/SM: same time, 8: spills
        } else if (getenv("CARTB200_GENERIC_DIAGONALS")) {  // the first, generic formulation (kept for cross-checks)
            dim3 grid(ceilDiv(a.W + a.H - 1, GPB), n);
            aggregate_path_kernel<D><<<grid, 128, 0, s>>>(a);
        } else {
            // all remaining diagonals of the requested range share one launch
            const int nz = p1 - p;
            const int nQuads = (a.W + ((a.H - 1 + 3) & ~3) + 3) >> 2;  // the larger of the two skew directions
            dim3 grid(ceilDiv(nQuads, GPB), n, nz);
            aggregate_diagonal_kernel<D><<<grid, 128, 0, s>>>(a, p, c->volPathStride);
            c->launches++;
            break;
        }
        c->launches++;
        if (pair) ++p;
    }
}

int launch_aggregate_range(cartb200_ctx* c, int n, int p0, int p1, cudaStream_t s);

int launch_aggregate(cartb200_ctx* c, int n, cudaStream_t s) { return launch_aggregate_range(c, n, 0, c->P, s); }

int launch_aggregate_range(cartb200_ctx* c, int n, int p0, int p1, cudaStream_t s) {
    PathArgs a;
    a.cenL = c->censusL;
    a.cenR = c->censusR;
    a.cenStride = c->cenRowWords;
    a.cenFrameStride = c->cenRowWords * c->H;
    a.cenMargin = c->cenMargin;
    a.vol = a.vol2 = nullptr;
    a.volFrameStride = c->volFrameStride;
    a.W = c->W;
    a.H = c->H;
    a.P1 = c->cfg.p1;
    a.P2 = c->cfg.p2;
    a.dx = a.dy = 0;
    a.P1v = (uint32_t)a.P1 * 0x10001u;
    a.P2v = (uint32_t)a.P2 * 0x10001u;
    a.negP1v = ((0x10000u - (uint32_t)a.P1) & 0xFFFFu) * 0x10001u;
    // CARTB200_PF_PIXELS / CARTB200_PF_ROWS (tuning aids): L2 prefetch distances, measured optimum as default
    static const int pfPixels = getenv("CARTB200_PF_PIXELS") ? atoi(getenv("CARTB200_PF_PIXELS")) : 32;
    static const int pfRows = getenv("CARTB200_PF_ROWS") ? atoi(getenv("CARTB200_PF_ROWS")) : 6;
    a.pfPixels = pfPixels;
    a.pfRows = pfRows;
    // More than one path kind: fork onto the context's auxiliary streams.  Every launch ends in a partially filled wave
    // (e.g. 3072 CTAs on 1184 slots = 2.6 waves for the horizontal pair of a 64-frame KITTI batch); CTAs of the next
    // path kind fill it instead of waiting for the launch to drain.  CARTB200_AGG_STREAMS=0 keeps one stream.
    static const bool multi = !(getenv("CARTB200_AGG_STREAMS") && atoi(getenv("CARTB200_AGG_STREAMS")) == 0);
    bool kinds[3] = {false, false, false};
    for (int p = p0; p < p1; ++p) kinds[kDirs[p][1] == 0 ? 0 : (kDirs[p][0] == 0 ? 1 : 2)] = true;
    const bool fork = multi && (int)kinds[0] + (int)kinds[1] + (int)kinds[2] > 1;
    cudaStream_t st[3] = {s, s, s};
    if (fork) {
        if (!c->aggFork) {
            for (int i = 0; i < 2; ++i) {
                CB_CHECK_CUDA(c, cudaStreamCreateWithFlags(&c->aggStream[i], cudaStreamNonBlocking));
                CB_CHECK_CUDA(c, cudaEventCreateWithFlags(&c->aggJoin[i], cudaEventDisableTiming));
            }
            CB_CHECK_CUDA(c, cudaEventCreateWithFlags(&c->aggFork, cudaEventDisableTiming));
        }
        CB_CHECK_CUDA(c, cudaEventRecord(c->aggFork, s));
        int next = 0;  // the first kind stays on the caller's stream
        bool first = true;
        for (int k = 0; k < 3; ++k) {
            if (!kinds[k]) continue;
            if (first) {
                first = false;
                continue;
            }
            st[k] = c->aggStream[next++];
            CB_CHECK_CUDA(c, cudaStreamWaitEvent(st[k], c->aggFork, 0));
        }
    }
    switch (c->D) {
        case 64: launch_paths_D<64>(c, a, n, p0, p1, st); break;
        case 128: launch_paths_D<128>(c, a, n, p0, p1, st); break;
        case 256: launch_paths_D<256>(c, a, n, p0, p1, st); break;
        default: c->err = "num_disparities must be 64, 128 or 256"; return CARTB200_E_UNSUPPORTED;
    }
    if (fork) {
        int next = 0;
        for (int k = 0; k < 3; ++k) {
            if (st[k] == s) continue;
            CB_CHECK_CUDA(c, cudaEventRecord(c->aggJoin[next], st[k]));
            CB_CHECK_CUDA(c, cudaStreamWaitEvent(s, c->aggJoin[next], 0));
            ++next;
        }
    }
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        c->err = std::string("aggregate launch: ") + cudaGetErrorString(e);
        return CARTB200_E_CUDA;
    }
    return CARTB200_OK;
}

// ---------------------------------------------------------------------------------------------
// Winner-takes-all (D5 + D6).  A group of LPP = D/16 lanes walks along a segment of one image row; each
// lane owns 16 disparities.  Per pixel the P path volumes are read once (one 16-byte load per lane and
// path: a group reads whole 128-byte lines), summed to u16x2, and turned into 32-bit keys (S << 8 | d):
//   * left image:  the two smallest keys by a min/max tournament in registers + log2(LPP) shuffle rounds,
//     uniqueness test and integer sub-pixel refinement on the group's first lane (the two neighbours of the
//     winner come from a per-group shared-memory stash of the summed vector);
//   * right image: dR(r) = argmin_d S(r + d, d) is accumulated systolically - a running minimum per
//     disparity slot that moves one slot up per pixel (RM[j] = min(RM[j-1], key[j]) in descending order: the shift is
//     free, one shuffle per step between lanes), so a right pixel leaves the last slot exactly when all its candidates
//     have been seen.  The step loop is only unrolled by the cp.async ring depth: a 16-fold unrolled body (56 KB of
//     code) made instruction fetch the kernel's first stall reason.  Results are merged into a u32 key image with atomicMin, which also joins the partial minima of
//     adjacent segments.  No shared-memory ring, no modulo arithmetic, no block barrier.
struct WtaArgs {
    const uint8_t* vol;
    size_t volPathStride, volFrameStride;
    int W, H;
    uint16_t* left;
    size_t pitch;  // elements
    uint32_t* rightKey;
    size_t rkPitch;  // elements
    float uniq;
    int nSeg, segLen, nItems;
};

constexpr uint32_t kKeyInf = 0xFFFFFFFFu;
constexpr uint32_t kKeyReal = 0x10000000u;  // real keys (S <= 255 * 8) are below; masked and empty ones above

// two smallest of 16 distinct keys: 8 compare-exchanges + 7 merges of sorted pairs
__device__ __forceinline__ void top2_of16(const uint32_t (&k)[16], uint32_t& b1, uint32_t& b2) {
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        lo[i] = min(k[2 * i], k[2 * i + 1]);
        hi[i] = max(k[2 * i], k[2 * i + 1]);
    }
#pragma unroll
    for (int n = 4; n >= 1; n >>= 1)
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const uint32_t m1 = min(lo[i], lo[i + n]);
            const uint32_t m2 = min(max(lo[i], lo[i + n]), min(hi[i], hi[i + n]));
            lo[i] = m1;
            hi[i] = m2;
        }
    b1 = lo[0];
    b2 = hi[0];
}

constexpr int kWtaDepth = 4;  // pixels in flight per lane (cp.async ring) = unroll factor of the step loop; divides 16
template <int D, int P>
__global__ void __launch_bounds__(128, 5) wta_walk_kernel(WtaArgs a) {
    constexpr int LPP = D / 16;
    constexpr int GPB = 128 / LPP;
    __shared__ __align__(16) uint16_t stash[2][GPB][D];
    extern __shared__ uint4 ring[];  // [kWtaDepth][P][128]
    const int lane = threadIdx.x % LPP, grp = threadIdx.x / LPP;
    const int item = blockIdx.x * GPB + grp;
    const bool live = item < a.nItems;
    int seg = 0, y = 0, f = 0;
    if (live) {
        seg = item % a.nSeg;
        const int row = item / a.nSeg;
        y = row % a.H;
        f = row / a.H;
    }
    const int W = a.W;
    const int x0 = seg * a.segLen;
    const int x1 = live ? min(W, x0 + a.segLen) : x0;
    const int T = a.segLen;  // multiple of 16; the same trip count for every group of the warp (shuffles inside)
    // loads are unconditional (the volume allocation is padded): steps past the segment end read valid
    // memory and are neutralised below
    const uint8_t* vp = a.vol + (size_t)f * a.volFrameStride + ((size_t)y * W + x0) * D + 16 * lane;
    uint16_t* outL = a.left + ((size_t)f * a.H + y) * a.pitch;
    uint32_t* rk = a.rightKey + ((size_t)f * a.H + y) * a.rkPitch;
    const uint32_t dbase = 16u * lane;
    // keys order by (S, d) with d < 256: one PRMT each - byte 0 = d from a constant register, bytes 1-2 = S,
    // byte 3 = the high byte of S again (keeps the order; masked off when the cost is extracted)
    uint32_t J[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) J[k] = (dbase + 4 * k) * 0x01010101u + 0x03020100u;

    uint32_t RM[16];  // running minima; slot j = disparity dbase + j
#pragma unroll
    for (int j = 0; j < 16; ++j) RM[j] = kKeyInf;
    // The volume vectors are staged through a per-thread shared-memory ring with cp.async (LDGSTS): kWtaDepth
    // pixels in flight per lane without holding them in registers.  Every thread only reads back the 16-byte
    // pieces it copied itself, so the only synchronisation is the thread's own cp.async wait.
    uint4* ringT = ring + threadIdx.x;  // slot (stage, path) at ringT[(stage * P + path) * 128]
#pragma unroll
    for (int st = 0; st < kWtaDepth; ++st) {
#pragma unroll
        for (int p = 0; p < P; ++p)
            __pipeline_memcpy_async(ringT + (st * P + p) * 128, vp + (size_t)p * a.volPathStride + st * D, 16);
        __pipeline_commit();
    }

    for (int t0 = 0; t0 < T; t0 += kWtaDepth, vp += kWtaDepth * D) {
        const bool blockActive = __all_sync(0xFFFFFFFFu, x0 + t0 + kWtaDepth <= x1);
#pragma unroll
        for (int t = 0; t < kWtaDepth; ++t) {
            const int x = x0 + t0 + t;
            const bool act = x < x1;
            __pipeline_wait_prior(kWtaDepth - 1);  // the oldest stage (this pixel) has landed
            uint4 v[P];
#pragma unroll
            for (int p = 0; p < P; ++p) v[p] = ringT[((t % kWtaDepth) * P + p) * 128];
            // ---- sum of the P path costs, u16x2 (d, d+1) in natural order ----
            uint32_t S[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) S[i] = 0;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                S[0] += __byte_perm(v[p].x, 0, 0x4140);
                S[1] += __byte_perm(v[p].x, 0, 0x4342);
                S[2] += __byte_perm(v[p].y, 0, 0x4140);
                S[3] += __byte_perm(v[p].y, 0, 0x4342);
                S[4] += __byte_perm(v[p].z, 0, 0x4140);
                S[5] += __byte_perm(v[p].z, 0, 0x4342);
                S[6] += __byte_perm(v[p].w, 0, 0x4140);
                S[7] += __byte_perm(v[p].w, 0, 0x4342);
            }
#pragma unroll
            for (int p = 0; p < P; ++p)  // refill this stage with the pixel kWtaDepth steps ahead
                __pipeline_memcpy_async(ringT + ((t % kWtaDepth) * P + p) * 128,
                                        vp + (size_t)p * a.volPathStride + (t + kWtaDepth) * D, 16);
            __pipeline_commit();
            if (!blockActive) {  // warp-uniform: only the ragged tail of a segment takes this path
                const uint32_t mask = act ? 0u : 0x7FFF7FFFu;  // inactive steps lose against every real candidate
#pragma unroll
                for (int i = 0; i < 8; ++i) S[i] |= mask;
            }
            uint4* st = reinterpret_cast<uint4*>(&stash[t & 1][grp][16 * lane]);
            st[0] = make_uint4(S[0], S[1], S[2], S[3]);
            st[1] = make_uint4(S[4], S[5], S[6], S[7]);
            uint32_t key[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                // selector nibbles (byte 0..3): constant byte (4 + j % 4), cost low byte, cost high byte, cost high byte
                key[2 * i] = __byte_perm(S[i], J[(2 * i) >> 2], 0x1104 + ((2 * i) & 3));
                key[2 * i + 1] = __byte_perm(S[i], J[(2 * i + 1) >> 2], 0x3324 + ((2 * i + 1) & 3));
            }
            // ---- left: two smallest keys of the pixel ----
            uint32_t b1, b2;
            top2_of16(key, b1, b2);
#pragma unroll
            for (int o = 1; o < LPP; o <<= 1) {
                const uint32_t o1 = __shfl_xor_sync(0xFFFFFFFFu, b1, o), o2 = __shfl_xor_sync(0xFFFFFFFFu, b2, o);
                const uint32_t m2 = min(max(b1, o1), min(b2, o2));
                b1 = min(b1, o1);
                b2 = m2;
            }
            __syncwarp();  // the stash of this step is complete
            {   // every lane holds the same (b1, b2): branch-free finish, only the group's first lane stores
                const int c1 = (int)((b1 >> 8) & 0xFFFF), d1 = (int)(b1 & 0xFF), c2 = (int)((b2 >> 8) & 0xFFFF), d2 = (int)(b2 & 0xFF);
                const bool reject = (__fmul_rn((float)c2, a.uniq) < (float)c1) && (abs(d1 - d2) > 1);
                const uint16_t* Sx = stash[t & 1][grp];
                const int l = Sx[max(d1 - 1, 0)], rr = Sx[min(d1 + 1, D - 1)];
                const int numer = l - rr, denom = l - 2 * c1 + rr;
                const bool interior = d1 > 0 && d1 < D - 1 && denom != 0;
                // trunc(((numer << 4) + denom) / (2 denom)) in fp32: |numerator| <= 21744 and the quotient's distance
                // to the next integer is >= 1 / |2 denom|, i.e. >= 4.6e-5 relative, so a reciprocal-multiply with
                // a 1e-5 relative nudge away from zero truncates exactly
                const float num = (float)((numer << 4) + denom), den = interior ? (float)(2 * denom) : 1.0f;
                const int q = interior ? __float2int_rz(__fdividef(num, den) * 1.00001f) : 0;
                if (act && lane == 0) outL[x] = reject ? (uint16_t)0xFFFF : (uint16_t)((d1 << 4) + q);
            }
            // ---- right: systolic running minima ----
            const uint32_t out = RM[15];  // slot 15 of the previous step: its right pixel moves on to the next lane
            uint32_t in = __shfl_up_sync(0xFFFFFFFFu, out, 1, LPP);
            if (lane == 0) in = kKeyInf;
            if (lane == LPP - 1 && out < kKeyReal) {
                const int r = x - 1 - (D - 1);  // complete: every candidate x' in [r, r + D) has been seen
                if (r >= 0) atomicMin(rk + r, out);
            }
#pragma unroll
            for (int j = 15; j >= 1; --j) RM[j] = min(RM[j - 1], key[j]);  // every minimum moves one slot up
            RM[0] = min(in, key[0]);
        }
    }
    // flush: slot j belongs to right pixel (x0 + T - 1) - d
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int r = x0 + T - 1 - (int)(dbase + j);
        const uint32_t val = RM[j];
        if (r >= 0 && r < W && val < kKeyReal) atomicMin(rk + r, val);
    }
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set from cartb200_create for the context's device
template <int D>
static cudaError_t wta_attributes_D() {
    cudaError_t e = cudaFuncSetAttribute(wta_walk_kernel<D, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(kWtaDepth * 4 * 128 * sizeof(uint4)));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(wta_walk_kernel<D, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(kWtaDepth * 8 * 128 * sizeof(uint4)));
}
cudaError_t sgm_set_kernel_attributes() {
    cudaError_t e;
    if ((e = wta_attributes_D<64>()) != cudaSuccess) return e;
    if ((e = wta_attributes_D<128>()) != cudaSuccess) return e;
    return wta_attributes_D<256>();
}

template <int D>
static int launch_wta_D(cartb200_ctx* c, WtaArgs& a, int n, cudaStream_t s) {
    constexpr int GPB = 128 / (D / 16);
    // segments per row: enough groups to fill the machine (~48 warps per SM), at most 8, at least D pixels long
    const long rows = (long)n * c->H;
    const long wantGroups = (long)kNumSMs * 48 * (32 / (D / 16));
    int nSeg = (int)std::min<long>(8, std::max<long>(1, (wantGroups + rows - 1) / rows));
    nSeg = std::max(1, std::min(nSeg, c->W / (2 * D)));
    a.nSeg = nSeg;
    a.segLen = (ceilDiv(c->W, nSeg) + 15) & ~15;
    a.nSeg = ceilDiv(c->W, a.segLen);
    a.nItems = (int)(rows * a.nSeg);
    CB_CHECK_CUDA(c, cudaMemsetAsync(c->wtaR, 0xFF, (size_t)n * c->H * c->rkPitch * sizeof(uint32_t), s));
    dim3 grid(ceilDiv(a.nItems, GPB));
    // (the dynamic shared memory limit of these kernels is raised per device in sgm_set_kernel_attributes)
    if (c->P == 4)
        wta_walk_kernel<D, 4><<<grid, 128, kWtaDepth * 4 * 128 * sizeof(uint4), s>>>(a);
    else
        wta_walk_kernel<D, 8><<<grid, 128, kWtaDepth * 8 * 128 * sizeof(uint4), s>>>(a);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

int launch_wta(cartb200_ctx* c, int n, cudaStream_t s) {
    WtaArgs a;
    a.vol = c->volumes;
    a.volPathStride = c->volPathStride;
    a.volFrameStride = c->volFrameStride;
    a.W = c->W;
    a.H = c->H;
    a.left = c->wtaL;
    a.pitch = c->dispPitch / 2;
    a.rightKey = c->wtaR;
    a.rkPitch = c->rkPitch;
    a.uniq = (float)(100 - c->cfg.uniqueness_ratio) / 100.0f;
    switch (c->D) {
        case 64: return launch_wta_D<64>(c, a, n, s);
        case 128: return launch_wta_D<128>(c, a, n, s);
        case 256: return launch_wta_D<256>(c, a, n, s);
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
template <typename T>  // T = uint16_t (left sub-pixel image) or uint32_t (right key image: disparity in the low byte)
__device__ __forceinline__ uint32_t median_at(const T* img, size_t pitch, int W, int H, int x, int y) {
    constexpr uint32_t mask = sizeof(T) == 2 ? 0xFFFFu : 0xFFu;
    if (x < 1 || y < 1 || x >= W - 1 || y >= H - 1) return __ldg(img + (size_t)y * pitch + x) & mask;
    uint32_t v[9];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) v[j * 3 + i] = __ldg(img + (size_t)(y + j - 1) * pitch + x + i - 1) & mask;
    return median9(v);
}

__global__ void __launch_bounds__(256) sgm_post_kernel(const uint16_t* __restrict__ wl, const uint32_t* __restrict__ wr,
                                                       size_t pitch, size_t rkPitch, const uint8_t* __restrict__ grayL, size_t grayPitch,
                                                       ImgBatch<int16_t> out, int W, int H, int minDisp) {
    const int f = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const uint16_t* L = wl + (size_t)f * H * pitch;
    const uint32_t* Rr = wr + (size_t)f * H * rkPitch;
    const uint32_t org = median_at(L, pitch, W, H, x, y);
    const int d = (int)org >> 4;
    const int k = x - d;
    bool invalid = grayL[((size_t)f * H + y) * grayPitch + x] == 0 || org == 0xFFFFu;
    if (!invalid && k >= 0 && k < W) invalid = abs((int)median_at(Rr, rkPitch, W, H, k, y) - d) > 1;
    out.frame(f).at(x, y) = invalid ? (int16_t)((minDisp - 1) * 16) : (int16_t)(uint16_t)(org + minDisp * 16);
}

__global__ void right_key_to_u16_kernel(const uint32_t* __restrict__ keys, size_t rkPitch, uint16_t* __restrict__ out,
                                        size_t pitch, int W, int rows) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < W && y < rows) out[(size_t)y * pitch + x] = (uint16_t)(keys[(size_t)y * rkPitch + x] & 0xFFu);
}

// parity/debug access: integer right disparities as a u16 image (cartb200_sgm_intermediate selector 4)
int launch_right_u16(cartb200_ctx* c, int n, uint16_t* out, cudaStream_t s) {
    dim3 grid(ceilDiv(c->W, 256), n * c->H);
    right_key_to_u16_kernel<<<grid, 256, 0, s>>>(c->wtaR, c->rkPitch, out, c->dispPitch / 2, c->W, n * c->H);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

uint32_t debug_median9_host(const uint16_t* v9) {
    uint32_t v[9];
    for (int i = 0; i < 9; ++i) v[i] = v9[i];
    return median9(v);
}

int launch_sgm_post(cartb200_ctx* c, int n, ImgBatch<int16_t> disp, cudaStream_t s) {
    dim3 grid(ceilDiv(c->W, 256), c->H, n);
    sgm_post_kernel<<<grid, 256, 0, s>>>(c->wtaL, c->wtaR, c->dispPitch / 2, c->rkPitch, c->grayL, c->grayPitch, disp, c->W, c->H,
                                         c->cfg.min_disparity);
    CB_LAUNCH_CHECK(c);
    return CARTB200_OK;
}

}  // namespace cb
