// TEST INFRASTRUCTURE ONLY (see tile.hpp header).
//
// Scalar restatement of the SGM stage that the reference obtains from a THIRD-PARTY dependency that is
// absent from /root/reference: cv::cuda::StereoSGM (opencv_contrib `cudastereo`, a port of Fixstars
// libSGM), created at /root/reference/include/modules/disparity.hpp:31-33 with
// createStereoSGM(minDisparity, numDisparities) + setUniquenessRatio(12) and invoked at
// /root/reference/src/modules/disparity/disparity.cu:71.  OpenCV is found with find_package(OpenCV
// REQUIRED) (/root/reference/CMakeLists.txt:19): NO pinned version, not vendored.
//
// *** PARITY UNPINNED *** for this stage: the reference ships no tests/golden vectors and the library
// cannot be built or run here, so this file is the normative specification (SURVEY.md Appendix A,
// resolved as below).  Decisions (each one is a recollection of the published libSGM / OpenCV code):
//   D1  gray  = (1868*B + 9617*G + 4899*R + 8192) >> 14                       (cvtColor BGR2GRAY, CUDA path)
//   D2  census: 9x7 centre-symmetric, 31 bits, border pixels (x<4, x>=W-4, y<3, y>=H-3) = 0
//   D3  cost  C(x,y,d) = popcount(cL(x,y) ^ cR(x-d-minDisp,y)), cR := 0 left of the image
//   D4  path  L(p,d) = C + min(L'(d), L'(d-1)+P1, L'(d+1)+P1, m+P2) - m, m = min_k L'(k); state before the
//       first in-image pixel of a path is all zero; diagonal paths start where they enter the image
//   D5  WTA left: two smallest packed (S<<16|d); reject iff (float)S2*u < (float)S1 and |d1-d2| > 1,
//       u = (float)(100-UR)/100; sub-pixel v = d1*16 + ((num<<4)+den)/(2*den), den==0 -> +0
//   D6  WTA right: dR(x) = argmin_d S(x+d,y,d) over x+d < W, ties -> smaller d; integer, never invalid
//   D7  3x3 median on both (u16, invalid = 0xFFFF sorts high); 1-pixel border copies the source
//   D8  L/R check: invalid iff left gray == 0, or already invalid, or (0 <= k=x-(v>>4) < W and |dR(k)-(v>>4)| > 1)
//       (k outside the image is NOT rejected - this deviates from SURVEY Appendix A6, following the
//       libSGM check_consistency kernel as recollected)
//   D9  range: invalid -> (minDisp-1)*16, valid -> v + minDisp*16
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {
inline int popc(uint32_t v) { return __builtin_popcount(v); }
}  // namespace

extern "C" {

// D1 - third-party cv::cuda::cvtColor(BGR2GRAY), call sites disparity.cu:66-67
int orc_gray(const uint8_t* bgr, int W, int H, uint8_t* gray) {
    for (long i = 0; i < (long)W * H; ++i) {
        unsigned b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        gray[i] = (uint8_t)((b * 1868u + g * 9617u + r * 4899u + 8192u) >> 14);
    }
    return 0;
}

// cv::cuda::cvtColor(BGR2YCrCb), call site superpixels.cu:82. Pinned bit-exact against CPU cv2 (tests).
int orc_ycrcb(const uint8_t* bgr, int W, int H, uint8_t* out) {
    for (long i = 0; i < (long)W * H; ++i) {
        int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        int Y = (b * 1868 + g * 9617 + r * 4899 + 8192) >> 14;
        int Cr = ((r - Y) * 11682 + (128 << 14) + 8192) >> 14;
        int Cb = ((b - Y) * 9241 + (128 << 14) + 8192) >> 14;
        out[3 * i] = (uint8_t)Y;
        out[3 * i + 1] = (uint8_t)std::min(255, std::max(0, Cr));
        out[3 * i + 2] = (uint8_t)std::min(255, std::max(0, Cb));
    }
    return 0;
}

// D2
int orc_census(const uint8_t* I, int W, int H, uint32_t* out) {
    std::memset(out, 0, sizeof(uint32_t) * (size_t)W * H);
    for (int y = 3; y < H - 3; ++y)
        for (int x = 4; x < W - 4; ++x) {
            uint32_t f = 0;
            for (int dy = -3; dy < 0; ++dy)
                for (int dx = -4; dx <= 4; ++dx)
                    f = (f << 1) | (uint32_t)(I[(size_t)(y + dy) * W + x + dx] > I[(size_t)(y - dy) * W + x - dx]);
            for (int dx = -4; dx < 0; ++dx) f = (f << 1) | (uint32_t)(I[(size_t)y * W + x + dx] > I[(size_t)y * W + x - dx]);
            out[(size_t)y * W + x] = f;
        }
    return 0;
}

// D3+D4: one aggregation path with direction (dx,dy) in {-1,0,1}^2 \ {0,0}.  L layout [y][x][d], u8.
int orc_sgm_path(const uint32_t* cl, const uint32_t* cr, int W, int H, int D, int minDisp, int P1, int P2, int dx,
                 int dy, uint8_t* L) {
    if (31 + P2 > 255) return -1;  // u8 volume would overflow
    std::vector<int> prev(D), cur(D);
    auto run = [&](int x, int y) {
        // walk from (x,y) along (dx,dy) until leaving the image
        std::fill(prev.begin(), prev.end(), 0);
        int m = 0;
        while (x >= 0 && x < W && y >= 0 && y < H) {
            const uint32_t l = cl[(size_t)y * W + x];
            int nm = 1 << 30;
            for (int d = 0; d < D; ++d) {
                const int xr = x - d - minDisp;
                const uint32_t r = (xr >= 0 && xr < W) ? cr[(size_t)y * W + xr] : 0u;
                int best = std::min(prev[d], m + P2);
                if (d > 0) best = std::min(best, prev[d - 1] + P1);
                if (d + 1 < D) best = std::min(best, prev[d + 1] + P1);
                const int v = popc(l ^ r) + best - m;
                cur[d] = v;
                nm = std::min(nm, v);
                L[((size_t)y * W + x) * D + d] = (uint8_t)v;
            }
            std::swap(prev, cur);
            m = nm;
            x += dx;
            y += dy;
        }
    };
    // every pixel whose predecessor (x-dx, y-dy) is outside the image starts a path
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int px = x - dx, py = y - dy;
            if (px < 0 || px >= W || py < 0 || py >= H) run(x, y);
        }
    return 0;
}

// D5+D6. Ls: P volumes [y][x][d].  left: u16 sub-pixel (0xFFFF invalid), right: u16 integer.
int orc_sgm_wta(const uint8_t* const* Ls, int P, int W, int H, int D, int uniquenessRatio, uint16_t* left,
                uint16_t* right) {
    const float u = (float)(100 - uniquenessRatio) / 100.0f;
    std::vector<uint16_t> S((size_t)W * D);
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x)
            for (int d = 0; d < D; ++d) {
                unsigned s = 0;
                for (int p = 0; p < P; ++p) s += Ls[p][((size_t)y * W + x) * D + d];
                S[(size_t)x * D + d] = (uint16_t)s;
            }
        for (int x = 0; x < W; ++x) {
            uint32_t b1 = 0xffffffffu, b2 = 0xffffffffu;
            for (int d = 0; d < D; ++d) {
                const uint32_t pk = ((uint32_t)S[(size_t)x * D + d] << 16) | (uint32_t)d;
                if (pk < b1) {
                    b2 = b1;
                    b1 = pk;
                } else if (pk < b2) {
                    b2 = pk;
                }
            }
            const int c1 = (int)(b1 >> 16), d1 = (int)(b1 & 0xffff), c2 = (int)(b2 >> 16), d2 = (int)(b2 & 0xffff);
            const bool reject = ((float)c2 * u < (float)c1) && (std::abs(d1 - d2) > 1);
            uint16_t v = 0xFFFF;
            if (!reject) {
                int subp = d1 << 4;
                if (d1 > 0 && d1 < D - 1) {
                    const int l = S[(size_t)x * D + d1 - 1], r = S[(size_t)x * D + d1 + 1];
                    const int numer = l - r, denom = l - 2 * c1 + r;
                    if (denom != 0) subp += ((numer << 4) + denom) / (2 * denom);
                }
                v = (uint16_t)subp;
            }
            left[(size_t)y * W + x] = v;
            // right
            uint32_t rb = 0xffffffffu;
            for (int d = 0; d < D && x + d < W; ++d) {
                const uint32_t pk = ((uint32_t)S[(size_t)(x + d) * D + d] << 16) | (uint32_t)d;
                rb = std::min(rb, pk);
            }
            right[(size_t)y * W + x] = (uint16_t)(rb & 0xffff);
        }
    }
    return 0;
}

// D7
int orc_median3(const uint16_t* in, int W, int H, uint16_t* out) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            if (x < 1 || y < 1 || x >= W - 1 || y >= H - 1) {
                out[(size_t)y * W + x] = in[(size_t)y * W + x];
                continue;
            }
            uint16_t b[9];
            int n = 0;
            for (int j = -1; j <= 1; ++j)
                for (int i = -1; i <= 1; ++i) b[n++] = in[(size_t)(y + j) * W + x + i];
            std::nth_element(b, b + 4, b + 9);
            out[(size_t)y * W + x] = b[4];
        }
    return 0;
}

// D8+D9
int orc_lr_check_range(const uint16_t* left, const uint16_t* right, const uint8_t* grayLeft, int W, int H, int minDisp,
                       int16_t* out) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const uint16_t org = left[(size_t)y * W + x];
            const int d = (int)org >> 4;
            const int k = x - d;
            bool invalid = grayLeft[(size_t)y * W + x] == 0 || org == 0xFFFF;
            if (!invalid && k >= 0 && k < W) invalid = std::abs((int)right[(size_t)y * W + k] - d) > 1;
            out[(size_t)y * W + x] = invalid ? (int16_t)((minDisp - 1) * 16) : (int16_t)(uint16_t)(org + minDisp * 16);
        }
    return 0;
}

static const int kDirs[8][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}, {1, 1}, {-1, 1}, {1, -1}, {-1, -1}};

int orc_sgm_dirs(int paths, int* out) {
    if (paths != 4 && paths != 8) return -1;
    for (int p = 0; p < paths; ++p) {
        out[2 * p] = kDirs[p][0];
        out[2 * p + 1] = kDirs[p][1];
    }
    return 0;
}

// whole stage: BGR pair -> CV_16SC1 disparity (x16).  Optional intermediates may be null.
int orc_sgm_compute(const uint8_t* leftBgr, const uint8_t* rightBgr, int W, int H, int D, int minDisp, int P1, int P2,
                    int uniquenessRatio, int paths, int16_t* disp, uint32_t* censusL, uint32_t* censusR,
                    uint8_t* volumes /* [paths][H][W][D] or null */, uint16_t* leftRaw, uint16_t* rightRaw) {
    if (paths != 4 && paths != 8) return -1;
    const size_t N = (size_t)W * H;
    std::vector<uint8_t> gl(N), gr(N);
    orc_gray(leftBgr, W, H, gl.data());
    orc_gray(rightBgr, W, H, gr.data());
    std::vector<uint32_t> cl(N), cr(N);
    orc_census(gl.data(), W, H, cl.data());
    orc_census(gr.data(), W, H, cr.data());
    if (censusL) std::memcpy(censusL, cl.data(), N * 4);
    if (censusR) std::memcpy(censusR, cr.data(), N * 4);
    std::vector<uint8_t> own;
    uint8_t* vol = volumes;
    if (!vol) {
        own.resize(N * D * paths);
        vol = own.data();
    }
    std::vector<const uint8_t*> Ls(paths);
    for (int p = 0; p < paths; ++p) {
        uint8_t* L = vol + (size_t)p * N * D;
        if (orc_sgm_path(cl.data(), cr.data(), W, H, D, minDisp, P1, P2, kDirs[p][0], kDirs[p][1], L)) return -1;
        Ls[p] = L;
    }
    std::vector<uint16_t> l0(N), r0(N), l1(N), r1(N);
    orc_sgm_wta(Ls.data(), paths, W, H, D, uniquenessRatio, l0.data(), r0.data());
    if (leftRaw) std::memcpy(leftRaw, l0.data(), N * 2);
    if (rightRaw) std::memcpy(rightRaw, r0.data(), N * 2);
    orc_median3(l0.data(), W, H, l1.data());
    orc_median3(r0.data(), W, H, r1.data());
    orc_lr_check_range(l1.data(), r1.data(), gl.data(), W, H, minDisp, disp);
    return 0;
}

}  // extern "C"
