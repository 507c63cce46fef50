// Closed-form evaluation of what the reference's cooperative tile loader
//   cart::copyToShared<T, XBatch, YBatch, Interpolate>  (/root/reference/include/utils/cuda.cuh:59-191)
// leaves at local position (lx, ly) of block (bx, by)'s shared tile, INCLUDING its deterministic
// defects (row shift for block rows >= 1, clamp over-count with row spill, halo indexing; SURVEY.md
// §8-Q Q1-Q9).  The reference fills the tile with a sequence of writes; this file inverts that sequence
// ("last writer wins", phases examined in reverse program order) so a kernel can fetch any element
// directly from global memory without replaying the loader.  Positions the reference leaves
// uninitialised, and image rows past the last row that it reads out of bounds, evaluate to `undef`.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define CB_HD __host__ __device__ __forceinline__
#else
#define CB_HD inline
#endif

namespace cb {

struct TileGeom {
    int W, H;          // image size
    int tileW, tileH;  // blockDim * batch
    int xPad, yPad;    // halo (note the reference's argument order is (yPadding, xPadding), Q9)
    int XB, YB;        // per-thread batch (needed by the halo phases)
    int alloc;         // elements of the shared array the reference kernel declares
};

template <typename T, typename Img>
struct TileEval {
    const Img& img;
    const TileGeom& g;
    int bx, by;
    T undef;
    int startX, startY;  // 32-bit throughout: images are < 65536 pixels per side (64-bit divisions are slow on the GPU)
    int pXS, pSXS, pYS, pSYS, xDim, yDim, S;

    CB_HD TileEval(const Img& img_, const TileGeom& g_, int bx_, int by_, T undef_)
        : img(img_), g(g_), bx(bx_), by(by_), undef(undef_) {
        startX = bx * g.tileW;
        startY = by * g.tileH;
        pXS = (int)(startX - g.xPad > 0 ? startX - g.xPad : 0);
        pSXS = pXS - (int)startX;
        pYS = (int)(startY - g.yPad > 0 ? startY - g.yPad : 0);
        pSYS = pYS - (int)startY;
        const unsigned fullW = (unsigned)(g.tileW + 2 * g.xPad), fullH = (unsigned)(g.tileH + 2 * g.yPad);
        const unsigned remW = (unsigned)(g.W - pXS), remH = (unsigned)(g.H - pYS);
        xDim = (int)(remW < fullW ? remW : fullW);
        yDim = (int)(remH < fullH ? remH : fullH);
        S = g.tileW + 2 * g.xPad;
    }
    CB_HD int index(int lx, int ly) const { return (ly + g.yPad) * S + (lx + g.xPad); }
    CB_HD T image(int x, int y) const {
        if (x < 0 || x >= g.W || y < 0 || y >= g.H) return undef;
        return img(x, y);
    }
    // value after the body copy only (cuda.cuh:91-96)
    CB_HD T body(int L) const {
        const int rel = L - index(pSXS, pSYS);
        if (rel < 0) return undef;
        const int i = (int)((unsigned)rel / (unsigned)S), e = rel - i * S;
        if (i >= yDim || e >= xDim) return undef;
        return image(pXS + e, startY + i);
    }
    CB_HD T body_at(int lx, int ly) const {
        const int L = index(lx, ly);
        if (L < 0 || L >= g.alloc) return undef;
        return body(L);
    }

    template <bool Interp>
    CB_HD T value(int lx, int ly) const {
        const int L = index(lx, ly);
        if (L < 0 || L >= g.alloc) return undef;
        const bool top = (int)startY - g.yPad < 0;
        const bool bottom = startY + g.tileH + g.yPad > g.H;
        const bool left = (int)startX - g.xPad < 0;
        const bool right = startX + g.tileW + g.xPad > g.W;
        if (Interp) {
            // right halo: (tileW+i, j), j < YB, startY + j < H  (cuda.cuh:174-186)
            if (right && lx >= g.tileW && lx < g.tileW + g.xPad && ly >= 0 && ly < g.YB && startY + ly < g.H) {
                const int i = lx - g.tileW;
                const T border = body_at(g.tileW - 1, ly), prev = body_at(g.tileW - 2 - i, ly);
                return (T)(border + (border - prev));
            }
            // left halo: (-i, j), j < YB  (cuda.cuh:151-163)
            if (left && lx < 0 && lx >= -g.xPad && ly >= 0 && ly < g.YB && startY + ly < g.H) {
                return body_at(-lx, ly);
            }
            // bottom halo: (c, tileH+i), startX + c < W  (cuda.cuh:128-140)
            if (bottom && ly >= g.tileH && ly < g.tileH + g.yPad && lx >= 0 && lx < g.tileW && startX + lx < g.W) {
                const int i = ly - g.tileH;
                const T border = body_at(lx, g.tileH - 1), prev = body_at(lx, g.tileH - 2 - i);
                return (T)(border + (border - prev));
            }
            // top halo: (c, -i) <- tile(c mod XB, i)  (cuda.cuh:106-118)
            if (top && ly < 0 && ly >= -g.yPad && lx >= 0 && lx < g.tileW && startX + lx < g.W) {
                return body_at(lx % g.XB, -ly);
            }
            return body(L);
        } else {
            // right: (tileW+i, j) <- img(W-1, startY+j), j < yDim  (cuda.cuh:169-173)
            if (right && lx >= g.tileW && lx < g.tileW + g.xPad && ly >= 0 && ly < yDim) return image(g.W - 1, startY + ly);
            // left: (-i, j) <- img(0, startY+j), j < yDim  (cuda.cuh:146-150)
            if (left && lx < 0 && lx >= -g.xPad && ly >= 0 && ly < yDim) return image(0, startY + ly);
            // bottom rows tileH+i <- image row H-1 (row copies, later i wins)  (cuda.cuh:125-127)
            if (bottom) {
                for (int i = g.yPad - 1; i >= 0; --i) {
                    const int rel = L - index(pSXS, g.tileH + i);
                    if (rel >= 0 && rel < xDim) return image(pXS + rel, g.H - 1);
                }
            }
            // top rows -i <- image row 0  (cuda.cuh:103-105)
            if (top) {
                for (int i = g.yPad; i >= 1; --i) {
                    const int rel = L - index(pSXS, -i);
                    if (rel >= 0 && rel < xDim) return image(pXS + rel, 0);
                }
            }
            return body(L);
        }
    }
};

}  // namespace cb
