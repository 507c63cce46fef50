// TEST INFRASTRUCTURE ONLY. Nothing under oracle/ may be linked, imported or executed by the
// product path (cart_slam_b200/); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, and only as the checker / CPU baseline.
//
// Literal CPU simulation of the reference's cooperative tile loader
//   cart::copyToShared<T, XBatch, YBatch, Interpolate>   (/root/reference/include/utils/cuda.cuh:59-191)
// including its deterministic defects (Q1 row shift, Q3 clamp over-count with row
// spill, Q4/Q6 halo indexing, Q5/Q7 halos at the tile end, Q8 label-tile halos).
//
// The simulation writes into a linear array exactly as the reference does (SHARED_INDEX, cuda.cuh:15)
// with the array size the calling kernel declares; writes that would land beyond that array are
// dropped (the reference's shared-memory overflow, Q3/Q17, is NOT reproduced) and reads of image rows
// past the last row (Q2) return `undef_value`.  Every element carries a `def` flag: 0 = the reference
// would hold an uninitialised / out-of-bounds value there.  Stage functions propagate the flag to a
// per-pixel "defined" mask that the golden-vector tests use when comparing with the real reference
// kernels; comparisons between the oracle and the CUDA path are exact everywhere (both use the
// canonical `undef_value`).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace orc {

template <typename T>
struct Tile {
    std::vector<T> s;
    std::vector<uint8_t> def;
    int tileW = 0, tileH = 0, xPad = 0, yPad = 0;
    long rowStride = 0;  // tileW + 2*xPad
    T undef{};

    // SHARED_INDEX (cuda.cuh:15)
    long index(int lx, int ly) const { return (long)(ly + yPad) * rowStride + (lx + xPad); }
    bool inAlloc(long idx) const { return idx >= 0 && idx < (long)s.size(); }
    T get(int lx, int ly, bool* defined = nullptr) const {
        long idx = index(lx, ly);
        if (!inAlloc(idx)) {
            if (defined) *defined = false;
            return undef;
        }
        if (defined) *defined = def[idx] != 0;
        return s[idx];
    }
    void put(long idx, T v, bool d) {
        if (!inAlloc(idx)) return;  // reference smem overflow: dropped (Q3/Q17)
        s[idx] = v;
        def[idx] = d ? 1 : 0;
    }
    void set(int lx, int ly, T v, bool d) { put(index(lx, ly), v, d); }
};

// img: tightly packed W x H.  (bx,by): block index; (bdx,bdy): blockDim; XB,YB: batch.
// allocElems: number of T elements the calling kernel declares for the shared array.
template <typename T, bool Interp>
void copy_to_shared(Tile<T>& t, const T* img, int W, int H, int bx, int by, int bdx, int bdy, int XB, int YB,
                    int yPadding, int xPadding, size_t allocElems, T undefValue) {
    t.tileW = XB * bdx;
    t.tileH = YB * bdy;
    t.xPad = xPadding;
    t.yPad = yPadding;
    t.rowStride = t.tileW + 2 * xPadding;
    t.undef = undefValue;
    t.s.assign(allocElems, undefValue);
    t.def.assign(allocElems, 0);

    const long startX = (long)bx * bdx * XB;  // cuda.cuh:78
    const long startY = (long)by * bdy * YB;  // cuda.cuh:79
    const int pXS = std::max(0L, startX - xPadding);
    const int pSXS = pXS - (int)startX;
    const int pYS = std::max(0L, startY - yPadding);
    const int pSYS = pYS - (int)startY;
    const int xDim = (int)std::min<unsigned>((unsigned)(W - pXS), (unsigned)(t.tileW + 2 * xPadding));  // :87
    const int yDim = (int)std::min<unsigned>((unsigned)(H - pYS), (unsigned)(t.tileH + 2 * yPadding));  // :88

    auto imgAt = [&](long x, long y, bool& d) -> T {
        if (y < 0 || y >= H || x < 0 || x >= W) {  // Q2: past the last row -> canonical undef
            d = false;
            return undefValue;
        }
        d = true;
        return img[(size_t)y * W + x];
    };
    // row copy = cg::memcpy_async of n elements, linear in shared memory
    auto rowCopy = [&](int lx, int ly, long sx, long sy, int n) {
        long base = t.index(lx, ly);
        for (int e = 0; e < n; ++e) {
            bool d;
            T v = imgAt(sx + e, sy, d);
            t.put(base + e, v, d);
        }
    };

    // body (cuda.cuh:91-96): note source row startY + i (Q1), yDim/xDim over-count when clamped (Q3)
    for (int i = 0; i < yDim; ++i) rowCopy(pSXS, pSYS + i, pXS, startY + i, xDim);

    // top halo (cuda.cuh:100-120)
    if ((int)startY - yPadding < 0) {
        for (int i = 1; i <= yPadding; ++i) {
            if (!Interp) {
                rowCopy(pSXS, -i, pXS, 0, xDim);
            } else {
                const int ty = 0;
                for (int tx = 0; tx < bdx; ++tx) {
                    const int sharedPixelX = tx * XB, sharedPixelY = ty * YB;
                    const long pixelX = ((long)bx * bdx + tx) * XB;
                    for (int j = 0; j < XB; ++j) {
                        if (pixelX + j >= W) break;
                        bool d0, d1;
                        T border = t.get(sharedPixelX + j, 0, &d0);
                        T next = t.get(sharedPixelY + j, i, &d1);  // Q4: column sharedPixelY + j
                        T value = (T)(border + (next - border));
                        t.set(sharedPixelX + j, sharedPixelY - i, value, d0 && d1);
                    }
                }
            }
        }
    }
    // bottom halo (cuda.cuh:122-142) - at the TILE end (Q5)
    if (startY + (long)YB * bdy + yPadding > H) {
        for (int i = 0; i < yPadding; ++i) {
            if (!Interp) {
                rowCopy(pSXS, YB * bdy + i, pXS, H - 1, xDim);
            } else {
                const int ty = bdy - 1;
                (void)ty;
                for (int tx = 0; tx < bdx; ++tx) {
                    const int sharedPixelX = tx * XB;
                    const long pixelX = ((long)bx * bdx + tx) * XB;
                    for (int j = 0; j < XB; ++j) {
                        if (pixelX + j >= W) break;
                        bool d0, d1;
                        T border = t.get(sharedPixelX + j, YB * bdy - 1, &d0);
                        T prev = t.get(sharedPixelX + j, YB * bdy - 2 - i, &d1);
                        T value = (T)(border + (border - prev));
                        t.set(sharedPixelX + j, YB * bdy + i, value, d0 && d1);
                    }
                }
            }
        }
    }
    // left halo (cuda.cuh:144-165)
    if ((int)startX - xPadding < 0) {
        for (int i = 1; i <= xPadding; ++i) {
            if (!Interp) {
                for (int j = 0; j < yDim; ++j) rowCopy(-i, j, 0, startY + j, 1);  // Q8: local row j, unshifted
            } else {
                // every thread with threadIdx.x == 0 writes the same local rows 0..YB-1 (Q6)
                for (int ty = 0; ty < bdy; ++ty) {
                    const long pixelY = ((long)by * bdy + ty) * YB;
                    for (int j = 0; j < YB; ++j) {
                        if (pixelY + j >= H) break;
                        bool d0, d1;
                        T border = t.get(0, j, &d0);
                        T next = t.get(i, j, &d1);
                        T value = (T)(border + (next - border));
                        t.set(-i, j, value, d0 && d1);
                    }
                }
            }
        }
    }
    // right halo (cuda.cuh:167-188) - at the TILE end (Q7)
    if (startX + (long)XB * bdx + xPadding > W) {
        for (int i = 0; i < xPadding; ++i) {
            if (!Interp) {
                for (int j = 0; j < yDim; ++j) rowCopy(XB * bdx + i, j, W - 1, startY + j, 1);
            } else {
                for (int ty = 0; ty < bdy; ++ty) {
                    const long pixelY = ((long)by * bdy + ty) * YB;
                    for (int j = 0; j < YB; ++j) {
                        if (pixelY + j >= H) break;
                        bool d0, d1;
                        T border = t.get(XB * bdx - 1, j, &d0);
                        T prev = t.get(XB * bdx - 2 - i, j, &d1);
                        T value = (T)(border + (border - prev));
                        t.set(XB * bdx + i, j, value, d0 && d1);
                    }
                }
            }
        }
    }
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace orc
