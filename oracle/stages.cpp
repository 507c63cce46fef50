// TEST INFRASTRUCTURE ONLY (see tile.hpp header).
//
// Scalar restatements of the reference's own CUDA kernels after the SGM stage:
//   interpolateKernel                 /root/reference/src/modules/disparity/interpolation.cu:17-82
//   calculateDirectionalDerivatives   /root/reference/src/modules/disparity/derivative.cu:27-97
//   mergeDerivativeHistograms         /root/reference/src/modules/disparity/derivative.cu:99-116
//   calculateDerivatives (naive)      /root/reference/src/modules/planeseg/planeseg.cu:31-142
//   classifyPlanes (naive)            /root/reference/src/modules/planeseg/planeseg.cu:160-243 (no temporal vote)
//   performSuperPixelClassifications  /root/reference/src/modules/planeseg/sp_planeseg.cu:25-134 (no temporal vote)
//   classifyPlanes (SP)               /root/reference/src/modules/planeseg/sp_planeseg.cu:136-184
//   HistogramPeak parameter update    /root/reference/src/modules/planeseg/planeseg.cu:405-458
//   util::findPeaks                   /root/reference/src/utils/peaks.cpp:12-72
// Tile semantics come from tile.hpp (copyToShared, bug-compatible).  Canonical choices where the
// reference races: Q10 low-pass is out-of-place, Q11 interpolation is Jacobi,
// Q17 histograms are recounted from the oracle's own derivative image.
// Pinned against the real reference kernels (built by oracle/ref/, run on a B200) through the golden
// vectors in tests/golden/ on the "defined" masks this file produces.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "oracle.h"
#include "tile.hpp"

using orc::ceil_div;
using orc::Tile;

static const int16_t INVALID = -32768;  // CARTSLAM_DISPARITY_INVALID, disparity.hpp:17

extern "C" {

// interpolation.cu:17-82, launch :85-99 (16x16 threads, 4x4 batch -> 64x64 tile, dynamic smem exactly sized).
// disp is updated in place; `defined` (may be null) gets 1 where the reference's result does not
// depend on uninitialised / out-of-image memory.
int orc_interpolate(int16_t* disp, int W, int H, int radius, int iterations, int minDisparity, int maxDisparity,
                    uint8_t* defined) {
    if (radius <= 0) return 0;
    const int BD = 16, XB = 4, YB = 4, TW = BD * XB, TH = BD * YB, pad = radius - 1;
    const size_t alloc = (size_t)(TW + 2 * pad) * (TH + 2 * pad);  // SHARED_SIZE(radius), :13
    const unsigned minCount = (unsigned)(radius * radius + 1);     // :33
    std::vector<int16_t> out(disp, disp + (size_t)W * H);
    Tile<int16_t> t;
    for (int by = 0; by < ceil_div(H, TH); ++by)
        for (int bx = 0; bx < ceil_div(W, TW); ++bx) {
            orc::copy_to_shared<int16_t, true>(t, disp, W, H, bx, by, BD, BD, XB, YB, pad, pad, alloc, INVALID);
            for (int it = 0; it < iterations; ++it) {
                // Q11 canonical schedule: Jacobi (all reads see the previous iteration)
                std::vector<int16_t> ns = t.s;
                std::vector<uint8_t> nd = t.def;
                for (int ly = 0; ly < TH; ++ly)
                    for (int lx = 0; lx < TW; ++lx) {
                        if (bx * TW + lx >= W || by * TH + ly >= H) continue;
                        int sum = 0, count = 0;
                        bool d = true;
                        for (int k = -radius + 1; k < radius; ++k)
                            for (int l = -radius + 1; l < radius; ++l) {
                                bool dd;
                                const int16_t v = t.get(lx + k, ly + l, &dd);
                                d = d && dd;
                                if (v > minDisparity && v < maxDisparity) {  // :52 (upper bound is NOT x16, Q11)
                                    sum += v;
                                    count++;
                                }
                            }
                        const long idx = t.index(lx, ly);
                        ns[idx] = ((unsigned)count > minCount) ? (int16_t)(sum / count) : INVALID;
                        nd[idx] = d;
                    }
                t.s.swap(ns);
                t.def.swap(nd);
            }
            for (int ly = 0; ly < TH; ++ly)
                for (int lx = 0; lx < TW; ++lx) {
                    const int x = bx * TW + lx, y = by * TH + ly;
                    if (x >= W || y >= H) continue;
                    bool d;
                    out[(size_t)y * W + x] = t.get(lx, ly, &d);
                    if (defined) defined[(size_t)y * W + x] = d;
                }
        }
    std::memcpy(disp, out.data(), sizeof(int16_t) * (size_t)W * H);
    return 0;
}

// derivative.cu:27-116. deriv: [H][W][2] (ch0 vertical, ch1 horizontal); hist: [256][2] int32;
// defined: [H][W][2] or null.
int orc_derivative(const int16_t* disp, int W, int H, int16_t* deriv, int32_t* hist, uint8_t* defined) {
    const int BD = 32, XB = 4, YB = 4, TW = 128, TH = 128, OFF = 2;
    const size_t alloc = (size_t)(OFF * 2 + TW) * (OFF * 2 + TH);  // SHARED_SIZE, :23
    std::memset(hist, 0, sizeof(int32_t) * 512);
    Tile<int16_t> t;
    for (int by = 0; by < ceil_div(H, TH); ++by)
        for (int bx = 0; bx < ceil_div(W, TW); ++bx) {
            orc::copy_to_shared<int16_t, true>(t, disp, W, H, bx, by, BD, BD, XB, YB, OFF, OFF, alloc, INVALID);
            for (int ly = 0; ly < TH; ++ly)
                for (int lx = 0; lx < TW; ++lx) {
                    const int x = bx * TW + lx, y = by * TH + ly;
                    if (x >= W || y >= H) continue;
                    bool du, dd, dl, dr;
                    const int16_t up = t.get(lx, ly - OFF, &du), dn = t.get(lx, ly + OFF, &dd);
                    const int16_t lf = t.get(lx - OFF, ly, &dl), rt = t.get(lx + OFF, ly, &dr);
                    const int16_t dv = (int16_t)(dn - up), dh = (int16_t)(rt - lf);
                    const bool vv = up != INVALID && dn != INVALID, hv = lf != INVALID && rt != INVALID;
                    const int16_t ov = vv ? dv : INVALID, oh = hv ? dh : INVALID;
                    const size_t o = ((size_t)y * W + x) * 2;
                    deriv[o] = ov;
                    deriv[o + 1] = oh;
                    if (defined) {
                        defined[o] = du && dd;
                        defined[o + 1] = dl && dr;
                    }
                    // Q17: histogram recounted from the oracle's own output
                    if (vv && dv >= -128 && dv <= 127) hist[2 * (dv + 128)]++;
                    if (hv && dh >= -128 && dh <= 127) hist[2 * (dh + 128) + 1]++;
                }
        }
    return 0;
}

// planeseg.cu:31-142 (calculateDerivatives). deriv: [H][W] s16; hist: this frame's 256-bin histogram
// (the reference adds it to a running total, planeseg.cu:144-158 - kept by the caller); defined or null.
int orc_naive_derivative(const int16_t* disp, int W, int H, int16_t* deriv, int32_t* hist, uint8_t* defined) {
    const int BD = 32, XB = 4, YB = 4, TW = 128, TH = 128, PADY = 2;
    const size_t alloc = (size_t)(XB * BD) * (YB * (PADY * 2 + BD));  // SHARED_SIZE, planeseg.cu:20 (over-allocated)
    std::memset(hist, 0, sizeof(int32_t) * 256);
    Tile<int16_t> t;
    for (int by = 0; by < ceil_div(H, TH); ++by)
        for (int bx = 0; bx < ceil_div(W, TW); ++bx) {
            orc::copy_to_shared<int16_t, true>(t, disp, W, H, bx, by, BD, BD, XB, YB, PADY, 0, alloc, INVALID);  // Q9
            // vertical 5-tap valid-mean over local rows 0..TH-1, all TW columns (no bounds test, :60-104).
            // Q10 canonical schedule: out-of-place (every tap reads the unfiltered tile).
            std::vector<int16_t> F((size_t)TW * TH);
            std::vector<uint8_t> Fd((size_t)TW * TH);
            for (int ly = 0; ly < TH; ++ly)
                for (int lx = 0; lx < TW; ++lx) {
                    int16_t sum = 0;  // derivative_t sum, :62 (16-bit accumulate)
                    int count = 0;
                    bool d = true;
                    for (int k = -PADY; k <= PADY; ++k) {
                        bool dd;
                        const int16_t v = t.get(lx, ly + k, &dd);
                        d = d && dd;
                        if (v != INVALID) {
                            sum = (int16_t)(sum + v);
                            count++;
                        }
                    }
                    F[(size_t)ly * TW + lx] = count == 0 ? INVALID : (int16_t)(sum / count);
                    Fd[(size_t)ly * TW + lx] = d;
                }
            auto Fget = [&](int lx, int ly, bool& d) -> int16_t {
                if (ly >= 0 && ly < TH) {  // filtered rows
                    d = Fd[(size_t)ly * TW + lx] != 0;
                    return F[(size_t)ly * TW + lx];
                }
                bool dd;  // Q10b: halo rows stay unfiltered
                const int16_t v = t.get(lx, ly, &dd);
                d = dd;
                return v;
            };
            for (int ly = 0; ly < TH; ++ly)
                for (int lx = 0; lx < TW; ++lx) {
                    const int x = bx * TW + lx, y = by * TH + ly;
                    if (x >= W || y >= H) continue;
                    bool d0, d1, d2;
                    const int16_t c = Fget(lx, ly, d0), n = Fget(lx, ly + 1, d1), p = Fget(lx, ly - 1, d2);
                    const int16_t dv = (int16_t)(n - p);
                    const bool valid = c != INVALID && n != INVALID && p != INVALID;
                    deriv[(size_t)y * W + x] = valid ? dv : INVALID;
                    if (defined) defined[(size_t)y * W + x] = d0 && d1 && d2;
                    if (valid && dv >= -128 && dv <= 127) hist[dv + 128]++;
                }
        }
    return 0;
}

// planeseg.cu:188-197 / sp_planeseg.cu:68-77: range rule on one derivative channel.
// deriv has `stride` int16 per pixel (1 naive, 2 for CV_16SC2 ch0). params = {hS,hE,vS,vE}.
int orc_classify(const int16_t* deriv, long n, int stride, int hS, int hE, int vS, int vE, uint8_t* planes) {
    for (long i = 0; i < n; ++i) {
        const int16_t d = deriv[i * stride];
        uint8_t p = 2;  // UNKNOWN
        if (d != INVALID && d >= hS && d < hE)
            p = 0;  // HORIZONTAL
        else if (d != INVALID && d >= vS && d < vE)
            p = 1;  // VERTICAL
        planes[i] = p;
    }
    return 0;
}

// sp_planeseg.cu:25-184 with previousPlanesCount == 0.
// labels u16 [H][W]; maxLabel = label COUNT (a11).  Returns -2 if a label >= maxLabel occurs,
// -3 when (maxLabel+1)*6 > 32768 (sp_planeseg.cu:327-331), -4 if a vote counter would exceed u16.
int orc_sp_planeseg(const int16_t* deriv2, const uint16_t* labels, int W, int H, int maxLabel, int hS, int hE, int vS,
                    int vE, uint8_t* planesUnsmoothed, uint8_t* planes) {
    if ((size_t)(maxLabel + 1) * 3 * sizeof(uint16_t) > 32768) return -3;
    const long N = (long)W * H;
    orc_classify(deriv2, N, 2, hS, hE, vS, vE, planesUnsmoothed);
    std::vector<uint32_t> votes((size_t)maxLabel * 3, 0);
    for (long i = 0; i < N; ++i) {
        if (labels[i] >= maxLabel) return -2;
        if (++votes[(size_t)labels[i] * 3 + planesUnsmoothed[i]] > 65535u) return -4;
    }
    std::vector<uint8_t> assign(maxLabel);
    for (int l = 0; l < maxLabel; ++l) {
        int maxVotes = (int)votes[(size_t)l * 3 + 2];
        uint8_t best = 2;
        const int v = (int)votes[(size_t)l * 3 + 1], h = (int)votes[(size_t)l * 3 + 0];
        if (v > maxVotes) {
            maxVotes = v;
            best = 1;
        }
        if (h > maxVotes) best = 0;
        assign[l] = best;
    }
    for (long i = 0; i < N; ++i) planes[i] = assign[labels[i]];
    return 0;
}

// ---- temporal smoothing vote ---------------------------------------------------------
// Per-pixel vote of classifyPlanes (planeseg.cu:199-240, mode 0) and of performSuperPixelClassifications
// (sp_planeseg.cu:79-117, mode 1).  prevPlanes[k] = "planes_unsmoothed" of frame id-(k+1), prevFlow[k] = "optflow"
// (CV_16SC2, S10.5) of frame id-k, both tightly packed here.  Every flow image is read at the CURRENT pixel, the
// position walks back by the accumulated integer flow (arithmetic >> 5), positions outside the image are skipped
// but stay accumulated.  Mode 0: current plane counts once, winner = H if votes[H] > votes[V] else V, UNKNOWN when
// the winner has no vote.  Mode 1: current plane counts twice, UNKNOWN when it outvotes the winner.
static inline uint8_t temporal_vote_px(int mode, uint8_t plane, int px, int py, int W, int H, int count,
                                       const uint8_t* const* prevPlanes, const int16_t* const* prevFlow) {
    int votes[3] = {0, 0, 0};
    votes[plane] += mode ? 2 : 1;
    int x = px, y = py;
    for (int k = 0; k < count; ++k) {
        const int16_t* fl = prevFlow[k] + ((size_t)py * W + px) * 2;
        x -= (int16_t)fl[0] >> 5;
        y -= (int16_t)fl[1] >> 5;
        if (x < 0 || y < 0 || x >= W || y >= H) continue;
        votes[prevPlanes[k][(size_t)y * W + x]]++;
    }
    int best = votes[0] > votes[1] ? 0 : 1;
    if (mode == 0) {
        if (votes[best] == 0) best = 2;
    } else {
        if (votes[best] < votes[2]) best = 2;
    }
    return (uint8_t)best;
}

// classifyPlanes with previousPlanesCount = count (planeseg.cu:160-243).  count == 0: the reference leaves the
// smoothed image untouched and the module returns the unsmoothed image under both keys (planeseg.cu:361-368);
// here smoothed = unsmoothed.  Returns -2 if a previous plane value is > 2 (the reference would index out of bounds).
int orc_classify_temporal(const int16_t* deriv, int W, int H, int stride, int hS, int hE, int vS, int vE, int count,
                          const uint8_t* const* prevPlanes, const int16_t* const* prevFlow, uint8_t* planesUnsmoothed,
                          uint8_t* planesSmoothed) {
    const long N = (long)W * H;
    for (int k = 0; k < count; ++k)
        for (long i = 0; i < N; ++i)
            if (prevPlanes[k][i] > 2) return -2;
    orc_classify(deriv, N, stride, hS, hE, vS, vE, planesUnsmoothed);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            planesSmoothed[i] = count > 0 ? temporal_vote_px(0, planesUnsmoothed[i], x, y, W, H, count, prevPlanes, prevFlow)
                                          : planesUnsmoothed[i];
        }
    return 0;
}

// performSuperPixelClassifications + classifyPlanes with previousPlanesCount = count (sp_planeseg.cu:25-184):
// the per-pixel vote result (not the unsmoothed class) feeds the superpixel counters.
int orc_sp_planeseg_temporal(const int16_t* deriv2, const uint16_t* labels, int W, int H, int maxLabel, int hS, int hE,
                             int vS, int vE, int count, const uint8_t* const* prevPlanes, const int16_t* const* prevFlow,
                             uint8_t* planesUnsmoothed, uint8_t* planes) {
    if ((size_t)(maxLabel + 1) * 3 * sizeof(uint16_t) > 32768) return -3;
    const long N = (long)W * H;
    for (int k = 0; k < count; ++k)
        for (long i = 0; i < N; ++i)
            if (prevPlanes[k][i] > 2) return -2;
    orc_classify(deriv2, N, 2, hS, hE, vS, vE, planesUnsmoothed);
    std::vector<uint32_t> votes((size_t)maxLabel * 3, 0);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            if (labels[i] >= maxLabel) return -2;
            const uint8_t p = count > 0 ? temporal_vote_px(1, planesUnsmoothed[i], x, y, W, H, count, prevPlanes, prevFlow)
                                        : planesUnsmoothed[i];
            if (++votes[(size_t)labels[i] * 3 + p] > 65535u) return -4;
        }
    for (long i = 0; i < N; ++i) {
        const uint32_t* v = &votes[(size_t)labels[i] * 3];
        int maxVotes = (int)v[2];
        uint8_t best = 2;
        if ((int)v[1] > maxVotes) {
            maxVotes = (int)v[1];
            best = 1;
        }
        if ((int)v[0] > maxVotes) best = 0;
        planes[i] = best;
    }
    return 0;
}

// ---- superpixel consumers of the plane fit ---------------------------------------------
// IS_VALID_DEPTH, planefit.cu:19: finite, <= 40 and > 0 (float compared against double constants).
static inline bool valid_depth(float z) { return std::isfinite(z) && (double)z <= 40.0 && (double)z > 0.0; }

// countPixels, planefit.cu:38-83: per label the pixel count and the count of pixels whose depth Z is invalid.
// The reference keeps both in uint16 counters updated with a non-carry-safe 16-bit atomic (cuda.cuh:31-36);
// this restatement counts in 32 bits and returns -4 when a count would not fit 16 bits (the reference's result is
// then corrupted by the carry into the neighbouring counter).  labels < nLabels (= maxLabelId + 1) or -2.
int orc_label_statistics(const uint16_t* labels, const float* xyz, int W, int H, int nLabels, uint32_t* pixelCount,
                         uint32_t* pixelCountInvalid) {
    for (int l = 0; l < nLabels; ++l) pixelCount[l] = pixelCountInvalid[l] = 0;
    for (long i = 0; i < (long)W * H; ++i) {
        const int l = labels[i];
        if (l >= nLabels) return -2;
        if (!valid_depth(xyz[3 * i + 2])) pixelCountInvalid[l]++;
        pixelCount[l]++;
    }
    for (int l = 0; l < nLabels; ++l)
        if (pixelCount[l] > 65535u) return -4;
    return 0;
}

// calculateRegionDistance, planefit.cu:85-138 with calculateDistanceFromPlane :34-36: inliers[plane][label] = number
// of valid-depth pixels of the label closer than `threshold` to plane (a, b, c, d); double arithmetic on float
// coordinates, evaluated left to right without contraction.  planes: nPlanes x 4 doubles.
int orc_region_inliers(const uint16_t* labels, const float* xyz, int W, int H, int nLabels, const double* planes,
                       int nPlanes, double threshold, uint32_t* inliers) {
    for (long i = 0; i < (long)nPlanes * nLabels; ++i) inliers[i] = 0;
    for (int p = 0; p < nPlanes; ++p) {
        const double a = planes[4 * p], b = planes[4 * p + 1], c = planes[4 * p + 2], d = planes[4 * p + 3];
        const double norm = std::sqrt(a * a + b * b + c * c);
        for (long i = 0; i < (long)W * H; ++i) {
            const int l = labels[i];
            if (l >= nLabels) return -2;
            const float* q = xyz + 3 * i;
            if (!valid_depth(q[2])) continue;
            const double dist = std::fabs(a * (double)q[0] + b * (double)q[1] + c * (double)q[2] + d) / norm;
            if (dist < threshold) inliers[(size_t)p * nLabels + l]++;
        }
    }
    return 0;
}

// ---- overlay kernels (visual QA) -------------------------------------------------------------
// overlayPlanes, planeseg_vis.cu:28-56 with the colour table :21-26 (PlaneColor, planeseg.hpp:44-66, halved):
// out = bgr / 2 + colour[plane] / 2 per channel.  Plane values > 2 index past the table in the reference: -2 here.
int orc_overlay_planes(const uint8_t* bgr, const uint8_t* planes, int W, int H, uint8_t* out) {
    static const int colors[3][3] = {{255 / 2, 0, 0}, {0, 255 / 2, 0}, {0, 0, 255 / 2}};  // B, G, R of H / V / UNKNOWN
    for (long i = 0; i < (long)W * H; ++i) {
        if (planes[i] > 2) return -2;
        for (int c = 0; c < 3; ++c) out[3 * i + c] = (uint8_t)(bgr[3 * i + c] / 2 + colors[planes[i]][c]);
    }
    return 0;
}

// overlayBoundaryVisualization, superpixels/visualization.cu:9-42: a pixel whose right or lower neighbour carries a
// different label becomes red (0, 0, 255), the others copy the image; the last row and the last column are NOT
// written (:20-22) - `out` keeps whatever it held there.
int orc_overlay_boundaries(const uint8_t* bgr, const uint16_t* labels, int W, int H, uint8_t* out) {
    for (int y = 0; y < H - 1; ++y)
        for (int x = 0; x < W - 1; ++x) {
            const size_t i = (size_t)y * W + x;
            const bool edge = labels[i] != labels[i + 1] || labels[i] != labels[i + W];
            out[3 * i] = edge ? 0 : bgr[3 * i];
            out[3 * i + 1] = edge ? 0 : bgr[3 * i + 1];
            out[3 * i + 2] = edge ? 255 : bgr[3 * i + 2];
        }
    return 0;
}

// ---- host-side parameter estimation --------------------------------------------------------------

struct Peak {
    int born, left, right, died;
    int persistence(const int32_t* h) const { return died == -1 ? INT_MAX : h[born] - h[died]; }
};

// peaks.cpp:12-72 (same std::sort calls, same comparators)
static std::vector<Peak> find_peaks(const int32_t* data, int n) {
    std::vector<Peak> peaks;
    std::vector<int> idxtopeak(n, -1), indices(n);
    for (int i = 0; i < n; ++i) indices[i] = i;
    std::sort(indices.begin(), indices.end(), [&](int a, int b) { return data[a] > data[b]; });
    for (int idx : indices) {
        const bool lftdone = idx > 0 && idxtopeak[idx - 1] != -1;
        const bool rgtdone = idx < n - 1 && idxtopeak[idx + 1] != -1;
        const int il = lftdone ? idxtopeak[idx - 1] : -1;
        const int ir = rgtdone ? idxtopeak[idx + 1] : -1;
        if (!lftdone && !rgtdone) {
            peaks.push_back(Peak{idx, idx, idx, -1});
            idxtopeak[idx] = (int)peaks.size() - 1;
        } else if (lftdone && !rgtdone) {
            peaks[il].right += 1;
            idxtopeak[idx] = il;
        } else if (!lftdone && rgtdone) {
            peaks[ir].left -= 1;
            idxtopeak[idx] = ir;
        } else {
            if (data[peaks[il].born] > data[peaks[ir].born]) {
                peaks[ir].died = idx;
                peaks[il].right = peaks[ir].right;
                idxtopeak[peaks[il].right] = idxtopeak[idx] = il;
            } else {
                peaks[il].died = idx;
                peaks[ir].left = peaks[il].left;
                idxtopeak[peaks[ir].left] = idxtopeak[idx] = ir;
            }
        }
    }
    std::sort(peaks.begin(), peaks.end(), [&](Peak a, Peak b) { return a.persistence(data) > b.persistence(data); });
    return peaks;
}

// out: up to maxPeaks rows of {born,left,right,died}; returns the number of peaks found.
int orc_find_peaks(const int32_t* hist, int n, int* out, int maxPeaks) {
    auto p = find_peaks(hist, n);
    for (int i = 0; i < (int)p.size() && i < maxPeaks; ++i) {
        out[4 * i] = p[i].born;
        out[4 * i + 1] = p[i].left;
        out[4 * i + 2] = p[i].right;
        out[4 * i + 3] = p[i].died;
    }
    return (int)p.size();
}

// planeseg.cu:405-458. params = {horizontalCenter, verticalCenter, hS, hE, vS, vE}, updated in place.
// Returns 1 if the ranges were updated, 0 if one of the early returns fired (ranges untouched; the
// centres are assigned before those returns, as in the reference).
int orc_histogram_peak_update(const int32_t* hist, int* params) {
    auto peaks = find_peaks(hist, 256);
    if (peaks.size() < 2) return 0;
    if (std::abs(peaks[0].born - 128) > std::abs(peaks[1].born - 128)) std::swap(peaks[0], peaks[1]);
    // NOTE: the reference assigns the centres before the early returns below (planeseg.cu:418-419)
    params[1] = peaks[0].born - 128;
    params[0] = peaks[1].born - 128;
    int minIndex = std::min(peaks[0].born, peaks[1].born);
    for (int i = minIndex; i < std::max(peaks[0].born, peaks[1].born); ++i)
        if (hist[i] < hist[minIndex]) minIndex = i;
    const int vDist = std::abs(minIndex - peaks[0].born), hDist = std::abs(minIndex - peaks[1].born);
    if (vDist == 0 || hDist == 0) return 0;
    const int vDer = (hist[peaks[0].born] - hist[minIndex]) / vDist;
    const int hDer = (hist[peaks[1].born] - hist[minIndex]) / hDist;
    if (vDer == 0 || hDer == 0) return 0;
    const int vWidth = hist[peaks[0].born] / vDer, hWidth = hist[peaks[1].born] / hDer;
    params[4] = peaks[0].born - vWidth - 128;
    params[5] = minIndex - 127;
    params[2] = minIndex - 127;
    params[3] = peaks[1].born + hWidth - 127;
    return 1;
}


// DepthModule::runInternal, /root/reference/src/modules/depth.cpp:9-25: disparity / 16 as float, then the
// THIRD-PARTY cv::cuda::reprojectImageTo3D(disparityFloat, depth, Q, 3) (opencv_contrib cudastereo, absent from
// /root/reference; recollection of its kernel: single precision, per pixel
//   q* = x Q[*][0] + y Q[*][1] + Q[*][3];  iW = 1 / (qw + Q[3][2] d);  out = ((q* + Q[*][2] d) iW) for x, y, z,
// no handling of invalid disparities).  Pinned within float tolerance against CPU cv2.reprojectImageTo3D
// (tests/test_oracle_cpu.py) - the CPU version accumulates in double, hence a tolerance and not bit equality.
int orc_depth(const int16_t* disp, int W, int H, const float* Q, float* xyz) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const float d = (float)disp[(size_t)y * W + x] * (1.0f / 16.0f);
            const float fx = (float)x, fy = (float)y;
            const float qx = fx * Q[0] + fy * Q[1] + Q[3], qy = fx * Q[4] + fy * Q[5] + Q[7];
            const float qz = fx * Q[8] + fy * Q[9] + Q[11], qw = fx * Q[12] + fy * Q[13] + Q[15];
            const float iW = 1.0f / (qw + Q[14] * d);
            float* o = xyz + ((size_t)y * W + x) * 3;
            o[0] = (qx + Q[2] * d) * iW;
            o[1] = (qy + Q[6] * d) * iW;
            o[2] = (qz + Q[10] * d) * iW;
        }
    return 0;
}

// cv::cuda::resize(src, dst, size, 0, 0, cv::INTER_LINEAR) on CV_8UC3, as the KITTI source applies it when the configured
// image size differs from the files' (/root/reference/src/sources/kitti.cpp:166-169).  The arithmetic lives in
// opencv_contrib/cudawarping (third party, absent): PARITY UNPINNED - this restates its resize_linear kernel from the
// published source as recollected: src = dst * (srcSize / dstSize) WITHOUT the half-pixel offset of the CPU cv::resize,
// x1 = floor, x2 = x1 + 1 (reads clamped to the last row / column), four float weights, accumulation in float in the
// order (y1,x1), (y1,x2), (y2,x1), (y2,x2), saturate_cast<uchar> = round to nearest even.  tests/test_oracle_cpu.py
// bounds the distance to CPU cv2.resize (which samples half a pixel further).
int orc_resize_bgr8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
    const float fx = (float)(1.0 / ((double)dw / (double)sw)), fy = (float)(1.0 / ((double)dh / (double)sh));
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            const float sx = (float)x * fx, sy = (float)y * fy;
            const int x1 = (int)std::floor(sx), y1 = (int)std::floor(sy);
            const int x2 = x1 + 1, y2 = y1 + 1;
            const int x2r = x2 < sw - 1 ? x2 : sw - 1, y2r = y2 < sh - 1 ? y2 : sh - 1;
            const int x1r = x1 < sw - 1 ? x1 : sw - 1, y1r = y1 < sh - 1 ? y1 : sh - 1;
            const float w11 = ((float)x2 - sx) * ((float)y2 - sy), w12 = (sx - (float)x1) * ((float)y2 - sy);
            const float w21 = ((float)x2 - sx) * (sy - (float)y1), w22 = (sx - (float)x1) * (sy - (float)y1);
            for (int c = 0; c < 3; ++c) {
                float out = 0.0f;
                out = out + (float)src[((size_t)y1r * sw + x1r) * 3 + c] * w11;
                out = out + (float)src[((size_t)y1r * sw + x2r) * 3 + c] * w12;
                out = out + (float)src[((size_t)y2r * sw + x1r) * 3 + c] * w21;
                out = out + (float)src[((size_t)y2r * sw + x2r) * 3 + c] * w22;
                const float r = std::nearbyint(out);  // default rounding mode: to nearest even
                dst[((size_t)y * dw + x) * 3 + c] = (uint8_t)(r < 0.0f ? 0.0f : (r > 255.0f ? 255.0f : r));
            }
        }
    return 0;
}

}  // extern "C"
