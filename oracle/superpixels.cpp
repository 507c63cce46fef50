// TEST INFRASTRUCTURE ONLY (see tile.hpp header).
//
// Scalar restatement of the reference's contour-relaxation superpixel refinement:
//   createBlockInitialization / performBlockIntialization  .../contourrelaxation/initialization.cu:12-59
//   ContourRelaxation::relax                               .../contourrelaxation/contourrelaxation.cu:349-447
//   findBorderPixels / performRelaxation / updateLabels    .../contourrelaxation/contourrelaxation.cu:146-301
//   getNeighbourLabels / calculateCliqueCost / calculateCost                     same file :72-144
//   CUDAGaussianFeature (Colour 3ch u8, Disparity 2ch s16) .../features/gaussian.cu:30-206
//   CUDACompactnessFeature                                 .../features/compactness.cu:28-197
//   featuresMinVariance = 1/12                             .../contourrelaxation/constants.hpp:35
// (paths relative to /root/reference/{src,include}/modules/superpixels).
// Canonical choices: Q12 statistics initialised over the floor(W/32)*32 x
// floor(H/32)*32 sub-rectangle only (reproduced); Q13 the stored per-label featureCost used for
// "third-party" neighbour labels is computed from the exact sums at the start of each iteration;
// Q14 unsigned pixelCount wrap reproduced; Q22 candidate order x-outer/y-inner, first minimum wins.
// Labels are integer results; costs are fp64 -> agreement (>= 99.9 %), not bit-exactness, is the
// contract between this oracle and the CUDA path.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "oracle.h"
#include "tile.hpp"

using orc::ceil_div;
using orc::Tile;

namespace {

const double kMinVariance = 1.0 / 12.0;
const uint16_t OUT_OF_BOUNDS = 1 << 14;  // contourrelaxation.cu:21

struct Stat {  // LabelStatisticsGauss, gaussian.cuh:18-28
    uint32_t n = 0;
    double sum = 0, sq = 0, cost = 0;
};

// The reference calls CUDA's device log() (gaussian.cu:41), whose last bits no CPU library reproduces.  The oracle
// therefore fixes a fully specified logarithm: the main path of fdlibm's e_log with IEEE +, -, *, / in a fixed order
// and no fused multiply-adds (this file is built with -ffp-contract=off).  Any implementation that keeps the order
// gets the same bits, which allows a bit-exact parity test of the label decisions.  Domain: x >= 2 pi / 12.
inline double detLog(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    uint64_t bits;
    std::memcpy(&bits, &x, 8);
    int32_t hx = (int32_t)(bits >> 32);
    const uint32_t lx = (uint32_t)bits;
    int32_t k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;
    const uint64_t mbits = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | lx;
    double m;
    std::memcpy(&m, &mbits, 8);
    k += i >> 20;
    const double f = m - 1.0;
    const double s = f / (2.0 + f);
    const double dk = (double)k;
    const double z = s * s;
    i = hx - 0x6147a;
    const double w = z * z;
    const int32_t j = 0x6b851 - hx;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    const double R = t2 + t1;
    if (i > 0) {
        const double hfsq = (0.5 * f) * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

inline void gaussCost(Stat& s) {  // deviceUpdateLabelFeatureCost, gaussian.cu:30-43
    const double n = (double)s.n;
    double variance = (s.sq / n) - ((s.sum / n) * (s.sum / n));
    variance = std::fmax(variance, kMinVariance);
    s.cost = (n / 2 * detLog(2 * M_PI * variance)) + (n / 2);
}
inline void compactCost(Stat& s) {  // updateCompactnessCost, compactness.cu:28-35
    if (s.n == 0) {
        s.cost = 0;
        return;
    }
    s.cost = s.sq - ((s.sum * s.sum) / (double)s.n);
}

struct Gauss {  // one Gaussian feature with C channels
    int C = 0;
    std::vector<Stat> st;  // [label][C]
    const uint8_t* u8 = nullptr;
    const int16_t* s16 = nullptr;
    int W = 0;
    double value(int x, int y, int ch) const {
        const size_t i = ((size_t)y * W + x) * C + ch;
        return u8 ? (double)u8[i] : (double)s16[i];
    }
    void refreshCosts() {
        for (auto& s : st)
            if (s.n != 0) gaussCost(s);
    }
    // gaussian.cu:96-174
    double cost(int x, int y, uint16_t oldL, uint16_t pretend, const uint16_t* nl, size_t nn) const {
        Stat o[3], p[3];
        for (int ch = 0; ch < C; ++ch) {
            o[ch] = st[(size_t)oldL * C + ch];
            p[ch] = st[(size_t)pretend * C + ch];
        }
        if (oldL != pretend) {
            for (int ch = 0; ch < C; ++ch) {
                o[ch].n--;  // Q14: unsigned wrap
                p[ch].n++;
                const double v = value(x, y, ch);
                o[ch].sum -= v;
                p[ch].sum += v;
                const double v2 = v * v;
                o[ch].sq -= v2;
                p[ch].sq += v2;
                gaussCost(o[ch]);
                gaussCost(p[ch]);
            }
        }
        double fc = 0;
        for (size_t i = 0; i < nn; ++i) {
            const Stat* row = &st[(size_t)nl[i] * C];
            if (nl[i] == oldL)
                row = o;
            else if (nl[i] == pretend)
                row = p;
            for (int ch = 0; ch < C; ++ch) {
                if (row[ch].n == 0) continue;
                fc += row[ch].cost;
            }
        }
        return fc / (double)C;
    }
    void move(int x, int y, uint16_t oldL, uint16_t newL) {  // gaussian.cu:176-206
        for (int ch = 0; ch < C; ++ch) {
            Stat& o = st[(size_t)oldL * C + ch];
            Stat& n = st[(size_t)newL * C + ch];
            o.n--;
            n.n++;
            const double v = value(x, y, ch);
            o.sum -= v;
            n.sum += v;
            o.sq -= v * v;
            n.sq += v * v;
        }
    }
};

struct Compact {
    std::vector<Stat> px, py;
    double progressive = 0;
    int H = 0;
    void refreshCosts() {
        for (auto& s : px) compactCost(s);
        for (auto& s : py) compactCost(s);
    }
    // compactness.cu:105-191
    double cost(int x, int y, uint16_t oldL, uint16_t pretend, const uint16_t* nl, size_t nn) const {
        Stat xo = px[oldL], xp = px[pretend], yo = py[oldL], yp = py[pretend];
        if (oldL != pretend) {
            xo.n--;
            xp.n++;
            xo.sum -= x;
            xp.sum += x;
            xo.sq -= x * x;
            xp.sq += x * x;
            yo.n--;
            yp.n++;
            yo.sum -= y;
            yp.sum += y;
            yo.sq -= y * y;
            yp.sq += y * y;
            compactCost(xo);
            compactCost(xp);
            compactCost(yo);
            compactCost(yp);
        }
        double fc = 0;
        for (size_t i = 0; i < nn; ++i) {
            const Stat* sx = &px[nl[i]];
            const Stat* sy = &py[nl[i]];
            if (nl[i] == oldL) {
                sx = &xo;
                sy = &yo;
            } else if (nl[i] == pretend) {
                sx = &xp;
                sy = &yp;
            }
            if (sx->n == 0) continue;
            fc += sx->cost + sy->cost;
        }
        if (progressive > 0.0) fc *= 1.0 + progressive * ((double)H - (double)y) / (double)H;
        return fc;
    }
    void move(int x, int y, uint16_t oldL, uint16_t newL) {  // compactness.cu:37-68
        px[oldL].n--;
        px[newL].n++;
        px[oldL].sum -= x;
        px[newL].sum += x;
        px[oldL].sq -= x * x;
        px[newL].sq += x * x;
        py[oldL].n--;
        py[newL].n++;
        py[oldL].sum -= y;
        py[newL].sum += y;
        py[oldL].sq -= y * y;
        py[newL].sq += y * y;
    }
};

}  // namespace

extern "C" {

// initialization.cu:12-59. Returns maxLabelId (a COUNT: numBlocksX * numBlocksY).
int orc_block_init(int W, int H, int bw, int bh, uint16_t* labels) {
    const int perRow = ceil_div(W, bw);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) labels[(size_t)y * W + x] = (uint16_t)((y / bh) * perRow + x / bw);
    return perRow * ceil_div(H, bh);
}

// findBorderPixels' membership test (contourrelaxation.cu:146-219) on the bug-compatible label tile.
// border[y*W+x] = 1 if the reference would list the pixel.  defined (nullable) as in tile.hpp.
int orc_border_map(const uint16_t* labels, int W, int H, uint8_t* border, uint8_t* defined) {
    const int BD = 16, XB = 4, YB = 4, TW = 64, TH = 64;
    const size_t alloc = (size_t)(XB * (2 + BD)) * (YB * (2 + BD));  // SHARED_SIZE(4,4) = 72*72, :17
    Tile<uint16_t> t;
    for (int by = 0; by < ceil_div(H, TH); ++by)
        for (int bx = 0; bx < ceil_div(W, TW); ++bx) {
            orc::copy_to_shared<uint16_t, false>(t, labels, W, H, bx, by, BD, BD, XB, YB, 1, 1, alloc, (uint16_t)0xFFFF);
            for (int ly = 0; ly < TH; ++ly)
                for (int lx = 0; lx < TW; ++lx) {
                    const int x = bx * TW + lx, y = by * TH + ly;
                    if (x >= W || y >= H) continue;
                    bool d, dall;
                    const uint16_t l = t.get(lx, ly, &dall);
                    bool b = false;
                    for (int k = -1; k <= 1; ++k)
                        for (int q = -1; q <= 1; ++q) {
                            if (k == 0 && q == 0) continue;
                            if (t.get(lx + k, ly + q, &d) != l) b = true;
                            dall = dall && d;
                        }
                    border[(size_t)y * W + x] = b;
                    if (defined) defined[(size_t)y * W + x] = dall;
                }
        }
    return 0;
}

// ContourRelaxation::relax (contourrelaxation.cu:349-447) for one frame.
// labels: persistent label image, updated in place. ycrcb: [H][W][3] u8. deriv2: [H][W][2] s16 or null
// (null <=> disparity weight <= 0).  Weights <= 0 disable a feature (contourrelaxation.hpp:54-61).
// Feature order: Compactness, Disparity, Colour (superpixels.cu:61-68).
// borderCounts (nullable): number of listed border pixels per iteration; moved (nullable): moves per iteration.
int orc_sp_relax(uint16_t* labels, int W, int H, int maxLabel, const uint8_t* ycrcb, const int16_t* deriv2,
                 int iterations, double directCost, double diagCost, double wCompact, double progressive, double wDisp,
                 double wImage, int32_t* borderCounts, int32_t* moved) {
    if (maxLabel >= OUT_OF_BOUNDS) return -2;
    const bool useC = wCompact > 0, useD = wDisp > 0 && deriv2 != nullptr, useI = wImage > 0;
    if (wDisp > 0 && !deriv2) return -1;
    Compact comp;
    Gauss gd, gi;
    const size_t NL = (size_t)maxLabel + 1;
    if (useC) {
        comp.px.assign(NL, Stat());
        comp.py.assign(NL, Stat());
        comp.progressive = progressive;
        comp.H = H;
    }
    if (useD) {
        gd.C = 2;
        gd.st.assign(NL * 2, Stat());
        gd.s16 = deriv2;
        gd.W = W;
    }
    if (useI) {
        gi.C = 3;
        gi.st.assign(NL * 3, Stat());
        gi.u8 = ycrcb;
        gi.W = W;
    }
    // initializeStatisticsKernel over floor(W/32) x floor(H/32) blocks of 32x32 (Q12), :379-381
    const int WS = (W / 32) * 32, HS = (H / 32) * 32;
    for (int y = 0; y < HS; ++y)
        for (int x = 0; x < WS; ++x) {
            const uint16_t l = labels[(size_t)y * W + x];
            if (l > maxLabel) return -2;
            if (useC) {
                comp.px[l].n++;
                comp.px[l].sum += x;
                comp.px[l].sq += (double)(x * x);
                comp.py[l].n++;
                comp.py[l].sum += y;
                comp.py[l].sq += (double)(y * y);
            }
            if (useD)
                for (int ch = 0; ch < 2; ++ch) {
                    Stat& s = gd.st[(size_t)l * 2 + ch];
                    const double v = gd.value(x, y, ch);
                    s.n++;
                    s.sum += v;
                    s.sq += v * v;
                }
            if (useI)
                for (int ch = 0; ch < 3; ++ch) {
                    Stat& s = gi.st[(size_t)l * 3 + ch];
                    const double v = gi.value(x, y, ch);
                    s.n++;
                    s.sum += v;
                    s.sq += v * v;
                }
        }

    std::vector<uint8_t> border((size_t)W * H);
    std::vector<uint16_t> newLabels((size_t)W * H);
    for (int it = 0; it < iterations; ++it) {
        if (useC) comp.refreshCosts();  // Q13 canonical: stored costs from the exact sums
        if (useD) gd.refreshCosts();
        if (useI) gi.refreshCosts();
        orc_border_map(labels, W, H, border.data(), nullptr);
        int nb = 0, nm = 0;
        // decide (Jacobi), contourrelaxation.cu:221-276
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                if (!border[(size_t)y * W + x]) continue;
                nb++;
                uint16_t nbh[9];
                for (int ox = -1; ox <= 1; ++ox)
                    for (int oy = -1; oy <= 1; ++oy) {
                        const int xc = x + ox, yc = y + oy;
                        nbh[(ox + 1) + (oy + 1) * 3] =
                            (xc < 0 || yc < 0 || xc >= W || yc >= H) ? OUT_OF_BOUNDS : labels[(size_t)yc * W + xc];
                    }
                uint16_t nl[9];
                size_t nn = 0;
                for (int i = -1; i <= 1; ++i)
                    for (int j = -1; j <= 1; ++j) {
                        const uint16_t l = nbh[(i + 1) + (j + 1) * 3];
                        if (l == OUT_OF_BOUNDS) continue;
                        bool found = false;
                        for (size_t k = 0; k < nn; ++k)
                            if (nl[k] == l) {
                                found = true;
                                break;
                            }
                        if (!found) nl[nn++] = l;
                    }
                const uint16_t cur = nbh[4];
                double minCost = DBL_MAX;
                uint16_t best = cur;
                for (size_t c = 0; c < nn; ++c) {
                    const uint16_t pl = nl[c];
                    auto clique = [&](int ox, int oy) {
                        const uint16_t n = nbh[(ox + 1) + (oy + 1) * 3];
                        return (int)(n != OUT_OF_BOUNDS && n != pl);
                    };
                    const int nd = clique(-1, 0) + clique(1, 0) + clique(0, -1) + clique(0, 1);
                    const int ng = clique(-1, -1) + clique(-1, 1) + clique(1, -1) + clique(1, 1);
                    double cost = nd * directCost + ng * diagCost;
                    if (useC) cost += wCompact * comp.cost(x, y, cur, pl, nl, nn);
                    if (useD) cost += wDisp * gd.cost(x, y, cur, pl, nl, nn);
                    if (useI) cost += wImage * gi.cost(x, y, cur, pl, nl, nn);
                    if (cost < minCost) {
                        minCost = cost;
                        best = pl;
                    }
                }
                newLabels[(size_t)y * W + x] = best;
            }
        // apply, contourrelaxation.cu:278-301
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                if (!border[(size_t)y * W + x]) continue;
                const uint16_t cur = labels[(size_t)y * W + x], nw = newLabels[(size_t)y * W + x];
                if (cur == nw) continue;
                nm++;
                if (useC) comp.move(x, y, cur, nw);
                if (useD) gd.move(x, y, cur, nw);
                if (useI) gi.move(x, y, cur, nw);
                labels[(size_t)y * W + x] = nw;
            }
        if (borderCounts) borderCounts[it] = nb;
        if (moved) moved[it] = nm;
    }
    return 0;
}

}  // extern "C"
