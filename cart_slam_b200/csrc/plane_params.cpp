// Host-side plane-parameter estimation for the `histogram_peak` provider:
//   HistogramPeakPlaneParameterProvider::updatePlaneParameters  /root/reference/src/modules/planeseg/planeseg.cu:405-458
//   util::findPeaks (persistent-homology 1-D peak detection)    /root/reference/src/utils/peaks.cpp:12-72
// 256 integers per update: not a GPU target.  The two std::sort calls use the reference's comparators on
// the same initial sequences, so ties resolve exactly as they do in a reference built with the same libstdc++.
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace cb {

namespace {
struct Component {  // one connected component of a superlevel set
    int born;       // index of its maximum
    int lo, hi;     // extent
    int died;       // index where it merged into an older component, -1 = never
};

inline int persistence(const Component& c, const int32_t* h) { return c.died < 0 ? INT_MAX : h[c.born] - h[c.died]; }

std::vector<Component> persistentPeaks(const int32_t* h, int n) {
    std::vector<int> order(n), owner(n, -1);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [h](int a, int b) { return h[a] > h[b]; });
    std::vector<Component> comps;
    for (int idx : order) {
        const int L = (idx > 0) ? owner[idx - 1] : -1;
        const int Rr = (idx < n - 1) ? owner[idx + 1] : -1;
        if (L < 0 && Rr < 0) {
            comps.push_back(Component{idx, idx, idx, -1});
            owner[idx] = (int)comps.size() - 1;
        } else if (L >= 0 && Rr < 0) {
            comps[L].hi += 1;
            owner[idx] = L;
        } else if (L < 0) {
            comps[Rr].lo -= 1;
            owner[idx] = Rr;
        } else if (h[comps[L].born] > h[comps[Rr].born]) {  // the older (higher) component survives
            comps[Rr].died = idx;
            comps[L].hi = comps[Rr].hi;
            owner[comps[L].hi] = owner[idx] = L;
        } else {
            comps[L].died = idx;
            comps[Rr].lo = comps[L].lo;
            owner[comps[Rr].lo] = owner[idx] = Rr;
        }
    }
    std::sort(comps.begin(), comps.end(),
              [h](Component a, Component b) { return persistence(a, h) > persistence(b, h); });
    return comps;
}
}  // namespace

// params = {horizontalCenter, verticalCenter, hStart, hEnd, vStart, vEnd}. Returns 1 when the ranges changed.
int histogram_peak_update(const int32_t* hist, int32_t* params) {
    std::vector<Component> pk = persistentPeaks(hist, 256);
    if (pk.size() < 2) return 0;
    int v = pk[0].born, hz = pk[1].born;  // vertical = the peak closer to bin 128
    if (std::abs(v - 128) > std::abs(hz - 128)) std::swap(v, hz);
    params[1] = v - 128;
    params[0] = hz - 128;
    const int from = std::min(v, hz), to = std::max(v, hz);
    int valley = from;
    for (int i = from; i < to; ++i)
        if (hist[i] < hist[valley]) valley = i;
    const int vDist = std::abs(valley - v), hDist = std::abs(valley - hz);
    if (vDist == 0 || hDist == 0) return 0;
    const int vSlope = (hist[v] - hist[valley]) / vDist, hSlope = (hist[hz] - hist[valley]) / hDist;
    if (vSlope == 0 || hSlope == 0) return 0;
    const int vWidth = hist[v] / vSlope, hWidth = hist[hz] / hSlope;
    params[4] = v - vWidth - 128;
    params[5] = valley - 127;
    params[2] = valley - 127;
    params[3] = hz + hWidth - 127;
    return 1;
}

}  // namespace cb
