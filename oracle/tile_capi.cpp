// TEST INFRASTRUCTURE ONLY (see tile.hpp header). C entry point for the tile-loader unit tests.
#include "oracle.h"
#include "tile.hpp"

extern "C" int orc_tile_i32(const int32_t* img, int W, int H, int bx, int by, int bdx, int bdy, int XB, int YB,
                            int yPad, int xPad, int interp, long allocElems, int32_t undef, int32_t* out,
                            uint8_t* def) {
    orc::Tile<int32_t> t;
    if (interp)
        orc::copy_to_shared<int32_t, true>(t, img, W, H, bx, by, bdx, bdy, XB, YB, yPad, xPad, (size_t)allocElems, undef);
    else
        orc::copy_to_shared<int32_t, false>(t, img, W, H, bx, by, bdx, bdy, XB, YB, yPad, xPad, (size_t)allocElems, undef);
    const int S = t.tileW + 2 * xPad, R = t.tileH + 2 * yPad;
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < S; ++c) {
            bool d;
            out[(size_t)r * S + c] = t.get(c - xPad, r - yPad, &d);
            def[(size_t)r * S + c] = d;
        }
    return 0;
}
