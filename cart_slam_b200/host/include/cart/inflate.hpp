// zlib-stream (RFC 1950 / 1951) decoder for the PNG reader of the KITTI source (src/sources/kitti.cpp:155-172 reads
// the files with cv::imread, i.e. libpng + zlib).  PNG decoding is what bounds the on-disk source (≈ 8 ms per KITTI
// image and core with zlib's inflate); this decoder is written for that one job: the whole input and the exact output
// size are known up front, so the hot loop works on a 64-bit bit buffer refilled by unaligned 8-byte loads, decodes
// literals and matches from two-level tables (11-bit / 8-bit primary) and copies matches in 8-byte steps, with one
// bounds test per symbol.  Output identical to zlib's (tests/test_sources.py compares both on random streams of every
// block type); the Adler-32 trailer is verified.
#pragma once
#include <cstddef>
#include <cstdint>

namespace cart {
namespace png {

// Decodes the zlib stream [in, in + inSize) into exactly outSize bytes at out.  Both buffers must be followed by
// kInflatePad readable (in) / writable (out) bytes.  Returns true on success; false on any malformed, truncated or
// mis-sized stream (nothing is thrown; out may have been written partially).
constexpr size_t kInflatePad = 16;
bool inflateZlib(const uint8_t* in, size_t inSize, uint8_t* out, size_t outSize);

}  // namespace png
}  // namespace cart
