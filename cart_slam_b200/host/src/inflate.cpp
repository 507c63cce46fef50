// zlib-stream decoder of the PNG reader (see cart/inflate.hpp).  RFC 1950 (container, Adler-32) and RFC 1951 (stored,
// fixed and dynamic Huffman blocks).  Written for whole-buffer decoding with a known output size.
#include "cart/inflate.hpp"

#include <zlib.h>  // adler32() only

#include <cstring>

namespace cart {
namespace png {
namespace {

// ---- decode tables ---------------------------------------------------------------------------------------------------
// Entry (32 bits): [7:0] bits to consume, [12:8] extra bits (length / distance entries) or index bits of a subtable,
// [13] literal, [14] end of block, [15] subtable pointer, [31:16] literal / base value / first entry of the subtable.
// An all-zero entry is an unused code: hitting it is a format error.
constexpr uint32_t kLit = 1u << 13, kEob = 1u << 14, kSub = 1u << 15;
constexpr int kLitBits = 11, kDistBits = 8, kPreBits = 7;
constexpr int kLitCap = (1 << kLitBits) + 288 * 16;   // every long code in a subtable of its own (15 - 11 = 4 bits) at worst
constexpr int kDistCap = (1 << kDistBits) + 32 * 128;  // 15 - 8 = 7 bits
constexpr int kPreCap = 1 << kPreBits;

const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

enum Alphabet { kLitLen, kDistance, kPreCode };

inline uint32_t symbolEntry(Alphabet a, int sym) {  // entry without its bit count; 0 = a symbol that must not occur
    switch (a) {
        case kLitLen:
            if (sym < 256) return kLit | ((uint32_t)sym << 16);
            if (sym == 256) return kEob;
            if (sym <= 285) return ((uint32_t)kLenBase[sym - 257] << 16) | ((uint32_t)kLenExtra[sym - 257] << 8);
            return 0;
        case kDistance:
            if (sym < 30) return ((uint32_t)kDistBase[sym] << 16) | ((uint32_t)kDistExtra[sym] << 8);
            return 0;
        default:
            return kLit | ((uint32_t)sym << 16);
    }
}

inline uint32_t reverseBits(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}

// canonical Huffman code of `n` symbols with lengths lens[] (0 = unused) -> two-level table indexed by the next bits
// of the stream, least significant bit first.  Fails on an over-subscribed code; an incomplete code is accepted (its
// unused patterns stay zero entries).
bool buildTable(Alphabet a, const uint8_t* lens, int n, uint32_t* table, int primaryBits, int cap) {
    int count[16] = {0};
    for (int s = 0; s < n; ++s) count[lens[s]]++;
    count[0] = 0;
    int left = 1;
    for (int l = 1; l <= 15; ++l) {
        left = (left << 1) - count[l];
        if (left < 0) return false;
    }
    uint32_t next[16];
    {
        uint32_t code = 0;
        for (int l = 1; l <= 15; ++l) {
            code = (code + (uint32_t)count[l - 1]) << 1;
            next[l] = code;
        }
    }
    const int primarySize = 1 << primaryBits;
    const uint32_t primaryMask = (uint32_t)primarySize - 1;
    std::memset(table, 0, sizeof(uint32_t) * (size_t)primarySize);
    uint16_t rcode[288];
    uint8_t subBits[1 << kLitBits];  // per primary index: index bits of its subtable (0 = none)
    bool anyLong = false;
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        rcode[s] = (uint16_t)reverseBits(next[l]++, l);
        if (l > primaryBits) anyLong = true;
    }
    if (anyLong) {
        std::memset(subBits, 0, (size_t)primarySize);
        for (int s = 0; s < n; ++s) {
            const int l = lens[s];
            if (l > primaryBits) {
                uint8_t& b = subBits[rcode[s] & primaryMask];
                if (l - primaryBits > b) b = (uint8_t)(l - primaryBits);
            }
        }
        int nextFree = primarySize;
        for (int p = 0; p < primarySize; ++p) {
            if (!subBits[p]) continue;
            const int size = 1 << subBits[p];
            if (nextFree + size > cap) return false;
            std::memset(table + nextFree, 0, sizeof(uint32_t) * (size_t)size);
            table[p] = kSub | ((uint32_t)nextFree << 16) | ((uint32_t)subBits[p] << 8) | (uint32_t)primaryBits;
            nextFree += size;
        }
    }
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t e = symbolEntry(a, s);
        if (l <= primaryBits) {
            if (!e) continue;  // stays an error entry
            for (uint32_t i = rcode[s]; i < (uint32_t)primarySize; i += 1u << l) table[i] = e | (uint32_t)l;
        } else {
            const uint32_t p = rcode[s] & primaryMask;
            const uint32_t start = table[p] >> 16, sb = subBits[p];
            if (!e) continue;
            for (uint32_t i = (uint32_t)rcode[s] >> primaryBits; i < (1u << sb); i += 1u << (l - primaryBits))
                table[start + i] = e | (uint32_t)(l - primaryBits);
        }
    }
    return true;
}

inline uint64_t load64(const uint8_t* p) {
    uint64_t v;
    std::memcpy(&v, p, 8);
#if defined(__BYTE_ORDER__) && __BYTE_ORDER__ == __ORDER_BIG_ENDIAN__
    v = __builtin_bswap64(v);
#endif
    return v;
}

struct Tables {
    uint32_t lit[kLitCap];
    uint32_t dist[kDistCap];
    uint32_t pre[kPreCap];
};

}  // namespace

bool inflateZlib(const uint8_t* in, size_t inSize, uint8_t* out, size_t outSize) {
    if (inSize < 2 + 4) return false;
    // RFC 1950 header: deflate, window <= 32 KiB, no preset dictionary, check bits
    if ((in[0] & 0x0F) != 8 || (in[0] >> 4) > 7 || (in[1] & 0x20) || ((in[0] << 8) | in[1]) % 31 != 0) return false;
    const uint8_t* const base = in;
    const uint8_t* const hardLimit = in + inSize + 8;  // a refill at or below this address stays inside the padding
    const uint8_t* ip = in + 2;
    uint8_t* op = out;
    uint8_t* const outEnd = out + outSize;
    uint64_t bits = 0;
    unsigned cnt = 0;
    static thread_local Tables* tables = nullptr;
    if (!tables) tables = new Tables;
    Tables& T = *tables;
    bool last = false;

#define REFILL()                         \
    do {                                 \
        bits |= load64(ip) << cnt;       \
        ip += (63 - cnt) >> 3;           \
        cnt |= 56;                       \
    } while (0)
#define DROP(n)     \
    do {            \
        bits >>= (n); \
        cnt -= (n); \
    } while (0)

    while (!last) {
        if (ip > hardLimit) return false;
        REFILL();
        last = bits & 1;
        const unsigned type = (unsigned)(bits >> 1) & 3;
        DROP(3);
        if (type == 0) {  // stored
            DROP(cnt & 7);
            const uint8_t* p = ip - (cnt >> 3);
            if (p + 4 > base + inSize) return false;
            const unsigned len = p[0] | (p[1] << 8), nlen = p[2] | (p[3] << 8);
            if ((len ^ nlen) != 0xFFFFu) return false;
            p += 4;
            if ((size_t)(base + inSize - p) < len || (size_t)(outEnd - op) < len) return false;
            std::memcpy(op, p, len);
            op += len;
            ip = p + len;
            bits = 0;
            cnt = 0;
            continue;
        }
        if (type == 3) return false;
        if (type == 1) {  // fixed code
            {  // (rebuilt per block: a dynamic block in between overwrites the tables; fixed blocks are rare in PNG files)
                uint8_t lens[288 + 32];
                for (int i = 0; i < 144; ++i) lens[i] = 8;
                for (int i = 144; i < 256; ++i) lens[i] = 9;
                for (int i = 256; i < 280; ++i) lens[i] = 7;
                for (int i = 280; i < 288; ++i) lens[i] = 8;
                for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
                if (!buildTable(kLitLen, lens, 288, T.lit, kLitBits, kLitCap) ||
                    !buildTable(kDistance, lens + 288, 32, T.dist, kDistBits, kDistCap))
                    return false;
            }
        } else {  // dynamic code
            const unsigned hlit = (unsigned)(bits & 31) + 257, hdist = (unsigned)((bits >> 5) & 31) + 1, hclen = (unsigned)((bits >> 10) & 15) + 4;
            DROP(14);
            if (hlit > 286 || hdist > 30) return false;
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t pre[19] = {0};
            for (unsigned i = 0; i < hclen; ++i) {
                if (cnt < 3) REFILL();
                pre[order[i]] = (uint8_t)(bits & 7);
                DROP(3);
            }
            if (!buildTable(kPreCode, pre, 19, T.pre, kPreBits, kPreCap)) return false;
            uint8_t lens[286 + 30 + 138];
            unsigned n = 0;
            const unsigned total = hlit + hdist;
            while (n < total) {
                if (ip > hardLimit) return false;
                REFILL();
                const uint32_t e = T.pre[bits & (kPreCap - 1)];
                if (!e) return false;
                DROP(e & 0xFF);
                const unsigned sym = e >> 16;
                if (sym < 16) {
                    lens[n++] = (uint8_t)sym;
                } else {
                    unsigned rep;
                    uint8_t v = 0;
                    if (sym == 16) {
                        if (n == 0) return false;
                        v = lens[n - 1];
                        rep = 3 + (unsigned)(bits & 3);
                        DROP(2);
                    } else if (sym == 17) {
                        rep = 3 + (unsigned)(bits & 7);
                        DROP(3);
                    } else {
                        rep = 11 + (unsigned)(bits & 127);
                        DROP(7);
                    }
                    if (n + rep > total) return false;
                    std::memset(lens + n, v, rep);
                    n += rep;
                }
            }
            if (lens[256] == 0) return false;  // no end-of-block code
            if (!buildTable(kLitLen, lens, (int)hlit, T.lit, kLitBits, kLitCap) ||
                !buildTable(kDistance, lens + hlit, (int)hdist, T.dist, kDistBits, kDistCap))
                return false;
        }
        // ---- symbols of the block ----
        // Fast loop: while at least 4 literals or one maximal match (258 bytes, copied in 8-byte steps) fit without a test
        // and the input pointer is inside the buffer; per iteration one refill, then up to four literals (15 + 3 x 11
        // bits) or one match (<= 48 bits).  It leaves through `break` (end of block) or when a bound comes near; the
        // careful loop below finishes the block.
        bool endOfBlock = false;
        if (outSize >= 300) {
            uint8_t* const outFast = outEnd - 300;
            constexpr uint32_t litMask = (1u << kLitBits) - 1, distMask = (1u << kDistBits) - 1;
            while (op < outFast && ip <= hardLimit) {
                REFILL();
                uint32_t e = T.lit[bits & litMask];
                if (e & kSub) {
                    DROP(kLitBits);
                    e = T.lit[(e >> 16) + (bits & ((1u << ((e >> 8) & 31)) - 1))];
                }
                if (e & kLit) {
                    DROP(e & 0xFF);
                    uint32_t e2 = T.lit[bits & litMask];
                    *op++ = (uint8_t)(e >> 16);
                    if (!(e2 & kLit)) continue;
                    DROP(e2 & 0xFF);
                    e = T.lit[bits & litMask];
                    *op++ = (uint8_t)(e2 >> 16);
                    if (!(e & kLit)) continue;
                    DROP(e & 0xFF);
                    e2 = T.lit[bits & litMask];
                    *op++ = (uint8_t)(e >> 16);
                    if (!(e2 & kLit)) continue;
                    DROP(e2 & 0xFF);
                    *op++ = (uint8_t)(e2 >> 16);
                    continue;
                }
                if (!e) return false;
                DROP(e & 0xFF);
                if (e & kEob) {
                    endOfBlock = true;
                    break;
                }
                const unsigned lx = (e >> 8) & 31;
                const size_t len = (e >> 16) + (size_t)(bits & ((1u << lx) - 1));
                DROP(lx);
                uint32_t d = T.dist[bits & distMask];
                if (d & kSub) {
                    DROP(kDistBits);
                    d = T.dist[(d >> 16) + (bits & ((1u << ((d >> 8) & 31)) - 1))];
                }
                if (!d) return false;
                DROP(d & 0xFF);
                const unsigned dx = (d >> 8) & 31;
                const size_t dist = (d >> 16) + (size_t)(bits & ((1u << dx) - 1));
                DROP(dx);
                if (dist > (size_t)(op - out)) return false;
                const uint8_t* src = op - dist;
                uint8_t* const end = op + len;
                if (dist >= 8) {
                    std::memcpy(op, src, 8);
                    std::memcpy(op + 8, src + 8, 8);
                    if (len > 16) {
                        op += 16;
                        src += 16;
                        do {
                            std::memcpy(op, src, 8);
                            op += 8;
                            src += 8;
                        } while (op < end);
                    }
                } else if (dist == 1) {
                    std::memset(op, *src, len);
                } else {
                    do {
                        *op++ = *src++;
                    } while (op < end);
                }
                op = end;
            }
        }
        if (endOfBlock) continue;
        for (;;) {
            if (ip > hardLimit) return false;
            REFILL();  // >= 56 bits: a literal/length code, its extra bits, a distance code and its extra bits need <= 48
            uint32_t e = T.lit[bits & ((1u << kLitBits) - 1)];
            if (e & kSub) {
                DROP(kLitBits);
                e = T.lit[(e >> 16) + (bits & ((1u << ((e >> 8) & 31)) - 1))];
            }
            if (e & kLit) {  // up to three literals per refill (3 x 15 bits)
                if (op >= outEnd) return false;
                DROP(e & 0xFF);
                *op++ = (uint8_t)(e >> 16);
                e = T.lit[bits & ((1u << kLitBits) - 1)];
                if (!(e & kLit) || op >= outEnd) continue;
                DROP(e & 0xFF);
                *op++ = (uint8_t)(e >> 16);
                e = T.lit[bits & ((1u << kLitBits) - 1)];
                if (!(e & kLit) || op >= outEnd) continue;
                DROP(e & 0xFF);
                *op++ = (uint8_t)(e >> 16);
                continue;
            }
            if (!e) return false;
            DROP(e & 0xFF);
            if (e & kEob) break;
            const unsigned lx = (e >> 8) & 31;
            const size_t len = (e >> 16) + (size_t)(bits & ((1u << lx) - 1));
            DROP(lx);
            uint32_t d = T.dist[bits & ((1u << kDistBits) - 1)];
            if (d & kSub) {
                DROP(kDistBits);
                d = T.dist[(d >> 16) + (bits & ((1u << ((d >> 8) & 31)) - 1))];
            }
            if (!d) return false;
            DROP(d & 0xFF);
            const unsigned dx = (d >> 8) & 31;
            const size_t dist = (d >> 16) + (size_t)(bits & ((1u << dx) - 1));
            DROP(dx);
            if (dist > (size_t)(op - out) || len > (size_t)(outEnd - op)) return false;
            const uint8_t* src = op - dist;
            uint8_t* const end = op + len;
            if (dist >= 8) {  // 8 bytes at a time; may write up to 7 bytes past `end` (inside the buffer or its padding)
                do {
                    std::memcpy(op, src, 8);
                    op += 8;
                    src += 8;
                } while (op < end);
            } else if (dist == 1) {
                std::memset(op, *src, len);
            } else {
                do {
                    *op++ = *src++;
                } while (op < end);
            }
            op = end;
        }
    }
#undef REFILL
#undef DROP
    // the stream ends at the next byte boundary; the Adler-32 of the output follows, big endian
    const uint8_t* tail = ip - (cnt >> 3);
    if (op != outEnd || tail + 4 > base + inSize) return false;
    const uint32_t want = ((uint32_t)tail[0] << 24) | ((uint32_t)tail[1] << 16) | ((uint32_t)tail[2] << 8) | tail[3];
    uLong a = adler32(0L, Z_NULL, 0);
    for (size_t off = 0; off < outSize;) {  // uInt lengths
        const size_t n = outSize - off < (1u << 30) ? outSize - off : (1u << 30);
        a = adler32(a, out + off, (uInt)n);
        off += n;
    }
    return (uint32_t)a == want;
}

}  // namespace png
}  // namespace cart
