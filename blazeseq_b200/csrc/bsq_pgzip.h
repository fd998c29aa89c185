// Parallel decoder for ordinary (single- or multi-member) gzip streams on host threads: the role of
// RapidgzipReader(parallelism) in the reference (blazeseq/io/readers.mojo:380-443; rapidgzip-mojo is an
// un-vendored dependency, so the algorithm follows rapidgzip's published two-stage scheme):
//
//   stage 1 (parallel, speculative)   the compressed file is cut into chunks at arbitrary byte offsets.  A worker
//       searches its chunk for the first bit offset that parses as a non-final dynamic-Huffman block header
//       (complete code-length code, valid repeat codes, complete literal/length code with an end-of-block symbol,
//       complete distance code) and inflates from there WITHOUT knowing the 32 KiB window before it: output
//       symbols are 16 bits wide, values >= 0x8000 are *markers* "byte i of the unknown window".  Once 32 KiB of
//       output hold no marker (or a new gzip member starts) the worker switches to plain byte output.  It stops at
//       the first block boundary at or after the next chunk's nominal start at which another such header begins.
//   sequencing (serial, cheap)        chunks are accepted in order: a chunk whose start is not where its
//       predecessor stopped is decoded again from the right bit (a false positive of the block finder, or a long
//       run of stored / fixed blocks); the 32 KiB window is handed from chunk to chunk.
//   stage 2 (parallel)                markers are replaced with the window bytes and the CRC-32 of every
//       member segment is computed; the sequencer folds the segment CRCs (crc32_combine) and checks each member's
//       CRC-32 and ISIZE.
//
// Host-only C++17 + zlib (crc32 / crc32_combine).  No CUDA in this file; bsq_capi.cu includes it for the stream
// reader thread and for the bsq_gzip_* entry points.
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace bsq_pgz {

constexpr uint32_t kWin = 32768;          // DEFLATE window
constexpr uint16_t kMarker = 0x8000;      // marker i = kMarker | i, i = index into the unknown window (0 = oldest byte)

// ---- bit reader (LSB first, RFC 1951 3.1.1) -----------------------------------------------------------------
struct BitReader {
    const uint8_t* base = nullptr;
    const uint8_t* ip = nullptr;
    const uint8_t* end = nullptr;
    uint64_t buf = 0;
    uint32_t cnt = 0;        // valid bits in buf
    uint32_t over = 0;       // zero bits supplied past the end of the input

    void init(const uint8_t* b, size_t n, uint64_t bit) {
        base = b; end = b + n; ip = b + (bit >> 3); buf = 0; cnt = 0; over = 0;
        if (ip > end) ip = end;
        refill();
        drop((uint32_t)(bit & 7));
    }
    inline void refill() {
        if (ip + 8 <= end) {
            uint64_t w;
            memcpy(&w, ip, 8);
            buf |= w << cnt;
            ip += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56) {
                if (ip < end) buf |= (uint64_t)*ip++ << cnt; else over += 8;
                cnt += 8;
            }
        }
    }
    inline uint32_t peek(uint32_t n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    inline void drop(uint32_t n) { buf >>= n; cnt -= n; }
    inline uint32_t take(uint32_t n) { const uint32_t v = peek(n); drop(n); return v; }
    // bits supplied beyond the end and already consumed?  (cnt < over: part of the padding was used)
    inline bool overrun() const { return over > cnt; }
    inline uint64_t bitpos() const { return (uint64_t)(ip - base) * 8 + over - cnt; }
    void align_byte() { drop(cnt & 7); }
};

// ---- Huffman tables --------------------------------------------------------------------------------------------
// entry: val (literal byte / base length / base distance / subtable start), len (code length; for a subtable pointer
// the subtable's index bits), kind | extra_bits << 4
struct Entry { uint16_t val; uint8_t len; uint8_t kx; };
enum : uint8_t { K_LIT = 0, K_LEN = 1, K_EOB = 2, K_SUB = 3, K_BAD = 4 };

constexpr int kLitBits = 11, kDistBits = 8;
constexpr int kLitSize = (1 << kLitBits) + 288 * 16, kDistSize = (1 << kDistBits) + 32 * 128;

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t rev_bits(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

// Canonical code check as zlib's inflate_table: over-subscribed -> invalid; incomplete -> invalid unless the longest
// code is 1 bit (a single code) and `allow_single`.  Returns the longest code length, -1 when invalid, 0 when no codes.
inline int check_lengths(const uint8_t* lens, int n, bool allow_single, uint16_t count[16]) {
    for (int i = 0; i < 16; ++i) count[i] = 0;
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    int maxl = 15;
    while (maxl > 0 && count[maxl] == 0) --maxl;
    if (maxl == 0) return 0;
    int left = 1;
    for (int l = 1; l <= 15; ++l) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return -1;
    }
    if (left > 0 && !(allow_single && maxl == 1)) return -1;
    return maxl;
}

// make(sym) fills val / kx for a symbol.  `root` = primary index bits.
template <class MakeEntry>
inline void build_table(const uint8_t* lens, int n, const uint16_t count[16], int root, Entry* T, uint8_t* subneed, MakeEntry make) {
    const int psize = 1 << root;
    const Entry bad{0, 1, K_BAD};   // the decoder stops on K_BAD
    for (int i = 0; i < psize; ++i) T[i] = bad;
    // canonical codes (RFC 1951 3.2.2), bit-reversed: the first bit read is the code's top bit
    uint16_t next[16], codes[320];
    {
        uint32_t c = 0;
        next[0] = 0;
        for (int l = 1; l <= 15; ++l) { c = (c + (l > 1 ? count[l - 1] : 0)) << 1; next[l] = (uint16_t)c; }
    }
    bool any_long = false;
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (l == 0) continue;
        const uint32_t r = rev_bits(next[l]++, l);
        codes[s] = (uint16_t)r;
        if (l > root) { any_long = true; continue; }
        Entry e = make(s);
        e.len = (uint8_t)l;
        for (uint32_t i = r; i < (uint32_t)psize; i += 1u << l) T[i] = e;
    }
    if (!any_long) return;
    // long codes: one subtable per primary prefix, as wide as its longest code needs
    memset(subneed, 0, (size_t)psize);
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (l <= root) continue;
        uint8_t& need = subneed[codes[s] & (uint32_t)(psize - 1)];
        need = std::max<uint8_t>(need, (uint8_t)(l - root));
    }
    uint32_t off = (uint32_t)psize;
    for (int i = 0; i < psize; ++i)
        if (subneed[i]) {
            T[i] = Entry{(uint16_t)off, subneed[i], K_SUB};
            for (uint32_t k = 0; k < (1u << subneed[i]); ++k) T[off + k] = bad;
            off += 1u << subneed[i];
        }
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (l <= root) continue;
        const uint32_t r = codes[s];
        const Entry ptr = T[r & (uint32_t)(psize - 1)];
        Entry e = make(s);
        e.len = (uint8_t)l;   // full code length: the decoder drops it in one go
        for (uint32_t i = r >> root; i < (1u << ptr.len); i += 1u << (l - root)) T[ptr.val + i] = e;
    }
}

struct Tables {
    Entry lit[kLitSize];
    Entry dist[kDistSize];
    uint8_t subneed[1 << kLitBits];
    bool have_dist = false;
};

inline Entry make_lit(int s) {
    if (s < 256) return Entry{(uint16_t)s, 0, K_LIT};
    if (s == 256) return Entry{0, 0, K_EOB};
    if (s < 286) return Entry{kLenBase[s - 257], 0, (uint8_t)(K_LEN | (kLenExtra[s - 257] << 4))};
    return Entry{0, 0, K_BAD};
}
inline Entry make_dist(int s) {
    if (s < 30) return Entry{kDistBase[s], 0, (uint8_t)(K_LEN | (kDistExtra[s] << 4))};
    return Entry{0, 0, K_BAD};
}

// builds both tables from the two length arrays; false when a code is invalid
inline bool build_tables(Tables& T, const uint8_t* ll, int nl, const uint8_t* dl, int nd) {
    uint16_t cl[16], cd[16];
    if (ll[256] == 0) return false;                                   // zlib: "missing end-of-block"
    const int ml = check_lengths(ll, nl, true, cl);
    if (ml <= 0) return false;
    const int md = check_lengths(dl, nd, true, cd);
    if (md < 0) return false;
    build_table(ll, nl, cl, kLitBits, T.lit, T.subneed, make_lit);
    T.have_dist = md > 0;
    if (md > 0) build_table(dl, nd, cd, kDistBits, T.dist, T.subneed, make_dist);
    return true;
}

inline const Tables& fixed_tables() {
    static const std::unique_ptr<Tables> F = [] {
        std::unique_ptr<Tables> t(new Tables());
        uint8_t ll[288], dl[32];
        for (int i = 0; i < 144; ++i) ll[i] = 8;
        for (int i = 144; i < 256; ++i) ll[i] = 9;
        for (int i = 256; i < 280; ++i) ll[i] = 7;
        for (int i = 280; i < 288; ++i) ll[i] = 8;
        for (int i = 0; i < 32; ++i) dl[i] = 5;
        uint16_t cl[16], cd[16];
        check_lengths(ll, 288, false, cl);
        check_lengths(dl, 32, false, cd);
        build_table(ll, 288, cl, kLitBits, t->lit, t->subneed, make_lit);
        build_table(dl, 32, cd, kDistBits, t->dist, t->subneed, make_dist);
        t->have_dist = true;
        return t;
    }();
    return *F;
}

// Reads a dynamic block's code lengths (RFC 1951 3.2.7) after the 3 header bits; false when anything is invalid.
inline bool read_dynamic_header(BitReader& br, Tables& T) {
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    br.refill();
    const uint32_t hlit = br.take(5) + 257, hdist = br.take(5) + 1, hclen = br.take(4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    uint8_t pl[19] = {0};
    br.refill();
    for (uint32_t i = 0; i < hclen; ++i) {
        if (i == 16) br.refill();
        pl[order[i]] = (uint8_t)br.take(3);
    }
    // the code-length code must be complete (zlib: type CODES)
    int left = 128;
    uint16_t cnt[8] = {0};
    for (int i = 0; i < 19; ++i) cnt[pl[i]]++;
    for (int l = 1; l <= 7; ++l) left -= (int)cnt[l] << (7 - l);
    if (left != 0) return false;
    // 7-bit direct table
    uint8_t psym[128], plen[128];
    {
        uint32_t code = 0, next[8];
        next[0] = 0;
        for (int l = 1; l <= 7; ++l) { code = (code + (l > 1 ? cnt[l - 1] : 0)) << 1; next[l] = code; }
        for (int s = 0; s < 19; ++s) {
            const int l = pl[s];
            if (!l) continue;
            const uint32_t r = rev_bits(next[l]++, l);
            for (uint32_t i = r; i < 128; i += 1u << l) { psym[i] = (uint8_t)s; plen[i] = (uint8_t)l; }
        }
    }
    uint8_t lens[286 + 30 + 138];
    const uint32_t total = hlit + hdist;
    uint32_t i = 0;
    while (i < total) {
        br.refill();
        if (br.overrun()) return false;
        const uint32_t x = br.peek(7);
        const uint32_t s = psym[x];
        br.drop(plen[x]);
        if (s < 16) { lens[i++] = (uint8_t)s; continue; }
        uint32_t rep, v = 0;
        if (s == 16) {
            if (i == 0) return false;
            v = lens[i - 1];
            rep = 3 + br.take(2);
        } else if (s == 17) rep = 3 + br.take(3);
        else rep = 11 + br.take(7);
        if (i + rep > total) return false;
        while (rep--) lens[i++] = (uint8_t)v;
    }
    return build_tables(T, lens, (int)hlit, lens + hlit, (int)hdist);
}

// ---- block finder -----------------------------------------------------------------------------------------------
// First bit offset in [from, to) at which a non-final dynamic block header parses; UINT64_MAX when none.
inline uint64_t find_block(const uint8_t* z, size_t n, uint64_t from, uint64_t to, Tables& scratch,
                           const std::atomic<bool>* cancel = nullptr) {
    const uint64_t last = (uint64_t)n * 8;
    for (uint64_t o = from; o < to; ++o) {
        if (cancel && (o & 0xFFFFu) == 0 && cancel->load(std::memory_order_relaxed)) return UINT64_MAX;
        if (o + 64 > last) {
            // near the end of the input: go through the bit reader for every offset
            BitReader br;
            br.init(z, n, o);
            if (br.peek(3) != 4u) continue;
            br.drop(3);
            if (read_dynamic_header(br, scratch) && !br.overrun()) return o;
            continue;
        }
        uint64_t w;
        memcpy(&w, z + (o >> 3), 8);
        w >>= (o & 7);
        // BFINAL = 0, BTYPE = 10b (bits: 0, 0, 1), HLIT <= 29, HDIST <= 29
        if ((w & 7u) != 4u) continue;
        if (((w >> 3) & 31u) > 29u || ((w >> 8) & 31u) > 29u) continue;
        BitReader br;
        br.init(z, n, o + 3);
        if (read_dynamic_header(br, scratch) && !br.overrun()) return o;
    }
    return UINT64_MAX;
}

// ---- inflate of a run of blocks -------------------------------------------------------------------------------
struct MemberEnd { uint64_t out_off; uint32_t crc, isize; };   // out_off: chunk output bytes before the member's end

struct ChunkOut {
    // output = resolve(m16[kWin..]) ++ b8[kWin..]   (each buffer starts with a kWin prefix: the window)
    std::vector<uint16_t> m16;
    std::vector<uint8_t> b8;
    size_t n16 = 0, n8 = 0;          // symbols / bytes after the prefixes
    std::vector<MemberEnd> ends;
    uint64_t start_bit = UINT64_MAX, end_bit = 0;
    bool ok = false, at_eof = false;
    std::string err;
    size_t out_bytes() const { return n16 + n8; }
};

// gzip member header (RFC 1952); leaves the reader at the first deflate bit.  0 ok, 1 clean end (no further member), -1 bad
inline int read_gzip_header(BitReader& br) {
    br.align_byte();
    br.refill();
    if (br.bitpos() >= (uint64_t)(br.end - br.base) * 8) return 1;
    auto byte = [&]() -> int { br.refill(); if (br.overrun()) return -1; const int v = (int)br.take(8); return br.overrun() ? -1 : v; };
    const int m0 = byte(), m1 = byte();
    if (m0 != 0x1f || m1 != 0x8b) return 1;        // trailing garbage is ignored, as gzip does
    if (byte() != 8) return -1;
    const int flg = byte();
    if (flg < 0 || (flg & 0xE0)) return -1;
    for (int i = 0; i < 6; ++i) if (byte() < 0) return -1;
    if (flg & 4) {
        const int a = byte(), b = byte();
        if (a < 0 || b < 0) return -1;
        for (int i = 0, k = a | (b << 8); i < k; ++i) if (byte() < 0) return -1;
    }
    if (flg & 8) for (;;) { const int c = byte(); if (c < 0) return -1; if (c == 0) break; }
    if (flg & 16) for (;;) { const int c = byte(); if (c < 0) return -1; if (c == 0) break; }
    if (flg & 2) { if (byte() < 0 || byte() < 0) return -1; }
    return 0;
}

enum BlockResult { BR_EOB = 0, BR_ERR = 1 };

// One compressed block's symbols into out[pos...]; T = uint16_t (marker mode) or uint8_t.  `grow` makes room.
// The reader's state and the output cursor live in locals for the duration of the block (byte stores may alias anything).
template <class T, class Grow>
inline BlockResult inflate_codes(BitReader& br, const Tables& tb, T*& out, size_t& pos, size_t& cap, Grow grow) {
    const Entry* const lit = tb.lit;
    const Entry* const dist = tb.dist;
    const bool have_dist = tb.have_dist;
    constexpr uint64_t LM = (1u << kLitBits) - 1, DM = (1u << kDistBits) - 1;
    constexpr uint32_t W = 16 / sizeof(T);   // elements per 16-byte move (the buffer keeps 320 slack elements)
    uint64_t buf = br.buf;
    uint32_t cnt = br.cnt, over = br.over;
    const uint8_t* ip = br.ip;
    const uint8_t* const end = br.end;
    T* obase = out;
    T* o = out + pos;
    T* olimit = out + cap - 320;
    BlockResult res = BR_ERR;
    for (;;) {
        if (o > olimit) {
            pos = (size_t)(o - obase);
            grow();
            obase = out; o = out + pos; olimit = out + cap - 320;
        }
        if (ip + 8 <= end) {
            uint64_t w;
            memcpy(&w, ip, 8);
            buf |= w << cnt;
            ip += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56) {
                if (ip < end) buf |= (uint64_t)*ip++ << cnt; else over += 8;
                cnt += 8;
            }
        }
        Entry e = lit[buf & LM];
        if (e.kx == K_LIT) {
            // up to three literals on the bits already loaded (<= 45 of >= 56)
            buf >>= e.len; cnt -= e.len;
            *o++ = (T)e.val;
            e = lit[buf & LM];
            if (e.kx == K_LIT) {
                buf >>= e.len; cnt -= e.len;
                *o++ = (T)e.val;
                e = lit[buf & LM];
                if (e.kx == K_LIT) {
                    buf >>= e.len; cnt -= e.len;
                    *o++ = (T)e.val;
                }
            }
            continue;
        }
        if (e.kx == K_SUB) {
            e = lit[e.val + ((buf >> kLitBits) & ((1u << e.len) - 1))];
            if (e.kx == K_LIT) { buf >>= e.len; cnt -= e.len; *o++ = (T)e.val; continue; }
        }
        if ((e.kx & 15) == K_LEN) {
            buf >>= e.len; cnt -= e.len;
            const uint32_t xl = e.kx >> 4;
            const uint32_t len = e.val + (uint32_t)(buf & ((1u << xl) - 1));
            buf >>= xl; cnt -= xl;
            if (!have_dist) break;
            Entry d = dist[buf & DM];
            if (d.kx == K_SUB) d = dist[d.val + ((buf >> kDistBits) & ((1u << d.len) - 1))];
            if ((d.kx & 15) != K_LEN) break;
            buf >>= d.len; cnt -= d.len;
            const uint32_t xd = d.kx >> 4;
            const uint32_t dd = d.val + (uint32_t)(buf & ((1u << xd) - 1));
            buf >>= xd; cnt -= xd;
            if (over > cnt) break;
            if (dd > kWin || dd > (size_t)(o - obase)) break;
            const T* src = o - dd;
            T* dst = o;
            o += len;
            if (dd >= W) {
                memcpy(dst, src, 16);
                if (len > W) {
                    memcpy(dst + W, src + W, 16);
                    for (uint32_t k = 2 * W; k < len; k += W) memcpy(dst + k, src + k, 16);
                }
            } else if (dd == 1) {
                const T v = *src;
                for (uint32_t k = 0; k < len; ++k) dst[k] = v;
            } else {
                for (uint32_t k = 0; k < len; ++k) dst[k] = src[k];
            }
            continue;
        }
        if (e.kx == K_EOB) { buf >>= e.len; cnt -= e.len; res = over > cnt ? BR_ERR : BR_EOB; }
        break;
    }
    br.buf = buf; br.cnt = cnt; br.ip = ip; br.over = over;
    pos = (size_t)(o - obase);
    return res;
}

// Decodes blocks from the reader's position until (a) a block boundary at or after `stop_bit` where a non-final
// dynamic block begins, or (b) the end of the data.  `window`: the 32 KiB before the first block when known (byte
// mode from the start), nullptr for marker mode.  `member_start`: the reader stands on a gzip header.
inline void inflate_chunk(const uint8_t* z, size_t n, uint64_t from_bit, uint64_t stop_bit, const uint8_t* window, bool member_start,
                          ChunkOut& C, Tables& tb) {
    BitReader br;
    br.init(z, n, from_bit);
    C.start_bit = from_bit;
    C.ok = false; C.at_eof = false; C.n16 = C.n8 = 0; C.ends.clear();
    bool bytes = window != nullptr || member_start;
    const uint64_t stop_eff = std::min<uint64_t>(stop_bit, (uint64_t)n * 8);
    const size_t est = (size_t)((stop_eff > from_bit ? (stop_eff - from_bit) / 8 : 0) * 4) + (1u << 16);
    uint16_t* o16 = nullptr; size_t p16 = kWin, c16 = 0;
    uint8_t* o8 = nullptr; size_t p8 = kWin, c8 = 0;
    // (the vectors may come from a pool with a size of their own: it is used as it is, and only ever grown)
    auto grow16 = [&]() {
        c16 = std::max<size_t>(c16 * 2, kWin + est);
        if (C.m16.size() < c16) C.m16.resize(c16); else c16 = C.m16.size();
        o16 = C.m16.data();
    };
    auto grow8 = [&]() {
        c8 = std::max<size_t>(c8 * 2, kWin + est);
        if (C.b8.size() < c8) C.b8.resize(c8); else c8 = C.b8.size();
        o8 = C.b8.data();
    };
    auto to_bytes = [&](bool fresh_member) {
        // byte mode from here on: the last kWin symbols (marker free, or irrelevant at a member start) become the prefix
        grow8();
        for (size_t k = 0; k < kWin; ++k) o8[k] = fresh_member ? 0 : (uint8_t)o16[p16 - kWin + k];
        bytes = true;
    };
    if (bytes) {
        grow8();
        if (window) memcpy(o8, window, kWin); else memset(o8, 0, kWin);
    } else {
        grow16();
        for (uint32_t k = 0; k < kWin; ++k) o16[k] = (uint16_t)(kMarker | k);
    }
    size_t checked = kWin;   // marker scan: symbols before `checked` were looked at
    bool need_header = member_start;
    for (;;) {
        if (need_header) {
            const int h = read_gzip_header(br);
            if (h == 1) { C.at_eof = true; break; }
            if (h < 0) { C.err = "bad gzip member header"; return; }
            need_header = false;
            if (!bytes) to_bytes(true);
        }
        br.refill();
        if (br.overrun()) { C.err = "unexpected end of deflate data"; return; }
        if (br.bitpos() >= stop_bit && br.peek(3) == 4u) break;   // a boundary the next chunk's finder accepts
        const uint32_t bfinal = br.take(1), btype = br.take(2);
        BlockResult r = BR_EOB;
        if (btype == 0) {
            br.align_byte();
            br.refill();
            const uint32_t len = br.take(16), nlen = br.take(16);
            if (br.overrun() || (len ^ 0xFFFFu) != nlen) { C.err = "bad stored block"; return; }
            // the bytes: first what the bit buffer holds, then straight from the input
            uint32_t left = len;
            if (bytes) { while (p8 + left + 320 > c8) grow8(); } else { while (p16 + left + 320 > c16) grow16(); }
            while (left && br.cnt >= 8) {
                const uint8_t v = (uint8_t)br.take(8);
                if (bytes) o8[p8++] = v; else o16[p16++] = v;
                --left;
            }
            if (br.overrun()) { C.err = "stored block past the end"; return; }
            if (left) {
                // the bit buffer is empty: ip is the next input byte
                if ((size_t)(br.end - br.ip) < left) { C.err = "stored block past the end"; return; }
                if (bytes) { memcpy(o8 + p8, br.ip, left); p8 += left; }
                else { for (uint32_t k = 0; k < left; ++k) o16[p16 + k] = br.ip[k]; p16 += left; }
                br.ip += left;
                br.buf = 0; br.cnt = 0;
            }
        } else if (btype == 1) {
            const Tables& F = fixed_tables();
            r = bytes ? inflate_codes<uint8_t>(br, F, o8, p8, c8, grow8) : inflate_codes<uint16_t>(br, F, o16, p16, c16, grow16);
        } else if (btype == 2) {
            if (!read_dynamic_header(br, tb)) { C.err = "bad dynamic block header"; return; }
            r = bytes ? inflate_codes<uint8_t>(br, tb, o8, p8, c8, grow8) : inflate_codes<uint16_t>(br, tb, o16, p16, c16, grow16);
        } else { C.err = "reserved block type"; return; }
        if (r != BR_EOB) { C.err = "bad deflate data"; return; }
        if (!bytes && p16 - checked >= kWin) {
            // marker free for 32 KiB?  (markers only ever enter through copies from the window or from markers)
            bool clean = true;
            size_t first_dirty = p16;
            for (size_t k = p16; k-- > p16 - kWin;) if (o16[k] & kMarker) { clean = false; first_dirty = k; break; }
            if (clean) to_bytes(false);
            else checked = first_dirty + 1;
        }
        if (bfinal) {
            br.align_byte();
            uint32_t f[2] = {0, 0};
            for (int w = 0; w < 2; ++w) { br.refill(); f[w] = br.take(32); }
            if (br.overrun()) { C.err = "truncated gzip trailer"; return; }
            C.ends.push_back(MemberEnd{(uint64_t)((p16 - kWin) + (bytes ? p8 - kWin : 0)), f[0], f[1]});
            need_header = true;   // the next member's header (or the end of the data) decides
        }
    }
    C.n16 = p16 - kWin;
    C.n8 = bytes ? p8 - kWin : 0;
    C.end_bit = br.bitpos();
    C.ok = true;
}

// markers -> window bytes, plain symbols -> bytes.  Most 16-symbol groups hold no marker (markers are what was copied,
// directly or through other copies, from before the chunk): those are packed with two vector instructions.
inline void resolve_markers(const uint16_t* s, size_t n, const uint8_t* w, uint8_t* d) {
    size_t i = 0;
#if defined(__SSE2__)
    for (; i + 16 <= n; i += 16) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 8));
        if (_mm_movemask_epi8(_mm_or_si128(a, b)) & 0xAAAA) {
            for (size_t k = i; k < i + 16; ++k) { const uint16_t v = s[k]; d[k] = (v & kMarker) ? w[v & (kMarker - 1)] : (uint8_t)v; }
        } else {
            _mm_storeu_si128(reinterpret_cast<__m128i*>(d + i), _mm_packus_epi16(a, b));
        }
    }
#endif
    for (; i < n; ++i) { const uint16_t v = s[i]; d[i] = (v & kMarker) ? w[v & (kMarker - 1)] : (uint8_t)v; }
}

// ---- the reader ---------------------------------------------------------------------------------------------------
class Reader {
  public:
    // threads <= 0: all cores.  chunk_bytes: compressed bytes per speculative chunk.
    bool open(const char* path, int threads, size_t chunk_bytes = 2u << 20) {
        close();
        fd_ = ::open(path, O_RDONLY);
        if (fd_ < 0) { err_ = std::string("cannot open ") + path; return false; }
        struct stat sb;
        if (fstat(fd_, &sb) != 0 || !S_ISREG(sb.st_mode)) { err_ = "not a regular file"; ::close(fd_); fd_ = -1; return false; }
        n_ = (size_t)sb.st_size;
        if (n_) {
            void* m = mmap(nullptr, n_, PROT_READ, MAP_PRIVATE, fd_, 0);
            if (m == MAP_FAILED) { err_ = "mmap failed"; ::close(fd_); fd_ = -1; return false; }
            z_ = static_cast<const uint8_t*>(m);
            madvise(const_cast<uint8_t*>(z_), n_, MADV_SEQUENTIAL);
        }
        return start(threads, chunk_bytes);
    }
    // decode from memory the caller keeps alive (tests)
    bool open_memory(const uint8_t* z, size_t n, int threads, size_t chunk_bytes = 2u << 20) {
        close();
        z_ = z; n_ = n; mapped_ = false;
        return start(threads, chunk_bytes);
    }
    ~Reader() { close(); }

    void close() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto& t : pool_) if (t.joinable()) t.join();
        pool_.clear();
        chunks_.clear();
        if (mapped_ && z_ && n_) munmap(const_cast<uint8_t*>(z_), n_);
        if (fd_ >= 0) ::close(fd_);
        fd_ = -1; z_ = nullptr; n_ = 0; mapped_ = true; stop_ = false;
    }

    // gzread semantics: up to `want` bytes; 0 at the end of the stream; -1 on a decode error (see error())
    int64_t read(uint8_t* dst, size_t want) {
        size_t got = 0;
        while (got < want) {
            if (failed_) return got ? (int64_t)got : -1;
            if (!cur_) {
                cur_ = next_ready();
                cur_off_ = 0;
                if (!cur_) { if (failed_ && !got) return -1; break; }
            }
            const size_t hn = cur_->out.n16;             // resolved symbols in `head`
            const size_t total = hn + cur_->out.n8;
            if (cur_off_ >= total) { retire(cur_); cur_ = nullptr; continue; }
            // [head | b8 after its prefix]
            size_t k;
            if (cur_off_ < hn) {
                k = std::min(want - got, hn - cur_off_);
                memcpy(dst + got, cur_->head.data() + cur_off_, k);
            } else {
                const size_t o = cur_off_ - hn;
                k = std::min(want - got, cur_->out.n8 - o);
                memcpy(dst + got, cur_->out.b8.data() + kWin + o, k);
            }
            got += k; cur_off_ += k;
        }
        return (int64_t)got;
    }
    const std::string& error() const { return err_; }
    int threads() const { return (int)pool_.size(); }
    uint64_t redecoded_chunks() const { return redecoded_; }
    uint64_t marker_symbols() const { return marker_symbols_; }

  private:
    struct Chunk {
        size_t idx = 0;
        uint64_t nominal_begin = 0, nominal_end = 0;
        ChunkOut out;
        int state = 0;       // 0 queued, 1 stage 1 running, 2 stage 1 done, 3 accepted (stage 2 queued/running), 4 ready
        uint8_t window[kWin];
        bool skipped = false;
        bool found = false;                 // stage 1 has a block start and is inflating from it
        std::atomic<bool> abandon{false};   // the sequencer got here first and decodes the chunk itself
        bool taken = false;
        std::vector<uint8_t> head;          // resolved m16
        std::vector<uint32_t> seg_crc;      // crc of [segment start, member end) pieces, then the open tail
        std::vector<uint64_t> seg_len;
    };

    bool start(int threads, size_t chunk_bytes) {
        if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
        threads = std::min(threads, 128);
        chunk_ = std::max<size_t>(chunk_bytes, 64u << 10);
        n_chunks_ = n_ ? (n_ + chunk_ - 1) / chunk_ : 0;
        next_s1_ = 0; next_seq_ = 0; next_out_ = 0; expect_bit_ = 0; expect_member_ = true;
        failed_ = false; finished_ = n_ == 0; err_.clear();
        crc_ = crc32(0L, Z_NULL, 0); member_len_ = 0; redecoded_ = 0; marker_symbols_ = 0; final_idx_ = 0; chunks_base_ = 0;
        cur_.reset(); cur_off_ = 0; s2_.clear(); pool_bufs_.clear();
        lookahead_ = std::min<size_t>((size_t)threads * 2 + 2, 64);   // chunks in flight (each holds its output: ~10-40 MB)
        memset(last_window_, 0, sizeof last_window_);
        if (n_ && n_ < 18) { err_ = "not a gzip stream"; failed_ = true; return false; }
        if (n_ && (z_[0] != 0x1f || z_[1] != 0x8b)) { err_ = "not a gzip stream"; failed_ = true; return false; }
        for (int t = 0; t < threads; ++t) pool_.emplace_back([this] { worker(); });
        return true;
    }

    void worker() {
        std::unique_ptr<Tables> tb(new Tables());
        for (;;) {
            std::shared_ptr<Chunk> c;
            int what = 0;
            {
                std::unique_lock<std::mutex> lk(mu_);
                for (;;) {
                    if (stop_) return;
                    if (!s2_.empty()) { c = s2_.front(); s2_.pop_front(); what = 2; break; }
                    if (!failed_ && !finished_ && next_s1_ < n_chunks_ && next_s1_ < next_out_ + lookahead_) {
                        c = std::make_shared<Chunk>();
                        c->idx = next_s1_++;
                        c->nominal_begin = (uint64_t)c->idx * chunk_ * 8;
                        c->nominal_end = std::min<uint64_t>((uint64_t)(c->idx + 1) * chunk_, n_) * 8;
                        c->state = 1;
                        if (!pool_bufs_.empty()) {
                            Bufs& b = pool_bufs_.back();
                            c->out.m16.swap(b.m16); c->out.b8.swap(b.b8); c->head.swap(b.head);
                            pool_bufs_.pop_back();
                        }
                        chunks_.push_back(c);
                        what = 1;
                        break;
                    }
                    cv_work_.wait(lk);
                }
            }
            if (what == 1) stage1(*c, *tb); else stage2(*c);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (what == 2) c->state = 4;
                else if (!c->taken) c->state = 2;
            }
            cv_done_.notify_all();
        }
    }

    // speculative decode of a chunk: chunk 0 starts on the gzip header with a known (empty) window
    void stage1(Chunk& c, Tables& tb) {
        const uint64_t stop = c.idx + 1 == n_chunks_ ? UINT64_MAX : c.nominal_end;
        if (c.idx == 0) { inflate_chunk(z_, n_, 0, stop, nullptr, true, c.out, tb); return; }
        uint64_t from = c.nominal_begin;
        for (int tries = 0; tries < 64; ++tries) {
            const uint64_t o = find_block(z_, n_, from, c.nominal_end, tb, &c.abandon);
            if (o == UINT64_MAX) break;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (c.abandon.load()) return;        // (the sequencer owns c.out from here on)
                c.found = true;
            }
            inflate_chunk(z_, n_, o, stop, nullptr, false, c.out, tb);
            if (c.out.ok) return;
            {
                std::lock_guard<std::mutex> lk(mu_);
                c.found = false;
            }
            from = o + 1;      // a false positive of the finder: the data after it does not decode
        }
        if (c.abandon.load()) return;
        c.out.ok = false;
        c.out.start_bit = UINT64_MAX;
    }

    // markers -> bytes, CRC-32 of the member segments
    void stage2(Chunk& c) {
        ChunkOut& o = c.out;
        if (c.head.size() < o.n16) c.head.resize(o.n16);              // (buffers come from the pool: grown, never shrunk)
        resolve_markers(o.m16.data() + kWin, o.n16, c.window, c.head.data());
        // segments: [0, end0), [end0, end1), ..., [end_last, total)
        c.seg_crc.clear(); c.seg_len.clear();
        uint64_t a = 0;
        const uint64_t total = o.n16 + o.n8;
        auto crc_range = [&](uint64_t lo, uint64_t hi) {
            uLong x = crc32(0L, Z_NULL, 0);
            while (lo < hi) {
                const uint8_t* p; uint64_t k;
                if (lo < o.n16) { p = c.head.data() + lo; k = std::min<uint64_t>(hi, o.n16) - lo; }
                else { p = o.b8.data() + kWin + (lo - o.n16); k = hi - lo; }
                k = std::min<uint64_t>(k, 1u << 30);
                x = crc32(x, p, (uInt)k);
                lo += k;
            }
            return (uint32_t)x;
        };
        for (const MemberEnd& e : o.ends) { c.seg_crc.push_back(crc_range(a, e.out_off)); c.seg_len.push_back(e.out_off - a); a = e.out_off; }
        c.seg_crc.push_back(crc_range(a, total)); c.seg_len.push_back(total - a);
    }

    // the last kWin bytes of (window ++ chunk output) without resolving the whole chunk
    void next_window(const Chunk& c, uint8_t* nw) const {
        const ChunkOut& o = c.out;
        const size_t total = o.n16 + o.n8;
        for (size_t k = 0; k < kWin; ++k) {
            // byte at distance kWin - k from the end
            const size_t back = kWin - k;
            if (back <= o.n8) nw[k] = o.b8[kWin + o.n8 - back];
            else if (back <= total) {
                const uint16_t v = o.m16[kWin + o.n16 - (back - o.n8)];
                nw[k] = (v & kMarker) ? c.window[v & (kMarker - 1)] : (uint8_t)v;
            } else nw[k] = c.window[kWin - (back - total)];
        }
    }

    // Accepts chunks in order (verification, window hand-over, stage 2 dispatch) as far as stage 1 has got, then returns
    // the next chunk whose stage 2 is complete; blocks while neither can make progress.  nullptr at the end / on failure.
    std::shared_ptr<Chunk> next_ready() {
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            if (failed_ || n_chunks_ == 0) return nullptr;
            // sequence
            bool progressed = false;
            while (!finished_ && next_seq_ < chunks_base_ + chunks_.size()) {
                std::shared_ptr<Chunk> c = chunks_[next_seq_ - chunks_base_];
                if (c->state >= 3) { ++next_seq_; continue; }
                if (c->nominal_end <= expect_bit_) {
                    // the predecessor ran through this whole chunk: nothing of it is needed (a worker still busy with it
                    // keeps its own reference and finds `taken`)
                    progressed = true;
                    c->abandon.store(true);
                    c->taken = true;
                    c->seg_crc.clear(); c->seg_len.clear();
                    c->state = 4; c->skipped = true;
                    if (c->idx + 1 == n_chunks_) { finished_ = true; final_idx_ = c->idx; }
                    ++next_seq_;
                    continue;
                }
                if (c->state < 2) {
                    // still searching for a block start (a stream of stored / fixed blocks has none): rather than wait for
                    // the search to run through the whole chunk, take the chunk over -- the start is known here
                    if (c->state == 1 && !c->found && c->idx != 0 && c->nominal_end > expect_bit_ && waited_) {
                        c->abandon.store(true);
                        c->taken = true;
                        c->out = ChunkOut();
                        c->state = 2;
                    } else break;
                }
                progressed = true;
                if (!c->out.ok || c->out.start_bit != expect_bit_) {
                    // not where the predecessor stopped (or nothing found): decode it here, window known
                    lk.unlock();
                    std::unique_ptr<Tables> tb(new Tables());
                    const uint64_t stop = c->idx + 1 == n_chunks_ ? UINT64_MAX : c->nominal_end;
                    ChunkOut fresh;
                    inflate_chunk(z_, n_, expect_bit_, stop, last_window_, expect_member_, fresh, *tb);
                    lk.lock();
                    if (c->idx != 0) ++redecoded_;
                    if (!fresh.ok) { fail(fresh.err.empty() ? "bad deflate data" : fresh.err); return nullptr; }
                    c->out = std::move(fresh);
                }
                memcpy(c->window, last_window_, kWin);
                marker_symbols_ += c->out.n16;
                {
                    uint8_t nw[kWin];
                    next_window(*c, nw);
                    memcpy(last_window_, nw, kWin);
                }
                expect_bit_ = c->out.end_bit;
                expect_member_ = false;
                if (c->out.at_eof || c->idx + 1 == n_chunks_) { finished_ = true; final_idx_ = c->idx; }
                c->state = 3;
                s2_.push_back(c);
                ++next_seq_;
                cv_work_.notify_one();
            }
            // output
            if (finished_ && next_out_ > final_idx_) return nullptr;
            if (next_out_ < chunks_base_ + chunks_.size()) {
                std::shared_ptr<Chunk> c = chunks_[next_out_ - chunks_base_];
                if (c->state == 4) {
                    if (!fold_crc(*c)) return nullptr;
                    ++next_out_;
                    chunks_.pop_front(); ++chunks_base_;
                    cv_work_.notify_all();
                    if (c->skipped) continue;      // (its buffers die with it: a worker may still be writing into them)
                    return c;
                }
            }
            if (progressed) { waited_ = false; continue; }
            cv_work_.notify_all();
            // (a short timed wait first: a worker that is merely finishing its chunk gets the chance to)
            waited_ = cv_done_.wait_for(lk, std::chrono::milliseconds(waited_ ? 50 : 3)) == std::cv_status::timeout;
        }
    }

    // running CRC-32 / length of the open member over the chunk's segments; every member end is checked
    bool fold_crc(const Chunk& c) {
        const ChunkOut& o = c.out;
        for (size_t i = 0; i < c.seg_crc.size(); ++i) {
            if (c.seg_len[i]) {
                crc_ = crc32_combine(crc_, c.seg_crc[i], (z_off_t)c.seg_len[i]);
                member_len_ += c.seg_len[i];
            }
            if (i < o.ends.size()) {
                if ((uint32_t)crc_ != o.ends[i].crc || (uint32_t)member_len_ != o.ends[i].isize) { fail("gzip CRC-32 / length check failed"); return false; }
                crc_ = crc32(0L, Z_NULL, 0); member_len_ = 0;
            }
        }
        return true;
    }

    // a consumed (or skipped) chunk's buffers go back to the pool: fresh allocations of this size are page-faulted and
    // zero-filled by the kernel, which costs a fifth of the decode time
    struct Bufs { std::vector<uint16_t> m16; std::vector<uint8_t> b8, head; };
    void retire(std::shared_ptr<Chunk>& c) {
        Bufs b;
        b.m16.swap(c->out.m16); b.b8.swap(c->out.b8); b.head.swap(c->head);
        std::lock_guard<std::mutex> lk(mu_);
        if (pool_bufs_.size() < lookahead_ + 2) pool_bufs_.push_back(std::move(b));
    }
    void fail(const std::string& e) { failed_ = true; if (err_.empty()) err_ = e; cv_work_.notify_all(); }

    int fd_ = -1;
    const uint8_t* z_ = nullptr;
    size_t n_ = 0;
    bool mapped_ = true;
    size_t chunk_ = 0, n_chunks_ = 0, lookahead_ = 0;
    std::vector<std::thread> pool_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    bool stop_ = false, failed_ = false, finished_ = false;
    std::deque<std::shared_ptr<Chunk>> chunks_;   // chunks_[i] is chunk chunks_base_ + i
    size_t chunks_base_ = 0;
    std::deque<std::shared_ptr<Chunk>> s2_;
    std::vector<Bufs> pool_bufs_;
    size_t next_s1_ = 0, next_seq_ = 0, next_out_ = 0;
    uint64_t expect_bit_ = 0;
    bool expect_member_ = true;
    uint8_t last_window_[kWin];
    uLong crc_ = 0;
    uint64_t member_len_ = 0;
    size_t final_idx_ = 0;
    bool waited_ = false;                  // the sequencer's last wait timed out with nothing done
    std::shared_ptr<Chunk> cur_;
    size_t cur_off_ = 0;
    std::string err_;
    uint64_t redecoded_ = 0, marker_symbols_ = 0;
};

}  // namespace bsq_pgz
