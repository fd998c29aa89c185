// bsq_inflate.cuh -- DEFLATE (RFC 1951) on the device for BGZF input (SAM specification 4.1).
//
// The reference reads gzip input through RapidgzipReader(parallelism) (blazeseq/io/readers.mojo:380-443), a
// parallel decoder on host threads.  A BGZF file is a series of gzip members of <= 64 KiB of payload each,
// every one an independent DEFLATE stream with its compressed size in the header and its inflated size in
// the trailer -- tens of thousands of independent streams per GiB.  Here the COMPRESSED members cross PCIe
// and ONE WARP inflates ONE member straight into the parse window in HBM:
//
//   k_inflate_members_uniform (shipped): every lane of the warp runs the decode loop on identical state -- one
//   broadcast load per input word, one broadcast shared-memory lookup per Huffman code (tables rebuilt per block
//   by lane 0) -- so that a match needs no hand-off: every lane already knows (position, length, distance) and
//   lane i copies byte i (LZ77 history is the member's own output, read back from L1 / L2); a literal is one
//   store; stored blocks are warp copies.
//   k_inflate_members (tuning builds, BSQ_INF_LANES lanes per member): the group's first lane decodes and
//   broadcasts every match to its group.
//
// The decode chain of one member is serial by nature; the parallelism is across members (a 256 MiB region is
// ~4,000 warps of work).  Output is bit-exact zlib (tests compare with zlib.decompress on the reference's own
// .bgz fixtures and synthetic files); a malformed or truncated member sets a per-member status and the
// stream reports BSQ_E_IO.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bsq {

struct InflateMember {
    uint64_t src;        // byte offset of the member's DEFLATE payload in the compressed buffer
    uint32_t src_len;    // payload bytes
    uint32_t isize;      // inflated size (gzip ISIZE)
    uint64_t dst;        // byte offset of the member's output in the destination buffer
    uint32_t crc;        // gzip CRC-32 of the inflated bytes (checked by k_crc32_members)
    uint32_t _pad;
};

#ifndef BSQ_INF_LANES
#define BSQ_INF_LANES 32
#endif
#ifndef BSQ_INF_WARPS
#define BSQ_INF_WARPS 8
#endif
constexpr int kInfWarps = BSQ_INF_WARPS;   // warps per CTA
constexpr int kInfLanes = BSQ_INF_LANES;   // lanes per member: one decodes, all copy
constexpr int kInfPerWarp = 32 / kInfLanes;
constexpr int kInfPerCta = kInfWarps * kInfPerWarp;   // members per CTA
constexpr int kLitBits = 9, kDistBits = 7;
// Decode table entry: [value:16][extra bits:8][kind:4][code length:4]
constexpr uint32_t kKindLit = 0u, kKindLen = 1u, kKindEob = 2u, kKindSlow = 3u, kKindBad = 15u;
constexpr uint32_t kEntryBad = kKindBad << 4;          // no such code

__device__ __forceinline__ uint32_t inf_entry(uint32_t value, uint32_t extra, uint32_t kind, uint32_t len) {
    return (value << 16) | (extra << 8) | (kind << 4) | len;
}

__device__ const uint16_t kInfLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115,
                                             131, 163, 195, 227, 258};
__device__ const uint8_t kInfLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__device__ const uint16_t kInfDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537,
                                              2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__device__ const uint8_t kInfDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__device__ const uint8_t kInfClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct InflateTables {
    uint32_t lit[1 << kLitBits];
    uint32_t dist[1 << kDistBits];
    // canonical form for codes longer than the lookup width (puff-style): symbols ordered by (length, symbol)
    uint16_t lit_sym[288], dist_sym[32];
    uint16_t lit_count[16], dist_count[16];
    uint8_t lens[320];                     // code lengths of the block being set up
    uint8_t dl[32];                        // ... and the distance code lengths while the literal table is built
};

struct BitReader {
    const uint32_t* wp;                    // next word to fetch
    const uint32_t* wend;                  // one past the last word that holds payload bytes
    const uint32_t* w0;                    // first word
    uint32_t overrun;                      // the reader ran words past the payload (a damaged stream)
    uint64_t buf;
    uint32_t cnt;                          // valid bits in buf
    uint32_t nextw;                        // prefetched word
    uint32_t lead;                         // bits of the first word that precede the payload (8 x misalignment)
    uint32_t base_bytes;                   // payload bytes before w0's payload start (after a stored block)
};

__device__ __forceinline__ void br_init(BitReader& b, const uint8_t* p, uint32_t nbytes, uint32_t base_bytes) {
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
    b.w0 = b.wp = reinterpret_cast<const uint32_t*>(p - a);
    b.wend = b.w0 + ((a + nbytes + 3u) >> 2);
    const uint32_t first = b.wp < b.wend ? __ldg(b.wp) : 0u;
    ++b.wp;
    b.buf = (uint64_t)(first >> (8u * a));
    b.cnt = 32u - 8u * a;
    b.lead = 8u * a;
    b.nextw = b.wp < b.wend ? __ldg(b.wp) : 0u;
    ++b.wp;
    b.base_bytes = base_bytes;
    b.overrun = 0;
}
// payload bits consumed so far (the prefetched word is not in buf yet)
__device__ __forceinline__ uint64_t br_consumed(const BitReader& b) {
    return (uint64_t)b.base_bytes * 8u + (uint64_t)(b.wp - b.w0 - 1) * 32u - b.lead - b.cnt;
}
// at least 33 valid bits afterwards
__device__ __forceinline__ void br_fill(BitReader& b) {
    if (b.cnt <= 32u) {
        b.buf |= (uint64_t)b.nextw << b.cnt;
        b.cnt += 32u;
        b.nextw = b.wp < b.wend ? __ldg(b.wp) : 0u;   // (past the end: zeros)
        if (b.wp > b.wend + 4) b.overrun = 1u;        // a damaged stream must end the member, not spin on zeros
        ++b.wp;
    }
}
__device__ __forceinline__ uint32_t br_peek(const BitReader& b, uint32_t n) { return (uint32_t)b.buf & ((1u << n) - 1u); }
__device__ __forceinline__ void br_skip(BitReader& b, uint32_t n) { b.buf >>= n; b.cnt -= n; }
__device__ __forceinline__ uint32_t br_take(BitReader& b, uint32_t n) {
    const uint32_t v = br_peek(b, n);
    br_skip(b, n);
    return v;
}

__device__ __forceinline__ uint32_t inf_rev(uint32_t code, uint32_t len) { return __brev(code) >> (32u - len); }

// Builds one decoding table from code lengths (one lane).  Returns false for an over-subscribed set; an incomplete
// set leaves unused entries (= invalid codes).
__device__ bool inf_build(const uint8_t* lens, uint32_t n, uint32_t* tab, uint32_t tbits, uint16_t* sym, uint16_t* count,
                          bool is_dist) {
    uint32_t offs[16];
    for (int i = 0; i < 16; ++i) count[i] = 0;
    for (uint32_t s = 0; s < n; ++s) count[lens[s]]++;
    count[0] = 0;
    int32_t left = 1;
    for (int l = 1; l < 16; ++l) {
        left = (left << 1) - (int32_t)count[l];
        if (left < 0) return false;                       // over-subscribed
    }
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
    for (uint32_t s = 0; s < n; ++s)
        if (lens[s]) sym[offs[lens[s]]++] = (uint16_t)s;
    for (uint32_t i = 0; i < (1u << tbits); ++i) tab[i] = kEntryBad;
    // canonical codes in (length, symbol) order
    uint32_t code = 0, idx = 0;
    for (uint32_t l = 1; l < 16; ++l) {
        for (uint32_t k = 0; k < count[l]; ++k, ++idx, ++code) {
            const uint32_t s = sym[idx];
            uint32_t e;
            if (is_dist) e = s < 30u ? inf_entry(kInfDistBase[s], kInfDistExtra[s], kKindLen, l) : kEntryBad;
            else if (s < 256u) e = inf_entry(s, 0u, kKindLit, l);
            else if (s == 256u) e = inf_entry(0u, 0u, kKindEob, l);
            else if (s < 286u) e = inf_entry(kInfLenBase[s - 257u], kInfLenExtra[s - 257u], kKindLen, l);
            else e = kEntryBad;
            if (l <= tbits) {
                for (uint32_t r = inf_rev(code, l); r < (1u << tbits); r += 1u << l) tab[r] = e;
            } else {
                // a long code: the table slot that holds its first tbits bits sends the decoder down the canonical path
                tab[inf_rev(code, l) & ((1u << tbits) - 1u)] = inf_entry(0u, 0u, kKindSlow, 0u);
            }
        }
        code <<= 1;
    }
    return true;
}

// canonical decode, one bit at a time (codes longer than the lookup width); returns the symbol or -1
__device__ int32_t inf_slow(BitReader& b, const uint16_t* sym, const uint16_t* count) {
    int32_t code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; ++l) {
        code |= (int32_t)br_take(b, 1);
        const int32_t c = count[l];
        if (code - c < first) return sym[index + (code - first)];
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}

// kInfLanes lanes per member (kInfPerWarp members per warp): the group's first lane decodes, every lane of the group
// copies.  Only one lane in kInfLanes does the serial Huffman work, so a warp instruction of the decode loop advances
// kInfPerWarp members at once -- the kernel is bound by instruction issue, not by memory.
// status[m]: 0 ok, 1 bad block type / table / code, 2 output overrun or bad distance, 3 input overrun, 4 inflated
// size differs from ISIZE (5: CRC mismatch, set by k_crc32_members).
__global__ void __launch_bounds__(kInfWarps * 32) k_inflate_members(const uint8_t* __restrict__ zbuf, uint8_t* __restrict__ out,
                                                                  const InflateMember* __restrict__ members, uint32_t n_members,
                                                                  uint32_t* __restrict__ status) {
    extern __shared__ __align__(16) uint8_t inf_smem[];
    constexpr uint32_t kFull = 0xFFFFFFFFu;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t g = lane / kInfLanes, gl = lane % kInfLanes, lead = lane - gl;
    const uint32_t m = (blockIdx.x * kInfWarps + warp) * kInfPerWarp + g;
    const bool valid = m < n_members;
    InflateTables& T = reinterpret_cast<InflateTables*>(inf_smem)[warp * kInfPerWarp + g];
    InflateMember M{};
    if (valid) M = members[m];
    uint8_t* const dst = out + M.dst;
    const uint8_t* const src0 = zbuf + M.src;
    const uint32_t cap = M.isize;
    const bool leader = gl == 0u && valid;
    BitReader br{};
    uint32_t pos = 0, err = 0;            // (leader)
    bool last = false, in_block = false, finished = false;   // (leader)
    int tables = 0;                        // 0 none, 1 fixed, 2 dynamic (leader)
    if (leader) br_init(br, src0, M.src_len, 0u);
    bool done = !valid;                    // (whole group)
    if (__all_sync(0xFFFFFFFFu, done)) return;   // a warp without members
    uint32_t pend_pos = 0xFFFFFFFFu;       // a match byte loaded but not yet stored (every lane)
    uint8_t pend_val = 0;

    while (true) {
        // ---- block headers: the leaders that stand between two blocks ----
        uint32_t st_len = 0, st_src = 0, st_go = 0;
        const bool at_header = leader && !done && !in_block && !finished && err == 0u;
        if (__any_sync(kFull, at_header)) {                    // (rare: a few blocks per member)
        if (at_header) {
            br_fill(br);
            last = br_take(br, 1) != 0u;
            const uint32_t btype = br_take(br, 2);
            if (btype == 0u) {
                br_skip(br, br.cnt & 7u);                             // to the byte boundary
                br_fill(br);
                const uint32_t len = br_take(br, 16);
                br_fill(br);
                const uint32_t nlen = br_take(br, 16);
                st_src = (uint32_t)(br_consumed(br) >> 3);            // payload offset of the raw bytes
                if ((len ^ nlen) != 0xFFFFu) err = 1u;
                else if (pos + len > cap) err = 2u;
                else if (st_src + len > M.src_len) err = 3u;
                else { st_len = len; st_go = 1u; }
            } else if (btype == 1u) {
                if (tables != 1) {
                    for (int s = 0; s < 144; ++s) T.lens[s] = 8;
                    for (int s = 144; s < 256; ++s) T.lens[s] = 9;
                    for (int s = 256; s < 280; ++s) T.lens[s] = 7;
                    for (int s = 280; s < 288; ++s) T.lens[s] = 8;
                    inf_build(T.lens, 288, T.lit, kLitBits, T.lit_sym, T.lit_count, false);
                    for (int s = 0; s < 30; ++s) T.lens[s] = 5;
                    inf_build(T.lens, 30, T.dist, kDistBits, T.dist_sym, T.dist_count, true);
                    tables = 1;
                }
                in_block = true;
            } else if (btype == 2u) {
                const uint32_t hlit = br_take(br, 5) + 257u, hdist = br_take(br, 5) + 1u, hclen = br_take(br, 4) + 4u;
                if (hlit > 286u || hdist > 30u) err = 1u;
                uint8_t* cl = T.lens + 300;                             // 19 code-length code lengths
                for (int i = 0; i < 19; ++i) cl[i] = 0;
                for (uint32_t i = 0; i < hclen && err == 0u; ++i) { br_fill(br); cl[kInfClOrder[i]] = (uint8_t)br_take(br, 3); }
                // the code-length code borrows the distance table's storage (7-bit lookup)
                if (err == 0u && !inf_build(cl, 19, T.dist, 7, T.dist_sym, T.dist_count, false)) err = 1u;
                uint32_t i = 0;
                while (i < hlit + hdist && err == 0u) {
                    br_fill(br);
                    const uint32_t e = T.dist[br_peek(br, 7)];
                    if ((e & 0xF0u) != 0u) { err = 1u; break; }        // not a (valid) symbol of the code-length code
                    br_skip(br, e & 15u);
                    const uint32_t s = e >> 16;
                    if (s < 16u) { T.lens[i++] = (uint8_t)s; continue; }
                    uint32_t rep, val = 0;
                    if (s == 16u) { if (i == 0u) { err = 1u; break; } val = T.lens[i - 1]; rep = 3u + br_take(br, 2); }
                    else if (s == 17u) rep = 3u + br_take(br, 3);
                    else rep = 11u + br_take(br, 7);
                    if (i + rep > hlit + hdist) { err = 1u; break; }
                    while (rep--) T.lens[i++] = (uint8_t)val;
                }
                if (err == 0u && T.lens[256] == 0) err = 1u;            // no end-of-block code
                if (err == 0u) {
                    for (uint32_t k = 0; k < 32u; ++k) T.dl[k] = k < hdist ? T.lens[hlit + k] : 0;
                    if (!inf_build(T.lens, hlit, T.lit, kLitBits, T.lit_sym, T.lit_count, false)) err = 1u;
                    if (err == 0u && !inf_build(T.dl, hdist, T.dist, kDistBits, T.dist_sym, T.dist_count, true)) err = 1u;
                    tables = 2;
                    in_block = true;
                }
            } else {
                err = 1u;
            }
        }
        // ---- stored blocks: a copy by the group ----
        st_go = __shfl_sync(kFull, st_go, lead);
        st_len = __shfl_sync(kFull, st_len, lead);
        st_src = __shfl_sync(kFull, st_src, lead);
        const uint32_t p0 = __shfl_sync(kFull, pos, lead);
        if (st_go != 0u)
            for (uint32_t i = gl; i < st_len; i += kInfLanes) dst[p0 + i] = src0[st_src + i];
        __syncwarp();
        if (leader && st_go != 0u) {
            pos += st_len;
            br_init(br, src0 + st_src + st_len, M.src_len - st_src - st_len, st_src + st_len);
            if (last) finished = true;
        }
        }
        // ---- compressed blocks: the leader decodes up to its next match ----
        uint32_t mlen = 0, mdist = 0;
        if (leader && !done && in_block && err == 0u) {
            while (true) {
                br_fill(br);
                uint32_t e = T.lit[br_peek(br, kLitBits)];
                if ((e & 0xF0u) == 0u) {                                  // a literal (the common case)
                    br_skip(br, e & 15u);
                    if (pos >= cap) { err = 2u; break; }
                    dst[pos++] = (uint8_t)(e >> 16);
                    continue;
                }
                uint32_t kind = (e >> 4) & 15u;
                uint32_t base, extra;
                if (kind == kKindSlow) {                                  // a code longer than the table width
                    const int32_t s = inf_slow(br, T.lit_sym, T.lit_count);
                    if (s < 0 || s >= 286) { err = 1u; break; }
                    if (s < 256) {
                        if (pos >= cap) { err = 2u; break; }
                        dst[pos++] = (uint8_t)s;
                        continue;
                    }
                    if (s == 256) { in_block = false; if (last) finished = true; break; }
                    base = kInfLenBase[s - 257]; extra = kInfLenExtra[s - 257];
                } else if (kind == kKindLen) {
                    br_skip(br, e & 15u);
                    base = e >> 16; extra = (e >> 8) & 255u;
                } else if (kind == kKindEob) {
                    br_skip(br, e & 15u);
                    in_block = false;
                    if (last) finished = true;
                    break;
                } else { err = 1u; break; }
                mlen = base + br_take(br, extra);
                br_fill(br);
                e = T.dist[br_peek(br, kDistBits)];
                kind = (e >> 4) & 15u;
                if (kind == kKindLen) {
                    br_skip(br, e & 15u);
                    base = e >> 16; extra = (e >> 8) & 255u;
                } else if (kind == kKindSlow) {
                    const int32_t d = inf_slow(br, T.dist_sym, T.dist_count);
                    if (d < 0 || d >= 30) { err = 1u; mlen = 0; break; }
                    base = kInfDistBase[d]; extra = kInfDistExtra[d];
                } else { err = 1u; mlen = 0; break; }
                br_fill(br);
                mdist = base + br_take(br, extra);
                if (mdist > pos || pos + mlen > cap || mdist > 32768u) { err = 2u; mlen = 0; }
                break;
            }
        }
        const uint32_t mm = __shfl_sync(kFull, mlen | (mdist << 16), lead);   // (mlen <= 258, mdist <= 32768)
        mlen = mm & 0xFFFFu; mdist = mm >> 16;
        const uint32_t p0 = __shfl_sync(kFull, pos, lead);
        // the previous match's bytes were only LOADED when it was decoded (the decoder does not need them to go on):
        // they are stored now, one decode run later, when the loads have long returned
        if (pend_pos != 0xFFFFFFFFu) { dst[pend_pos] = pend_val; pend_pos = 0xFFFFFFFFu; }
        __syncwarp();                                                 // ... and the leader's literal stores: visible to the group
        if (mlen != 0u) {
            const uint8_t* s = dst + p0 - mdist;
            if (mlen <= (uint32_t)kInfLanes) {                         // one byte per lane: load now, store at the next sync point
                if (gl < mlen) { pend_val = s[mdist >= mlen ? gl : gl % mdist]; pend_pos = p0 + gl; }
            } else {
                if (mdist >= mlen) {
                    for (uint32_t i = gl; i < mlen; i += kInfLanes) dst[p0 + i] = s[i];
                } else {
                    for (uint32_t i = gl; i < mlen; i += kInfLanes) dst[p0 + i] = s[i % mdist];   // the pattern repeats
                }
            }
        }
        __syncwarp();                                                 // (outside every branch: groups differ in what they do)
        if (leader && mlen != 0u) pos = p0 + mlen;
        // ---- members that are through ----
        if (leader && br.overrun) err = 3u;
        if (__any_sync(kFull, leader && !done && (finished || err != 0u))) {   // (rare: somebody is through)
            const uint32_t fin = __shfl_sync(kFull, (uint32_t)(finished || err != 0u), lead);
            if (fin != 0u && !done) {
                if (pend_pos != 0xFFFFFFFFu) { dst[pend_pos] = pend_val; pend_pos = 0xFFFFFFFFu; }
                done = true;
                if (leader) {
                    uint32_t st = err;
                    if (st == 0u && br_consumed(br) > (uint64_t)M.src_len * 8u) st = 3u;
                    if (st == 0u && pos != cap) st = 4u;
                    status[m] = st;
                }
            }
            if (__all_sync(kFull, done)) break;
        }
    }
}

#ifndef BSQ_INF_UNIFORM
#define BSQ_INF_UNIFORM 1
#endif

// ---- bit reader of the shipped kernel: a 64-bit window (lo, hi) of the payload and a bit offset into it.  One funnel shift
// yields the next 32 bits; consuming bits is an add to the offset.  br32_fill() moves the window on by a word once 32 bits are
// used up, so after it `off < 32` and br32_window() holds 32 valid bits -- enough for a literal/length code with its extra bits
// (<= 20) or a distance code with its extra bits (<= 28).  The word after the window is prefetched.
struct Br32 {
    const uint32_t* wp;                    // next word to fetch
    const uint32_t* wend;                  // one past the last word that holds payload bytes
    const uint32_t* w0;                    // first word
    uint32_t lo, hi, next;                 // window and the prefetched word
    uint32_t off;                          // bits of (lo, hi) already consumed (< 64)
    uint32_t lead;                         // bits of the first word that precede the payload (8 x misalignment)
    uint32_t base_bytes;                   // payload bytes before w0's payload start (after a stored block)
    uint32_t overrun;                      // the reader ran words past the payload (a damaged stream)
};
__device__ __forceinline__ uint32_t br32_word(Br32& b) {
    const uint32_t v = b.wp < b.wend ? __ldg(b.wp) : 0u;          // (past the end: zeros)
    ++b.wp;
    return v;
}
__device__ __forceinline__ void br32_init(Br32& b, const uint8_t* p, uint32_t nbytes, uint32_t base_bytes) {
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
    b.w0 = b.wp = reinterpret_cast<const uint32_t*>(p - a);
    b.wend = b.w0 + ((a + nbytes + 3u) >> 2);
    b.lo = br32_word(b); b.hi = br32_word(b); b.next = br32_word(b);
    b.off = 8u * a; b.lead = 8u * a;
    b.base_bytes = base_bytes;
    b.overrun = 0;
}
__device__ __forceinline__ void br32_fill(Br32& b) {
    if (b.off >= 32u) {
        b.lo = b.hi; b.hi = b.next;
        if (b.wp > b.wend + 4) b.overrun = 1u;                    // a damaged stream must end the member, not spin on zeros
        b.next = br32_word(b);
        b.off -= 32u;
    }
}
__device__ __forceinline__ uint32_t br32_window(const Br32& b) { return __funnelshift_r(b.lo, b.hi, b.off); }
// payload bits consumed so far
__device__ __forceinline__ uint64_t br32_consumed(const Br32& b) {
    return (uint64_t)b.base_bytes * 8u + (uint64_t)(b.wp - b.w0 - 3) * 32u + b.off - b.lead;
}
// n <= 16 bits (block headers)
__device__ __forceinline__ uint32_t br32_take(Br32& b, uint32_t n) {
    br32_fill(b);
    const uint32_t v = br32_window(b) & ((1u << n) - 1u);
    b.off += n;
    return v;
}
__device__ __forceinline__ uint32_t inf_lds(uint32_t addr) {          // one table entry (32-bit shared-memory address)
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// canonical decode of a code longer than the lookup width from the window's bits; returns the symbol or -1
__device__ int32_t inf_slow32(uint32_t w, const uint16_t* sym, const uint16_t* count, uint32_t& used) {
    int32_t code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; ++l) {
        code |= (int32_t)(w & 1u);
        w >>= 1;
        const int32_t c = count[l];
        if (code - c < first) { used = (uint32_t)l; return sym[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    used = 15u;
    return -1;
}

// The shipped form (one warp per member, kInfLanes == 32): EVERY lane runs the decode loop on identical state -- the
// same input words (one broadcast load), the same table lookups (one shared-memory broadcast) -- so the warp never
// leaves the loop: no leader hand-off, no shuffles, no votes.  A redundant lane costs nothing on a SIMT machine (the
// warp instruction issues once either way), and every lane already knows (position, length, distance) when a match
// comes up: lane i copies byte i.  A literal is stored by every lane (one address, one value: one transaction).  Only the
// block headers (a few per member) are parsed by lane 0 alone and the reader state is broadcast afterwards.  As before, a
// match's bytes are loaded when it is decoded and stored at the next match (the decoder never waits for them).
__global__ void __launch_bounds__(kInfWarps * 32) k_inflate_members_uniform(const uint8_t* __restrict__ zbuf, uint8_t* __restrict__ out,
                                                                          const InflateMember* __restrict__ members, uint32_t n_members,
                                                                          uint32_t* __restrict__ status) {
    extern __shared__ __align__(16) uint8_t inf_smem[];
    constexpr uint32_t kFull = 0xFFFFFFFFu, kNone = 0xFFFFFFFFu;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t m = blockIdx.x * kInfWarps + warp;
    if (m >= n_members) return;                       // (whole warp)
    InflateTables& T = reinterpret_cast<InflateTables*>(inf_smem)[warp];
    const InflateMember M = members[m];
    uint8_t* const dst = out + M.dst;
    const uint8_t* const src0 = zbuf + M.src;
    const uint32_t cap = M.isize;
    // shared-memory addresses of the two lookup tables, kept in registers (opaque to the compiler, which otherwise rebuilds
    // them from the thread index in front of every lookup)
    uint32_t lit_base = (uint32_t)__cvta_generic_to_shared(T.lit), dist_base = (uint32_t)__cvta_generic_to_shared(T.dist);
    asm volatile("" : "+r"(lit_base), "+r"(dist_base));
    Br32 br{};
    br32_init(br, src0, M.src_len, 0u);
    uint32_t pos = 0, err = 0;
    bool last = false, finished = false;
    int tables = 0;                                   // 0 none, 1 fixed, 2 dynamic
    uint32_t pend_pos = kNone;                        // a match byte loaded but not yet stored (per lane)
    uint32_t pend_val = 0;

    while (!finished && err == 0u) {
        // ---- block header: every lane reads the three header bits; lane 0 alone sets the tables up ----
        last = br32_take(br, 1) != 0u;
        const uint32_t btype = br32_take(br, 2);
        if (btype == 0u) {
            br.off = (br.off + 7u) & ~7u;                                 // to the byte boundary
            const uint32_t len = br32_take(br, 16);
            const uint32_t nlen = br32_take(br, 16);
            const uint32_t st_src = (uint32_t)(br32_consumed(br) >> 3);   // payload offset of the raw bytes
            if ((len ^ nlen) != 0xFFFFu) { err = 1u; break; }
            if (pos + len > cap) { err = 2u; break; }
            if (st_src + len > M.src_len) { err = 3u; break; }
            if (pend_pos != kNone) { dst[pend_pos] = (uint8_t)pend_val; pend_pos = kNone; }
            for (uint32_t i = lane; i < len; i += 32u) dst[pos + i] = src0[st_src + i];
            pos += len;
            br32_init(br, src0 + st_src + len, M.src_len - st_src - len, st_src + len);
            if (last) finished = true;
            continue;
        }
        if (btype == 3u) { err = 1u; break; }
        {
            // lane 0 parses the rest of the header and builds the tables; its reader state and verdict are broadcast
            uint32_t herr = 0;
            __syncwarp();                                                 // every lane is through with the previous block's tables
            if (lane == 0u) {
                if (btype == 1u) {
                    if (tables != 1) {
                        for (int s = 0; s < 144; ++s) T.lens[s] = 8;
                        for (int s = 144; s < 256; ++s) T.lens[s] = 9;
                        for (int s = 256; s < 280; ++s) T.lens[s] = 7;
                        for (int s = 280; s < 288; ++s) T.lens[s] = 8;
                        inf_build(T.lens, 288, T.lit, kLitBits, T.lit_sym, T.lit_count, false);
                        for (int s = 0; s < 30; ++s) T.lens[s] = 5;
                        inf_build(T.lens, 30, T.dist, kDistBits, T.dist_sym, T.dist_count, true);
                    }
                } else {
                    const uint32_t hlit = br32_take(br, 5) + 257u, hdist = br32_take(br, 5) + 1u, hclen = br32_take(br, 4) + 4u;
                    if (hlit > 286u || hdist > 30u) herr = 1u;
                    uint8_t* cl = T.lens + 300;                             // 19 code-length code lengths
                    for (int i = 0; i < 19; ++i) cl[i] = 0;
                    for (uint32_t i = 0; i < hclen && herr == 0u; ++i) cl[kInfClOrder[i]] = (uint8_t)br32_take(br, 3);
                    // the code-length code borrows the distance table's storage (7-bit lookup)
                    if (herr == 0u && !inf_build(cl, 19, T.dist, 7, T.dist_sym, T.dist_count, false)) herr = 1u;
                    uint32_t i = 0;
                    while (i < hlit + hdist && herr == 0u) {
                        br32_fill(br);
                        const uint32_t e = T.dist[br32_window(br) & 127u];
                        if ((e & 0xF0u) != 0u) { herr = 1u; break; }        // not a (valid) symbol of the code-length code
                        br.off += e & 15u;
                        const uint32_t s = e >> 16;
                        if (s < 16u) { T.lens[i++] = (uint8_t)s; continue; }
                        uint32_t rep, val = 0;
                        if (s == 16u) { if (i == 0u) { herr = 1u; break; } val = T.lens[i - 1]; rep = 3u + br32_take(br, 2); }
                        else if (s == 17u) rep = 3u + br32_take(br, 3);
                        else rep = 11u + br32_take(br, 7);
                        if (i + rep > hlit + hdist) { herr = 1u; break; }
                        while (rep--) T.lens[i++] = (uint8_t)val;
                    }
                    if (herr == 0u && T.lens[256] == 0) herr = 1u;            // no end-of-block code
                    if (herr == 0u) {
                        for (uint32_t k = 0; k < 32u; ++k) T.dl[k] = k < hdist ? T.lens[hlit + k] : 0;
                        if (!inf_build(T.lens, hlit, T.lit, kLitBits, T.lit_sym, T.lit_count, false)) herr = 1u;
                        if (herr == 0u && !inf_build(T.dl, hdist, T.dist, kDistBits, T.dist_sym, T.dist_count, true)) herr = 1u;
                    }
                }
            }
            tables = (int)btype;
            __syncwarp();                                                 // the tables are in shared memory for every lane
            herr = __shfl_sync(kFull, herr, 0);
            if (btype == 2u) {
                // the reader moved in lane 0 only
                const uint32_t adv = __shfl_sync(kFull, (uint32_t)(br.wp - br.w0), 0);
                br.wp = br.w0 + adv;
                br.lo = __shfl_sync(kFull, br.lo, 0);
                br.hi = __shfl_sync(kFull, br.hi, 0);
                br.next = __shfl_sync(kFull, br.next, 0);
                br.off = __shfl_sync(kFull, br.off, 0);
                br.overrun = __shfl_sync(kFull, br.overrun, 0);
            }
            if (herr != 0u) { err = herr; break; }
        }
        // ---- the block's symbols: every lane decodes, every lane stores the literal, lane i copies byte i of a match ----
        while (true) {
            br32_fill(br);
            uint32_t w = br32_window(br);
            uint32_t e = inf_lds(lit_base + ((w & ((1u << kLitBits) - 1u)) << 2));
            if ((e & 0xF0u) == 0u) {                                  // a literal (the common case)
                br.off += e & 15u;
                if (pos >= cap) { err = 2u; break; }
                dst[pos] = (uint8_t)(e >> 16);
                ++pos;
                continue;
            }
            uint32_t kind = (e >> 4) & 15u;
            uint32_t base, extra, used;
            if (kind == kKindLen) {
                used = e & 15u; base = e >> 16; extra = (e >> 8) & 255u;
            } else if (kind == kKindEob) {
                br.off += e & 15u;
                if (last) finished = true;
                break;
            } else if (kind == kKindSlow) {                           // a code longer than the table width
                const int32_t s = inf_slow32(w, T.lit_sym, T.lit_count, used);
                if (s < 0 || s >= 286) { err = 1u; break; }
                if (s < 256) {
                    br.off += used;
                    if (pos >= cap) { err = 2u; break; }
                    dst[pos] = (uint8_t)s;
                    ++pos;
                    continue;
                }
                if (s == 256) { br.off += used; if (last) finished = true; break; }
                base = kInfLenBase[s - 257]; extra = kInfLenExtra[s - 257];
            } else { err = 1u; break; }
            const uint32_t mlen = base + ((w >> used) & ((1u << extra) - 1u));
            br.off += used + extra;                                   // <= 20 bits of the window
            br32_fill(br);
            w = br32_window(br);
            e = inf_lds(dist_base + ((w & ((1u << kDistBits) - 1u)) << 2));
            kind = (e >> 4) & 15u;
            if (kind == kKindLen) {
                used = e & 15u; base = e >> 16; extra = (e >> 8) & 255u;
            } else if (kind == kKindSlow) {
                const int32_t d = inf_slow32(w, T.dist_sym, T.dist_count, used);
                if (d < 0 || d >= 30) { err = 1u; break; }
                base = kInfDistBase[d]; extra = kInfDistExtra[d];
            } else { err = 1u; break; }
            const uint32_t mdist = base + ((w >> used) & ((1u << extra) - 1u));
            br.off += used + extra;                                   // <= 28 bits of the window
            if (mdist > pos || pos + mlen > cap || mdist > 32768u) { err = 2u; break; }
            // the previous match's bytes were only LOADED when it was decoded: they are stored now, when the loads have long
            // returned; then the literal stores and these are visible to the whole warp
            if (pend_pos != kNone) { dst[pend_pos] = (uint8_t)pend_val; pend_pos = kNone; }
            __syncwarp();
            const uint8_t* s = dst + pos - mdist;
            if (mlen <= 32u) {                                         // one byte per lane: load now, store at the next match
                if (lane < mlen) {
                    uint32_t i = lane;
                    if (mdist < mlen) i = lane % mdist;               // the pattern repeats
                    pend_val = s[i];
                    pend_pos = pos + lane;
                }
            } else if (mdist >= mlen) {
                for (uint32_t i = lane; i < mlen; i += 32u) dst[pos + i] = s[i];
            } else {
                for (uint32_t i = lane; i < mlen; i += 32u) dst[pos + i] = s[i % mdist];
            }
            pos += mlen;
        }
        if (br.overrun) err = 3u;
    }
    if (pend_pos != kNone) dst[pend_pos] = (uint8_t)pend_val;
    if (lane == 0u) {
        uint32_t st = err;
        if (st == 0u && br.overrun) st = 3u;
        if (st == 0u && br32_consumed(br) > (uint64_t)M.src_len * 8u) st = 3u;
        if (st == 0u && pos != cap) st = 4u;
        status[m] = st;
    }
}

// CRC-32 (gzip, reflected 0xEDB88320) of every member's inflated bytes: one warp per member, each lane a
// contiguous slice, the slices' CRCs combined with x^(8 n) mod P multiplications.
__device__ __forceinline__ uint32_t crc_mul(uint32_t a, uint32_t b) {      // a * b mod P, bit-reflected operands
    uint32_t p = 0;
    for (int i = 0; i < 32; ++i) {
        if (a & 0x80000000u) p ^= b;
        a <<= 1;
        b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);
    }
    return p;
}
__device__ __forceinline__ uint32_t crc_xpow8n(uint32_t n) {               // x^(8 n) mod P
    uint32_t r = 0x80000000u, base = 0x00800000u;                          // 1, x^8
    while (n) {
        if (n & 1u) r = crc_mul(r, base);
        base = crc_mul(base, base);
        n >>= 1;
    }
    return r;
}
constexpr int kCrcWarps = 4;
__global__ void __launch_bounds__(kCrcWarps * 32) k_crc32_members(const uint8_t* __restrict__ out, const InflateMember* __restrict__ members,
                                                                uint32_t n_members, uint32_t* __restrict__ status) {
    __shared__ uint32_t tab[256];
    for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? 0xEDB88320u : 0u);
        tab[i] = c;
    }
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t m = blockIdx.x * kCrcWarps + warp;
    if (m >= n_members) return;
    const InflateMember M = members[m];
    // lane 0 takes the first per + (n mod 32) bytes, every other lane `per` bytes: all right-hand operands of the
    // combine tree are multiples of `per` long, so one power of x serves a whole level
    const uint32_t n = M.isize, per = n / 32u, rem = n - 32u * per;
    const uint32_t a = lane == 0u ? 0u : rem + lane * per, b = rem + (lane + 1u) * per;
    const uint8_t* p = out + M.dst;
    uint32_t c = 0;                                                        // raw CRC register of the slice (init 0)
    for (uint32_t i = a; i < b; ++i) c = tab[(c ^ p[i]) & 255u] ^ (c >> 8);
    // crc(A || B) = crc(A) * x^(8 |B|) + crc(B) on raw (init 0, no final xor) registers
    uint32_t xp = crc_xpow8n(per);                                         // x^(8 per), squared per level
    for (uint32_t d = 1; d < 32u; d <<= 1) {
        const uint32_t c2 = __shfl_down_sync(0xFFFFFFFFu, c, d);
        if ((lane & (2u * d - 1u)) == 0u) c = crc_mul(c, xp) ^ c2;
        xp = crc_mul(xp, xp);
    }
    if (lane == 0) {
        // standard CRC = raw(init 0xFFFFFFFF) ^ 0xFFFFFFFF, and raw(init I) = raw(init 0) ^ I * x^(8 n)
        const uint32_t crc = c ^ crc_mul(0xFFFFFFFFu, crc_xpow8n(n)) ^ 0xFFFFFFFFu;
        if (status[m] == 0u && crc != M.crc) status[m] = 5u;
    }
}

}  // namespace bsq
