// tile_math.h -- the rank algebra shared by the device kernels and the host-side shard stitcher.
//
// BlazeSeq's parser is newline-count based (blazeseq/utils.mojo:470-551: a record is "the next
// four '\n'"; no resync), so the line class of every byte is (newline rank) mod 4 and the whole
// path is a prefix scan over newline ranks.  A contiguous byte range ("run": the tiles one CTA
// owns, or the shard one GPU owns) is summarised by a small monoid element; an exclusive scan of
// those elements gives every run the state it must start from.  Everything is modulo 2^32 on
// window-relative positions (a window is < 2^31 bytes).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BSQ_HD __host__ __device__ __forceinline__
#else
#define BSQ_HD inline
#endif

// Summary of a byte range.  `count` newlines; last[i] = position of the (i+1)-th most recent
// newline (valid for i < count); first[i] = position of the i-th newline; P[r] = sum of the
// positions of the newlines whose index within the range is == r (mod 4); flags = OR of
// BSQ_SUM_* bits.
struct BsqSummary {
    uint32_t count;
    uint32_t last[4];
    uint32_t P[4];
    uint32_t flags;
    uint32_t first[4];  // positions of the first four newlines (valid for i < count)
    uint32_t _pad[2];
};
static_assert(sizeof(BsqSummary) == 64, "BsqSummary is one 64-byte record");

#define BSQ_SUM_ID_MAY_STRIP 1u  // some header line may need _strip_spaces (utils.mojo:221-242)

// State a run starts from (all newlines before it, in window order).
struct BsqPrefix {
    uint32_t rank;      // newline rank of the run's first newline (window-relative, base == 0 mod 4)
    uint32_t prev[3];   // positions of the 3 newlines before the run (prev[0] most recent)
    uint32_t cum_seq;   // bytes of the sequence lines that ended before the run
    uint32_t cum_qual;  // bytes of the quality lines that ended before the run
    uint32_t cum_id;    // UNSTRIPPED id bytes of the header lines that ended before the run
    uint32_t prev3;     // position of the 4th newline before the run (the start of a record whose last three
                        // lines end in the run: its length is checked against the buffer limit)
};
static_assert(sizeof(BsqPrefix) == 32, "BsqPrefix is 32 bytes");

BSQ_HD BsqSummary bsq_summary_identity() {
    BsqSummary s;
    s.count = 0;
    for (int i = 0; i < 4; ++i) { s.last[i] = 0; s.P[i] = 0; }
    s.flags = 0;
    for (int i = 0; i < 4; ++i) s.first[i] = 0;
    s._pad[0] = s._pad[1] = 0;
    return s;
}

// The state before the first byte of a window that begins (at byte `begin`) on a record
// boundary: no newline seen, but the record's header starts after a VIRTUAL newline at begin-1
// (header_start = previous newline + 1, utils.mojo:423-432 with header_start the record base).
BSQ_HD BsqSummary bsq_summary_window_init(uint32_t begin) {
    BsqSummary s = bsq_summary_identity();
    s.last[0] = begin - 1u;
    return s;
}

// a followed by b.
BSQ_HD BsqSummary bsq_combine(const BsqSummary& a, const BsqSummary& b) {
    BsqSummary r;
    r.count = a.count + b.count;
    const uint32_t cb = b.count < 4u ? b.count : 4u;
    for (uint32_t i = 0; i < 4u; ++i) r.last[i] = i < cb ? b.last[i] : a.last[i - cb];
    const uint32_t rot = a.count & 3u;
    for (uint32_t k = 0; k < 4u; ++k) r.P[k] = a.P[k] + b.P[(k + 4u - rot) & 3u];
    r.flags = a.flags | b.flags;
    const uint32_t ca = a.count < 4u ? a.count : 4u;
    for (uint32_t i = 0; i < 4u; ++i) r.first[i] = i < ca ? a.first[i] : b.first[i - ca];
    r._pad[0] = r._pad[1] = 0;
    return r;
}

// Prefix for a run from the exclusive state E = init (+) all earlier runs.  E.last[0] is the
// virtual newline when E.count == 0.  `begin` = window begin offset.
BSQ_HD BsqPrefix bsq_prefix_from(const BsqSummary& E, uint32_t begin) {
    BsqPrefix p;
    const uint32_t G = E.count;
    p.rank = G;
    p.prev[0] = E.last[0]; p.prev[1] = E.last[1]; p.prev[2] = E.last[2];
    const uint32_t ph = G & 3u;
    // class-1 newlines (end of sequence line) before the run: ranks 1,5,9,... < G
    const uint32_t n1 = (G + 2u) >> 2, n3 = G >> 2, n0 = (G + 3u) >> 2;
    // a class-0 newline whose successor is not yet seen does not open a finished sequence line
    p.cum_seq = E.P[1] - (E.P[0] - (ph == 1u ? E.last[0] : 0u)) - n1;
    p.cum_qual = E.P[3] - (E.P[2] - (ph == 3u ? E.last[0] : 0u)) - n3;
    // header line r (class 0) spans (pos[r-1], pos[r]); pos[-1] is the virtual newline begin-1
    const uint32_t open3 = (ph == 0u && G > 0u) ? E.last[0] : 0u;
    p.cum_id = E.P[0] - (E.P[3] - open3) - (G > 0u ? (begin - 1u) : 0u) - 2u * n0;
    p.prev3 = E.last[3];
    return p;
}

// Totals over the COMPLETE records of a window from the state after its last byte.
struct BsqTotals {
    uint32_t newlines, records, consumed_end, seq_bytes, qual_bytes, id_bytes_unstripped, flags, _pad;
};

BSQ_HD BsqTotals bsq_totals_from(const BsqSummary& E_end, uint32_t begin) {
    BsqTotals t;
    t.newlines = E_end.count;
    t.records = E_end.count >> 2;
    const uint32_t rem = E_end.count & 3u;
    // drop the trailing `rem` newlines (they belong to an incomplete record)
    BsqSummary c = E_end;
    for (uint32_t i = 0; i < rem; ++i) c.P[rem - 1u - i] -= E_end.last[i];  // classes rem-1 .. 0
    c.count = E_end.count - rem;
    for (uint32_t i = 0; i < 4u; ++i) c.last[i] = (i + rem < 4u) ? E_end.last[i + rem] : 0u;
    // with rem == 3 the 4th-last newline is last[3]; for c.count > 0 it is the last record end
    if (c.count == 0u) c.last[0] = begin - 1u;
    BsqPrefix p = bsq_prefix_from(c, begin);
    t.consumed_end = t.records > 0u ? c.last[0] + 1u : begin;
    t.seq_bytes = p.cum_seq;
    t.qual_bytes = p.cum_qual;
    // all complete: the last class-3 newline opens no counted header, bsq_prefix_from excluded it
    t.id_bytes_unstripped = p.cum_id;
    t.flags = E_end.flags;
    t._pad = 0;
    return t;
}

// ---- the part of a BsqSummary that crosses tiles in the single-pass kernel (decoupled look-back) ----
// count, the last four newline positions and the position sums: what bsq_prefix_from and
// bsq_totals_from read.  lb_combine is bsq_combine restricted to those fields, branch-free.
// NOTE: the window-init state (count 0, last[0] = begin-1) is only valid as the LEFTMOST operand.
struct LbState { uint32_t count, last[4], P[4]; };

BSQ_HD LbState lb_identity() {
    LbState s;
    s.count = 0;
    for (int i = 0; i < 4; ++i) { s.last[i] = 0; s.P[i] = 0; }
    return s;
}
// a followed by b
BSQ_HD LbState lb_combine(const LbState& a, const LbState& b) {
    LbState r;
    r.count = a.count + b.count;
    const uint32_t cb = b.count < 4u ? b.count : 4u;
    r.last[0] = cb > 0u ? b.last[0] : a.last[0];
    r.last[1] = cb > 1u ? b.last[1] : (cb == 1u ? a.last[0] : a.last[1]);
    r.last[2] = cb > 2u ? b.last[2] : (cb == 2u ? a.last[0] : (cb == 1u ? a.last[1] : a.last[2]));
    r.last[3] = cb > 3u ? b.last[3] : (cb == 3u ? a.last[0] : (cb == 2u ? a.last[1] : (cb == 1u ? a.last[2] : a.last[3])));
    // r.P[k] = a.P[k] + b.P[(k - a.count) mod 4]: rotate b.P right by (a.count & 3)
    uint32_t q0 = b.P[0], q1 = b.P[1], q2 = b.P[2], q3 = b.P[3];
    if (a.count & 1u) { const uint32_t t = q3; q3 = q2; q2 = q1; q1 = q0; q0 = t; }
    if (a.count & 2u) { uint32_t t = q0; q0 = q2; q2 = t; t = q1; q1 = q3; q3 = t; }
    r.P[0] = a.P[0] + q0; r.P[1] = a.P[1] + q1; r.P[2] = a.P[2] + q2; r.P[3] = a.P[3] + q3;
    return r;
}
BSQ_HD BsqSummary lb_to_summary(const LbState& v) {
    BsqSummary s = bsq_summary_identity();
    s.count = v.count;
    for (int i = 0; i < 4; ++i) { s.last[i] = v.last[i]; s.P[i] = v.P[i]; }
    return s;
}

// is_posix_space, utils.mojo:266-289: {9,10,11,12,13,28,29,30,32}
BSQ_HD bool bsq_is_space(uint32_t c) {
    return c == 32u || (c < 32u && ((0x70003E00u >> c) & 1u));
}

// ---- byte-lane SIMD-in-register helpers: results have 0x80 in each byte lane that matches ----

// bytes equal to '\n' (0x0A): exact zero-byte test on w ^ 0x0A0A0A0A
BSQ_HD uint32_t bsq_nl_flags(uint32_t w) {
    const uint32_t x = w ^ 0x0A0A0A0Au;
    const uint32_t t = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x;
    return ~t & 0x80808080u;
}
// bytes with bit 7 set (_check_ascii, utils.mojo:245-263)
BSQ_HD uint32_t bsq_hi_flags(uint32_t w) { return w & 0x80808080u; }
// bytes outside [lower, upper], upper < 128 (Validator._validate_quality_range, record.mojo:76-104,
// inclusive bounds).  addlo = (128-lower)*0x01010101, addup = (127-upper)*0x01010101.
BSQ_HD uint32_t bsq_badq_flags(uint32_t w, uint32_t addlo, uint32_t addup) {
    const uint32_t l = w & 0x7F7F7F7Fu;
    const uint32_t ge_lo = l + addlo;   // bit7 set iff l >= lower
    const uint32_t gt_up = l + addup;   // bit7 set iff l >  upper
    return (w | ~ge_lo | gt_up) & 0x80808080u;
}
// gather the four lane flags (bits 7,15,23,31) into the top nibble, byte order preserved
BSQ_HD uint32_t bsq_gather_top(uint32_t f) { return f * 0x00204081u; }
