// tile_model.cpp -- CPU walk-through of the device decomposition (TEST INFRASTRUCTURE ONLY).
//
// Runs the same rank algebra the kernels use (blazeseq_b200/csrc/tile_math.h: run summaries,
// their exclusive scan, the derived per-run prefix) over byte ranges of arbitrary size, one run
// at a time and with no knowledge of the bytes outside the run except through BsqPrefix.  The
// tests compare what it resolves (record offsets, SoA destinations, totals) with the oracle, so
// the formulas are pinned on the CPU before any kernel runs.  It is not a parser fallback: the
// product never links it.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../blazeseq_b200/csrc/tile_math.h"

extern "C" {

// pass 1: summary of [lo, hi) with positions relative to the window
static BsqSummary summarize(const uint8_t* d, uint32_t lo, uint32_t hi) {
    BsqSummary s = bsq_summary_identity();
    for (uint32_t p = lo; p < hi; ++p) {
        if (d[p] != '\n') continue;
        s.P[s.count & 3u] += p;
        if (s.count < 4u) s.first[s.count] = p;
        s.last[3] = s.last[2]; s.last[2] = s.last[1]; s.last[1] = s.last[0]; s.last[0] = p;
        s.count++;
    }
    return s;
}

struct tm_record {
    uint32_t header_start, seq_start, sep_start, qual_start, record_end;
    uint32_t seq_dst, qual_dst, id_dst_unstripped;  // exclusive cumulative lengths before the record
    uint32_t code;                                   // structure code 0..3
};

// Returns the number of complete records; fills recs (cap entries) and totals[8] (BsqTotals).
int64_t tm_parse(const uint8_t* d, uint32_t begin, uint32_t end, uint32_t run_bytes,
                 tm_record* recs, int64_t cap, uint32_t* totals_out) {
    if (run_bytes == 0) run_bytes = 1;
    std::vector<uint32_t> lo, hi;
    // runs are aligned to multiples of run_bytes like tiles are aligned to the window base
    for (uint64_t a = (begin / run_bytes) * (uint64_t)run_bytes; a < end; a += run_bytes) {
        uint32_t l = a < begin ? begin : (uint32_t)a;
        uint32_t h = a + run_bytes < end ? (uint32_t)(a + run_bytes) : end;
        lo.push_back(l); hi.push_back(h);
    }
    size_t nr = lo.size();
    std::vector<BsqSummary> sum(nr);
    for (size_t r = 0; r < nr; ++r) sum[r] = summarize(d, lo[r], hi[r]);
    // exclusive scan
    std::vector<BsqPrefix> pre(nr);
    BsqSummary E = bsq_summary_window_init(begin);
    for (size_t r = 0; r < nr; ++r) { pre[r] = bsq_prefix_from(E, begin); E = bsq_combine(E, sum[r]); }
    BsqTotals tot = bsq_totals_from(E, begin);
    memcpy(totals_out, &tot, sizeof tot);
    // pass 2: every run resolves the lines that END inside it, from its prefix only
    for (size_t r = 0; r < nr; ++r) {
        BsqPrefix p = pre[r];
        uint32_t rank = p.rank, q1 = p.prev[0], q2 = p.prev[1], q3 = p.prev[2];
        uint32_t cs = p.cum_seq, cq = p.cum_qual, ci = p.cum_id;
        for (uint32_t x = lo[r]; x < hi[r]; ++x) {
            if (d[x] != '\n') continue;
            uint32_t cls = rank & 3u;
            int64_t k = rank >> 2;
            bool live = k < (int64_t)tot.records && k < cap;
            uint32_t len = x - q1 - 1u;
            if (cls == 0) {
                if (live) {
                    recs[k].header_start = q1 + 1u; recs[k].seq_start = x + 1u;
                    recs[k].id_dst_unstripped = ci; recs[k].code = 0;
                    if (d[q1 + 1u] != '@') recs[k].code = 1;
                }
                ci += len - 1u;  // id = line minus the '@'
            } else if (cls == 1) {
                if (live) { recs[k].sep_start = x + 1u; recs[k].seq_dst = cs; }
                cs += len;
            } else if (cls == 2) {
                if (live) {
                    recs[k].qual_start = x + 1u;
                    if (recs[k].code == 0 && d[q1 + 1u] != '+') recs[k].code = 2;
                }
            } else {
                if (live) {
                    recs[k].record_end = x; recs[k].qual_dst = cq;
                    uint32_t seq_len = q2 - q3 - 1u;
                    if (recs[k].code == 0 && seq_len != len) recs[k].code = 3;
                }
                cq += len;
            }
            q3 = q2; q2 = q1; q1 = x; rank++;
        }
    }
    return tot.records;
}

// Ordered 32-wide tree reductions of lb_combine (the non-commutative monoid k_scan_runs scans with warp
// shuffles), exercised the way a decoupled look-back would use them: tile `ti` combines 32 predecessors per
// round, nearest first, up to the nearest one whose INCLUSIVE state is known (`inc_every`: every n-th tile
// counts as inclusive; 0 = none, so that the walk reaches the window-init state).  Returns the number of tiles whose prefix or whose inclusive
// state differs from the sequential scan (0 = the algebra holds).
int64_t tm_lookback_check(const uint8_t* d, uint32_t begin, uint32_t end, uint32_t tile_bytes, uint32_t inc_every) {
    if (tile_bytes == 0) tile_bytes = 1;
    std::vector<LbState> agg, inc;
    std::vector<BsqPrefix> want;
    BsqSummary E = bsq_summary_window_init(begin);
    for (uint64_t a = (begin / tile_bytes) * (uint64_t)tile_bytes; a < end; a += tile_bytes) {
        const uint32_t l = a < begin ? begin : (uint32_t)a;
        const uint32_t h = a + tile_bytes < end ? (uint32_t)(a + tile_bytes) : end;
        const BsqSummary s = summarize(d, l, h);
        LbState v = lb_identity();
        v.count = s.count;
        for (int i = 0; i < 4; ++i) { v.last[i] = s.last[i]; v.P[i] = s.P[i]; }
        agg.push_back(v);
        want.push_back(bsq_prefix_from(E, begin));
        E = bsq_combine(E, s);
        LbState w = lb_identity();
        w.count = E.count;
        for (int i = 0; i < 4; ++i) { w.last[i] = E.last[i]; w.P[i] = E.P[i]; }
        inc.push_back(w);
    }
    int64_t bad = 0;
    for (size_t ti = 0; ti < agg.size(); ++ti) {
        LbState acc = lb_identity();
        bool have = false;
        int64_t base = (int64_t)ti - 1;
        while (true) {
            LbState v[32];
            uint32_t incmask = 0;
            for (int lane = 0; lane < 32; ++lane) {
                const int64_t t = base - lane;
                v[lane] = lb_identity();
                if (t < 0) {
                    incmask |= 1u << lane;
                    if (t == -1) v[lane].last[0] = begin - 1u;
                } else if (inc_every != 0 && (uint64_t)t % inc_every == 0) {
                    incmask |= 1u << lane;
                    v[lane] = inc[(size_t)t];
                } else {
                    v[lane] = agg[(size_t)t];
                }
            }
            const uint32_t nearest = incmask ? (uint32_t)__builtin_ctz(incmask) : 31u;
            for (uint32_t dd = 1; dd < 32u; dd <<= 1) {
                LbState nv[32];
                for (uint32_t lane = 0; lane < 32u; ++lane)
                    nv[lane] = (lane + dd <= nearest) ? lb_combine(v[lane + dd], v[lane]) : v[lane];
                for (uint32_t lane = 0; lane < 32u; ++lane) v[lane] = nv[lane];
            }
            acc = have ? lb_combine(v[0], acc) : v[0];
            have = true;
            if (incmask) break;
            base -= 32;
        }
        const BsqPrefix got = bsq_prefix_from(lb_to_summary(acc), begin);
        const LbState I = lb_combine(acc, agg[ti]);
        bool ok = memcmp(&got, &want[ti], sizeof got) == 0 && I.count == inc[ti].count;
        for (int i = 0; i < 4; ++i) {
            ok = ok && I.P[i] == inc[ti].P[i];
            if ((uint32_t)i < I.count + 1u) ok = ok && I.last[i] == inc[ti].last[i];   // +1: the virtual newline
        }
        if (!ok) ++bad;
    }
    return bad;
}

// 64-byte BsqSummary of d[lo, hi) with shard-relative positions (what bsq_summarize_device returns)
void tm_summarize(const uint8_t* d, uint32_t lo, uint32_t hi, uint32_t* out16) {
    BsqSummary s = summarize(d + lo, 0, hi - lo);
    memcpy(out16, &s, sizeof s);
}

int tm_is_space(uint32_t c) { return bsq_is_space(c) ? 1 : 0; }

// byte-lane helpers against a scalar definition; returns the number of mismatching words
int64_t tm_check_flags(const uint32_t* words, int64_t n, uint32_t lower, uint32_t upper) {
    int64_t bad = 0;
    uint32_t addlo = (128u - lower) * 0x01010101u, addup = (127u - upper) * 0x01010101u;
    for (int64_t i = 0; i < n; ++i) {
        uint32_t w = words[i], nl = 0, hi = 0, bq = 0, nib_nl = 0;
        for (int b = 0; b < 4; ++b) {
            uint32_t c = (w >> (8 * b)) & 255u;
            if (c == 10u) { nl |= 0x80u << (8 * b); nib_nl |= 1u << b; }
            if (c & 0x80u) hi |= 0x80u << (8 * b);
            if (c < lower || c > upper) bq |= 0x80u << (8 * b);
        }
        if (bsq_nl_flags(w) != nl || bsq_hi_flags(w) != hi || bsq_badq_flags(w, addlo, addup) != bq ||
            (bsq_gather_top(nl) >> 28) != nib_nl)
            bad++;
    }
    return bad;
}
}
