/*
 * bsq_oracle.c -- CPU restatement of BlazeSeq's FASTQ hot path (see bsq_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  Never by the product.
 *
 * Parity status: PINNED against the reference's own literal test streams
 * (tests/fastq/test_parser.mojo, tests/test_error_context.mojo), its 70-file
 * corpus expectations (tests/test_data/fastq_parser/README.md +
 * tests/fastq/test_fastq_parser_correctness.mojo) and the Python-binding ids
 * (tests/test_python_bindings.py).  The reference itself (Mojo) cannot be run
 * in this image, so there is no oracle/_ref build; behaviours no reference
 * test pins follow the reference source line by line
 * and are marked "unpinned" where they are implemented.
 *
 * All file:line citations are relative to the reference repo root.
 */
#include "bsq_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#define ORA_NL 10 /* CONSTS.mojo:14 */
#define ORA_CR 13 /* CONSTS.mojo:15 */

/* ------------------------------------------------------------------------ */
/* config / schema                                                           */
/* ------------------------------------------------------------------------ */

void ora_default_config(ora_config* cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->buffer_capacity = 256 * 1024;      /* CONSTS.mojo:26 */
    cfg->buffer_max_capacity = 1LL << 30;   /* CONSTS.mojo:27-28 */
    cfg->buffer_growth_enabled = 0;         /* parser.mojo:64 */
    cfg->check_ascii = 0;                   /* parser.mojo:65 */
    cfg->check_quality = 0;                 /* parser.mojo:66 */
    cfg->q_lower = 33;                      /* generic_schema, quality_schema.mojo:26 */
    cfg->q_upper = 126;
    cfg->q_offset = 33;
    cfg->compat_simd_width = 0;
    cfg->reader_max_read = 0;
}

/* _parse_schema, utils.mojo:612-637; table quality_schema.mojo:26-31 */
int ora_parse_schema(const char* name, uint8_t* lower, uint8_t* upper, uint8_t* offset) {
    static const struct { const char* n; uint8_t lo, up, off; } tab[] = {
        {"sanger", 33, 126, 33},       {"solexa", 59, 126, 64},
        {"illumina_1.3", 64, 126, 64}, {"illumina_1.5", 66, 126, 64},
        {"illumina_1.8", 33, 126, 33}, {"generic", 33, 126, 33},
    };
    for (size_t i = 0; i < sizeof(tab) / sizeof(tab[0]); ++i) {
        if (name && strcmp(name, tab[i].n) == 0) {
            *lower = tab[i].lo; *upper = tab[i].up; *offset = tab[i].off;
            return 0;
        }
    }
    *lower = 33; *upper = 126; *offset = 33; /* falls back to generic, utils.mojo:630-636 */
    return 1;
}

/* ------------------------------------------------------------------------ */
/* small helpers shared by both parsers                                      */
/* ------------------------------------------------------------------------ */

/* is_posix_space, utils.mojo:266-289: {9,10,11,12,13,28,29,30,32} */
static inline int ora_is_space(uint8_t c) {
    if (c > 32) return 0;
    const uint64_t mask = (1ULL << 9) | (1ULL << 10) | (1ULL << 11) | (1ULL << 12) |
                          (1ULL << 13) | (1ULL << 28) | (1ULL << 29) | (1ULL << 30) |
                          (1ULL << 32);
    return (int)((mask >> c) & 1);
}

/* _strip_spaces, utils.mojo:221-242.  In/out: [*start, *start + *len). */
static void ora_strip(const uint8_t* d, int64_t* start, int64_t* len) {
    int64_t s = *start, n = *len;
    if (n <= 0) { *len = n < 0 ? 0 : n; return; }
    if (!ora_is_space(d[s]) && !ora_is_space(d[s + n - 1])) return;
    int64_t a = 0;
    while (a < n && ora_is_space(d[s + a])) a++;
    int64_t e = n;
    while (e > a && ora_is_space(d[s + e - 1])) e--;
    *start = s + a;
    *len = e - a;
}

/* _check_ascii, utils.mojo:245-263 (the SIMD/scalar split has no semantic
 * effect: any byte with bit 7 set fails). */
static int ora_has_high_bit(const uint8_t* p, int64_t n) {
    for (int64_t i = 0; i < n; ++i)
        if (p[i] & 0x80) return 1;
    return 0;
}

/* Validator._validate_quality_range, record.mojo:76-104.
 * compat_w == 0: documented intent LOWER..UPPER inclusive (record.mojo:78).
 * compat_w == W: the first floor(n/W)*W bytes use `>=` (record.mojo:92-94),
 * the tail uses `>` (record.mojo:100).  [unpinned: SURVEY App. A Q5] */
static int ora_quality_bad(const uint8_t* q, int64_t n, uint8_t lower, uint8_t upper,
                           int compat_w) {
    uint8_t span = (uint8_t)(upper - lower);
    int64_t i = 0;
    if (compat_w > 0) {
        for (; i + compat_w <= n; i += compat_w)
            for (int j = 0; j < compat_w; ++j)
                if ((uint8_t)(q[i + j] - lower) >= span) return 1;
    }
    for (; i < n; ++i)
        if ((uint8_t)(q[i] - lower) > span) return 1;
    return 0;
}

/* Validator._validate(FastqView), record.mojo:162-172: ASCII over id, seq,
 * qual in that order (record.mojo:106-116), then the quality range. */
static int ora_validate_view(const uint8_t* d, const ora_view* v, const ora_config* c) {
    if (c->check_ascii) {
        if (ora_has_high_bit(d + v->id_start, v->id_len)) return ORA_ASCII_INVALID;
        if (ora_has_high_bit(d + v->seq_start, v->seq_len)) return ORA_ASCII_INVALID;
        if (ora_has_high_bit(d + v->qual_start, v->qual_len)) return ORA_ASCII_INVALID;
    }
    if (c->check_quality &&
        ora_quality_bad(d + v->qual_start, v->qual_len, c->q_lower, c->q_upper,
                        c->compat_simd_width))
        return ORA_QUALITY_OUT_OF_RANGE;
    return ORA_OK;
}

/* _message_for_code, errors.mojo:71-90 */
static const char* ora_code_message(int code) {
    switch (code) {
    case ORA_ID_NO_AT: return "Sequence id line does not start with '@'";
    case ORA_SEP_NO_PLUS: return "Separator line does not start with '+'";
    case ORA_SEQ_QUAL_LEN_MISMATCH: return "Quality and sequence line do not match in length";
    case ORA_ASCII_INVALID: return "Non ASCII letters found";
    case ORA_QUALITY_OUT_OF_RANGE: return "Corrupt quality score according to provided schema";
    case ORA_UNEXPECTED_EOF: return "Unexpected end of file in FASTQ record";
    case ORA_BUFFER_EXCEEDED: return "FASTQ record exceeds buffer capacity";
    case ORA_BUFFER_AT_MAX: return "FASTQ record exceeds maximum buffer capacity";
    default: return "Parse or validation error";
    }
}

typedef struct { char* p; size_t cap, len; } ora_sb;
static void sb_bytes(ora_sb* s, const void* b, size_t n) {
    if (s->len + n >= s->cap) n = s->cap - 1 - s->len;
    memcpy(s->p + s->len, b, n);
    s->len += n;
    s->p[s->len] = 0;
}
static void sb_str(ora_sb* s, const char* z) { sb_bytes(s, z, strlen(z)); }
static void sb_i64(ora_sb* s, int64_t v) {
    char t[32];
    snprintf(t, sizeof t, "%lld", (long long)v);
    sb_str(s, t);
}

static void ora_err_clear(ora_error* e, int code) {
    if (!e) return;
    e->code = code;
    e->record_number = e->line_number = e->file_position = 0;
    e->message[0] = 0;
}

/* ParseError.write_to, errors.mojo:178-192 (via raise_parse_error :318-333 or
 * format_parse_error_from_code :93-116). */
static void ora_fill_parse_error(ora_error* e, int code, int64_t rec, int64_t line,
                                 int64_t pos, const uint8_t* snip, int64_t snip_len) {
    if (!e) return;
    ora_err_clear(e, code);
    e->record_number = rec; e->line_number = line; e->file_position = pos;
    ora_sb s = {e->message, sizeof e->message, 0};
    sb_str(&s, ora_code_message(code));
    if (rec > 0) { sb_str(&s, "\n  Record number: "); sb_i64(&s, rec); }
    if (line > 0) { sb_str(&s, "\n  Line number: "); sb_i64(&s, line); }
    if (pos > 0) { sb_str(&s, "\n  File position: "); sb_i64(&s, pos); }
    if (snip_len > 0) { sb_str(&s, "\n  Record snippet: "); sb_bytes(&s, snip, (size_t)snip_len); }
}

/* ValidationError.write_to, errors.mojo:223-234, raised from next_view
 * (parser.mojo:163-169) with field "" and the snippet of
 * _get_record_snippet (parser.mojo:597-610). */
static void ora_fill_validation_error(ora_error* e, int code, int64_t rec,
                                      const uint8_t* d, const ora_view* v) {
    if (!e) return;
    ora_err_clear(e, code);
    e->record_number = rec;
    uint8_t snip[512];
    int64_t sl = 0;
    if (v->id_len > 0) {
        int64_t n = v->id_len > 400 ? 400 : v->id_len;
        memcpy(snip, d + v->id_start, (size_t)n);
        sl = n;
        if (sl < 200) snip[sl++] = '\n';
    }
    if (sl < 200 && v->seq_len > 0) {
        int64_t n = v->seq_len < 200 - sl ? v->seq_len : 200 - sl;
        memcpy(snip + sl, d + v->seq_start, (size_t)n);
        sl += n;
    }
    if (sl > 200) { sl = 197; memcpy(snip + sl, "...", 3); sl = 200; }
    ora_sb s = {e->message, sizeof e->message, 0};
    sb_str(&s, ora_code_message(code));
    if (rec > 0) { sb_str(&s, "\n  Record number: "); sb_i64(&s, rec); }
    if (sl > 0) { sb_str(&s, "\n  Record snippet: "); sb_bytes(&s, snip, (size_t)sl); }
}

static void ora_fill_plain_error(ora_error* e, int code, const char* msg) {
    if (!e) return;
    ora_err_clear(e, code);
    ora_sb s = {e->message, sizeof e->message, 0};
    sb_str(&s, msg);
}

/* ------------------------------------------------------------------------ */
/* streaming model: MemoryReader -> BufferedReader -> FastqParser            */
/* ------------------------------------------------------------------------ */

typedef struct { int64_t header_start, seq_start, sep_start, qual_start, record_end; } ora_offsets;

struct ora_parser {
    ora_config cfg;
    /* MemoryReader, readers.mojo:140-223 */
    const uint8_t* data;
    int64_t n, src_pos;
    /* BufferedReader, buffered.mojo:129-149 */
    uint8_t* ptr;
    int64_t len, head, end, stream_position;
    int is_eof;
    /* FastqParser, parser.mojo:82-87 */
    int64_t current_line_number;
};

/* MemoryReader.read_to_buffer, readers.mojo:171-213 (+ reader_max_read to
 * model short reads). */
static int64_t rd_read(ora_parser* p, uint8_t* dst, int64_t amt) {
    if (p->src_pos >= p->n) return 0;
    int64_t avail = p->n - p->src_pos;
    int64_t k = amt < avail ? amt : avail;
    if (p->cfg.reader_max_read > 0 && k > p->cfg.reader_max_read) k = p->cfg.reader_max_read;
    if (k > 0) { memcpy(dst, p->data + p->src_pos, (size_t)k); p->src_pos += k; }
    return k;
}

/* buffered.mojo:262-281 */
static int64_t br_fill(ora_parser* p) {
    if (p->is_eof) return 0;
    int64_t space = p->len - p->end;
    if (space == 0) return 0;
    int64_t amt = rd_read(p, p->ptr + p->end, space);
    p->end += amt;
    if (amt == 0) p->is_eof = 1;
    return amt;
}

/* buffered.mojo:239-260 */
static void br_compact_from(ora_parser* p, int64_t from_pos) {
    if (from_pos == 0) return;
    if (from_pos >= p->end) {
        p->stream_position += p->end;
        p->head = 0; p->end = 0;
        return;
    }
    p->stream_position += from_pos;
    int64_t remaining = p->end - from_pos;
    memmove(p->ptr, p->ptr + from_pos, (size_t)remaining);
    p->head = p->head < from_pos ? 0 : p->head - from_pos;
    p->end = remaining;
}

static inline int64_t br_available(const ora_parser* p) { return p->end - p->head; }

/* buffered.mojo:211-217,292-299 */
static void br_resize(ora_parser* p, int64_t additional, int64_t max_capacity) {
    int64_t nc = p->len + additional;
    if (nc > max_capacity) nc = max_capacity;
    uint8_t* np = (uint8_t*)malloc((size_t)(nc > 0 ? nc : 1));
    memcpy(np, p->ptr, (size_t)(p->len < nc ? p->len : nc));
    free(p->ptr);
    p->ptr = np; p->len = nc;
}

ora_parser* ora_open(const uint8_t* data, size_t n, const ora_config* cfg) {
    ora_parser* p = (ora_parser*)calloc(1, sizeof(*p));
    if (!p) return NULL;
    if (cfg) p->cfg = *cfg; else ora_default_config(&p->cfg);
    p->data = data; p->n = (int64_t)n;
    p->len = p->cfg.buffer_capacity > 0 ? p->cfg.buffer_capacity : 1;
    p->ptr = (uint8_t*)malloc((size_t)p->len);
    br_fill(p); /* buffered.mojo:149 */
    return p;
}

void ora_close(ora_parser* p) {
    if (!p) return;
    free(p->ptr);
    free(p);
}

const uint8_t* ora_data(const ora_parser* p) { return p->data; }

/* parser.mojo:156-157 */
int ora_has_more(const ora_parser* p) { return br_available(p) > 0 || !p->is_eof; }

/* _validate_fastq_structure, utils.mojo:448-462 */
static int ora_structure(const uint8_t* view, const ora_offsets* o) {
    if (view[o->header_start] != '@') return ORA_ID_NO_AT;
    if (view[o->sep_start] != '+') return ORA_SEP_NO_PLUS;
    if (o->sep_start - o->seq_start - 1 != o->record_end - o->qual_start)
        return ORA_SEQ_QUAL_LEN_MISMATCH;
    return ORA_OK;
}

/* _scan_record, utils.mojo:470-551 (scalar restatement: the SIMD body and the
 * scalar tail find the same newlines; offsets per _store_newline_offset
 * :408-432; resume point per _phase_start_offset :332-353). */
static int ora_scan_record(const uint8_t* view, int64_t view_len, ora_offsets* o, int* phase,
                           int* code) {
    int64_t start_rel;
    switch (*phase) {
    case 0: start_rel = o->header_start; break;
    case 1: start_rel = o->seq_start; break;
    case 2: start_rel = o->sep_start; break;
    default: start_rel = o->qual_start; break;
    }
    *code = ORA_OK;
    if (view_len - start_rel <= 0) return 0;
    int found = *phase;
    const uint8_t* cur = view + start_rel;
    const uint8_t* endp = view + view_len;
    while (found < 4) {
        const uint8_t* nl = (const uint8_t*)memchr(cur, ORA_NL, (size_t)(endp - cur));
        if (!nl) break;
        int64_t abs_pos = (nl - view) + 1;
        found++;
        if (found == 1) o->seq_start = abs_pos;
        else if (found == 2) o->sep_start = abs_pos;
        else if (found == 3) o->qual_start = abs_pos;
        else o->record_end = abs_pos - 1;
        cur = nl + 1;
    }
    if (found == 4) {
        *code = ora_structure(view, o);
        *phase = 0;
        return 1;
    }
    *phase = found;
    return 0;
}

/* _check_end_qual, utils.mojo:292-329 */
static int ora_check_end_qual(ora_parser* p, int64_t base, ora_offsets* o) {
    int64_t rest_start = base + o->qual_start;
    int all_blank = 1;
    for (int64_t i = rest_start; i < p->end; ++i) {
        uint8_t b = p->ptr[i];
        if (b != ORA_NL && b != ORA_CR && b != ' ' && b != '\t') { all_blank = 0; break; }
    }
    if (all_blank) return 0;
    o->record_end = p->end - base;
    return 1;
}

/* _next_ref_complete, parser.mojo:451-522.  Returns the refill code; sets
 * *complete. */
static int ora_next_ref_complete(ora_parser* p, int64_t base, ora_offsets* o, int* phase,
                                 int* complete) {
    int64_t new_base = base;
    *complete = 0;
    for (;;) {
        if (br_available(p) < p->len && p->is_eof) {          /* :464 */
            if (*phase == 3) {                                  /* :465-475 */
                *complete = ora_check_end_qual(p, new_base, o);
                return ORA_OK;
            }
            return ORA_UNEXPECTED_EOF;                          /* :476-482 */
        }
        if (new_base == 0) {                                    /* :484 */
            if (!p->cfg.buffer_growth_enabled) return ORA_BUFFER_EXCEEDED; /* :486-492 */
            int64_t cur = p->len, mx = p->cfg.buffer_max_capacity;
            if (cur >= mx) return ORA_BUFFER_AT_MAX;            /* :495-501 */
            int64_t growth = cur < mx - cur ? cur : mx - cur;   /* :502 */
            br_resize(p, growth, mx);
        } else {
            br_compact_from(p, new_base);                       /* :505-506 */
            new_base = 0;
        }
        int64_t filled = br_fill(p);                            /* :508 */
        if (filled == 0 && br_available(p) == 0) return ORA_EOF; /* :509-510 */
        int code;
        int done = ora_scan_record(p->ptr + new_base, p->end - new_base, o, phase, &code);
        if (done) { *complete = 1; return code; }               /* :521-522 */
    }
}

/* _find_and_consume_ref_record, parser.mojo:311-379.  Returns ORA_OK and the
 * view (absolute stream offsets), or the error. */
static int ora_find_and_consume(ora_parser* p, ora_view* out, ora_error* err) {
    if (br_available(p) == 0) {                                 /* :314-315 */
        br_compact_from(p, p->head);
        br_fill(p);
    }
    if (!ora_has_more(p)) {                                     /* :316-317 */
        ora_fill_plain_error(err, ORA_EOF, "EOF");
        return ORA_EOF;
    }
    int64_t base = p->head;
    ora_offsets o = {0, 0, 0, 0, 0};
    int phase = 0, code = ORA_OK;
    const uint8_t* scan_view = p->ptr + base;
    int64_t scan_len = p->end - base;
    int complete = ora_scan_record(scan_view, scan_len, &o, &phase, &code);
    if (code != ORA_OK) {                                       /* :332-338 */
        int64_t rec = p->current_line_number / 4 + 1;
        int64_t line = p->current_line_number + 1;
        int64_t pos = p->stream_position + p->head;
        int64_t sn = o.record_end + 1 < scan_len ? o.record_end + 1 : scan_len; /* utils.mojo:436-445 */
        if (sn > 200) sn = 200;
        if (sn < 0) sn = 0;
        ora_fill_parse_error(err, code, rec, line, pos, scan_view, sn);
        return code;
    }
    if (!complete) {                                            /* :339-351 */
        int refill = ora_next_ref_complete(p, base, &o, &phase, &complete);
        base = 0;
        if (refill == ORA_EOF && !complete) {
            ora_fill_plain_error(err, ORA_EOF, "EOF");
            return ORA_EOF;
        } else if (refill != ORA_OK) {
            /* _refill_error_message, parser.mojo:276-309 */
            if (refill == ORA_ID_NO_AT || refill == ORA_SEP_NO_PLUS ||
                refill == ORA_SEQ_QUAL_LEN_MISMATCH) {
                int64_t rec = p->current_line_number / 4 + 1;
                int64_t line = p->current_line_number + 1;
                int64_t pos = p->stream_position + p->head;
                int64_t vl = br_available(p);
                int64_t sn = o.record_end + 1 < vl ? o.record_end + 1 : vl;
                if (sn > 200) sn = 200;
                if (sn < 0) sn = 0;
                ora_fill_parse_error(err, refill, rec, line, pos, p->ptr + p->head, sn);
            } else if (refill == ORA_UNEXPECTED_EOF) {
                char m[96];
                snprintf(m, sizeof m, "Unexpected end of file in FASTQ record at phase %d", phase);
                ora_fill_plain_error(err, refill, m);
            } else if (refill == ORA_BUFFER_EXCEEDED) {
                char m[160];
                snprintf(m, sizeof m,
                         "FASTQ record exceeds buffer capacity (%lld bytes). Enable buffer "
                         "growth or increase buffer_capacity.", (long long)p->len);
                ora_fill_plain_error(err, refill, m);
            } else {
                char m[160];
                snprintf(m, sizeof m,
                         "FASTQ record exceeds maximum buffer capacity (%lld bytes). Enable "
                         "buffer growth or increase max_capacity.",
                         (long long)p->cfg.buffer_max_capacity);
                ora_fill_plain_error(err, refill, m);
            }
            return refill;
        }
        if (!complete) {                                        /* :350-351 `raise Error()` */
            ora_fill_plain_error(err, ORA_EMPTY_ERROR, "");
            return ORA_EMPTY_ERROR;
        }
    }
    /* spans, parser.mojo:353-373; offsets are relative to view()[0] = head */
    int64_t abs0 = p->stream_position + p->head;
    out->header_start = abs0 + o.header_start;
    out->seq_start = abs0 + o.seq_start;
    out->sep_start = abs0 + o.sep_start;
    out->qual_start = abs0 + o.qual_start;
    out->record_end = abs0 + o.record_end;
    int64_t id_s = out->header_start + 1;
    int64_t id_l = o.seq_start - o.header_start - 2;
    if (id_l < 0) id_l = 0; /* the reference would build a negative-length Span here */
    ora_strip(p->data, &id_s, &id_l);
    out->id_start = id_s; out->id_len = id_l;
    out->seq_len = o.sep_start - o.seq_start - 1;
    out->qual_len = o.record_end - o.qual_start;
    /* consume, parser.mojo:375-377 */
    int64_t to_consume = o.record_end + 1;
    int64_t lim = p->end - base;
    if (to_consume > lim) to_consume = lim;
    int64_t av = br_available(p);
    p->head += to_consume < av ? to_consume : av;
    p->current_line_number += 4;
    return ORA_OK;
}

/* next_view, parser.mojo:160-170 */
int ora_next_view(ora_parser* p, ora_view* out, ora_error* err) {
    ora_err_clear(err, ORA_OK);
    int rc = ora_find_and_consume(p, out, err);
    if (rc != ORA_OK) return rc;
    int code = ora_validate_view(p->data, out, &p->cfg);
    if (code != ORA_OK) {
        ora_fill_validation_error(err, code, p->current_line_number / 4, p->data, out);
        return code;
    }
    return ORA_OK;
}

/* next_record, parser.mojo:189-211 */
int ora_next_record(ora_parser* p, ora_view* out, ora_error* err) {
    ora_err_clear(err, ORA_OK);
    if (!ora_has_more(p)) {
        ora_fill_plain_error(err, ORA_EOF, "EOF");
        return ORA_EOF;
    }
    return ora_next_view(p, out, err);
}

/* next_batch, parser.mojo:239-251 */
int ora_next_batch(ora_parser* p, int64_t max_records, ora_view* views, int64_t* n_out,
                   ora_error* err) {
    int64_t limit = max_records ? max_records : 4096; /* DEFAULT_BATCH_SIZE, CONSTS.mojo:31 */
    int64_t k = 0;
    ora_err_clear(err, ORA_OK);
    while (k < limit && ora_has_more(p)) {
        ora_view v;
        int rc = ora_next_view(p, &v, err);
        if (rc == ORA_EOF) { ora_err_clear(err, ORA_OK); break; } /* :248-249 */
        if (rc != ORA_OK) { *n_out = k; return rc; }              /* :250 re-raise */
        views[k++] = v;
    }
    *n_out = k;
    return ORA_OK;
}

/* FastqBatch.add(view), record_batch.mojo:77-87 */
void ora_build_batch(const uint8_t* data, const ora_view* views, int64_t n, uint8_t* id_bytes,
                     uint8_t* seq_bytes, uint8_t* qual_bytes, int64_t* id_ends, int64_t* ends) {
    int64_t io = 0, so = 0, qo = 0;
    for (int64_t i = 0; i < n; ++i) {
        const ora_view* v = &views[i];
        memcpy(qual_bytes + qo, data + v->qual_start, (size_t)v->qual_len); qo += v->qual_len;
        memcpy(seq_bytes + so, data + v->seq_start, (size_t)v->seq_len);   so += v->seq_len;
        memcpy(id_bytes + io, data + v->id_start, (size_t)v->id_len);      io += v->id_len;
        id_ends[i] = io;
        ends[i] = qo; /* quality length, record_batch.mojo:84,87 */
    }
}

/* ------------------------------------------------------------------------ */
/* canonical whole-stream parse                       */
/* ------------------------------------------------------------------------ */

/* Four newlines at/after p.  The AVX2 body mirrors the reference's single
 * forward sweep (utils.mojo:519-531: load W, eq '\n', movemask, ctz loop). */
#if defined(__x86_64__)
__attribute__((target("avx2")))
static int ora_find4_avx2(const uint8_t* p, const uint8_t* end, const uint8_t** nl) {
    int found = 0;
    const __m256i nlv = _mm256_set1_epi8(ORA_NL);
    while (p + 32 <= end) {
        uint32_t m = (uint32_t)_mm256_movemask_epi8(
            _mm256_cmpeq_epi8(_mm256_loadu_si256((const __m256i*)p), nlv));
        while (m) {
            nl[found++] = p + __builtin_ctz(m);
            if (found == 4) return 4;
            m &= m - 1;
        }
        p += 32;
    }
    for (; p < end; ++p)
        if (*p == ORA_NL) { nl[found++] = p; if (found == 4) return 4; }
    return found;
}
#endif

static int ora_find4_memchr(const uint8_t* p, const uint8_t* end, const uint8_t** nl) {
    int found = 0;
    while (found < 4 && p < end) {
        const uint8_t* q = (const uint8_t*)memchr(p, ORA_NL, (size_t)(end - p));
        if (!q) break;
        nl[found++] = q;
        p = q + 1;
    }
    return found;
}

typedef int (*ora_find4_fn)(const uint8_t*, const uint8_t*, const uint8_t**);
static ora_find4_fn ora_pick_find4(void) {
#if defined(__x86_64__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx2")) return ora_find4_avx2;
#endif
    return ora_find4_memchr;
}

int64_t ora_parse_all(const uint8_t* data, size_t n_, const ora_config* cfg, ora_view* views,
                      int64_t cap, int64_t* bases, ora_error* err) {
    ora_find4_fn find4 = ora_pick_find4();
    const int64_t n = (int64_t)n_;
    const uint8_t* end = data + n;
    int64_t pos = 0, lines = 0, nrec = 0, nb = 0;
    ora_err_clear(err, ORA_EOF);
    if (err) strcpy(err->message, "EOF");
    /* A record must fit the BufferedReader: buffer_capacity bytes, or buffer_max_capacity once growth has
     * run its course (parser.mojo:484-503; the streaming model above reaches the same verdict for every
     * record that ends in a newline -- cross-checked in tests/test_oracle_golden.py). */
    const int64_t limit = cfg->buffer_growth_enabled ? cfg->buffer_max_capacity : cfg->buffer_capacity;
    char limit_msg[200];
    if (cfg->buffer_growth_enabled)
        snprintf(limit_msg, sizeof limit_msg, "FASTQ record exceeds maximum buffer capacity (%lld bytes). Enable "
                 "buffer growth or increase max_capacity.", (long long)cfg->buffer_max_capacity);
    else
        snprintf(limit_msg, sizeof limit_msg, "FASTQ record exceeds buffer capacity (%lld bytes). Enable buffer "
                 "growth or increase buffer_capacity.", (long long)cfg->buffer_capacity);
    const int limit_code = cfg->buffer_growth_enabled ? ORA_BUFFER_AT_MAX : ORA_BUFFER_EXCEEDED;
    for (;;) {
        if (pos == n) break;                                     /* parser.mojo:314-317 */
        const uint8_t* nl[4];
        int found = find4(data + pos, end, nl);
        ora_view v;
        v.header_start = pos;
        if (found == 4) {
            v.seq_start = (nl[0] - data) + 1;
            v.sep_start = (nl[1] - data) + 1;
            v.qual_start = (nl[2] - data) + 1;
            v.record_end = nl[3] - data;
            if (limit > 0 && v.record_end + 1 - pos > limit) {   /* the 4th newline is never seen in the buffer */
                ora_fill_plain_error(err, limit_code, limit_msg);
                break;
            }
            int code = ORA_OK;                                   /* utils.mojo:448-462 */
            if (data[pos] != '@') code = ORA_ID_NO_AT;
            else if (data[v.sep_start] != '+') code = ORA_SEP_NO_PLUS;
            else if (v.sep_start - v.seq_start - 1 != v.record_end - v.qual_start)
                code = ORA_SEQ_QUAL_LEN_MISMATCH;
            if (code != ORA_OK) {
                int64_t sn = v.record_end + 1 - pos;
                if (sn > 200) sn = 200;
                ora_fill_parse_error(err, code, lines / 4 + 1, lines + 1, pos, data + pos, sn);
                break;
            }
        } else {
            /* tail rule, SURVEY App. A.2 (parser.mojo:460-492, utils.mojo:292-329) */
            if (pos == 0 && !cfg->buffer_growth_enabled) {
                /* Q2: first record in the buffer incomplete -> BUFFER_EXCEEDED */
                char m[160];
                snprintf(m, sizeof m,
                         "FASTQ record exceeds buffer capacity (%lld bytes). Enable buffer "
                         "growth or increase buffer_capacity.", (long long)cfg->buffer_capacity);
                ora_fill_plain_error(err, ORA_BUFFER_EXCEEDED, m);
                break;
            }
            if (limit > 0 && n - pos > limit) {                  /* a tail that fills the buffer is not at "EOF" */
                ora_fill_plain_error(err, limit_code, limit_msg);
                break;
            }
            if (found < 3) {
                char m[96];
                snprintf(m, sizeof m, "Unexpected end of file in FASTQ record at phase %d", found);
                ora_fill_plain_error(err, ORA_UNEXPECTED_EOF, m);
                break;
            }
            v.seq_start = (nl[0] - data) + 1;
            v.sep_start = (nl[1] - data) + 1;
            v.qual_start = (nl[2] - data) + 1;
            int all_blank = 1;
            for (int64_t i = v.qual_start; i < n; ++i) {
                uint8_t b = data[i];
                if (b != ORA_NL && b != ORA_CR && b != ' ' && b != '\t') { all_blank = 0; break; }
            }
            if (all_blank) { ora_fill_plain_error(err, ORA_EMPTY_ERROR, ""); break; }
            v.record_end = n; /* accepted WITHOUT the structure check (Q1, unpinned) */
        }
        v.id_start = pos + 1;
        v.id_len = v.seq_start - pos - 2;
        if (v.id_len < 0) v.id_len = 0;
        ora_strip(data, &v.id_start, &v.id_len);
        v.seq_len = v.sep_start - v.seq_start - 1;
        v.qual_len = v.record_end - v.qual_start;
        pos = v.record_end + 1 < n ? v.record_end + 1 : n;       /* parser.mojo:375-376 */
        lines += 4;
        int vcode = ora_validate_view(data, &v, cfg);
        if (vcode != ORA_OK) {
            ora_fill_validation_error(err, vcode, lines / 4, data, &v);
            break;
        }
        if (views && nrec < cap) views[nrec] = v;
        nrec++;
        nb += v.seq_len;
    }
    if (bases) *bases = nb;
    return nrec;
}

/* ------------------------------------------------------------------------ */
/* multi-threaded CPU baseline (shard by newline rank)             */
/* ------------------------------------------------------------------------ */

typedef struct {
    const uint8_t* data; int64_t lo, hi, n;
    const ora_config* cfg; int mode; int64_t batch_size;
    int64_t nl_count;        /* pass 1 */
    int64_t start;           /* pass 2: first record start inside the shard */
    int64_t stop;            /* pass 2: parse records with start < stop */
    int64_t recs, bases; int code;
} ora_shard;

static int64_t ora_count_nl(const uint8_t* p, const uint8_t* end) {
    int64_t c = 0;
    while (p < end) {
        const uint8_t* q = (const uint8_t*)memchr(p, ORA_NL, (size_t)(end - p));
        if (!q) break;
        c++; p = q + 1;
    }
    return c;
}

static void* ora_shard_count(void* a) {
    ora_shard* s = (ora_shard*)a;
    s->nl_count = ora_count_nl(s->data + s->lo, s->data + s->hi);
    return NULL;
}

static void* ora_shard_parse(void* a) {
    ora_shard* s = (ora_shard*)a;
    ora_find4_fn find4 = ora_pick_find4();
    const uint8_t* d = s->data;
    const uint8_t* end = d + s->n;
    int64_t pos = s->start, recs = 0, bases = 0;
    int64_t bs = s->batch_size > 0 ? s->batch_size : 4096;
    /* batches mode: FastqBatch(batch_size) reserves 150*batch per array
     * (record_batch.mojo:29-42); List.extend grows geometrically. */
    uint8_t *bq = NULL, *bsq = NULL, *bi = NULL; int64_t *be = NULL, *bie = NULL;
    int64_t capb = 0, capi = 0, qo = 0, io = 0, nb = 0;
    if (s->mode == 1) {
        capb = 150 * bs; capi = 150 * bs;
        bq = (uint8_t*)malloc((size_t)capb); bsq = (uint8_t*)malloc((size_t)capb);
        bi = (uint8_t*)malloc((size_t)capi);
        be = (int64_t*)malloc(sizeof(int64_t) * (size_t)bs);
        bie = (int64_t*)malloc(sizeof(int64_t) * (size_t)bs);
    }
    s->code = ORA_OK;
    while (pos < s->stop) {
        const uint8_t* nl[4];
        if (find4(d + pos, end, nl) < 4) break;
        ora_view v;
        v.header_start = pos;
        v.seq_start = (nl[0] - d) + 1; v.sep_start = (nl[1] - d) + 1;
        v.qual_start = (nl[2] - d) + 1; v.record_end = nl[3] - d;
        v.seq_len = v.sep_start - v.seq_start - 1;
        v.qual_len = v.record_end - v.qual_start;
        if (d[pos] != '@') { s->code = ORA_ID_NO_AT; break; }
        if (d[v.sep_start] != '+') { s->code = ORA_SEP_NO_PLUS; break; }
        if (v.seq_len != v.qual_len) { s->code = ORA_SEQ_QUAL_LEN_MISMATCH; break; }
        v.id_start = pos + 1; v.id_len = v.seq_start - pos - 2;
        ora_strip(d, &v.id_start, &v.id_len);
        int vc = ora_validate_view(d, &v, s->cfg);
        if (vc != ORA_OK) { s->code = vc; break; }
        if (s->mode == 1) {
            while (qo + v.qual_len > capb) {
                capb *= 2; bq = (uint8_t*)realloc(bq, (size_t)capb); bsq = (uint8_t*)realloc(bsq, (size_t)capb);
            }
            while (io + v.id_len > capi) { capi *= 2; bi = (uint8_t*)realloc(bi, (size_t)capi); }
            memcpy(bq + qo, d + v.qual_start, (size_t)v.qual_len);
            memcpy(bsq + qo, d + v.seq_start, (size_t)v.seq_len);
            memcpy(bi + io, d + v.id_start, (size_t)v.id_len);
            qo += v.qual_len; io += v.id_len;
            be[nb] = qo; bie[nb] = io;
            if (++nb == bs) { nb = 0; qo = 0; io = 0; } /* next FastqBatch */
        }
        recs++; bases += v.seq_len;
        pos = v.record_end + 1;
    }
    free(bq); free(bsq); free(bi); free(be); free(bie);
    s->recs = recs; s->bases = bases;
    return NULL;
}

int64_t ora_baseline_mt(const uint8_t* data, size_t n_, const ora_config* cfg, int mode,
                        int64_t batch_size, int threads, int64_t* bases,
                        int32_t* first_error_code) {
    int64_t n = (int64_t)n_;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    ora_shard sh[256];
    pthread_t th[256];
    if (first_error_code) *first_error_code = ORA_OK;
    if (threads == 1) {
        memset(&sh[0], 0, sizeof sh[0]);
        sh[0].data = data; sh[0].n = n; sh[0].cfg = cfg; sh[0].mode = mode;
        sh[0].batch_size = batch_size; sh[0].start = 0; sh[0].stop = n;
        ora_shard_parse(&sh[0]);
        if (bases) *bases = sh[0].bases;
        if (first_error_code) *first_error_code = sh[0].code;
        return sh[0].recs;
    }
    for (int t = 0; t < threads; ++t) {
        memset(&sh[t], 0, sizeof sh[t]);
        sh[t].data = data; sh[t].n = n; sh[t].cfg = cfg; sh[t].mode = mode;
        sh[t].batch_size = batch_size;
        sh[t].lo = n * t / threads; sh[t].hi = n * (t + 1) / threads;
        pthread_create(&th[t], NULL, ora_shard_count, &sh[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    /* exclusive newline rank at each shard start -> the shard's first record
     * starts after its (4 - rank%4)%4-th newline (rank%4 == 0 and the shard
     * starting right after a newline means the shard start IS a record start
     * only if the previous byte is '\n'; otherwise skip to the next boundary). */
    int64_t rank = 0;
    for (int t = 0; t < threads; ++t) {
        int64_t lo = sh[t].lo;
        if (t == 0) sh[t].start = 0;
        else {
            /* newlines to skip so that rank becomes a multiple of 4 */
            int64_t need = (4 - (rank % 4)) % 4;
            int64_t p = lo;
            if (need == 0 && data[lo - 1] != ORA_NL) need = 4;
            while (need > 0 && p < n) {
                const uint8_t* q = (const uint8_t*)memchr(data + p, ORA_NL, (size_t)(n - p));
                if (!q) { p = n; break; }
                p = (q - data) + 1; need--;
            }
            sh[t].start = p;
        }
        rank += sh[t].nl_count;
    }
    for (int t = 0; t < threads; ++t) sh[t].stop = t + 1 < threads ? sh[t + 1].start : n;
    for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, ora_shard_parse, &sh[t]);
    int64_t recs = 0, nb = 0;
    for (int t = 0; t < threads; ++t) {
        pthread_join(th[t], NULL);
        recs += sh[t].recs; nb += sh[t].bases;
        if (first_error_code && *first_error_code == ORA_OK) *first_error_code = sh[t].code;
    }
    if (bases) *bases = nb;
    return recs;
}

/* ------------------------------------------------------------------------ */
/* synthetic generator, utils.mojo:640-678,707-917                            */
/* ------------------------------------------------------------------------ */

static int ora_ndigits(int64_t v) { int d = 1; while (v >= 10) { v /= 10; d++; } return d; }

/* compute_num_reads_for_size, utils.mojo:640-678 */
int64_t ora_compute_num_reads_for_size(int64_t target, int64_t min_length, int64_t max_length) {
    if (target <= 0) return 0;
    int64_t avg = (min_length + max_length) / 2;
    int64_t est = target / (15 + 2 * avg + 4);
    if (est <= 0) return 0;
    int64_t digits = est > 1 ? ora_ndigits(est - 1) : 1;
    return target / (6 + digits + 1 + 2 * avg + 4);
}

static inline int64_t ora_read_len(int64_t i, int64_t mn, int64_t mx) {
    return mn == mx ? mn : mn + ((i * 31 + 7) % (mx - mn + 1)); /* utils.mojo:753-757 */
}

int64_t ora_synth_size(int64_t num_reads, int64_t mn, int64_t mx) {
    if (num_reads <= 0) return 0;
    int64_t digits = num_reads > 1 ? ora_ndigits(num_reads - 1) : 1;
    int64_t total = 0;
    if (mn == mx) return num_reads * (6 + digits + 1 + 2 * mn + 4);
    for (int64_t i = 0; i < num_reads; ++i) total += 6 + digits + 1 + 2 * ora_read_len(i, mn, mx) + 4;
    return total;
}

int64_t ora_synth_generate(int64_t num_reads, int64_t first, int64_t count, int64_t mn,
                           int64_t mx, int64_t min_phred, int64_t max_phred, uint8_t q_lower,
                           uint8_t q_upper, uint8_t q_offset, int gc_slots, uint8_t* out) {
    if (num_reads <= 0) return 0;
    const uint64_t M63 = 0x7FFFFFFFFFFFFFFFULL;
    /* _build_gc_biased_base_lut, utils.mojo:707-733 (gc_bias 0.5 -> 4 slots) */
    if (gc_slots < 0) gc_slots = 4;
    if (gc_slots > 8) gc_slots = 8;
    uint8_t lut[8];
    int k = 0;
    for (int j = 0; j < gc_slots; ++j) lut[k++] = (j % 2 == 0) ? 'G' : 'C';
    for (int j = 0; j < 8 - gc_slots; ++j) lut[k++] = (j % 2 == 0) ? 'A' : 'T';
    int digits = num_reads > 1 ? ora_ndigits(num_reads - 1) : 1;   /* :880-882 */
    int64_t q_start = max_phred, q_range = max_phred - min_phred;  /* :891-893 */
    int64_t noise_amp = q_range / 6 + 1;                           /* :897 */
    uint8_t* w = out;
    int64_t last = first + count;
    if (last > num_reads) last = num_reads;
    for (int64_t i = first; i < last; ++i) {
        int64_t rl = ora_read_len(i, mn, mx);
        /* header, :764-768 */
        memcpy(w, "@read_", 6); w += 6;
        { int64_t v = i; for (int d = digits - 1; d >= 0; --d) { w[d] = (uint8_t)('0' + v % 10); v /= 10; } w += digits; }
        *w++ = '\n';
        /* sequence, :773-784 */
        uint64_t s = ((uint64_t)i * 6364136223846793005ULL + 1442695040888963407ULL) & M63;
        for (int64_t p = 0; p < rl; ++p) {
            s = (s * 6364136223846793005ULL + 1442695040888963407ULL) & M63;
            *w++ = lut[(s >> 33) % 8];
        }
        *w++ = '\n'; *w++ = '+'; *w++ = '\n';
        /* quality, :795-827 */
        uint64_t q = ((uint64_t)i * 2654435761ULL + 1013904223ULL) & M63;
        int64_t lm1 = rl - 1;
        for (int64_t p = 0; p < rl; ++p) {
            int64_t mean = lm1 == 0 ? q_start : q_start - (q_range * p + lm1 / 2) / lm1;
            q = (q * 1664525ULL + 1013904223ULL) & M63;
            int64_t noise_raw = (int64_t)((q >> 17) % (uint64_t)(2 * noise_amp + 1));
            int64_t ph = mean + noise_raw - noise_amp;
            if (ph < min_phred) ph = min_phred; else if (ph > max_phred) ph = max_phred;
            int64_t a = (int64_t)q_offset + ph;
            if (a < q_lower) a = q_lower; else if (a > q_upper) a = q_upper;
            *w++ = (uint8_t)a;
        }
        *w++ = '\n';
    }
    return (int64_t)(w - out);
}

/* ------------------------------------------------------------------------ */
/* SHA-256 (FIPS 180-4) for the generator known-answer tests                  */
/* ------------------------------------------------------------------------ */

static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

#define ROR(x, r) (((x) >> (r)) | ((x) << (32 - (r))))
static void sha256_block(uint32_t h[8], const uint8_t* b) {
    uint32_t w[64];
    for (int i = 0; i < 16; ++i)
        w[i] = ((uint32_t)b[4 * i] << 24) | ((uint32_t)b[4 * i + 1] << 16) |
               ((uint32_t)b[4 * i + 2] << 8) | b[4 * i + 3];
    for (int i = 16; i < 64; ++i) {
        uint32_t s0 = ROR(w[i - 15], 7) ^ ROR(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = ROR(w[i - 2], 17) ^ ROR(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; ++i) {
        uint32_t S1 = ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + K256[i] + w[i];
        uint32_t S0 = ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22);
        uint32_t mj = (a & bb) ^ (a & c) ^ (bb & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
    }
    h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

void ora_sha256(const uint8_t* data, size_t n, uint8_t out[32]) {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                     0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    size_t i = 0;
    for (; i + 64 <= n; i += 64) sha256_block(h, data + i);
    uint8_t tail[128];
    size_t r = n - i;
    memcpy(tail, data + i, r);
    tail[r++] = 0x80;
    size_t padlen = r <= 56 ? 64 : 128;
    memset(tail + r, 0, padlen - r);
    uint64_t bits = (uint64_t)n * 8;
    for (int j = 0; j < 8; ++j) tail[padlen - 1 - j] = (uint8_t)(bits >> (8 * j));
    sha256_block(h, tail);
    if (padlen == 128) sha256_block(h, tail + 64);
    for (int j = 0; j < 8; ++j) {
        out[4 * j] = (uint8_t)(h[j] >> 24); out[4 * j + 1] = (uint8_t)(h[j] >> 16);
        out[4 * j + 2] = (uint8_t)(h[j] >> 8); out[4 * j + 3] = (uint8_t)h[j];
    }
}

/* ------------------------------------------------------------------------ */
/* FASTA: blazeseq/fasta/parser.mojo:60-200 over LineIterator   */
/* (io/buffered.mojo:600-638: lines without '\n', a trailing '\r' trimmed,   */
/* the last line may lack its newline; line numbers count lines read; the    */
/* file position is the stream offset at the start of the last next_line).   */
/* ------------------------------------------------------------------------ */

typedef struct { const uint8_t* d; int64_t n, pos, line_no, file_pos; } ora_lines;

/* LineIterator.next_line: 1 and [*s, *e) on a line, 0 at EOF */
static int ora_next_line(ora_lines* L, int64_t* s, int64_t* e) {
    L->file_pos = L->pos;                                   /* buffered.mojo:606 */
    if (L->pos >= L->n) return 0;                           /* :609-616 EOFError */
    const uint8_t* q = (const uint8_t*)memchr(L->d + L->pos, ORA_NL, (size_t)(L->n - L->pos));
    int64_t end = q ? q - L->d : L->n;
    *s = L->pos;
    *e = end;
    if (*e > *s && L->d[*e - 1] == ORA_CR) --*e;            /* _trim_trailing_cr */
    L->pos = q ? end + 1 : L->n;
    L->line_no++;
    return 1;
}

static void ora_strip_range(const uint8_t* d, int64_t* s, int64_t* e) {   /* _strip_spaces, utils.mojo:221-242 */
    int64_t len = *e - *s;
    ora_strip(d, s, &len);
    *e = *s + len;
}

/* FastaParser.next_record in a loop (fasta/parser.mojo:123-172, _read_header_line :182-203).
 * ids: (start, len) pairs per record; seq_off: n+1 cumulative offsets into seq (sequence r = seq[seq_off[r], seq_off[r+1]));
 * returns the records delivered before the stop; *err = the stop (ORA_EOF on a clean end). */
int64_t ora_fasta_parse(const uint8_t* data, size_t n_, int check_ascii, int64_t* ids, int64_t* seq_off, uint8_t* seq,
                        int64_t cap_records, ora_error* err) {
    ora_lines L = {data, (int64_t)n_, 0, 0, 0};
    int64_t nrec = 0, nseq = 0;
    int have_pending = 0;
    int64_t pend_s = 0, pend_e = 0;
    ora_err_clear(err, ORA_EOF);
    if (err) strcpy(err->message, "EOF");
    if (seq_off) seq_off[0] = 0;
    for (;;) {
        int64_t id_s, id_e, s, e;
        /* has_more (:103-105) is `pending or lines.has_more()`; _read_header_line raises EOF itself when only blank lines remain */
        if (have_pending) { id_s = pend_s; id_e = pend_e; have_pending = 0; }
        else {
            int got = 0;
            while (ora_next_line(&L, &s, &e)) {                       /* :189-203 */
                ora_strip_range(data, &s, &e);
                if (e == s) continue;
                if (data[s] != '>') {
                    ora_err_clear(err, ORA_OTHER);
                    err->record_number = nrec; err->line_number = L.line_no; err->file_position = L.file_pos;
                    ora_sb b = {err->message, sizeof err->message, 0};
                    sb_str(&b, "FASTA: sequence id line does not start with '>'");
                    if (nrec > 0) { sb_str(&b, "\n  Record number: "); sb_i64(&b, nrec); }
                    if (L.line_no > 0) { sb_str(&b, "\n  Line number: "); sb_i64(&b, L.line_no); }
                    if (L.file_pos > 0) { sb_str(&b, "\n  File position: "); sb_i64(&b, L.file_pos); }
                    return nrec;
                }
                id_s = s + 1; id_e = e;
                ora_strip_range(data, &id_s, &id_e);
                got = 1;
                break;
            }
            if (!got) return nrec;                                   /* EOFError */
        }
        const int64_t seq_start_line = L.line_no + 1;                /* :134 */
        const int64_t seq_begin = nseq;
        while (ora_next_line(&L, &s, &e)) {                           /* :136-149 */
            ora_strip_range(data, &s, &e);
            if (e > s && data[s] == '>') {
                pend_s = s + 1; pend_e = e;
                ora_strip_range(data, &pend_s, &pend_e);
                have_pending = 1;
                break;
            }
            if (seq) memcpy(seq + nseq, data + s, (size_t)(e - s));
            nseq += e - s;
        }
        if (nseq == seq_begin) {                                     /* :152-160 */
            ora_err_clear(err, ORA_OTHER);
            err->record_number = nrec + 1; err->line_number = seq_start_line; err->file_position = L.file_pos;
            ora_sb b = {err->message, sizeof err->message, 0};
            sb_str(&b, "FASTA record has empty sequence");
            sb_str(&b, "\n  Record number: "); sb_i64(&b, nrec + 1);
            sb_str(&b, "\n  Line number: "); sb_i64(&b, seq_start_line);
            if (L.file_pos > 0) { sb_str(&b, "\n  File position: "); sb_i64(&b, L.file_pos); }
            return nrec;
        }
        if (check_ascii) {                                           /* :162-163, Validator :40-58 */
            int bad = 0;
            for (int64_t i = id_s; i < id_e && !bad; ++i) bad = data[i] & 0x80;
            for (int64_t i = seq_begin; i < nseq && !bad; ++i) bad = seq ? (seq[i] & 0x80) : 0;
            if (bad) {
                ora_err_clear(err, ORA_ASCII_INVALID);
                err->record_number = nrec;
                ora_sb b = {err->message, sizeof err->message, 0};
                sb_str(&b, ora_code_message(ORA_ASCII_INVALID));
                if (nrec > 0) { sb_str(&b, "\n  Record number: "); sb_i64(&b, nrec); }
                return nrec;
            }
        }
        if (nrec < cap_records) {
            if (ids) { ids[2 * nrec] = id_s; ids[2 * nrec + 1] = id_e - id_s; }
            if (seq_off) seq_off[nrec + 1] = nseq;
        }
        nrec++;
    }
}
