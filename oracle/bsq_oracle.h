/*
 * bsq_oracle.h -- CPU restatement of BlazeSeq's FASTQ hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path (blazeseq_b200/) never calls into it.
 *
 * The reference (MoSafi2/BlazeSeq @ 66ddbd1) is Mojo and cannot be built in
 * this image (no mojo/pixi toolchain), so there is no oracle/_ref build.  The
 * restatement is pinned against the reference's own literal test streams,
 * its 70-file corpus expectations and its Python-binding test ids
 * (tests/test_oracle_*.py, tests/golden/).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference repo root).
 */
#ifndef BSQ_ORACLE_H
#define BSQ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* FastxErrorCode -- blazeseq/errors.mojo:43-56 */
enum {
    ORA_OK = 0,
    ORA_ID_NO_AT = 1,
    ORA_SEP_NO_PLUS = 2,
    ORA_SEQ_QUAL_LEN_MISMATCH = 3,
    ORA_ASCII_INVALID = 4,
    ORA_QUALITY_OUT_OF_RANGE = 5,
    ORA_EOF = 6,
    ORA_UNEXPECTED_EOF = 7,
    ORA_BUFFER_EXCEEDED = 8,
    ORA_BUFFER_AT_MAX = 9,
    ORA_OTHER = 10,
    ORA_EMPTY_ERROR = 11 /* `raise Error()` with empty text, parser.mojo:350-351 */
};

/* ParserConfig (parser.mojo:33-74) + the resolved QualitySchema
 * (quality_schema.mojo:26-31) + oracle-only knobs. */
typedef struct ora_config {
    int64_t buffer_capacity;     /* DEFAULT_CAPACITY = 256 KiB (CONSTS.mojo:26) */
    int64_t buffer_max_capacity; /* MAX_CAPACITY = 2^30 (CONSTS.mojo:28) */
    int32_t buffer_growth_enabled;
    int32_t check_ascii;
    int32_t check_quality;
    uint8_t q_lower, q_upper, q_offset, _pad;
    /* 0: documented intent, byte valid iff LOWER <= b <= UPPER everywhere.
     * W (16/32/64): emulate record.mojo:90-102, where the first floor(n/W)*W
     * quality bytes use `(b-LOWER) >= span` and only the tail uses `>`. */
    int32_t compat_simd_width;
    /* max bytes a single Reader.read_to_buffer call returns; 0 = unlimited.
     * Models short reads. */
    int64_t reader_max_read;
} ora_config;

/* One parsed record.  All offsets are absolute positions in the input
 * stream (the reference's buffer-relative RecordOffsets, utils.mojo:37-93,
 * rebased by BufferedReader.stream_position, buffered.mojo:176-182). */
typedef struct ora_view {
    int64_t header_start; /* '@' byte */
    int64_t seq_start;
    int64_t sep_start;
    int64_t qual_start;
    int64_t record_end;   /* one past the last quality byte */
    int64_t id_start;     /* after _strip_spaces (utils.mojo:221-242) */
    int64_t id_len;
    int64_t seq_len;
    int64_t qual_len;
} ora_view;

typedef struct ora_error {
    int32_t code;           /* ORA_* */
    int64_t record_number;  /* 0 = not printed */
    int64_t line_number;    /* 0 = not printed */
    int64_t file_position;  /* 0 = not printed */
    char    message[1024];  /* full text as String(Error) would be in the ref */
} ora_error;

typedef struct ora_parser ora_parser;

void ora_default_config(ora_config* cfg);
/* _parse_schema (utils.mojo:612-637).  Returns 0 if known, 1 if the name is
 * unknown (reference prints a warning and falls back to generic). */
int ora_parse_schema(const char* name, uint8_t* lower, uint8_t* upper, uint8_t* offset);

/* Streaming model: BufferedReader (buffered.mojo:115-327) over a MemoryReader
 * (readers.mojo:140-223) driving FastqParser (parser.mojo:77-625).  `data`
 * must outlive the parser. */
ora_parser* ora_open(const uint8_t* data, size_t n, const ora_config* cfg);
void ora_close(ora_parser* p);
int  ora_has_more(const ora_parser* p);                 /* parser.mojo:156-157 */
/* next_view (parser.mojo:160-170).  Returns ORA_OK and fills *out, or an error
 * code (ORA_EOF for end of input) and fills *err. */
int  ora_next_view(ora_parser* p, ora_view* out, ora_error* err);
/* next_record (parser.mojo:189-211): same record content; differs from
 * next_view only by the leading has_more() check. */
int  ora_next_record(ora_parser* p, ora_view* out, ora_error* err);
/* next_batch (parser.mojo:239-251).  Fills up to max_records views into
 * `views` (caller array).  Returns ORA_OK (n_out may be < max at EOF) or the
 * error code that made the reference re-raise (the partial batch is lost in
 * the reference; n_out still reports how many were collected). */
int  ora_next_batch(ora_parser* p, int64_t max_records, ora_view* views,
                    int64_t* n_out, ora_error* err);
/* Pointer to the input byte at absolute stream offset (convenience). */
const uint8_t* ora_data(const ora_parser* p);

/* FastqBatch SoA (record_batch.mojo:22-27,77-87) built from `n` views.
 * Caller provides buffers sized from the views' lengths.  ends/id_ends are
 * inclusive cumulative Int64 restarting at 0 for this batch. */
void ora_build_batch(const uint8_t* data, const ora_view* views, int64_t n,
                     uint8_t* id_bytes, uint8_t* seq_bytes, uint8_t* qual_bytes,
                     int64_t* id_ends, int64_t* ends);

/* Whole-stream canonical parse (SURVEY App. A.1/A.2 = the streaming model with
 * buffer_capacity > n, derived and cross-checked in tests).  Fast path used
 * for the CPU baseline: memchr newline search, no buffer model.
 * Writes up to `cap` views (views may be NULL to only count).  Returns number
 * of records parsed before the stop; *err holds the stop reason (ORA_EOF on a
 * clean end).  `bases` accumulates sequence lengths. */
int64_t ora_parse_all(const uint8_t* data, size_t n, const ora_config* cfg,
                      ora_view* views, int64_t cap, int64_t* bases, ora_error* err);

/* Multi-threaded CPU baseline (newline-rank sharding): counts
 * records/bases (mode 0 = views) or additionally packs FastqBatch SoA chunks
 * of `batch_size` into per-thread scratch (mode 1 = batches).  Returns the
 * record count; input must be structurally valid with a trailing newline. */
int64_t ora_baseline_mt(const uint8_t* data, size_t n, const ora_config* cfg,
                        int mode, int64_t batch_size, int threads, int64_t* bases,
                        int32_t* first_error_code);

/* Synthetic generator (utils.mojo:640-678,707-917). */
int64_t ora_compute_num_reads_for_size(int64_t target_size_bytes, int64_t min_length,
                                       int64_t max_length);
/* Size in bytes of generate_synthetic_fastq_buffer(...) output. */
int64_t ora_synth_size(int64_t num_reads, int64_t min_length, int64_t max_length);
/* Writes records [first, first+count) of the num_reads-record stream into out
 * (which must hold them); returns bytes written.  gc_bias is fixed at the
 * default 0.5 unless gc_slots (0..8) is given as >= 0. */
int64_t ora_synth_generate(int64_t num_reads, int64_t first, int64_t count,
                           int64_t min_length, int64_t max_length, int64_t min_phred,
                           int64_t max_phred, uint8_t q_lower, uint8_t q_upper,
                           uint8_t q_offset, int gc_slots, uint8_t* out);

/* FASTA: FastaParser.next_record in a loop (blazeseq/fasta/parser.mojo:60-200).  ids: (start, len) pairs of the
 * stripped id in `data`; seq_off: n+1 cumulative offsets into `seq` (all line breaks and surrounding blanks
 * removed); `seq` must hold n bytes.  Returns the records delivered before the stop; *err = the stop reason
 * (ORA_EOF on a clean end; ORA_OTHER with the reference's text for the two FASTA parse errors). */
int64_t ora_fasta_parse(const uint8_t* data, size_t n, int check_ascii, int64_t* ids, int64_t* seq_off, uint8_t* seq,
                        int64_t cap_records, ora_error* err);

/* SHA-256 helper for generator known-answer tests. */
void ora_sha256(const uint8_t* data, size_t n, uint8_t out[32]);

#ifdef __cplusplus
}
#endif
#endif
