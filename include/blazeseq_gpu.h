/*
 * blazeseq_gpu.h -- C ABI of the B200 FASTQ record tokenizer / validator / SoA packer.
 *
 * Drop-in boundary for the hot path of MoSafi2/BlazeSeq (reference paths are relative to its
 * repository root).  The reference has no FFI for this path -- it is a generic Mojo struct
 * consumed in-process -- so every entry point names the reference symbol it stands in for.  Mojo
 * reaches C through OwnedDLHandle(...).get_function (blazeseq/io/readers.mojo:242-280), hence:
 * plain C, POD structs, no callbacks, no exceptions; every function returns a bsq_status.
 *
 * Model.  A parser parses one contiguous region of a FASTQ byte stream per call ("pass").  The
 * pass finds every complete 4-line record, checks structure ('@', '+', equal seq/qual length),
 * optionally validates ASCII and the quality range, and leaves on the device:
 *   - the offsets table  (views():   blazeseq/fastq/parser.mojo:160-170,311-379)
 *   - the FastqBatch SoA (batches(): blazeseq/fastq/parser.mojo:239-251,
 *                                    blazeseq/fastq/record_batch.mojo:19-87,210-220)
 * It reports how many bytes it consumed (through the last complete record), exactly like
 * BufferedReader.consume (blazeseq/io/buffered.mojo:156-164); the caller re-presents the
 * unconsumed tail in front of the next region (the reference's _compact_from, :239-260).
 * All compute runs in CUDA kernels on the parser's device; there is no CPU fallback.
 */
#ifndef BLAZESEQ_GPU_H
#define BLAZESEQ_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSQ_ABI_VERSION 3

/* >= 0: FastxErrorCode, identical to blazeseq/errors.mojo:43-56.  < 0: library failures. */
typedef int32_t bsq_status;
enum {
    BSQ_OK = 0,
    BSQ_ID_NO_AT = 1,
    BSQ_SEP_NO_PLUS = 2,
    BSQ_SEQ_QUAL_LEN_MISMATCH = 3,
    BSQ_ASCII_INVALID = 4,
    BSQ_QUALITY_OUT_OF_RANGE = 5,
    BSQ_EOF = 6,
    BSQ_UNEXPECTED_EOF = 7,
    BSQ_BUFFER_EXCEEDED = 8,
    BSQ_BUFFER_AT_MAX = 9,
    BSQ_OTHER = 10,
    BSQ_EMPTY_ERROR = 11,      /* `raise Error()` with empty text, parser.mojo:350-351 */
    BSQ_E_CUDA = -1,           /* a CUDA runtime call failed; see bsq_last_error_text */
    BSQ_E_ARG = -2,            /* bad argument */
    BSQ_E_NO_DEVICE = -3,      /* no usable CUDA device: this library never parses on the CPU */
    BSQ_E_NOMEM = -4,
    BSQ_E_STATE = -5,          /* call sequence error (e.g. results requested before a pass) */
    BSQ_E_IO = -6              /* bsq_stream_*: reading or inflating the file failed */
};

/* what a pass materialises */
#define BSQ_WANT_OFFSETS 1u    /* views(): line-end table + stripped id spans */
#define BSQ_WANT_BATCHES 2u    /* batches(): FastqBatch SoA */
#define BSQ_WANT_WHOLE_BATCHES 4u /* with BSQ_WANT_BATCHES on a region that does not end the stream (is_last = 0): the
                                  records of a trailing partial batch stay unconsumed (bytes_consumed ends at the last whole
                                  batch), so that batches(m) cut region by region are the batches of the whole stream */

/* ParserConfig (parser.mojo:33-74) + the resolved QualitySchema (quality_schema.mojo:9-31).
 * check_ascii / check_quality select the kernel instantiation (the reference's comptime
 * toggles). */
typedef struct bsq_config {
    int32_t device_id;             /* CUDA ordinal */
    int32_t check_ascii;           /* ParserConfig.check_ascii */
    int32_t check_quality;         /* ParserConfig.check_quality */
    uint8_t q_lower, q_upper, q_offset, _pad0; /* QualitySchema.LOWER/UPPER/OFFSET */
    int64_t buffer_capacity;       /* ParserConfig.buffer_capacity: a record longer than this many bytes ends the
                                      parse with BSQ_BUFFER_EXCEEDED when growth is off (parser.mojo:484-492) */
    int64_t buffer_max_capacity;   /* ParserConfig.buffer_max_capacity: ... than this, BSQ_BUFFER_AT_MAX with
                                      growth on (parser.mojo:493-503) */
    int32_t buffer_growth_enabled; /* ParserConfig.buffer_growth_enabled */
    int32_t batch_size;            /* FastqParser._batch_size, DEFAULT_BATCH_SIZE = 4096 */
    int64_t h2d_chunk_bytes;       /* staging chunk for bsq_parse_host (default 64 MiB) */
    int32_t force_id_slow_path;    /* tests: always take the id strip pipeline */
    int32_t inflate_threads;       /* bsq_stream_*: host threads that decode an ordinary gzip stream (1 = zlib's gzread) /
                                      inflate BGZF members with host_inflate / read slices of a plain file (0 = all cores,
                                      at most 8 for plain reads); the parallelism argument of RapidgzipReader,
                                      readers.mojo:380-443 */
    int32_t compat_q5_width;       /* 0: quality bytes are valid iff LOWER <= b <= UPPER (the documented intent).
                                      W > 0: reproduce Validator._validate_quality_range as written
                                      (fastq/record.mojo:90-102) for a host whose SIMD width is W bytes: the first
                                      floor(n / W) * W quality bytes of a record also fail when b == UPPER */
    int32_t host_inflate;          /* bsq_stream_*: 0 = the members of a BGZF file cross PCIe compressed and are inflated on the
                                      device (k_inflate_members); 1 = by inflate_threads host threads (zlib) */
} bsq_config;

/* The first error of a pass, with the context the reference prints
 * (errors.mojo:178-192,223-234; parser.mojo:332-338,163-169). */
typedef struct bsq_error {
    int32_t code;                  /* BSQ_OK / BSQ_EOF when the pass ended cleanly */
    int32_t _pad;
    int64_t record_number;         /* 1-based; 0 = not reported */
    int64_t line_number;           /* 1-based; 0 = not reported */
    int64_t file_position;         /* stream offset; 0 = not reported */
    char message[1024];            /* String(e) of the reference, NUL-terminated */
} bsq_error;

/* Result of one pass. */
typedef struct bsq_pass_result {
    int64_t n_records;             /* records before the stop (EOF or first error) */
    int64_t n_bases;               /* sum of sequence lengths over those records */
    int64_t bytes_consumed;        /* through the last complete record (== n when is_last) */
    int64_t n_newlines;            /* '\n' count of the region */
    int64_t n_batches;             /* ceil(n_records / batch_size) when BSQ_WANT_BATCHES */
    int32_t n_windows;             /* offsets tables are per window (<= 2 GiB of input each) */
    int32_t id_slow_path;          /* 1 if some id needed _strip_spaces (utils.mojo:221-242) */
    bsq_error stop;                /* why the pass stopped: BSQ_EOF, BSQ_OK (more input needed)
                                      or the first error */
} bsq_pass_result;

/* views(): one window of the offsets table.  Record i of the window (global record
 * first_record + i) has, relative to the window base (stream offset stream_base):
 *   header_start = line_ends[4i]+1   seq_start = line_ends[4i+1]+1   sep_start = line_ends[4i+2]+1
 *   qual_start   = line_ends[4i+3]+1 record_end = line_ends[4i+4]
 * (RecordOffsets, utils.mojo:37-93: u32 arithmetic, the leading sentinel makes i = 0 regular)
 * and the id after _strip_spaces is [id_spans[2i], id_spans[2i] + id_spans[2i+1]).
 * Pointers are DEVICE pointers, valid until the next pass or bsq_destroy. */
typedef struct bsq_offsets_view {
    int64_t stream_base;
    int64_t first_record;
    int64_t n_records;
    const uint32_t* line_ends;     /* 4*n_records + 1 entries */
    const uint32_t* id_spans;      /* 2*n_records entries */
    const uint8_t* window_bytes;   /* device address of stream_base (NULL for host passes whose
                                      input was streamed through the staging ring) */
} bsq_offsets_view;

/* batches(): one FastqBatch / DeviceFastqBatch (record_batch.mojo:22-27,210-220).  ends and
 * id_ends are inclusive cumulative Int64 restarting at 0 for the batch (:82-87).  DEVICE
 * pointers, valid until the next pass or bsq_destroy. */
typedef struct bsq_batch_view {
    int64_t num_records;
    int64_t seq_len;               /* == ends[num_records-1] */
    int64_t total_id_bytes;
    uint8_t quality_offset;        /* always 33 on this path (parser.mojo:243) */
    uint8_t _pad[7];
    int64_t sequence_bytes;        /* bytes in sequence_buffer: == seq_len, except when the batch ends
                                      with a stream's last record that has no trailing newline -- it is
                                      accepted without the length check, and `ends` counts QUALITY
                                      bytes (record_batch.mojo:83-87) */
    const uint8_t* sequence_buffer;
    const uint8_t* qual_buffer;
    const uint8_t* id_buffer;
    const int64_t* ends;
    const int64_t* id_ends;
} bsq_batch_view;

/* ---- configuration helpers (host only) --------------------------------------------------- */

/* ParserConfig() defaults (parser.mojo:60-74) with generic_schema (quality_schema.mojo:26). */
void bsq_default_config(bsq_config* cfg);
/* _parse_schema (utils.mojo:612-637).  Returns 0, or 1 for an unknown name (falls back to
 * generic like the reference, which prints a warning). */
int32_t bsq_parse_schema(const char* name, uint8_t* lower, uint8_t* upper, uint8_t* offset);
uint32_t bsq_abi_version(void);

/* ---- parser lifetime ------------------------------------------------------------------------ */

typedef struct bsq_parser bsq_parser;

/* FastqParser.__init__ (parser.mojo:89-145).  One parser = one device, one stream pair. */
bsq_status bsq_create(const bsq_config* cfg, bsq_parser** out);
void bsq_destroy(bsq_parser* p);
/* Text of the last library failure (status < 0) on this parser ("" if none). */
const char* bsq_last_error_text(const bsq_parser* p);
/* FastqParser.batches(max_records) / next_batch(max_records) (parser.mojo:239-274): the batch
 * size used by the passes that follow. */
bsq_status bsq_set_batch_size(bsq_parser* p, int32_t batch_size);

/* ---- passes ------------------------------------------------------------------------------------ */

/* Parse n bytes that are already in DEVICE memory.  stream_offset = position of dev_bytes[0] in
 * the stream (only used for error context); first_record = number of records before it
 * (record/line numbers in errors, batch cutting).  is_last: the region ends the stream, so the
 * tail rule applies (parser.mojo:460-492, utils.mojo:292-329).  `want` = BSQ_WANT_* bits.
 * Synchronous.  Stands in for the _find_and_consume_ref_record loop (parser.mojo:311-379) over a
 * MemoryReader-style buffer. */
bsq_status bsq_parse_device(bsq_parser* p, const uint8_t* dev_bytes, uint64_t n,
                            int64_t stream_offset, int64_t first_record, int32_t is_last,
                            uint32_t want, bsq_pass_result* out);

/* Same, for n bytes in HOST memory (pinned or pageable).  The bytes are copied to the device in
 * h2d_chunk_bytes pieces on a side stream, double-buffered through pinned staging when the source
 * is pageable, overlapped with the first scan pass.  Replaces Reader.read_to_buffer +
 * BufferedReader (readers.mojo:51-79, buffered.mojo:115-327) + the parse loop. */
bsq_status bsq_parse_host(bsq_parser* p, const uint8_t* host_bytes, uint64_t n,
                          int64_t stream_offset, int64_t first_record, int32_t is_last,
                          uint32_t want, bsq_pass_result* out);

/* ---- streaming from a file ----------------------------------------------------------------------
 * FileReader / GZFile / RapidgzipReader (io/readers.mojo:86-137,283-377,380-443) + BufferedReader
 * (io/buffered.mojo:115-327) + the parse loop, as a pipeline: a reader thread reads (or inflates,
 * zlib) the next region into pinned memory while the GPU parses the current one; the H2D copies of
 * a region run on the copy stream, overlapped with its first scan pass.  The unconsumed tail of a
 * region (the partial record; with BSQ_WANT_BATCHES also the records of a trailing partial batch) is
 * carried in front of the next region, like BufferedReader._compact_from (:239-260). */
typedef struct bsq_stream bsq_stream;
#define BSQ_SOURCE_PLAIN 0
#define BSQ_SOURCE_GZIP 1          /* gzip: BGZF members are inflated on the device, one warp per member (or block-parallel by
                                      cfg.inflate_threads host threads with cfg.host_inflate); any other gzip stream is
                                      decoded by cfg.inflate_threads host threads with the two-stage speculative decoder
                                      of bsq_gzip_* below (one zlib thread when inflate_threads == 1) */
#define BSQ_SOURCE_AUTO 2          /* by suffix: .gz / .bgz -> gzip (python/blazeseq_parser.mojo:100-114) */

typedef struct bsq_stream_stats {
    uint64_t bytes_read;           /* decompressed bytes delivered by the reader thread */
    uint64_t regions;
    double reader_busy_s;          /* reader thread: time spent reading / inflating */
    double parse_s;                /* caller: time inside the GPU passes (incl. H2D) */
    double wait_reader_s;          /* caller: time blocked waiting for the reader thread */
    double h2d_s;                  /* device-inflated BGZF: device time of the compressed bytes' H2D copies ... */
    double inflate_s;              /* ... and of k_inflate_members + k_crc32_members (CUDA events) */
    uint64_t compressed_bytes;     /* ... and the compressed bytes that crossed PCIe */
    double launch_s;               /* ... caller: time spent enqueueing copies / kernels (incl. buffer growth) */
    double wait_inflate_s;         /* ... caller: time blocked until a region was inflated */
} bsq_stream_stats;

bsq_status bsq_stream_open(bsq_parser* p, const char* path, int32_t source_kind, uint64_t region_bytes,
                           bsq_stream** out);
/* Parses the next region.  out->stop.code == BSQ_OK: more regions follow; BSQ_EOF: clean end;
 * anything else: the first error (no further regions).  Result views (bsq_get_offsets, bsq_get_batch,
 * ...) refer to this region until the next call.  first_record / stream offsets are global. */
bsq_status bsq_stream_next(bsq_stream* s, uint32_t want, bsq_pass_result* out);
/* Host bytes of the region just parsed (offset tables index these); *stream_offset = position of
 * byte 0 in the file's (decompressed) stream; *first_record = records before it. */
const uint8_t* bsq_stream_region(const bsq_stream* s, uint64_t* n, int64_t* stream_offset, int64_t* first_record);
/* The same numbers without the bytes: a region that was inflated on the device (BGZF) is copied to the host only by
 * bsq_stream_region; callers that consume device batches never pay for that copy. */
void bsq_stream_region_info(const bsq_stream* s, uint64_t* n, int64_t* stream_offset, int64_t* first_record);
bsq_status bsq_stream_get_stats(const bsq_stream* s, bsq_stream_stats* out);
void bsq_stream_close(bsq_stream* s);

/* ---- RapidgzipReader (blazeseq/io/readers.mojo:380-443) ------------------------------------ */

/* Parallel decoder for ordinary gzip files on host threads -- `RapidgzipReader(path, parallelism)`: the compressed file
 * is cut into chunks, every chunk is inflated speculatively from the first deflate block found in it (references into
 * the unknown 32 KiB window are kept as markers), chunks are stitched in order, markers resolved and every member's
 * CRC-32 / ISIZE verified.  Host-only: it needs no parser and no device.  parallelism 0 = all cores.
 * bsq_gzip_read is `Reader.read_to_buffer` (readers.mojo:421-443) / gzread: *got = bytes written, 0 at the end. */
typedef struct bsq_gzip bsq_gzip;
/* chunk_bytes: compressed bytes per speculative chunk (0 = 2 MiB, at least 64 KiB) */
bsq_status bsq_gzip_open(const char* path, int32_t parallelism, uint64_t chunk_bytes, bsq_gzip** out);
bsq_status bsq_gzip_read(bsq_gzip* g, uint8_t* dst, uint64_t n, uint64_t* got);
const char* bsq_gzip_error(const bsq_gzip* g);
void bsq_gzip_close(bsq_gzip* g);

/* ---- results of the last pass -------------------------------------------------------------- */

/* A consumer of the device-resident SoA of the last batches() pass -- the hand-off BlazeSeq's GPU story is
 * about (DeviceFastqBatch, fastq/record_batch.mojo:210-220, consumed by a kernel in
 * examples/nw_gpu/kernels.mojo:21-89): per-record sum of Phred scores (quality byte - cfg.q_offset) for the
 * arena records [first_record, first_record + count), computed on the device from the quality arena and the
 * per-batch `ends`.  The int32 results go to out_device (device pointer, may be NULL: a library buffer is
 * used) and, if out_host is not NULL, are copied there.  The parsed bytes never visit the host. */
bsq_status bsq_quality_sums(bsq_parser* p, int64_t first_record, int64_t count, int32_t* out_device, int32_t* out_host);

/* FastqRecord.write / byte_len (fastq/record.mojo:384-402) over the device SoA of the last batches() pass: records
 * [first_record, first_record + count) serialised back to four-line FASTQ text ('@' id '\n' sequence '\n' '+' '\n' quality '\n')
 * by one warp per record -- the device side of records() and of the writer round trip (tests/fastq/test_fastq_integration.mojo).
 * out_device (capacity bytes) and / or out_host receive the text; both NULL: only *bytes_written (the size to allocate) and the
 * offsets are computed.  offsets_host (optional, count + 1 entries): byte offset of every record in the text. */
bsq_status bsq_write_records(bsq_parser* p, int64_t first_record, int64_t count, uint8_t* out_device, uint64_t capacity,
                             uint8_t* out_host, uint64_t* offsets_host, uint64_t* bytes_written);

bsq_status bsq_get_offsets(const bsq_parser* p, int32_t window, bsq_offsets_view* out);
/* FastqParser.next_batch(max_records) restricted to the pass: batch b holds records
 * [b*batch_size, min((b+1)*batch_size, n_records)). */
bsq_status bsq_get_batch(const bsq_parser* p, int64_t batch_index, bsq_batch_view* out);
/* Whole-pass SoA (all batches back to back; ends/id_ends are the per-batch rebased values). */
bsq_status bsq_get_soa(const bsq_parser* p, bsq_batch_view* out);
/* DeviceFastqBatch.copy_to_host (record_batch.mojo:222-241): copies one batch into caller
 * arrays sized from bsq_batch_view (sequence_bytes, seq_len, total_id_bytes, num_records,
 * num_records). */
bsq_status bsq_batch_to_host(bsq_parser* p, int64_t batch_index, uint8_t* seq, uint8_t* qual,
                             uint8_t* id, int64_t* ends, int64_t* id_ends);
/* The whole pass as ONE host FastqBatch-shaped SoA (all batches back to back, ends / id_ends rebased per batch
 * as in bsq_get_soa): what a caller that wants the reference's host FastqBatch product, not the
 * DeviceFastqBatch, pays for.  Arrays sized from bsq_get_soa (sequence_bytes, seq_len, total_id_bytes,
 * num_records, num_records); pinned destinations make the copies asynchronous.  NULL skips an array. */
bsq_status bsq_soa_to_host(bsq_parser* p, uint8_t* seq, uint8_t* qual, uint8_t* id, int64_t* ends,
                           int64_t* id_ends);
/* Copies one window's offsets table to the host (line_ends: 4n+1, id_spans: 2n entries). */
bsq_status bsq_offsets_to_host(bsq_parser* p, int32_t window, uint32_t* line_ends,
                               uint32_t* id_spans);
/* Device address of the input of the last pass (the staged copy for host passes). */
const uint8_t* bsq_pass_device_input(const bsq_parser* p);

/* ---- FASTA (blazeseq/fasta/parser.mojo:60-200) ---------------------------------------------------
 * FastaParser.next_record in a loop over one region (<= 2 GiB - 1 MiB): every line is stripped of blanks at both
 * ends, a line that then begins with '>' opens a record (id = the rest, stripped), every other line is sequence and
 * is appended without its line break.  Results stay on the device: one contiguous sequence arena, n+1 offsets into
 * it, and the span of every id in the input.  The stop is BSQ_EOF, or the reference's error with its context:
 * BSQ_OTHER "FASTA: sequence id line does not start with '>'" / "FASTA record has empty sequence", or
 * BSQ_ASCII_INVALID under cfg.check_ascii (fasta/parser.mojo:40-58). */
typedef struct bsq_fasta_result {
    int64_t n_records;             /* records before the stop */
    int64_t n_bases;               /* sequence bytes of those records */
    int64_t n_lines;
    bsq_error stop;
} bsq_fasta_result;
typedef struct bsq_fasta_view {    /* DEVICE pointers, valid until the next pass */
    int64_t n_records;
    int64_t sequence_bytes;
    const uint8_t* sequence;       /* record r: sequence[seq_starts[r] .. seq_starts[r+1]) */
    const uint64_t* seq_starts;    /* n_records + 1 */
    const uint32_t* id_start;      /* id of record r: input[id_start[r] .. + id_len[r]) */
    const uint32_t* id_len;
    const uint8_t* input;
} bsq_fasta_view;
bsq_status bsq_fasta_parse_device(bsq_parser* p, const uint8_t* dev_bytes, uint64_t n, bsq_fasta_result* out);
bsq_status bsq_fasta_parse_host(bsq_parser* p, const uint8_t* host_bytes, uint64_t n, bsq_fasta_result* out);
bsq_status bsq_fasta_get(const bsq_parser* p, bsq_fasta_view* out);
/* Host copies: seq (sequence_bytes), seq_starts (n+1), the ids packed back to back (ids, with n+1 offsets). */
bsq_status bsq_fasta_to_host(bsq_parser* p, uint8_t* seq, uint64_t* seq_starts, uint8_t* ids, uint64_t* id_starts);

/* ---- measurement hooks ------------------------------------------------------------------------ */

/* Device time (ms, CUDA events on the parser's stream) of the kernels of the last pass:
 * [0] summarise runs, [1] scan, [2] resolve/validate/pack, [3] id pipeline + tail + rebase,
 * [4] whole pass including H2D for host passes.  n_launches = kernels launched by the pass. */
bsq_status bsq_last_timing(const bsq_parser* p, float ms[5], int64_t* n_launches);

/* ---- synthetic input (blazeseq/utils.mojo:640-678,831-917), generated on the device -------- */

int64_t bsq_compute_num_reads_for_size(int64_t target_size_bytes, int64_t min_length,
                                       int64_t max_length);
int64_t bsq_synth_size(int64_t num_reads, int64_t min_length, int64_t max_length);
/* Writes records [first, first+count) of generate_synthetic_fastq_buffer(num_reads, ...) to
 * dev_out (device memory, capacity bytes); returns bytes written through *written. */
bsq_status bsq_synth_device(bsq_parser* p, uint8_t* dev_out, uint64_t capacity, int64_t num_reads,
                            int64_t first, int64_t count, int64_t min_length, int64_t max_length,
                            int64_t min_phred, int64_t max_phred, uint8_t q_lower,
                            uint8_t q_upper, uint8_t q_offset, uint64_t* written);

/* ---- multi-GPU: shard stitching (host arithmetic) ------------------------------- */

/* Summary of a byte range under the newline-rank algebra (csrc/tile_math.h).  A rank summarises
 * its shard with bsq_summarize_device, all-gathers the 64-byte records, and every rank derives
 * where its first record starts with bsq_shard_prefix -- only 64 bytes per rank cross GPUs. */
typedef struct bsq_summary { uint32_t w[16]; } bsq_summary;
typedef struct bsq_shard_start {
    int64_t newline_rank;          /* rank of the shard's first newline in the whole stream */
    int64_t first_record;          /* index of the first record that STARTS in the shard */
    int64_t skip_bytes;            /* leading bytes that belong to a record of the previous shard;
                                      == shard size when no record starts in the shard */
    int32_t phase;                 /* newline_rank mod 4: lines of the open record already seen */
    int32_t _pad;
} bsq_shard_start;
bsq_status bsq_summarize_device(bsq_parser* p, const uint8_t* dev_bytes, uint64_t n,
                                bsq_summary* out);
/* shards[0..n_shards) in stream order with their sizes; fills start[i] for every shard.  A record
 * is owned by the shard that holds its first byte, so rank i parses the bytes
 * [skip_bytes[i], shard_bytes[i] + skip_bytes[i+1]) of its shard (+ halo) as one region. */
bsq_status bsq_shard_prefix(const bsq_summary* shards, const uint64_t* shard_bytes, int32_t n_shards,
                            bsq_shard_start* start);

#ifdef __cplusplus
}
#endif
#endif /* BLAZESEQ_GPU_H */
