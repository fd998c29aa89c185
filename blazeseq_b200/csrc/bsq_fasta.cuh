// bsq_fasta.cuh -- multi-line FASTA on the newline machinery of the FASTQ path.
//
// Reference: blazeseq/fasta/parser.mojo:60-200.  A FASTA stream is line oriented: every line is stripped of blanks
// at both ends (_strip_spaces, utils.mojo:221-242); a stripped line that begins with '>' opens a record (id = the
// rest of it, stripped again); every other line is sequence, appended without its line break; blank lines vanish.
// On the device that is: the table of ALL newline positions (k_summarize + k_resolve, the views() pass), one
// thread per line to strip and classify it (k_fa_lines), two prefix sums over the lines (record index = headers
// so far, sequence offset = sequence bytes so far), and one warp per line to move the sequence bytes to their
// place in one contiguous arena (k_fa_pack).  Errors are the reference's: a first non-blank line that is not a
// header, a record without sequence bytes, a byte >= 0x80 under check_ascii.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tile_math.h"

namespace bsq {

struct FastaLines {
    const uint8_t* base;         // window base; positions below are relative to it
    const uint32_t* line_ends;   // [0] = begin - 1, [1 + j] = position of newline j
    uint32_t n_newlines, n_lines, end;   // lines = newlines (+ 1 when the stream does not end in '\n')
    uint32_t* hdr;               // [line] 1 = header line
    unsigned long long* slen;    // [line] sequence bytes of the line (0 for headers and blank lines)
    uint32_t* start;             // [line] first byte after stripping (header: of the id)
    uint32_t* len;               // [line] stripped length (header: of the id)
};

// one thread per line: strip, classify
__global__ void __launch_bounds__(256) k_fa_lines(const FastaLines F) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < F.n_lines; i += stride) {
        uint32_t a = F.line_ends[i] + 1u;
        uint32_t z = i < F.n_newlines ? F.line_ends[i + 1] : F.end;
        const uint8_t* B = F.base;
        while (a < z && bsq_is_space(B[a])) ++a;              // (covers the '\r' LineIterator trims, buffered.mojo:621)
        while (z > a && bsq_is_space(B[z - 1u])) --z;
        const bool header = z > a && B[a] == '>';
        if (header) {                                          // id = the line after '>', stripped again (parser.mojo:140-141)
            ++a;
            while (a < z && bsq_is_space(B[a])) ++a;
        }
        F.hdr[i] = header ? 1u : 0u;
        F.slen[i] = header ? 0ull : (unsigned long long)(z - a);
        F.start[i] = a;
        F.len[i] = z - a;
    }
}

struct FastaPack {
    const uint8_t* base;
    uint32_t n_lines, n_records, check_ascii;
    const uint32_t* hdr; const uint32_t* hcum;               // inclusive count of headers up to and including the line
    const unsigned long long* soff;                          // exclusive sum of the sequence bytes
    const uint32_t* start; const uint32_t* len;
    uint8_t* seq_out;
    unsigned long long* seq_start;                           // [record] offset of its sequence (+ [n_records] = total)
    uint32_t* id_start; uint32_t* id_len; uint32_t* hdr_line;   // [record]
    unsigned long long total_seq;
    // first offenders (atomicMin): [0] line before any header that is not blank, [1] record with a non-ASCII byte
    uint32_t* err;
};

// one warp per line: sequence lines are copied to seq_out[soff ...], header lines fill the record table
__global__ void __launch_bounds__(256) k_fa_pack(const FastaPack K) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < K.n_lines; i += warps) {
        const uint32_t a = K.start[i], n = K.len[i], h = K.hdr[i], rec1 = K.hcum[i];   // rec1 = record index + 1
        const uint8_t* s = K.base + a;
        uint32_t hi = 0;
        if (h) {
            const uint32_t r = rec1 - 1u;
            if (lane == 0) {
                K.seq_start[r] = K.soff[i];
                K.id_start[r] = a; K.id_len[r] = n; K.hdr_line[r] = i;
                if (r + 1u == K.n_records) K.seq_start[K.n_records] = K.total_seq;
            }
            if (K.check_ascii)
                for (uint32_t x = lane; x < n; x += 32u) hi |= s[x];
        } else if (n != 0u) {
            if (rec1 == 0u) {                                 // sequence before the first header (parser.mojo:193-197)
                if (lane == 0) atomicMin(&K.err[0], i);
                continue;
            }
            uint8_t* d = K.seq_out + K.soff[i];
            for (uint32_t x = lane; x < n; x += 32u) {
                const uint8_t b = s[x];
                d[x] = b;
                hi |= b;
            }
        }
        if (K.check_ascii) {
            hi = __reduce_or_sync(0xFFFFFFFFu, hi & 0x80u);
            if (hi != 0u && lane == 0 && rec1 != 0u) atomicMin(&K.err[1], rec1 - 1u);
        }
    }
}

// records without a single sequence byte (parser.mojo:152-160): the first one
__global__ void __launch_bounds__(256) k_fa_empty(const unsigned long long* __restrict__ seq_start, uint32_t n_records,
                                                  uint32_t* __restrict__ err) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_records; r += stride)
        if (seq_start[r + 1] == seq_start[r]) atomicMin(&err[2], r);
}

}  // namespace bsq
