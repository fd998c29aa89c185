// bsq_aux.cuh -- cold-path and bookkeeping kernels: the tail record, per-batch rebasing of the
// cumulative ends, the id strip pipeline, and the synthetic FASTQ generator.
#pragma once
#include "bsq_device.cuh"

namespace bsq {

// ------------------------------------------------------------------------------------------------
// k_tail: the last record of a stream that does not end in '\n' (three newlines found).
// Reference: _next_ref_complete + _check_end_qual (parser.mojo:460-475, utils.mojo:292-329): if
// the bytes after the third newline are all of {\n, \r, ' ', \t} there is no record (the parser
// then raises an empty Error, parser.mojo:350-351); otherwise the record ends at the end of the
// stream and is accepted WITHOUT the structure check, but the validators still run
// (parser.mojo:160-170).
// ------------------------------------------------------------------------------------------------

struct TailParams {
    const uint8_t* base;         // window base
    uint32_t hs, nl0, nl1, nl2, end;  // record start, the three newlines, end of stream
    uint32_t k;                  // window-local record index (== n_complete)
    uint32_t check_ascii, check_quality, lower, upper, q5_width;
    uint32_t want_offsets, want_pack, id_fast;
    uint32_t seq_rel, qual_rel, id_rel;   // window-relative stream offsets (totals)
    uint32_t newline_rank;       // rank of the virtual 4th newline (4k + 3)
};

struct TailOut {
    uint32_t status;             // 0: record emitted, 11: blank remainder (BSQ_EMPTY_ERROR)
    uint32_t id_len, seq_len, qual_len;
};

__global__ void __launch_bounds__(256, 1) k_tail(const TailParams T, const ResolveParams P, TailOut* __restrict__ out) {
    __shared__ uint32_t s_flag[3];
    __shared__ uint32_t s_id[2];
    const uint32_t tid = threadIdx.x;
    if (tid < 3) s_flag[tid] = 0;
    __syncthreads();
    const uint8_t* B = T.base;
    const uint32_t qs = T.nl2 + 1u;
    uint32_t nonblank = 0;
    for (uint32_t x = qs + tid; x < T.end; x += blockDim.x) {
        const uint32_t b = B[x];
        if (b != '\n' && b != '\r' && b != ' ' && b != '\t') nonblank = 1;
    }
    if (nonblank) s_flag[0] = 1;
    __syncthreads();
    if (!s_flag[0]) {
        if (tid == 0) { out->status = 11u; out->id_len = out->seq_len = out->qual_len = 0; }
        return;
    }
    // spans (parser.mojo:355-366); a negative id length (empty header line) is clamped to 0
    const uint32_t seq_s = T.nl0 + 1u, seq_len = T.nl1 - T.nl0 - 1u, qual_len = T.end - qs;
    if (tid == 0) {
        uint32_t a = T.hs + 1u, e = T.nl0;
        if (e < a) e = a;
        while (a < e && bsq_is_space(B[a])) ++a;          // _strip_spaces, utils.mojo:221-242
        while (e > a && bsq_is_space(B[e - 1u])) --e;
        s_id[0] = a; s_id[1] = e - a;
    }
    __syncthreads();
    const uint32_t id_s = s_id[0], id_len = s_id[1];
    // validators (record.mojo:106-116, 76-104)
    uint32_t hi = 0, bad = 0;
    if (T.check_ascii) {
        for (uint32_t x = tid; x < id_len; x += blockDim.x) hi |= B[id_s + x] & 0x80u;
        for (uint32_t x = tid; x < seq_len; x += blockDim.x) hi |= B[seq_s + x] & 0x80u;
        for (uint32_t x = tid; x < qual_len; x += blockDim.x) hi |= B[qs + x] & 0x80u;
    }
    if (T.check_quality) {
        const uint32_t body = T.q5_width ? qual_len - qual_len % T.q5_width : 0u;   // record.mojo:90-102 as written
        for (uint32_t x = tid; x < qual_len; x += blockDim.x) {
            const uint32_t b = B[qs + x];
            if (b < T.lower || b > T.upper || (x < body && b == T.upper)) bad = 1;
        }
    }
    if (hi) s_flag[1] = 1;
    if (bad) s_flag[2] = 1;
    __syncthreads();
    if (tid == 0) {
        if (s_flag[1]) report(P, T.k, 4u);
        else if (s_flag[2]) report(P, T.k, 5u);
        out->status = 0; out->id_len = id_len; out->seq_len = seq_len; out->qual_len = qual_len;
        if (T.want_offsets) P.line_ends[1u + T.newline_rank] = T.end;  // record_end
        if (T.want_offsets || T.want_pack) { P.id_spans[2u * T.k] = id_s; P.id_spans[2u * T.k + 1u] = id_len; }
        if (T.want_pack) {
            const int64_t gk = P.rec_base + (int64_t)T.k;
            const int64_t endv = P.qual_base64 + (int64_t)T.qual_rel + (int64_t)qual_len;
            P.ends_abs[gk] = endv;
            if ((gk + 1) % P.batch_size == 0) P.ends_base[(gk + 1) / P.batch_size] = endv;   // cold: one thread
            if (T.id_fast) {
                const int64_t iend = P.id_base64 + (int64_t)T.id_rel + (int64_t)id_len;
                P.id_ends_abs[gk] = iend;
                if ((gk + 1) % P.batch_size == 0) P.id_ends_base[(gk + 1) / P.batch_size] = iend;
            }
        }
    }
    if (T.want_pack) {
        uint8_t* so = P.seq_out + P.seq_base64 + T.seq_rel;
        uint8_t* qo = P.qual_out + P.qual_base64 + T.qual_rel;
        for (uint32_t x = tid; x < seq_len; x += blockDim.x) so[x] = B[seq_s + x];
        for (uint32_t x = tid; x < qual_len; x += blockDim.x) qo[x] = B[qs + x];
        if (T.id_fast) {
            uint8_t* io = P.id_out + P.id_base64 + T.id_rel;
            for (uint32_t x = tid; x < id_len; x += blockDim.x) io[x] = B[id_s + x];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// per-batch rebasing: FastqBatch._ends / _id_ends restart at 0 for every batch
// (record_batch.mojo:82-87).  *_base[b] = cumulative value at the end of record b*m - 1.
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_rebase(int64_t* __restrict__ ends, int64_t* __restrict__ id_ends,
                                                const int64_t* __restrict__ ends_base,
                                                const int64_t* __restrict__ id_ends_base, int64_t n,
                                                int32_t batch_size) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t b = i / batch_size;
        ends[i] -= ends_base[b];
        id_ends[i] -= id_ends_base[b];
    }
}

// ------------------------------------------------------------------------------------------------
// A consumer of the device-resident SoA (the role of examples/nw_gpu/kernels.mojo:21-89, which takes a
// DeviceFastqBatch's qual_buffer + ends): per-record sum of Phred scores, one warp per record, straight
// from the quality arena -- no host round trip between the parse and the consumer.
// ends[] are per-batch rebased (record_batch.mojo:82-87); ends_base[b] is the arena offset of batch b.
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_quality_sums(const uint8_t* __restrict__ qual, const int64_t* __restrict__ ends,
                                                      const int64_t* __restrict__ ends_base, int64_t first, int64_t count,
                                                      int32_t batch_size, uint32_t offset, int32_t* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31u;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < count; r += warps) {
        const int64_t k = first + r, b = k / batch_size;
        const int64_t base = ends_base[b];
        const int64_t lo = base + (k % batch_size == 0 ? 0 : ends[k - 1]), hi = base + ends[k];
        // aligned 4-byte words covering [lo, hi); bytes outside the record are masked off
        const uint32_t* words = reinterpret_cast<const uint32_t*>(qual + (lo & ~int64_t(3)));
        const int64_t w_lo = lo & ~int64_t(3);
        const int64_t n_words = ((hi + 3) >> 2) - (w_lo >> 2);
        uint32_t sum = 0;
        for (int64_t w = lane; w < n_words; w += 32) {
            uint32_t v = __ldg(words + w);
            const int64_t p0 = w_lo + 4 * w;                      // arena offset of the word's first byte
            if (p0 < lo) v &= 0xFFFFFFFFu << (8u * (uint32_t)(lo - p0));
            if (p0 + 4 > hi) v &= 0xFFFFFFFFu >> (8u * (uint32_t)(p0 + 4 - hi));
            sum = __dp4a(v, 0x01010101u, sum);                    // sum of the four (unsigned) bytes
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
        if (lane == 0) out[r] = (int32_t)sum - (int32_t)(offset * (uint32_t)(hi - lo));
    }
}

// ------------------------------------------------------------------------------------------------
// FastqRecord.write (fastq/record.mojo:390-402, byte_len :384-388) over the device SoA: records [first, first + count)
// serialised back to four-line FASTQ text on the device -- '@' id '\n' sequence '\n' '+' '\n' quality '\n'.
// k_write_sizes: byte_len of every record; (exclusive scan by the caller); k_write_records: one warp per record.
// ------------------------------------------------------------------------------------------------

struct WriteParams {
    const uint8_t* seq; const uint8_t* qual; const uint8_t* id;
    const int64_t* ends; const int64_t* id_ends; const int64_t* ends_base; const int64_t* id_ends_base;
    int64_t first, count;
    int32_t batch_size;
};

__device__ __forceinline__ void write_span(const WriteParams& W, int64_t k, int64_t& lo, int64_t& hi, int64_t& id_lo, int64_t& id_hi) {
    const int64_t b = k / W.batch_size;
    const bool head = k % W.batch_size == 0;
    const int64_t base = W.ends_base[b], id_base = W.id_ends_base[b];
    lo = base + (head ? 0 : W.ends[k - 1]); hi = base + W.ends[k];
    id_lo = id_base + (head ? 0 : W.id_ends[k - 1]); id_hi = id_base + W.id_ends[k];
}

__global__ void __launch_bounds__(256) k_write_sizes(const WriteParams W, unsigned long long* __restrict__ sizes) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= W.count; r += stride) {
        if (r == W.count) { sizes[r] = 0ull; continue; }           // (the scan leaves the total here)
        int64_t lo, hi, id_lo, id_hi;
        write_span(W, W.first + r, lo, hi, id_lo, id_hi);
        sizes[r] = (unsigned long long)(1 + (id_hi - id_lo) + 2 * (hi - lo) + 5);
    }
}

__global__ void __launch_bounds__(256) k_write_records(const WriteParams W, const unsigned long long* __restrict__ offs,
                                                       uint8_t* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31u;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < W.count; r += warps) {
        int64_t lo, hi, id_lo, id_hi;
        write_span(W, W.first + r, lo, hi, id_lo, id_hi);
        const int64_t n = hi - lo, ni = id_hi - id_lo;
        uint8_t* o = out + offs[r];
        // [ '@' | id | '\n' | seq | '\n' '+' '\n' | qual | '\n' ]
        if (lane == 0) { o[0] = '@'; o[1 + ni] = '\n'; o[2 + ni + n] = '\n'; o[3 + ni + n] = '+'; o[4 + ni + n] = '\n'; o[5 + ni + 2 * n] = '\n'; }
        for (int64_t i = lane; i < ni; i += 32) o[1 + i] = W.id[id_lo + i];
        uint8_t* os = o + 2 + ni;
        uint8_t* oq = o + 5 + ni + n;
        for (int64_t i = lane; i < n; i += 32) { os[i] = W.seq[lo + i]; oq[i] = W.qual[lo + i]; }
    }
}

// ------------------------------------------------------------------------------------------------
// id strip pipeline (taken only when some header needs _strip_spaces, e.g. CRLF input)
// ------------------------------------------------------------------------------------------------

struct WindowTable {
    const uint8_t* base[kMaxWindows];
    int64_t rec_base[kMaxWindows + 1];   // arena index of each window's first record; [n] = total
    int32_t n;
};

__global__ void __launch_bounds__(256) k_id_lens(const uint32_t* __restrict__ id_spans, int64_t* __restrict__ id_ends,
                                                 int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        id_ends[i] = (int64_t)id_spans[2 * i + 1];
}

__global__ void __launch_bounds__(256) k_id_bases(const int64_t* __restrict__ id_ends, int64_t* __restrict__ id_ends_base,
                                                  int64_t n, int32_t batch_size) {
    const int64_t nb = (n + batch_size - 1) / batch_size;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= nb; b += stride) {
        const int64_t last = b * batch_size - 1;
        id_ends_base[b] = (b == 0) ? 0 : id_ends[last < n ? last : n - 1];
    }
}

// One thread per record: the ids are short (tens of bytes), so the copies of a warp's 32 records are 32
// independent byte streams in flight (a warp per record left 19 of 32 lanes idle and serialised the
// span -> source -> destination dependency of every record).
__global__ void __launch_bounds__(256) k_id_copy(const WindowTable WT, const uint32_t* __restrict__ id_spans,
                                                 const int64_t* __restrict__ id_ends, uint8_t* __restrict__ id_out,
                                                 int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int w = 0;
        while (w + 1 < WT.n && i >= WT.rec_base[w + 1]) ++w;
        const uint2 sp = *reinterpret_cast<const uint2*>(id_spans + 2 * i);
        const uint32_t len = sp.y;
        const uint8_t* src = WT.base[w] + sp.x;
        uint8_t* dst = id_out + (id_ends[i] - (int64_t)len);
        uint32_t x = 0;
        for (; x + 4u <= len; x += 4u) {
            const uint8_t b0 = src[x], b1 = src[x + 1], b2 = src[x + 2], b3 = src[x + 3];
            dst[x] = b0; dst[x + 1] = b1; dst[x + 2] = b2; dst[x + 3] = b3;
        }
        for (; x < len; ++x) dst[x] = src[x];
    }
}

// ------------------------------------------------------------------------------------------------
// synthetic FASTQ, generate_synthetic_fastq_buffer (utils.mojo:736-917), one warp per record
// ------------------------------------------------------------------------------------------------

struct SynthParams {
    int64_t num_reads, first, count;
    int64_t min_len, max_len, min_phred, max_phred;
    int32_t digits;
    uint32_t q_lower, q_upper, q_offset;
    uint64_t origin;              // byte offset of record `first` in the full stream
    const uint64_t* len_prefix;   // [m+1]: sum_{j<x} ((31 j + 7) mod m), m = max_len - min_len + 1
    uint64_t period_sum;          // len_prefix[m]
    uint64_t a32_seq, c32_seq;    // 32-step jump of the sequence LCG (mod 2^64)
    uint64_t a32_q, c32_q;        // 32-step jump of the quality LCG
};

__device__ __forceinline__ uint64_t synth_offset(const SynthParams& G, uint64_t i) {
    const uint64_t m = (uint64_t)(G.max_len - G.min_len + 1);
    const uint64_t lens = i * (uint64_t)G.min_len + (i / m) * G.period_sum + G.len_prefix[i % m];
    return i * (uint64_t)(6 + G.digits + 1 + 4) + 2ull * lens;
}

__global__ void __launch_bounds__(256) k_synth(const SynthParams G, uint8_t* __restrict__ out) {
    const uint64_t M63 = 0x7FFFFFFFFFFFFFFFull;
    const uint64_t A = 6364136223846793005ull, C = 1442695040888963407ull;  // utils.mojo:773-779
    const uint64_t QA = 1664525ull, QC = 1013904223ull;                     // utils.mojo:808
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    const uint8_t lut[8] = {'G', 'C', 'G', 'C', 'A', 'T', 'A', 'T'};        // gc_bias 0.5, utils.mojo:707-733
    const int64_t q_start = G.max_phred, q_range = G.max_phred - G.min_phred;
    const int64_t noise_amp = q_range / 6 + 1;
    for (int64_t r = warp; r < G.count; r += nwarps) {
        const uint64_t i = (uint64_t)(G.first + r);
        const int64_t m = G.max_len - G.min_len + 1;
        const int64_t rl = G.min_len + (int64_t)((i * 31ull + 7ull) % (uint64_t)m);
        uint8_t* w = out + (synth_offset(G, i) - G.origin);
        // header "@read_<i zero padded>\n"
        const int hdr = 6 + G.digits + 1;
        if ((int)lane < hdr) {
            uint8_t ch;
            if (lane < 6u) ch = (uint8_t)"@read_"[lane];
            else if ((int)lane == hdr - 1) ch = '\n';
            else {
                uint64_t v = i;
                for (int d = G.digits - 1 - ((int)lane - 6); d > 0; --d) v /= 10ull;
                ch = (uint8_t)('0' + v % 10ull);
            }
            w[lane] = ch;
        }
        w += hdr;
        // sequence: state_0 = (i*A + C); base p uses state_{p+1}
        uint64_t s = i * A + C;
        for (uint32_t k = 0; k <= lane; ++k) s = s * A + C;
        for (int64_t p = lane; p < rl; p += 32) {
            w[p] = lut[((s & M63) >> 33) & 7ull];
            s = s * G.a32_seq + G.c32_seq;
        }
        if (lane == 0) { w[rl] = '\n'; w[rl + 1] = '+'; w[rl + 2] = '\n'; w[2 * rl + 3] = '\n'; }
        w += rl + 3;
        // quality: linear decay + LCG noise (utils.mojo:795-827)
        uint64_t q = i * 2654435761ull + 1013904223ull;
        for (uint32_t k = 0; k <= lane; ++k) q = q * QA + QC;
        const int64_t lm1 = rl - 1;
        for (int64_t p = lane; p < rl; p += 32) {
            const int64_t mean = lm1 == 0 ? q_start : q_start - (q_range * p + lm1 / 2) / lm1;
            const int64_t noise_raw = (int64_t)(((q & M63) >> 17) % (uint64_t)(2 * noise_amp + 1));
            int64_t ph = mean + noise_raw - noise_amp;
            ph = ph < G.min_phred ? G.min_phred : (ph > G.max_phred ? G.max_phred : ph);
            int64_t a = (int64_t)G.q_offset + ph;
            a = a < (int64_t)G.q_lower ? (int64_t)G.q_lower : (a > (int64_t)G.q_upper ? (int64_t)G.q_upper : a);
            w[p] = (uint8_t)a;
            q = q * G.a32_q + G.c32_q;
        }
    }
}

}  // namespace bsq
