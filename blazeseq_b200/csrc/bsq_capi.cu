// bsq_capi.cu -- the C ABI (include/blazeseq_gpu.h): parser object, pass driver, result views.
//
// Host logic only decides WHERE kernels run (windows, runs, buffers) and formats errors; every
// byte of the FASTQ stream is examined on the device.  There is no CPU parsing path.
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include <zlib.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/blazeseq_gpu.h"
#include "bsq_aux.cuh"
#include "bsq_device.cuh"
#include "bsq_inflate.cuh"
#include "bsq_fasta.cuh"
#include "bsq_pgzip.h"

using namespace bsq;

static_assert(sizeof(bsq_summary) == sizeof(BsqSummary), "ABI summary is a BsqSummary");

namespace {

constexpr uint64_t kWindowMax = (1ull << 31) - (1ull << 20);  // bytes per window (u32 offsets)
constexpr uint64_t kHostWindow = 256ull << 20;                // window size while bytes stream in

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes, size_t slack = 0) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + slack;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Window {
    const uint8_t* base = nullptr;   // device, 16-byte aligned
    uint64_t region_off = 0;         // region offset of the window's first valid byte
    WinParams wp{};
    ScanOut scan{};
    int64_t rec_base = 0;            // arena index of first record
    int64_t seq_base = 0, qual_base = 0, id_base = 0;
    DevBuf line_ends;                // views(): newlines + 2 entries
    DevBuf run_pre;                  // BsqPrefix per run (kept for the resolve pass)
    DevBuf nl_count, nl_list;        // per tile: newline count and ordered list (k_summarize -> k_resolve)
};

struct HostMirror {                  // pinned
    ScanOut scan;
    unsigned long long err[4];       // [0] min error key, [1] "some id needed stripping", [2] bases
    TailOut tail;
};

}  // namespace

struct bsq_parser {
    bsq_config cfg{};
    int sm_count = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev[6]{};             // timing marks
    cudaEvent_t ev_copy[2]{};
    DevBuf run_sum, scan_out, err_word, tail_out, cub_tmp, len_prefix;   // err_word: [0] error key, [1] strip flag
    DevBuf seq_out, qual_out, id_out, ends, id_ends, ends_base, id_ends_base, id_spans;
    DevBuf host_input;               // device copy of a host pass
    DevBuf write_offs, write_buf;    // bsq_write_records: record offsets, text when the caller gives no device buffer
    // buffers of the last device-inflating stream (inflated regions, compressed bytes, member tables, pinned status): the next
    // stream of this parser takes them over instead of allocating its own
    DevBuf inf_rdev[2], inf_zdev[2], inf_mdev[2], inf_sdev[2];
    uint32_t* inf_status[2] = {nullptr, nullptr};
    size_t inf_status_cap[2] = {0, 0};
    void* pinned_stage[2] = {nullptr, nullptr};
    size_t pinned_stage_bytes = 0;
    HostMirror* hm = nullptr;        // pinned
    std::vector<Window> win;
    std::vector<int64_t> h_ends_base, h_id_ends_base;
    // last pass
    bool have_pass = false;
    uint32_t want = 0;
    bsq_pass_result res{};
    const uint8_t* pass_input = nullptr;
    int64_t pass_stream_offset = 0;
    int64_t total_records = 0;       // arena records (complete + tail)
    int64_t total_seq = 0, total_qual = 0, total_id = 0;
    bool tail_emitted = false;       // the arena's last record is a newline-less tail
    int64_t tail_seq = 0, tail_qual = 0;
    float ms[5] = {0, 0, 0, 0, 0};
    int64_t n_launches = 0;
    std::string last_error;
    // FASTA (bsq_fasta_*): line tables, record tables and the sequence arena of the last FASTA pass
    DevBuf fa_hdr, fa_slen, fa_start, fa_len, fa_hcum, fa_soff, fa_seq, fa_seq_start, fa_id_start, fa_id_len, fa_hdr_line, fa_err;
    Window fa_win;
    bool fa_have = false;
    int64_t fa_records = 0, fa_total_records = 0;
    uint64_t fa_seq_bytes = 0;
    const uint8_t* fa_input = nullptr;
};

namespace {

bsq_status fail_cuda(bsq_parser* p, cudaError_t e, const char* what) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    p->last_error = buf;
    cudaGetLastError();
    return BSQ_E_CUDA;
}
#define CK(call)                                              \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return fail_cuda(p, _e, #call); \
    } while (0)

// errors.mojo:71-90
const char* code_message(int code) {
    switch (code) {
        case BSQ_ID_NO_AT: return "Sequence id line does not start with '@'";
        case BSQ_SEP_NO_PLUS: return "Separator line does not start with '+'";
        case BSQ_SEQ_QUAL_LEN_MISMATCH: return "Quality and sequence line do not match in length";
        case BSQ_ASCII_INVALID: return "Non ASCII letters found";
        case BSQ_QUALITY_OUT_OF_RANGE: return "Corrupt quality score according to provided schema";
        case BSQ_UNEXPECTED_EOF: return "Unexpected end of file in FASTQ record";
        case BSQ_BUFFER_EXCEEDED: return "FASTQ record exceeds buffer capacity";
        case BSQ_BUFFER_AT_MAX: return "FASTQ record exceeds maximum buffer capacity";
        default: return "Parse or validation error";
    }
}

struct Msg {
    char* p; size_t cap, len;
    void bytes(const void* b, size_t n) {
        if (len + n >= cap) n = cap - 1 - len;
        memcpy(p + len, b, n); len += n; p[len] = 0;
    }
    void str(const char* z) { bytes(z, strlen(z)); }
    void i64(int64_t v) { char t[32]; snprintf(t, sizeof t, "%lld", (long long)v); str(t); }
};

void set_plain_error(bsq_error* e, int code, const char* text) {
    memset(e, 0, sizeof *e);
    e->code = code;
    Msg m{e->message, sizeof e->message, 0};
    m.str(text);
}

// k_resolve: kResolveCtas / SM when it packs; without the SoA stage (views, count/validate-only) kViewCtas / SM
size_t smem_bytes(bool pack = true, bool validate = true) {
#if BSQ_COPY_STAGED
    if (pack) return sizeof(TileSmem) + 128;
    return offsetof(TileSmem, stage) + (validate ? 2 * kWords * 4 : 0) + 128;
#else
    (void)pack;   // the validation bitmaps are the tail of TileSmem, only allocated by validating passes
    return (validate ? sizeof(TileSmem) : offsetof(TileSmem, bm_hi)) + 128;
#endif
}
size_t smem_bytes_summarize(bool sums) {   // k_summarize: SumShape<kSums>
    return (sums ? sizeof(SumSmemT<SumShape<true>::kStagesOf>) : sizeof(SumSmemT<SumShape<false>::kStagesOf>)) + 128;
}

template <typename K>
cudaError_t opt_in_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

void plan_window(bsq_parser* p, Window& w, const uint8_t* first_byte, uint64_t bytes, bool pack) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(first_byte);
    const uintptr_t al = a & ~uintptr_t(15);
    w.base = reinterpret_cast<const uint8_t*>(al);
    w.wp.base = w.base;
    w.wp.begin = (uint32_t)(a - al);
    w.wp.end = w.wp.begin + (uint32_t)bytes;
    w.wp.first_tile = w.wp.begin / kTile;
    w.wp.n_tiles = (w.wp.end + kTile - 1) / kTile;
    uint32_t tiles = w.wp.n_tiles - w.wp.first_tile;
    if (tiles == 0) tiles = 1, w.wp.n_tiles = w.wp.first_tile + 1;
    // runs per window: a whole number of waves for the dominant kernel of the pass
    uint32_t max_runs = (uint32_t)std::min<int>((pack ? kRunsPerSmPack : kRunsPerSmView) * p->sm_count, kMaxRuns);
    uint32_t runs = std::min(tiles, max_runs);
    w.wp.tiles_per_run = (tiles + runs - 1) / runs;
    w.wp.n_runs = (tiles + w.wp.tiles_per_run - 1) / w.wp.tiles_per_run;
}

using ResolveKernel = void (*)(const WinParams, const ResolveParams);

ResolveKernel pick_resolve(bool ascii, bool qual, bool offs, bool pack) {
    // ParserConfig(check_ascii, check_quality) x {views, batches}: one instantiation each
    static const ResolveKernel table[16] = {
        k_resolve<false, false, false, false>, k_resolve<false, false, false, true>,
        k_resolve<false, false, true, false>,  k_resolve<false, false, true, true>,
        k_resolve<false, true, false, false>,  k_resolve<false, true, false, true>,
        k_resolve<false, true, true, false>,   k_resolve<false, true, true, true>,
        k_resolve<true, false, false, false>,  k_resolve<true, false, false, true>,
        k_resolve<true, false, true, false>,   k_resolve<true, false, true, true>,
        k_resolve<true, true, false, false>,   k_resolve<true, true, false, true>,
        k_resolve<true, true, true, false>,    k_resolve<true, true, true, true>,
    };
    return table[(ascii ? 8 : 0) | (qual ? 4 : 0) | (offs ? 2 : 0) | (pack ? 1 : 0)];
}

using SummarizeKernel = void (*)(const WinParams, BsqSummary*, uint32_t, uint32_t);
SummarizeKernel pick_summarize(bool sums, bool hi, bool bad) {
    static const SummarizeKernel table[8] = {
        k_summarize<false, false, false>, k_summarize<false, false, true>, k_summarize<false, true, false>,
        k_summarize<false, true, true>,   k_summarize<true, false, false>, k_summarize<true, false, true>,
        k_summarize<true, true, false>,   k_summarize<true, true, true>,
    };
    return table[(sums ? 4 : 0) | (hi ? 2 : 0) | (bad ? 1 : 0)];
}

bsq_status setup_kernels(bsq_parser* p) {
    for (int i = 0; i < 8; ++i) CK(opt_in_smem(pick_summarize(i & 4, i & 2, i & 1), smem_bytes_summarize(i & 4)));
    for (int i = 0; i < 16; ++i) CK(opt_in_smem(pick_resolve(i & 8, i & 4, i & 2, i & 1), smem_bytes()));
    return BSQ_OK;
}

// Summarise + scan one window; leaves the ScanOut in w.scan (host) after a stream sync.
bsq_status summarize_window(bsq_parser* p, Window& w, bool sums, bool hand_off = false) {
    CK(p->run_sum.ensure(sizeof(BsqSummary) * kMaxRuns));
    CK(w.run_pre.ensure(sizeof(BsqPrefix) * kMaxRuns));
    CK(p->scan_out.ensure(sizeof(ScanOut)));
    if (hand_off) {   // the per-tile newline lists k_resolve picks up
        const size_t tiles = w.wp.n_tiles - w.wp.first_tile;
        CK(w.nl_count.ensure(4 * tiles, 1 << 16));
        CK(w.nl_list.ensure(2ull * kNlCap * tiles, 1 << 20));
        w.wp.nl_count = w.nl_count.as<uint32_t>();
        w.wp.nl_list = w.nl_list.as<uint16_t>();
    } else {
        w.wp.nl_count = nullptr; w.wp.nl_list = nullptr;
    }
    // with the hand-off, a validating pass also screens every tile for HI / BAD bytes (k_resolve skips clean tiles)
    const bool hi = hand_off && p->cfg.check_ascii, bad = hand_off && p->cfg.check_quality;
    pick_summarize(sums, hi, bad)<<<w.wp.n_runs, kThreads, smem_bytes_summarize(sums), p->stream>>>(
        w.wp, p->run_sum.as<BsqSummary>(), (uint32_t)p->cfg.q_lower,
        (uint32_t)p->cfg.q_upper - (p->cfg.compat_q5_width > 0 ? 1u : 0u));   // (compat: UPPER itself is suspicious)
    k_scan_runs<<<1, kScanThreads, 0, p->stream>>>(p->run_sum.as<BsqSummary>(), w.wp.n_runs, w.wp.begin,
                                          w.run_pre.as<BsqPrefix>(), p->scan_out.as<ScanOut>());
    p->n_launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&p->hm->scan, p->scan_out.p, sizeof(ScanOut), cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    w.scan = p->hm->scan;
    return BSQ_OK;
}

// D2H of a few bytes of the region (error snippets)
bsq_status fetch_bytes(bsq_parser* p, const uint8_t* dev, size_t n, std::vector<uint8_t>& out) {
    out.resize(n);
    if (n) CK(cudaMemcpy(out.data(), dev, n, cudaMemcpyDeviceToHost));
    return BSQ_OK;
}

// Offsets of window-local record k (5 values) read back from the line-end table.
bsq_status fetch_record_offsets(bsq_parser* p, Window& w, uint32_t k, uint32_t o[5]) {
    uint32_t le[5];
    CK(cudaMemcpy(le, w.line_ends.as<uint32_t>() + 4ull * k, sizeof le, cudaMemcpyDeviceToHost));
    o[0] = le[0] + 1u; o[1] = le[1] + 1u; o[2] = le[2] + 1u; o[3] = le[3] + 1u; o[4] = le[4];
    return BSQ_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// configuration helpers
// ------------------------------------------------------------------------------------------------

extern "C" void bsq_default_config(bsq_config* c) {
    memset(c, 0, sizeof *c);
    c->device_id = 0;
    c->q_lower = 33; c->q_upper = 126; c->q_offset = 33;   // generic_schema, quality_schema.mojo:26
    c->buffer_capacity = 256 * 1024;                       // CONSTS.mojo:26
    c->buffer_max_capacity = 1ll << 30;                    // CONSTS.mojo:27-28
    c->batch_size = 4096;                                  // CONSTS.mojo:31
    c->h2d_chunk_bytes = 64ll << 20;
}

extern "C" int32_t bsq_parse_schema(const char* name, uint8_t* lower, uint8_t* upper, uint8_t* offset) {
    static const struct { const char* n; uint8_t lo, up, off; } tab[] = {
        {"sanger", 33, 126, 33},       {"solexa", 59, 126, 64},       {"illumina_1.3", 64, 126, 64},
        {"illumina_1.5", 66, 126, 64}, {"illumina_1.8", 33, 126, 33}, {"generic", 33, 126, 33}};
    for (auto& t : tab)
        if (name && strcmp(name, t.n) == 0) { *lower = t.lo; *upper = t.up; *offset = t.off; return 0; }
    *lower = 33; *upper = 126; *offset = 33;
    return 1;
}

extern "C" uint32_t bsq_abi_version(void) { return BSQ_ABI_VERSION; }

// ------------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------------

extern "C" bsq_status bsq_create(const bsq_config* cfg, bsq_parser** out) {
    if (!out) return BSQ_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return BSQ_E_NO_DEVICE; }
    bsq_parser* p = new (std::nothrow) bsq_parser();
    if (!p) return BSQ_E_NOMEM;
    if (cfg) p->cfg = *cfg; else bsq_default_config(&p->cfg);
    if (p->cfg.batch_size <= 0) p->cfg.batch_size = 4096;
    if (p->cfg.h2d_chunk_bytes <= 0) p->cfg.h2d_chunk_bytes = 64ll << 20;
    if (p->cfg.q_upper >= 128 || p->cfg.q_lower > p->cfg.q_upper || p->cfg.device_id < 0 || p->cfg.device_id >= ndev) {
        delete p;
        return BSQ_E_ARG;
    }
    bsq_status st = BSQ_OK;
    auto init = [&]() -> bsq_status {
        CK(cudaSetDevice(p->cfg.device_id));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, p->cfg.device_id));
        if (prop.major < 10) { p->last_error = "this build targets sm_100a (B200)"; return BSQ_E_NO_DEVICE; }
        p->sm_count = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
        for (auto& e : p->ev) CK(cudaEventCreate(&e));
        for (auto& e : p->ev_copy) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CK(cudaHostAlloc(reinterpret_cast<void**>(&p->hm), sizeof(HostMirror), cudaHostAllocDefault));
        CK(p->err_word.ensure(sizeof(unsigned long long)));
        CK(p->tail_out.ensure(sizeof(TailOut)));
        return setup_kernels(p);
    };
    st = init();
    if (st != BSQ_OK) { bsq_destroy(p); return st; }
    *out = p;
    return BSQ_OK;
}

extern "C" void bsq_destroy(bsq_parser* p) {
    if (!p) return;
    cudaSetDevice(p->cfg.device_id);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->copy_stream) cudaStreamSynchronize(p->copy_stream);
    for (auto& w : p->win) { w.line_ends.release(); w.run_pre.release(); w.nl_count.release(); w.nl_list.release(); }
    DevBuf* bufs[] = {&p->run_sum, &p->scan_out, &p->err_word, &p->tail_out, &p->cub_tmp, &p->len_prefix,
                      &p->seq_out, &p->qual_out, &p->id_out, &p->ends, &p->id_ends, &p->ends_base,
                      &p->id_ends_base, &p->id_spans, &p->host_input, &p->write_offs, &p->write_buf, &p->fa_hdr, &p->fa_slen, &p->fa_start, &p->fa_len, &p->fa_hcum,
                      &p->fa_soff, &p->fa_seq, &p->fa_seq_start, &p->fa_id_start, &p->fa_id_len, &p->fa_hdr_line, &p->fa_err};
    p->fa_win.line_ends.release(); p->fa_win.run_pre.release(); p->fa_win.nl_count.release(); p->fa_win.nl_list.release();
    for (auto* b : bufs) b->release();
    for (int i = 0; i < 2; ++i) {
        p->inf_rdev[i].release(); p->inf_zdev[i].release(); p->inf_mdev[i].release(); p->inf_sdev[i].release();
        if (p->inf_status[i]) cudaFreeHost(p->inf_status[i]);
    }
    for (auto& s : p->pinned_stage) if (s) cudaFreeHost(s);
    if (p->hm) cudaFreeHost(p->hm);
    for (auto& e : p->ev) if (e) cudaEventDestroy(e);
    for (auto& e : p->ev_copy) if (e) cudaEventDestroy(e);
    if (p->stream) cudaStreamDestroy(p->stream);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    delete p;
}

extern "C" const char* bsq_last_error_text(const bsq_parser* p) { return p ? p->last_error.c_str() : ""; }

extern "C" bsq_status bsq_set_batch_size(bsq_parser* p, int32_t batch_size) {
    if (!p || batch_size <= 0) return BSQ_E_ARG;
    p->cfg.batch_size = batch_size;
    p->have_pass = false;  // views of the previous pass were cut with the old size
    return BSQ_OK;
}

// ------------------------------------------------------------------------------------------------
// the pass driver
// ------------------------------------------------------------------------------------------------

namespace {

// `ready_upto(region_bytes)` makes sure region bytes [0, region_bytes) are on the device before a
// kernel that reads them is enqueued on p->stream (no-op for device passes).
struct InputFeed {
    virtual bsq_status ready_upto(uint64_t) { return BSQ_OK; }
    virtual ~InputFeed() {}
};

// Two passes: k_summarize + k_scan_runs over every window (the host learns where the next window starts and
// the exact output sizes), then k_resolve.
bsq_status run_pass(bsq_parser* p, const uint8_t* d, uint64_t n, int64_t stream_offset, int64_t first_record,
                    int32_t is_last, uint32_t want, uint64_t window_bytes, InputFeed& feed, bsq_pass_result* out) {
    const bsq_config& cfg = p->cfg;
    const bool want_offs = (want & BSQ_WANT_OFFSETS) != 0, want_pack = (want & BSQ_WANT_BATCHES) != 0;
    p->have_pass = false;
    p->want = want;
    p->pass_input = d;
    p->pass_stream_offset = stream_offset;
    p->n_launches = 0;
    memset(&p->res, 0, sizeof p->res);
    bsq_pass_result& R = p->res;
    const int32_t m = cfg.batch_size;
    const size_t smem_res = smem_bytes(want_pack, cfg.check_ascii || cfg.check_quality);
    // k_summarize hands every tile's ordered newline list (and, when validating, its HI / BAD screen) to k_resolve
    const bool list_hand_off = true;
    CK(cudaEventRecord(p->ev[0], p->stream));

    bool id_fast = cfg.force_id_slow_path == 0;   // optimistic: k_resolve raises the strip flag when an id needs stripping
    // the reference holds one record in a buffer of buffer_capacity bytes (growing up to buffer_max_capacity when
    // growth is enabled): a longer record ends the parse with BUFFER_EXCEEDED / BUFFER_AT_MAX (parser.mojo:484-503)
    const int64_t rec_limit = std::max<int64_t>(1, cfg.buffer_growth_enabled ? cfg.buffer_max_capacity : cfg.buffer_capacity);
    const int limit_code = cfg.buffer_growth_enabled ? BSQ_BUFFER_AT_MAX : BSQ_BUFFER_EXCEEDED;

    // error word, arenas and tables for `rec` records (+1) and the given byte counts
    auto prepare_outputs = [&](int64_t rec, int64_t seq_bytes, int64_t qual_bytes, int64_t id_bytes) -> bsq_status {
        const int64_t nb_cap = rec / m + 2;
        CK(p->err_word.ensure(32));
        CK(cudaMemsetAsync(p->err_word.p, 0xFF, 8, p->stream));
        CK(cudaMemsetAsync(p->err_word.as<uint8_t>() + 8, 0, 24, p->stream));
        if (want_offs || want_pack) CK(p->id_spans.ensure(8ull * (rec + 1), 1 << 20));
        if (want_pack) {
            CK(p->seq_out.ensure((size_t)seq_bytes + 64, 1 << 20));
            CK(p->qual_out.ensure((size_t)qual_bytes + 64, 1 << 20));
            CK(p->id_out.ensure((size_t)id_bytes, 1 << 20));
            CK(p->ends.ensure(8ull * (rec + 1), 1 << 20));
            CK(p->id_ends.ensure(8ull * (rec + 1), 1 << 20));
            CK(p->ends_base.ensure(8ull * nb_cap, 1 << 16));
            CK(p->id_ends_base.ensure(8ull * nb_cap, 1 << 16));
            CK(cudaMemsetAsync(p->ends_base.p, 0, 8ull * nb_cap, p->stream));
            CK(cudaMemsetAsync(p->id_ends_base.p, 0, 8ull * nb_cap, p->stream));
        }
        return BSQ_OK;
    };
    auto make_params = [&](Window& w) {
        ResolveParams P{};
        P.run_pre = w.run_pre.as<BsqPrefix>();
        P.n_complete = w.scan.totals.records;
        P.id_fast = id_fast ? 1u : 0u;
        P.strip_flag = reinterpret_cast<uint32_t*>(p->err_word.as<uint8_t>() + 8);
        P.bases = p->err_word.as<unsigned long long>() + 2;
        P.rec_mod = (uint32_t)(w.rec_base % m);
        P.rec_div = w.rec_base / m;
        P.rec_base = w.rec_base;
        P.first_record = first_record;
        P.line_ends = w.line_ends.as<uint32_t>();
        P.id_spans = p->id_spans.p ? p->id_spans.as<uint32_t>() + 2 * w.rec_base : nullptr;
        P.seq_out = p->seq_out.as<uint8_t>(); P.qual_out = p->qual_out.as<uint8_t>(); P.id_out = p->id_out.as<uint8_t>();
        P.seq_base64 = w.seq_base; P.qual_base64 = w.qual_base; P.id_base64 = w.id_base;
        P.ends_abs = p->ends.as<int64_t>(); P.id_ends_abs = p->id_ends.as<int64_t>();
        P.ends_base = p->ends_base.as<int64_t>(); P.id_ends_base = p->id_ends_base.as<int64_t>();
        P.id_cap = (int64_t)p->id_out.cap;
        P.batch_size = m;
        P.lower = cfg.q_lower; P.upper = cfg.q_upper;
        P.rec_limit = (uint32_t)std::min<int64_t>(rec_limit, 0xFFFFFFFFll);
        P.q5_width = cfg.compat_q5_width > 0 ? (uint32_t)cfg.compat_q5_width : 0u;
        P.err = p->err_word.as<unsigned long long>();
        return P;
    };
    size_t nw = 0;
    uint64_t pos = 0;
    bool oversize = false;
    // ---- pass 1: summarise every window (host learns where the next window starts) ----------
    while (pos < n) {
        if (nw >= (size_t)kMaxWindows) { p->last_error = "region needs more than kMaxWindows windows"; return BSQ_E_ARG; }
        if (p->win.size() <= nw) p->win.emplace_back();
        Window& w = p->win[nw];
        const uint64_t bytes = std::min<uint64_t>(n - pos, window_bytes);
        bsq_status st = feed.ready_upto(pos + bytes);
        if (st != BSQ_OK) return st;
        plan_window(p, w, d + pos, bytes, want_pack);
        w.region_off = pos;
        st = summarize_window(p, w, want_pack, list_hand_off);
        if (st != BSQ_OK) return st;
        ++nw;
        const uint64_t consumed = w.scan.totals.consumed_end - w.wp.begin;
        const bool reaches_end = pos + bytes == n;
        if (reaches_end) break;
        if (consumed == 0) {
            if (window_bytes < kWindowMax) { window_bytes = std::min<uint64_t>(window_bytes * 4, kWindowMax); --nw; continue; }
            oversize = true;  // a single record larger than a window
            break;
        }
        pos += consumed;
    }
    CK(cudaEventRecord(p->ev[1], p->stream));

    // ---- totals, tail classification ----------------------------------------------------------
    int64_t nrec = 0, nseq = 0, nqual = 0, nid = 0, nnl = 0;
    for (size_t i = 0; i < nw; ++i) {
        Window& w = p->win[i];
        w.rec_base = nrec; w.seq_base = nseq; w.qual_base = nqual; w.id_base = nid;
        nrec += w.scan.totals.records; nseq += w.scan.totals.seq_bytes; nqual += w.scan.totals.qual_bytes;
        nid += w.scan.totals.id_bytes_unstripped;
        nnl += (i + 1 < nw) ? 4ll * w.scan.totals.records : w.scan.totals.newlines;
    }
    uint64_t consumed_total = 0;
    uint32_t rem = 0;
    bool have_tail_candidate = false;
    Window* lw = nw ? &p->win[nw - 1] : nullptr;
    // SURVEY App. A Q2 (parser.mojo:484-492): the stream's first record is incomplete and the
    // buffer may not grow -> BUFFER_EXCEEDED, whatever the tail looks like
    bool q2 = false, tail_too_long = false;
    // ids: the scanned prefix counts an empty header line as -1; such a line is a structure error,
    // so only records after the first error are affected -- but the arena must still hold them
    int64_t id_room = 64;
    for (size_t i = 0; i < nw; ++i) {
        const Window& w = p->win[i];
        const int64_t bytes = (int64_t)w.wp.end - w.wp.begin;
        id_room += std::min<int64_t>(w.scan.totals.id_bytes_unstripped, bytes) + w.scan.totals.records + 1;
    }
    if (lw) {
        consumed_total = lw->region_off + (lw->scan.totals.consumed_end - lw->wp.begin);
        rem = lw->scan.totals.newlines & 3u;
        q2 = is_last && consumed_total < n && first_record + nrec == 0 && stream_offset == 0 && !cfg.buffer_growth_enabled;
        // (the position sums exist only in pack passes; a wrapped id total means an empty header
        //  line earlier in the window, i.e. an error before the tail -- nothing to append then)
        // an unterminated last record (or fragment) that would not fit the reference's buffer either
        tail_too_long = is_last && !oversize && !q2 && consumed_total < n && (int64_t)(n - consumed_total) > rec_limit;
        have_tail_candidate = is_last && !oversize && !q2 && !tail_too_long && consumed_total < n && rem == 3u &&
                              (!want_pack || (int64_t)lw->scan.totals.id_bytes_unstripped <= (int64_t)lw->wp.end);
    }
    uint32_t tail_seq = 0, tail_qual = 0, tail_id_max = 0;
    TailParams T{};
    if (have_tail_candidate) {
        const BsqSummary& E = lw->scan.end_state;
        T.base = lw->base; T.hs = lw->scan.totals.consumed_end;
        T.nl0 = E.last[2]; T.nl1 = E.last[1]; T.nl2 = E.last[0]; T.end = lw->wp.end;
        T.k = lw->scan.totals.records;
        T.newline_rank = 4u * T.k + 3u;
        tail_seq = T.nl1 - T.nl0 - 1u; tail_qual = T.end - T.nl2 - 1u;
        tail_id_max = T.nl0 > T.hs ? T.nl0 - T.hs - 1u : 0u;
    }

    // ---- outputs -------------------------------------------------------------------------------
    const int64_t arena_rec = nrec + (have_tail_candidate ? 1 : 0);
    ResolveKernel kern = pick_resolve(cfg.check_ascii, cfg.check_quality, want_offs, want_pack);
    bsq_status stp = prepare_outputs(arena_rec, nseq + tail_seq, nqual + tail_qual, id_room + tail_id_max);
    if (stp != BSQ_OK) return stp;

    // ---- pass 2: resolve / validate / pack ------------------------------------------------------
    for (size_t i = 0; i < nw; ++i) {
        Window& w = p->win[i];
        if (want_offs) CK(w.line_ends.ensure(4ull * ((size_t)w.scan.totals.newlines + 2), 1 << 16));
        ResolveParams P = make_params(w);
        kern<<<w.wp.n_runs, kThreads, smem_res, p->stream>>>(w.wp, P);
        p->n_launches += 1;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(p->ev[2], p->stream));
    if (want_pack && id_fast) {
        // did the optimistic id packing hold?  (CRLF input and padded ids need _strip_spaces.)  k_resolve wrote
        // every record's stripped id span, so only the ids are redone, by the strip pipeline below
        CK(cudaMemcpyAsync(p->hm->err, p->err_word.p, 32, cudaMemcpyDeviceToHost, p->stream));
        CK(cudaStreamSynchronize(p->stream));
        if (p->hm->err[1] != 0) id_fast = false;
    }

    if (have_tail_candidate) {
        ResolveParams P = make_params(*lw);
        T.check_ascii = cfg.check_ascii; T.check_quality = cfg.check_quality;
        T.lower = cfg.q_lower; T.upper = cfg.q_upper; T.q5_width = cfg.compat_q5_width > 0 ? (uint32_t)cfg.compat_q5_width : 0u;
        T.want_offsets = want_offs; T.want_pack = want_pack; T.id_fast = id_fast;
        T.seq_rel = lw->scan.totals.seq_bytes; T.qual_rel = lw->scan.totals.qual_bytes;
        T.id_rel = lw->scan.totals.id_bytes_unstripped;
        k_tail<<<1, 256, 0, p->stream>>>(T, P, p->tail_out.as<TailOut>());
        p->n_launches += 1;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&p->hm->tail, p->tail_out.p, sizeof(TailOut), cudaMemcpyDeviceToHost, p->stream));
    }
    CK(cudaMemcpyAsync(p->hm->err, p->err_word.p, 32, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));

    // ---- how far did the pass get? ------------------------------------------------------------
    bsq_error stop;
    memset(&stop, 0, sizeof stop);
    int64_t good = nrec;             // records before the stop
    bool tail_emitted = false;
    if (have_tail_candidate && p->hm->tail.status == 0) tail_emitted = true;
    const unsigned long long key = p->hm->err[0];
    const bool have_err = key != ~0ull;
    int64_t err_rec = -1; int err_code = 0;
    if (have_err) {
        err_rec = (int64_t)(key >> 8) - first_record; err_code = (int)(key & 0xFF);
        if (err_code == 0) err_code = limit_code;   // reported with the lowest code so that it precedes the record's other errors
    }
    int64_t arena_final = nrec + (tail_emitted ? 1 : 0);
    if (have_err && err_rec < arena_final) {
        good = err_rec;
    } else {
        good = arena_final;
        err_code = 0;
    }

    // ---- id strip pipeline + rebase over the records that count ------------------------------
    if (want_pack && arena_final > 0) {
        const int grid = std::max(1, std::min<int>(p->sm_count * 8, (int)((arena_final + 255) / 256)));
        if (!id_fast) {
            k_id_lens<<<grid, 256, 0, p->stream>>>(p->id_spans.as<uint32_t>(), p->id_ends.as<int64_t>(), arena_final);
            size_t tmp = 0;
            CK(cub::DeviceScan::InclusiveSum(nullptr, tmp, p->id_ends.as<int64_t>(), p->id_ends.as<int64_t>(),
                                             (int)arena_final, p->stream));
            CK(p->cub_tmp.ensure(tmp + 16));
            CK(cub::DeviceScan::InclusiveSum(p->cub_tmp.p, tmp, p->id_ends.as<int64_t>(), p->id_ends.as<int64_t>(),
                                             (int)arena_final, p->stream));
            k_id_bases<<<std::max<int>(1, (int)((arena_final / m + 256) / 256)), 256, 0, p->stream>>>(
                p->id_ends.as<int64_t>(), p->id_ends_base.as<int64_t>(), arena_final, m);
            WindowTable WT{};
            WT.n = (int32_t)nw;
            for (size_t i = 0; i < nw; ++i) { WT.base[i] = p->win[i].base; WT.rec_base[i] = p->win[i].rec_base; }
            WT.rec_base[nw] = arena_final;
            const int grid_rec = (int)std::max<int64_t>(1, std::min<int64_t>((arena_final + 255) / 256, 1 << 20));
            k_id_copy<<<grid_rec, 256, 0, p->stream>>>(WT, p->id_spans.as<uint32_t>(), p->id_ends.as<int64_t>(),
                                                   p->id_out.as<uint8_t>(), arena_final);
            p->n_launches += 5;
        }
        // cumulative totals at the end (before rebasing) -> batch directory on the host
        const int64_t nbat = (arena_final + m - 1) / m;
        p->h_ends_base.assign(nbat + 1, 0);
        p->h_id_ends_base.assign(nbat + 1, 0);
        CK(cudaMemcpyAsync(p->h_ends_base.data(), p->ends_base.p, 8ull * nbat, cudaMemcpyDeviceToHost, p->stream));
        CK(cudaMemcpyAsync(p->h_id_ends_base.data(), p->id_ends_base.p, 8ull * nbat, cudaMemcpyDeviceToHost, p->stream));
        CK(cudaMemcpyAsync(&p->h_ends_base[nbat], p->ends.as<int64_t>() + (arena_final - 1), 8, cudaMemcpyDeviceToHost, p->stream));
        CK(cudaMemcpyAsync(&p->h_id_ends_base[nbat], p->id_ends.as<int64_t>() + (arena_final - 1), 8, cudaMemcpyDeviceToHost, p->stream));
        k_rebase<<<grid, 256, 0, p->stream>>>(p->ends.as<int64_t>(), p->id_ends.as<int64_t>(), p->ends_base.as<int64_t>(),
                                              p->id_ends_base.as<int64_t>(), arena_final, m);
        p->n_launches += 1;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(p->ev[3], p->stream));
    CK(cudaStreamSynchronize(p->stream));
    p->total_records = arena_final;
    p->tail_emitted = tail_emitted; p->tail_seq = tail_seq; p->tail_qual = tail_qual;
    p->total_seq = (int64_t)p->hm->err[2] + (tail_emitted ? tail_seq : 0);   // counted by k_resolve
    p->total_qual = nqual + (tail_emitted ? tail_qual : 0);

    // ---- the stop reason, with the reference's context and text -------------------------------
    char limit_text[200];
    if (cfg.buffer_growth_enabled)
        snprintf(limit_text, sizeof limit_text, "FASTQ record exceeds maximum buffer capacity (%lld bytes). Enable buffer growth or increase max_capacity.",
                 (long long)cfg.buffer_max_capacity);
    else
        snprintf(limit_text, sizeof limit_text, "FASTQ record exceeds buffer capacity (%lld bytes). Enable buffer growth or increase buffer_capacity.",
                 (long long)cfg.buffer_capacity);
    R.n_newlines = nnl;
    R.n_windows = (int32_t)nw;
    R.id_slow_path = id_fast ? 0 : 1;
    uint64_t consumed_final = consumed_total;
    if (good < arena_final || (have_err && err_code != 0)) {
        // first error at arena record `good`: find its window and offsets
        size_t wi = 0;
        while (wi + 1 < nw && good >= p->win[wi + 1].rec_base) ++wi;
        Window& w = p->win[wi];
        const uint32_t k = (uint32_t)(good - w.rec_base);
        const bool is_tail_rec = tail_emitted && good == nrec;
        uint32_t o[5];
        if (is_tail_rec) {
            o[0] = T.hs; o[1] = T.nl0 + 1u; o[2] = T.nl1 + 1u; o[3] = T.nl2 + 1u; o[4] = T.end;
        } else {
            if (!want_offs) {  // cold path: materialise this window's line-end table
                CK(w.line_ends.ensure(4ull * ((size_t)w.scan.totals.newlines + 2), 1 << 16));
                CK(p->id_spans.ensure(8ull * (arena_rec + 1), 1 << 20));
                ResolveParams P = make_params(w);
                DevBuf scratch;
                CK(scratch.ensure(8));
                CK(cudaMemset(scratch.p, 0xFF, 8));
                P.err = scratch.as<unsigned long long>();
                k_resolve<false, false, true, false><<<w.wp.n_runs, kThreads, smem_bytes(false, false), p->stream>>>(w.wp, P);
                p->n_launches += 1;
                CK(cudaStreamSynchronize(p->stream));
                scratch.release();
            }
            bsq_status st = fetch_record_offsets(p, w, k, o);
            if (st != BSQ_OK) return st;
        }
        const int64_t win_stream_base = stream_offset + (int64_t)w.region_off - (int64_t)w.wp.begin;
        const int64_t gidx = first_record + good;  // 0-based global record index
        consumed_final = w.region_off + (o[0] - w.wp.begin);
        memset(&stop, 0, sizeof stop);
        stop.code = err_code;
        Msg msg{stop.message, sizeof stop.message, 0};
        msg.str(code_message(err_code));
        if (err_code == BSQ_BUFFER_EXCEEDED || err_code == BSQ_BUFFER_AT_MAX) {
            set_plain_error(&stop, err_code, limit_text);   // parser.mojo:299-309: no context
        } else if (err_code <= 3) {
            // ParseError: parser.mojo:332-338 + errors.mojo:178-192
            stop.record_number = gidx + 1; stop.line_number = 4 * gidx + 1; stop.file_position = win_stream_base + o[0];
            std::vector<uint8_t> sn;
            size_t len = std::min<size_t>((size_t)(o[4] + 1u - o[0]), 200);
            bsq_status st = fetch_bytes(p, w.base + o[0], len, sn);
            if (st != BSQ_OK) return st;
            msg.str("\n  Record number: "); msg.i64(stop.record_number);
            msg.str("\n  Line number: "); msg.i64(stop.line_number);
            if (stop.file_position > 0) { msg.str("\n  File position: "); msg.i64(stop.file_position); }
            if (!sn.empty()) { msg.str("\n  Record snippet: "); msg.bytes(sn.data(), sn.size()); }
        } else {
            // ValidationError: parser.mojo:163-169,597-610 + errors.mojo:223-234
            stop.record_number = gidx + 1;
            uint32_t sp[2];
            CK(cudaMemcpy(sp, p->id_spans.as<uint32_t>() + 2 * good, sizeof sp, cudaMemcpyDeviceToHost));
            const uint32_t seq_len = o[2] - o[1] - 1u;
            std::vector<uint8_t> idb, sqb;
            bsq_status st = fetch_bytes(p, w.base + sp[0], std::min<size_t>(sp[1], 400), idb);
            if (st != BSQ_OK) return st;
            std::string snip(idb.begin(), idb.end());
            if (!snip.empty() && snip.size() < 200) snip.push_back('\n');
            if (snip.size() < 200 && seq_len > 0) {
                st = fetch_bytes(p, w.base + o[1], std::min<size_t>(seq_len, 200 - snip.size()), sqb);
                if (st != BSQ_OK) return st;
                snip.append(sqb.begin(), sqb.end());
            }
            if (snip.size() > 200) { snip.resize(197); snip += "..."; }
            msg.str("\n  Record number: "); msg.i64(stop.record_number);
            if (!snip.empty()) { msg.str("\n  Record snippet: "); msg.bytes(snip.data(), snip.size()); }
        }
    } else if (oversize) {
        char t[200];
        snprintf(t, sizeof t, "FASTQ record exceeds maximum buffer capacity (%lld bytes). Enable buffer growth or increase max_capacity.",
                 (long long)kWindowMax);
        set_plain_error(&stop, BSQ_BUFFER_AT_MAX, t);
    } else if (tail_too_long) {
        set_plain_error(&stop, limit_code, limit_text);
    } else if (!is_last) {
        stop.code = BSQ_OK;  // more input needed; the caller re-presents the unconsumed tail
    } else if (tail_emitted) {
        consumed_final = n;
        set_plain_error(&stop, BSQ_EOF, "EOF");
    } else if (consumed_total == n) {
        set_plain_error(&stop, BSQ_EOF, "EOF");       // parser.mojo:316-317
    } else if (q2) {
        char t[200];
        snprintf(t, sizeof t, "FASTQ record exceeds buffer capacity (%lld bytes). Enable buffer growth or increase buffer_capacity.",
                 (long long)cfg.buffer_capacity);
        set_plain_error(&stop, BSQ_BUFFER_EXCEEDED, t);
    } else if (rem < 3u) {
        char t[96];                                   // parser.mojo:295-298
        snprintf(t, sizeof t, "Unexpected end of file in FASTQ record at phase %u", rem);
        set_plain_error(&stop, BSQ_UNEXPECTED_EOF, t);
    } else {
        set_plain_error(&stop, BSQ_EMPTY_ERROR, "");  // parser.mojo:350-351
    }

    R.n_records = good;
    R.bytes_consumed = (int64_t)consumed_final;
    R.n_batches = want_pack ? (good + m - 1) / m : 0;
    R.stop = stop;
    // bases: sum of sequence lengths of the good records
    if (good == arena_final) {
        R.n_bases = p->total_seq;
    } else if (want_pack && good > 0) {
        // cumulative quality == sequence length for structurally valid records
        int64_t v = 0;
        const int64_t b = (good - 1) / m;
        CK(cudaMemcpy(&v, p->ends.as<int64_t>() + (good - 1), 8, cudaMemcpyDeviceToHost));
        R.n_bases = v + p->h_ends_base[b];
    } else {
        R.n_bases = -1;  // not materialised (views only, error before the end)
    }
    if (want_pack) {
        // total id bytes of the arena
        p->total_id = p->h_id_ends_base.empty() ? 0 : p->h_id_ends_base.back();
    }
    CK(cudaEventRecord(p->ev[4], p->stream));
    CK(cudaEventSynchronize(p->ev[4]));
    cudaEventElapsedTime(&p->ms[0], p->ev[0], p->ev[1]);
    cudaEventElapsedTime(&p->ms[2], p->ev[1], p->ev[2]);
    cudaEventElapsedTime(&p->ms[3], p->ev[2], p->ev[3]);
    cudaEventElapsedTime(&p->ms[4], p->ev[0], p->ev[4]);
    p->ms[1] = 0.f;
    p->have_pass = true;
    if (out) *out = R;
    return BSQ_OK;
}

struct HostFeed : InputFeed {
    bsq_parser* p; const uint8_t* h; uint8_t* d; uint64_t n; uint64_t sent = 0; uint64_t ahead = 0; bool pinned = false; int slot = 0;
    bsq_status send_upto(uint64_t upto) {
        const uint64_t chunk = (uint64_t)p->cfg.h2d_chunk_bytes;
        if (upto > n) upto = n;
        while (sent < upto) {
            const uint64_t len = std::min<uint64_t>(chunk, n - sent);
            if (pinned) {
                CK(cudaMemcpyAsync(d + sent, h + sent, len, cudaMemcpyHostToDevice, p->copy_stream));
            } else {
                // pageable source: bounce through two pinned buffers
                CK(cudaEventSynchronize(p->ev_copy[slot]));
                memcpy(p->pinned_stage[slot], h + sent, len);
                CK(cudaMemcpyAsync(d + sent, p->pinned_stage[slot], len, cudaMemcpyHostToDevice, p->copy_stream));
                CK(cudaEventRecord(p->ev_copy[slot], p->copy_stream));
                slot ^= 1;
            }
            sent += len;
        }
        return BSQ_OK;
    }
    bsq_status ready_upto(uint64_t upto) override {
        bsq_status st = send_upto(upto);
        if (st != BSQ_OK) return st;
        // the scan stream may read the bytes once the copies enqueued so far have landed ...
        CK(cudaEventRecord(p->ev[5], p->copy_stream));
        CK(cudaStreamWaitEvent(p->stream, p->ev[5], 0));
        // ... and the next window's bytes travel while this one is scanned
        return send_upto(upto + ahead);
    }
};

}  // namespace

static bsq_status trim_to_whole_batches(bsq_parser* p, uint32_t want, bool is_last, uint32_t m, bsq_pass_result* out);

// BSQ_WANT_WHOLE_BATCHES: a region that does not end the stream keeps its trailing partial batch unconsumed (the cut
// between two records needs the offsets table, so such a pass also fills it)
static inline uint32_t pass_want(uint32_t want, int32_t is_last) {
    uint32_t w = want & (BSQ_WANT_OFFSETS | BSQ_WANT_BATCHES);
    if ((want & BSQ_WANT_WHOLE_BATCHES) && (want & BSQ_WANT_BATCHES) && !is_last) w |= BSQ_WANT_OFFSETS;
    return w;
}

extern "C" bsq_status bsq_parse_device(bsq_parser* p, const uint8_t* dev_bytes, uint64_t n, int64_t stream_offset,
                                       int64_t first_record, int32_t is_last, uint32_t want, bsq_pass_result* out) {
    if (!p || (!dev_bytes && n)) return BSQ_E_ARG;
    CK(cudaSetDevice(p->cfg.device_id));
    InputFeed none;
    bsq_status st = run_pass(p, dev_bytes, n, stream_offset, first_record, is_last, pass_want(want, is_last), kWindowMax, none, out);
    if (st == BSQ_OK && (want & BSQ_WANT_WHOLE_BATCHES)) st = trim_to_whole_batches(p, want, is_last != 0, (uint32_t)p->cfg.batch_size, out);
    return st;
}

extern "C" bsq_status bsq_parse_host(bsq_parser* p, const uint8_t* host_bytes, uint64_t n, int64_t stream_offset,
                                     int64_t first_record, int32_t is_last, uint32_t want, bsq_pass_result* out) {
    if (!p || (!host_bytes && n)) return BSQ_E_ARG;
    CK(cudaSetDevice(p->cfg.device_id));
    CK(p->host_input.ensure(n + 256, 1 << 20));
    HostFeed feed;
    feed.p = p; feed.h = host_bytes; feed.d = p->host_input.as<uint8_t>(); feed.n = n;
    cudaPointerAttributes attr{};
    if (n && cudaPointerGetAttributes(&attr, host_bytes) == cudaSuccess && attr.type == cudaMemoryTypeHost) feed.pinned = true;
    cudaGetLastError();
    if (!feed.pinned && n) {
        const size_t chunk = (size_t)p->cfg.h2d_chunk_bytes;
        if (p->pinned_stage_bytes < chunk) {
            for (auto& s : p->pinned_stage) { if (s) cudaFreeHost(s); s = nullptr; }
            for (auto& s : p->pinned_stage) CK(cudaHostAlloc(&s, chunk, cudaHostAllocDefault));
            p->pinned_stage_bytes = chunk;
        }
    }
    const uint64_t window = std::min<uint64_t>(kWindowMax, std::max<uint64_t>(kHostWindow, (uint64_t)p->cfg.h2d_chunk_bytes));
    feed.ahead = window;
    bsq_status st = run_pass(p, feed.d, n, stream_offset, first_record, is_last, pass_want(want, is_last), window, feed, out);
    if (st == BSQ_OK && (want & BSQ_WANT_WHOLE_BATCHES)) st = trim_to_whole_batches(p, want, is_last != 0, (uint32_t)p->cfg.batch_size, out);
    return st;
}

// ------------------------------------------------------------------------------------------------
// streaming from a file: reader thread -> pinned regions -> passes
// ------------------------------------------------------------------------------------------------

// the inflate kernel of the build: every lane decodes (one warp per member), or the leader-lane form for the tuning
// builds with several members per warp (BSQ_INF_LANES < 32)
#if BSQ_INF_UNIFORM && BSQ_INF_LANES == 32
static constexpr auto kInflateKernel = k_inflate_members_uniform;
#else
static constexpr auto kInflateKernel = k_inflate_members;
#endif

struct bsq_stream {
    bsq_parser* p = nullptr;
    int kind = BSQ_SOURCE_PLAIN;
    FILE* fp = nullptr;
    gzFile gz = nullptr;
    std::unique_ptr<bsq_pgz::Reader> pgz;   // ordinary gzip, decoded by host threads (bsq_pgzip.h)
    uint64_t region_bytes = 0, carry_cap = 0;
    struct Buf { uint8_t* mem = nullptr; uint64_t n_new = 0; bool eof = false; int state = 0; /* 0 free, 1 ready, 2 in use */ };
    static constexpr int kMaxSlots = 3;
    Buf buf[kMaxSlots];
    int n_slots = 2;                 // pinned buffers the reader cycles through (3 when regions are inflated on the device)
    std::thread reader;
    std::mutex mu;
    std::condition_variable cv;
    bool stop = false, read_error = false;
    int next_fill = 0, next_take = 0;
    // caller side
    int cur = -1;                    // buffer of the region last parsed
    const uint8_t* region_ptr = nullptr;
    uint64_t region_n = 0;
    std::vector<uint8_t> carry;      // unconsumed tail of the previous region
    std::vector<uint8_t> big;        // a region whose carry did not fit in front of the pinned buffer
    int64_t stream_pos = 0;          // stream offset of carry[0]
    int64_t records_done = 0;
    bool finished = false;
    bsq_stream_stats st{};

    // ---- BGZF (blocked gzip, SAM spec 4.1): every member is <= 64 KiB, carries its compressed size in
    // the 'BC' extra field and its uncompressed size in ISIZE, so members inflate independently.  The
    // reader thread walks the member headers and `inflate_threads` workers inflate a region's members in
    // parallel, each straight into its place in the pinned region (the parallel-decoder role of
    // RapidgzipReader(parallelism), readers.mojo:380-443; ordinary gzip goes through the speculative decoder of bsq_pgzip.h).
    bool bgzf = false;
    int inflate_threads = 1, io_threads = 1;
    int64_t file_size = -1, file_pos = 0;   // regular plain files: parallel pread
    FILE* zfp = nullptr;                 // the compressed file, read raw
    std::vector<uint8_t> zbuf;           // compressed members of the region being filled
    std::vector<uint8_t> zpend;          // a member read for the previous region that did not fit it
    struct Member { size_t coff; uint32_t clen; uint64_t ooff; uint32_t isize; };
    std::vector<Member> members;

    // ---- BGZF inflated on the device (cfg.gpu_inflate): the reader thread only READS -- it fills a pinned buffer
    // with whole compressed members (parallel pread + a walk over the member headers); bsq_stream_next sends the
    // compressed bytes over PCIe and k_inflate_members writes the region straight into HBM, in front of which the
    // previous region's unconsumed tail is copied device to device.  The inflated bytes visit the host only when
    // the caller asks for them (bsq_stream_region).
    bool gpu_inflate = false;
    struct ZBuf { uint8_t* mem = nullptr; uint64_t n = 0; std::vector<bsq::InflateMember> members; uint64_t out_bytes = 0; };
    ZBuf zb[kMaxSlots];
    uint64_t zcap = 0;                   // compressed bytes a pinned buffer holds
    int regions_filled = 0;
    std::vector<uint8_t> zcarry;         // compressed bytes read but not yet handed out (a partial member, or over budget)
    int64_t zfile_pos = 0;
    // Two regions are in flight: while the pass of region k runs on the parser's stream, the compressed bytes of region
    // k+1 travel and are inflated on the copy stream (when the reader has them ready).
    struct InfJob {
        bool launched = false, eof = false;
        int slot = -1;
        uint32_t nm = 0;
        uint64_t out_bytes = 0, z_bytes = 0, room = 0;   // room: bytes kept free in front of the inflated data for the carry
        DevBuf zdev, mdev, sdev;
        uint32_t* status = nullptr;      // pinned: the D2H copy must not hold the host up
        size_t status_cap = 0;
        cudaEvent_t e_start = nullptr, e_h2d = nullptr, e_k0 = nullptr, e_done = nullptr;
    };
    cudaStream_t h2d_stream = nullptr;   // compressed bytes of region k+1 travel while region k is still inflating
    InfJob job[2];
    int jcur = 0;
    cudaEvent_t e_carry = nullptr;
    DevBuf rdev[2];                      // inflated regions (ping-pong): [room | inflated members]
    int rcur = 0;
    uint64_t region_dev_off = 0;         // offset of the current region in rdev[rcur]
    uint64_t dev_carry_off = 0, dev_carry_len = 0;   // unconsumed tail of the region in rdev[rcur] (offset in the buffer)
    std::vector<uint8_t> region_host;    // the current region on the host, fetched on demand
    bool region_host_valid = false;
    uint64_t region_dev_n = 0;

    // fills zb[which] with whole members: at most region_bytes of output, at most zcap compressed bytes
    bool fill_bgzf_raw(ZBuf& z, bool* eof) {
        z.members.clear(); z.n = 0; z.out_bytes = 0;
        // the first regions are short (1/8, 1/4, 1/2 of a region): the read -> copy -> inflate -> parse pipeline has nothing to
        // overlap with until its first stages have run once, so it is filled with small pieces
        const int shift = regions_filled < 3 ? 3 - regions_filled : 0;
        ++regions_filled;
        const uint64_t out_limit = std::max<uint64_t>(region_bytes >> shift, std::min<uint64_t>(region_bytes, 1ull << 20));
        const uint64_t z_limit = std::max<uint64_t>(zcap >> shift, std::min<uint64_t>(zcap, 1ull << 20));
        uint64_t have = zcarry.size();
        if (have) memcpy(z.mem, zcarry.data(), have);
        zcarry.clear();
        // parallel pread of the next compressed bytes
        const uint64_t want = have >= z_limit ? 0 : (uint64_t)std::min<int64_t>((int64_t)(z_limit - have), file_size - zfile_pos);
        if (want > 0) {
            const int nt = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)io_threads, want >> 22));
            const uint64_t slice = (want + (uint64_t)nt - 1) / (uint64_t)nt;
            std::atomic<bool> bad{false};
            auto work = [&](int t) {
                uint64_t a = std::min<uint64_t>(want, slice * (uint64_t)t), b2 = std::min<uint64_t>(want, a + slice);
                while (a < b2) {
                    const ssize_t k = pread(fileno(zfp), z.mem + have + a, (size_t)std::min<uint64_t>(b2 - a, 1u << 30), (off_t)(zfile_pos + (int64_t)a));
                    if (k <= 0) { bad = true; return; }
                    a += (uint64_t)k;
                }
            };
            std::vector<std::thread> pool;
            for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
            work(0);
            for (auto& th : pool) th.join();
            if (bad.load()) return false;
            zfile_pos += (int64_t)want;
            have += want;
        }
        // walk the member headers
        uint64_t pos = 0;
        while (pos + 18 <= have) {
            const uint8_t* h = z.mem + pos;
            if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4) || h[10] != 6 || h[11] != 0 || h[12] != 'B' || h[13] != 'C' ||
                h[14] != 2 || h[15] != 0)
                return false;
            const uint32_t total = ((uint32_t)h[16] | ((uint32_t)h[17] << 8)) + 1u;
            if (total < 18u + 8u) return false;
            if (pos + total > have) break;                              // the member continues in bytes not read yet
            const uint8_t* t = h + total - 8;
            const uint32_t crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
            const uint32_t isz = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
            if (isz > (1u << 16)) return false;
            if (z.out_bytes + isz > out_limit && !z.members.empty()) break;      // belongs to the next region
            if (isz > 0) z.members.push_back(bsq::InflateMember{pos + 18, total - 18u - 8u, isz, z.out_bytes, crc, 0u});
            z.out_bytes += isz;
            pos += total;
        }
        if (pos == 0 && have > 0 && (have >= z_limit || zfile_pos >= file_size)) return false;   // a member that never completes
        z.n = pos;
        zcarry.assign(z.mem + pos, z.mem + have);
        *eof = zfile_pos >= file_size && zcarry.empty();
        return true;
    }

    // next member (header + payload) appended to `to`; returns 1 ok, 0 clean EOF, -1 malformed
    int read_member(std::vector<uint8_t>& to, uint32_t* clen, uint32_t* isize) {
        uint8_t h[18];
        const size_t k = fread(h, 1, 18, zfp);
        if (k == 0) return 0;
        if (k != 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4) || h[10] != 6 || h[11] != 0 ||
            h[12] != 'B' || h[13] != 'C' || h[14] != 2 || h[15] != 0)
            return -1;
        const uint32_t total = ((uint32_t)h[16] | ((uint32_t)h[17] << 8)) + 1u;
        if (total < 18u + 8u) return -1;
        const size_t at = to.size();
        to.resize(at + total);
        memcpy(to.data() + at, h, 18);
        if (fread(to.data() + at + 18, 1, total - 18u, zfp) != total - 18u) return -1;
        const uint8_t* t = to.data() + at + total - 4;
        *isize = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        *clen = total;
        return 1;
    }
    // fills dst[0, region_bytes) with whole members; false on a malformed / corrupt member
    bool fill_bgzf(uint8_t* dst, uint64_t* got, bool* eof) {
        zbuf.clear(); members.clear();
        uint64_t out = 0;
        if (!zpend.empty()) {
            const uint8_t* t = zpend.data() + zpend.size() - 4;
            const uint32_t isz = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
            if (isz > region_bytes) return false;
            zbuf = zpend; zpend.clear();
            members.push_back(Member{0, (uint32_t)zbuf.size(), 0, isz});
            out = isz;
        }
        for (;;) {
            const size_t at = zbuf.size();
            uint32_t clen = 0, isz = 0;
            const int r = read_member(zbuf, &clen, &isz);
            if (r == 0) { *eof = true; break; }
            if (r < 0 || isz > (1u << 16)) return false;
            if (out + isz > region_bytes) {            // belongs to the next region
                zpend.assign(zbuf.begin() + (ptrdiff_t)at, zbuf.end());
                zbuf.resize(at);
                break;
            }
            members.push_back(Member{at, clen, out, isz});
            out += isz;
        }
        std::atomic<size_t> next{0};
        std::atomic<bool> bad{false};
        auto work = [&]() {
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) { bad = true; return; }
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= members.size() || bad.load()) break;
                const Member& m = members[i];
                const uint8_t* c = zbuf.data() + m.coff;
                inflateReset(&zs);
                zs.next_in = const_cast<Bytef*>(c + 18); zs.avail_in = m.clen - 18u - 8u;
                zs.next_out = dst + m.ooff; zs.avail_out = m.isize;
                const int rc = m.isize ? inflate(&zs, Z_FINISH) : (inflate(&zs, Z_FINISH), Z_STREAM_END);
                const uint8_t* t = c + m.clen - 8;
                const uint32_t want_crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
                if (rc != Z_STREAM_END || zs.total_out != m.isize ||
                    (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst + m.ooff, m.isize) != want_crc)
                    bad = true;
            }
            inflateEnd(&zs);
        };
        const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)inflate_threads, members.size() / 4 + 1));
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; ++t) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
        *got = out;
        return !bad.load();
    }

    void reader_main() {
        for (;;) {
            Buf* b;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || buf[next_fill].state == 0; });
                if (stop) return;
                b = &buf[next_fill];
            }
            const auto t0 = std::chrono::steady_clock::now();
            uint64_t got = 0;
            bool eof = false, err = false;
            uint8_t* dst = b->mem ? b->mem + carry_cap : nullptr;
            if (gpu_inflate) {
                ZBuf& z = zb[next_fill];
                if (!fill_bgzf_raw(z, &eof)) err = true;
                got = z.out_bytes;
            } else if (bgzf) {
                if (!fill_bgzf(dst, &got, &eof)) err = true;
            } else if (kind == BSQ_SOURCE_PLAIN && file_size >= 0) {
                // regular file: the region is read as `io_threads` slices with pread (one memcpy-bound
                // thread tops out near 5 GB/s from the page cache)
                const uint64_t want = std::min<uint64_t>(region_bytes, (uint64_t)(file_size - file_pos));
                const int nt = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)io_threads, want >> 22));
                const uint64_t slice = (want + (uint64_t)nt - 1) / (uint64_t)nt;
                std::atomic<bool> bad{false};
                auto work = [&](int t) {
                    uint64_t a = std::min<uint64_t>(want, slice * (uint64_t)t), b2 = std::min<uint64_t>(want, a + slice);
                    while (a < b2) {
                        const ssize_t k = pread(fileno(fp), dst + a, (size_t)std::min<uint64_t>(b2 - a, 1u << 30), (off_t)(file_pos + (int64_t)a));
                        if (k <= 0) { bad = true; return; }
                        a += (uint64_t)k;
                    }
                };
                std::vector<std::thread> pool;
                for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
                work(0);
                for (auto& th : pool) th.join();
                if (bad.load()) err = true;
                got = want;
                file_pos += (int64_t)want;
                eof = file_pos >= file_size;
            } else
            while (got < region_bytes) {
                const size_t ask = (size_t)std::min<uint64_t>(region_bytes - got, 1u << 30);
                long k;
                if (pgz) k = (long)pgz->read(dst + got, ask);
                else if (kind == BSQ_SOURCE_GZIP) k = gzread(gz, dst + got, (unsigned)std::min<size_t>(ask, 1u << 30));
                else k = (long)fread(dst + got, 1, ask, fp);
                if (k < 0) { err = true; break; }
                if (k == 0) { eof = true; break; }
                got += (uint64_t)k;
            }
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            {
                std::lock_guard<std::mutex> lk(mu);
                b->n_new = got; b->eof = eof || err; b->state = 1;
                if (err) read_error = true;
                st.reader_busy_s += dt; st.bytes_read += got;
                next_fill = (next_fill + 1) % n_slots;
            }
            cv.notify_all();
            if (eof || err) return;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// RapidgzipReader: the parallel gzip decoder on its own (host only)
// ------------------------------------------------------------------------------------------------

struct bsq_gzip { bsq_pgz::Reader r; std::string err; };

extern "C" bsq_status bsq_gzip_open(const char* path, int32_t parallelism, uint64_t chunk_bytes, bsq_gzip** out) {
    if (!path || !out) return BSQ_E_ARG;
    *out = nullptr;
    bsq_gzip* g = new (std::nothrow) bsq_gzip();
    if (!g) return BSQ_E_NOMEM;
    if (!g->r.open(path, parallelism, chunk_bytes ? (size_t)chunk_bytes : (2u << 20))) { delete g; return BSQ_E_IO; }
    *out = g;
    return BSQ_OK;
}

extern "C" bsq_status bsq_gzip_read(bsq_gzip* g, uint8_t* dst, uint64_t n, uint64_t* got) {
    if (!g || (!dst && n) || !got) return BSQ_E_ARG;
    const int64_t k = g->r.read(dst, (size_t)n);
    if (k < 0) { *got = 0; return BSQ_E_IO; }
    *got = (uint64_t)k;
    return BSQ_OK;
}

extern "C" const char* bsq_gzip_error(const bsq_gzip* g) { return g ? g->r.error().c_str() : ""; }

extern "C" void bsq_gzip_close(bsq_gzip* g) { delete g; }

extern "C" bsq_status bsq_stream_open(bsq_parser* p, const char* path, int32_t source_kind, uint64_t region_bytes,
                                      bsq_stream** out) {
    if (!p || !path || !out) return BSQ_E_ARG;
    *out = nullptr;
    CK(cudaSetDevice(p->cfg.device_id));
    bsq_stream* s = new (std::nothrow) bsq_stream();
    if (!s) return BSQ_E_NOMEM;
    s->p = p;
    if (source_kind == BSQ_SOURCE_AUTO) {
        const size_t n = strlen(path);
        auto ends = [&](const char* suf) { const size_t k = strlen(suf); return n >= k && strcmp(path + n - k, suf) == 0; };
        source_kind = (ends(".gz") || ends(".bgz")) ? BSQ_SOURCE_GZIP : BSQ_SOURCE_PLAIN;
    }
    s->kind = source_kind;
    if (source_kind == BSQ_SOURCE_GZIP) {
        // BGZF?  (first member: FEXTRA with the 'BC' subfield in the canonical position)
        FILE* f = fopen(path, "rb");
        uint8_t h[18];
        if (f && fread(h, 1, 18, f) == 18 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && h[10] == 6 &&
            h[11] == 0 && h[12] == 'B' && h[13] == 'C' && h[14] == 2 && h[15] == 0) {
            rewind(f);
            s->bgzf = true;
            s->zfp = f;
            int nt = p->cfg.inflate_threads;
            if (nt <= 0) nt = (int)std::max(1u, std::thread::hardware_concurrency());
            s->inflate_threads = std::min(nt, 64);
            struct stat sb;
            if (!p->cfg.host_inflate && fstat(fileno(f), &sb) == 0 && S_ISREG(sb.st_mode)) {
                s->gpu_inflate = true;           // the members cross PCIe compressed and are inflated by k_inflate_members
                s->file_size = (int64_t)sb.st_size;
                s->io_threads = std::min(nt, 8);
            }
        } else {
            // an ordinary gzip stream: `inflate_threads` host threads decode it speculatively in parallel
            // (RapidgzipReader(parallelism), readers.mojo:380-443); one thread = zlib's gzread
            const bool is_gzip = f && h[0] == 0x1f && h[1] == 0x8b;
            if (f) fclose(f);
            if (is_gzip && p->cfg.inflate_threads != 1) {
                s->pgz.reset(new bsq_pgz::Reader());
                if (!s->pgz->open(path, p->cfg.inflate_threads)) s->pgz.reset();
            }
            if (!s->pgz) {
                s->gz = gzopen(path, "rb");
                if (s->gz) gzbuffer(s->gz, 1 << 20);
            }
        }
    } else {
        s->fp = fopen(path, "rb");
        struct stat sb;
        if (s->fp && fstat(fileno(s->fp), &sb) == 0 && S_ISREG(sb.st_mode)) {
            s->file_size = (int64_t)sb.st_size;
            int nt = p->cfg.inflate_threads;
            if (nt <= 0) nt = (int)std::max(1u, std::thread::hardware_concurrency());
            s->io_threads = std::min(nt, 8);
        }
    }
    if (!s->gz && !s->fp && !s->zfp && !s->pgz) { p->last_error = std::string("cannot open ") + path; delete s; return BSQ_E_ARG; }
    s->region_bytes = region_bytes ? region_bytes : (256ull << 20);
    // room in front of every pinned region for the previous region's unconsumed tail (a larger tail takes the
    // `big` path of bsq_stream_next)
    s->carry_cap = std::max<uint64_t>(std::min<uint64_t>(s->region_bytes / 4, 64ull << 20), 4096);
    if (s->gpu_inflate) {
        s->zcap = std::max<uint64_t>(s->region_bytes / 2 + (1ull << 20), 256ull << 10);
        s->n_slots = bsq_stream::kMaxSlots;
        for (auto& j : s->job)
            if (cudaEventCreate(&j.e_start) != cudaSuccess || cudaEventCreate(&j.e_h2d) != cudaSuccess ||
                cudaEventCreate(&j.e_k0) != cudaSuccess ||
                cudaEventCreate(&j.e_done) != cudaSuccess) { bsq_stream_close(s); return fail_cuda(p, cudaGetLastError(), "cudaEventCreate"); }
        if (cudaStreamCreateWithFlags(&s->h2d_stream, cudaStreamNonBlocking) != cudaSuccess) { bsq_stream_close(s); return fail_cuda(p, cudaGetLastError(), "cudaStreamCreate"); }
        if (cudaEventCreateWithFlags(&s->e_carry, cudaEventDisableTiming) != cudaSuccess) { bsq_stream_close(s); return fail_cuda(p, cudaGetLastError(), "cudaEventCreate"); }
        for (auto& z : s->zb) {
            cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&z.mem), s->zcap + 64, cudaHostAllocDefault);
            if (e != cudaSuccess) { bsq_stream_close(s); return fail_cuda(p, e, "cudaHostAlloc(compressed region)"); }
        }
        for (int i = 0; i < 2; ++i) {            // the buffers of the parser's previous device-inflating stream
            s->job[i].zdev = p->inf_zdev[i]; s->job[i].mdev = p->inf_mdev[i]; s->job[i].sdev = p->inf_sdev[i]; s->rdev[i] = p->inf_rdev[i];
            s->job[i].status = p->inf_status[i]; s->job[i].status_cap = p->inf_status_cap[i];
            p->inf_zdev[i] = DevBuf(); p->inf_mdev[i] = DevBuf(); p->inf_sdev[i] = DevBuf(); p->inf_rdev[i] = DevBuf();
            p->inf_status[i] = nullptr; p->inf_status_cap[i] = 0;
        }
        cudaError_t e = opt_in_smem(kInflateKernel, sizeof(InflateTables) * kInfPerCta);
        if (e != cudaSuccess) { bsq_stream_close(s); return fail_cuda(p, e, "k_inflate_members shared memory"); }
    } else
    for (auto& b : s->buf) {
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&b.mem), s->carry_cap + s->region_bytes + 64, cudaHostAllocDefault);
        if (e != cudaSuccess) { bsq_stream_close(s); return fail_cuda(p, e, "cudaHostAlloc(stream region)"); }
    }
    s->reader = std::thread([s] { s->reader_main(); });
    *out = s;
    return BSQ_OK;
}

extern "C" void bsq_stream_close(bsq_stream* s) {
    if (!s) return;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->stop = true;
    }
    s->cv.notify_all();
    if (s->reader.joinable()) s->reader.join();
    if (s->gz) gzclose(s->gz);
    if (s->fp) fclose(s->fp);
    if (s->zfp) fclose(s->zfp);
    for (auto& b : s->buf) if (b.mem) cudaFreeHost(b.mem);
    for (auto& z : s->zb) if (z.mem) cudaFreeHost(z.mem);
    if (s->p) {
        cudaSetDevice(s->p->cfg.device_id);
        if (s->h2d_stream) cudaStreamSynchronize(s->h2d_stream);
        if (s->p->copy_stream) cudaStreamSynchronize(s->p->copy_stream);   // a prefetched region may still be inflating
        if (s->p->stream) cudaStreamSynchronize(s->p->stream);
    }
    for (int i = 0; i < 2; ++i) {
        auto& j = s->job[i];
        if (s->p && s->gpu_inflate) {          // kept for the parser's next stream
            bsq_parser* p = s->p;
            p->inf_zdev[i].release(); p->inf_mdev[i].release(); p->inf_sdev[i].release(); p->inf_rdev[i].release();
            if (p->inf_status[i]) cudaFreeHost(p->inf_status[i]);
            p->inf_zdev[i] = j.zdev; p->inf_mdev[i] = j.mdev; p->inf_sdev[i] = j.sdev; p->inf_rdev[i] = s->rdev[i];
            p->inf_status[i] = j.status; p->inf_status_cap[i] = j.status_cap;
            j.zdev = DevBuf(); j.mdev = DevBuf(); j.sdev = DevBuf(); s->rdev[i] = DevBuf(); j.status = nullptr; j.status_cap = 0;
        }
        j.zdev.release(); j.mdev.release(); j.sdev.release();
        if (j.status) cudaFreeHost(j.status);
        if (j.e_start) cudaEventDestroy(j.e_start);
        if (j.e_h2d) cudaEventDestroy(j.e_h2d);
        if (j.e_k0) cudaEventDestroy(j.e_k0);
        if (j.e_done) cudaEventDestroy(j.e_done);
    }
    if (s->e_carry) cudaEventDestroy(s->e_carry);
    if (s->h2d_stream) cudaStreamDestroy(s->h2d_stream);
    s->rdev[0].release(); s->rdev[1].release();
    delete s;
}

// keep batches whole across regions: the trailing partial batch (all of the region's records when it holds fewer than one
// batch) is left unconsumed, to be re-presented with the next region
static bsq_status trim_to_whole_batches(bsq_parser* p, uint32_t want, bool is_last, uint32_t m, bsq_pass_result* out) {
    if (!is_last && out->stop.code == BSQ_OK && (want & BSQ_WANT_BATCHES) && out->n_records % m != 0) {
        const int64_t keep = out->n_records - out->n_records % m;
        int64_t cut = 0;
        if (keep > 0) {
            int wi = 0;
            while (wi + 1 < p->res.n_windows && keep >= p->win[wi + 1].rec_base) ++wi;
            uint32_t le = 0;
            CK(cudaMemcpy(&le, p->win[wi].line_ends.as<uint32_t>() + 4ull * (keep - p->win[wi].rec_base), 4, cudaMemcpyDeviceToHost));
            cut = (int64_t)p->win[wi].region_off - (int64_t)p->win[wi].wp.begin + (int64_t)(le + 1u);
        }
        out->n_records = keep;
        out->n_batches = keep / m;
        out->bytes_consumed = cut;
        out->n_bases = -1;
        p->res.n_records = keep; p->res.n_batches = keep / m; p->res.bytes_consumed = cut;
    }
    return BSQ_OK;
}

// ---- BGZF regions inflated on the device, two in flight ----------------------------------------------------------
// H2D of the compressed members + k_inflate_members + k_crc32_members of one region, on the copy stream
static bsq_status launch_inflate(bsq_stream* s, bsq_stream::InfJob& J, int slot, int rbuf) {
    bsq_parser* p = s->p;
    bsq_stream::ZBuf& z = s->zb[slot];
    J.slot = slot; J.eof = s->buf[slot].eof; J.nm = (uint32_t)z.members.size(); J.out_bytes = z.out_bytes; J.z_bytes = z.n;
    // room for the carry of the region before it: generous (device memory is not the constraint), grown with what was seen
    J.room = std::max<uint64_t>(32ull << 20, 2 * s->dev_carry_len + (1ull << 20));
    J.room = (J.room + 255) & ~255ull;
    CK(s->rdev[rbuf].ensure(J.room + J.out_bytes + 256, 1 << 20));
    // the compressed bytes travel on their own stream (the job's buffers were last read by the inflate two regions ago, whose
    // end the caller has waited for), so the copy runs while the previous region is still inflating on the copy stream
    CK(cudaEventRecord(J.e_start, s->h2d_stream));
    if (J.nm) {
        CK(J.zdev.ensure(z.n + 64, 1 << 20));
        CK(J.mdev.ensure(sizeof(InflateMember) * J.nm, 1 << 16));
        CK(J.sdev.ensure(4ull * J.nm, 1 << 12));
        CK(cudaMemcpyAsync(J.zdev.p, z.mem, z.n, cudaMemcpyHostToDevice, s->h2d_stream));
        CK(cudaMemcpyAsync(J.mdev.p, z.members.data(), sizeof(InflateMember) * J.nm, cudaMemcpyHostToDevice, s->h2d_stream));
    }
    CK(cudaEventRecord(J.e_h2d, s->h2d_stream));
    CK(cudaStreamWaitEvent(p->copy_stream, J.e_h2d, 0));
    CK(cudaEventRecord(J.e_k0, p->copy_stream));
    if (J.nm) {
        uint8_t* dst = s->rdev[rbuf].as<uint8_t>() + J.room;
        kInflateKernel<<<(J.nm + kInfPerCta - 1) / kInfPerCta, kInfWarps * 32, sizeof(InflateTables) * kInfPerCta, p->copy_stream>>>(
            J.zdev.as<uint8_t>(), dst, J.mdev.as<InflateMember>(), J.nm, J.sdev.as<uint32_t>());
        k_crc32_members<<<(J.nm + kCrcWarps - 1) / kCrcWarps, kCrcWarps * 32, 0, p->copy_stream>>>(
            dst, J.mdev.as<InflateMember>(), J.nm, J.sdev.as<uint32_t>());
        CK(cudaGetLastError());
        if (J.nm > J.status_cap) {
            if (J.status) cudaFreeHost(J.status);
            J.status = nullptr; J.status_cap = 0;
            CK(cudaHostAlloc(reinterpret_cast<void**>(&J.status), 4ull * (J.nm + 1024), cudaHostAllocDefault));
            J.status_cap = J.nm + 1024;
        }
        CK(cudaMemcpyAsync(J.status, J.sdev.p, 4ull * J.nm, cudaMemcpyDeviceToHost, p->copy_stream));
    }
    CK(cudaEventRecord(J.e_done, p->copy_stream));
    J.launched = true;
    return BSQ_OK;
}

static bsq_status stream_next_device_inflate(bsq_stream* s, uint32_t want, bsq_pass_result* out) {
    bsq_parser* p = s->p;
    bsq_stream::InfJob& J = s->job[s->jcur];
    bsq_stream::InfJob& N = s->job[s->jcur ^ 1];
    const int rb = s->jcur, rb_prev = s->jcur ^ 1;           // region buffers: this region / the previous (and the next) one
    const auto t0 = std::chrono::steady_clock::now();
    if (!J.launched) {                                       // the first region, or the reader was not ahead
        int slot;
        {
            std::unique_lock<std::mutex> lk(s->mu);
            s->cv.wait(lk, [&] { return s->buf[s->next_take].state == 1; });
            slot = s->next_take;
            s->buf[slot].state = 2;
            s->next_take = (s->next_take + 1) % s->n_slots;
        }
        if (s->read_error) { p->last_error = "read error / malformed BGZF member"; s->finished = true; return BSQ_E_IO; }
        bsq_status st = launch_inflate(s, J, slot, rb);
        if (st != BSQ_OK) { s->finished = true; return st; }
    }
    const auto t1 = std::chrono::steady_clock::now();
    s->st.wait_reader_s += std::chrono::duration<double>(t1 - t0).count();
    s->cur = J.slot;
    // the unconsumed tail of the previous region goes in front of this one, device to device
    uint64_t region_off = J.room - s->dev_carry_len;
    if (s->dev_carry_len > J.room) {
        // (a carry larger than the room left for it: wait for this region, then rebuild it in a larger buffer)
        CK(cudaEventSynchronize(J.e_done));
        DevBuf big;
        CK(big.ensure(s->dev_carry_len + J.out_bytes + 512, 1 << 20));
        CK(cudaMemcpyAsync(big.as<uint8_t>() + s->dev_carry_len, s->rdev[rb].as<uint8_t>() + J.room, J.out_bytes, cudaMemcpyDeviceToDevice, p->stream));
        CK(cudaStreamSynchronize(p->stream));
        std::swap(big, s->rdev[rb]);
        big.release();
        J.room = s->dev_carry_len;
        region_off = 0;
    }
    if (s->dev_carry_len)
        CK(cudaMemcpyAsync(s->rdev[rb].as<uint8_t>() + region_off, s->rdev[rb_prev].as<uint8_t>() + s->dev_carry_off, s->dev_carry_len,
                           cudaMemcpyDeviceToDevice, p->stream));
    CK(cudaEventRecord(s->e_carry, p->stream));
    // the next region, if the reader has it: its compressed bytes travel and inflate while this region is parsed
    if (!J.eof && !N.launched) {
        int slot = -1;
        {
            std::lock_guard<std::mutex> lk(s->mu);
            if (s->buf[s->next_take].state == 1 && !s->read_error) {
                slot = s->next_take;
                s->buf[slot].state = 2;
                s->next_take = (s->next_take + 1) % s->n_slots;
            }
        }
        if (slot >= 0) {
            CK(cudaStreamWaitEvent(p->copy_stream, s->e_carry, 0));   // it overwrites the buffer the carry was just read from
            bsq_status st = launch_inflate(s, N, slot, rb_prev);
            if (st != BSQ_OK) { s->finished = true; return st; }
        }
    }
    // this region: inflated?
    const auto tw0 = std::chrono::steady_clock::now();
    s->st.launch_s += std::chrono::duration<double>(tw0 - t1).count();
    CK(cudaEventSynchronize(J.e_done));
    s->st.wait_inflate_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - tw0).count();
    float ms_copy = 0.f, ms_inf = 0.f;
    cudaEventElapsedTime(&ms_copy, J.e_start, J.e_h2d);
    cudaEventElapsedTime(&ms_inf, J.e_k0, J.e_done);
    s->st.h2d_s += ms_copy * 1e-3; s->st.inflate_s += ms_inf * 1e-3; s->st.compressed_bytes += J.z_bytes;
    for (uint32_t i = 0; i < J.nm; ++i)
        if (J.status[i] != 0u) {
            char t[128];
            snprintf(t, sizeof t, "BGZF member %u of the region does not inflate (status %u)", i, J.status[i]);
            p->last_error = t; s->finished = true;
            return BSQ_E_IO;
        }
    const uint64_t n = s->dev_carry_len + J.out_bytes;
    uint8_t* region = s->rdev[rb].as<uint8_t>() + region_off;
    const bool is_last = J.eof;
    const uint32_t m = (uint32_t)p->cfg.batch_size;
    uint32_t w = want;
    if (!is_last && (want & BSQ_WANT_BATCHES)) w |= BSQ_WANT_OFFSETS;   // the cut between regions needs offsets
    InputFeed none;
    bsq_status rc = run_pass(p, region, n, s->stream_pos, s->records_done, is_last ? 1 : 0, w, kWindowMax, none, out);
    if (rc != BSQ_OK) { s->finished = true; return rc; }
    rc = trim_to_whole_batches(p, want, is_last, m, out);
    if (rc != BSQ_OK) { s->finished = true; return rc; }
    s->rcur = rb; s->region_dev_off = region_off;
    s->region_dev_n = n; s->region_host_valid = false;
    s->region_ptr = nullptr; s->region_n = n;
    s->st.parse_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    s->st.regions += 1;
    const uint64_t consumed = (uint64_t)out->bytes_consumed;
    s->dev_carry_off = region_off + consumed; s->dev_carry_len = n - consumed;
    s->stream_pos += (int64_t)consumed;
    s->records_done += out->n_records;
    if (out->stop.code != BSQ_OK) s->finished = true;
    // the compressed bytes of this region are on the device: its pinned buffer goes back to the reader
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->buf[J.slot].state = 0;
    }
    s->cv.notify_all();
    J.launched = false;
    s->jcur ^= 1;
    return BSQ_OK;
}

extern "C" bsq_status bsq_stream_next(bsq_stream* s, uint32_t want, bsq_pass_result* out) {
    if (!s || !out) return BSQ_E_ARG;
    bsq_parser* p = s->p;
    CK(cudaSetDevice(p->cfg.device_id));
    if (s->finished) { p->last_error = "stream already finished"; return BSQ_E_STATE; }
    if (s->gpu_inflate) return stream_next_device_inflate(s, want, out);
    // the previous region's buffer goes back to the reader
    if (s->cur >= 0) {
        {
            std::lock_guard<std::mutex> lk(s->mu);
            s->buf[s->cur].state = 0;
        }
        s->cv.notify_all();
        s->cur = -1;
    }
    const auto t0 = std::chrono::steady_clock::now();
    bsq_stream::Buf* b;
    {
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv.wait(lk, [&] { return s->buf[s->next_take].state == 1; });
        b = &s->buf[s->next_take];
        b->state = 2;
        s->cur = s->next_take;
        s->next_take = (s->next_take + 1) % s->n_slots;
    }
    const auto t1 = std::chrono::steady_clock::now();
    s->st.wait_reader_s += std::chrono::duration<double>(t1 - t0).count();
    if (s->read_error) { p->last_error = "read / inflate error"; s->finished = true; return BSQ_E_IO; }
    uint8_t* region;
    if (s->carry.size() <= s->carry_cap) {
        region = b->mem + s->carry_cap - s->carry.size();
        if (!s->carry.empty()) memcpy(region, s->carry.data(), s->carry.size());
    } else {
        // the carry (a partial record plus, for batches, the records of a trailing partial batch: ~ batch_size x
        // record length, e.g. 4096 long reads) outgrew the area in front of the pinned region: this region is
        // assembled in a separate buffer.  A record that fits nowhere is reported by the pass itself.
        s->big.resize(s->carry.size() + b->n_new);
        memcpy(s->big.data(), s->carry.data(), s->carry.size());
        memcpy(s->big.data() + s->carry.size(), b->mem + s->carry_cap, b->n_new);
        region = s->big.data();
    }
    const uint64_t n = s->carry.size() + b->n_new;
    const bool is_last = b->eof;
    const uint32_t m = (uint32_t)p->cfg.batch_size;
    uint32_t w = want;
    if (!is_last && (want & BSQ_WANT_BATCHES)) w |= BSQ_WANT_OFFSETS;   // the cut between regions needs offsets
    bsq_status rc = bsq_parse_host(p, region, n, s->stream_pos, s->records_done, is_last ? 1 : 0, w, out);
    if (rc != BSQ_OK) { s->finished = true; return rc; }
    rc = trim_to_whole_batches(p, want, is_last, m, out);
    if (rc != BSQ_OK) { s->finished = true; return rc; }
    s->region_ptr = region; s->region_n = n;
    s->st.parse_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    s->st.regions += 1;
    // carry for the next region
    const uint64_t consumed = (uint64_t)out->bytes_consumed;
    s->carry.assign(region + consumed, region + n);
    // (stream_pos / records_done describe the NEXT region from here on; bsq_stream_region reports this one)
    const int64_t this_pos = s->stream_pos, this_first = s->records_done;
    s->stream_pos += (int64_t)consumed;
    s->records_done += out->n_records;
    if (out->stop.code != BSQ_OK) s->finished = true;
    s->carry.shrink_to_fit();
    (void)this_pos; (void)this_first;
    return BSQ_OK;
}

extern "C" void bsq_stream_region_info(const bsq_stream* s, uint64_t* n, int64_t* stream_offset, int64_t* first_record) {
    if (n) *n = (s && s->cur >= 0) ? s->region_n : 0;
    if (stream_offset) *stream_offset = (s && s->cur >= 0) ? s->p->pass_stream_offset : 0;
    if (first_record) *first_record = (s && s->cur >= 0) ? s->records_done - s->p->res.n_records : 0;
}

extern "C" const uint8_t* bsq_stream_region(const bsq_stream* cs, uint64_t* n, int64_t* stream_offset, int64_t* first_record) {
    bsq_stream* s = const_cast<bsq_stream*>(cs);
    if (!s || s->cur < 0) return nullptr;
    if (n) *n = s->region_n;
    if (stream_offset) *stream_offset = s->p->pass_stream_offset;
    if (first_record) *first_record = s->records_done - s->p->res.n_records;
    if (s->gpu_inflate) {
        // the region was inflated on the device: its bytes come to the host only on this request
        if (!s->region_host_valid) {
            s->region_host.resize(s->region_dev_n ? s->region_dev_n : 1);
            cudaSetDevice(s->p->cfg.device_id);
            if (s->region_dev_n &&
                cudaMemcpy(s->region_host.data(), s->rdev[s->rcur].as<uint8_t>() + s->region_dev_off, s->region_dev_n, cudaMemcpyDeviceToHost) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
            s->region_host_valid = true;
        }
        return s->region_host.data();
    }
    return s->region_ptr;
}

extern "C" bsq_status bsq_stream_get_stats(const bsq_stream* s, bsq_stream_stats* out) {
    if (!s || !out) return BSQ_E_ARG;
    std::lock_guard<std::mutex> lk(const_cast<bsq_stream*>(s)->mu);
    *out = s->st;
    return BSQ_OK;
}

// ------------------------------------------------------------------------------------------------
// result views
// ------------------------------------------------------------------------------------------------

extern "C" bsq_status bsq_get_offsets(const bsq_parser* p, int32_t window, bsq_offsets_view* out) {
    if (!p || !out) return BSQ_E_ARG;
    if (!p->have_pass || !(p->want & BSQ_WANT_OFFSETS)) return BSQ_E_STATE;
    if (window < 0 || window >= p->res.n_windows) return BSQ_E_ARG;
    const Window& w = p->win[window];
    const int64_t next_base = (window + 1 < p->res.n_windows) ? p->win[window + 1].rec_base : p->total_records;
    int64_t nrec = next_base - w.rec_base;
    if (w.rec_base + nrec > p->res.n_records) nrec = std::max<int64_t>(0, p->res.n_records - w.rec_base);
    out->stream_base = p->pass_stream_offset + (int64_t)w.region_off - (int64_t)w.wp.begin;
    out->first_record = w.rec_base;
    out->n_records = nrec;
    out->line_ends = w.line_ends.as<uint32_t>();
    out->id_spans = p->id_spans.as<uint32_t>() + 2 * w.rec_base;
    out->window_bytes = w.base;
    return BSQ_OK;
}

static bsq_status fill_batch(const bsq_parser* p, int64_t first, int64_t count, int64_t b0, bsq_batch_view* out) {
    memset(out, 0, sizeof *out);
    out->quality_offset = 33;  // parser.mojo:243 never passes the schema offset
    out->num_records = count;
    if (count <= 0) return BSQ_OK;
    const int32_t m = p->cfg.batch_size;
    const int64_t qb = p->h_ends_base[b0], ib = p->h_id_ends_base[b0];
    out->sequence_buffer = p->seq_out.as<uint8_t>() + qb;
    out->qual_buffer = p->qual_out.as<uint8_t>() + qb;
    out->id_buffer = p->id_out.as<uint8_t>() + ib;
    out->ends = p->ends.as<int64_t>() + first;
    out->id_ends = p->id_ends.as<int64_t>() + first;
    // sizes: a full batch ends where the next one begins
    const int64_t last = first + count;  // exclusive
    int64_t q_end, i_end;
    if (last % m == 0 || last == p->total_records) {
        const int64_t nb = (last + m - 1) / m;
        q_end = p->h_ends_base[nb]; i_end = p->h_id_ends_base[nb];
    } else {
        // a batch cut short by an error: read the cumulative values of its last record
        int64_t v[2];
        if (cudaMemcpy(&v[0], p->ends.as<int64_t>() + (last - 1), 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(&v[1], p->id_ends.as<int64_t>() + (last - 1), 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
            cudaGetLastError();
            return BSQ_E_CUDA;
        }
        q_end = qb + v[0]; i_end = ib + v[1];
    }
    out->seq_len = q_end - qb;
    out->sequence_bytes = out->seq_len;
    if (p->tail_emitted && last == p->total_records) out->sequence_bytes += p->tail_seq - p->tail_qual;
    out->total_id_bytes = i_end - ib;
    return BSQ_OK;
}

extern "C" bsq_status bsq_get_batch(const bsq_parser* p, int64_t b, bsq_batch_view* out) {
    if (!p || !out) return BSQ_E_ARG;
    if (!p->have_pass || !(p->want & BSQ_WANT_BATCHES)) return BSQ_E_STATE;
    if (b < 0 || b >= p->res.n_batches) return BSQ_E_ARG;
    const int32_t m = p->cfg.batch_size;
    const int64_t first = b * m;
    const int64_t count = std::min<int64_t>(m, p->res.n_records - first);
    return fill_batch(p, first, count, b, out);
}

extern "C" bsq_status bsq_get_soa(const bsq_parser* p, bsq_batch_view* out) {
    if (!p || !out) return BSQ_E_ARG;
    if (!p->have_pass || !(p->want & BSQ_WANT_BATCHES)) return BSQ_E_STATE;
    memset(out, 0, sizeof *out);
    out->quality_offset = 33;
    out->num_records = p->res.n_records;
    if (p->res.n_records == 0) return BSQ_OK;
    out->sequence_buffer = p->seq_out.as<uint8_t>();
    out->qual_buffer = p->qual_out.as<uint8_t>();
    out->id_buffer = p->id_out.as<uint8_t>();
    out->ends = p->ends.as<int64_t>();
    out->id_ends = p->id_ends.as<int64_t>();
    bsq_batch_view last;
    const int64_t lb = p->res.n_batches - 1;
    bsq_status st = bsq_get_batch(p, lb, &last);
    if (st != BSQ_OK) return st;
    out->seq_len = p->h_ends_base[lb] + last.seq_len;
    out->sequence_bytes = p->h_ends_base[lb] + last.sequence_bytes;
    out->total_id_bytes = p->h_id_ends_base[lb] + last.total_id_bytes;
    return BSQ_OK;
}

extern "C" bsq_status bsq_batch_to_host(bsq_parser* p, int64_t b, uint8_t* seq, uint8_t* qual, uint8_t* id,
                                        int64_t* ends, int64_t* id_ends) {
    bsq_batch_view v;
    bsq_status st = bsq_get_batch(p, b, &v);
    if (st != BSQ_OK) return st;
    if (v.num_records == 0) return BSQ_OK;
    CK(cudaSetDevice(p->cfg.device_id));
    if (seq) CK(cudaMemcpyAsync(seq, v.sequence_buffer, v.sequence_bytes, cudaMemcpyDeviceToHost, p->stream));
    if (qual) CK(cudaMemcpyAsync(qual, v.qual_buffer, v.seq_len, cudaMemcpyDeviceToHost, p->stream));
    if (id) CK(cudaMemcpyAsync(id, v.id_buffer, v.total_id_bytes, cudaMemcpyDeviceToHost, p->stream));
    if (ends) CK(cudaMemcpyAsync(ends, v.ends, 8 * v.num_records, cudaMemcpyDeviceToHost, p->stream));
    if (id_ends) CK(cudaMemcpyAsync(id_ends, v.id_ends, 8 * v.num_records, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return BSQ_OK;
}

extern "C" bsq_status bsq_soa_to_host(bsq_parser* p, uint8_t* seq, uint8_t* qual, uint8_t* id, int64_t* ends,
                                      int64_t* id_ends) {
    bsq_batch_view v;
    bsq_status st = bsq_get_soa(p, &v);
    if (st != BSQ_OK) return st;
    if (v.num_records == 0) return BSQ_OK;
    CK(cudaSetDevice(p->cfg.device_id));
    // two copy queues so that the big arrays travel back to back on the D2H engine
    if (seq) CK(cudaMemcpyAsync(seq, v.sequence_buffer, v.sequence_bytes, cudaMemcpyDeviceToHost, p->copy_stream));
    if (qual) CK(cudaMemcpyAsync(qual, v.qual_buffer, v.seq_len, cudaMemcpyDeviceToHost, p->copy_stream));
    if (id) CK(cudaMemcpyAsync(id, v.id_buffer, v.total_id_bytes, cudaMemcpyDeviceToHost, p->stream));
    if (ends) CK(cudaMemcpyAsync(ends, v.ends, 8 * v.num_records, cudaMemcpyDeviceToHost, p->stream));
    if (id_ends) CK(cudaMemcpyAsync(id_ends, v.id_ends, 8 * v.num_records, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->copy_stream));
    CK(cudaStreamSynchronize(p->stream));
    return BSQ_OK;
}

extern "C" bsq_status bsq_write_records(bsq_parser* p, int64_t first_record, int64_t count, uint8_t* out_device, uint64_t capacity,
                                        uint8_t* out_host, uint64_t* offsets_host, uint64_t* bytes_written) {
    if (!p || first_record < 0 || count < 0 || !bytes_written) return BSQ_E_ARG;
    if (!p->have_pass || !(p->want & BSQ_WANT_BATCHES)) return BSQ_E_STATE;
    if (first_record + count > p->res.n_records) return BSQ_E_ARG;
    *bytes_written = 0;
    if (count == 0) { if (offsets_host) offsets_host[0] = 0; return BSQ_OK; }
    CK(cudaSetDevice(p->cfg.device_id));
    WriteParams W{};
    W.seq = p->seq_out.as<uint8_t>(); W.qual = p->qual_out.as<uint8_t>(); W.id = p->id_out.as<uint8_t>();
    W.ends = p->ends.as<int64_t>(); W.id_ends = p->id_ends.as<int64_t>();
    W.ends_base = p->ends_base.as<int64_t>(); W.id_ends_base = p->id_ends_base.as<int64_t>();
    W.first = first_record; W.count = count; W.batch_size = p->cfg.batch_size;
    CK(p->write_offs.ensure(8ull * (size_t)(count + 1), 1 << 16));
    unsigned long long* offs = p->write_offs.as<unsigned long long>();
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)p->sm_count * 8, (count + 256) / 256));
    k_write_sizes<<<grid, 256, 0, p->stream>>>(W, offs);
    size_t tmp = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, offs, offs, (int)(count + 1), p->stream));
    CK(p->cub_tmp.ensure(tmp + 16));
    CK(cub::DeviceScan::ExclusiveSum(p->cub_tmp.p, tmp, offs, offs, (int)(count + 1), p->stream));
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, offs + count, 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    *bytes_written = total;
    if (offsets_host) CK(cudaMemcpyAsync(offsets_host, offs, 8ull * (size_t)(count + 1), cudaMemcpyDeviceToHost, p->stream));
    if (out_device || out_host) {
        uint8_t* dst = out_device;
        if (dst) { if (capacity < total) return BSQ_E_ARG; }
        else { CK(p->write_buf.ensure((size_t)total + 16, 1 << 20)); dst = p->write_buf.as<uint8_t>(); }
        const int gridw = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)p->sm_count * 8, (count + 7) / 8));
        k_write_records<<<gridw, 256, 0, p->stream>>>(W, offs, dst);
        CK(cudaGetLastError());
        if (out_host) CK(cudaMemcpyAsync(out_host, dst, (size_t)total, cudaMemcpyDeviceToHost, p->stream));
    }
    CK(cudaStreamSynchronize(p->stream));
    p->n_launches += 3;
    return BSQ_OK;
}

extern "C" bsq_status bsq_quality_sums(bsq_parser* p, int64_t first_record, int64_t count, int32_t* out_device,
                                       int32_t* out_host) {
    if (!p || first_record < 0 || count < 0) return BSQ_E_ARG;
    if (!p->have_pass || !(p->want & BSQ_WANT_BATCHES)) return BSQ_E_STATE;
    if (first_record + count > p->total_records) return BSQ_E_ARG;
    if (count == 0) return BSQ_OK;
    CK(cudaSetDevice(p->cfg.device_id));
    int32_t* dst = out_device;
    if (!dst) {
        CK(p->cub_tmp.ensure(4ull * (size_t)count, 1 << 16));
        dst = p->cub_tmp.as<int32_t>();
    }
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)p->sm_count * 8, (count + 7) / 8));
    k_quality_sums<<<grid, 256, 0, p->stream>>>(p->qual_out.as<uint8_t>(), p->ends.as<int64_t>(), p->ends_base.as<int64_t>(),
                                                 first_record, count, p->cfg.batch_size, (uint32_t)p->cfg.q_offset, dst);
    CK(cudaGetLastError());
    if (out_host) CK(cudaMemcpyAsync(out_host, dst, 4ull * (size_t)count, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return BSQ_OK;
}

extern "C" bsq_status bsq_offsets_to_host(bsq_parser* p, int32_t window, uint32_t* line_ends, uint32_t* id_spans) {
    bsq_offsets_view v;
    bsq_status st = bsq_get_offsets(p, window, &v);
    if (st != BSQ_OK) return st;
    CK(cudaSetDevice(p->cfg.device_id));
    if (line_ends) CK(cudaMemcpyAsync(line_ends, v.line_ends, 4 * (4 * v.n_records + 1), cudaMemcpyDeviceToHost, p->stream));
    if (id_spans && v.n_records) CK(cudaMemcpyAsync(id_spans, v.id_spans, 8 * v.n_records, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return BSQ_OK;
}

extern "C" const uint8_t* bsq_pass_device_input(const bsq_parser* p) { return p && p->have_pass ? p->pass_input : nullptr; }

extern "C" bsq_status bsq_last_timing(const bsq_parser* p, float ms[5], int64_t* n_launches) {
    if (!p || !p->have_pass) return BSQ_E_STATE;
    if (ms) memcpy(ms, p->ms, sizeof p->ms);
    if (n_launches) *n_launches = p->n_launches;
    return BSQ_OK;
}


// ------------------------------------------------------------------------------------------------
// FASTA
// ------------------------------------------------------------------------------------------------

namespace {

// the table of every newline position of one window (the views() pass of the FASTQ path, its checks ignored)
bsq_status materialize_line_ends(bsq_parser* p, Window& w) {
    bsq_status st = summarize_window(p, w, false);
    if (st != BSQ_OK) return st;
    CK(w.line_ends.ensure(4ull * ((size_t)w.scan.totals.newlines + 2), 1 << 16));
    CK(p->id_spans.ensure(8ull * ((size_t)w.scan.totals.records + 1), 1 << 20));
    CK(p->err_word.ensure(32));
    ResolveParams P{};
    P.run_pre = w.run_pre.as<BsqPrefix>();
    P.n_complete = w.scan.totals.records;
    P.strip_flag = reinterpret_cast<uint32_t*>(p->err_word.as<uint8_t>() + 8);
    P.bases = p->err_word.as<unsigned long long>() + 2;
    P.line_ends = w.line_ends.as<uint32_t>();
    P.id_spans = p->id_spans.as<uint32_t>();
    P.batch_size = 4096;
    P.rec_limit = 0xFFFFFFFFu;
    P.err = p->err_word.as<unsigned long long>();   // (FASTQ structure reports land here and are ignored)
    k_resolve<false, false, true, false><<<w.wp.n_runs, kThreads, smem_bytes(false, false), p->stream>>>(w.wp, P);
    p->n_launches += 1;
    CK(cudaGetLastError());
    return BSQ_OK;
}

bsq_status fasta_pass(bsq_parser* p, const uint8_t* d, uint64_t n, bsq_fasta_result* out) {
    memset(out, 0, sizeof *out);
    p->fa_have = false;
    p->have_pass = false;            // the FASTQ result views share buffers with this pass
    p->fa_input = d;
    if (n > kWindowMax) { p->last_error = "a FASTA pass is limited to one window (2 GiB - 1 MiB)"; return BSQ_E_ARG; }
    set_plain_error(&out->stop, BSQ_EOF, "EOF");
    if (n == 0) { p->fa_have = true; p->fa_records = p->fa_total_records = 0; p->fa_seq_bytes = 0; return BSQ_OK; }
    Window& w = p->fa_win;
    plan_window(p, w, d, n, false);
    w.region_off = 0;
    bsq_status st = materialize_line_ends(p, w);
    if (st != BSQ_OK) return st;
    const uint32_t nnl = w.scan.totals.newlines;
    uint8_t last_byte = 0;
    CK(cudaMemcpyAsync(&last_byte, d + n - 1, 1, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    const uint32_t L = nnl + (last_byte != '\n' ? 1u : 0u);
    out->n_lines = L;
    CK(p->fa_hdr.ensure(4ull * (L + 1), 1 << 16)); CK(p->fa_hcum.ensure(4ull * (L + 1), 1 << 16));
    CK(p->fa_slen.ensure(8ull * (L + 1), 1 << 16)); CK(p->fa_soff.ensure(8ull * (L + 1), 1 << 16));
    CK(p->fa_start.ensure(4ull * (L + 1), 1 << 16)); CK(p->fa_len.ensure(4ull * (L + 1), 1 << 16));
    CK(p->fa_err.ensure(16));
    CK(cudaMemsetAsync(p->fa_err.p, 0xFF, 16, p->stream));
    FastaLines F{};
    F.base = w.base; F.line_ends = w.line_ends.as<uint32_t>(); F.n_newlines = nnl; F.n_lines = L; F.end = w.wp.end;
    F.hdr = p->fa_hdr.as<uint32_t>(); F.slen = p->fa_slen.as<unsigned long long>();
    F.start = p->fa_start.as<uint32_t>(); F.len = p->fa_len.as<uint32_t>();
    const int grid = (int)std::max<uint32_t>(1, std::min<uint32_t>((L + 255) / 256, (uint32_t)p->sm_count * 32));
    k_fa_lines<<<grid, 256, 0, p->stream>>>(F);
    size_t t1 = 0, t2 = 0;
    CK(cub::DeviceScan::InclusiveSum(nullptr, t1, F.hdr, p->fa_hcum.as<uint32_t>(), (int)L, p->stream));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, t2, F.slen, p->fa_soff.as<unsigned long long>(), (int)L, p->stream));
    CK(p->cub_tmp.ensure(std::max(t1, t2) + 16));
    t1 = t2 = p->cub_tmp.cap;
    CK(cub::DeviceScan::InclusiveSum(p->cub_tmp.p, t1, F.hdr, p->fa_hcum.as<uint32_t>(), (int)L, p->stream));
    CK(cub::DeviceScan::ExclusiveSum(p->cub_tmp.p, t2, F.slen, p->fa_soff.as<unsigned long long>(), (int)L, p->stream));
    uint32_t n_rec = 0;
    unsigned long long soff_last = 0, slen_last = 0;
    CK(cudaMemcpyAsync(&n_rec, p->fa_hcum.as<uint32_t>() + (L - 1), 4, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(&soff_last, p->fa_soff.as<unsigned long long>() + (L - 1), 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaMemcpyAsync(&slen_last, p->fa_slen.as<unsigned long long>() + (L - 1), 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    const unsigned long long total_seq = soff_last + slen_last;
    CK(p->fa_seq.ensure(total_seq + 64, 1 << 20));
    CK(p->fa_seq_start.ensure(8ull * (n_rec + 2), 1 << 12));
    CK(p->fa_id_start.ensure(4ull * (n_rec + 1), 1 << 12)); CK(p->fa_id_len.ensure(4ull * (n_rec + 1), 1 << 12));
    CK(p->fa_hdr_line.ensure(4ull * (n_rec + 1), 1 << 12));
    FastaPack K{};
    K.base = w.base; K.n_lines = L; K.n_records = n_rec; K.check_ascii = p->cfg.check_ascii ? 1u : 0u;
    K.hdr = F.hdr; K.hcum = p->fa_hcum.as<uint32_t>(); K.soff = p->fa_soff.as<unsigned long long>();
    K.start = F.start; K.len = F.len; K.seq_out = p->fa_seq.as<uint8_t>();
    K.seq_start = p->fa_seq_start.as<unsigned long long>(); K.id_start = p->fa_id_start.as<uint32_t>();
    K.id_len = p->fa_id_len.as<uint32_t>(); K.hdr_line = p->fa_hdr_line.as<uint32_t>();
    K.total_seq = total_seq; K.err = p->fa_err.as<uint32_t>();
    const int gridw = (int)std::max<uint32_t>(1, std::min<uint32_t>((L + 7) / 8, (uint32_t)p->sm_count * 64));
    k_fa_pack<<<gridw, 256, 0, p->stream>>>(K);
    if (n_rec)
        k_fa_empty<<<std::max(1, std::min<int>((int)((n_rec + 255) / 256), p->sm_count * 8)), 256, 0, p->stream>>>(
            K.seq_start, n_rec, K.err);
    p->n_launches += 5;
    CK(cudaGetLastError());
    uint32_t herr[4];
    CK(cudaMemcpyAsync(herr, p->fa_err.p, 16, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    // ---- the first stop, in the order the reference meets them ----
    const uint32_t none = 0xFFFFFFFFu, begin = w.wp.begin;
    auto line_start = [&](uint32_t line, int64_t* pos) -> bsq_status {   // stream offset of the first byte of a line
        uint32_t le = 0;
        CK(cudaMemcpy(&le, w.line_ends.as<uint32_t>() + line, 4, cudaMemcpyDeviceToHost));
        *pos = (int64_t)(le + 1u) - (int64_t)begin;
        return BSQ_OK;
    };
    int64_t good = n_rec;
    if (herr[0] != none) {
        // a non-blank line before the first header: _read_header_line, parser.mojo:193-197 (context: records so far = 0)
        good = 0;
        int64_t pos = 0;
        st = line_start(herr[0], &pos);
        if (st != BSQ_OK) return st;
        memset(&out->stop, 0, sizeof out->stop);
        out->stop.code = BSQ_OTHER; out->stop.line_number = (int64_t)herr[0] + 1; out->stop.file_position = pos;
        Msg m{out->stop.message, sizeof out->stop.message, 0};
        m.str("FASTA: sequence id line does not start with '>'");
        m.str("\n  Line number: "); m.i64(out->stop.line_number);
        if (pos > 0) { m.str("\n  File position: "); m.i64(pos); }
    } else {
        const uint32_t e_empty = herr[2], e_ascii = herr[1];
        const uint32_t first = std::min(e_empty, e_ascii);
        if (first != none) {
            good = first;
            memset(&out->stop, 0, sizeof out->stop);
            Msg m{out->stop.message, sizeof out->stop.message, 0};
            if (e_empty == first) {                       // parser.mojo:152-160
                uint32_t hl[2] = {0, 0};
                CK(cudaMemcpy(hl, p->fa_hdr_line.as<uint32_t>() + first, first + 1 < n_rec ? 8 : 4, cudaMemcpyDeviceToHost));
                int64_t pos = (int64_t)n;                 // the line read last: the next header, or the end of the stream
                if (first + 1 < n_rec) { st = line_start(hl[1], &pos); if (st != BSQ_OK) return st; }
                out->stop.code = BSQ_OTHER; out->stop.record_number = (int64_t)first + 1;
                out->stop.line_number = (int64_t)hl[0] + 2; out->stop.file_position = pos;
                m.str("FASTA record has empty sequence");
                m.str("\n  Record number: "); m.i64(out->stop.record_number);
                m.str("\n  Line number: "); m.i64(out->stop.line_number);
                if (pos > 0) { m.str("\n  File position: "); m.i64(pos); }
            } else {                                      // parser.mojo:162-163, errors.mojo:223-234
                out->stop.code = BSQ_ASCII_INVALID; out->stop.record_number = first;
                m.str(code_message(BSQ_ASCII_INVALID));
                if (first > 0) { m.str("\n  Record number: "); m.i64(first); }
            }
        }
    }
    p->fa_total_records = n_rec;
    p->fa_records = good;
    p->fa_seq_bytes = total_seq;
    out->n_records = good;
    if (good == (int64_t)n_rec && herr[0] == none) out->n_bases = (int64_t)total_seq;
    else if (good > 0) {
        unsigned long long v = 0;
        CK(cudaMemcpy(&v, p->fa_seq_start.as<unsigned long long>() + good, 8, cudaMemcpyDeviceToHost));
        out->n_bases = (int64_t)v;
    }
    p->fa_have = true;
    return BSQ_OK;
}

}  // namespace

extern "C" bsq_status bsq_fasta_parse_device(bsq_parser* p, const uint8_t* dev_bytes, uint64_t n, bsq_fasta_result* out) {
    if (!p || !out || (!dev_bytes && n)) return BSQ_E_ARG;
    CK(cudaSetDevice(p->cfg.device_id));
    p->n_launches = 0;
    return fasta_pass(p, dev_bytes, n, out);
}

extern "C" bsq_status bsq_fasta_parse_host(bsq_parser* p, const uint8_t* host_bytes, uint64_t n, bsq_fasta_result* out) {
    if (!p || !out || (!host_bytes && n)) return BSQ_E_ARG;
    CK(cudaSetDevice(p->cfg.device_id));
    p->n_launches = 0;
    CK(p->host_input.ensure(n + 256, 1 << 20));
    if (n) CK(cudaMemcpyAsync(p->host_input.p, host_bytes, n, cudaMemcpyHostToDevice, p->stream));
    return fasta_pass(p, p->host_input.as<uint8_t>(), n, out);
}

extern "C" bsq_status bsq_fasta_get(const bsq_parser* p, bsq_fasta_view* out) {
    if (!p || !out) return BSQ_E_ARG;
    if (!p->fa_have) return BSQ_E_STATE;
    memset(out, 0, sizeof *out);
    out->n_records = p->fa_records;
    if (p->fa_total_records == 0) return BSQ_OK;
    out->sequence = p->fa_seq.as<uint8_t>();
    out->seq_starts = p->fa_seq_start.as<uint64_t>();
    out->id_start = p->fa_id_start.as<uint32_t>();
    out->id_len = p->fa_id_len.as<uint32_t>();
    out->input = p->fa_win.base;
    out->sequence_bytes = (int64_t)p->fa_seq_bytes;
    return BSQ_OK;
}

extern "C" bsq_status bsq_fasta_to_host(bsq_parser* p, uint8_t* seq, uint64_t* seq_starts, uint8_t* ids, uint64_t* id_starts) {
    bsq_fasta_view v;
    bsq_status st = bsq_fasta_get(p, &v);
    if (st != BSQ_OK) return st;
    const int64_t n = v.n_records;
    if (n == 0) { if (seq_starts) seq_starts[0] = 0; if (id_starts) id_starts[0] = 0; return BSQ_OK; }
    CK(cudaSetDevice(p->cfg.device_id));
    std::vector<uint64_t> ss(n + 1);
    CK(cudaMemcpy(ss.data(), v.seq_starts, 8ull * (n + 1), cudaMemcpyDeviceToHost));
    if (seq_starts) memcpy(seq_starts, ss.data(), 8ull * (n + 1));
    if (seq && ss[n]) CK(cudaMemcpy(seq, v.sequence, ss[n], cudaMemcpyDeviceToHost));
    if (ids || id_starts) {
        std::vector<uint32_t> a(n), l(n);
        CK(cudaMemcpy(a.data(), v.id_start, 4ull * n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(l.data(), v.id_len, 4ull * n, cudaMemcpyDeviceToHost));
        uint64_t off = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (id_starts) id_starts[i] = off;
            if (ids && l[i]) CK(cudaMemcpy(ids + off, v.input + a[i], l[i], cudaMemcpyDeviceToHost));
            off += l[i];
        }
        if (id_starts) id_starts[n] = off;
    }
    return BSQ_OK;
}

// ------------------------------------------------------------------------------------------------
// synthetic input
// ------------------------------------------------------------------------------------------------

static int ndigits(int64_t v) { int d = 1; while (v >= 10) { v /= 10; ++d; } return d; }

extern "C" int64_t bsq_compute_num_reads_for_size(int64_t target, int64_t mn, int64_t mx) {
    if (target <= 0) return 0;                       // utils.mojo:640-678
    const int64_t avg = (mn + mx) / 2;
    const int64_t est = target / (15 + 2 * avg + 4);
    if (est <= 0) return 0;
    const int64_t digits = est > 1 ? ndigits(est - 1) : 1;
    return target / (6 + digits + 1 + 2 * avg + 4);
}

static void len_prefix_table(int64_t mn, int64_t mx, std::vector<uint64_t>& t) {
    const int64_t m = mx - mn + 1;
    t.assign(m + 1, 0);
    for (int64_t j = 0; j < m; ++j) t[j + 1] = t[j] + (uint64_t)((j * 31 + 7) % m);
}

static uint64_t synth_offset_host(int64_t i, int64_t mn, int64_t mx, int digits, const std::vector<uint64_t>& t) {
    const uint64_t m = (uint64_t)(mx - mn + 1);
    const uint64_t lens = (uint64_t)i * mn + ((uint64_t)i / m) * t[m] + t[(uint64_t)i % m];
    return (uint64_t)i * (6 + digits + 1 + 4) + 2 * lens;
}

extern "C" int64_t bsq_synth_size(int64_t num_reads, int64_t mn, int64_t mx) {
    if (num_reads <= 0 || mn > mx || mn < 0) return 0;
    std::vector<uint64_t> t;
    len_prefix_table(mn, mx, t);
    return (int64_t)synth_offset_host(num_reads, mn, mx, num_reads > 1 ? ndigits(num_reads - 1) : 1, t);
}

extern "C" bsq_status bsq_synth_device(bsq_parser* p, uint8_t* dev_out, uint64_t capacity, int64_t num_reads,
                                       int64_t first, int64_t count, int64_t mn, int64_t mx, int64_t min_phred,
                                       int64_t max_phred, uint8_t q_lower, uint8_t q_upper, uint8_t q_offset,
                                       uint64_t* written) {
    if (!p || !dev_out || num_reads <= 0 || first < 0 || count < 0 || first + count > num_reads || mn < 0 || mn > mx ||
        min_phred < 0 || min_phred > max_phred || mx - mn > (1 << 24))
        return BSQ_E_ARG;
    CK(cudaSetDevice(p->cfg.device_id));
    std::vector<uint64_t> t;
    len_prefix_table(mn, mx, t);
    SynthParams G{};
    G.num_reads = num_reads; G.first = first; G.count = count;
    G.min_len = mn; G.max_len = mx; G.min_phred = min_phred; G.max_phred = max_phred;
    G.digits = num_reads > 1 ? ndigits(num_reads - 1) : 1;
    G.q_lower = q_lower; G.q_upper = q_upper; G.q_offset = q_offset;
    G.origin = synth_offset_host(first, mn, mx, G.digits, t);
    const uint64_t end = synth_offset_host(first + count, mn, mx, G.digits, t);
    if (end - G.origin > capacity) return BSQ_E_ARG;
    CK(p->len_prefix.ensure(8 * t.size()));
    CK(cudaMemcpyAsync(p->len_prefix.p, t.data(), 8 * t.size(), cudaMemcpyHostToDevice, p->stream));
    G.len_prefix = p->len_prefix.as<uint64_t>();
    G.period_sum = t.back();
    // 32-step jump of x -> a x + c (mod 2^64)
    auto jump = [](uint64_t a, uint64_t c, uint64_t& a32, uint64_t& c32) {
        a32 = 1; c32 = 0;
        for (int i = 0; i < 32; ++i) { c32 = c32 * a + c; a32 = a32 * a; }
    };
    jump(6364136223846793005ull, 1442695040888963407ull, G.a32_seq, G.c32_seq);
    jump(1664525ull, 1013904223ull, G.a32_q, G.c32_q);
    if (count > 0) {
        const int64_t warps_per_block = 8;
        const int grid = (int)std::min<int64_t>((count + warps_per_block - 1) / warps_per_block, (int64_t)p->sm_count * 32);
        k_synth<<<grid, 256, 0, p->stream>>>(G, dev_out);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(p->stream));
    if (written) *written = end - G.origin;
    return BSQ_OK;
}

// ------------------------------------------------------------------------------------------------
// shard stitching
// ------------------------------------------------------------------------------------------------

extern "C" bsq_status bsq_summarize_device(bsq_parser* p, const uint8_t* dev_bytes, uint64_t n, bsq_summary* out) {
    if (!p || !out || (!dev_bytes && n)) return BSQ_E_ARG;
    CK(cudaSetDevice(p->cfg.device_id));
    // positions in the summary are relative to the shard start; shards larger than a window are
    // folded window by window with the positions rebased on the host
    BsqSummary acc = bsq_summary_identity();
    uint64_t pos = 0;
    Window w;
    while (pos < n) {
        const uint64_t bytes = std::min<uint64_t>(n - pos, kWindowMax);
        plan_window(p, w, dev_bytes + pos, bytes, true);
        bsq_status st = summarize_window(p, w, true);
        if (st != BSQ_OK) { w.run_pre.release(); return st; }
        BsqSummary s = w.scan.region;
        // rebase window-relative positions to shard-relative (mod 2^32, like all rank algebra)
        const uint32_t shift = (uint32_t)(pos - w.wp.begin);
        for (int i = 0; i < 4; ++i) { s.last[i] += shift; s.first[i] += shift; }
        for (uint32_t k = 0; k < 4; ++k) {
            const uint32_t cnt_k = (s.count + 3u - k) >> 2;  // newlines with index == k (mod 4)
            s.P[k] += shift * cnt_k;
        }
        acc = bsq_combine(acc, s);
        pos += bytes;
    }
    w.run_pre.release();
    memcpy(out, &acc, sizeof acc);
    return BSQ_OK;
}

extern "C" bsq_status bsq_shard_prefix(const bsq_summary* shards, const uint64_t* shard_bytes, int32_t n_shards,
                                       bsq_shard_start* start) {
    if (!shards || !shard_bytes || !start || n_shards <= 0) return BSQ_E_ARG;
    int64_t rank = 0;
    bool prev_ends_with_newline = true;  // the stream start is a record start
    for (int32_t i = 0; i < n_shards; ++i) {
        BsqSummary s;
        memcpy(&s, &shards[i], sizeof s);
        start[i].newline_rank = rank;
        start[i].phase = (int32_t)(rank & 3);
        start[i]._pad = 0;
        // records are owned by the shard that holds their first byte
        uint64_t skip;
        const bool at_record_start = (rank & 3) == 0 && prev_ends_with_newline;
        if (at_record_start) {
            skip = 0;
        } else {
            const uint32_t j = (uint32_t)((3 - (rank & 3)) & 3);  // first newline of class 3 in the shard
            skip = j < s.count ? (uint64_t)s.first[j] + 1u : shard_bytes[i];
        }
        start[i].skip_bytes = (int64_t)skip;
        // records completed before the first byte this shard owns
        start[i].first_record = at_record_start ? (rank >> 2) : (rank >> 2) + 1;
        if (s.count > 0) prev_ends_with_newline = (uint64_t)s.last[0] + 1u == shard_bytes[i];
        else if (shard_bytes[i] > 0) prev_ends_with_newline = false;
        rank += s.count;
    }
    return BSQ_OK;
}
