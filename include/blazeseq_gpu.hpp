// blazeseq_gpu.hpp -- C++17 host mirror of BlazeSeq's parser API over the C ABI of blazeseq_gpu.h (header only).
//
// The reference is compiled code (Mojo); a compiled caller that wants the same surface without Python gets it here:
//
//   ParserConfig                     blazeseq/fastq/parser.mojo:33-74
//   FastqParser::has_more / next_view / next_batch / views() / batches()
//                                    blazeseq/fastq/parser.mojo:147-274,628-735
//   FastqView                        blazeseq/fastq/record.mojo:431-550   (spans into the current region, valid until the
//                                                                          parser moves to its next region)
//   FastqBatch / DeviceFastqBatch    blazeseq/fastq/record_batch.mojo:19-87,210-244
//   Error (what(), code, record / line / position)   blazeseq/errors.mojo:43-90,178-234
//   EOF = the text "EOF"             blazeseq/CONSTS.mojo:19, io/buffered.mojo:102-112
//
// Sources: a file path (FileReader / GZFile / RapidgzipReader, chosen by suffix like python/blazeseq_parser.mojo:100-114 --
// the library's stream pipeline reads, inflates and parses region by region) or bytes in host memory (MemoryReader).
// All byte work happens on the device; without a usable B200 the constructor throws -- there is no CPU parse.
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "blazeseq_gpu.h"

namespace blazeseq {

struct ParserConfig {                       // parser.mojo:33-74
    bool check_ascii = false;
    bool check_quality = false;
    int64_t buffer_capacity = 256 * 1024;   // CONSTS.mojo:26
    int64_t buffer_max_capacity = 1ll << 30;
    bool buffer_growth_enabled = false;
    int32_t parallelism = 0;                // RapidgzipReader(parallelism): 0 = all cores
    int32_t device_id = 0;
};

class Error : public std::runtime_error {   // String(e) of the reference, with its context
  public:
    Error(const bsq_error& e) : std::runtime_error(e.message), code(e.code), record_number(e.record_number),
                                line_number(e.line_number), file_position(e.file_position) {}
    Error(int32_t c, const std::string& what) : std::runtime_error(what), code(c) {}
    int32_t code = BSQ_OTHER;
    int64_t record_number = 0, line_number = 0, file_position = 0;
};
struct EOFError : Error {                   // buffered.mojo:102-112: the text is "EOF"
    EOFError() : Error(BSQ_EOF, "EOF") {}
};

struct FastqView {                          // record.mojo:431-550
    std::string_view id, sequence, quality;
    uint8_t phred_offset = 33;
    size_t size() const { return sequence.size(); }                              // __len__ = bases
    size_t byte_len() const { return 1 + id.size() + sequence.size() + quality.size() + 5; }
};

// host FastqBatch: three byte arrays + two inclusive cumulative Int64 arrays restarting at 0 (record_batch.mojo:22-27,77-87)
struct FastqBatch {
    std::vector<uint8_t> sequence, quality, id;
    std::vector<int64_t> ends, id_ends;
    uint8_t quality_offset = 33;
    int64_t num_records() const { return (int64_t)ends.size(); }
    int64_t seq_len() const { return ends.empty() ? 0 : ends.back(); }
    FastqView get_ref(int64_t i) const {
        const int64_t a = i ? ends[(size_t)i - 1] : 0, b = ends[(size_t)i], ia = i ? id_ends[(size_t)i - 1] : 0, ib = id_ends[(size_t)i];
        return FastqView{{reinterpret_cast<const char*>(id.data()) + ia, (size_t)(ib - ia)},
                         {reinterpret_cast<const char*>(sequence.data()) + a, (size_t)(b - a)},
                         {reinterpret_cast<const char*>(quality.data()) + a, (size_t)(b - a)}, quality_offset};
    }
};

// device-resident batch: pointers into the parser's arena, valid until its next region (record_batch.mojo:210-220)
struct DeviceFastqBatch {
    bsq_batch_view view{};
    bsq_parser* parser = nullptr;
    int64_t index = 0;
    int64_t num_records() const { return view.num_records; }
    int64_t seq_len() const { return view.seq_len; }
    FastqBatch copy_to_host() const {                                            // record_batch.mojo:222-241
        FastqBatch b;
        b.sequence.resize((size_t)view.sequence_bytes); b.quality.resize((size_t)view.seq_len); b.id.resize((size_t)view.total_id_bytes);
        b.ends.resize((size_t)view.num_records); b.id_ends.resize((size_t)view.num_records);
        b.quality_offset = view.quality_offset;
        if (view.num_records && bsq_batch_to_host(parser, index, b.sequence.data(), b.quality.data(), b.id.data(), b.ends.data(),
                                                  b.id_ends.data()) != BSQ_OK)
            throw Error(BSQ_OTHER, std::string("bsq_batch_to_host: ") + bsq_last_error_text(parser));
        return b;
    }
};

class FastqParser {
  public:
    // FastqParser[FileReader | GZFile | RapidgzipReader, config](reader, schema, batch_size): the file is read, inflated and parsed
    // region by region by the library's stream pipeline
    FastqParser(const std::string& path, const std::string& quality_schema = "generic", const ParserConfig& config = ParserConfig(),
                int32_t batch_size = 4096, uint64_t region_bytes = 512ull << 20) {
        create(quality_schema, config, batch_size);
        if (bsq_stream_open(p_, path.c_str(), BSQ_SOURCE_AUTO, region_bytes, &stream_) != BSQ_OK) {
            const std::string why = bsq_last_error_text(p_);
            bsq_destroy(p_);
            p_ = nullptr;
            throw Error(BSQ_OTHER, "cannot open " + path + (why.empty() ? "" : ": " + why));
        }
    }
    // FastqParser[MemoryReader, config]: bytes the caller keeps alive
    FastqParser(const uint8_t* data, size_t n, const std::string& quality_schema = "generic", const ParserConfig& config = ParserConfig(),
                int32_t batch_size = 4096) {
        create(quality_schema, config, batch_size);
        mem_ = data; mem_n_ = n;
    }
    FastqParser(const FastqParser&) = delete;
    FastqParser& operator=(const FastqParser&) = delete;
    ~FastqParser() {
        if (stream_) bsq_stream_close(stream_);
        if (p_) bsq_destroy(p_);
    }

    // parser.mojo:147-158
    bool has_more() {
        if (cursor_ < region_records_) return true;
        if (finished_) return false;
        if (want_ == 0) want_ = BSQ_WANT_OFFSETS;
        load(want_);
        return cursor_ < region_records_;
    }

    // parser.mojo:160-170: the next record as spans into the region; throws EOFError at the end, Error on the first
    // malformed / invalid record (after every record before it has been delivered)
    FastqView next_view() {
        if (want_ != BSQ_WANT_OFFSETS) { if (started_) throw Error(BSQ_OTHER, "views() and batches() cannot be mixed on one parser"); want_ = BSQ_WANT_OFFSETS; }
        while (cursor_ >= region_records_) {
            if (finished_) raise_stop();
            load(BSQ_WANT_OFFSETS);
        }
        // window of the record (a region has one window per 2 GiB)
        while (win_ + 1 < (int)wins_.size() && cursor_ >= wins_[(size_t)win_ + 1].first) ++win_;
        const Win& w = wins_[(size_t)win_];
        const int64_t i = cursor_ - w.first;
        const uint32_t* le = w.line_ends.data() + 4 * i;
        const char* base = reinterpret_cast<const char*>(region_) + (w.stream_base - region_offset_);
        // u32 arithmetic relative to the window base (the leading sentinel is begin - 1)
        const uint32_t seq0 = le[1] + 1u, qual0 = le[3] + 1u, rec_end = le[4];
        const uint32_t id0 = w.id_spans[(size_t)(2 * i)], idn = w.id_spans[(size_t)(2 * i + 1)];
        ++cursor_;
        return FastqView{{base + id0, idn}, {base + seq0, le[2] - seq0}, {base + qual0, rec_end - qual0}, q_offset_};
    }

    // parser.mojo:239-251: the next (up to) batch_size records as a device batch; throws EOFError at the end
    DeviceFastqBatch next_device_batch() {
        if (want_ != BSQ_WANT_BATCHES) { if (started_) throw Error(BSQ_OTHER, "views() and batches() cannot be mixed on one parser"); want_ = BSQ_WANT_BATCHES; }
        while (batch_ >= region_batches_) {
            if (finished_) raise_stop();
            load(BSQ_WANT_BATCHES);
        }
        DeviceFastqBatch b;
        b.parser = p_; b.index = batch_;
        if (bsq_get_batch(p_, batch_, &b.view) != BSQ_OK) throw Error(BSQ_OTHER, std::string("bsq_get_batch: ") + bsq_last_error_text(p_));
        ++batch_;
        cursor_ += b.view.num_records;
        return b;
    }
    FastqBatch next_batch() { return next_device_batch().copy_to_host(); }

    // views() / batches(): for_each forms of the reference's iterators (parser.mojo:628-735): they end at EOF and, like the
    // reference's iterators, report a parse error after the records before it (here: by rethrowing it)
    template <class F> void views(F&& f) {
        for (;;) {
            FastqView v;
            try { v = next_view(); } catch (const EOFError&) { return; }
            f(v);
        }
    }
    template <class F> void device_batches(F&& f) {
        for (;;) {
            DeviceFastqBatch b;
            try { b = next_device_batch(); } catch (const EOFError&) { return; }
            f(b);
        }
    }
    template <class F> void batches(F&& f) {
        device_batches([&](const DeviceFastqBatch& d) { f(d.copy_to_host()); });
    }

    bsq_parser* handle() const { return p_; }

  private:
    struct Win { int64_t stream_base = 0, first = 0, n = 0; std::vector<uint32_t> line_ends, id_spans; };

    void create(const std::string& schema, const ParserConfig& c, int32_t batch_size) {
        bsq_config cfg;
        bsq_default_config(&cfg);
        cfg.device_id = c.device_id;
        cfg.check_ascii = c.check_ascii; cfg.check_quality = c.check_quality;
        bsq_parse_schema(schema.c_str(), &cfg.q_lower, &cfg.q_upper, &cfg.q_offset);    // unknown names fall back to generic
        cfg.buffer_capacity = c.buffer_capacity; cfg.buffer_max_capacity = c.buffer_max_capacity;
        cfg.buffer_growth_enabled = c.buffer_growth_enabled;
        cfg.batch_size = batch_size;
        cfg.inflate_threads = c.parallelism;
        q_offset_ = cfg.q_offset;
        const bsq_status st = bsq_create(&cfg, &p_);
        if (st == BSQ_E_NO_DEVICE) throw Error(BSQ_OTHER, "blazeseq_gpu: no usable CUDA device (this library never parses on the CPU)");
        if (st != BSQ_OK) throw Error(BSQ_OTHER, "blazeseq_gpu: bsq_create failed");
    }

    // the next region: one pass of the stream pipeline, or the whole memory buffer
    void load(uint32_t want) {
        started_ = true;
        bsq_pass_result r;
        std::memset(&r, 0, sizeof r);
        bsq_status st;
        if (stream_) {
            st = bsq_stream_next(stream_, want, &r);
            if (st == BSQ_OK && (want & BSQ_WANT_OFFSETS)) {
                uint64_t n = 0; int64_t off = 0, first = 0;
                region_ = bsq_stream_region(stream_, &n, &off, &first);
                region_offset_ = off;
            }
        } else {
            st = bsq_parse_host(p_, mem_, mem_n_, 0, 0, 1, want, &r);
            region_ = mem_; region_offset_ = 0;
        }
        if (st != BSQ_OK) throw Error(BSQ_OTHER, std::string("blazeseq_gpu: ") + bsq_last_error_text(p_));
        stop_ = r.stop;
        finished_ = r.stop.code != BSQ_OK;
        cursor_ = 0; region_records_ = r.n_records;
        batch_ = 0; region_batches_ = r.n_batches;
        wins_.clear(); win_ = 0;
        if (want & BSQ_WANT_OFFSETS) {
            int64_t first = 0;
            for (int32_t w = 0; w < r.n_windows; ++w) {
                bsq_offsets_view v;
                if (bsq_get_offsets(p_, w, &v) != BSQ_OK) break;
                Win x;
                x.stream_base = v.stream_base; x.first = first; x.n = v.n_records;
                x.line_ends.resize((size_t)(4 * v.n_records + 1)); x.id_spans.resize((size_t)(2 * v.n_records));
                if (bsq_offsets_to_host(p_, w, x.line_ends.data(), v.n_records ? x.id_spans.data() : nullptr) != BSQ_OK)
                    throw Error(BSQ_OTHER, std::string("bsq_offsets_to_host: ") + bsq_last_error_text(p_));
                first += v.n_records;
                wins_.push_back(std::move(x));
            }
        }
    }

    [[noreturn]] void raise_stop() {
        if (stop_.code == BSQ_EOF || stop_.code == BSQ_OK) throw EOFError();
        throw Error(stop_);
    }

    bsq_parser* p_ = nullptr;
    bsq_stream* stream_ = nullptr;
    const uint8_t* mem_ = nullptr;
    size_t mem_n_ = 0;
    uint32_t want_ = 0;
    uint8_t q_offset_ = 33;
    bool started_ = false, finished_ = false;
    bsq_error stop_{};
    const uint8_t* region_ = nullptr;        // host bytes of the current region
    int64_t region_offset_ = 0;              // stream offset of region_[0]
    int64_t cursor_ = 0, region_records_ = 0;
    int64_t batch_ = 0, region_batches_ = 0;
    std::vector<Win> wins_;
    int win_ = 0;
};

}  // namespace blazeseq
