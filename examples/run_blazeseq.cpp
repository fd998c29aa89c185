// The reference's benchmark runners (benchmark/fastq-parser/run_blazeseq.mojo, run_blazeseq_batch.mojo,
// run_blazeseq_gzip.mojo) over the B200 library, through the C++ host mirror of include/blazeseq_gpu.hpp: reads a FASTQ file
// (.fastq / .fq, .gz / .bgz by suffix), counts records and base pairs and prints exactly "records base_pairs" on one line, the
// cross-check line of benchmark/fastq-parser/run_benchmarks.sh:317-337.
//
//   run_blazeseq <path> [views|batches|device_batches] [batch_size] [validate]
//
// build: g++ -std=c++17 -O2 -Iinclude examples/run_blazeseq.cpp -Lblazeseq_b200/lib -lblazeseq_gpu -Wl,-rpath,$PWD/blazeseq_b200/lib
#include <cstdio>
#include <cstdlib>
#include <string>

#include "blazeseq_gpu.hpp"

int main(int argc, char** argv) {
    if (argc < 2) {
        std::printf("Usage: run_blazeseq <path.fastq[.gz]> [views|batches|device_batches] [batch_size] [validate]\n");
        return 0;
    }
    const std::string path = argv[1], mode = argc > 2 ? argv[2] : "views";
    const int batch_size = argc > 3 ? std::atoi(argv[3]) : 4096;
    if (batch_size <= 0) { std::printf("batch_size must be positive\n"); return 0; }
    blazeseq::ParserConfig config;                      // ParserConfig(check_ascii=False, check_quality=False, ...)
    config.check_ascii = config.check_quality = argc > 4 && std::string(argv[4]) == "validate";
    long long total_reads = 0, total_base_pairs = 0;
    try {
        blazeseq::FastqParser parser(path, "generic", config, batch_size);
        if (mode == "views") {
            parser.views([&](const blazeseq::FastqView& record) { total_reads += 1; total_base_pairs += (long long)record.size(); });
        } else if (mode == "batches") {
            parser.batches([&](const blazeseq::FastqBatch& batch) { total_reads += batch.num_records(); total_base_pairs += batch.seq_len(); });
        } else if (mode == "device_batches") {
            parser.device_batches([&](const blazeseq::DeviceFastqBatch& batch) { total_reads += batch.num_records(); total_base_pairs += batch.seq_len(); });
        } else {
            std::printf("Invalid mode. Expected 'views', 'batches' or 'device_batches'.\n");
            return 0;
        }
    } catch (const blazeseq::Error& e) {
        // the reference's iterators print the error and stop; the counts so far are still reported
        std::printf("%lld %lld\n", total_reads, total_base_pairs);
        std::fprintf(stderr, "%s\n", e.what());
        return e.code > 0 ? 2 : 1;
    }
    std::printf("%lld %lld\n", total_reads, total_base_pairs);
    return 0;
}
