// query-p on the GPU path, end to end from a file: the reference's `lphash query-p -i index -q file`
// (src/query.cpp:35-96) with the kseq loop replaced by the streaming ingest of include/lphash_b200_fastx.hpp:
// one background thread inflates and splits records chunk by chunk while the calling thread sends the
// previous chunk through lphb_query_stream (or lphb_query_stream_runs), so inflate + parse overlap the
// copies and the kernels.
//
//   lphb_query <index.lph> <kmer_bits: 64|128> <query.fa|.fq[.gz]> [device] [chunk_MB] [runs] [nofold]
//
// Prints one CSV line: query file, index file, total k-mers, ns per k-mer end to end (file open to the last
// code on the host), ns per k-mer of the GPU calls alone, the 64-bit FNV-1a-style fold of all hash codes in
// file order for cross-checking against the reference, bases per second end to end.
// With `runs` the codes come back as run records (about 2 bytes per k-mer over PCIe) and are expanded on
// the host for the fold.  `nofold` leaves the codes (or run records) untouched in the host buffer, like the
// reference's loop, which discards them (src/query.cpp:54); the fold column is then 0.  The result buffers are
// pinned (lphb_host_alloc), the bases go up from the parser's pageable vector.
//
//   g++ -std=c++17 -O2 -DLPHASH_B200_WITH_ZLIB -Iinclude examples/lphb_query.cpp -o lphb_query
//       (continued) -Llphash_b200 -llphash_b200 -Wl,-rpath,$PWD/lphash_b200 -lz -pthread
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "lphash_b200.hpp"
#include "lphash_b200_fastx.hpp"

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <index.lph> <kmer_bits> <query.fa|.fq[.gz]> [device] [chunk_MB] [runs]\n", argv[0]);
        return 1;  // usage error, like the reference CLI
    }
    const int bits = std::atoi(argv[2]);
    const int device = argc > 4 ? std::atoi(argv[4]) : 0;
    const size_t chunk = (argc > 5 ? size_t(std::atoll(argv[5])) : 4) << 20;  // a few MB: small enough to overlap, large enough to fill the GPU
    bool use_runs = false, fold = true;
    for (int i = 6; i < argc; ++i) {
        if (std::strcmp(argv[i], "runs") == 0) use_runs = true;
        if (std::strcmp(argv[i], "nofold") == 0) fold = false;
    }
    try {
        using clock = std::chrono::steady_clock;
        lphb_mphf* f = nullptr;
        if (lphb_mphf_load_file(argv[1], bits, device, &f) != LPHB_OK) {
            std::fprintf(stderr, "%s\n", lphb_last_error());
            return 2;
        }
        lphb_info info{};
        lphb_mphf_info(f, &info);
        std::vector<uint64_t> code_off;
        uint64_t* codes = nullptr;        // pinned
        unsigned char* runs = nullptr;    // pinned
        uint64_t codes_cap = 0, runs_cap = 0;
        auto grow = [&](void** p, uint64_t& have, uint64_t want_bytes) {
            if (want_bytes <= have) return;
            if (*p) lphb_host_free(*p);
            *p = nullptr;
            if (lphb_host_alloc(p, want_bytes + want_bytes / 4) != LPHB_OK) throw std::runtime_error(lphb_last_error());
            have = want_bytes + want_bytes / 4;
        };
        uint64_t total = 0, total_bases = 0, h = 0xcbf29ce484222325ULL;
        double call_ns = 0;
        int rc = LPHB_OK;
        const auto t0 = clock::now();
        lphash_b200::fastx::stream_file(argv[3], chunk, [&](lphash_b200::fastx::Batch& b) {
            uint64_t cap = 0;
            for (uint64_t c = 0; c + 1 < b.offsets.size(); ++c) {
                const uint64_t len = b.offsets[c + 1] - b.offsets[c];
                if (len >= info.m) cap += len - info.m + 1;  // room for the streaming quirk's extras
            }
            if (!use_runs || fold) grow(reinterpret_cast<void**>(&codes), codes_cap, (cap + 1) * 8);
            if (code_off.size() < b.offsets.size()) code_off.resize(b.offsets.size());
            uint64_t n_codes = 0;
            const auto t1 = clock::now();
            if (use_runs) {
                grow(reinterpret_cast<void**>(&runs), runs_cap, (cap + 1) * 12);
                uint64_t n_runs = 0;
                rc = lphb_query_stream_runs(f, b.bases.data(), b.offsets.data(), b.n_records(), runs, cap + 1, &n_runs,
                                            code_off.data(), &n_codes);
                call_ns += std::chrono::duration<double, std::nano>(clock::now() - t1).count();
                if (rc == LPHB_OK && fold) rc = lphb_expand_runs(runs, n_runs, codes, cap + 1, &n_codes, 4);
            } else {
                rc = lphb_query_stream(f, b.bases.data(), b.offsets.data(), b.n_records(), codes, cap + 1,
                                       code_off.data(), &n_codes);
                call_ns += std::chrono::duration<double, std::nano>(clock::now() - t1).count();
            }
            if (rc != LPHB_OK) throw std::runtime_error(lphb_last_error());
            if (fold)
                for (uint64_t i = 0; i < n_codes; ++i) h = (h ^ codes[i]) * 0x100000001b3ULL;
            total += n_codes;
            total_bases += b.bases.size();
        });
        const double total_ns = std::chrono::duration<double, std::nano>(clock::now() - t0).count();
        if (codes) lphb_host_free(codes);
        if (runs) lphb_host_free(runs);
        std::printf("%s,%s,%llu,%.4f,%.4f,%016llx,%.4g\n", argv[3], argv[1], (unsigned long long)total,
                    total ? total_ns / double(total) : 0.0, total ? call_ns / double(total) : 0.0, (unsigned long long)(fold ? h : 0),
                    total_ns > 0 ? double(total_bases) / (total_ns * 1e-9) : 0.0);
        lphb_mphf_free(f);
        return 0;
    } catch (std::exception const& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }
}
