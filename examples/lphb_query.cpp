// query-p on the GPU path, end to end from a file: the reference's `lphash query-p -i index -q file`
// (src/query.cpp:35-96) with the kseq loop replaced by one batched call.
//
//   lphb_query <index.lph> <kmer_bits: 64|128> <query.fa|.fq[.gz]> [device]
//
// Prints one CSV line: query file, index file, total k-mers, ns per k-mer (parse + H2D + kernels + D2H),
// ns per k-mer of the GPU call alone, and the 64-bit FNV-1a-style fold of all hash codes in file order
//  for cross-checking against the reference.
//
//   g++ -std=c++17 -O2 -DLPHASH_B200_WITH_ZLIB -Iinclude examples/lphb_query.cpp -o lphb_query
//       (continued) -Llphash_b200 -llphash_b200 -Wl,-rpath,$PWD/lphash_b200 -lz
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "lphash_b200.hpp"
#include "lphash_b200_fastx.hpp"

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <index.lph> <kmer_bits> <query.fa|.fq[.gz]> [device]\n", argv[0]);
        return 1;  // usage error, like the reference CLI
    }
    const int bits = std::atoi(argv[2]);
    const int device = argc > 4 ? std::atoi(argv[4]) : 0;
    try {
        using clock = std::chrono::steady_clock;
        lphb_mphf* f = nullptr;
        if (lphb_mphf_load_file(argv[1], bits, device, &f) != LPHB_OK) {
            std::fprintf(stderr, "%s\n", lphb_last_error());
            return 2;
        }
        auto t0 = clock::now();
        lphash_b200::fastx::Batch b;
        lphash_b200::fastx::read_file(argv[3], b);
        lphb_info info{};
        lphb_mphf_info(f, &info);
        uint64_t cap = 0;
        for (uint64_t c = 0; c + 1 < b.offsets.size(); ++c) {
            const uint64_t len = b.offsets[c + 1] - b.offsets[c];
            if (len >= info.m) cap += len - info.m + 1;  // room for the streaming quirk's extras
        }
        std::vector<uint64_t> codes(cap ? cap : 1), code_off(b.offsets.size());
        uint64_t n_codes = 0;
        auto t1 = clock::now();
        int rc = lphb_query_stream(f, b.bases.data(), b.offsets.data(), b.n_records(), codes.data(), codes.size(),
                                   code_off.data(), &n_codes);
        auto t2 = clock::now();
        if (rc != LPHB_OK) {
            std::fprintf(stderr, "%s\n", lphb_last_error());
            lphb_mphf_free(f);
            return 3;
        }
        uint64_t h = 0xcbf29ce484222325ULL;
        for (uint64_t i = 0; i < n_codes; ++i) h = (h ^ codes[i]) * 0x100000001b3ULL;
        const double total_ns = std::chrono::duration<double, std::nano>(t2 - t0).count();
        const double call_ns = std::chrono::duration<double, std::nano>(t2 - t1).count();
        std::printf("%s,%s,%llu,%.4f,%.4f,%016llx\n", argv[3], argv[1], (unsigned long long)n_codes,
                    n_codes ? total_ns / double(n_codes) : 0.0, n_codes ? call_ns / double(n_codes) : 0.0,
                    (unsigned long long)h);
        lphb_mphf_free(f);
        return 0;
    } catch (std::exception const& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }
}
