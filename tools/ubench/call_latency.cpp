// Per-call cost of lphb_query_stream for one record per call (what the reference's query loop does, src/query.cpp:48-56),
// as a function of the record length.  usage: call_latency <index.lph> <kmer_bits>
//   g++ -std=c++17 -O2 -Iinclude tools/ubench/call_latency.cpp -o call_latency -Llphash_b200 -llphash_b200 -Wl,-rpath,$PWD/lphash_b200
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "lphash_b200.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 1;
    lphash_b200::mphf f;
    if (std::atoi(argv[2]) == 128) return 2;  // 64-bit flavour only (LPHASH_B200_KMER_BITS default)
    f.load(argv[1]);
    std::mt19937_64 rng(1);
    for (size_t len : {100, 1000, 8000, 64000, 500000}) {
        std::string s(len, 'A');
        for (auto& c : s) c = "ACGT"[rng() & 3];
        const uint64_t offsets[2] = {0, len};
        std::vector<uint64_t> codes(len), coff(2);
        uint64_t n = 0;
        for (int i = 0; i < 20; ++i) lphb_query_stream(f.handle(), s.data(), offsets, 1, codes.data(), codes.size(), coff.data(), &n);
        const int reps = 200;
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < reps; ++i) lphb_query_stream(f.handle(), s.data(), offsets, 1, codes.data(), codes.size(), coff.data(), &n);
        double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
        auto t1 = std::chrono::steady_clock::now();
        for (int i = 0; i < reps; ++i) { auto v = f(s.data(), len); n = v.size(); }
        double us2 = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t1).count() / reps;
        lphb_stats st{};
        lphb_mphf_stats(f.handle(), &st);
        std::printf("record %7zu bases: %8.1f us per lphb_query_stream call, %8.1f us per operator() call (%.1f ns per k-mer), kernel %.1f us\n", len, us, us2,
                    us2 * 1e3 / double(n), st.kernel_ms * 1e3);
    }
    return 0;
}
