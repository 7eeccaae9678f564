// Test driver for include/lphash_b200_fastx.hpp: parses argv[1], writes u64 n, offsets[n+1], bases to argv[2].
// With a third argument (chunk size in bytes) the file goes through the streaming ingest (stream_file: chunked
// reads, records held back across chunk ends) and the batches are concatenated.
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "lphash_b200_fastx.hpp"

int main(int argc, char** argv) {
    if (argc != 3 && argc != 4) return 1;
    lphash_b200::fastx::Batch b;
    try {
        if (argc == 4) {
            b.offsets.push_back(0);
            lphash_b200::fastx::stream_file(argv[1], size_t(std::atoll(argv[3])), [&](lphash_b200::fastx::Batch& part) {
                const uint64_t base = b.bases.size();
                b.bases.insert(b.bases.end(), part.bases.begin(), part.bases.end());
                for (size_t i = 1; i < part.offsets.size(); ++i) b.offsets.push_back(base + part.offsets[i]);
            });
        } else {
            lphash_b200::fastx::read_file(argv[1], b);
        }
    } catch (std::exception const& e) {
        std::fprintf(stderr, "fastx_check: %s\n", e.what());
        return 3;
    }
    if (b.offsets.empty()) b.offsets.push_back(0);
    std::ofstream out(argv[2], std::ios::binary);
    uint64_t n = b.n_records();
    out.write(reinterpret_cast<const char*>(&n), 8);
    out.write(reinterpret_cast<const char*>(b.offsets.data()), std::streamsize(8 * b.offsets.size()));
    out.write(b.bases.data(), std::streamsize(b.bases.size()));
    return 0;
}
