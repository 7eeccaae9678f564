// Exercises include/lphash_b200.hpp the way the reference's driver and builder use lphash::mphf and
// namespace minimizer (src/query.cpp:36-56, src/partitioned_mphf.cpp:70-77, 120-129):
//   shim_check <file.lph> <batch.bin> <out.bin>
// batch.bin: u64 n_contigs, u64 offsets[n+1], bases;  then u64 k, m, n_ids, ids[n_ids] and the
// index batch (u64 n_contigs, offsets, bases) for the build-side calls.
// out.bin:   u64 n_codes, codes (one operator() call per contig, concatenated), u64 n_records,
//            records (18 B each), u64 n_kmers, u64 mm_count, u64 n_colliding, colliding k-mers.
// Exit code 3 = std::runtime_error from the library (e.g. no CUDA device): printed to stderr.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "lphash_b200.hpp"

namespace lb = lphash_b200;

static std::vector<char> slurp(const char* path) {
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) throw std::runtime_error(std::string("cannot open ") + path);
    std::vector<char> buf(size_t(in.tellg()));
    in.seekg(0);
    in.read(buf.data(), std::streamsize(buf.size()));
    return buf;
}

struct Batch {
    uint64_t n = 0;
    const uint64_t* offsets = nullptr;
    const char* bases = nullptr;
};

static const char* read_batch(const char* p, Batch& b) {
    b.n = *reinterpret_cast<const uint64_t*>(p);
    b.offsets = reinterpret_cast<const uint64_t*>(p + 8);
    b.bases = p + 8 + 8 * (b.n + 1);
    uint64_t nbytes = b.offsets[b.n];
    return b.bases + ((nbytes + 7) & ~uint64_t(7));
}

int main(int argc, char** argv) {
    if (argc != 4) return 1;
    try {
        std::vector<char> in = slurp(argv[2]);
        Batch q, idx;
        const char* p = read_batch(in.data(), q);
        const uint64_t* hdr = reinterpret_cast<const uint64_t*>(p);
        uint64_t k = hdr[0], m = hdr[1], n_ids = hdr[2];
        std::vector<uint64_t> ids(hdr + 3, hdr + 3 + n_ids);
        read_batch(reinterpret_cast<const char*>(hdr + 3 + n_ids), idx);

        lb::mphf hf;
        hf.load(argv[1]);
        std::vector<uint64_t> codes;
        for (uint64_t c = 0; c < q.n; ++c) {  // one record at a time, like query<MPHF>
            auto h = hf(q.bases + q.offsets[c], q.offsets[c + 1] - q.offsets[c], true);
            codes.insert(codes.end(), h.begin(), h.end());
        }
        // the whole batch in one call must agree
        std::vector<uint64_t> bcodes, boff;
        hf.query_batch(q.bases, q.offsets, q.n, bcodes, boff);
        if (bcodes != codes) throw std::logic_error("batch call differs from per-contig calls");

        std::vector<lb::mm_record_t> records;
        uint64_t mm_count = 0, n_kmers = 0;
        for (uint64_t c = 0; c < idx.n; ++c)
            n_kmers += lb::minimizer::from_string(idx.bases + idx.offsets[c], idx.offsets[c + 1] - idx.offsets[c],
                                                  uint32_t(k), uint32_t(m), 42, false, mm_count, records);
        std::vector<lb::kmer_t> coll;
        uint64_t mm2 = 0;
        auto it = ids.cbegin(), stop = ids.cend();
        for (uint64_t c = 0; c < idx.n; ++c)
            lb::minimizer::get_colliding_kmers(idx.bases + idx.offsets[c], idx.offsets[c + 1] - idx.offsets[c],
                                               uint32_t(k), uint32_t(m), 42, false, it, stop, mm2, coll);
        if (it != stop || mm2 != mm_count) throw std::logic_error("colliding-id stream not consumed");
        // classify over the scan's stream must reproduce the id list the test handed in
        std::vector<lb::mm_triplet_t> uniq;
        std::vector<uint64_t> coll_ids;
        lb::minimizer::classify(records, uniq, coll_ids);
        if (coll_ids != ids) throw std::logic_error("classify: colliding ids differ");
        for (size_t i = 1; i < uniq.size(); ++i)
            if (!(uniq[i - 1].itself < uniq[i].itself)) throw std::logic_error("classify: triplets not strictly ascending");
        lb::minimizer::release_workspace();

        std::ofstream out(argv[3], std::ios::binary);
        auto put = [&](uint64_t v) { out.write(reinterpret_cast<const char*>(&v), 8); };
        put(codes.size());
        out.write(reinterpret_cast<const char*>(codes.data()), std::streamsize(codes.size() * 8));
        put(records.size());
        out.write(reinterpret_cast<const char*>(records.data()), std::streamsize(records.size() * sizeof(lb::mm_record_t)));
        put(n_kmers);
        put(mm_count);
        put(coll.size());
        out.write(reinterpret_cast<const char*>(coll.data()), std::streamsize(coll.size() * sizeof(lb::kmer_t)));
        put(uniq.size());
        out.write(reinterpret_cast<const char*>(uniq.data()), std::streamsize(uniq.size() * sizeof(lb::mm_triplet_t)));
        std::printf("ok k=%u m=%u kmers=%llu codes=%zu records=%zu colliding=%zu\n", hf.get_k(), hf.get_m(),
                    (unsigned long long)hf.get_kmer_count(), codes.size(), records.size(), coll.size());
        return 0;
    } catch (std::logic_error const& e) {
        std::cerr << "shim_check: " << e.what() << "\n";
        return 4;
    } catch (std::runtime_error const& e) {
        std::cerr << "shim_check: " << e.what() << "\n";
        return 3;
    }
}
