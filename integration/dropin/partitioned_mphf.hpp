// SHADOW of the reference's include/partitioned_mphf.hpp for the literal drop-in build
// (integration/dropin/build_dropin.sh): the reference's own src/lphash.cpp, src/query.cpp and src/build.cpp
// are compiled UNMODIFIED against this file, so `lphash::mphf` - the class `query-p`, `build-p` and
// `--check` use - answers its queries on the GPU through include/lphash_b200.hpp and builds on the GPU
// everything of build-p except the two PTHash constructions (out of scope of the GPU path: they run on the
// CPU through the reference's own pthash headers, exactly as src/partitioned_mphf.cpp:147-160 calls them).
// Save / load of the `.lph` file and the statistics are forwarded to the reference's class, which lives on
// under the name lphash::mphf_reference and receives the GPU-built image through essentials' visitor.
//
// How the two coexist: the build recipe keeps the original header as partitioned_mphf_reference.hpp and
// compiles the reference's own translation units with -DLPHASH_B200_REFERENCE_TU -Dmphf=mphf_reference, so
// in them every `mphf` token names the renamed class; src/lphash.cpp sees both classes.
#pragma once
#ifdef LPHASH_B200_REFERENCE_TU
#include "partitioned_mphf_reference.hpp"
#else
#define mphf mphf_reference
#include "partitioned_mphf_reference.hpp"
#undef mphf

#include <cstdlib>
#include <iostream>
#include <ostream>
#include <string>
#include <type_traits>
#include <vector>

#include "lphash_b200.hpp"
#include "lphash_b200_fastx.hpp"

namespace lphash {

class mphf {
public:
    mphf() = default;

    // build-p (src/partitioned_mphf.cpp:33-145).  Part 1 (minimizer::from_string over every record) and Part 2's
    // sort + minimizer::classify: lphb_scan_classify.  Part 3 (re-key + build_inverted_index):
    // lphb_build_inverted_index.  Part 4's k-mer extraction: lphb_colliding_kmers.  The two PTHash functions are
    // built on the CPU by the reference's own calls; lphb_lph_assemble lays out the serialized object, which the
    // reference's class then loads - so save, num_bits and print_statistics see exactly what they would after the
    // CPU build.  LPHASH_B200_CPU_BUILD=1 forwards the whole build to the reference instead.
    void build(configuration const& config, std::ostream& res_strm) {
        stale_ = true;
        if (const char* e = std::getenv("LPHASH_B200_CPU_BUILD"); e && e[0] == '1') {
            ref_.build(config, res_strm);
            return;
        }
        const int device = 0;
        const uint32_t k = uint32_t(config.k), m = uint32_t(config.m);
        auto check = [](int rc) {
            if (rc != LPHB_OK) throw std::runtime_error(std::string("lphash_b200: ") + lphb_last_error());
        };
        pthash::build_configuration cfg;  // src/partitioned_mphf.cpp:45-52
        cfg.minimal_output = true;
        cfg.seed = constants::default_pthash_seed;
        cfg.c = config.c;
        cfg.alpha = 0.94;
        cfg.verbose_output = config.verbose;
        cfg.num_threads = config.num_threads;
        cfg.ram = static_cast<uint64_t>(config.max_memory) * essentials::GB;
        cfg.tmp_dir = config.tmp_dirname;

        if (config.verbose) std::cerr << "Part 1: file reading and info gathering\n";
        lphash_b200::fastx::Batch input;
        lphash_b200::fastx::read_file(config.input_filename.c_str(), input);
        const uint64_t n_records = input.n_records();
        uint64_t cap = 1;
        for (uint64_t c = 0; c < n_records; ++c) {
            const uint64_t len = input.offsets[c + 1] - input.offsets[c];
            if (len >= k) cap += len - k + 1;
        }
        std::vector<lphash_b200::mm_triplet_t> triplets(cap);
        std::vector<uint64_t> coll_ids(cap);
        uint64_t mm_count = 0, n_triplets = 0, n_ids = 0, nkmers = 0;
        check(lphb_scan_classify(device, k, m, config.mm_seed, input.bases.data(), input.offsets.data(), n_records, &mm_count,
                                 triplets.data(), cap, &n_triplets, coll_ids.data(), cap, &n_ids, &nkmers));
        triplets.resize(n_triplets);
        coll_ids.resize(n_ids);

        if (config.verbose) std::cerr << "Part 2: build MPHF\n";
        std::vector<unsigned char> minimizer_order, fallback;
        {
            std::vector<uint64_t> keys(n_triplets);
            for (uint64_t i = 0; i < n_triplets; ++i) keys[i] = triplets[i].itself;
            pthash_minimizers_mphf_t f;
            f.build_in_external_memory(keys.begin(), n_triplets, cfg);  // src/partitioned_mphf.cpp:147-153
            lphash_b200::memory_saver saver;
            saver.visit(f);
            minimizer_order.swap(saver.bytes);
        }

        if (config.verbose) std::cerr << "Part 3: build inverted index\n";
        lphb_inverted_index index{};
        std::vector<unsigned char> body(lphb_inverted_index_bound(n_triplets));
        uint64_t body_bytes = 0;
        check(lphb_build_inverted_index(device, k, m, minimizer_order.data(), minimizer_order.size(), triplets.data(),
                                        n_triplets, body.data(), body.size(), &body_bytes, &index));
        body.resize(body_bytes);

        if (config.verbose) std::cerr << "Part 4: build fallback MPHF\n";
        {
            uint64_t n_coll_kmers = 0, mm_again = 0;
            uint64_t kcap = 0;  // a colliding super-k-mer holds at most k - m + 1 k-mers
            kcap = n_ids * (uint64_t(k) - m + 1) + 1;
            std::vector<kmer_t> kmers(kcap);
            check(lphb_colliding_kmers(device, k, m, config.mm_seed, input.bases.data(), input.offsets.data(), n_records,
                                       &mm_again, coll_ids.data(), n_ids, int(sizeof(kmer_t) * 8), kmers.data(), kcap,
                                       &n_coll_kmers));
            kmers.resize(n_coll_kmers);
            pthash_fallback_mphf_t f;
            f.build_in_external_memory(kmers.begin(), n_coll_kmers, cfg);  // src/partitioned_mphf.cpp:155-160
            lphash_b200::memory_saver saver;
            saver.visit(f);
            fallback.swap(saver.bytes);
        }

        std::vector<unsigned char> image(58 + minimizer_order.size() + body.size() + fallback.size());
        uint64_t image_bytes = 0;
        check(lphb_lph_assemble(k, m, config.mm_seed, nkmers, n_triplets, &index, minimizer_order.data(),
                                minimizer_order.size(), body.data(), body.size(), fallback.data(), fallback.size(),
                                image.data(), image.size(), &image_bytes));
        lphash_b200::memory_loader loader(image.data(), image_bytes);
        loader.visit(ref_);

        // the CSV line of src/partitioned_mphf.cpp:137-144
        const uint64_t total_minimizers = (n_triplets - index.colliding_minimizers) + n_ids;  // records of Part 1
        const uint64_t total_contigs = n_records ? n_records - 1 : 0;
        res_strm << config.input_filename << "," << static_cast<uint32_t>(k) << "," << static_cast<uint32_t>(m) << ","
                 << static_cast<double>(n_ids) / n_triplets << "," << 2.0 / ((k - m + 1) + 1) << ","
                 << static_cast<double>(total_minimizers) / nkmers << "," << static_cast<double>(total_contigs) / nkmers
                 << "," << static_cast<double>(ref_.num_bits()) / nkmers;
        res_strm << "\n";
    }
    uint64_t get_minimizer_L0() const noexcept { return ref_.get_minimizer_L0(); }
    uint64_t get_kmer_count() const noexcept { return ref_.get_kmer_count(); }
    uint64_t num_bits() const noexcept { return ref_.num_bits(); }
    void print_statistics() const noexcept { ref_.print_statistics(); }

    // ★ the hot call of query-p (src/query.cpp:52, :72) and of --check (include/mphf_utils.hpp:56, :85-86)
    template <typename MinimizerHasher = pthash::murmurhash2_64>
    std::vector<uint64_t> operator()(const char* contig, std::size_t length, bool streaming = true) const {
        upload();
        return gpu_(contig, length, streaming);
    }
    template <typename MinimizerHasher = pthash::murmurhash2_64>
    std::vector<uint64_t> operator()(std::string const& contig, bool streaming = true) const {
        return (*this)(contig.c_str(), contig.length(), streaming);
    }

    // essentials::load / essentials::save walk the reference's own field order (the file format).  A load replaces
    // the function: the device image is rebuilt right away, so that - as with the reference, whose load is all the
    // set-up there is - the first query of a timed loop (src/query.cpp:48-56) does not pay for it.
    template <typename Visitor>
    void visit(Visitor& visitor) {
        ref_.visit(visitor);
        stale_ = true;
        if constexpr (std::is_same<Visitor, essentials::loader>::value) upload();
    }

    friend std::ostream& operator<<(std::ostream& os, const mphf& obj) { return os << obj.ref_; }

private:
    // the reference object serialized in memory (byte for byte its `.lph` file) -> device image
    void upload() const {
        if (!stale_) return;
        lphash_b200::memory_saver saver;
        saver.visit(const_cast<mphf_reference&>(ref_));
        gpu_.load(saver.bytes.data(), saver.bytes.size());
        stale_ = false;
    }
    mphf_reference ref_;
    mutable lphash_b200::mphf gpu_;
    mutable bool stale_ = true;
};

}  // namespace lphash
#endif
