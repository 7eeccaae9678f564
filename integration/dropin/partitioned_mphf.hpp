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
#include <memory>
#include <ostream>
#include <string>
#include <type_traits>
#include <vector>

#include "gpu_build.hpp"

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
        if (gpu_build::cpu_requested()) {
            ref_.build(config, res_strm);
            return;
        }
        gpu_build::Parts p;
        p.scan_and_order(config);  // Parts 1 + 2
        if (config.verbose) std::cerr << "Part 3: build inverted index\n";
        lphb_inverted_index index{};
        const uint64_t body_cap = lphb_inverted_index_bound(p.triplets.size());  // worst case; untouched beyond body_bytes
        std::unique_ptr<unsigned char[]> body(new unsigned char[body_cap]);
        uint64_t body_bytes = 0;
        gpu_build::check(lphb_build_inverted_index(p.device, p.k, p.m, p.minimizer_order.data(), p.minimizer_order.size(),
                                                   p.triplets.data(), p.triplets.size(), body.get(), body_cap,
                                                   &body_bytes, &index));
        p.fallback_function(config);  // Part 4
        std::vector<unsigned char> image(58 + p.minimizer_order.size() + body_bytes + p.fallback.size());
        uint64_t image_bytes = 0;
        gpu_build::check(lphb_lph_assemble(p.k, p.m, config.mm_seed, p.nkmers, p.triplets.size(), &index,
                                           p.minimizer_order.data(), p.minimizer_order.size(), body.get(), body_bytes,
                                           p.fallback.data(), p.fallback.size(), image.data(), image.size(), &image_bytes));
        lphash_b200::memory_loader loader(image.data(), image_bytes);
        loader.visit(ref_);
        p.csv_line(config, res_strm, index.colliding_minimizers, static_cast<double>(ref_.num_bits()) / p.nkmers);
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
