// SHADOW of the reference's include/unpartitioned_mphf.hpp for the literal drop-in build
// (integration/dropin/build_dropin.sh), the `build-u` / `query-u` twin of partitioned_mphf.hpp in this directory:
// `lphash::mphf_alt` answers its queries on the GPU (same kernels as the partitioned function; the device image is
// built from the reference object serialized in memory) and builds on the GPU everything of build-u except the two
// PTHash constructions; save / load / statistics stay with the reference's class, which the recipe keeps under the
// name lphash::mphf_alt_reference (-Dmphf_alt=mphf_alt_reference in the reference's own translation units).
#pragma once
#ifdef LPHASH_B200_REFERENCE_TU
#include "unpartitioned_mphf_reference.hpp"
#else
#define mphf_alt mphf_alt_reference
#include "unpartitioned_mphf_reference.hpp"
#undef mphf_alt

#include <memory>
#include <ostream>
#include <string>
#include <type_traits>
#include <vector>

#include "gpu_build.hpp"

namespace lphash {

class mphf_alt {
public:
    mphf_alt() = default;

    // build-u (src/unpartitioned_mphf.cpp:31-140): Parts 1, 2 and 4 as for build-p; Part 3 = re-key by minimizer_order
    // + the two Elias-Fano sequences (positions, sizes): lphb_build_inverted_index_alt.
    void build(configuration const& config, std::ostream& res_strm) {
        stale_ = true;
        if (gpu_build::cpu_requested()) {
            ref_.build(config, res_strm);
            return;
        }
        gpu_build::Parts p;
        p.scan_and_order(config);
        if (config.verbose) std::cerr << "Part 3: build inverted index\n";
        lphb_inverted_index_alt index{};
        const uint64_t body_cap = lphb_inverted_index_bound(p.triplets.size());  // worst case; untouched beyond body_bytes
        std::unique_ptr<unsigned char[]> body(new unsigned char[body_cap]);
        uint64_t body_bytes = 0;
        gpu_build::check(lphb_build_inverted_index_alt(p.device, p.minimizer_order.data(), p.minimizer_order.size(),
                                                       p.triplets.data(), p.triplets.size(), body.get(), body_cap,
                                                       &body_bytes, &index));
        p.fallback_function(config);
        std::vector<unsigned char> image(34 + p.minimizer_order.size() + body_bytes + p.fallback.size());
        uint64_t image_bytes = 0;
        gpu_build::check(lphb_lph_assemble_alt(p.k, p.m, config.mm_seed, p.nkmers, p.triplets.size(), &index,
                                               p.minimizer_order.data(), p.minimizer_order.size(), body.get(), body_bytes,
                                               p.fallback.data(), p.fallback.size(), image.data(), image.size(),
                                               &image_bytes));
        lphash_b200::memory_loader loader(image.data(), image_bytes);
        loader.visit(ref_);
        uint64_t colliding = 0;
        for (auto const& t : p.triplets) colliding += t.size == 0;
        p.csv_line(config, res_strm, colliding, static_cast<double>(ref_.num_bits()) / ref_.get_kmer_count());
    }
    uint64_t get_minimizer_L0() const noexcept { return ref_.get_minimizer_L0(); }
    uint64_t get_kmer_count() const noexcept { return ref_.get_kmer_count(); }
    uint64_t num_bits() const noexcept { return ref_.num_bits(); }
    void print_statistics() const noexcept { ref_.print_statistics(); }

    // the hot call of query-u (src/query.cpp:52, :72) and of --check (include/mphf_utils.hpp:56, :85-86)
    template <typename MinimizerHasher = pthash::murmurhash2_64>
    std::vector<uint64_t> operator()(const char* contig, std::size_t length, bool streaming) const {
        upload();
        return gpu_(contig, length, streaming);
    }
    template <typename MinimizerHasher = pthash::murmurhash2_64>
    std::vector<uint64_t> operator()(std::string const& contig, bool streaming) const {
        return (*this)(contig.c_str(), contig.length(), streaming);
    }

    template <typename Visitor>
    void visit(Visitor& visitor) {
        ref_.visit(visitor);
        stale_ = true;
        if constexpr (std::is_same<Visitor, essentials::loader>::value) upload();
    }

    friend std::ostream& operator<<(std::ostream& os, const mphf_alt& obj) { return os << obj.ref_; }

private:
    void upload() const {
        if (!stale_) return;
        lphash_b200::memory_saver saver;
        saver.visit(const_cast<mphf_alt_reference&>(ref_));
        gpu_.load_alt(saver.bytes.data(), saver.bytes.size());
        stale_ = false;
    }
    mphf_alt_reference ref_;
    mutable lphash_b200::mphf gpu_;
    mutable bool stale_ = true;
};

}  // namespace lphash
#endif
