// SHADOW of the reference's include/partitioned_mphf.hpp for the literal drop-in build
// (integration/dropin/build_dropin.sh): the reference's own src/lphash.cpp, src/query.cpp and src/build.cpp
// are compiled UNMODIFIED against this file, so `lphash::mphf` - the class `query-p`, `build-p` and
// `--check` use - answers its queries on the GPU through include/lphash_b200.hpp, while everything that is
// out of scope of the GPU path (build = PTHash construction, save / load of the `.lph` file, statistics)
// is forwarded to the reference's class, which lives on under the name lphash::mphf_reference.
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

#include <ostream>
#include <string>
#include <vector>

#include "lphash_b200.hpp"

namespace lphash {

class mphf {
public:
    mphf() = default;

    // build-p stays on the CPU (src/partitioned_mphf.cpp:33-145); the image goes to the GPU at the first query
    void build(configuration const& config, std::ostream& res_strm) {
        ref_.build(config, res_strm);
        stale_ = true;
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

    // essentials::load / essentials::save walk the reference's own field order (the file format); a load
    // invalidates the device image
    template <typename Visitor>
    void visit(Visitor& visitor) {
        ref_.visit(visitor);
        stale_ = true;
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
