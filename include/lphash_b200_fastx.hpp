// Host ingest for the query / build drivers: FASTA / FASTQ text -> the batch layout the C ABI consumes
// (concatenated bases + uint64 offsets[n_records + 1]).
//
// Replaces, for batched calls, the record loop of the reference's drivers
//     while (kseq_read(seq) >= 0) hf(seq->seq.s, seq->seq.l, true);            src/query.cpp:51-55
//     while (kseq_read(seq) >= 0) minimizer::from_string(seq->seq.s, ...)      src/partitioned_mphf.cpp:70-77
// and follows kseq.h's record grammar exactly, so that record i here is the i-th `seq->seq.s`:
//   * a record starts at the first '>' or '@'; the rest of that line (name, comment) is skipped;
//   * sequence lines follow until a line whose FIRST character is '>', '@' or '+' (markers are only
//     recognised at line starts); line ends ("\n", "\r\n") are removed, nothing else is altered
//     (case, N, any other byte are preserved);
//   * '+' opens a FASTQ quality section: the rest of that line is skipped, then quality lines are
//     consumed until they hold at least as many characters as the sequence (multi-line FASTQ works);
//   * after a FASTQ record, bytes up to the next '>' or '@' are skipped.
// kseq reads through gzread one byte-buffer at a time (~9 ns/base); this works on the whole text in
// memory with memchr (a few GB/s), which is what keeps the GPU path from waiting on the parser.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef LPHASH_B200_WITH_ZLIB
#include <zlib.h>
#endif

namespace lphash_b200 {
namespace fastx {

struct Batch {
    std::vector<char> bases;         // all records' sequences, concatenated
    std::vector<uint64_t> offsets;   // n_records + 1 (offsets[0] = 0)
    uint64_t n_records() const { return offsets.empty() ? 0 : offsets.size() - 1; }
};

// Appends every record of `data[0..n)` to `out`.  Returns the number of records appended.
inline uint64_t parse(const char* data, size_t n, Batch& out) {
    if (out.offsets.empty()) out.offsets.push_back(0);
    const char* p = data;
    const char* const end = data + n;
    uint64_t added = 0;
    auto line_end = [&](const char* q) -> const char* {
        const void* e = q < end ? std::memchr(q, '\n', size_t(end - q)) : nullptr;
        return e ? static_cast<const char*>(e) : end;
    };
    // skip to the first header character
    while (p < end && *p != '>' && *p != '@') ++p;
    while (p < end) {
        // p is at a header character: skip the header line
        if (p + 1 == end) break;  // a bare header character at the very end: kseq reports end of file
        p = line_end(p);
        if (p < end) ++p;
        const size_t seq_begin = out.bases.size();
        // sequence lines
        while (p < end && *p != '>' && *p != '@' && *p != '+') {
            if (*p == '\n') {
                ++p;
                continue;
            }
            const char* e = line_end(p);
            size_t len = size_t(e - p);
            // kseq drops the '\r' of "\r\n", but only once the accumulated sequence is longer than one
            // character (ks_getuntil2: `str->l > 1`): a lone "\r" first line stays a base
            if (len && p[len - 1] == '\r' && (out.bases.size() - seq_begin) + len > 1) --len;
            out.bases.insert(out.bases.end(), p, p + len);
            p = e < end ? e + 1 : end;
        }
        const size_t seq_len = out.bases.size() - seq_begin;
        out.offsets.push_back(out.bases.size());
        ++added;
        if (p < end && *p == '+') {  // FASTQ: skip the '+' line, then quality up to the sequence length
            p = line_end(p);
            bool ok = p < end;  // kseq: no newline after '+' = "no quality string" (-2)
            if (p < end) ++p;
            size_t qual = 0;
            // at least one quality line is read, then more while it is shorter than the sequence
            while (ok && p < end) {
                const char* e = line_end(p);
                size_t len = size_t(e - p);
                if (len && p[len - 1] == '\r' && qual + len > 1) --len;  // same rule on the quality string
                qual += len;
                p = e < end ? e + 1 : end;
                if (qual >= seq_len) break;
            }
            if (!ok || qual != seq_len) {
                // kseq_read returns -2 here and the reference's `while (kseq_read(seq) >= 0)` loop ends:
                // the malformed record and everything after it are not processed
                out.bases.resize(seq_begin);
                out.offsets.pop_back();
                --added;
                break;
            }
            while (p < end && *p != '>' && *p != '@') ++p;  // to the next header character
        }
    }
    return added;
}

// Whole file (plain text, or gzip when built with -DLPHASH_B200_WITH_ZLIB -lz) -> records appended
// to `out`.  Throws std::runtime_error if the file cannot be read.
inline uint64_t read_file(std::string const& path, Batch& out) {
    std::vector<char> text;
#ifdef LPHASH_B200_WITH_ZLIB
    gzFile f = gzopen(path.c_str(), "rb");  // transparently reads uncompressed files too
    if (!f) throw std::runtime_error("cannot open " + path);
    gzbuffer(f, 1u << 20);
    size_t used = 0;
    text.resize(size_t(1) << 24);
    for (;;) {
        if (used == text.size()) text.resize(text.size() * 2);
        const size_t want = text.size() - used;
        int got = gzread(f, text.data() + used, unsigned(want > (1u << 30) ? (1u << 30) : want));
        if (got < 0) {
            gzclose(f);
            throw std::runtime_error("read error in " + path);
        }
        if (got == 0) break;
        used += size_t(got);
    }
    gzclose(f);
    text.resize(used);
#else
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    text.resize(sz > 0 ? size_t(sz) : 0);
    size_t got = text.empty() ? 0 : std::fread(text.data(), 1, text.size(), f);
    std::fclose(f);
    if (got != text.size()) throw std::runtime_error("short read in " + path);
    if (text.size() >= 2 && (unsigned char)text[0] == 0x1f && (unsigned char)text[1] == 0x8b)
        throw std::runtime_error(path + " is gzip-compressed: build with -DLPHASH_B200_WITH_ZLIB -lz");
#endif
    return parse(text.data(), text.size(), out);
}

}  // namespace fastx
}  // namespace lphash_b200
