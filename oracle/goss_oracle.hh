// goss_oracle.hh -- CPU restatement of the data61/gossamer build-graph / build-kmer-set
// hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, loaded by or called from
// the product (gossamer_b200/, include/).  It exists so that tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs have a checker and a CPU timing arm.
//
// Parity status: PINNED.  (1) every known-answer value the reference's own tests hold for this
// path (tests/test_oracle_known_answers.py, with file:line); (2) byte-for-byte against the REAL
// reference -- its own Graph/KmerSet/SparseArray builders, its own readers, and the whole
// GossCmdBuildGraph / GossCmdBuildKmerSet / GossCmdTrimGraph commands -- compiled unmodified from
// /root/reference/src with a std::-only Boost shim (oracle/ref/, tests/test_oracle_vs_reference.py,
// 56 cases).  See oracle/README.md.
//
// Every function cites the reference file:line (relative to /root/reference) it restates.
// It is written from the reference's *behaviour*; the sequential writer state machines are
// necessarily the same algorithm because the on-disk bytes are the contract.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace goss_oracle {

typedef unsigned __int128 u128;

static inline u128 mk128(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | lo; }
// 4^n as a 128-bit value (0 when it does not fit; callers range-check k first)
static inline u128 pow4(uint64_t n) { return n >= 64 ? (u128)0 : (((u128)1) << (2 * n)); }
static inline uint64_t lo64(u128 x) { return (uint64_t)x; }
static inline uint64_t hi64(u128 x) { return (uint64_t)(x >> 64); }

// ---------------------------------------------------------------------------------------
// Key arithmetic
// ---------------------------------------------------------------------------------------

// Base-4 digit reversal of one 64-bit word.  Reference: Gossamer::rev, src/Utils.hh:377-396.
static inline uint64_t rev_base4(uint64_t x) {
    x = __builtin_bswap64(x);                                                     // bytes
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);  // nibbles
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);  // base pairs
    return x;
}

// Reverse complement of a k-symbol value held in the low 2k bits of a 128-bit integer.
// Reference: BigInteger<2>::reverseComplement, src/BigInteger.hh:204-217 (complement all
// 128 bits, reverse each word base-4 while swapping the words, shift right by 128-2k).
static inline u128 reverse_complement(u128 x, unsigned k) {
    u128 y = ~x;
    u128 r = mk128(rev_base4(lo64(y)), rev_base4(hi64(y)));
    unsigned sh = 128 - 2 * k;
    return sh >= 128 ? (u128)0 : (r >> sh);
}

// FNV-1a-64 over the 16 little-endian bytes (low word first).
// Reference: BigInteger::hash / wordHash, src/BigInteger.hh:528-536,572-582.
static inline uint64_t fnv_hash(u128 x) {
    uint64_t h = 14695981039346656037ULL;
    for (int i = 0; i < 16; ++i) {
        h ^= (uint64_t)(x & 0xFF);
        x >>= 8;
        h *= 1099511628211ULL;
    }
    return h;
}

// Canonical representative used by build-kmer-set: smaller hash wins, tie -> smaller value.
// Reference: position_type::normalize, src/RankSelect.hh:126-140.
static inline u128 normalize(u128 x, unsigned k) {
    u128 rc = reverse_complement(x, k);
    uint64_t h0 = fnv_hash(x), h1 = fnv_hash(rc);
    if (h0 > h1) return rc;
    if (h0 == h1 && rc < x) return rc;
    return x;
}

// ASCII -> 2-bit code, -1 for anything that is not ACGTacgt.
// Reference: GossReadBaseString::getBase, src/GossReadBaseString.hh:133-170.
static inline int base_code(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

static inline std::string kmer_to_string(unsigned k, u128 x) {
    std::string s(k, 'A');
    for (unsigned i = 0; i < k; ++i) {
        s[k - 1 - i] = "ACGT"[(unsigned)(x & 3)];
        x >>= 2;
    }
    return s;
}

// ---------------------------------------------------------------------------------------
// Line splitting and record framing
// ---------------------------------------------------------------------------------------

enum Format { FMT_FASTA = 0, FMT_FASTQ = 1, FMT_LINE = 2 };

struct ParseError : public std::runtime_error {
    explicit ParseError(const std::string& m) : std::runtime_error(m) {}
};

// A cursor over the lines of a buffer with std::getline semantics.
// Reference: PlainLineSource, src/LineSource.cc:17-48 -- a line is valid while the stream is
// good or the line is non-empty, so "a\nb" has two lines, "a\nb\n" has two, "a\n\n" has "a","".
struct LineCursor {
    const char* p; const char* end;
    const char* line; size_t len; bool ok;
    LineCursor(const char* b, size_t n) : p(b), end(b + n), line(b), len(0), ok(false) { advance(); }
    bool valid() const { return ok; }
    void advance() {
        if (p >= end) { ok = false; len = 0; return; }
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        line = p;
        if (nl) { len = (size_t)(nl - p); p = nl + 1; }
        else    { len = (size_t)(end - p); p = end; }
        ok = true;
    }
};

typedef std::function<void(const char*, size_t)> ReadSink;

// FASTA framing.  Reference: FastaParser::next, src/FastaParser.hh:51-87 -- header line must
// start with '>', every following line up to the next '>' line is appended verbatim (no '\r'
// strip); line counter starts at 0.
static inline void for_each_fasta_read(const char* buf, size_t n, const ReadSink& sink) {
    LineCursor src(buf, n);
    uint64_t line_num = 0;
    std::string seq;
    while (src.valid()) {
        if (!(src.len > 0 && src.line[0] == '>'))
            throw ParseError("expected '>' at beginning of line " + std::to_string(line_num));
        seq.clear();
        const char* single = nullptr; size_t single_len = 0; int pieces = 0;
        for (;;) {
            src.advance(); ++line_num;
            if (!src.valid()) break;
            if (src.len > 0 && src.line[0] == '>') break;
            if (src.len == 0) continue;
            if (pieces == 0) { single = src.line; single_len = src.len; }
            else {
                if (pieces == 1) seq.assign(single, single_len);
                seq.append(src.line, src.len);
            }
            ++pieces;
        }
        if (pieces <= 1) sink(single, single_len); else sink(seq.data(), seq.size());
    }
}

// FASTQ framing.  Reference: FastqParser::next, src/FastqParser.hh:62-176 -- one trailing
// '\r' is stripped from every line; sequence lines run until a line starting '@' or '+',
// which must be '+'; a non-empty '+' label must equal the '@' label; quality lines run until
// a line starts '@'/'+' *and* at least as much quality as sequence has been seen; lengths
// must match.  Line counter starts at 1.
static inline void for_each_fastq_read(const char* buf, size_t n, const ReadSink& sink) {
    LineCursor src(buf, n);
    uint64_t line_num = 1;
    std::string seq;
    auto stripped = [](const LineCursor& c) { return (c.len > 0 && c.line[c.len - 1] == '\r') ? c.len - 1 : c.len; };
    while (src.valid()) {
        size_t l = stripped(src);
        if (!(l > 0 && src.line[0] == '@'))
            throw ParseError("expected '@' at beginning of line " + std::to_string(line_num));
        const char* label = src.line + 1; size_t label_len = l - 1;
        seq.clear();
        const char* single = nullptr; size_t single_len = 0; int pieces = 0; size_t seq_len = 0;
        for (;;) {
            src.advance(); ++line_num;
            if (!src.valid())
                throw ParseError("expected sequence data or quality header at line " + std::to_string(line_num));
            l = stripped(src);
            if (l > 0 && (src.line[0] == '@' || src.line[0] == '+')) break;
            if (l == 0) continue;
            if (pieces == 0) { single = src.line; single_len = l; }
            else {
                if (pieces == 1) seq.assign(single, single_len);
                seq.append(src.line, l);
            }
            ++pieces; seq_len += l;
        }
        if (!(l > 0 && src.line[0] == '+'))
            throw ParseError("expected '+' at beginning of line " + std::to_string(line_num));
        if (l - 1 > 0 && !(l - 1 == label_len && memcmp(src.line + 1, label, label_len) == 0))
            throw ParseError("quality title does not match sequence title at line " + std::to_string(line_num));
        size_t qual_len = 0;
        for (;;) {
            src.advance(); ++line_num;
            if (!src.valid()) break;
            l = stripped(src);
            if (l > 0 && (src.line[0] == '@' || src.line[0] == '+')) {
                if (qual_len >= seq_len) break;
            }
            qual_len += l;
        }
        if (seq_len != qual_len)
            throw ParseError("length mistmatch between sequence and quality data just before line " + std::to_string(line_num));
        if (pieces <= 1) sink(single, single_len); else sink(seq.data(), seq.size());
    }
}

// One read per line, empty lines included.  Reference: LineParser::next, src/LineParser.hh:71-82.
static inline void for_each_line_read(const char* buf, size_t n, const ReadSink& sink) {
    for (LineCursor src(buf, n); src.valid(); src.advance()) sink(src.line, src.len);
}

static inline void for_each_read(const char* buf, size_t n, int fmt, const ReadSink& sink) {
    switch (fmt) {
        case FMT_FASTA: for_each_fasta_read(buf, n, sink); break;
        case FMT_FASTQ: for_each_fastq_read(buf, n, sink); break;
        case FMT_LINE:  for_each_line_read(buf, n, sink); break;
        default: throw std::runtime_error("unknown format");
    }
}

// ---------------------------------------------------------------------------------------
// Window enumeration
// ---------------------------------------------------------------------------------------

// Every window of `w` consecutive ACGT bases, first base most significant; a non-ACGT byte
// restarts the search after it.  Reference: GossReadBaseString::firstKmer/nextKmer,
// src/GossReadBaseString.hh:52-103 (+ getEdge :172-188), driven by GossRead::Iterator,
// src/GossRead.hh:57-114.
template <typename Emit>
static inline void for_each_window(const char* s, size_t len, unsigned w, Emit&& emit) {
    if (len < w) return;
    const u128 mask = (w >= 64) ? ~(u128)0 : ((((u128)1) << (2 * w)) - 1);
    u128 x = 0; unsigned run = 0;
    for (size_t i = 0; i < len; ++i) {
        int b = base_code(s[i]);
        if (b < 0) { run = 0; x = 0; continue; }
        x = ((x << 2) | (unsigned)b) & mask;
        if (++run >= w) emit(x);
    }
}

enum Mode {
    MODE_GRAPH = 0,    // x then rc(x): ReverseComplementAdapter, src/ReverseComplementAdapter.hh:34-55
    MODE_KMERSET = 1,  // normalize(x): KmerizingAdapter + src/GossCmdBuildKmerSet.tcc:248-250
    MODE_FORWARD = 2   // x only (GossRead::Iterator on its own; used by unit checks)
};

struct Input { const char* data; size_t size; int format; };

// Reads are consumed line files first, then FASTA, then FASTQ
// (src/GossCmdBuildGraph.cc:284-300); the order cannot change the counted result.
static inline void extract_keys(const std::vector<Input>& inputs, unsigned w, int mode,
                                std::vector<u128>& out, uint64_t* n_reads = nullptr) {
    uint64_t reads = 0;
    static const int order[3] = {FMT_LINE, FMT_FASTA, FMT_FASTQ};
    for (int f = 0; f < 3; ++f)
        for (const Input& in : inputs) {
            if (in.format != order[f]) continue;
            for_each_read(in.data, in.size, in.format, [&](const char* s, size_t len) {
                ++reads;
                for_each_window(s, len, w, [&](u128 x) {
                    if (mode == MODE_GRAPH) { out.push_back(x); out.push_back(reverse_complement(x, w)); }
                    else if (mode == MODE_KMERSET) out.push_back(normalize(x, w));
                    else out.push_back(x);
                });
            });
        }
    if (n_reads) *n_reads = reads;
}

// ---------------------------------------------------------------------------------------
// Counting: multiset -> sorted (key, count).  This is the *result* of BackyardHash::insert +
// sort + the duplicate-merging emit loop (src/BackyardHash.cc:115-271,
// src/GossCmdBuildGraph.cc:239-258), not its mechanism.
// ---------------------------------------------------------------------------------------

template <typename K>
static inline void sort_keys_parallel(std::vector<K>& keys, unsigned key_bits, int threads) {
    if (threads <= 1 || keys.size() < (1u << 16)) { std::sort(keys.begin(), keys.end()); return; }
    // bucket by the top 8 significant bits, then sort buckets independently
    const unsigned sh = key_bits > 8 ? key_bits - 8 : 0;
    const size_t n = keys.size();
    std::vector<size_t> start(257, 0);
    for (size_t i = 0; i < n; ++i) ++start[(size_t)((keys[i] >> sh) & 0xFF) + 1];
    for (int b = 0; b < 256; ++b) start[b + 1] += start[b];
    std::vector<K> tmp(n);
    {
        std::vector<size_t> pos(start.begin(), start.end() - 1);
        for (size_t i = 0; i < n; ++i) tmp[pos[(size_t)((keys[i] >> sh) & 0xFF)]++] = keys[i];
    }
    keys.swap(tmp);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            for (int b = t; b < 256; b += threads) std::sort(keys.begin() + start[b], keys.begin() + start[b + 1]);
        });
    for (auto& th : pool) th.join();
}

template <typename K>
static inline void run_length_reduce(const std::vector<K>& sorted, uint64_t min_count,
                                     std::vector<u128>& keys, std::vector<uint64_t>& counts) {
    size_t i = 0, n = sorted.size();
    while (i < n) {
        size_t j = i + 1;
        while (j < n && sorted[j] == sorted[i]) ++j;
        if ((uint64_t)(j - i) >= min_count) { keys.push_back((u128)sorted[i]); counts.push_back((uint64_t)(j - i)); }
        i = j;
    }
}

// ---------------------------------------------------------------------------------------
// In-memory output store (the StringFileFactory analogue, src/StringFileFactory.hh:25-75)
// ---------------------------------------------------------------------------------------

typedef std::map<std::string, std::string> MemFS;

template <typename T>
static inline void put(std::string& f, T v) { f.append(reinterpret_cast<const char*>(&v), sizeof(T)); }

// ---------------------------------------------------------------------------------------
// Writers
// ---------------------------------------------------------------------------------------

// Raw little-endian uint64[] bitmap, zero words materialised lazily.
// Reference: WordyBitVector::Builder, src/WordyBitVector.hh:54-134, src/WordyBitVector.cc:18-29.
class BitmapWriter {
public:
    explicit BitmapWriter(std::string& f) : out_(f) {}
    void set(uint64_t pos) { seek(pos); cur_ |= 1ULL << (pos & 63); pos_ = pos + 1; }
    void pad(uint64_t pos) { seek(pos + 1); }
    void finish() { spill(); }
private:
    void seek(uint64_t pos) {
        uint64_t w = pos >> 6;
        if (cur_word_ < w) { spill(); cur_word_ = w; cur_ = 0; }
        pos_ = pos;
    }
    void spill() {
        while (file_words_ < cur_word_) { put<uint64_t>(out_, 0); ++file_words_; }
        put<uint64_t>(out_, cur_); ++file_words_;
    }
    std::string& out_;
    uint64_t pos_ = 0, file_words_ = 0, cur_word_ = 0, cur_ = 0;
};

// Two-level select directory.  Reference: DenseSelect::Builder::{push_back,flush,end} and
// ctor, src/DenseArray.cc:446-694; header layout src/DenseArray.hh:98-136; constants :82-96;
// block type tags :188-196.
class DenseSelectWriter {
public:
    enum { T_SMALL = 0, T_SPILL64 = 1, T_SPILL32 = 2, T_SPILL16 = 3, T_SPILL8 = 4, T_INTERMEDIATE = 5 };
    static const uint64_t kBlock = 8192, kLogBlock = 13, kSample = 64, kLogSample = 6;
    struct Header {
        uint64_t version, flags, indexArrayOffset, rankArrayOffset;
        uint64_t logBlockSize, blockSize, logSampleRate, sampleRate;
        uint64_t numBlocks, indexSize, smallBlocks, smallBlocksSize;
        uint64_t intermediateBlocks, intermediateBlocksSize, largeBlocks, largeBlocksSize;
    };
    DenseSelectWriter(std::string& f, bool invert) : out_(f) {
        memset(&h_, 0, sizeof(h_));
        h_.version = 2012092701ULL; h_.flags = invert ? 1 : 0;
        h_.logBlockSize = kLogBlock; h_.blockSize = kBlock; h_.logSampleRate = kLogSample; h_.sampleRate = kSample;
        out_.append(reinterpret_cast<const char*>(&h_), sizeof(h_));
        out_.resize(4096, '\0');
        blk_.reserve(kBlock);
    }
    void push_back(uint64_t pos) { blk_.push_back(pos); if (blk_.size() == kBlock) flush(); }
    void finish() {
        flush();
        align(15);
        h_.indexArrayOffset = out_.size();
        for (uint64_t v : index_) put(out_, v);
        h_.indexSize += 8 * index_.size();
        h_.rankArrayOffset = out_.size();
        for (uint64_t v : rank_) put(out_, v);
        h_.indexSize += 8 * rank_.size();
        memcpy(&out_[0], &h_, sizeof(h_));
    }
private:
    void align(uint64_t mask) { while (out_.size() & mask) out_.push_back('\0'); }
    void flush() {
        if (blk_.empty()) return;
        const uint64_t at = out_.size();
        const uint64_t first = blk_.front(), span = blk_.back() - first;
        rank_.push_back(first);
        if (span >= (1ULL << 24) || blk_.size() < kBlock) {
            if (span < (1ULL << 32)) {
                for (uint64_t p : blk_) { put<uint32_t>(out_, (uint32_t)(p - first)); h_.largeBlocksSize += 4; }
                index_.push_back(at | T_SPILL32);
            } else {
                for (uint64_t p : blk_) { put<uint64_t>(out_, p); h_.largeBlocksSize += 8; }   // absolute, by design
                index_.push_back(at | T_SPILL64);
            }
            ++h_.largeBlocks;
        } else if (span >= (1ULL << 16)) {
            const size_t ns = blk_.size() / kSample;
            std::vector<uint32_t> sub_span(ns);
            std::vector<uint16_t> ptr(ns);
            for (size_t s = 0; s < ns; ++s) {
                sub_span[s] = (uint32_t)(blk_[s * kSample + kSample - 1] - blk_[s * kSample]);
                put<uint32_t>(out_, (uint32_t)(blk_[s * kSample] - first));
                h_.intermediateBlocksSize += 4;
            }
            uint64_t base = (ns * (4 + 2) + 7) & ~7ULL;
            for (size_t s = 0; s < ns; ++s) {
                uint16_t ip = 0;
                if (sub_span[s] <= (kBlock >> kLogSample)) ip = T_SMALL;
                else if (sub_span[s] < (1u << 8))  { ip = (uint16_t)base | T_SPILL8;  base += kSample * 1; }
                else if (sub_span[s] < (1u << 16)) { ip = (uint16_t)base | T_SPILL16; base += kSample * 2; }
                else                               { ip = (uint16_t)base | T_SPILL32; base += kSample * 4; }
                ptr[s] = ip;
                put<uint16_t>(out_, ip);
                h_.intermediateBlocksSize += 2;
                base = (base + 7) & ~7ULL;
            }
            if (base > (1ULL << 16)) throw std::runtime_error("Intermediate sub-blocks too large");
            for (size_t s = 0; s < ns; ++s) {
                if (!ptr[s]) continue;
                const uint64_t origin = blk_[s * kSample];
                align(7);
                for (size_t j = s * kSample; j < (s + 1) * kSample; ++j) {
                    uint64_t d = blk_[j] - origin;
                    switch (ptr[s] & 7) {
                        case T_SPILL8:  put<uint8_t>(out_, (uint8_t)d);   h_.intermediateBlocksSize += 1; break;
                        case T_SPILL16: put<uint16_t>(out_, (uint16_t)d); h_.intermediateBlocksSize += 2; break;
                        case T_SPILL32: put<uint32_t>(out_, (uint32_t)d); h_.intermediateBlocksSize += 4; break;
                    }
                }
            }
            index_.push_back(at | T_INTERMEDIATE);
            ++h_.intermediateBlocks;
        } else {
            for (size_t s = 0; s < blk_.size(); s += kSample) { put<uint16_t>(out_, (uint16_t)(blk_[s] - first)); h_.smallBlocksSize += 2; }
            index_.push_back(at | T_SMALL);
            ++h_.smallBlocks;
        }
        blk_.clear();
        align(7);
        ++h_.numBlocks;
    }
    std::string& out_;
    Header h_;
    std::vector<uint64_t> blk_, index_, rank_;
};

// Fixed-width integers split into nested native-type planes.
// Reference: IntegerArray::builder, src/IntegerArray.cc:259-357; StackedArray::Builder,
// src/StackedArray.hh:152-178 (upr = value >> bits(lwr)); MappedArray::Builder,
// src/MappedArray.hh:65-88 (raw native values, no header).
struct PlaneSpec { std::string suffix; unsigned shift; unsigned bytes; };

static inline std::vector<PlaneSpec> integer_array_planes(unsigned bits) {
    switch (bits) {
        case 8:   return {{"", 0, 1}};
        case 16:  return {{"", 0, 2}};
        case 24:  return {{".upr", 16, 1}, {".lwr", 0, 2}};
        case 32:  return {{"", 0, 4}};
        case 40:  return {{".upr", 32, 1}, {".lwr", 0, 4}};
        case 48:  return {{".upr", 32, 2}, {".lwr", 0, 4}};
        case 56:  return {{".upr", 48, 1}, {".lwr.upr", 32, 2}, {".lwr.lwr", 0, 4}};
        case 64:  return {{"", 0, 8}};
        case 72:  return {{".upr", 64, 1}, {".lwr", 0, 8}};
        case 80:  return {{".upr", 64, 2}, {".lwr", 0, 8}};
        case 88:  return {{".upr", 80, 1}, {".lwr.upr", 64, 2}, {".lwr.lwr", 0, 8}};
        case 96:  return {{".upr", 64, 4}, {".lwr", 0, 8}};
        case 104: return {{".upr", 96, 1}, {".lwr.upr", 64, 4}, {".lwr.lwr", 0, 8}};
        case 112: return {{".upr", 96, 2}, {".lwr.upr", 64, 4}, {".lwr.lwr", 0, 8}};
        case 120: return {{".upr.upr", 112, 1}, {".upr.lwr", 96, 2}, {".lwr.upr", 64, 4}, {".lwr.lwr", 0, 8}};
        case 128: return {{".upr", 64, 8}, {".lwr", 0, 8}};
        default: throw std::runtime_error("IntegerArray::builder: unsupported integer width " + std::to_string(bits));
    }
}

class IntegerArrayWriter {
public:
    IntegerArrayWriter(MemFS& fs, const std::string& base, unsigned bits) : planes_(integer_array_planes(bits)) {
        for (const PlaneSpec& p : planes_) files_.push_back(&fs[base + p.suffix]);
    }
    void push_back(u128 v) {
        for (size_t i = 0; i < planes_.size(); ++i) {
            u128 piece = v >> planes_[i].shift;
            files_[i]->append(reinterpret_cast<const char*>(&piece), planes_[i].bytes);   // little-endian truncation
        }
    }
private:
    std::vector<PlaneSpec> planes_;
    std::vector<std::string*> files_;
};

// Split parameter.  Reference: SparseArray::Builder::d, src/SparseArray.cc:47-72 (double
// arithmetic, ceil, clamp to [8,128]; the cast of a negative/NaN double is whatever the
// x86-64 cvttsd2si path gives, which the clamp turns into 128 or 8 -- k>=15 in all tests).
static inline uint64_t sparse_array_d(u128 n_universe, uint64_t m_est) {
    // BigInteger::asDouble sums word*2^64 terms in double, src/BigInteger.hh
    double n = (double)hi64(n_universe) * 18446744073709551616.0 + (double)lo64(n_universe);
    double m = (double)m_est;
    double d0 = log2(n / ((1 + m) * 1.4426950408889634));
    uint64_t d = (uint64_t)ceil(d0);
    if (d < 8) d = 8; else if (d > 128) d = 128;
    return d;
}

// Elias-Fano set.  Reference: SparseArray::Builder::{push_back,end}, src/SparseArray.hh:87-118,
// src/SparseArray.cc:75-117; header src/SparseArray.hh:60-72, src/SparseArray.cc:11-15.
class SparseArrayWriter {
public:
    struct Header { uint64_t version, D, quantizedD; uint64_t dmask[2]; uint64_t size[2]; uint64_t count; };
    SparseArrayWriter(MemFS& fs, const std::string& base, u128 n_universe, uint64_t m_est)
        : SparseArrayWriter(fs, base, sparse_array_d(n_universe, m_est)) {}
    SparseArrayWriter(MemFS& fs, const std::string& base, uint64_t D)
        : D_(D), qD_(8 * ((D + 7) / 8)), mask_(D >= 128 ? ~(u128)0 : ((((u128)1) << D) - 1)),
          high_(fs[base + ".high-bits"]), d0_(fs[base + "-d0"], true), d1_(fs[base + "-d1"], false),
          low_(fs, base + ".low-bits", (unsigned)qD_), header_(fs[base + ".header"]) {}
    void push_back(u128 pos) {
        u128 nd = D_ >= 128 ? (u128)0 : (pos >> D_);
        if (hi64(nd)) throw std::runtime_error("SparseArray::end()");
        uint64_t h = lo64(nd) + count_;
        high_.set(h);
        while (next_bit_ < h) d0_.push_back(next_bit_++);
        d1_.push_back(h);
        next_bit_ = h + 1;
        low_.push_back(pos & mask_);
        ++count_;
    }
    void finish(u128 n_universe) {
        u128 nd = D_ >= 128 ? (u128)0 : (n_universe >> D_);
        if (hi64(nd)) throw std::runtime_error("Internal error in SparseArray; nd too large");
        uint64_t h = lo64(nd) + count_ + 2;
        while (next_bit_ < h) d0_.push_back(next_bit_++);
        high_.pad(next_bit_);
        high_.finish();
        d0_.finish();
        d1_.finish();
        Header hd;
        hd.version = 2012030501ULL; hd.D = D_; hd.quantizedD = qD_;
        hd.dmask[0] = lo64(mask_); hd.dmask[1] = hi64(mask_);
        hd.size[0] = lo64(n_universe); hd.size[1] = hi64(n_universe);
        hd.count = count_;
        header_.append(reinterpret_cast<const char*>(&hd), sizeof(hd));
    }
    uint64_t D() const { return D_; }
private:
    uint64_t D_, qD_;
    u128 mask_;
    BitmapWriter high_;
    DenseSelectWriter d0_, d1_;
    IntegerArrayWriter low_;
    std::string& header_;
    uint64_t count_ = 0, next_bit_ = 0;
};

// Counts.  Reference: VariableByteArray::Builder, src/VariableByteArray.hh:76-118,
// src/VariableByteArray.cc:21-43 (both presence sets sized N=numItems, M=floor(0.001*numItems);
// the pFrac argument is ignored).
class VariableByteArrayWriter {
public:
    VariableByteArrayWriter(MemFS& fs, const std::string& base, uint64_t num_items)
        : ord0_(fs[base + ".ord0"]),
          ord1p_(fs, base + ".ord1p", (u128)num_items, (uint64_t)(num_items * 0.001)),
          ord1_(fs[base + ".ord1"]),
          ord2p_(fs, base + ".ord2p", (u128)num_items, (uint64_t)(num_items * 0.001)),
          ord2_(fs[base + ".ord2"]) {}
    void push_back(uint32_t v) {
        uint64_t pos = n0_++;
        put<uint8_t>(ord0_, (uint8_t)(v & 0xFF));
        if (!(v >>= 8)) return;
        ord1p_.push_back((u128)pos);
        pos = n1_++;
        put<uint8_t>(ord1_, (uint8_t)(v & 0xFF));
        if (!(v >>= 8)) return;
        ord2p_.push_back((u128)pos);
        put<uint16_t>(ord2_, (uint16_t)(v & 0xFFFF));
    }
    void finish() { ord1p_.finish((u128)n0_); ord2p_.finish((u128)n1_); }
private:
    std::string& ord0_;
    SparseArrayWriter ord1p_;
    std::string& ord1_;
    SparseArrayWriter ord2p_;
    std::string& ord2_;
    uint64_t n0_ = 0, n1_ = 0;
};

// Graph.  Reference: Graph::Builder ctor/push_back/end, src/Graph.cc:115-167,
// src/Graph.hh:95-127; header src/Graph.hh:73-83.  The count histogram uses the 64-bit
// count, the counts array the value truncated to uint32 (src/BackgroundBlockConsumer.hh:20).
class GraphWriter {
public:
    static const uint64_t kMaxK = 62;
    GraphWriter(MemFS& fs, const std::string& base, uint64_t k, uint64_t m_est)
        : fs_(fs), base_(base), k_(k),
          edges_(fs, base + "-edges", pow4(check_k(k) + 1), m_est),
          counts_(fs, base + "-counts", m_est) {
        std::string& h = fs[base + ".header"];
        put<uint64_t>(h, 2011101014ULL); put<uint64_t>(h, k); put<uint64_t>(h, 0);
    }
    static uint64_t check_k(uint64_t k) {
        if (k > kMaxK) throw std::runtime_error("unable to build a graph with k=" + std::to_string(k));
        return k;
    }
    void push_back(u128 edge, uint64_t count) {
        edges_.push_back(edge);
        counts_.push_back((uint32_t)count);
        ++hist_[count];
    }
    void finish() {
        edges_.finish(pow4(k_ + 1));
        counts_.finish();
        std::string& t = fs_[base_ + "-counts-hist.txt"];
        for (auto& kv : hist_) t += std::to_string(kv.first) + "\t" + std::to_string(kv.second) + "\n";
    }
private:
    MemFS& fs_; std::string base_; uint64_t k_;
    SparseArrayWriter edges_;
    VariableByteArrayWriter counts_;
    std::map<uint64_t, uint64_t> hist_;
};

// KmerSet.  Reference: KmerSet::Builder, src/KmerSet.hh:61-103; header :32-43 (written at end()).
class KmerSetWriter {
public:
    static const uint64_t kMaxK = 63;
    KmerSetWriter(MemFS& fs, const std::string& base, uint64_t k, uint64_t m_est)
        : fs_(fs), base_(base), k_(k), kmers_(fs, base + ".kmers", pow4(check_k(k)), m_est) {}
    static uint64_t check_k(uint64_t k) {
        if (k > kMaxK) throw std::runtime_error("unable to build a graph with k=" + std::to_string(k));
        return k;
    }
    void push_back(u128 kmer) { kmers_.push_back(kmer); ++count_; }
    void finish() {
        kmers_.finish(pow4(k_));
        std::string& h = fs_[base_ + ".header"];
        put<uint64_t>(h, 2011101701ULL); put<uint64_t>(h, k_); put<uint64_t>(h, count_);
    }
private:
    MemFS& fs_; std::string base_; uint64_t k_;
    SparseArrayWriter kmers_;
    uint64_t count_ = 0;
};

// ---------------------------------------------------------------------------------------
// Readers (restated so that reference-style code is shown to open what we write)
// ---------------------------------------------------------------------------------------

static inline const std::string& fs_get(const MemFS& fs, const std::string& name) {
    auto it = fs.find(name);
    if (it == fs.end()) throw std::runtime_error("missing file " + name);
    return it->second;
}

// Reference: WordyBitVector::select, src/WordyBitVector.tcc:17-54.
class BitmapReader {
public:
    explicit BitmapReader(const std::string& f) : w_(reinterpret_cast<const uint64_t*>(f.data())), n_(f.size() / 8) {}
    uint64_t words() const { return n_; }
    bool get(uint64_t pos) const { return (w_[pos >> 6] >> (pos & 63)) & 1; }
    uint64_t select(bool invert, uint64_t from, uint64_t count) const {
        uint64_t w = from >> 6, b = from & 63;
        if (w >= n_) throw std::runtime_error("WordyBitVector::select out of range");
        uint64_t x = (invert ? ~w_[w] : w_[w]) >> b;
        uint64_t c = count, p = (uint64_t)__builtin_popcountll(x);
        while (c >= p) {
            c -= p; ++w; b = 0;
            if (w >= n_) throw std::runtime_error("WordyBitVector::select out of range");
            x = invert ? ~w_[w] : w_[w];
            p = (uint64_t)__builtin_popcountll(x);
        }
        for (uint64_t i = 0; i < c; ++i) x &= x - 1;
        return w * 64 + b + (uint64_t)__builtin_ctzll(x);
    }
private:
    const uint64_t* w_; uint64_t n_;
};

// Reference: DenseSelect ctor checks + select + lookupSubBlock, src/DenseArray.cc:36-91,135-248.
class DenseSelectReader {
public:
    DenseSelectReader(const BitmapReader& bits, const std::string& f, bool invert) : bits_(bits), d_(reinterpret_cast<const uint8_t*>(f.data())), invert_(invert) {
        if (f.size() < 4096) throw std::runtime_error("DenseSelect file too small");
        memcpy(&h_, d_, sizeof(h_));
        if (h_.version != 2012092701ULL) throw std::runtime_error("DenseSelect version mismatch");
        if ((1ULL << h_.logBlockSize) != h_.blockSize || (1ULL << h_.logSampleRate) != h_.sampleRate ||
            h_.smallBlocks + h_.intermediateBlocks + h_.largeBlocks != h_.numBlocks)
            throw std::runtime_error("Corrupt DenseSelect index header");
        if ((h_.flags & 1) != (invert ? 1u : 0u)) throw std::runtime_error("DenseSelect index does not have the expected sense");
        if (h_.flags >> 1) throw std::runtime_error("Reserved DenseSelect flag set");
        if (h_.rankArrayOffset + 8 * h_.numBlocks > f.size()) throw std::runtime_error("DenseSelect truncated");
        index_ = reinterpret_cast<const uint64_t*>(d_ + h_.indexArrayOffset);
        rank_ = reinterpret_cast<const uint64_t*>(d_ + h_.rankArrayOffset);
    }
    const DenseSelectWriter::Header& header() const { return h_; }
    uint64_t select(uint64_t i) const {
        uint64_t b = i >> h_.logBlockSize;
        if (b >= h_.numBlocks) throw std::runtime_error("DenseSelect::select out of range");
        uint64_t start = rank_[b], il = index_[b];
        const uint8_t* blk = d_ + (il & ~7ULL);
        uint64_t within = i & (h_.blockSize - 1), sb = within >> h_.logSampleRate, r = i & (h_.sampleRate - 1);
        switch (il & 7) {
            case DenseSelectWriter::T_SMALL:
                return bits_.select(invert_, start + reinterpret_cast<const uint16_t*>(blk)[sb], r);
            case DenseSelectWriter::T_SPILL64: return reinterpret_cast<const uint64_t*>(blk)[within];
            case DenseSelectWriter::T_SPILL32: return start + reinterpret_cast<const uint32_t*>(blk)[within];
            case DenseSelectWriter::T_SPILL16: return start + reinterpret_cast<const uint16_t*>(blk)[within];
            case DenseSelectWriter::T_SPILL8:  return start + blk[within];
            case DenseSelectWriter::T_INTERMEDIATE: {
                const uint32_t* s = reinterpret_cast<const uint32_t*>(blk);
                const uint16_t* ptrs = reinterpret_cast<const uint16_t*>(blk + (4ULL << (h_.logBlockSize - h_.logSampleRate)));
                uint64_t origin = start + s[sb];
                uint16_t ip = ptrs[sb];
                if (!ip) return bits_.select(invert_, origin, r);
                const uint8_t* sub = blk + (ip & ~7u);
                switch (ip & 7) {
                    case DenseSelectWriter::T_SPILL32: return origin + reinterpret_cast<const uint32_t*>(sub)[r];
                    case DenseSelectWriter::T_SPILL16: return origin + reinterpret_cast<const uint16_t*>(sub)[r];
                    case DenseSelectWriter::T_SPILL8:  return origin + sub[r];
                    default: throw std::runtime_error("Corrupt DenseSelect index (intermediate-level)");
                }
            }
            default: throw std::runtime_error("Corrupt DenseSelect index (top-level)");
        }
    }
private:
    const BitmapReader& bits_;
    const uint8_t* d_;
    bool invert_;
    DenseSelectWriter::Header h_;
    const uint64_t* index_; const uint64_t* rank_;
};

// Reference: StackedArray::operator[], src/StackedArray.hh:217-245 (value = upr << bits(lwr) | lwr).
class IntegerArrayReader {
public:
    IntegerArrayReader(const MemFS& fs, const std::string& base, unsigned bits) : planes_(integer_array_planes(bits)) {
        for (const PlaneSpec& p : planes_) files_.push_back(&fs_get(fs, base + p.suffix));
        size_ = files_[0]->size() / planes_[0].bytes;
        for (size_t i = 0; i < planes_.size(); ++i)
            if (files_[i]->size() != size_ * planes_[i].bytes) throw std::runtime_error("IntegerArray planes disagree on length: " + base);
    }
    uint64_t size() const { return size_; }
    u128 operator[](uint64_t i) const {
        u128 v = 0;
        for (size_t p = 0; p < planes_.size(); ++p) {
            u128 piece = 0;
            memcpy(&piece, files_[p]->data() + i * planes_[p].bytes, planes_[p].bytes);
            v |= piece << planes_[p].shift;
        }
        return v;
    }
private:
    std::vector<PlaneSpec> planes_;
    std::vector<const std::string*> files_;
    uint64_t size_;
};

// Reference: SparseArray ctor, select, rank, accessAndRank, findLowOrderGroup,
// src/SparseArray.cc:175-194, src/SparseArray.hh:246-364.
class SparseArrayReader {
public:
    SparseArrayReader(const MemFS& fs, const std::string& base)
        : hd_(read_header(fs_get(fs, base + ".header"))), high_(fs_get(fs, base + ".high-bits")),
          d0_(high_, fs_get(fs, base + "-d0"), true), d1_(high_, fs_get(fs, base + "-d1"), false),
          low_(fs, base + ".low-bits", (unsigned)hd_.quantizedD) {
        if (low_.size() != hd_.count) throw std::runtime_error("SparseArray low-bits length != count: " + base);
    }
    uint64_t count() const { return hd_.count; }
    u128 size() const { return mk128(hd_.size[1], hd_.size[0]); }
    uint64_t D() const { return hd_.D; }
    u128 select(uint64_t r) const {
        u128 pos = 0;
        if (hd_.D < 128) { pos = (u128)(d1_.select(r) - r); pos <<= hd_.D; }
        return pos | low_[r];
    }
    uint64_t rank(u128 pos) const {
        if (pos >= size()) return hd_.count;
        uint64_t lo, hi; group(pos, lo, hi);
        return lower_bound(lo, hi, pos & mask());
    }
    bool access_and_rank(u128 pos, uint64_t& r) const {
        uint64_t lo, hi; group(pos, lo, hi);
        r = lower_bound(lo, hi, pos & mask());
        return r < hi && low_[r] == (pos & mask());
    }
    // sequential decode straight off the bitmap (SparseArray::LazyIterator, src/SparseArray.hh:179-225)
    std::vector<u128> decode_all() const {
        std::vector<u128> out; out.reserve(hd_.count);
        uint64_t i = 0;
        for (uint64_t w = 0; w < high_.words() && i < hd_.count; ++w)
            for (uint64_t b = 0; b < 64 && i < hd_.count; ++b)
                if (high_.get(w * 64 + b)) {
                    u128 pos = hd_.D < 128 ? ((u128)(w * 64 + b - i) << hd_.D) : (u128)0;
                    out.push_back(pos | low_[i]); ++i;
                }
        if (i != hd_.count) throw std::runtime_error("SparseArray bitmap has fewer ones than count");
        return out;
    }
    const DenseSelectReader& d0() const { return d0_; }
    const DenseSelectReader& d1() const { return d1_; }
    const BitmapReader& high() const { return high_; }
private:
    static SparseArrayWriter::Header read_header(const std::string& f) {
        SparseArrayWriter::Header h;
        if (f.size() != sizeof(h)) throw std::runtime_error("SparseArray header size");
        memcpy(&h, f.data(), sizeof(h));
        if (h.version != 2012030501ULL) throw std::runtime_error("SparseArray version mismatch");
        return h;
    }
    u128 mask() const { return mk128(hd_.dmask[1], hd_.dmask[0]); }
    void group(u128 pos, uint64_t& lo, uint64_t& hi) const {
        if (hd_.D >= 128) { lo = 0; hi = low_.size(); return; }
        uint64_t pd = lo64(pos >> hd_.D);
        if (!pd) { lo = 0; hi = d0_.select(0); return; }
        uint64_t a = d0_.select(pd - 1) + 1, b = d0_.select(pd);
        lo = a >= pd ? a - pd : 0; hi = b >= pd ? b - pd : 0;
    }
    uint64_t lower_bound(uint64_t lo, uint64_t hi, u128 v) const {
        while (lo < hi) { uint64_t mid = lo + (hi - lo) / 2; if (low_[mid] < v) lo = mid + 1; else hi = mid; }
        return lo;
    }
    SparseArrayWriter::Header hd_;
    BitmapReader high_;
    DenseSelectReader d0_, d1_;
    IntegerArrayReader low_;
};

// Reference: VariableByteArray::operator[], src/VariableByteArray.hh:227-247.
class VariableByteArrayReader {
public:
    VariableByteArrayReader(const MemFS& fs, const std::string& base)
        : ord0_(fs_get(fs, base + ".ord0")), p1_(fs, base + ".ord1p"), ord1_(fs_get(fs, base + ".ord1")),
          p2_(fs, base + ".ord2p"), ord2_(fs_get(fs, base + ".ord2")) {}
    uint64_t size() const { return ord0_.size(); }
    uint32_t operator[](uint64_t i) const {
        uint32_t v = (uint8_t)ord0_[i];
        uint64_t r1;
        if (!p1_.access_and_rank((u128)i, r1)) return v;
        v |= (uint32_t)(uint8_t)ord1_[r1] << 8;
        uint64_t r2;
        if (!p2_.access_and_rank((u128)r1, r2)) return v;
        uint16_t top; memcpy(&top, ord2_.data() + 2 * r2, 2);
        return v | ((uint32_t)top << 16);
    }
private:
    const std::string& ord0_;
    SparseArrayReader p1_;
    const std::string& ord1_;
    SparseArrayReader p2_;
    const std::string& ord2_;
};

// Reference: Graph::open / getAndVerifyHeader / LazyIterator, src/Graph.cc:89-112,195-216,366-396.
struct GraphContents { uint64_t k; std::vector<u128> edges; std::vector<uint32_t> counts; uint64_t hist_total; };

static inline GraphContents read_graph(const MemFS& fs, const std::string& base, bool exercise_select = true) {
    GraphContents g;
    const std::string& h = fs_get(fs, base + ".header");
    if (h.size() != 24) throw std::runtime_error("Graph header size");
    uint64_t ver; memcpy(&ver, h.data(), 8); memcpy(&g.k, h.data() + 8, 8);
    if (ver != 2011101014ULL) throw std::runtime_error("Graph version mismatch");
    SparseArrayReader edges(fs, base + "-edges");
    VariableByteArrayReader counts(fs, base + "-counts");
    g.edges = edges.decode_all();
    if (counts.size() != g.edges.size()) throw std::runtime_error("counts/edges length mismatch");
    g.counts.resize(g.edges.size());
    for (uint64_t i = 0; i < g.edges.size(); ++i) g.counts[i] = counts[i];
    if (exercise_select)
        for (uint64_t i = 0; i < g.edges.size(); ++i) {
            if (edges.select(i) != g.edges[i]) throw std::runtime_error("select(i) disagrees with bitmap decode at " + std::to_string(i));
            if (edges.rank(g.edges[i]) != i) throw std::runtime_error("rank(select(i)) != i at " + std::to_string(i));
        }
    g.hist_total = 0;
    const std::string& t = fs_get(fs, base + "-counts-hist.txt");
    size_t p = 0;
    while (p < t.size()) {
        size_t tab = t.find('\t', p), nl = t.find('\n', p);
        if (tab == std::string::npos || nl == std::string::npos) break;
        g.hist_total += std::stoull(t.substr(tab + 1, nl - tab - 1));
        p = nl + 1;
    }
    return g;
}

struct KmerSetContents { uint64_t k; uint64_t count; std::vector<u128> kmers; };

static inline KmerSetContents read_kmer_set(const MemFS& fs, const std::string& base, bool exercise_select = true) {
    KmerSetContents s;
    const std::string& h = fs_get(fs, base + ".header");
    if (h.size() != 24) throw std::runtime_error("KmerSet header size");
    uint64_t ver; memcpy(&ver, h.data(), 8); memcpy(&s.k, h.data() + 8, 8); memcpy(&s.count, h.data() + 16, 8);
    if (ver != 2011101701ULL) throw std::runtime_error("KmerSet version mismatch");
    SparseArrayReader kmers(fs, base + ".kmers");
    s.kmers = kmers.decode_all();
    if (exercise_select)
        for (uint64_t i = 0; i < s.kmers.size(); ++i)
            if (kmers.select(i) != s.kmers[i] || kmers.rank(s.kmers[i]) != i) throw std::runtime_error("kmer-set select/rank mismatch at " + std::to_string(i));
    return s;
}

// ---------------------------------------------------------------------------------------
// Whole commands
// ---------------------------------------------------------------------------------------

struct BuildStats { uint64_t n_reads, n_instances, n_distinct, n_kept; double t_extract, t_sort, t_emit; };

static inline double now_s() {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

// Extraction split across threads at read granularity (the reference does this part on one
// thread, src/GossCmdBuildGraph.cc:335-380; threads here only make the CPU baseline fairer).
static inline void extract_keys_mt(const std::vector<Input>& inputs, unsigned w, int mode, int threads,
                                   std::vector<u128>& out, uint64_t* n_reads) {
    if (threads <= 1) { extract_keys(inputs, w, mode, out, n_reads); return; }
    // framing is sequential (it validates the whole file); windows are farmed out per read batch
    struct Piece { const char* s; size_t len; std::string own; };
    std::vector<std::vector<Piece>> shards(threads);
    uint64_t reads = 0;
    static const int order[3] = {FMT_LINE, FMT_FASTA, FMT_FASTQ};
    for (int f = 0; f < 3; ++f)
        for (const Input& in : inputs) {
            if (in.format != order[f]) continue;
            const char* lo = in.data; const char* hi = in.data + in.size;
            for_each_read(in.data, in.size, in.format, [&](const char* s, size_t len) {
                Piece p; p.len = len;
                if (s >= lo && s < hi) p.s = s; else { p.own.assign(s ? s : "", len); p.s = nullptr; }
                shards[reads % threads].push_back(std::move(p));
                ++reads;
            });
        }
    std::vector<std::vector<u128>> outs(threads);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            for (const Piece& p : shards[t]) {
                const char* s = p.s ? p.s : p.own.data();
                for_each_window(s, p.len, w, [&](u128 x) {
                    if (mode == MODE_GRAPH) { outs[t].push_back(x); outs[t].push_back(reverse_complement(x, w)); }
                    else if (mode == MODE_KMERSET) outs[t].push_back(normalize(x, w));
                    else outs[t].push_back(x);
                });
            }
        });
    for (auto& th : pool) th.join();
    size_t total = 0; for (auto& o : outs) total += o.size();
    out.reserve(out.size() + total);
    for (auto& o : outs) { out.insert(out.end(), o.begin(), o.end()); std::vector<u128>().swap(o); }
    if (n_reads) *n_reads = reads;
}

static inline void count_keys(std::vector<u128>& raw, unsigned key_bits, uint64_t min_count, int threads,
                              std::vector<u128>& keys, std::vector<uint64_t>& counts) {
    if (key_bits <= 64) {
        std::vector<uint64_t> k64(raw.size());
        for (size_t i = 0; i < raw.size(); ++i) k64[i] = lo64(raw[i]);
        std::vector<u128>().swap(raw);
        sort_keys_parallel(k64, key_bits, threads);
        run_length_reduce(k64, min_count, keys, counts);
    } else {
        sort_keys_parallel(raw, key_bits, threads);
        run_length_reduce(raw, min_count, keys, counts);
    }
}

// build-graph (+ the fused min-count filter that is bit-identical to `trim-graph -C m-1`:
// keep count > C, rebuild with the exact kept count as the size estimate,
// src/GossCmdTrimGraph.cc:97-124).  Single-pass regime: M_est = number of distinct edges
// (src/GossCmdBuildGraph.cc:225-233).
static inline BuildStats build_graph(const std::vector<Input>& inputs, unsigned k, uint64_t min_count, int threads,
                                     const std::string& base, MemFS& fs) {
    if (k > GraphWriter::kMaxK) throw std::runtime_error("unable to build a graph with k=" + std::to_string(k));
    bool any = false; for (const Input& in : inputs) any = any || in.size > 0;
    if (!any) throw std::runtime_error("No valid reads.");   // src/ReverseComplementAdapter.hh:77-86
    BuildStats st{};
    double t0 = now_s();
    std::vector<u128> raw;
    extract_keys_mt(inputs, k + 1, MODE_GRAPH, threads, raw, &st.n_reads);
    st.n_instances = raw.size();
    double t1 = now_s();
    std::vector<u128> keys; std::vector<uint64_t> counts;
    std::vector<u128> all_keys; std::vector<uint64_t> all_counts;
    count_keys(raw, 2 * (k + 1), 1, threads, all_keys, all_counts);
    st.n_distinct = all_keys.size();
    if (min_count > 1) {
        for (size_t i = 0; i < all_keys.size(); ++i)
            if (all_counts[i] >= min_count) { keys.push_back(all_keys[i]); counts.push_back(all_counts[i]); }
    } else { keys.swap(all_keys); counts.swap(all_counts); }
    st.n_kept = keys.size();
    double t2 = now_s();
    GraphWriter g(fs, base, k, keys.size());
    for (size_t i = 0; i < keys.size(); ++i) g.push_back(keys[i], counts[i]);
    g.finish();
    double t3 = now_s();
    st.t_extract = t1 - t0; st.t_sort = t2 - t1; st.t_emit = t3 - t2;
    return st;
}

// build-kmer-set.  Reference: src/GossCmdBuildKmerSet.tcc:167-210,226-250.
static inline BuildStats build_kmer_set(const std::vector<Input>& inputs, unsigned k, int threads,
                                        const std::string& base, MemFS& fs) {
    if (k > KmerSetWriter::kMaxK) throw std::runtime_error("unable to build a graph with k=" + std::to_string(k));
    BuildStats st{};
    double t0 = now_s();
    std::vector<u128> raw;
    extract_keys_mt(inputs, k, MODE_KMERSET, threads, raw, &st.n_reads);
    st.n_instances = raw.size();
    double t1 = now_s();
    std::vector<u128> keys; std::vector<uint64_t> counts;
    count_keys(raw, 2 * k, 1, threads, keys, counts);
    st.n_distinct = st.n_kept = keys.size();
    double t2 = now_s();
    KmerSetWriter s(fs, base, k, keys.size());
    for (u128 x : keys) s.push_back(x);
    s.finish();
    double t3 = now_s();
    st.t_extract = t1 - t0; st.t_sort = t2 - t1; st.t_emit = t3 - t2;
    return st;
}

// ---------------------------------------------------------------------------------------
// xenome index, steps 3 and 4 (src/XenoApp.cc:62-76)
// ---------------------------------------------------------------------------------------

// A plain bit vector written one bit at a time.  Reference: WordyBitVector::Builder::push_backX / end,
// src/WordyBitVector.hh:90-116: ceil(n / 64) little-endian words, one (zero) word when nothing was pushed.
static inline std::string write_bit_vector(const std::vector<bool>& bits) {
    std::vector<uint64_t> words(std::max<size_t>(1, (bits.size() + 63) / 64), 0);
    for (uint64_t i = 0; i < bits.size(); ++i)
        if (bits[i]) words[i >> 6] |= 1ULL << (i & 63);
    return std::string(reinterpret_cast<const char*>(words.data()), words.size() * 8);
}
static inline bool bit_vector_get(const std::string& f, uint64_t i) {
    uint64_t w; memcpy(&w, f.data() + 8 * (i >> 6), 8);
    return (w >> (i & 63)) & 1;
}

struct AnnotateStats { uint64_t n_lhs, n_rhs, n_common, n_out; };

// merge-and-annotate-kmer-sets.  Reference: src/GossCmdMergeAndAnnotateKmerSets.cc:27-207 -- the union of two kmer sets
// (size estimate = exact size of the union, :121) plus one membership bit per element and side.
static inline AnnotateStats merge_and_annotate(MemFS& fs, const std::string& lhs_base, const std::string& rhs_base, const std::string& out_base) {
    KmerSetContents a = read_kmer_set(fs, lhs_base, false), b = read_kmer_set(fs, rhs_base, false);
    if (a.count == 0 || b.count == 0 || a.k != b.k) throw std::runtime_error("nonsense");   // :41-49 (a bare `throw "nonsense"` there)
    std::vector<u128> u; std::vector<bool> lb, rb;
    uint64_t l = 0, r = 0, c = 0;
    while (l < a.kmers.size() || r < b.kmers.size()) {
        const bool take_l = r >= b.kmers.size() || (l < a.kmers.size() && a.kmers[l] <= b.kmers[r]);
        const bool take_r = l >= a.kmers.size() || (r < b.kmers.size() && b.kmers[r] <= a.kmers[l]);
        u.push_back(take_l ? a.kmers[l] : b.kmers[r]);
        lb.push_back(take_l); rb.push_back(take_r);
        if (take_l && take_r) ++c;
        if (take_l) ++l;
        if (take_r) ++r;
    }
    KmerSetWriter w(fs, out_base, a.k, u.size());
    for (u128 x : u) w.push_back(x);
    w.finish();
    fs[out_base + ".lhs-bits"] = write_bit_vector(lb);
    fs[out_base + ".rhs-bits"] = write_bit_vector(rb);
    return AnnotateStats{(uint64_t)a.kmers.size(), (uint64_t)b.kmers.size(), c, (uint64_t)u.size()};
}

// compute-near-kmers.  Reference: src/GossCmdComputeNearKmers.cc:57-118,158-225.  A k-mer that belongs to exactly one
// side turns "gray" (both bits cleared) when one of its variants y = x ^ (b << j), 0 <= j < K, 0 <= b < 4, is a member that
// belongs to exactly one side and to the OTHER side than x.  Decisions use the ORIGINAL bits throughout (:71,:99).
// Two things in the reference are restated as they ARE, not as they may have been meant (the files must match):
//   * the variant mask is shifted by j BITS, not by j bases (:82-83);
//   * `mKmerSet.normalize(y);` (:96) calls GraphEssentials::normalize(const Edge&) const (src/GraphEssentials.hh:152-157),
//     which RETURNS the normalised edge; the result is discarded, so y is looked up as it is.
static inline uint64_t compute_near_kmers(MemFS& fs, const std::string& base) {
    KmerSetContents s = read_kmer_set(fs, base, false);
    const std::string lf = fs_get(fs, base + ".lhs-bits"), rf = fs_get(fs, base + ".rhs-bits");
    const uint64_t n = s.kmers.size();
    std::vector<bool> nl(n), nr(n);
    uint64_t gray = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const bool li = bit_vector_get(lf, i), ri = bit_vector_get(rf, i);
        nl[i] = li; nr[i] = ri;
        if (li == ri) continue;
        const u128 x = s.kmers[i];
        bool found = false;
        for (uint64_t j = 0; !found && j < s.k; ++j) {
            for (uint64_t b = 0; !found && b < 4; ++b) {
                u128 y = x ^ ((u128)b << j);
                if (y == x) continue;
                auto it = std::lower_bound(s.kmers.begin(), s.kmers.end(), y);
                if (it == s.kmers.end() || *it != y) continue;
                const uint64_t r = (uint64_t)(it - s.kmers.begin());
                if (bit_vector_get(lf, r) != bit_vector_get(rf, r) && li != bit_vector_get(lf, r)) found = true;
            }
        }
        if (found) { ++gray; nl[i] = false; nr[i] = false; }
    }
    fs[base + ".lhs-bits"] = write_bit_vector(nl);
    fs[base + ".rhs-bits"] = write_bit_vector(nr);
    return gray;
}

}  // namespace goss_oracle
