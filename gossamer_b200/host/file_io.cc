// file_io.cc -- see file_io.hh
#include "file_io.hh"

#include <dirent.h>
#include <dlfcn.h>
#include <sys/stat.h>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cerrno>
#include <cstring>
#include <fstream>
#include <vector>

namespace goss {

static bool ends_with(const std::string& s, const char* suf) {
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

// ---- bzip2 input (src/PhysicalFileFactory.cc:272-275: the reference pushes Boost.Iostreams' bzip2_decompressor, i.e.
// libbz2, in front of the file).  The image ships libbz2.so.1 without its header, so the library's (stable, documented)
// low-level C interface is declared here and bound at run time.
namespace {
struct BzStream {                                   // bz_stream of bzlib.h
    char* next_in; unsigned avail_in; unsigned total_in_lo32, total_in_hi32;
    char* next_out; unsigned avail_out; unsigned total_out_lo32, total_out_hi32;
    void* state;
    void* (*bzalloc)(void*, int, int); void (*bzfree)(void*, void*); void* opaque;
};
enum { kBzOk = 0, kBzStreamEnd = 4 };
struct BzApi {
    int (*init)(BzStream*, int, int) = nullptr;
    int (*decompress)(BzStream*) = nullptr;
    int (*end)(BzStream*) = nullptr;
    bool ok = false;
};
BzApi& bz_api() {
    static BzApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libbz2.so.1.0", RTLD_NOW);
        if (!h) h = dlopen("libbz2.so.1", RTLD_NOW);
        if (h) {
            api.init = (int (*)(BzStream*, int, int))dlsym(h, "BZ2_bzDecompressInit");
            api.decompress = (int (*)(BzStream*))dlsym(h, "BZ2_bzDecompress");
            api.end = (int (*)(BzStream*))dlsym(h, "BZ2_bzDecompressEnd");
            api.ok = api.init && api.decompress && api.end;
        }
    }
    return api;
}
struct BzState {
    BzStream strm;
    std::vector<char> in;
    bool stream_open = false, eof = false;
};
}  // namespace

InputFile::InputFile(const std::string& name) : name_(name) {
    if (ends_with(name, ".bz2")) {
        if (!bz_api().ok) throw Error{"\t'" + name + "': bzip2 input needs libbz2.so.1, which could not be loaded\n"};
        fd_ = open(name.c_str(), O_RDONLY);
        if (fd_ < 0) throw Error{"\t'" + name + "': " + strerror(errno) + "\n"};
        BzState* st = new BzState();
        memset(&st->strm, 0, sizeof(st->strm));
        st->in.resize(1 << 20);
        bz_ = st;
        return;
    }
    if (name == "-") { fd_ = 0; return; }
    if (ends_with(name, ".gz")) {
        gz_ = gzopen(name.c_str(), "rb");
        if (!gz_) throw Error{"\t'" + name + "': " + strerror(errno) + "\n"};
        gzbuffer((gzFile)gz_, 1 << 20);
        return;
    }
    fd_ = open(name.c_str(), O_RDONLY);
    if (fd_ < 0) throw Error{"\t'" + name + "': " + strerror(errno) + "\n"};
#ifdef POSIX_FADV_SEQUENTIAL
    posix_fadvise(fd_, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
}

InputFile::~InputFile() {
    if (bz_) {
        BzState* st = (BzState*)bz_;
        if (st->stream_open) bz_api().end(&st->strm);
        delete st;
    }
    if (gz_) gzclose((gzFile)gz_);
    if (fd_ > 0) close(fd_);
}

// concatenated bzip2 streams are decoded one after the other, as bzip2(1) and Boost's multi-stream filter do
size_t InputFile::read_bz2(void* dst, size_t n) {
    BzState* st = (BzState*)bz_;
    BzApi& api = bz_api();
    size_t got = 0;
    while (got < n && !st->eof) {
        if (st->strm.avail_in == 0) {
            long r;
            do { r = ::read(fd_, st->in.data(), st->in.size()); } while (r < 0 && errno == EINTR);
            if (r < 0) throw Error{"\t'" + name_ + "': read error\n"};
            if (r == 0) {
                if (st->stream_open) throw Error{"\t'" + name_ + "': unexpected end of bzip2 data\n"};
                st->eof = true;
                break;
            }
            st->strm.next_in = st->in.data();
            st->strm.avail_in = (unsigned)r;
        }
        if (!st->stream_open) {
            char* keep_in = st->strm.next_in; const unsigned keep_avail = st->strm.avail_in;
            memset(&st->strm, 0, sizeof(st->strm));
            st->strm.next_in = keep_in; st->strm.avail_in = keep_avail;
            if (api.init(&st->strm, 0, 0) != kBzOk) throw Error{"\t'" + name_ + "': cannot start the bzip2 decoder\n"};
            st->stream_open = true;
        }
        st->strm.next_out = (char*)dst + got;
        st->strm.avail_out = (unsigned)std::min<size_t>(n - got, 1u << 30);
        const unsigned before = st->strm.avail_out;
        const int rc = api.decompress(&st->strm);
        got += before - st->strm.avail_out;
        if (rc == kBzStreamEnd) { api.end(&st->strm); st->stream_open = false; continue; }
        if (rc != kBzOk) throw Error{"\t'" + name_ + "': corrupt bzip2 data\n"};
    }
    return got;
}

size_t InputFile::read(void* dst, size_t n) {
    if (bz_) return read_bz2(dst, n);
    size_t got = 0;
    while (got < n) {
        long r;
        if (gz_) r = gzread((gzFile)gz_, (char*)dst + got, (unsigned)std::min<size_t>(n - got, 1u << 30));
        else r = ::read(fd_, (char*)dst + got, n - got);
        if (r < 0) {
            if (!gz_ && errno == EINTR) continue;
            throw Error{"\t'" + name_ + "': read error\n"};
        }
        if (r == 0) break;
        got += (size_t)r;
    }
    return got;
}

BlockReader::BlockReader(const std::string& name, int format, size_t block_bytes) : in_(name), format_(format), cap_(block_bytes) {
    for (int i = 0; i < 2; ++i) {
        void* p = nullptr;
        if (gsb_host_alloc(cap_, &p) != GSB_OK) throw Error{std::string("\tcannot allocate pinned staging buffer: ") + gsb_last_error(nullptr) + "\n"};
        buf_[i] = (uint8_t*)p;
    }
}

BlockReader::~BlockReader() {
    for (int i = 0; i < 2; ++i) gsb_host_free(buf_[i]);
}

// Index one past the last byte of the block to hand over (the rest is carried to the next block).
size_t BlockReader::cut_point(size_t filled, bool eof) const {
    const uint8_t* b = buf_[cur_];
    if (eof) return filled;
    // last newline
    size_t nl = filled;
    while (nl > 0 && b[nl - 1] != '\n') --nl;
    if (nl == 0) return 0;                                   // no complete line in the buffer
    if (format_ != GSB_FMT_FASTQ) return nl;
    // FASTQ: walk back over line starts until one looks like a record header
    size_t line_end = nl;                                    // exclusive end (just after a '\n')
    for (int guard = 0; guard < 100000 && line_end > 0; ++guard) {
        size_t ls = line_end - 1;                            // start of the line ending at line_end
        while (ls > 0 && b[ls - 1] != '\n') --ls;
        if (b[ls] == '@') {
            // find the start of line+2
            size_t p = ls; int seen = 0;
            while (p < nl && seen < 2) { if (b[p] == '\n') ++seen; ++p; }
            if (seen == 2 && p < nl && b[p] == '+') return ls; // records before this header are complete
        }
        line_end = ls;
    }
    // Not the four-line layout (wrapped sequence / quality lines, or very short lines): find the record boundaries the way
    // the reference's parser does (src/FastqParser.hh:78-176), forwards from the start of the buffer -- which IS a record
    // boundary.  Sequence lines run up to the first line that starts with '@' or '+'; quality lines run up to the first
    // such line seen once the quality is at least as long as the sequence; that line is the next record's header.  The last
    // record in the buffer may continue in the next block, so the cut is the start of the last header found.
    size_t pos = 0, last_header = 0;
    auto line_end_of = [&](size_t from) { size_t e = from; while (e < nl && b[e] != '\n') ++e; return e; };   // nl is just after a '\n'
    while (pos < nl) {
        if (b[pos] != '@') return nl;                        // malformed: hand everything over, the device parser reports it precisely
        last_header = pos;
        pos = line_end_of(pos) + 1;
        size_t seq = 0, qual = 0;
        bool plus = false;
        while (pos < nl) {                                   // sequence lines
            const size_t e = line_end_of(pos);
            if (e > pos && (b[pos] == '@' || b[pos] == '+')) { plus = b[pos] == '+'; if (plus) pos = e + 1; break; }
            seq += e - pos;
            pos = e + 1;
        }
        if (!plus) { if (pos < nl) return nl; break; }       // '@' where '+' was due: malformed; or the buffer ended inside the sequence
        bool next_record = false;
        while (pos < nl) {                                   // quality lines
            const size_t e = line_end_of(pos);
            if (e > pos && (b[pos] == '@' || b[pos] == '+') && qual >= seq) { next_record = true; break; }
            qual += e - pos;
            pos = e + 1;
        }
        if (!next_record) break;                             // the buffer ended inside (or right behind) this record
    }
    return last_header;
}

bool BlockReader::next(const uint8_t*& data, size_t& size, bool& last) {
    if (done_) return false;
    for (;;) {
        uint8_t* b = buf_[cur_];
        // The tail carried over from the previous block is placed only now: that block may still be on its way to the
        // device (GSB_BLOCK_ASYNC), and this buffer -- the one of the block before it -- is free again once the push
        // of the previous block has returned.
        if (carry_) memcpy(b, carry_store_.data(), carry_);
        size_t filled = carry_;
        if (!eof_) {
            size_t got = in_.read(b + filled, cap_ - filled);
            filled += got;
            if (filled < cap_) eof_ = true;
        }
        size_t cut = cut_point(filled, eof_);
        if (cut == 0 && !eof_)
            throw Error{"\t'" + in_.name() + "': a single record is larger than the " + std::to_string(cap_ >> 20) + " MiB input block; raise --block-mb\n"};
        // keep the tail (an incomplete line / record) for the next block
        carry_ = filled - cut;
        carry_store_.assign(b + cut, b + cut + carry_);
        data = b; size = cut; last = eof_ && carry_ == 0;
        cur_ ^= 1;
        if (last) done_ = true;
        return true;
    }
}

OutputFiles::OutputFiles(bool shared) : shared_(shared) {
    sink_.user = this;
    sink_.open = &OutputFiles::s_open;
    sink_.pwrite = &OutputFiles::s_pwrite;
    sink_.close = &OutputFiles::s_close;
}

int OutputFiles::s_open(void* user, const char* name, uint64_t size_hint, void** handle) {
    OutputFiles* self = (OutputFiles*)user;
    int fd = ::open(name, O_WRONLY | O_CREAT | (self->shared_ ? 0 : O_TRUNC), 0644);
    if (fd < 0) return -1;
    if (size_hint || self->shared_) { if (ftruncate(fd, (off_t)size_hint) != 0) { /* best effort */ } }
    self->names_.push_back(name);
    *handle = (void*)(intptr_t)(fd + 1);
    return 0;
}

int OutputFiles::s_pwrite(void* user, void* handle, uint64_t offset, const void* data, uint64_t len) {
    OutputFiles* self = (OutputFiles*)user;
    int fd = (int)(intptr_t)handle - 1;
    uint64_t done = 0;
    while (done < len) {
        ssize_t w = ::pwrite(fd, (const char*)data + done, len - done, (off_t)(offset + done));
        if (w < 0) { if (errno == EINTR) continue; return -1; }
        done += (uint64_t)w;
    }
    self->bytes_ += len;
    return 0;
}

int OutputFiles::s_close(void*, void* handle) {
    int fd = (int)(intptr_t)handle - 1;
    return ::close(fd) == 0 ? 0 : -1;
}

InputFiles::InputFiles() {
    src_.user = this;
    src_.size = &InputFiles::s_size;
    src_.pread = &InputFiles::s_pread;
}

int InputFiles::s_size(void*, const char* name, uint64_t* size_out) {
    struct stat st;
    if (::stat(name, &st) != 0 || !S_ISREG(st.st_mode)) return -1;
    *size_out = (uint64_t)st.st_size;
    return 0;
}

int InputFiles::s_pread(void*, const char* name, uint64_t offset, void* dst, uint64_t len) {
    int fd = ::open(name, O_RDONLY);
    if (fd < 0) return -1;
    uint64_t done = 0;
    while (done < len) {
        ssize_t r = ::pread(fd, (char*)dst + done, len - done, (off_t)(offset + done));
        if (r < 0) { if (errno == EINTR) continue; ::close(fd); return -1; }
        if (r == 0) { ::close(fd); return -1; }
        done += (uint64_t)r;
    }
    ::close(fd);
    return 0;
}

void remove_file_set(const std::string& prefix) {
    const size_t slash = prefix.rfind('/');
    const std::string dir = slash == std::string::npos ? "." : prefix.substr(0, slash == 0 ? 1 : slash);
    const std::string stem = slash == std::string::npos ? prefix : prefix.substr(slash + 1);
    DIR* d = ::opendir(dir.c_str());
    if (!d) return;
    std::vector<std::string> victims;
    while (struct dirent* e = ::readdir(d)) {
        const std::string n = e->d_name;
        if (n.size() > stem.size() && n.compare(0, stem.size(), stem) == 0 && (n[stem.size()] == '.' || n[stem.size()] == '-')) victims.push_back(dir + "/" + n);
    }
    ::closedir(d);
    for (const std::string& v : victims) ::unlink(v.c_str());
}

void check_output_prefix(const std::string& prefix) {
    const std::string probe = prefix + ".test";
    int fd = ::open(probe.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) throw Error{"cannot create filenames with prefix '" + prefix + "'\n"};
    ::close(fd);
    ::unlink(probe.c_str());
}

void check_readable(const std::string& name) {
    if (name == "-") return;
    if (access(name.c_str(), R_OK) != 0) throw Error{"\t'" + name + "': " + strerror(errno) + "\n"};
}

std::vector<std::string> expand_file_list(const std::string& list_name) {
    std::ifstream in(list_name);
    if (!in) throw Error{"\t'" + list_name + "': " + strerror(errno) + "\n"};
    std::vector<std::string> out;
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty()) out.push_back(line);
    }
    return out;
}

}  // namespace goss
