// file_io.hh -- host-side I/O seam of the `goss` CLI: the FileFactory analogue
// (reference: src/FileFactory.hh:80-164, src/PhysicalFileFactory.cc:261-280) and the block reader
// that replaces LineSource / BackgroundLineSource (src/LineSource.cc:17-76): instead of handing
// std::string lines to a parser thread it streams raw text blocks, cut at record boundaries,
// from pinned buffers into the CUDA library.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../include/gossamer_b200.h"

namespace goss {

struct Error {
    std::string text;     // already formatted the way src/App.cc:328-407 prints it
};

// Plain, ".gz" (zlib), ".bz2" (libbz2, bound at run time) or "-" (stdin) input, chosen by suffix like PhysicalFileFactory::in
// (src/PhysicalFileFactory.cc:261-280).
class InputFile {
public:
    explicit InputFile(const std::string& name);
    ~InputFile();
    InputFile(const InputFile&) = delete;
    InputFile& operator=(const InputFile&) = delete;
    // reads up to n bytes; returns the number read (0 at end of file)
    size_t read(void* dst, size_t n);
    const std::string& name() const { return name_; }
private:
    size_t read_bz2(void* dst, size_t n);
    std::string name_;
    void* gz_ = nullptr;
    void* bz_ = nullptr;           // .bz2: decoder state (libbz2's bz_stream + the compressed-side buffer)
    int fd_ = -1;
};

// Streams one file as blocks that end at record boundaries.
//   FASTA / line files: cut after the last '\n' of the buffer (a FASTA record may continue).
//   FASTQ: cut before the last line that starts a record ('@' line whose line+2 starts with '+').
class BlockReader {
public:
    BlockReader(const std::string& name, int format, size_t block_bytes);
    ~BlockReader();
    // Next block: pointer into pinned memory, size, and whether it is the last of the file.
    bool next(const uint8_t*& data, size_t& size, bool& last);
private:
    size_t cut_point(size_t filled, bool eof) const;
    InputFile in_;
    int format_;
    size_t cap_;
    uint8_t* buf_[2] = {nullptr, nullptr};
    int cur_ = 0;
    size_t carry_ = 0;       // bytes carried from the previous block (held in carry_store_ until the next call)
    std::vector<uint8_t> carry_store_;
    bool eof_ = false, done_ = false;
};

// Output files under a prefix, written with pwrite (PhysicalFileFactory::out).
class OutputFiles {
public:
    // shared = true: several worker processes write their own pieces of the same files (multi-GPU emission): the files
    // are created without truncation and sized with ftruncate; the launcher removes stale files beforehand
    explicit OutputFiles(bool shared = false);
    gsb_sink* sink() { return &sink_; }
    uint64_t bytes_written() const { return bytes_; }
    std::vector<std::string> names() const { return names_; }
private:
    static int s_open(void* user, const char* name, uint64_t size_hint, void** handle);
    static int s_pwrite(void* user, void* handle, uint64_t offset, const void* data, uint64_t len);
    static int s_close(void* user, void* handle);
    gsb_sink sink_;
    uint64_t bytes_ = 0;
    std::vector<std::string> names_;
    bool shared_ = false;
};

// Existing files behind the gsb_source callbacks (PhysicalFileFactory::in for the readers of a file set).
class InputFiles {
public:
    InputFiles();
    gsb_source* source() { return &src_; }
private:
    static int s_size(void* user, const char* name, uint64_t* size_out);
    static int s_pread(void* user, const char* name, uint64_t offset, void* dst, uint64_t len);
    gsb_source src_;
};

// removes every file `<prefix>.*` / `<prefix>-*` in the prefix's directory (Graph::remove, src/Graph.cc:380-387, widened to
// whatever an earlier run left under the prefix)
void remove_file_set(const std::string& prefix);

// "-O prefix" check: create and remove <prefix>.test (src/GossOptionChecker.hh:80-105)
void check_output_prefix(const std::string& prefix);
// readable-file check for inputs
void check_readable(const std::string& name);
// -F / -f: a file of file names, one per line (src/GossOptionChecker.hh:405-425)
std::vector<std::string> expand_file_list(const std::string& list_name);

}  // namespace goss
