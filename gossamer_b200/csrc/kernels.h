// kernels.h -- internal interface between the CUDA modules (ingest / sort / emit) and the
// context that owns memory and orchestrates them.  Not part of the public ABI.
#pragma once
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/gossamer_b200.h"
#include "common.cuh"

namespace gsb {

// parse / capacity errors raised on the device, rendered by the host with the reference's text
enum {
    GSB_PE_NONE = 0,
    GSB_PE_FASTA_EXPECT_GT = 1,        // "expected '>' at beginning of line N"            src/FastaParser.hh:60-67
    GSB_PE_FASTQ_EXPECT_AT = 2,        // "expected '@' at beginning of line N"            src/FastqParser.hh:89-97
    GSB_PE_FASTQ_EXPECT_SEQ = 3,       // "expected sequence data or quality header at line N"   :107-113
    GSB_PE_FASTQ_EXPECT_PLUS = 4,      // "expected '+' at beginning of line N"            :122-129
    GSB_PE_FASTQ_TITLE_MISMATCH = 5,   // "quality title does not match sequence title at line N" :132-140
    GSB_PE_FASTQ_LEN_MISMATCH = 6,     // "length mistmatch between sequence and quality data just before line N" :167-174
    GSB_PE_KEY_OVERFLOW = 100
};

struct IngestStatus {
    int error;
    int fastq_irregular;
    u64 error_line;
    u64 n_reads;
    u64 n_self_rc;             // graph mode: windows that are their own reverse complement (rare; see fold.cu)
};

// Device memory for one stream: a caching allocator.  Freed blocks are kept in size-ordered free
// lists and handed out again to later requests of a similar size, so a steady-state step performs
// no driver allocation at all.  (The first version used cudaMallocAsync; with the multi-GB buffers
// of this path the pool kept re-mapping memory and emission time varied between 6 and 44 ms.)
// Reuse is safe without events because every user of a Workspace enqueues on its single stream.
struct Workspace {
    cudaStream_t stream = nullptr;
    int device = 0;
    int sm_count = 148;
    u64 launches = 0;
    u64 live_bytes = 0, peak_bytes = 0, reserved_bytes = 0;
    u64 device_allocs = 0;       // cudaMalloc calls (the caching allocator's misses + the exchange windows)
    void* alloc(size_t bytes);
    void release(void* p, size_t bytes);
    void sync();
    void trim();                 // give every cached block back to the driver
    ~Workspace();
    Workspace() {}
    Workspace(const Workspace&) = delete;
    Workspace& operator=(const Workspace&) = delete;
private:
    std::multimap<size_t, void*> free_;
    std::unordered_map<void*, size_t> size_of_;
};

template <typename T>
struct DevBuf {
    Workspace* ws = nullptr;
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(Workspace* w, size_t count) { reset(w, count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : ws(o.ws), p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { free(); ws = o.ws; p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DevBuf() { free(); }
    void reset(Workspace* w, size_t count) { free(); ws = w; n = count; p = (T*)w->alloc((count ? count : 1) * sizeof(T)); }
    void free() { if (p) { ws->release(p, (n ? n : 1) * sizeof(T)); p = nullptr; n = 0; } }
    size_t bytes() const { return n * sizeof(T); }
};

// ---- ingest.cu -------------------------------------------------------------------------------
u32 ingest_newline_tiles(u64 nbytes);
void ingest_count_newlines(const u8* text, u64 n, u32* tile_counts, cudaStream_t s, u64* launches);
void ingest_scan_tiles(u32* tile_counts, u32 tiles, u32* total, cudaStream_t s, u64* launches);
void ingest_fill_line_starts(const u8* text, u64 n, const u32* tile_offsets, u32* line_start, u32 n_lines, cudaStream_t s, u64* launches);
void ingest_classify(const u8* text, const u32* line_start, u32 n_lines, int format, int file_start, u64 line_base,
                     u8* kind, u32* nsym, IngestStatus* st_dev, cudaStream_t s, u64* launches, u32* fq_scratch = nullptr);
u64 ingest_fastq_scratch_words(u32 n_lines);     // scratch the irregular-FASTQ framing wants (u32 words); without it: one thread
void ingest_symbol_offsets(const u32* nsym, u32* sym_off, u32 n_lines, u32* total_dev, u32* tmp, cudaStream_t s, u64* launches);
void ingest_pack(const u8* text, u64 text_bytes, const u32* line_start, const u8* kind, const u32* sym_off, u32 n_lines,
                 const u8* carry, u32 n_carry, u64 n_sym_total, u64* codes, u32* valid, u64 n_words, u32* word_line /* scratch [n_words] */,
                 cudaStream_t s, u64* launches);
void ingest_save_carry(const u64* codes, const u32* valid, u64 n_sym_total, u32 want, u8* carry_out, cudaStream_t s, u64* launches);
// mix != 0 (graph mode only): the folded key is stored bit-mixed (key_mix, common.cuh)
void ingest_extract(int kind, int key_bytes, const u64* codes, const u32* valid, u64 p_begin, u64 p_end, int w, int passes, int mix,
                    void* out, u64* cursor, u64 capacity, u64* digit_hist, IngestStatus* st, int sm_count, cudaStream_t s, u64* launches);

// ---- sort.cu ---------------------------------------------------------------------------------
u64 sort_tile_keys(int key_bytes, bool with_values = false);
void sort_set_tuning(int id);
void sort_fill_random(int key_bytes, void* keys, u64 n, int key_bits, u64 seed, cudaStream_t s);
u64 sort_lookback_bytes(int key_bytes, u64 n, bool with_values = false);
void sort_digit_hist(int key_bytes, const void* keys, u64 n, int passes, u64* hist, int sm_count, cudaStream_t s, u64* launches);
void sort_digit_base(const u64* hist, u64* base, int passes, cudaStream_t s, u64* launches);
void sort_pass(int key_bytes, const void* in, void* out, const u64* vin, u64* vout, u64 n, int pass, const u64* digit_base_all,
               void* lookback, cudaStream_t s, u64* launches, cudaEvent_t ev_begin = nullptr, cudaEvent_t ev_end = nullptr);
u64 rle_lookback_bytes(u64 n);
u64 rle_tiles(u64 n);
void sort_rle_count(int key_bytes, const void* keys, u64 n, u64 min_count, int fold_w, u32* tile_kept, u64* total_heads /* u64[4] */, cudaStream_t s, u64* launches);
void sort_rle_emit(int key_bytes, const void* keys, const u64* csum, u64 n, u64 min_count, int fold_w, const u64* tile_off, void* out_keys, u64* out_counts,
                   cudaStream_t s, u64* launches);
void sort_filter(int key_bytes, const void* keys, const u64* counts, const u64* pos, u64 m, u64 min_count, void* out_keys, u64* out_counts,
                 void* lookback, u64* total_dev, cudaStream_t s, u64* launches);
void sort_scan_weights(const u64* w, u64* csum, u64 n, u64* tmp, cudaStream_t s, u64* launches);
u64 sort_scan_tmp_elems(u64 n);

// Full sort of n keys (optionally with a u64 payload).  `a` holds the input; `b` is scratch of the
// same size.  Digit histograms are computed here unless hist_dev (already accumulated, [passes][256])
// is given.  Returns 0 if the result is in a, 1 if in b.  passes_run gets the number of sweeps.
// Only the digits [digit_begin, digit_end) are swept (default: all): a stable LSD sort on those bits.
int sort_keys(Workspace& ws, int key_bytes, int key_bits, void* a, void* b, u64* va, u64* vb, u64 n,
              const u64* hist_dev, int* passes_run, double* sweep_ms = nullptr, int digit_begin = 0, int digit_end = -1);

// sorted keys (+ optional weights) -> distinct keys and summed counts, min-count filtered.
// Outputs are freshly allocated; *m_distinct is the count before the filter.
struct ReducedRun {
    DevBuf<u8> keys;        // key_bytes * m
    DevBuf<u64> counts;     // m
    u64 m = 0;
};
// fold_w > 0 (only without weights): the keys are strand-folded windows of fold_w symbols -- the counts of
// self-complementary keys are doubled and the filter applies to the doubled count; *n_self_rc gets the
// number of distinct self-complementary keys (before the filter).
void reduce_sorted(Workspace& ws, int key_bytes, const void* sorted, const u64* weights, u64 n, u64 min_count,
                   ReducedRun& out, u64* m_distinct, int fold_w = 0, u64* n_self_rc = nullptr);

// keys[i] = key_unmix(keys[i])
void sort_unmix_inplace(int key_bytes, void* keys, u64 n, int sm_count, cudaStream_t s, u64* launches);
// number of elements the descriptors {first index, length} cover -> *total_dev (sort.cu)
void sort_desc_total(const ulonglong2* desc, u64 n_desc, u64* total_dev, cudaStream_t s, u64* launches);

// ---- partition.cu ------------------------------------------------------------------------------
static const int kMaxRanks = 32;
static const int kTopHistBits = 10;        // the extraction kernel histograms the top 10 bits of the low (mixed) key word
struct PartitionPlan { int total_bits = 0, levels = 0; int bits[8] = {0, 0, 0, 0, 0, 0, 0, 0}; u32 max_slots = 0; };
PartitionPlan partition_plan(int key_bytes, u64 n_all_ranks);
u32 partition_tile_keys(int key_bytes);
void partition_set_debug(u32 max_slots, int total_bits);          // test-only geometry overrides (0 = default)
void pairsort_set_debug(u32 cap, int bits);                       // test-only: bucket capacity (0 = default), partition bits (< 0 = default)
struct PartitionTiming { double ms_partition = 0, ms_count = 0, ms_scatter = 0; int levels = 0, total_bits = 0; u64 scatter_launches = 0, n_overflow_keys = 0; };
struct PartitionInput {
    void* keys = nullptr; void* scratch = nullptr;
    u64 n = 0;                       // one parent of n keys (host-known) ...
    const u64* hist_top = nullptr;   // ... optionally with the [2^kTopHistBits] histogram of its top bits
    DevBuf<u64> cstart;              // ... or parents made by earlier passes: [n_parents + 1] starts (device), taken over
    u64 n_parents = 1, n_cap = 0;    //     n_cap: upper bound of the key count (allocation sizes)
    int consumed_bits = 0;
    // ... or a first pass run block by block (partition_block_level0): keys = the buffer the blocks were split INTO, n keys in all
    const u64* runs = nullptr;       //     [runs_n_src][2^runs_bits0 + 1] absolute starts of every block's children
    int runs_n_src = 0, runs_bits0 = 0;
    const u64* runs_hist_next = nullptr;   // [2^(runs_bits0 + kTopHistBits)] histogram of the next bits of every child, summed over the blocks
};
int partition_stream_bits0();
PartitionPlan partition_plan_streamed(int key_bytes, u64 n, int bits0);
void partition_block_level0(Workspace& ws, int key_bytes, void* keys_block, void* alt_block, u64 block_off, u64 n_block, int bits0, const u64* hist_block_top,
                            u64* runs_out, u64* hist_next, cudaEvent_t* e0_out = nullptr, cudaEvent_t* e1_out = nullptr);
// Counting by partitioning (partition.cu); see there for the contract.
bool count_partitioned(Workspace& ws, int key_bytes, int key_bits, PartitionInput& in, const PartitionPlan& plan, u64 min_count, int fold_w,
                       ReducedRun& out, u64* m_distinct, u64* n_self_rc, int* where_keys = nullptr, PartitionTiming* timing = nullptr);
// first pass of a multi-GPU build: children stored straight into their owners' windows (exchange.cu)
void partition_scatter_to_peers(Workspace& ws, int key_bytes, const void* in, u64 n, int bits, u64* cursor, u32 cstride,
                                void* const* peer_base, int n_peers, const u32* abort_flag);
void partition_fold_hist(Workspace& ws, const u64* hist_top, int bits, u64* out);
// pieces of the multi-GPU PULL exchange (partition.cu; orchestrated by exchange.cu)
void partition_local_level0(Workspace& ws, int key_bytes, void* in, void* out, u64 n, int bits, const u64* hist_top, DevBuf<u64>& cstart_out,
                            cudaEvent_t* e0_out = nullptr, cudaEvent_t* e1_out = nullptr, u64 in_off = 0, u64 out_off = 0);
void partition_next_hist(Workspace& ws, int key_bytes, const void* keys, const u64* cstart, int bits0, u64 n, int bits, u64* hist, bool accumulate = false);
void partition_pull_check(Workspace& ws, u64* hist_all, const u64* gathered, int bits0, int bits1, int n_ranks, int rank, u64 cap_keys,
                          u32* abort_flag, u64* n_recv, u64* n_remote);
void partition_pull_level(Workspace& ws, int key_bytes, const void* const* src_base, int n_src, const u64* gathered, int bits0, u32 c_lo, u32 n_parents,
                          u64 n_cap, int bits, const u64* hist_slice, void* out, const u32* abort_flag, DevBuf<u64>& cstart_out, cudaEvent_t e0, cudaEvent_t e1,
                          bool same_base = false);

// (key, count) pairs with distinct keys, arbitrary order -> ordered by key (most-significant-digit passes + a shared-memory
// sort per bucket, partition.cu); fold_w > 0: the reverse complements join the set first.  false = declined, radix-sort instead.
bool sort_pairs_msd(Workspace& ws, int key_bytes, int key_bits, const void* keys, const u64* counts, u64 m, int fold_w, ReducedRun& out);
// its pieces (the multi-GPU re-partition of the survivors runs the first pass across NVLink, exchange.cu)
struct PairSortPlan { int bits = 0, levels = 0; int lb[8] = {0, 0, 0, 0, 0, 0, 0, 0}; u32 cap = 0; };
u32 pairsort_elem_bytes(int key_bytes);
PairSortPlan pairsort_plan(int key_bytes, int key_bits, u64 n_cap, int first_bits);
void pairsort_pack(Workspace& ws, int key_bytes, int key_bits, const void* keys, const u64* counts, u64 m, int fold_w, void* elems, u64* n_dev, u64* hist0, int bits0);
bool pairsort_finish(Workspace& ws, int key_bytes, int key_bits, void* cur, void* other, u64 n_cap, const u64* n_dev, DevBuf<u64>& cstart, u64 n_parents,
                     int consumed, const PairSortPlan& plan, int first_level, const u64* hist_first, ReducedRun& out);
void pairsort_plan_owners(Workspace& ws, const u64* hist_all, int n_ranks, int rank, int bits0, u64* cursor, u32 cstride, u8* owner, u64* totals, u32* range,
                          u64* cstart_local);
void pairsort_scatter_to_peers(Workspace& ws, int key_bytes, int key_bits, const void* elems, u64 n_cap, const u64* n_dev, int bits0, u64* cursor, u32 cstride,
                               const u8* owner, void* const* peer_base, int n_peers);

// ---- fold.cu ---------------------------------------------------------------------------------
// Strand folding (graph mode): instances are counted as min(x, rc x); these restore both strands.
// merged, still folded run with raw occurrence counts -> doubled counts for self-complementary keys;
// returns how many of the m keys are self-complementary
u64 fold_double_self_rc(Workspace& ws, int key_bytes, int w, const void* keys, u64* counts, u64 m);
// appends (rc y, count) of every key y of the run that is not self-complementary to out_* (room for m pairs,
// arbitrary order); returns how many were written
u64 unfold_append_rc(Workspace& ws, int key_bytes, int w, const void* keys, const u64* counts, u64 m, void* out_keys, u64* out_counts);
// folded, filtered run (final counts) -> the full sorted run: every key y plus rc(y) with the same count
// sorted_input = false: the run is in arbitrary order (reduce_groups): run ++ rc(run) is sorted as a whole instead
// of sorting rc(run) and merging
void unfold_run(Workspace& ws, int key_bytes, int key_bits, int w, ReducedRun& run, bool sorted_input = true);
// two sorted runs with disjoint key sets -> one sorted run (merge path)
void merge_disjoint_runs(Workspace& ws, int key_bytes, const void* ka, const u64* ca, u64 na, const void* kb, const u64* cb, u64 nb,
                         void* out_keys, u64* out_counts);

// ---- emit.cu ---------------------------------------------------------------------------------
// Device -> sink pipeline of the emitter: a ring of staging slots (device + pinned host), a copy stream and a WRITER THREAD
// -- the analogue of the reference's writer threads behind Graph::Builder (src/Graph.cc:148-150).  The emitting thread only
// copies a finished file chunk device-to-device into a free slot on the library's stream (the producer's buffer is free
// again at once) and queues a job; the writer thread moves the chunk across PCIe on the copy stream (copies are queued ahead
// of the callbacks), and makes EVERY sink call (open, pwrite, close), in queue order, so a sink never sees two threads.
struct EmitRing {
    static const int kSlots = 8;
    static const size_t kChunk = 16ull << 20;
    // One unit of work for the writer thread, delivered strictly in the order it was queued.
    struct Job {
        std::string name;
        u64 size_hint = 0;             // open(name, size_hint) ...
        bool open = false, close = false;   // ... before / close after this job's write
        u64 file_off = 0, len = 0;
        int slot = -1;                 // >= 0: the bytes come from this staging slot (device -> pinned host); -1: from `bytes`
        bool issued = false;           // the D2H copy has been queued on the copy stream
        std::vector<u8> bytes;         // host data (slot < 0), or the host prefix that overrides the start of the chunk
    };
    cudaStream_t copy = nullptr;
    int device = 0;
    u8* dev[kSlots] = {};
    u8* host[kSlots] = {};
    cudaEvent_t ready[kSlots], done[kSlots];
    bool slot_busy[kSlots] = {};
    bool created = false;
    // writer thread state (guarded by mu)
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv_work, cv_idle;
    std::deque<Job> queue;
    std::map<std::string, void*> handles;      // files currently open at the sink (writer thread only)
    const gsb_sink* sink = nullptr;
    bool stop = false;
    std::string error;                         // first failure on the writer thread; later jobs are dropped
    void create(Workspace& ws);
    void destroy(Workspace& ws);
    void drop();                   // waits until the writer is idle, forgets errors and open handles
    int acquire_slot();            // blocks until a staging slot is free
    void push(Job&& j);
    void wait_idle();              // every queued job has reached the sink (or was dropped after an error)
private:
    void run();
};

struct Emitter {
    Workspace* ws;
    const gsb_sink* sink;          // nullptr: build on the device, count the bytes, drop
    u64 bytes_out = 0;
    u8* pinned = nullptr;          // staging for device -> sink copies when there is no ring
    size_t pinned_bytes = 0;
    EmitRing* ring = nullptr;      // optional: overlapped delivery by a writer thread (gsb_emit)
    // whole file from host memory
    void put_host(const std::string& name, const void* data, u64 len);
    // whole file = optional host prefix that overrides the first prefix_len bytes + device payload
    void put_device(const std::string& name, const void* dev, u64 len, const void* host_prefix = nullptr, u64 prefix_len = 0);
    // one piece of a file of `total` bytes (multi-GPU emission: every rank hands over its own pieces)
    void put_device_at(const std::string& name, u64 total, u64 offset, const void* dev, u64 len);
    void put_host_at(const std::string& name, u64 total, u64 offset, const void* data, u64 len);
    // waits until everything has reached the sink; throws what the writer thread ran into.  Must be called before the
    // sink's owner looks at the files
    void flush();
private:
    void ring_put(const std::string& name, u64 size_hint, u64 file_off, const void* dev, u64 len, const void* host_prefix, u64 prefix_len);
    void ring_put_host(const std::string& name, u64 size_hint, u64 file_off, const void* data, u64 len);
};

struct U128 { u64 lo, hi; };

// SparseArray::Builder::d (src/SparseArray.cc:47-72), host side, double arithmetic
u64 sparse_array_d(U128 universe, u64 m_est);

// Elias-Fano set of `m` sorted distinct positions: writes base.header, base.high-bits, base-d0,
// base-d1, base.low-bits[...] (src/SparseArray.{hh,cc}).
void emit_sparse_array(Emitter& em, int key_bytes, const void* keys, u64 m, U128 universe_ctor, u64 m_est, U128 universe_end,
                       const std::string& base);
// VariableByteArray (src/VariableByteArray.{hh,cc}) from 64-bit counts (truncated to u32 as the reference does)
void emit_counts(Emitter& em, const u64* counts, u64 m, u64 m_est, const std::string& base);
// "count\tfrequency\n" lines in ascending count order (src/Graph.cc:127-133)
void emit_count_histogram(Emitter& em, const u64* counts, u64 m, const std::string& name);

// ---- reader.cu -------------------------------------------------------------------------------
void read_sparse_array(Workspace& ws, const gsb_source* src, const std::string& base, int key_bytes, u8* pinned, size_t pinned_bytes,
                       DevBuf<u8>& keys_out, u64* m_out);
void read_counts(Workspace& ws, const gsb_source* src, const std::string& base, u64 m, u8* pinned, size_t pinned_bytes, DevBuf<u64>& counts_out);
void fill_ones(Workspace& ws, u64* counts, u64 m);
void fill_value(Workspace& ws, u64* counts, u64 m, u64 v);
void dump_text(Workspace& ws, int key_bytes, const void* keys, const u64* counts, u64 m, int w, DevBuf<u8>& text_out, u64* bytes_out);

// ---- xeno.cu ---------------------------------------------------------------------------------
u64 bit_vector_words(u64 m);
void xeno_annotate_bits(Workspace& ws, const u64* weight, u64 m, DevBuf<u64>& lhs, DevBuf<u64>& rhs);
u64 xeno_near_kmers(Workspace& ws, int key_bytes, int k, const void* keys, u64 m, const u64* lhs, const u64* rhs, DevBuf<u64>& new_lhs, DevBuf<u64>& new_rhs);

struct ParseFailure { int code; u64 line; };
struct StatusError { int status; std::string message; };

}  // namespace gsb
