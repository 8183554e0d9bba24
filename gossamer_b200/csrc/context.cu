// context.cu -- the C ABI of include/gossamer_b200.h: context, memory, per-block pipeline,
// batching/merging, emission of the Graph / KmerSet file sets.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>

#include <sched.h>

#include "exchange.h"
#include "kernels.h"

namespace gsb {

// ------------------------------------------------------------------------------------------
// Workspace
// ------------------------------------------------------------------------------------------
// Size classes: powers of two below 64 KiB, multiples of 64 KiB below 1 MiB; above, multiples of max(2 MiB, 1/8 of the
// largest power of two below the request) -- at most 12.5 % slack.  A request takes the smallest cached block that is at
// least its class and at most twice (small blocks: four times) as large.  On one GPU every step repeats the same sequence
// of sizes and hits its own blocks exactly; with several GPUs the sizes move by a few per cent from step to step (how many
// keys a rank receives, how many survive), so a MISS allocates one class more than was asked for: the block then also
// serves the slightly larger requests of later steps, and a steady-state step makes no driver allocation at all
// (cudaMalloc costs tens of milliseconds with eight peer-mapped devices; gsb_stats.device_allocs counts them).
static size_t round_block(size_t bytes) {
    if (bytes < 512) return 512;
    if (bytes < (64u << 10)) { size_t p = 512; while (p < bytes) p <<= 1; return p; }
    if (bytes < (1u << 20)) return (bytes + 65535) & ~(size_t)65535;
    size_t p2 = (size_t)1 << (63 - __builtin_clzll((unsigned long long)bytes));
    size_t g = std::max<size_t>((size_t)2 << 20, p2 / 8);
    return (bytes + g - 1) / g * g;
}

void* Workspace::alloc(size_t bytes) {
    const size_t want = round_block(bytes);
    void* p = nullptr;
    auto it = free_.lower_bound(want);
    if (it != free_.end() && it->first <= (want >= (1u << 20) ? 2 * want : 4 * want)) {
        p = it->second;
        free_.erase(it);
    } else {
        const size_t get = want >= (1u << 20) ? round_block(want + want / 8) : want;   // head room for the next, slightly larger request
        cudaError_t e = cudaMalloc(&p, get);
        if (e != cudaSuccess) {                                            // out of memory: drop the cache and retry once, without head room
            cudaGetLastError();
            GSB_CUDA_TRY(cudaStreamSynchronize(stream));
            trim();
            e = cudaMalloc(&p, want);
            if (e == cudaSuccess) { size_of_[p] = want; reserved_bytes += want; ++device_allocs; }
        } else {
            size_of_[p] = get; reserved_bytes += get; ++device_allocs;
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            throw StatusError{GSB_ENOMEM, std::string("device allocation of ") + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e)};
        }
    }
    live_bytes += size_of_[p];
    peak_bytes = std::max(peak_bytes, live_bytes);
    return p;
}

void Workspace::release(void* p, size_t) {
    auto it = size_of_.find(p);
    if (it == size_of_.end()) return;
    live_bytes -= it->second;
    free_.insert({it->second, p});
}

void Workspace::trim() {
    for (auto& kv : free_) { cudaFree(kv.second); reserved_bytes -= kv.first; size_of_.erase(kv.second); }
    free_.clear();
}

Workspace::~Workspace() {
    if (stream) cudaStreamSynchronize(stream);
    trim();
}

void Workspace::sync() { GSB_CUDA_TRY(cudaStreamSynchronize(stream)); }

// Device time per phase: CUDA events on the library's stream around each phase, resolved LAZILY (gsb_get_stats /
// gsb_reset) -- recording an interval never makes the host wait for the device.
struct PhaseTimer {
    struct Interval { cudaEvent_t e0, e1; double* acc; };
    std::vector<Interval> open_, free_;
    cudaStream_t s = nullptr;
    Interval cur_{nullptr, nullptr, nullptr};
    bool running_ = false;
    void init(cudaStream_t st) { s = st; }
    void destroy() {
        for (auto& v : {&open_, &free_}) { for (auto& i : *v) { cudaEventDestroy(i.e0); cudaEventDestroy(i.e1); } v->clear(); }
    }
    void start() {
        if (running_) throw StatusError{GSB_EINVAL, "internal: phase timer started twice"};
        running_ = true;
        if (free_.empty()) {
            Interval i{nullptr, nullptr, nullptr};
            GSB_CUDA_TRY(cudaEventCreate(&i.e0)); GSB_CUDA_TRY(cudaEventCreate(&i.e1));
            free_.push_back(i);
        }
        cur_ = free_.back(); free_.pop_back();
        GSB_CUDA_TRY(cudaEventRecord(cur_.e0, s));
    }
    void stop(double& acc_ms) {
        running_ = false;
        GSB_CUDA_TRY(cudaEventRecord(cur_.e1, s));
        cur_.acc = &acc_ms;
        open_.push_back(cur_);
    }
    // an interval whose events were recorded elsewhere (partition.cu); the events are recycled like the timer's own
    void adopt(cudaEvent_t e0, cudaEvent_t e1, double& acc_ms) { if (e0 && e1) open_.push_back(Interval{e0, e1, &acc_ms}); }
    // adds every finished interval to its accumulator (waits for the last recorded event)
    void resolve() {
        for (auto& i : open_) {
            float ms = 0;
            if (cudaEventSynchronize(i.e1) == cudaSuccess && cudaEventElapsedTime(&ms, i.e0, i.e1) == cudaSuccess) *i.acc += ms;
            else cudaGetLastError();
            free_.push_back(i);
        }
        open_.clear();
    }
};

static const u64 kMaxBlockBytes = 1ull << 30;

}  // namespace gsb

using namespace gsb;

struct gsb_ctx {
    gsb_config cfg;
    Workspace ws;
    std::string err;
    int key_bytes = 8, key_bits = 0, window = 0, passes = 0;
    u64 self_rc_windows = 0;       // windows seen so far that equal their own reverse complement
    bool any_self_rc = true;       // (all ranks) whether any exist: if not, doubling / the special filter threshold are skipped
    bool mix = false;              // graph mode with a min-count filter: instances are stored bit-mixed (key_mix) for the
                                   // partial-sort counting of sort.cu; any other use of a batch un-mixes it first
    int fold_w = 0;                // graph mode: instances are strand-folded windows of this many symbols (fold.cu); 0 = kmer set

    // instance keys of the current batch
    DevBuf<u8> keys;
    u64 keys_cap = 0, n_keys = 0;
    DevBuf<u8> alt;                // second sort buffer (grow-only, kept across steps)
    u64 alt_cap = 0;
    DevBuf<u8> third;              // receive buffer of the instance all-to-all (multi-GPU only)
    u64 third_cap = 0;
    DevBuf<u64> cursor;            // device-side append cursor
    DevBuf<u64> hist;              // histograms fused into the extraction: [256] top byte of the low key word (partition counting) or
                                   // [passes][256] every digit (legacy LSD counting)
    bool hist_valid = false;       // ... describe exactly the current batch
    // streamed first pass (single GPU, partition counting): every block is split by its top bits as soon as it has been
    // extracted -- while the next block is still crossing PCIe -- into `alt` at the offsets it has in `keys`
    DevBuf<u64> runs;              // [blocks][2^bits0 + 1] absolute starts of every block's children
    u64 runs_cap = 0, n_runs = 0;  // blocks: capacity / split so far in this batch
    DevBuf<u64> hist_next;         // [2^(bits0 + kTopHistBits)] next-bits histogram of every child, summed over the blocks
    u64 n_streamed = 0;            // keys covered by the blocks split so far (== n_keys when the batch is fully streamed)
    DevBuf<IngestStatus> status;
    u64 max_batch_keys = 0;

    // per-format open-file state
    bool file_open[3] = {false, false, false};
    u64 line_base[3] = {0, 0, 0};
    DevBuf<u8> carry;              // last window-1 symbols of an unfinished FASTA file
    u32 n_carry = 0;

    // host<->device staging
    u8* pinned = nullptr;
    size_t pinned_bytes = 0;
    EmitRing ring;                   // overlapped device -> sink delivery of gsb_emit (created on first use)
    DevBuf<u8> text_dev;
    size_t text_cap = 0;
    // GSB_BLOCK_ASYNC: two text buffers filled by a copy stream while the main stream works on the other one
    cudaStream_t copy_stream = nullptr;
    DevBuf<u8> atext[2];
    size_t atext_cap[2] = {0, 0};
    cudaEvent_t copied[2] = {nullptr, nullptr}, processed[2] = {nullptr, nullptr}, copy_begin = nullptr;
    bool processed_valid[2] = {false, false};
    struct { bool valid = false; int buf = 0; size_t nbytes = 0; int format = 0; u32 flags = 0; } pending;
    int next_abuf = 0;

    ReducedRun acc;                // merged (key,count) run of all flushed batches
    bool have_acc = false;
    bool counted = false;
    gsb_counts counts;
    gsb_stats stats;
    PhaseTimer timer;
    cudaEvent_t user_e0 = nullptr, user_e1 = nullptr;
    Exchange* comm = nullptr;
    bool exchanged_instances = false;
    u64 instances_before_exchange = 0;
    u8* batch_src = nullptr;       // where the current batch's instances are when not in `keys` (receive window)
    DistRun dist;                  // multi-GPU: the final run as published in the ranks' peer-mapped windows
    bool dist_ready = false;       // ... is valid: gsb_emit writes this rank's byte ranges of every file
    bool acc_unsorted = false;     // acc came out of reduce_groups: folded, filtered, final counts, arbitrary order
    bool gathered = false;         // gsb_gather_to_root was called: rank 0 holds and emits everything
    u64 m_est = 0;                 // size estimate the builders are parameterised with (0 = the number of items emitted)
    u64 loaded_items = 0;          // gsb_graph_load: sum of the sizes of the file sets loaded so far

    void log(int sev, const std::string& m) { if (cfg.log) cfg.log(cfg.log_user, sev, m.c_str()); }
};

namespace {

// test / profiling switch (gsb_debug_set_tuning bit 16): count by the full LSD sort of raw keys instead of by partitioning
int g_legacy_counting = 0;
int g_sampled_survivors = 0;        // test switch: re-partition the survivors with sampled splitters (the fallback path)

thread_local std::string g_create_error;

// leaves only when the emitter's writer thread has handed over (or dropped) everything queued: the caller's sink may go away
struct RingGuard {
    EmitRing* ring;
    ~RingGuard() { if (ring && ring->created) ring->wait_idle(); }
};

template <typename F>
int guarded(gsb_ctx* ctx, F&& body) {
    try {
        if (ctx) { GSB_CUDA_TRY(cudaSetDevice(ctx->ws.device)); ctx->timer.running_ = false; }   // an earlier call may have thrown inside a phase
        body();
        return GSB_OK;
    } catch (const StatusError& e) {
        if (ctx) ctx->err = e.message; else g_create_error = e.message;
        return e.status;
    } catch (const CudaError& e) {
        std::string m = std::string("CUDA error: ") + cudaGetErrorString(e.code) + " at " + e.file + ":" + std::to_string(e.line) + " (" + e.expr + ")";
        cudaGetLastError();
        if (ctx) ctx->err = m; else g_create_error = m;
        return GSB_ECUDA;
    } catch (const std::bad_alloc&) {
        if (ctx) ctx->err = "host allocation failed"; else g_create_error = "host allocation failed";
        return GSB_ENOMEM;
    } catch (const std::exception& e) {
        if (ctx) ctx->err = e.what(); else g_create_error = e.what();
        return GSB_EINVAL;
    }
}

std::string parse_message(int code, u64 line) {
    const std::string n = std::to_string(line);
    switch (code) {
        case GSB_PE_FASTA_EXPECT_GT: return "expected '>' at beginning of line " + n;
        case GSB_PE_FASTQ_EXPECT_AT: return "expected '@' at beginning of line " + n;
        case GSB_PE_FASTQ_EXPECT_SEQ: return "expected sequence data or quality header at line " + n;
        case GSB_PE_FASTQ_EXPECT_PLUS: return "expected '+' at beginning of line " + n;
        case GSB_PE_FASTQ_TITLE_MISMATCH: return "quality title does not match sequence title at line " + n;
        case GSB_PE_FASTQ_LEN_MISMATCH: return "length mistmatch between sequence and quality data just before line " + n;
        default: return "internal ingest error " + std::to_string(code);
    }
}

void reset_batch(gsb_ctx* c) {
    cudaStream_t s = c->ws.stream;
    GSB_CUDA_TRY(cudaMemsetAsync(c->cursor.p, 0, 8, s));
    GSB_CUDA_TRY(cudaMemsetAsync(c->hist.p, 0, c->hist.bytes(), s));
    c->hist_valid = c->mix || c->comm == nullptr;              // legacy counting with a communicator: the histograms are taken after the exchange
    c->n_keys = 0;
    c->n_runs = 0; c->n_streamed = 0;
    if (c->hist_next.p) GSB_CUDA_TRY(cudaMemsetAsync(c->hist_next.p, 0, c->hist_next.bytes(), s));
}

// merge two sorted reduced runs (merge path, fold.cu), then sum the counts of equal keys -- what AsyncMerge / PairMerge
// do one item at a time (src/AsyncMerge.tcc:267-324, src/GossCmdMerge.tcc:84-145).  Linear in the sizes of the runs.
void merge_runs(gsb_ctx* c, ReducedRun& into, ReducedRun& other) {
    Workspace& ws = c->ws;
    const u64 n = into.m + other.m;
    const int kb = c->key_bytes;
    if (other.m == 0) return;
    if (into.m == 0) { into = std::move(other); other.m = 0; return; }
    DevBuf<u8> mk(&ws, n * kb);
    DevBuf<u64> mc(&ws, n);
    merge_disjoint_runs(ws, kb, into.keys.p, into.counts.p, into.m, other.keys.p, other.counts.p, other.m, mk.p, mc.p);
    into.keys.free(); into.counts.free(); other.keys.free(); other.counts.free();
    ReducedRun merged; u64 distinct = 0;
    reduce_sorted(ws, kb, mk.p, mc.p, n, 1, merged, &distinct);
    into = std::move(merged);
    other.m = 0;
}

void ensure_alt(gsb_ctx* c, u64 n) {
    if (n <= c->alt_cap) return;
    c->alt.free();
    c->alt_cap = std::max<u64>(n, c->keys_cap);
    c->alt.reset(&c->ws, c->alt_cap * c->key_bytes + 64);         // slack: bulk tile loads are rounded up to 16 bytes
}

static bool streaming(const gsb_ctx* c) { return c->mix && !c->comm; }

// First partition pass over the block that has just been extracted: keys[off, off + n_blk) -> alt (same offsets).
void stream_block(gsb_ctx* c, u64 off, u64 n_blk) {
    Workspace& ws = c->ws;
    const int kb = c->key_bytes;
    const int bits0 = partition_stream_bits0();
    const u64 C = 1ull << bits0;
    if (c->alt_cap < c->keys_cap) {                              // alt mirrors keys; what earlier blocks put there is kept
        DevBuf<u8> bigger(&ws, c->keys_cap * kb + 64);
        if (off && c->alt.p) GSB_CUDA_TRY(cudaMemcpyAsync(bigger.p, c->alt.p, off * kb, cudaMemcpyDeviceToDevice, ws.stream));
        c->alt = std::move(bigger);
        c->alt_cap = c->keys_cap;
    }
    if (c->n_runs + 1 > c->runs_cap) {
        const u64 want = std::max<u64>(16, 2 * c->runs_cap);
        DevBuf<u64> bigger(&ws, want * (C + 1));
        if (c->n_runs) GSB_CUDA_TRY(cudaMemcpyAsync(bigger.p, c->runs.p, c->n_runs * (C + 1) * 8, cudaMemcpyDeviceToDevice, ws.stream));
        c->runs = std::move(bigger);
        c->runs_cap = want;
    }
    if (!c->hist_next.p || c->hist_next.n != ((size_t)C << kTopHistBits)) {
        if (c->n_runs) throw StatusError{GSB_EINVAL, "internal: the first-pass width changed inside a batch"};
        c->hist_next.reset(&ws, (size_t)C << kTopHistBits);
        GSB_CUDA_TRY(cudaMemsetAsync(c->hist_next.p, 0, c->hist_next.bytes(), ws.stream));
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    c->timer.start();
    partition_block_level0(ws, kb, c->keys.p, c->alt.p, off, n_blk, bits0, c->hist.p, c->runs.p + c->n_runs * (C + 1), c->hist_next.p, &e0, &e1);
    c->timer.stop(c->stats.ms_sort);
    c->timer.adopt(e0, e1, c->stats.ms_sort_sweeps);             // the scatter launch alone
    if (e0) c->stats.sort_passes += 1;
    ++c->n_runs;
    c->n_streamed = off + n_blk;
}

// order a run of distinct (key, count) pairs by key
void sort_run(gsb_ctx* c, ReducedRun& run) {
    if (run.m < 2) return;
    Workspace& ws = c->ws;
    {
        ReducedRun sorted;
        if (sort_pairs_msd(ws, c->key_bytes, c->key_bits, run.keys.p, run.counts.p, run.m, 0, sorted)) { run = std::move(sorted); return; }
    }
    DevBuf<u8> kalt(&ws, run.m * c->key_bytes);
    DevBuf<u64> calt(&ws, run.m);
    const int where = sort_keys(ws, c->key_bytes, c->key_bits, run.keys.p, kalt.p, run.counts.p, calt.p, run.m, nullptr, nullptr);
    if (where) { run.keys = std::move(kalt); run.counts = std::move(calt); }
}

// count the buffered instance keys; fold the result into the accumulated run
// pre / pre_plan: the batch has been through the multi-GPU exchange already (level 0 of the partition counting)
void flush_batch(gsb_ctx* c, bool final_and_only, PartitionInput* pre = nullptr, const PartitionPlan* pre_plan = nullptr) {
    if (c->n_keys == 0 && !pre) return;
    Workspace& ws = c->ws;
    const int kb = c->key_bytes;
    if (!pre) ensure_alt(c, c->n_keys);
    u8* const src = pre ? (u8*)pre->keys : (c->batch_src ? c->batch_src : c->keys.p);   // the batch
    u8* const alt = pre ? (u8*)pre->scratch : c->alt.p;                                // scratch of the same capacity
    c->batch_src = nullptr;
    int passes_run = 0;
    ReducedRun run; u64 distinct = 0, n_self_rc = 0;
    // The only batch of a build (on this rank, after any instance exchange): doubling of the
    // self-complementary keys and the min-count filter are fused into the run-length reduce.
    // Otherwise the run stays folded with raw occurrence counts until every batch has been merged.
    const bool fused_final = final_and_only && (!c->comm || c->exchanged_instances);
    const u64 min_count = (fused_final && c->cfg.kind == GSB_KIND_GRAPH) ? std::max<u64>(1, c->cfg.min_count) : 1;
    const int fold_w = (fused_final && c->any_self_rc) ? c->fold_w : 0;
    // Counting by partitioning (partition.cu): the instances are stored bit-mixed; two most-significant-digit passes bring
    // equal keys into one bucket of ~3000 instances, which one CTA counts exactly in shared memory.  The survivors come back
    // un-mixed in arbitrary order; whoever needs them ordered sorts the (few) distinct keys afterwards.
    int where = 0;
    bool reduced = false;
    bool streamed_in = false;
    if (c->mix) {
        PartitionTiming pt;
        PartitionInput in;
        PartitionPlan plan;
        if (pre) {
            in = std::move(*pre);
            plan = *pre_plan;
        } else if (c->n_runs && c->n_streamed == c->n_keys && src == c->keys.p) {
            // every block of the batch has been through the first pass already (stream_block): the blocks' runs lie in alt
            in.keys = alt; in.scratch = src; in.n = c->n_keys;
            in.runs = c->runs.p; in.runs_n_src = (int)c->n_runs; in.runs_bits0 = partition_stream_bits0(); in.runs_hist_next = c->hist_next.p;
            plan = partition_plan_streamed(kb, c->n_keys, in.runs_bits0);
            streamed_in = true;
        } else {
            in.keys = src; in.scratch = alt; in.n = c->n_keys;
            // the fused top-bit histogram describes exactly this batch only if it came straight out of the extraction
            in.hist_top = (c->hist_valid && !c->exchanged_instances) ? c->hist.p : nullptr;
            plan = partition_plan(kb, c->n_keys);
        }
        reduced = count_partitioned(ws, kb, c->key_bits, in, plan, min_count, fold_w, run, &distinct, &n_self_rc, &where, &pt);
        if (streamed_in) where ^= 1;                               // where is relative to in.keys, which was alt here
        if (pre && !reduced) c->n_keys = exchange_partition_received(c->comm);   // rare: the full sort below needs the exact number of keys that arrived
        c->stats.ms_sort += pt.ms_partition;
        c->stats.ms_reduce += pt.ms_count;
        c->stats.ms_sort_sweeps += pt.ms_scatter;
        c->stats.sort_passes += pt.scatter_launches;
        passes_run = 0;
        if (reduced) {
            if (fused_final) {
                c->acc_unsorted = true;
            } else if (run.m) {                                    // a batch that will be merged with others: order it by key
                c->timer.start();
                sort_run(c, run);
                c->timer.stop(c->stats.ms_sort);
            }
        } else {
            c->log(0, "partition counting declined (hardly any duplication under a min-count filter): sorting by the full key instead");
        }
    }
    if (!reduced) {
        u8* const from = where ? alt : src;
        u8* const other = where ? src : alt;
        const u64* hist = nullptr;
        c->timer.start();
        if (c->mix) sort_unmix_inplace(kb, from, c->n_keys, ws.sm_count, ws.stream, &ws.launches);   // real keys again
        else if (c->hist_valid) hist = c->hist.p;
        passes_run = 0;
        const int where2 = sort_keys(ws, kb, c->key_bits, from, other, nullptr, nullptr, c->n_keys, hist, &passes_run, &c->stats.ms_sort_sweeps);
        where ^= where2;
        c->timer.stop(c->stats.ms_sort);
        c->stats.sort_passes += passes_run;
    }
    c->stats.sort_passes_model += c->passes;
    c->stats.n_batches += 1;
    if (!reduced) {
        c->timer.start();
        reduce_sorted(ws, kb, where ? alt : src, nullptr, c->n_keys, min_count, run, &distinct, fold_w, &n_self_rc);
        c->timer.stop(c->stats.ms_reduce);
    }
    c->counts.n_instances += c->n_keys * (c->fold_w ? 2 : 1);  // the reference counts both strands (src/ReverseComplementAdapter.hh:34-55)
    c->stats.n_sorted_keys += c->n_keys;
    if (fused_final) c->counts.n_distinct = c->fold_w ? 2 * distinct - n_self_rc : distinct;
    if (c->have_acc) {
        c->timer.start();
        merge_runs(c, c->acc, run);
        c->timer.stop(c->stats.ms_merge);
    } else {
        c->acc = std::move(run);
        c->have_acc = true;
    }
    reset_batch(c);
}

void ensure_key_capacity(gsb_ctx* c, u64 extra) {
    const int kb = c->key_bytes;
    if (c->n_keys + extra <= c->keys_cap) return;
    if (c->n_keys + extra > c->max_batch_keys && c->n_keys > 0) {
        c->log(0, "key buffer full: sorting and reducing a batch of " + std::to_string(c->n_keys) + " keys");
        flush_batch(c, false);
    }
    if (c->n_keys + extra <= c->keys_cap) return;
    u64 want = std::max<u64>(c->n_keys + extra, std::min<u64>(c->keys_cap * 2, c->max_batch_keys));
    DevBuf<u8> bigger(&c->ws, want * kb + 64);
    if (c->n_keys) GSB_CUDA_TRY(cudaMemcpyAsync(bigger.p, c->keys.p, c->n_keys * kb, cudaMemcpyDeviceToDevice, c->ws.stream));
    c->keys = std::move(bigger);
    c->keys_cap = want;
}

// one raw text block already on the device
void process_block(gsb_ctx* c, const u8* text, u64 n, int format, u32 flags) {
    Workspace& ws = c->ws;
    cudaStream_t s = ws.stream;
    if (format < 0 || format > 2) throw StatusError{GSB_EINVAL, "unknown input format"};
    if (n > kMaxBlockBytes) throw StatusError{GSB_EINVAL, "input blocks are limited to 1 GiB; split the file at record boundaries"};
    if (c->counted) throw StatusError{GSB_EINVAL, "gsb_push_block after gsb_finish_counting (call gsb_reset first)"};
    const bool file_start = !c->file_open[format];
    const bool last = (flags & GSB_BLOCK_LAST_OF_FILE) != 0;
    c->stats.bytes_in += n;

    // K1: line table
    c->timer.start();
    GSB_CUDA_TRY(cudaMemsetAsync(c->status.p, 0, sizeof(IngestStatus), s));
    const u32 tiles = ingest_newline_tiles(n);
    DevBuf<u32> tile_counts(&ws, (size_t)tiles + 1), scalars(&ws, 4);
    ingest_count_newlines(text, n, tile_counts.p, s, &ws.launches);
    ingest_scan_tiles(tile_counts.p, tiles, scalars.p, s, &ws.launches);
    u32 n_newlines = 0; u8 last_byte = '\n';
    GSB_CUDA_TRY(cudaMemcpyAsync(&n_newlines, scalars.p, 4, cudaMemcpyDeviceToHost, s));
    if (n) GSB_CUDA_TRY(cudaMemcpyAsync(&last_byte, text + n - 1, 1, cudaMemcpyDeviceToHost, s));
    ws.sync();
    const u32 n_lines = n_newlines + ((n > 0 && last_byte != '\n') ? 1u : 0u);
    DevBuf<u32> line_start(&ws, (size_t)n_lines + 1), nsym(&ws, (size_t)n_lines + 1), sym_off(&ws, (size_t)n_lines + 1);
    DevBuf<u8> kind(&ws, (size_t)n_lines + 1);
    DevBuf<u32> scan_tmp(&ws, (size_t)(n_lines / 2048 + 2));
    GSB_CUDA_TRY(cudaMemsetAsync(line_start.p, 0, 4, s));
    ingest_fill_line_starts(text, n, tile_counts.p, line_start.p, n_lines, s, &ws.launches);

    // K1b: framing
    {
        DevBuf<u32> fq_scratch;                                // irregular FASTQ layouts are framed in parallel (speculate + verify, ingest.cu)
        if (format == GSB_FMT_FASTQ && n_lines) fq_scratch.reset(&ws, ingest_fastq_scratch_words(n_lines));
        ingest_classify(text, line_start.p, n_lines, format, file_start ? 1 : 0, c->line_base[format], kind.p, nsym.p, c->status.p, s, &ws.launches, fq_scratch.p);
    }
    ingest_symbol_offsets(nsym.p, sym_off.p, n_lines, scalars.p + 1, scan_tmp.p, s, &ws.launches);
    GSB_CUDA_TRY(cudaMemcpyAsync(sym_off.p + n_lines, scalars.p + 1, 4, cudaMemcpyDeviceToDevice, s));
    IngestStatus st; u32 block_syms = 0;
    GSB_CUDA_TRY(cudaMemcpyAsync(&st, c->status.p, sizeof(st), cudaMemcpyDeviceToHost, s));
    GSB_CUDA_TRY(cudaMemcpyAsync(&block_syms, scalars.p + 1, 4, cudaMemcpyDeviceToHost, s));
    c->timer.stop(c->stats.ms_scan);
    if (st.error) throw StatusError{GSB_EPARSE, parse_message(st.error, st.error_line)};
    c->counts.n_reads += st.n_reads;

    ensure_key_capacity(c, (u64)block_syms);           // one (folded) key per window at most; may sort+reduce the current batch first

    // K2: packed symbol stream = [64 pad][carry][block]
    c->timer.start();
    const u32 n_carry = (format == GSB_FMT_FASTA && !file_start) ? c->n_carry : 0;
    const u64 n_sym_total = 64 + (u64)n_carry + block_syms;
    const u64 n_words = (n_sym_total + 31) / 32;
    DevBuf<u64> codes(&ws, n_words);
    DevBuf<u32> valid(&ws, n_words);
    {
        DevBuf<u32> word_line(&ws, n_words);
        ingest_pack(text, n, line_start.p, kind.p, sym_off.p, n_lines, c->carry.p, n_carry, n_sym_total, codes.p, valid.p, n_words, word_line.p, s, &ws.launches);
    }
    line_start.free(); nsym.free(); sym_off.free(); kind.free();
    c->stats.n_symbols += block_syms;

    // K3: windows -> keys (+ fused digit histograms)
    // (with a communicator attached the instances are exchanged before the sort and the digit histograms are taken from
    // what arrives, so the fused histograms -- the dominant cost of the kernel -- are switched off)
    const u64 keys_before = c->n_keys;
    if (streaming(c)) {                                          // the fused histogram describes THIS block: its first pass runs right behind the extraction
        GSB_CUDA_TRY(cudaMemsetAsync(c->hist.p, 0, c->hist.bytes(), s));
        c->hist_valid = false;
    }
    ingest_extract(c->cfg.kind, c->key_bytes, codes.p, valid.p, 64 + n_carry, n_sym_total, c->window, c->mix ? -1 : (c->comm ? 0 : c->passes), c->mix ? 1 : 0,
                   c->keys.p, c->cursor.p, c->keys_cap, c->hist.p, c->status.p, ws.sm_count, s, &ws.launches);
    u64 cur = 0;
    GSB_CUDA_TRY(cudaMemcpyAsync(&cur, c->cursor.p, 8, cudaMemcpyDeviceToHost, s));
    GSB_CUDA_TRY(cudaMemcpyAsync(&st, c->status.p, sizeof(st), cudaMemcpyDeviceToHost, s));
    // carry for a FASTA file that continues in the next block
    if (format == GSB_FMT_FASTA && !last) {
        const u32 want = (u32)std::min<u64>((u64)c->window - 1, n_sym_total - 64);
        ingest_save_carry(codes.p, valid.p, n_sym_total, want, c->carry.p, s, &ws.launches);
        c->n_carry = want;
    } else if (format == GSB_FMT_FASTA) {
        c->n_carry = 0;
    }
    c->timer.stop(c->stats.ms_extract);
    if (st.error) throw StatusError{GSB_EINVAL, "key buffer overflow (internal sizing error)"};
    c->self_rc_windows += st.n_self_rc;
    c->n_keys = cur;
    if (streaming(c) && cur > keys_before) stream_block(c, keys_before, cur - keys_before);
    c->file_open[format] = !last;
    c->line_base[format] = last ? 0 : c->line_base[format] + n_lines;
}

void write_graph_files(gsb_ctx* c, Emitter& em, const std::string& prefix) {
    const u64 k = (u64)c->cfg.k;
    const u64 m = c->acc.m;
    const u64 m_est = c->m_est ? c->m_est : m;           // pNumEdges of Graph::Builder (src/Graph.cc:145-158)
    // Graph::Builder ctor writes the header first (src/Graph.cc:159-166)
    u64 header[3] = {2011101014ull, k, 0};
    em.put_host(prefix + ".header", header, sizeof(header));
    const unsigned rho2 = 2 * (unsigned)(k + 1);
    U128 universe = rho2 < 64 ? U128{1ull << rho2, 0} : U128{0, 1ull << (rho2 - 64)};
    emit_sparse_array(em, c->key_bytes, c->acc.keys.p, m, universe, m_est, universe, prefix + "-edges");
    emit_counts(em, c->acc.counts.p, m, m_est, prefix + "-counts");
    emit_count_histogram(em, c->acc.counts.p, m, prefix + "-counts-hist.txt");
}

void write_kmer_set_files(gsb_ctx* c, Emitter& em, const std::string& prefix) {
    const u64 k = (u64)c->cfg.k;
    const u64 m = c->acc.m;
    const unsigned bits = 2 * (unsigned)k;
    U128 universe = bits < 64 ? U128{1ull << bits, 0} : U128{0, 1ull << (bits - 64)};
    emit_sparse_array(em, c->key_bytes, c->acc.keys.p, m, universe, c->m_est ? c->m_est : m, universe, prefix + ".kmers");
    u64 header[3] = {2011101701ull, k, m};             // KmerSet::Builder::end, src/KmerSet.hh:76-83
    em.put_host(prefix + ".header", header, sizeof(header));
}

// multi-GPU: every rank writes its own byte ranges of every file (emit.cu, "Multi-GPU emission")
void write_graph_files_dist(gsb_ctx* c, Emitter& em, const std::string& prefix) {
    const u64 k = (u64)c->cfg.k;
    const u64 m = c->dist.off[c->dist.n];
    if (c->dist.rank == 0) {
        u64 header[3] = {2011101014ull, k, 0};
        em.put_host(prefix + ".header", header, sizeof(header));
    }
    const unsigned rho2 = 2 * (unsigned)(k + 1);
    U128 universe = rho2 < 64 ? U128{1ull << rho2, 0} : U128{0, 1ull << (rho2 - 64)};
    emit_sparse_array_dist(em, c->comm, c->key_bytes, c->dist, universe, m, universe, prefix + "-edges");
    emit_counts_dist(em, c->comm, c->dist, m, prefix + "-counts");
    emit_count_histogram_dist(em, c->comm, c->dist, prefix + "-counts-hist.txt");
}

void write_kmer_set_files_dist(gsb_ctx* c, Emitter& em, const std::string& prefix) {
    const u64 k = (u64)c->cfg.k;
    const u64 m = c->dist.off[c->dist.n];
    const unsigned bits = 2 * (unsigned)k;
    U128 universe = bits < 64 ? U128{1ull << bits, 0} : U128{0, 1ull << (bits - 64)};
    emit_sparse_array_dist(em, c->comm, c->key_bytes, c->dist, universe, m, universe, prefix + ".kmers");
    if (c->dist.rank == 0) {
        u64 header[3] = {2011101701ull, k, m};
        em.put_host(prefix + ".header", header, sizeof(header));
    }
}

void init_ctx(gsb_ctx* c) {
    const gsb_config& cfg = c->cfg;
    if (cfg.abi_version != GSB_ABI_VERSION) throw StatusError{GSB_EINVAL, "ABI version mismatch"};
    if (cfg.kind != GSB_KIND_GRAPH && cfg.kind != GSB_KIND_KMERSET) throw StatusError{GSB_EINVAL, "unknown kind"};
    const int max_k = cfg.kind == GSB_KIND_GRAPH ? 62 : 63;
    if (cfg.k < 1 || cfg.k > max_k)
        throw StatusError{GSB_ERANGE, "unable to build a graph with k=" + std::to_string(cfg.k)};
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        throw StatusError{GSB_ECUDA, "no CUDA device available (this library has no CPU path)"};
    }
    if (cfg.device < 0 || cfg.device >= n_dev) throw StatusError{GSB_EINVAL, "bad device ordinal"};
    c->ws.device = cfg.device;
    GSB_CUDA_TRY(cudaSetDevice(cfg.device));
    cudaDeviceProp prop;
    GSB_CUDA_TRY(cudaGetDeviceProperties(&prop, cfg.device));
    if (prop.major < 10) throw StatusError{GSB_ECUDA, std::string("device ") + prop.name + " is not sm_100-class; this build targets B200 only"};
    c->ws.sm_count = prop.multiProcessorCount;
    GSB_CUDA_TRY(cudaStreamCreateWithFlags(&c->ws.stream, cudaStreamNonBlocking));
    c->timer.init(c->ws.stream);
    GSB_CUDA_TRY(cudaEventCreate(&c->user_e0));
    GSB_CUDA_TRY(cudaEventCreate(&c->user_e1));

    c->window = cfg.kind == GSB_KIND_GRAPH ? cfg.k + 1 : cfg.k;
    c->fold_w = cfg.kind == GSB_KIND_GRAPH ? c->window : 0;
    c->mix = !g_legacy_counting;
    c->key_bits = 2 * c->window;
    c->key_bytes = c->key_bits <= 64 ? 8 : 16;
    c->passes = (c->key_bits + 7) / 8;
    c->cursor.reset(&c->ws, 1);
    c->hist.reset(&c->ws, std::max<size_t>((size_t)c->passes * 256, (size_t)1 << kTopHistBits));
    c->status.reset(&c->ws, 1);
    c->carry.reset(&c->ws, 64);
    size_t free_b = 0, total_b = 0;
    GSB_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    // per buffered key: two sort buffers + run-length positions
    const u64 per_key = 2 * (u64)c->key_bytes + 8;
    c->max_batch_keys = cfg.max_batch_keys ? cfg.max_batch_keys : (u64)(free_b * 0.7) / per_key;
    c->pinned_bytes = 64ull << 20;
    GSB_CUDA_TRY(cudaMallocHost((void**)&c->pinned, c->pinned_bytes));
    memset(&c->counts, 0, sizeof(c->counts));
    memset(&c->stats, 0, sizeof(c->stats));
    c->stats.sort_key_bytes = c->key_bytes;
    reset_batch(c);
    c->ws.sync();
}

// process the block whose asynchronous copy was started by the previous GSB_BLOCK_ASYNC push
void drain_pending(gsb_ctx* c) {
    if (!c->pending.valid) return;
    const int b = c->pending.buf;
    c->pending.valid = false;
    GSB_CUDA_TRY(cudaStreamWaitEvent(c->ws.stream, c->copied[b], 0));
    process_block(c, c->atext[b].p, c->pending.nbytes, c->pending.format, c->pending.flags & ~GSB_BLOCK_ASYNC);
    GSB_CUDA_TRY(cudaEventRecord(c->processed[b], c->ws.stream));
    c->processed_valid[b] = true;
}

void push_async(gsb_ctx* c, const void* data, size_t nbytes, int format, u32 flags) {
    if (!c->copy_stream) {
        GSB_CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            GSB_CUDA_TRY(cudaEventCreateWithFlags(&c->copied[i], cudaEventDisableTiming));
            GSB_CUDA_TRY(cudaEventCreateWithFlags(&c->processed[i], cudaEventDisableTiming));
        }
    }
    const int b = c->next_abuf;
    c->next_abuf ^= 1;
    if (nbytes + 16 > c->atext_cap[b]) {
        if (c->processed_valid[b]) GSB_CUDA_TRY(cudaEventSynchronize(c->processed[b]));
        c->atext[b].free();
        c->atext_cap[b] = std::max<size_t>(nbytes + 16, 1 << 20);
        c->atext[b].reset(&c->ws, c->atext_cap[b]);
        c->ws.sync();                                            // the block may be recycled memory still in use on the main stream
    }
    // buffer b was last read by the block pushed two calls ago
    if (c->processed_valid[b]) GSB_CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->processed[b], 0));
    if (nbytes) GSB_CUDA_TRY(cudaMemcpyAsync(c->atext[b].p, data, nbytes, cudaMemcpyHostToDevice, c->copy_stream));
    GSB_CUDA_TRY(cudaEventRecord(c->copied[b], c->copy_stream));
    drain_pending(c);                                            // device work of the previous block overlaps this copy
    c->pending.valid = true; c->pending.buf = b; c->pending.nbytes = nbytes; c->pending.format = format; c->pending.flags = flags;
}

}  // namespace

extern "C" {

int gsb_create(const gsb_config* cfg, gsb_ctx** out) {
    if (!cfg || !out) { g_create_error = "null argument"; return GSB_EINVAL; }
    *out = nullptr;
    gsb_ctx* c = new (std::nothrow) gsb_ctx();
    if (!c) { g_create_error = "host allocation failed"; return GSB_ENOMEM; }
    c->cfg = *cfg;
    int rc = guarded(nullptr, [&] { init_ctx(c); });
    if (rc != GSB_OK) { gsb_destroy(c); return rc; }
    *out = c;
    return GSB_OK;
}

void gsb_destroy(gsb_ctx* c) {
    if (!c) return;
    if (c->ws.stream) {
        cudaSetDevice(c->ws.device);
        cudaStreamSynchronize(c->ws.stream);
    }
    if (c->comm) { exchange_destroy(c->comm); c->comm = nullptr; }
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        for (int i = 0; i < 2; ++i) { if (c->copied[i]) cudaEventDestroy(c->copied[i]); if (c->processed[i]) cudaEventDestroy(c->processed[i]); c->atext[i].free(); }
        cudaStreamDestroy(c->copy_stream);
        c->copy_stream = nullptr;
    }
    c->keys.free(); c->alt.free(); c->third.free(); c->cursor.free(); c->hist.free(); c->status.free(); c->carry.free(); c->text_dev.free();
    c->acc.keys.free(); c->acc.counts.free();
    c->timer.destroy();
    if (c->user_e0) cudaEventDestroy(c->user_e0);
    if (c->user_e1) cudaEventDestroy(c->user_e1);
    c->ring.destroy(c->ws);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->ws.stream) { cudaStreamSynchronize(c->ws.stream); c->ws.trim(); cudaStreamDestroy(c->ws.stream); c->ws.stream = nullptr; }
    delete c;
}

const char* gsb_last_error(const gsb_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int gsb_push_block(gsb_ctx* c, const void* data, size_t nbytes, int format, uint32_t flags) {
    if (!c || (!data && nbytes)) return GSB_EINVAL;
    return guarded(c, [&] {
        if (nbytes > kMaxBlockBytes) throw StatusError{GSB_EINVAL, "input blocks are limited to 1 GiB; split the file at record boundaries"};
        if (c->counted) throw StatusError{GSB_EINVAL, "gsb_push_block after gsb_finish_counting (call gsb_reset first)"};
        if (flags & GSB_BLOCK_ASYNC) { push_async(c, data, nbytes, format, flags); return; }
        drain_pending(c);
        if (nbytes + 16 > c->text_cap) {
            c->text_dev.free();
            c->text_cap = std::max<size_t>(nbytes + 16, 1 << 20);
            c->text_dev.reset(&c->ws, c->text_cap);
        }
        c->timer.start();
        if (nbytes) GSB_CUDA_TRY(cudaMemcpyAsync(c->text_dev.p, data, nbytes, cudaMemcpyHostToDevice, c->ws.stream));
        c->timer.stop(c->stats.ms_h2d);
        process_block(c, c->text_dev.p, nbytes, format, flags);
    });
}

int gsb_push_device_block(gsb_ctx* c, const void* device_data, size_t nbytes, int format, uint32_t flags) {
    if (!c || (!device_data && nbytes)) return GSB_EINVAL;
    return guarded(c, [&] {
        drain_pending(c);
        const u8* text = (const u8*)device_data;
        if (((uintptr_t)text & 15) != 0) {                     // vector loads need 16-byte alignment: take a private copy
            if (nbytes + 16 > c->text_cap) {
                c->text_dev.free();
                c->text_cap = std::max<size_t>(nbytes + 16, 1 << 20);
                c->text_dev.reset(&c->ws, c->text_cap);
            }
            GSB_CUDA_TRY(cudaMemcpyAsync(c->text_dev.p, device_data, nbytes, cudaMemcpyDeviceToDevice, c->ws.stream));
            text = c->text_dev.p;
        }
        process_block(c, text, nbytes, format, flags);
    });
}

int gsb_finish_counting(gsb_ctx* c, gsb_counts* out) {
    if (!c) return GSB_EINVAL;
    return guarded(c, [&] {
        drain_pending(c);
        if (!c->counted) {
            // Every rank must take the same sequence of collectives: whether ANY rank has flushed a batch already (then all
            // ranks merge runs instead of exchanging raw instances) and whether any rank saw a self-complementary window are
            // agreed on first.
            u64 spilled = c->have_acc ? 1 : 0, self_rc_all = c->self_rc_windows, n_total = c->n_keys;
            std::vector<u64> n_keys_all;
            if (c->comm) {
                const u64 mine[3] = {spilled, self_rc_all, c->n_keys};
                std::vector<u64> all;
                c->timer.start();
                exchange_allgather_u64(c->comm, c->ws, mine, 3, all);
                c->timer.stop(c->stats.ms_exchange_agree);
                spilled = 0; self_rc_all = 0; n_total = 0;
                for (int r = 0; r < exchange_size(c->comm); ++r) {
                    spilled += all[3 * r]; self_rc_all += all[3 * r + 1]; n_total += all[3 * r + 2];
                    n_keys_all.push_back(all[3 * r + 2]);
                }
            }
            const bool single = spilled == 0;
            c->any_self_rc = self_rc_all > 0;
            bool exchanged_instances = false, counted_batch = false;
            if (c->comm && single) {
                // Multi-GPU, everything still buffered as raw instances: route each instance to the rank that owns its key
                // range FIRST (one all-to-all of raw keys over NVLink), then count locally exactly as on one GPU.
                const u64 local_instances = c->n_keys;
                c->instances_before_exchange = local_instances;
                if (c->mix) {
                    // the exchange IS the first pass of the partition counting: children of the top bits of the mixed key are
                    // stored straight into their owners' windows; no sampling, no host round trip (exchange.cu)
                    int min_bits = 1;
                    while ((1 << min_bits) < exchange_size(c->comm)) ++min_bits;
                    PartitionPlan plan = partition_plan(c->key_bytes, n_total);
                    if (plan.levels == 0) { plan.levels = 1; plan.bits[0] = min_bits; plan.total_bits = min_bits; }
                    else if (plan.bits[0] < min_bits) { plan.total_bits += min_bits - plan.bits[0]; plan.bits[0] = min_bits; }
                    // this rank's share of a uniform split + 1/8 + a little (every rank computes the same figure)
                    const u64 share = n_total / (u64)exchange_size(c->comm);
                    const u64 out_cap = share + share / 8 + (1u << 16);
                    ensure_alt(c, out_cap);
                    if (out_cap > c->third_cap) { c->third.free(); c->third_cap = out_cap; c->third.reset(&c->ws, out_cap * c->key_bytes + 64); }
                    PartitionedInstances pi;
                    c->timer.start();
                    const bool fast = exchange_partition_pull(c->comm, c->ws, c->key_bytes, c->keys.p, c->n_keys, c->hist.p, n_keys_all, plan.bits[0],
                                                              plan.levels > 1 ? plan.bits[1] : 0, c->alt.p, out_cap, &pi);
                    c->timer.stop(c->stats.ms_exchange);
                    if (fast) {
                        PartitionInput in;
                        in.keys = pi.recv; in.scratch = c->third.p; in.cstart = std::move(pi.cstart); in.n_parents = pi.n_parents; in.n_cap = pi.n_cap;
                        in.consumed_bits = pi.bits;
                        c->exchanged_instances = true;
                        flush_batch(c, true, &in, &plan);              // synchronises with the stream
                        if (exchange_partition_aborted(c->comm)) {
                            // some rank's window was too small for its share (a massively repeated k-mer): nothing was moved and
                            // nothing counted, on every rank alike; take the sampled exchange below instead
                            c->log(1, "partition exchange declined (a receive window would overflow): using the sampled exchange");
                            c->acc.keys.free(); c->acc.counts.free(); c->acc.m = 0; c->have_acc = false; c->acc_unsorted = false;
                            c->counts.n_instances = 0; c->counts.n_distinct = 0;
                            c->stats.n_batches -= 1; c->stats.n_sorted_keys = 0; c->stats.sort_passes_model -= c->passes;
                            c->n_keys = local_instances;
                            c->exchanged_instances = false;
                        } else {
                            c->stats.ms_all_to_all += exchange_partition_scatter_ms(c->comm);
                            c->stats.ms_sort_sweeps += exchange_partition_level0_ms(c->comm);   // the HBM-bound local pass
                            c->stats.sort_passes += 1;
                            c->stats.exchange_bytes_sent += exchange_partition_bytes_sent(c->comm);
                            c->stats.exchange_peer_memory = 1;
                            exchanged_instances = true;
                            counted_batch = true;
                        }
                    }
                }
                if (!exchanged_instances) {
                    c->timer.start();
                    u64 n_recv = 0;
                    ExchangeTiming et;
                    ensure_alt(c, c->n_keys);
                    u8* recv_ptr = nullptr;
                    exchange_instances(c->comm, c->ws, c->key_bytes, c->keys.p, c->n_keys, c->alt.p, c->third, &c->third_cap, &recv_ptr, &n_recv, &et, nullptr);
                    c->stats.ms_all_to_all += et.ms_all_to_all;
                    c->stats.exchange_bytes_sent += et.bytes_sent_remote;
                    c->stats.exchange_peer_memory = et.used_peer_memory ? 1 : 0;
                    c->batch_src = recv_ptr;                           // the received instances are the batch to count
                    c->n_keys = n_recv;
                    if (!c->mix) {                                     // legacy LSD counting wants the digit histograms of what arrived
                        GSB_CUDA_TRY(cudaMemsetAsync(c->hist.p, 0, c->hist.bytes(), c->ws.stream));
                        sort_digit_hist(c->key_bytes, recv_ptr, c->n_keys, c->passes, c->hist.p, c->ws.sm_count, c->ws.stream, &c->ws.launches);
                        c->hist_valid = true;
                    }
                    c->timer.stop(c->stats.ms_exchange);
                    exchanged_instances = true;
                }
            }
            c->exchanged_instances = exchanged_instances;
            if (!counted_batch) flush_batch(c, single);
            if (!c->have_acc) { c->acc.keys.reset(&c->ws, 0); c->acc.counts.reset(&c->ws, 0); c->acc.m = 0; c->have_acc = true; }
            if (c->comm && !exchanged_instances) {
                c->timer.start();
                exchange_runs(c->comm, c->ws, c->key_bytes, c->key_bits, c->acc);
                c->timer.stop(c->stats.ms_exchange);
            }
            const u64 min_count = c->cfg.kind == GSB_KIND_GRAPH ? std::max<u64>(1, c->cfg.min_count) : 1;
            const bool filtered_already = single && (!c->comm || exchanged_instances);
            bool stats_summed = false;                         // n_instances / n_distinct already summed over the ranks
            // after an instance exchange the keys of a rank are spread over the whole (real) key space -- they were routed by
            // their mixed value -- and a graph's reverse complements belong to other ranks anyway: one more re-partition
            const bool repartition = c->comm && (c->fold_w || exchanged_instances);
            const bool fast_p2p = exchanged_instances && exchange_peer_memory_usable(c->comm);   // sums ride on a later all-gather
            if (exchanged_instances && !fast_p2p) c->counts.n_distinct = exchange_sum(c->comm, c->ws, c->counts.n_distinct);
            if (!filtered_already) {
                // merged batches (and/or exchanged reduced runs): sorted, still folded, raw occurrence counts
                if (c->acc_unsorted) throw StatusError{GSB_EINVAL, "internal: unsorted run outside the single-batch path"};
                u64 local_distinct = c->acc.m;
                if (c->fold_w && c->acc.m && c->any_self_rc) {
                    c->timer.start();
                    const u64 n_self = fold_double_self_rc(c->ws, c->key_bytes, c->fold_w, c->acc.keys.p, c->acc.counts.p, c->acc.m);
                    c->timer.stop(c->stats.ms_unfold);
                    local_distinct = 2 * c->acc.m - n_self;
                } else if (c->fold_w) {
                    local_distinct = 2 * c->acc.m;
                }
                if (min_count > 1 && c->acc.m) {
                    c->timer.start();
                    DevBuf<u8> fk(&c->ws, c->acc.m * c->key_bytes);
                    DevBuf<u64> fc(&c->ws, c->acc.m), total(&c->ws, 1);
                    DevBuf<u8> lb(&c->ws, rle_lookback_bytes(c->acc.m));
                    sort_filter(c->key_bytes, c->acc.keys.p, c->acc.counts.p, nullptr, c->acc.m, min_count, fk.p, fc.p, lb.p, total.p, c->ws.stream, &c->ws.launches);
                    u64 kept = 0;
                    GSB_CUDA_TRY(cudaMemcpyAsync(&kept, total.p, 8, cudaMemcpyDeviceToHost, c->ws.stream));
                    c->ws.sync();
                    c->acc.keys = std::move(fk); c->acc.counts = std::move(fc); c->acc.m = kept;
                    c->timer.stop(c->stats.ms_reduce);
                }
                c->counts.n_distinct = c->comm ? exchange_sum(c->comm, c->ws, local_distinct) : local_distinct;
            }
            // acc: final counts, filtered, folded (graphs); sorted by key unless acc_unsorted
            bool done = false;
            // publishing the slices: the global view, the summed statistics, and the barrier that makes every slice sorted and
            // in place before anyone reads a neighbour's -- one all-gather
            auto publish = [&](const std::vector<u64>& totals) {
                c->timer.start();
                exchange_view(c->comm, c->key_bytes, totals, &c->dist);
                const u64 mine_stats[2] = {c->counts.n_instances, exchanged_instances ? c->counts.n_distinct : 0};
                std::vector<u64> all_stats;
                exchange_allgather_u64(c->comm, c->ws, mine_stats, 2, all_stats);
                u64 inst = 0, dist = 0;
                for (int r = 0; r < exchange_size(c->comm); ++r) { inst += all_stats[2 * r]; dist += all_stats[2 * r + 1]; }
                c->counts.n_instances = inst;
                if (exchanged_instances) c->counts.n_distinct = dist;
                stats_summed = true;
                c->dist_ready = true;
                c->timer.stop(c->stats.ms_exchange_publish);
            };
            if (repartition && exchange_peer_memory_usable(c->comm) && !g_sampled_survivors) {
                // The survivors (folded for graphs: the reverse complements join here) go to the ranks that own their range
                // of the FINAL order as the first pass of the pair sort, stored straight into the owners' windows; each
                // owner finishes the sort of its slice locally (exchange.cu).
                c->timer.start();
                ReducedRun sorted;
                std::vector<u64> totals;
                const bool ok = exchange_pairs_msd(c->comm, c->ws, c->key_bytes, c->key_bits, c->acc.keys.p, c->acc.counts.p, c->acc.m, c->fold_w, sorted, &totals);
                c->timer.stop(c->stats.ms_exchange_survivors);
                if (ok) {
                    c->acc = std::move(sorted);
                    publish(totals);
                    done = true;
                }
            }
            if (!done && repartition && exchange_peer_memory_usable(c->comm)) {
                // U = acc (++ rc(acc) for graphs) is built unsorted, every pair is stored straight into the window of the rank
                // that owns its range of the FINAL order (splitters sampled from U itself, so the slices are balanced), and
                // the owner sorts what it received -- the slice is then already published for the emitters.
                const u64 m = c->acc.m;
                const int kb = c->key_bytes;
                c->timer.start();
                const u64 u_cap = c->fold_w ? 2 * m : m;
                DevBuf<u8> uk(&c->ws, u_cap * kb);
                DevBuf<u64> uc(&c->ws, u_cap);
                if (m) {
                    GSB_CUDA_TRY(cudaMemcpyAsync(uk.p, c->acc.keys.p, m * kb, cudaMemcpyDeviceToDevice, c->ws.stream));
                    GSB_CUDA_TRY(cudaMemcpyAsync(uc.p, c->acc.counts.p, m * 8, cudaMemcpyDeviceToDevice, c->ws.stream));
                }
                const u64 n_rc = c->fold_w ? unfold_append_rc(c->ws, kb, c->fold_w, c->acc.keys.p, c->acc.counts.p, m, uk.p + m * kb, uc.p + m) : 0;
                c->timer.stop(c->stats.ms_unfold);
                c->timer.start();
                u8* rk = nullptr; u64* rc = nullptr;
                std::vector<u64> totals;
                const bool ok = exchange_pairs_p2p(c->comm, c->ws, kb, uk.p, uc.p, m + n_rc, &rk, &rc, &totals);
                c->timer.stop(c->stats.ms_exchange_survivors);
                if (ok) {
                    uk.free(); uc.free();
                    const u64 mine = totals[exchange_rank(c->comm)];
                    c->timer.start();
                    const cudaMemcpyKind d2d = cudaMemcpyDeviceToDevice;
                    ReducedRun sorted;
                    if (mine && sort_pairs_msd(c->ws, kb, c->key_bits, rk, rc, mine, 0, sorted)) {
                        // one copy stays in the window (for the peers), one is acc
                        GSB_CUDA_TRY(cudaMemcpyAsync(rk, sorted.keys.p, mine * kb, d2d, c->ws.stream));
                        GSB_CUDA_TRY(cudaMemcpyAsync(rc, sorted.counts.p, mine * 8, d2d, c->ws.stream));
                        c->acc = std::move(sorted);
                    } else {
                        DevBuf<u8> bk(&c->ws, mine * kb);
                        DevBuf<u64> bc(&c->ws, mine);
                        const int where = sort_keys(c->ws, kb, c->key_bits, rk, bk.p, rc, bc.p, mine, nullptr, nullptr);
                        if (mine) {
                            GSB_CUDA_TRY(cudaMemcpyAsync(where ? (void*)rk : (void*)bk.p, where ? (void*)bk.p : (void*)rk, mine * kb, d2d, c->ws.stream));
                            GSB_CUDA_TRY(cudaMemcpyAsync(where ? rc : bc.p, where ? bc.p : rc, mine * 8, d2d, c->ws.stream));
                        }
                        c->acc.keys = std::move(bk); c->acc.counts = std::move(bc); c->acc.m = mine;
                    }
                    c->timer.stop(c->stats.ms_unfold);
                    publish(totals);
                    done = true;
                }
            }
            if (!done) {
                c->timer.start();
                if (c->fold_w) unfold_run(c->ws, c->key_bytes, c->key_bits, c->fold_w, c->acc, !c->acc_unsorted);   // acc := sorted(acc U rc(acc))
                else if (c->acc_unsorted) sort_run(c, c->acc);
                c->timer.stop(c->stats.ms_unfold);
                if (repartition) {
                    c->timer.start();
                    exchange_runs(c->comm, c->ws, c->key_bytes, c->key_bits, c->acc);
                    c->timer.stop(c->stats.ms_exchange);
                }
            }
            c->acc_unsorted = false;
            if (c->comm) {
                // publish the slice in this rank's peer-mapped window: the distributed emitters work from there
                if (!c->dist_ready) {
                    c->timer.start();
                    c->dist_ready = exchange_publish(c->comm, c->ws, c->key_bytes, c->acc, &c->dist);
                    c->timer.stop(c->stats.ms_exchange);
                }
                c->counts.n_kept = c->dist_ready ? c->dist.off[c->dist.n] : exchange_sum(c->comm, c->ws, c->acc.m);
                if (!stats_summed) {
                    c->counts.n_instances = exchange_sum(c->comm, c->ws, c->counts.n_instances);
                    if (fast_p2p) c->counts.n_distinct = exchange_sum(c->comm, c->ws, c->counts.n_distinct);   // the p2p path declined after all
                }
            } else {
                c->counts.n_kept = c->acc.m;
            }
            c->counted = true;
        }
        if (out) *out = c->counts;
    });
}

int gsb_emit(gsb_ctx* c, const char* prefix, const gsb_sink* sink) {
    if (!c || !prefix || (sink && (!sink->open || !sink->pwrite || !sink->close))) return GSB_EINVAL;
    return guarded(c, [&] {
        if (!c->counted) throw StatusError{GSB_EINVAL, "gsb_emit before gsb_finish_counting"};
        Emitter em;
        em.ws = &c->ws; em.sink = sink; em.pinned = c->pinned; em.pinned_bytes = c->pinned_bytes;
        RingGuard ring_guard{sink ? &c->ring : nullptr};           // whatever happens below, the writer thread is done with `sink` on return
        if (sink) { c->ring.create(c->ws); c->ring.drop(); c->ring.sink = sink; em.ring = &c->ring; }
        if (c->comm && !c->gathered && !c->dist_ready) {
            // peer memory is not available here: fall back to shipping every slice to rank 0
            c->timer.start();
            exchange_gather(c->comm, c->ws, c->key_bytes, c->acc);
            c->timer.stop(c->stats.ms_exchange);
            c->gathered = true;
        }
        c->timer.start();
        if (c->comm && !c->gathered) {
            if (c->cfg.kind == GSB_KIND_GRAPH) write_graph_files_dist(c, em, prefix);
            else write_kmer_set_files_dist(c, em, prefix);
            exchange_barrier(c->comm, c->ws);                    // nobody reuses its window while a peer still reads it
        } else if (!c->comm || exchange_rank(c->comm) == 0) {
            if (c->cfg.kind == GSB_KIND_GRAPH) write_graph_files(c, em, prefix);
            else write_kmer_set_files(c, em, prefix);
        }
        c->timer.stop(c->stats.ms_emit);
        em.flush();                                              // the last chunks cross PCIe and reach the sink
        c->stats.bytes_out += em.bytes_out;
    });
}

// ---- existing file sets ---------------------------------------------------------------------------------------------
static void read_host_file(const gsb_source* src, const std::string& name, std::vector<u8>& out) {
    uint64_t size = 0;
    if (src->size(src->user, name.c_str(), &size) != 0) throw StatusError{GSB_EIO, "cannot open " + name};
    out.resize(size);
    if (size && src->pread(src->user, name.c_str(), 0, out.data(), size) != 0) throw StatusError{GSB_EIO, "read failed for " + name};
}

static void peek_graph(const std::string& prefix, const gsb_source* src, int kind, gsb_graph_info* out) {
    std::vector<u8> h, sa;
    read_host_file(src, prefix + ".header", h);
    if (h.size() < 24) throw StatusError{GSB_EIO, prefix + ".header is truncated"};
    u64 w[3];
    memcpy(w, h.data(), 24);
    const u64 want = kind == GSB_KIND_GRAPH ? 2011101014ull : 2011101701ull;
    if (w[0] != want) throw StatusError{GSB_EINVAL, prefix + ": version mismatch " + std::to_string(w[0]) + " vs " + std::to_string(want)};
    read_host_file(src, prefix + (kind == GSB_KIND_GRAPH ? "-edges.header" : ".kmers.header"), sa);
    if (sa.size() < 64) throw StatusError{GSB_EIO, prefix + ": SparseArray header is truncated"};
    u64 count = 0;
    memcpy(&count, sa.data() + 56, 8);
    out->version = w[0]; out->k = w[1]; out->flags = w[2]; out->n_items = count;
}

int gsb_graph_peek(const char* prefix, const gsb_source* src, int kind, gsb_graph_info* out, char* err, size_t errcap) {
    if (!prefix || !src || !src->size || !src->pread || !out) return GSB_EINVAL;
    int rc = guarded(nullptr, [&] { peek_graph(prefix, src, kind, out); });
    if (rc != GSB_OK && err && errcap) { strncpy(err, g_create_error.c_str(), errcap - 1); err[errcap - 1] = 0; }
    return rc;
}

static void absorb_run(gsb_ctx* c, ReducedRun& run) {
    if (c->have_acc) {
        c->timer.start();
        merge_runs(c, c->acc, run);
        c->timer.stop(c->stats.ms_merge);
    } else {
        c->acc = std::move(run);
        c->have_acc = true;
    }
}

int gsb_graph_load(gsb_ctx* c, const char* prefix, const gsb_source* src) {
    if (!c || !prefix || !src || !src->size || !src->pread) return GSB_EINVAL;
    return guarded(c, [&] {
        if (c->counted) throw StatusError{GSB_EINVAL, "gsb_graph_load after the run was finished (call gsb_reset first)"};
        if (c->comm) throw StatusError{GSB_EINVAL, "gsb_graph_load is a single-GPU operation"};
        if (c->n_keys) throw StatusError{GSB_EINVAL, "gsb_graph_load cannot be mixed with gsb_push_block in one build"};
        gsb_graph_info info;
        peek_graph(prefix, src, c->cfg.kind, &info);
        if ((int)info.k != c->cfg.k) throw StatusError{GSB_EINVAL, std::string(prefix) + " has k=" + std::to_string(info.k) + ", this context was created for k=" + std::to_string(c->cfg.k)};
        if (c->cfg.kind == GSB_KIND_GRAPH && (info.flags & 1)) throw StatusError{GSB_EINVAL, "Asymmetric graphs not yet handled"};
        ReducedRun run;
        c->timer.start();
        const std::string p(prefix);
        read_sparse_array(c->ws, src, p + (c->cfg.kind == GSB_KIND_GRAPH ? "-edges" : ".kmers"), c->key_bytes, c->pinned, c->pinned_bytes, run.keys, &run.m);
        if (c->cfg.kind == GSB_KIND_GRAPH) read_counts(c->ws, src, p + "-counts", run.m, c->pinned, c->pinned_bytes, run.counts);
        else { run.counts.reset(&c->ws, run.m); fill_ones(c->ws, run.counts.p, run.m); }
        c->timer.stop(c->stats.ms_scan);
        c->loaded_items += run.m;
        absorb_run(c, run);
    });
}

int gsb_graph_load_pairs(gsb_ctx* c, const uint64_t* key_lo, const uint64_t* key_hi, const uint64_t* counts, uint64_t m) {
    if (!c || (m && (!key_lo || !counts))) return GSB_EINVAL;
    return guarded(c, [&] {
        if (c->counted) throw StatusError{GSB_EINVAL, "gsb_graph_load_pairs after the run was finished (call gsb_reset first)"};
        ReducedRun raw, run;
        Workspace& ws = c->ws;
        const int kb = c->key_bytes;
        raw.keys.reset(&ws, m * kb);
        raw.counts.reset(&ws, m);
        if (m) {
            if (kb == 8) {
                GSB_CUDA_TRY(cudaMemcpyAsync(raw.keys.p, key_lo, m * 8, cudaMemcpyHostToDevice, ws.stream));
            } else {
                std::vector<u64> inter(2 * m);
                for (u64 i = 0; i < m; ++i) { inter[2 * i] = key_lo[i]; inter[2 * i + 1] = key_hi ? key_hi[i] : 0; }
                GSB_CUDA_TRY(cudaMemcpyAsync(raw.keys.p, inter.data(), m * 16, cudaMemcpyHostToDevice, ws.stream));
                ws.sync();
            }
            GSB_CUDA_TRY(cudaMemcpyAsync(raw.counts.p, counts, m * 8, cudaMemcpyHostToDevice, ws.stream));
            ws.sync();
        }
        raw.m = m;
        // any order, possibly repeated keys: sort by key, sum equal keys
        DevBuf<u8> kalt(&ws, m * kb);
        DevBuf<u64> calt(&ws, m);
        const int where = sort_keys(ws, kb, c->key_bits, raw.keys.p, kalt.p, raw.counts.p, calt.p, m, nullptr, nullptr);
        u64 distinct = 0;
        reduce_sorted(ws, kb, where ? kalt.p : raw.keys.p, where ? calt.p : raw.counts.p, m, 1, run, &distinct);
        c->loaded_items += m;
        absorb_run(c, run);
    });
}

int gsb_graph_finish(gsb_ctx* c, uint64_t cutoff, uint64_t m_est, gsb_counts* out) {
    if (!c) return GSB_EINVAL;
    return guarded(c, [&] {
        if (c->counted) throw StatusError{GSB_EINVAL, "the run is finished already"};
        if (c->n_keys) throw StatusError{GSB_EINVAL, "gsb_graph_finish cannot be mixed with gsb_push_block in one build"};
        if (!c->have_acc) { c->acc.keys.reset(&c->ws, 0); c->acc.counts.reset(&c->ws, 0); c->acc.m = 0; c->have_acc = true; }
        c->counts.n_distinct = c->acc.m;
        if (cutoff > 0 && c->acc.m) {
            c->timer.start();
            DevBuf<u8> fk(&c->ws, c->acc.m * c->key_bytes);
            DevBuf<u64> fc(&c->ws, c->acc.m), total(&c->ws, 1);
            DevBuf<u8> lb(&c->ws, rle_lookback_bytes(c->acc.m));
            sort_filter(c->key_bytes, c->acc.keys.p, c->acc.counts.p, nullptr, c->acc.m, cutoff + 1, fk.p, fc.p, lb.p, total.p, c->ws.stream, &c->ws.launches);
            u64 kept = 0;
            GSB_CUDA_TRY(cudaMemcpyAsync(&kept, total.p, 8, cudaMemcpyDeviceToHost, c->ws.stream));
            c->ws.sync();
            c->acc.keys = std::move(fk); c->acc.counts = std::move(fc); c->acc.m = kept;
            c->timer.stop(c->stats.ms_reduce);
        }
        c->counts.n_kept = c->acc.m;
        c->counts.n_instances = c->loaded_items;
        c->m_est = m_est;
        c->acc_unsorted = false;
        c->counted = true;
        if (out) *out = c->counts;
    });
}

// ---- xenome index, steps 3 and 4 (xeno.cu) --------------------------------------------------------------------------------
static void load_kmer_set_weighted(gsb_ctx* c, const std::string& prefix, const gsb_source* src, u64 weight, u64* m_out) {
    gsb_graph_info info;
    peek_graph(prefix, src, GSB_KIND_KMERSET, &info);
    if ((int)info.k != c->cfg.k) throw StatusError{GSB_EINVAL, prefix + " has k=" + std::to_string(info.k) + ", this context was created for k=" + std::to_string(c->cfg.k)};
    ReducedRun run;
    read_sparse_array(c->ws, src, prefix + ".kmers", c->key_bytes, c->pinned, c->pinned_bytes, run.keys, &run.m);
    run.counts.reset(&c->ws, run.m);
    fill_value(c->ws, run.counts.p, run.m, weight);           // the summed weight of an element of the union is its membership
    *m_out = run.m;
    absorb_run(c, run);
}

int gsb_kmerset_merge_annotate(gsb_ctx* c, const char* lhs_prefix, const char* rhs_prefix, const gsb_source* src, const char* out_prefix,
                               const gsb_sink* sink, uint64_t* stats) {
    if (!c || !lhs_prefix || !rhs_prefix || !out_prefix || !src || !src->size || !src->pread) return GSB_EINVAL;
    if (sink && (!sink->open || !sink->pwrite || !sink->close)) return GSB_EINVAL;
    return guarded(c, [&] {
        if (c->cfg.kind != GSB_KIND_KMERSET) throw StatusError{GSB_EINVAL, "merge-and-annotate-kmer-sets wants a kmer-set context"};
        if (c->comm) throw StatusError{GSB_EINVAL, "merge-and-annotate-kmer-sets is a single-GPU operation"};
        if (c->counted || c->have_acc || c->n_keys) throw StatusError{GSB_EINVAL, "the context is in use (call gsb_reset first)"};
        u64 n_lhs = 0, n_rhs = 0;
        load_kmer_set_weighted(c, lhs_prefix, src, 1, &n_lhs);      // (absorb_run times the merge itself: the phase timer does not nest)
        load_kmer_set_weighted(c, rhs_prefix, src, 2, &n_rhs);
        // src/GossCmdMergeAndAnnotateKmerSets.cc:41-49: a bare `throw "nonsense"` for an empty side (k is checked above)
        if (n_lhs == 0 || n_rhs == 0) throw StatusError{GSB_EINVAL, "nonsense"};
        const u64 n = c->acc.m;
        c->counts.n_instances = n_lhs + n_rhs; c->counts.n_distinct = n; c->counts.n_kept = n;
        c->m_est = n;                                            // KmerSet::Builder bld(K, out, fac, n), :121
        c->counted = true;
        Emitter em;
        em.ws = &c->ws; em.sink = sink; em.pinned = c->pinned; em.pinned_bytes = c->pinned_bytes;
        RingGuard ring_guard{sink ? &c->ring : nullptr};           // whatever happens below, the writer thread is done with `sink` on return
        if (sink) { c->ring.create(c->ws); c->ring.drop(); c->ring.sink = sink; em.ring = &c->ring; }
        c->timer.start();
        const std::string out(out_prefix);
        write_kmer_set_files(c, em, out);
        DevBuf<u64> lhs, rhs;
        xeno_annotate_bits(c->ws, c->acc.counts.p, n, lhs, rhs);
        em.put_device(out + ".lhs-bits", lhs.p, bit_vector_words(n) * 8);
        em.put_device(out + ".rhs-bits", rhs.p, bit_vector_words(n) * 8);
        c->timer.stop(c->stats.ms_emit);
        em.flush();
        c->stats.bytes_out += em.bytes_out;
        if (stats) { stats[0] = n_lhs; stats[1] = n_rhs; stats[2] = n_lhs + n_rhs - n; stats[3] = n; }
    });
}

int gsb_kmerset_near_kmers(gsb_ctx* c, const char* prefix, const gsb_source* src, const gsb_sink* sink, uint64_t* n_gray) {
    if (!c || !prefix || !src || !src->size || !src->pread) return GSB_EINVAL;
    if (sink && (!sink->open || !sink->pwrite || !sink->close)) return GSB_EINVAL;
    return guarded(c, [&] {
        if (c->cfg.kind != GSB_KIND_KMERSET) throw StatusError{GSB_EINVAL, "compute-near-kmers wants a kmer-set context"};
        if (c->comm) throw StatusError{GSB_EINVAL, "compute-near-kmers is a single-GPU operation"};
        const std::string p(prefix);
        gsb_graph_info info;
        peek_graph(p, src, GSB_KIND_KMERSET, &info);
        if ((int)info.k != c->cfg.k) throw StatusError{GSB_EINVAL, p + " has k=" + std::to_string(info.k) + ", this context was created for k=" + std::to_string(c->cfg.k)};
        Workspace& ws = c->ws;
        DevBuf<u8> keys; u64 m = 0;
        c->timer.start();
        read_sparse_array(ws, src, p + ".kmers", c->key_bytes, c->pinned, c->pinned_bytes, keys, &m);
        const u64 words = bit_vector_words(m);
        DevBuf<u64> bits[2];
        const char* suffix[2] = {".lhs-bits", ".rhs-bits"};
        for (int side = 0; side < 2; ++side) {
            const std::string name = p + suffix[side];
            uint64_t size = 0;
            if (src->size(src->user, name.c_str(), &size) != 0) throw StatusError{GSB_EIO, "missing file " + name};
            if (size < words * 8) throw StatusError{GSB_EINVAL, name + " is shorter than the kmer set it annotates"};
            bits[side].reset(&ws, words);
            for (u64 off = 0; off < words * 8; off += c->pinned_bytes) {
                const u64 chunk = std::min<u64>(c->pinned_bytes, words * 8 - off);
                if (src->pread(src->user, name.c_str(), off, c->pinned, chunk) != 0) throw StatusError{GSB_EIO, "read failed for " + name};
                GSB_CUDA_TRY(cudaMemcpyAsync((u8*)bits[side].p + off, c->pinned, chunk, cudaMemcpyHostToDevice, ws.stream));
                ws.sync();
            }
        }
        c->timer.stop(c->stats.ms_scan);
        c->timer.start();
        DevBuf<u64> nl, nr;
        const u64 gray = xeno_near_kmers(ws, c->key_bytes, c->cfg.k, keys.p, m, bits[0].p, bits[1].p, nl, nr);
        c->timer.stop(c->stats.ms_reduce);
        Emitter em;
        em.ws = &ws; em.sink = sink; em.pinned = c->pinned; em.pinned_bytes = c->pinned_bytes;
        RingGuard ring_guard{sink ? &c->ring : nullptr};
        if (sink) { c->ring.create(ws); c->ring.drop(); c->ring.sink = sink; em.ring = &c->ring; }
        em.put_device(p + ".lhs-bits", nl.p, words * 8);
        em.put_device(p + ".rhs-bits", nr.p, words * 8);
        em.flush();
        c->stats.bytes_out += em.bytes_out;
        if (n_gray) *n_gray = gray;
    });
}

int gsb_graph_dump(gsb_ctx* c, const char* name, const gsb_sink* sink) {
    if (!c || !name || !sink || !sink->open || !sink->pwrite || !sink->close) return GSB_EINVAL;
    return guarded(c, [&] {
        if (!c->counted) throw StatusError{GSB_EINVAL, "gsb_graph_dump before the run was finished"};
        if (c->cfg.kind != GSB_KIND_GRAPH) throw StatusError{GSB_EINVAL, "dump-graph wants a graph"};
        Emitter em;
        em.ws = &c->ws; em.sink = sink; em.pinned = c->pinned; em.pinned_bytes = c->pinned_bytes;
        DevBuf<u8> text;
        u64 bytes = 0;
        dump_text(c->ws, c->key_bytes, c->acc.keys.p, c->acc.counts.p, c->acc.m, c->cfg.k + 1, text, &bytes);
        // '#' version, then K <tab> count <tab> flags (src/GossCmdDumpGraph.cc:49-50)
        const std::string head = "#2011101014\n" + std::to_string(c->cfg.k) + "\t" + std::to_string(c->acc.m) + "\t0\n";
        void* h = nullptr;
        const std::string nm(name);
        if (sink->open(sink->user, name, head.size() + bytes, &h) != 0) throw StatusError{GSB_EIO, "open failed for " + nm};
        if (sink->pwrite(sink->user, h, 0, head.data(), head.size()) != 0) throw StatusError{GSB_EIO, "pwrite failed for " + nm};
        for (u64 off = 0; off < bytes; off += c->pinned_bytes) {
            const u64 chunk = std::min<u64>(c->pinned_bytes, bytes - off);
            GSB_CUDA_TRY(cudaMemcpyAsync(c->pinned, text.p + off, chunk, cudaMemcpyDeviceToHost, c->ws.stream));
            c->ws.sync();
            if (sink->pwrite(sink->user, h, head.size() + off, c->pinned, chunk) != 0) throw StatusError{GSB_EIO, "pwrite failed for " + nm};
        }
        if (sink->close(sink->user, h) != 0) throw StatusError{GSB_EIO, "close failed for " + nm};
        c->stats.bytes_out += head.size() + bytes;
    });
}

int gsb_timer_begin(gsb_ctx* c) {
    if (!c) return GSB_EINVAL;
    return guarded(c, [&] { GSB_CUDA_TRY(cudaEventRecord(c->user_e0, c->ws.stream)); });
}

int gsb_timer_end(gsb_ctx* c, double* ms_out) {
    if (!c || !ms_out) return GSB_EINVAL;
    return guarded(c, [&] {
        GSB_CUDA_TRY(cudaEventRecord(c->user_e1, c->ws.stream));
        GSB_CUDA_TRY(cudaEventSynchronize(c->user_e1));
        float ms = 0;
        GSB_CUDA_TRY(cudaEventElapsedTime(&ms, c->user_e0, c->user_e1));
        *ms_out = ms;
    });
}

// Binds the calling thread to the CPUs of the NUMA node the device hangs off (sysfs: the device's PCI address ->
// numa_node -> that node's cpulist), so that the pinned block buffers allocated afterwards -- first touched by this
// thread -- are local to the GPU's root port.  With eight ranks copying 0.5 GB blocks at once the host side is what
// limits the end-to-end rate.  Best effort: GSB_OK also when the topology cannot be read (nothing is changed then).
int gsb_host_bind_near_device(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return GSB_OK; }
    for (char* p = bus; *p; ++p) *p = (char)tolower(*p);
    auto read_line = [](const std::string& path, std::string& out) {
        FILE* f = fopen(path.c_str(), "r");
        if (!f) return false;
        char buf[4096];
        const bool ok = fgets(buf, sizeof(buf), f) != nullptr;
        fclose(f);
        if (ok) { out = buf; while (!out.empty() && (out.back() == '\n' || out.back() == ' ')) out.pop_back(); }
        return ok;
    };
    std::string node, cpus;
    if (!read_line(std::string("/sys/bus/pci/devices/") + bus + "/numa_node", node)) return GSB_OK;
    const int nid = atoi(node.c_str());
    if (nid < 0) return GSB_OK;                                   // single-node box (or the firmware does not say)
    if (!read_line("/sys/devices/system/node/node" + std::to_string(nid) + "/cpulist", cpus) || cpus.empty()) return GSB_OK;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n_set = 0;
    size_t i = 0;
    while (i < cpus.size()) {                                     // "0-15,64-79"
        char* end = nullptr;
        const long a = strtol(cpus.c_str() + i, &end, 10);
        long b = a;
        i = (size_t)(end - cpus.c_str());
        if (i < cpus.size() && cpus[i] == '-') { b = strtol(cpus.c_str() + i + 1, &end, 10); i = (size_t)(end - cpus.c_str()); }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, &set); ++n_set; }
        if (i < cpus.size() && cpus[i] == ',') ++i; else if (i < cpus.size() && !isdigit((unsigned char)cpus[i])) break;
    }
    if (n_set) sched_setaffinity(0, sizeof(set), &set);
    return GSB_OK;
}

int gsb_host_alloc(size_t nbytes, void** out) {
    if (!out) return GSB_EINVAL;
    *out = nullptr;
    return guarded(nullptr, [&] {
        cudaError_t e = cudaMallocHost(out, nbytes ? nbytes : 1);
        if (e != cudaSuccess) { cudaGetLastError(); throw StatusError{e == cudaErrorMemoryAllocation ? GSB_ENOMEM : GSB_ECUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e)}; }
    });
}

void gsb_host_free(void* p) { if (p) cudaFreeHost(p); }

int gsb_get_stats(const gsb_ctx* cc, gsb_stats* out) {
    if (!cc || !out) return GSB_EINVAL;
    gsb_ctx* c = const_cast<gsb_ctx*>(cc);
    cudaSetDevice(c->ws.device);
    c->timer.resolve();
    *out = c->stats;
    out->kernel_launches = c->ws.launches;
    out->hbm_peak_bytes = c->ws.peak_bytes;
    out->device_allocs = c->ws.device_allocs;
    return GSB_OK;
}

int gsb_reset(gsb_ctx* c) {
    if (!c) return GSB_EINVAL;
    return guarded(c, [&] {
        c->acc.keys.free(); c->acc.counts.free(); c->acc.m = 0;
        c->have_acc = false; c->counted = false;
        c->dist_ready = false; c->gathered = false;
        c->self_rc_windows = 0; c->any_self_rc = true;
        c->m_est = 0; c->loaded_items = 0;
        c->exchanged_instances = false; c->batch_src = nullptr; c->acc_unsorted = false;
        if (c->pending.valid) { GSB_CUDA_TRY(cudaStreamSynchronize(c->copy_stream)); c->pending.valid = false; }
        for (int f = 0; f < 3; ++f) { c->file_open[f] = false; c->line_base[f] = 0; }
        c->n_carry = 0;
        memset(&c->counts, 0, sizeof(c->counts));
        const u64 launches = c->ws.launches;
        c->timer.resolve();                                    // accumulators of the step that ends here
        memset(&c->stats, 0, sizeof(c->stats));
        c->stats.sort_key_bytes = c->key_bytes;
        c->ws.launches = launches;
        reset_batch(c);
    });
}

int gsb_comm_make_id(void* id_out) {
    if (!id_out) return GSB_EINVAL;
    return guarded(nullptr, [&] { exchange_make_id(id_out); });
}

int gsb_comm_attach(gsb_ctx* c, const void* id, int n_ranks, int rank) {
    if (!c || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return GSB_EINVAL;
    return guarded(c, [&] {
        if (c->comm) throw StatusError{GSB_EINVAL, "communicator already attached"};
        if (n_ranks > kMaxRanks) throw StatusError{GSB_EINVAL, "at most " + std::to_string(kMaxRanks) + " ranks are supported"};
        if (c->n_keys || c->have_acc) throw StatusError{GSB_EINVAL, "gsb_comm_attach after input has been pushed (attach first, or gsb_reset)"};
        c->comm = exchange_create(id, n_ranks, rank, c->ws);
        c->hist_valid = c->mix;
    });
}

int gsb_plan_splitters(const uint64_t* samples, uint64_t n_samples, int n_ranks, uint64_t* splitters_out) {
    if (!samples || !splitters_out || n_ranks < 1 || n_samples == 0) return GSB_EINVAL;
    return guarded(nullptr, [&] { plan_splitters((const u64*)samples, n_samples, n_ranks, (u64*)splitters_out); });
}

uint32_t gsb_samples_per_rank(void) { return kExchangeSamplesPerRank; }

int gsb_gather_to_root(gsb_ctx* c) {
    if (!c) return GSB_EINVAL;
    return guarded(c, [&] {
        if (!c->counted) throw StatusError{GSB_EINVAL, "gsb_gather_to_root before gsb_finish_counting"};
        if (!c->comm) return;
        if (c->gathered) return;
        c->timer.start();
        exchange_gather(c->comm, c->ws, c->key_bytes, c->acc);
        c->timer.stop(c->stats.ms_exchange);
        c->gathered = true;
    });
}

int64_t gsb_debug_copy_counts(gsb_ctx* c, uint64_t* key_lo, uint64_t* key_hi, uint64_t* counts, uint64_t cap) {
    if (!c) return GSB_EINVAL;
    int64_t result = 0;
    int rc = guarded(c, [&] {
        if (!c->counted) throw StatusError{GSB_EINVAL, "gsb_debug_copy_counts before gsb_finish_counting"};
        const u64 m = c->acc.m, take = std::min<u64>(m, cap);
        std::vector<u64> raw(take * (c->key_bytes / 8));
        GSB_CUDA_TRY(cudaMemcpyAsync(raw.data(), c->acc.keys.p, take * c->key_bytes, cudaMemcpyDeviceToHost, c->ws.stream));
        if (counts) GSB_CUDA_TRY(cudaMemcpyAsync(counts, c->acc.counts.p, take * 8, cudaMemcpyDeviceToHost, c->ws.stream));
        c->ws.sync();
        for (u64 i = 0; i < take; ++i) {
            if (c->key_bytes == 8) { if (key_lo) key_lo[i] = raw[i]; if (key_hi) key_hi[i] = 0; }
            else { if (key_lo) key_lo[i] = raw[2 * i]; if (key_hi) key_hi[i] = raw[2 * i + 1]; }
        }
        result = (int64_t)m;
    });
    return rc == GSB_OK ? result : rc;
}

// ---- test-only entry points -----------------------------------------------------------------

int64_t gsb_debug_extract(int device, const void* text, size_t nbytes, int format, int kind, int k,
                          uint64_t* key_lo, uint64_t* key_hi, uint64_t cap, uint64_t* n_reads,
                          char* err, size_t errcap) {
    gsb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = GSB_ABI_VERSION; cfg.kind = kind; cfg.k = k; cfg.device = device;
    gsb_ctx* c = nullptr;
    int rc = gsb_create(&cfg, &c);
    auto report = [&](const char* m) { if (err && errcap) { strncpy(err, m, errcap - 1); err[errcap - 1] = 0; } };
    if (rc != GSB_OK) { report(gsb_last_error(nullptr)); return rc; }
    rc = gsb_push_block(c, text, nbytes, format, GSB_BLOCK_LAST_OF_FILE);
    if (rc != GSB_OK) { report(gsb_last_error(c)); gsb_destroy(c); return rc; }
    int64_t n = (int64_t)c->n_keys;
    rc = guarded(c, [&] {
        const u64 take = std::min<u64>(c->n_keys, cap);
        std::vector<u64> raw(take * (c->key_bytes / 8));
        GSB_CUDA_TRY(cudaMemcpyAsync(raw.data(), c->keys.p, take * c->key_bytes, cudaMemcpyDeviceToHost, c->ws.stream));
        c->ws.sync();
        for (u64 i = 0; i < take; ++i) {                           // the instances are stored bit-mixed: hand back the real keys
            if (c->key_bytes == 8) {
                const u64 k = c->mix ? key_unmix(raw[i]) : raw[i];
                if (key_lo) key_lo[i] = k; if (key_hi) key_hi[i] = 0;
            } else {
                Key128 k; k.lo = raw[2 * i]; k.hi = raw[2 * i + 1];
                if (c->mix) k = key_unmix(k);
                if (key_lo) key_lo[i] = k.lo; if (key_hi) key_hi[i] = k.hi;
            }
        }
        if (n_reads) *n_reads = c->counts.n_reads;
    });
    if (rc != GSB_OK) { report(gsb_last_error(c)); n = rc; }
    gsb_destroy(c);
    return n;
}

namespace {
struct DebugDevice {
    Workspace ws;
    u8* pinned = nullptr;
    void open(int device) {
        int n_dev = 0;
        if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); throw StatusError{GSB_ECUDA, "no CUDA device available (this library has no CPU path)"}; }
        ws.device = device;
        GSB_CUDA_TRY(cudaSetDevice(device));
        cudaDeviceProp prop;
        GSB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        ws.sm_count = prop.multiProcessorCount;
        GSB_CUDA_TRY(cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking));
        GSB_CUDA_TRY(cudaMallocHost((void**)&pinned, 16 << 20));
    }
    ~DebugDevice() {
        if (ws.stream) { cudaStreamSynchronize(ws.stream); ws.trim(); cudaStreamDestroy(ws.stream); ws.stream = nullptr; }
        if (pinned) cudaFreeHost(pinned);
    }
};

// host (lo, hi) arrays -> device key array
void upload_keys(Workspace& ws, int key_bytes, const uint64_t* lo, const uint64_t* hi, u64 n, DevBuf<u8>& out) {
    out.reset(&ws, n * key_bytes);
    if (!n) return;
    if (key_bytes == 8) {
        GSB_CUDA_TRY(cudaMemcpyAsync(out.p, lo, n * 8, cudaMemcpyHostToDevice, ws.stream));
    } else {
        std::vector<u64> inter(2 * n);
        for (u64 i = 0; i < n; ++i) { inter[2 * i] = lo[i]; inter[2 * i + 1] = hi ? hi[i] : 0; }
        GSB_CUDA_TRY(cudaMemcpyAsync(out.p, inter.data(), n * 16, cudaMemcpyHostToDevice, ws.stream));
        ws.sync();
    }
    ws.sync();
}
}  // namespace

int64_t gsb_debug_sort_keys(int device, uint64_t* key_lo, uint64_t* key_hi, uint64_t n, int key_bits) {
    int64_t passes = 0;
    int rc = guarded(nullptr, [&] {
        DebugDevice d; d.open(device);
        const int kb = key_bits <= 64 ? 8 : 16;
        DevBuf<u8> a, b(&d.ws, n * kb);
        upload_keys(d.ws, kb, key_lo, key_hi, n, a);
        int run = 0;
        int where = sort_keys(d.ws, kb, key_bits, a.p, b.p, nullptr, nullptr, n, nullptr, &run);
        std::vector<u64> raw(n * (kb / 8));
        if (n) GSB_CUDA_TRY(cudaMemcpyAsync(raw.data(), where ? b.p : a.p, n * kb, cudaMemcpyDeviceToHost, d.ws.stream));
        d.ws.sync();
        for (u64 i = 0; i < n; ++i) {
            if (kb == 8) key_lo[i] = raw[i];
            else { key_lo[i] = raw[2 * i]; key_hi[i] = raw[2 * i + 1]; }
        }
        passes = run;
        a.free(); b.free();
    });
    return rc == GSB_OK ? passes : rc;
}

// sorts `iters` fresh random key sets of n keys on the device; returns the mean sweep time and the
// mean whole-sort time (histogram + sweeps) in milliseconds
int gsb_debug_sort_bench(int device, uint64_t n, int key_bits, int iters, int tuning, double* sweep_ms, double* sort_ms, int* sweeps) {
    return guarded(nullptr, [&] {
        DebugDevice d; d.open(device);
        sort_set_tuning(tuning);
        const int kb = key_bits <= 64 ? 8 : 16;
        DevBuf<u8> a(&d.ws, n * kb), b(&d.ws, n * kb);
        cudaEvent_t e0, e1;
        GSB_CUDA_TRY(cudaEventCreate(&e0)); GSB_CUDA_TRY(cudaEventCreate(&e1));
        double sw = 0, tot = 0; int run = 0, total_run = 0;
        for (int it = 0; it < iters + 1; ++it) {
            sort_fill_random(kb, a.p, n, key_bits, 1234567ull * (it + 1), d.ws.stream);
            double sw_it = 0;
            GSB_CUDA_TRY(cudaEventRecord(e0, d.ws.stream));
            sort_keys(d.ws, kb, key_bits, a.p, b.p, nullptr, nullptr, n, nullptr, &run, &sw_it);
            GSB_CUDA_TRY(cudaEventRecord(e1, d.ws.stream));
            GSB_CUDA_TRY(cudaEventSynchronize(e1));
            float ms = 0; GSB_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
            if (it > 0) { sw += sw_it; tot += ms; total_run += run; }     // first iteration is warm-up
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (sweep_ms) *sweep_ms = total_run ? sw / total_run : 0;
        if (sort_ms) *sort_ms = tot / iters;
        if (sweeps) *sweeps = total_run / (iters ? iters : 1);
        a.free(); b.free();
        sort_set_tuning(0);
    });
}

int gsb_debug_set_partition(int max_slots, int total_bits) { partition_set_debug((u32)(max_slots < 0 ? 0 : max_slots), total_bits); return GSB_OK; }

// host-only: the pass widths the counting / the pair sort would use (no device needed).  out = {levels, total bits, table
// slots or bucket capacity, bits[0..7]}
int gsb_debug_plan(int what, int key_bytes, int key_bits, uint64_t n, int first_bits, uint32_t* out) {
    if (!out || (key_bytes != 8 && key_bytes != 16)) return GSB_EINVAL;
    if (what == 0 || what == 1) {
        const PartitionPlan p = what == 0 ? partition_plan(key_bytes, n) : partition_plan_streamed(key_bytes, n, first_bits);
        out[0] = (uint32_t)p.levels; out[1] = (uint32_t)p.total_bits; out[2] = p.max_slots;
        for (int i = 0; i < 8; ++i) out[3 + i] = (uint32_t)p.bits[i];
        return GSB_OK;
    }
    if (what == 2) {
        const PairSortPlan p = pairsort_plan(key_bytes, key_bits, n, first_bits);
        out[0] = (uint32_t)p.levels; out[1] = (uint32_t)p.bits; out[2] = p.cap;
        for (int i = 0; i < 8; ++i) out[3 + i] = (uint32_t)p.lb[i];
        return GSB_OK;
    }
    return GSB_EINVAL;
}

int gsb_debug_set_pairsort(int cap, int bits) { pairsort_set_debug((u32)(cap < 0 ? 0 : cap), bits); return GSB_OK; }

int gsb_debug_set_tuning(int id) { g_legacy_counting = (id >> 16) & 1; g_sampled_survivors = (id >> 17) & 1; sort_set_tuning(id & 0xFFFF); return GSB_OK; }

int gsb_debug_emit_sparse_array(int device, const uint64_t* key_lo, const uint64_t* key_hi, uint64_t m,
                                uint64_t universe_lo, uint64_t universe_hi, uint64_t m_est,
                                const char* base, const gsb_sink* sink) {
    return guarded(nullptr, [&] {
        DebugDevice d; d.open(device);
        const int kb = key_hi ? 16 : 8;
        DevBuf<u8> keys;
        upload_keys(d.ws, kb, key_lo, key_hi, m, keys);
        Emitter em; em.ws = &d.ws; em.sink = sink; em.pinned = d.pinned; em.pinned_bytes = 16 << 20;
        U128 u{universe_lo, universe_hi};
        emit_sparse_array(em, kb, keys.p, m, u, m_est, u, base);
        keys.free();
    });
}

int gsb_debug_emit_graph(int device, const uint64_t* key_lo, const uint64_t* key_hi, const uint64_t* counts,
                         uint64_t m, int k, const char* prefix, const gsb_sink* sink) {
    gsb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = GSB_ABI_VERSION; cfg.kind = GSB_KIND_GRAPH; cfg.k = k; cfg.device = device;
    gsb_ctx* c = nullptr;
    int rc = gsb_create(&cfg, &c);
    if (rc != GSB_OK) return rc;
    rc = guarded(c, [&] {
        upload_keys(c->ws, c->key_bytes, key_lo, key_hi, m, c->acc.keys);
        c->acc.counts.reset(&c->ws, m);
        if (m) GSB_CUDA_TRY(cudaMemcpyAsync(c->acc.counts.p, counts, m * 8, cudaMemcpyHostToDevice, c->ws.stream));
        c->ws.sync();
        c->acc.m = m; c->have_acc = true; c->counted = true;
    });
    if (rc == GSB_OK) rc = gsb_emit(c, prefix, sink);
    if (rc != GSB_OK) g_create_error = c->err;
    gsb_destroy(c);
    return rc;
}

}  // extern "C"
