/* gossamer_b200.h -- C ABI of the B200-native graph-construction path.
 *
 * This is the drop-in boundary for the reference's `build-graph` / `build-kmer-set` hot path.
 * The reference has no FFI; its seam for this path is the command object
 *   GossCmdBuildGraph::operator()(const GossCmdContext&)      src/GossCmdBuildGraph.cc:270-426
 *   GossCmdBuildKmerSet::operator()(const GossCmdContext&)    src/GossCmdBuildKmerSet.tcc:212-332
 * whose body is: read blocks -> (k+1)-mer/k-mer stream -> BackyardHash count -> sort ->
 * Graph::Builder / KmerSet::Builder -> files through a FileFactory.  Each entry point below
 * names the piece of that body it replaces.  INTEGRATION.md shows the patch a maintainer
 * would apply to GossCmdBuildGraph.cc to call it.
 *
 * Conventions: plain C, no exceptions, no STL, no torch types.  Every call returns 0 or a
 * negative gsb_status; gsb_last_error() has the detail text.  One gsb_ctx is driven by one
 * host thread at a time.  The library owns all device memory and its own pinned staging; the
 * caller owns the input buffers (until the call returns) and the file handles behind the sink.
 * There is NO CPU fallback: without a usable CUDA device every call fails with GSB_ECUDA.
 */
#ifndef GOSSAMER_B200_H
#define GOSSAMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSB_ABI_VERSION 1u

typedef struct gsb_ctx gsb_ctx;

typedef enum gsb_status {
    GSB_OK = 0,
    GSB_EINVAL = -1,  /* bad argument / call order */
    GSB_EPARSE = -2,  /* malformed FASTA/FASTQ; message matches the reference's text (src/FastqParser.hh:89-175) */
    GSB_EIO = -3,     /* sink callback failed */
    GSB_ENOMEM = -4,
    GSB_ECUDA = -5,   /* no device, kernel or runtime failure */
    GSB_ENCCL = -6,
    GSB_ERANGE = -7   /* k out of range: Graph::MaxK = 62 (src/Graph.hh:89), KmerSet::MaxK = 63 (src/KmerSet.hh:30) */
} gsb_status;

typedef enum gsb_kind {
    GSB_KIND_GRAPH = 0,   /* build-graph: every (k+1)-mer and its reverse complement, with counts */
    GSB_KIND_KMERSET = 1  /* build-kmer-set: FNV-normalised k-mers, no counts */
} gsb_kind;

typedef enum gsb_format {
    GSB_FMT_FASTA = 0,  /* -I / -F   src/FastaParser.hh:51-87  */
    GSB_FMT_FASTQ = 1,  /* -i / -f   src/FastqParser.hh:78-176 */
    GSB_FMT_LINE = 2    /* --line-in src/LineParser.hh:71-82   */
} gsb_format;

/* gsb_push_* flags */
#define GSB_BLOCK_LAST_OF_FILE 1u /* the file ends with this block; without it the next block of the same
                                     format continues the same file (FASTA records may straddle blocks) */
#define GSB_BLOCK_ASYNC 2u        /* gsb_push_block only: start the host-to-device copy of this block and return after
                                    processing the PREVIOUS asynchronous block, so that the copy overlaps the device
                                    work (scan, pack, extraction) of its predecessor.  `data` must be page-locked
                                    (gsb_host_alloc) and stay unchanged until the next gsb_push_block /
                                    gsb_finish_counting on this context returns; a parse error in this block is
                                    reported by that later call. */

typedef void (*gsb_log_fn)(void* user, int severity /*0 info,1 warning,2 error*/, const char* message);

typedef struct gsb_config {
    uint32_t abi_version;      /* GSB_ABI_VERSION */
    int32_t kind;              /* gsb_kind */
    int32_t k;                 /* -k  (src/GossCmdBuildGraph.cc:433-434) */
    int32_t device;            /* CUDA ordinal */
    uint64_t min_count;        /* 0/1 = keep all.  m>1 == build-graph followed by `trim-graph -C m-1`
                                  (src/GossCmdTrimGraph.cc:97-124); ignored for kmer sets */
    uint64_t max_batch_keys;   /* 0 = choose from free HBM.  Keys buffered before a count+merge
                                  round (the -B analogue, src/GossCmdBuildGraph.cc:436-447) */
    gsb_log_fn log;            /* may be NULL */
    void* log_user;
} gsb_config;

typedef struct gsb_counts {
    uint64_t n_reads;      /* records framed */
    uint64_t n_instances;  /* keys the reference feeds to counting (both strands for graphs) */
    uint64_t n_distinct;   /* distinct keys (all ranks when a communicator is attached) */
    uint64_t n_kept;       /* after the min-count filter == number of edges / k-mers emitted */
} gsb_counts;

/* Output seam: the FileFactory::out analogue (src/FileFactory.hh:80-164).  The library composes
 * the reference's file names (`P.header`, `P-edges.high-bits`, ...) and hands each file over
 * as one or more contiguous pieces at increasing offsets. */
typedef struct gsb_sink {
    void* user;
    int (*open)(void* user, const char* name, uint64_t size_hint, void** handle);
    int (*pwrite)(void* user, void* handle, uint64_t offset, const void* data, uint64_t len);
    int (*close)(void* user, void* handle);
} gsb_sink;

typedef struct gsb_stats {
    /* device time per phase, milliseconds, CUDA events on the library's stream */
    double ms_h2d, ms_scan, ms_extract, ms_sort, ms_reduce, ms_merge, ms_emit, ms_d2h, ms_exchange;
    /* ms_sort = the partition passes of the counting (+ sorts of merged batches); ms_reduce = the shared-memory bucket count
     * with the min-count filter; ms_unfold = reverse complements + pair sort of the survivors */
    double ms_sort_sweeps;      /* sum of the partition (or radix sweep) kernels alone (events around each launch) */
    double ms_all_to_all;       /* multi-GPU: the pass that pulls the instances over NVLink, alone (inside ms_exchange) */
    double ms_unfold;           /* graph mode: reverse complements of the folded, filtered run + ordering by key */
    uint64_t exchange_bytes_sent; /* instance bytes this rank moved from / to OTHER ranks in the exchange */
    uint64_t exchange_peer_memory; /* 1 = exchange fused into the partition passes over peer windows (NVLink), 0 = NCCL send/recv */
    uint64_t bytes_in;          /* raw text bytes pushed */
    uint64_t bytes_out;         /* bytes handed to the sink */
    uint64_t n_symbols;         /* bases + separators in the packed symbol stream */
    uint64_t sort_key_bytes;    /* 8 or 16 */
    uint64_t sort_passes;       /* partition (or radix) passes actually launched */
    uint64_t sort_passes_model; /* ceil(keybits/8): the 8-bit LSD passes SURVEY 8d's B_sort charges */
    uint64_t n_batches;         /* count (+ merge) rounds */
    uint64_t kernel_launches;   /* launches of this library's kernels so far */
    uint64_t hbm_peak_bytes;    /* high-water mark of device allocations */
    uint64_t n_sorted_keys;     /* keys that went through the counting passes (graph mode: one per window,
                                   = n_instances / 2, because the two strands are folded) */
    uint64_t device_allocs;     /* cudaMalloc calls made by this context so far (NOT reset by gsb_reset): a steady-state
                                   step makes none -- every buffer comes from the context's caching allocator */
    /* multi-GPU: parts of ms_exchange (the rest is the instance exchange itself: local first pass, histograms, pull pass) */
    double ms_exchange_agree;     /* the collective that agrees on the finish path */
    double ms_exchange_survivors; /* re-partition of the surviving (key,count) pairs into the final order's slices */
    double ms_exchange_publish;   /* statistics all-gather / barrier that publishes the slices */
} gsb_stats;

/* Replaces: GossCmdFactoryBuildGraph::create parameter checks (src/GossCmdBuildGraph.cc:428-477)
 * and the BackyardHash / consumer-pool construction (:315-322). */
int gsb_create(const gsb_config* cfg, gsb_ctx** out);
void gsb_destroy(gsb_ctx* ctx);
const char* gsb_last_error(const gsb_ctx* ctx); /* ctx may be NULL for gsb_create failures */

/* Replaces: the ingest + k-merising loop, HOT LOOP A/B (src/GossCmdBuildGraph.cc:335-380;
 * LineSource/Fast{a,q}Parser/GossReadBaseString/ReverseComplementAdapter beneath it).
 * `data` is host memory (pinned for best throughput); it is copied to the device inside the
 * call.  A block must end at a line boundary (or end of file); a FASTQ / line block must end at
 * a record boundary; a FASTA record may continue into the next block. */
int gsb_push_block(gsb_ctx* ctx, const void* data, size_t nbytes, int format, uint32_t flags);
/* Same, for text already resident in device memory (no copy). */
int gsb_push_device_block(gsb_ctx* ctx, const void* device_data, size_t nbytes, int format, uint32_t flags);

/* Replaces: BackyardHash::insert + sort + the duplicate-merging emit loop's counting
 * (src/BackyardHash.cc:115-271, src/GossCmdBuildGraph.cc:239-258), flushNaked + AsyncMerge for
 * multi-batch input (:171-220, src/AsyncMerge.tcc:267-324), and trim-graph's predicate. */
int gsb_finish_counting(gsb_ctx* ctx, gsb_counts* out);

/* Replaces: Graph::Builder / KmerSet::Builder and everything under them (src/Graph.cc:115-167,
 * src/KmerSet.hh:61-103, SparseArray / DenseSelect / WordyBitVector / IntegerArray /
 * VariableByteArray builders).  Files are byte-identical to the reference's writers. */
int gsb_emit(gsb_ctx* ctx, const char* prefix, const gsb_sink* sink);
/* sink == NULL builds every file in device memory and drops it (device-only timing; bytes_out still counts). */

/* ---- existing file sets: trim-graph, merge-graphs / merge-kmer-sets, dump-graph, restore-graph ------------------------
 * Input seam: the FileFactory::in analogue (src/FileFactory.hh:80-164).  size() returns 0 and the size if the file exists. */
typedef struct gsb_source {
    void* user;
    int (*size)(void* user, const char* name, uint64_t* size_out);
    int (*pread)(void* user, const char* name, uint64_t offset, void* dst, uint64_t len);
} gsb_source;

typedef struct gsb_graph_info {
    uint64_t version;   /* 2011101014 (Graph, src/Graph.hh:73-83) or 2011101701 (KmerSet, src/KmerSet.hh:32-43) */
    uint64_t k;
    uint64_t flags;     /* Graph: bit 0 = asymmetric; KmerSet: the stored count */
    uint64_t n_items;   /* edges / k-mers in the set (the SparseArray header's count) */
} gsb_graph_info;

/* Host only: what `prefix` holds (kind: gsb_kind).  Replaces Graph::LazyIterator's header checks (src/Graph.cc:195-216). */
int gsb_graph_peek(const char* prefix, const gsb_source* src, int kind, gsb_graph_info* out, char* err, size_t errcap);
/* Decodes the file set `prefix` (Elias-Fano edges / k-mers, VariableByteArray counts) into a sorted (key, count) run on the
 * device and merges it into the context's run, summing the counts of equal keys -- the read side of
 * Graph::LazyIterator / CursorMerge / PairMerge (src/GossCmdMerge.tcc:29-145).  The context's k must match the file's. */
int gsb_graph_load(gsb_ctx* ctx, const char* prefix, const gsb_source* src);
/* Same, from host arrays (restore-graph's parsed lines, src/GossCmdRestoreGraph.cc:70-128); any order. */
int gsb_graph_load_pairs(gsb_ctx* ctx, const uint64_t* key_lo, const uint64_t* key_hi, const uint64_t* counts, uint64_t m);
/* Ends loading: keeps the items with count > cutoff (trim-graph's predicate, src/GossCmdTrimGraph.cc:119; 0 keeps all) and
 * fixes the size estimate the builders are parameterised with: m_est = 0 means "the number kept" (trim-graph passes the
 * exact n), merge passes the sum of the inputs' sizes (src/GossCmdMerge.tcc:224-263,296), restore the header's n.
 * gsb_emit then writes the file set. */
int gsb_graph_finish(gsb_ctx* ctx, uint64_t cutoff, uint64_t m_est, gsb_counts* out);
/* xenome index, steps 3 and 4 (src/XenoApp.cc:62-76), on a GSB_KIND_KMERSET context of the sets' k.
 * Replaces: GossCmdMergeAndAnnotateKmerSets::operator() (src/GossCmdMergeAndAnnotateKmerSets.cc:27-207): the union of
 * the kmer sets <lhs_prefix> and <rhs_prefix> (read through `src`) is written as kmer set <out_prefix> plus the membership
 * bit vectors <out_prefix>.lhs-bits / .rhs-bits.  stats (optional) = {n_lhs, n_rhs, n_common, n_out}. */
int gsb_kmerset_merge_annotate(gsb_ctx* ctx, const char* lhs_prefix, const char* rhs_prefix, const gsb_source* src, const char* out_prefix,
                               const gsb_sink* sink, uint64_t* stats);
/* Replaces: GossCmdComputeNearKmers::operator() (src/GossCmdComputeNearKmers.cc:57-225): k-mers of exactly one side that
 * have a one-sided member of the OTHER side among their variants turn gray (both bits cleared); <prefix>.lhs-bits /
 * .rhs-bits are rewritten through the sink.  *n_gray (optional) = how many turned gray. */
int gsb_kmerset_near_kmers(gsb_ctx* ctx, const char* prefix, const gsb_source* src, const gsb_sink* sink, uint64_t* n_gray);

/* dump-graph's text (src/GossCmdDumpGraph.cc:31-60) of the finished run, formatted on the device, as one file `name`. */
int gsb_graph_dump(gsb_ctx* ctx, const char* name, const gsb_sink* sink);

/* Device-side stopwatch on the library's stream (CUDA events): begin records, end records +
 * synchronises and returns the elapsed milliseconds between the two. */
int gsb_timer_begin(gsb_ctx* ctx);
int gsb_timer_end(gsb_ctx* ctx, double* ms_out);

/* Best effort: binds the calling thread to the CPUs of the NUMA node `device` hangs off, so that pinned buffers allocated
 * afterwards are local to the GPU (matters when several ranks stream blocks at once).  Call before gsb_host_alloc. */
int gsb_host_bind_near_device(int device);
/* Page-locked host buffers for gsb_push_block (so that the host program need not link CUDA itself). */
int gsb_host_alloc(size_t nbytes, void** out);
void gsb_host_free(void* p);

int gsb_get_stats(const gsb_ctx* ctx, gsb_stats* out);
/* Forget all input but keep buffers: lets a benchmark loop reuse one context. */
int gsb_reset(gsb_ctx* ctx);

/* Multi-GPU (one process per GPU).  Rank 0 makes an id, the host program broadcasts its 128
 * bytes (torch.distributed / MPI / a file), every rank attaches.  After that
 *   - gsb_finish_counting is a COLLECTIVE: the instances are range-partitioned by splitters sampled
 *     from all ranks and stored straight into the owners' peer-mapped receive windows over NVLink
 *     (CUDA IPC; grouped ncclSend/ncclRecv if peer memory cannot be mapped); each rank counts its
 *     range, and the survivors are re-partitioned the same way so that rank r ends up holding the
 *     r-th contiguous slice of the global sorted (edge, count) run;
 *   - gsb_emit is a COLLECTIVE too: every rank writes its own byte ranges of every file -- its sink
 *     receives pieces as open(name, size of the WHOLE file) / pwrite(offset, ...) / close, a file may
 *     be opened more than once, and bytes that no rank writes are zero (open must not truncate a
 *     file another rank has written to: O_CREAT without O_TRUNC + ftruncate(size));
 *   - gsb_gather_to_root (optional) moves all slices to rank 0 instead; gsb_emit then hands the
 *     complete files to rank 0's sink only.  It is also what gsb_emit falls back to without peer
 *     memory. */
#define GSB_NCCL_ID_BYTES 128
int gsb_comm_make_id(void* id_out /* GSB_NCCL_ID_BYTES */);
int gsb_comm_attach(gsb_ctx* ctx, const void* id, int n_ranks, int rank);
int gsb_gather_to_root(gsb_ctx* ctx);
/* Host-only piece of the exchange (no device needed): splitters at equal quantiles of the pooled
 * sample of all ranks' keys, given as (lo,hi) pairs; writes n_ranks-1 (lo,hi) pairs.  Each rank
 * contributes gsb_samples_per_rank() keys taken at positions floor((i+0.5)*m/S) of its sorted run. */
int gsb_plan_splitters(const uint64_t* samples, uint64_t n_samples, int n_ranks, uint64_t* splitters_out);
uint32_t gsb_samples_per_rank(void);

/* Test-only: copy the reduced (key,count) run of this rank to host arrays. */
int64_t gsb_debug_copy_counts(gsb_ctx* ctx, uint64_t* key_lo, uint64_t* key_hi, uint64_t* counts, uint64_t cap);
/* Test-only: run one component on host-provided arrays (inputs copied to the device, the
 * component's kernels run, results returned through the sink or out arrays). */
int64_t gsb_debug_sort_keys(int device, uint64_t* key_lo, uint64_t* key_hi, uint64_t n, int key_bits);
int gsb_debug_sort_bench(int device, uint64_t n, int key_bits, int iters, int tuning, double* sweep_ms, double* sort_ms, int* sweeps);
int gsb_debug_set_tuning(int id);   /* bits 0-7 sweep tile shape, 8-15 profiling ablations, bit 16: contexts created afterwards count with the
                                       full LSD sort of raw keys (the round-1 path) instead of by partitioning */
int gsb_debug_set_partition(int max_slots, int total_bits); /* bucket geometry of the partition counting, 0 = default; results never depend on it */
int gsb_debug_plan(int what, int key_bytes, int key_bits, uint64_t n, int first_bits, uint32_t* out); /* host-only: pass widths of the counting (what = 0), of a streamed build (1), of the pair sort (2): out[11] = {levels, bits, slots|capacity, bits[8]} */
int gsb_debug_set_pairsort(int cap, int bits); /* bucket capacity (0 = default) and partition bits (< 0 = default) of the pair sort; results never depend on them */
int gsb_debug_emit_sparse_array(int device, const uint64_t* key_lo, const uint64_t* key_hi, uint64_t m,
                                uint64_t universe_lo, uint64_t universe_hi, uint64_t m_est,
                                const char* base, const gsb_sink* sink);
int gsb_debug_emit_graph(int device, const uint64_t* key_lo, const uint64_t* key_hi, const uint64_t* counts,
                         uint64_t m, int k, const char* prefix, const gsb_sink* sink);
int64_t gsb_debug_extract(int device, const void* text, size_t nbytes, int format, int kind, int k,
                          uint64_t* key_lo, uint64_t* key_hi, uint64_t cap, uint64_t* n_reads,
                          char* err, size_t errcap);

#ifdef __cplusplus
}
#endif
#endif /* GOSSAMER_B200_H */
