// exchange.h -- multi-GPU range partition + all-to-all of reduced (key,count) runs over NCCL.
#pragma once
#include "kernels.h"

namespace gsb {

struct Exchange;

static const u32 kExchangeSamplesPerRank = 2048;

// The final run of a multi-GPU build: globally sorted and range-partitioned, rank r's slice sitting in
// r's peer-mapped window (CUDA IPC over NVLink), so that every rank can read any element.  This is what
// the distributed emitters (emit.cu) work from: each rank writes its own byte ranges of every file and
// fetches the few elements it needs from its neighbours' slices straight from their memory.
struct DistRun {
    int n = 1, rank = 0;
    const u8* keys[kMaxRanks];
    const u64* counts[kMaxRanks];
    u64 off[kMaxRanks + 1];         // off[r] = global index of rank r's first element, off[n] = total
};
void plan_splitters(const u64* samples, u64 n_samples, int n_ranks, u64* splitters_out);

void exchange_make_id(void* id_out /* GSB_NCCL_ID_BYTES */);
Exchange* exchange_create(const void* id, int n_ranks, int rank, Workspace& ws);
void exchange_destroy(Exchange* x);
// Range-partition `run` (sorted distinct keys + counts) by sampled splitters, all-to-all, merge:
// afterwards rank r holds the r-th contiguous slice of the global order.
void exchange_runs(Exchange* x, Workspace& ws, int key_bytes, int key_bits, ReducedRun& run);
struct ExchangeTiming { double ms_all_to_all = 0; u64 bytes_sent_remote = 0; bool used_peer_memory = false; };
// Range-partition raw instance keys by sampled splitters and exchange them with one all-to-all.
// Preferred path: ONE kernel partitions the instances and stores each destination's run straight into
// that rank's receive window over NVLink (CUDA-IPC-mapped peer memory); an all-reduce is the only
// collective (a barrier).  If the windows cannot be mapped, it falls back to a staged partition
// (`parted_buf`, scratch for n_keys keys) + grouped ncclSend/ncclRecv into `recv`.
// *recv_ptr_out is where the received instances are (the window or recv.p).
void exchange_instances(Exchange* x, Workspace& ws, int key_bytes, const void* keys, u64 n_keys, u8* parted_buf,
                        DevBuf<u8>& recv, u64* recv_cap, u8** recv_ptr_out, u64* n_recv, ExchangeTiming* timing,
                        u64* piggyback_sum = nullptr /* in: this rank's number, out: the sum over all ranks (rides on the sample all-gather) */);
// Instance exchange fused into the partition counting (exchange.cu): every rank runs the first pass (top bits0 bits of the
// mixed key) locally into its peer-mapped window; rank r owns the children [ceil(r*2^bits0/n), ceil((r+1)*2^bits0/n)) and its
// second pass (bits1 more bits) pulls them tile by tile out of all ranks' windows over NVLink into out_local (capacity
// out_cap_keys).  Collective; no host synchronisation inside.  false (on every rank): peer memory unusable.
struct PartitionedInstances { u8* recv = nullptr; DevBuf<u64> cstart; u64 n_parents = 0, n_cap = 0; int bits = 0; };
bool exchange_partition_pull(Exchange* x, Workspace& ws, int key_bytes, void* keys, u64 n_keys, const u64* hist_top, const std::vector<u64>& n_keys_all,
                             int bits0, int bits1, void* out_local, u64 out_cap_keys, PartitionedInstances* out);
// the following are valid once the host has synchronised with the stream
bool exchange_partition_aborted(const Exchange* x);      // some rank's buffer was too small for its share: nothing was pulled
u64 exchange_partition_bytes_sent(const Exchange* x);
u64 exchange_partition_received(const Exchange* x);      // keys that arrived in this rank's window
double exchange_partition_scatter_ms(const Exchange* x);  // the pulling pass
double exchange_partition_level0_ms(const Exchange* x);   // the local first pass (small kernels around it included)
int exchange_rank(const Exchange* x);
int exchange_size(const Exchange* x);
// Collective.  Copies this rank's slice into its window and returns the global view; false if peer
// memory cannot be mapped here (the caller then gathers to rank 0 instead).
bool exchange_publish(Exchange* x, Workspace& ws, int key_bytes, const ReducedRun& run, DistRun* out);
bool exchange_peer_memory_usable(const Exchange* x);
// Collective.  Range-partition UNSORTED (key,count) pairs by sampled splitters and store each pair straight
// into its owner's window over NVLink; this rank's pairs end up at *recv_keys / *recv_counts (inside the
// window, arbitrary order), totals[r] = pairs received by rank r.  false: peer memory unavailable.
bool exchange_pairs_p2p(Exchange* x, Workspace& ws, int key_bytes, const void* keys, const u64* counts, u64 n_pairs,
                        u8** recv_keys, u64** recv_counts, std::vector<u64>* totals);
// Collective.  The same result with the first pass of the pair sort (partition.cu) crossing NVLink: children of the top key
// bits are dealt to the ranks as balanced contiguous ranges (one all-gather of histograms, no samples); `sorted` = this
// rank's slice, ordered, also copied into its window.  false: peer memory unavailable (take exchange_pairs_p2p).
bool exchange_pairs_msd(Exchange* x, Workspace& ws, int key_bytes, int key_bits, const void* keys, const u64* counts, u64 m, int fold_w,
                        ReducedRun& sorted, std::vector<u64>* totals);
// global view of slices already sitting in the windows (layout of exchange_pairs_p2p / exchange_publish)
void exchange_view(const Exchange* x, int key_bytes, const std::vector<u64>& totals, DistRun* out);
// Collective: n_words u64 from every rank, concatenated in rank order on the host of every rank.
void exchange_allgather_u64(Exchange* x, Workspace& ws, const u64* mine_host, size_t n_words, std::vector<u64>& all_host);
// Collective: when it returns, every rank has finished the work it enqueued before the call.
void exchange_barrier(Exchange* x, Workspace& ws);
// Collective: rank 0 receives the concatenation (rank order) of every rank's bytes[r] bytes at `mine`.
void exchange_gatherv_root(Exchange* x, Workspace& ws, const void* mine, const std::vector<u64>& bytes, DevBuf<u8>& out_root);
// sum of one u64 over all ranks
u64 exchange_sum(Exchange* x, Workspace& ws, u64 v);
// concatenate all ranks' runs on rank 0 in rank order (other ranks end up empty)
void exchange_gather(Exchange* x, Workspace& ws, int key_bytes, ReducedRun& run);

// ---- emit.cu: multi-GPU emission from the published run (collective: every rank calls them) ----
void emit_sparse_array_dist(Emitter& em, Exchange* x, int key_bytes, const DistRun& run, U128 universe_ctor, u64 m_est, U128 universe_end,
                            const std::string& base);
void emit_counts_dist(Emitter& em, Exchange* x, const DistRun& run, u64 m_est, const std::string& base);
void emit_count_histogram_dist(Emitter& em, Exchange* x, const DistRun& run, const std::string& name);

}  // namespace gsb
