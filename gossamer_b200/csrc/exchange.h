// exchange.h -- multi-GPU range partition + all-to-all of reduced (key,count) runs over NCCL.
#pragma once
#include "kernels.h"

namespace gsb {

struct Exchange;

static const u32 kExchangeSamplesPerRank = 2048;
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
                        DevBuf<u8>& recv, u64* recv_cap, u8** recv_ptr_out, u64* n_recv, ExchangeTiming* timing);
// sum of one u64 over all ranks
u64 exchange_sum(Exchange* x, Workspace& ws, u64 v);
// concatenate all ranks' runs on rank 0 in rank order (other ranks end up empty)
void exchange_gather(Exchange* x, Workspace& ws, int key_bytes, ReducedRun& run);

}  // namespace gsb
