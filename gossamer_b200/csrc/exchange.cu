// exchange.cu -- multi-GPU steps of the counting phase (one process per GPU).
//
// The reference has no distributed mode; its scale-out advice is "build
// parts, then merge-graphs" (docs/goss.md:315-321).  Here every rank extracts the folded instance
// keys of its own share of the reads; the instances, and later the surviving (key,count) pairs, are
// range-partitioned by splitters taken from a sample of all ranks' keys and stored STRAIGHT INTO THE
// OWNER'S MEMORY over NVLink: every rank has a receive window (cudaMalloc + CUDA IPC, mapped by all
// peers), the partition kernel groups a tile's elements by destination in shared memory and writes
// each run into the owner's window at the offset a count matrix reserved for this source.  Rank r
// ends up owning -- and publishing in its window -- the r-th contiguous slice of the global sorted
// edge set, which the distributed emitters (emit.cu) read.  NCCL carries only the small all-gathers
// (samples, count matrices, statistics) and barriers; grouped ncclSend/ncclRecv is the data-path
// fallback when peer memory cannot be mapped, and exchange_runs / exchange_gather are the
// sorted-run variants used for merged batches and for gsb_gather_to_root.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch already loaded when the
// host program is a torchrun rank), so single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <ctime>

#include "exchange.h"
#include "scan.cuh"

namespace gsb {

static const int kPtExchangeMaxChildren = 1 << kTopHistBits;

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.lib) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
    if (!api.lib) throw StatusError{GSB_ENCCL, std::string("cannot load libnccl.so.2: ") + dlerror()};
#define GSB_SYM(field, sym)                                                                    \
    *(void**)(&api.field) = dlsym(api.lib, sym);                                              \
    if (!api.field) throw StatusError{GSB_ENCCL, std::string("libnccl lacks ") + sym};
    GSB_SYM(GetUniqueId, "ncclGetUniqueId")
    GSB_SYM(CommInitRank, "ncclCommInitRank")
    GSB_SYM(CommDestroy, "ncclCommDestroy")
    GSB_SYM(AllGather, "ncclAllGather")
    GSB_SYM(AllReduce, "ncclAllReduce")
    GSB_SYM(Send, "ncclSend")
    GSB_SYM(Recv, "ncclRecv")
    GSB_SYM(GroupStart, "ncclGroupStart")
    GSB_SYM(GroupEnd, "ncclGroupEnd")
    GSB_SYM(GetErrorString, "ncclGetErrorString")
#undef GSB_SYM
    return api;
}

void check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw StatusError{GSB_ENCCL, std::string(what) + ": " + nccl().GetErrorString(r)};
}


// sample[i] = key at position floor((i + 0.5) * m / S), as (lo, hi) pairs; slot S holds m
template <typename K>
__global__ void sample_kernel(const K* __restrict__ keys, u64 m, u32 S, u64* __restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S && m) {
        u64 p = (u64)(((double)i + 0.5) * (double)m / (double)S);
        if (p >= m) p = m - 1;
        K k = keys[p];
        out[2 * i] = KeyOps<K>::lo(k); out[2 * i + 1] = KeyOps<K>::hi(k);
    }
    if (i == 0) { out[2 * S] = m; out[2 * S + 1] = 0; }
}

// bounds[j + 1] = first index whose key is >= splitter j
template <typename K>
__global__ void bounds_kernel(const K* __restrict__ keys, u64 m, const u64* __restrict__ splitters, u32 n_split, u64* __restrict__ bounds) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_split) return;
    K sp = KeyOps<K>::make(splitters[2 * j], splitters[2 * j + 1]);
    u64 lo = 0, hi = m;
    while (lo < hi) { u64 mid = lo + ((hi - lo) >> 1); if (KeyOps<K>::lt(keys[mid], sp)) lo = mid + 1; else hi = mid; }
    bounds[j + 1] = lo;
}

struct Splitters { u64 lo[kMaxRanks], hi[kMaxRanks]; int n; };      // n = number of splitters (ranks - 1)

template <typename K>
__device__ __forceinline__ int dest_of(const K& k, const Splitters& sp) {
    int d = 0;                                                  // rank r owns [splitter[r-1], splitter[r])
    for (int j = 0; j < sp.n; ++j) d += !KeyOps<K>::lt(k, KeyOps<K>::make(sp.lo[j], sp.hi[j]));
    return d;
}

// strided sample of the (unsorted) instance keys
template <typename K>
__global__ void sample_strided_kernel(const K* __restrict__ keys, u64 n, u32 S, u64* __restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S && n) {
        u64 p = (u64)(((double)i + 0.5) * (double)n / (double)S);
        if (p >= n) p = n - 1;
        K k = keys[p];
        out[2 * i] = KeyOps<K>::lo(k); out[2 * i + 1] = KeyOps<K>::hi(k);
    }
    if (i == 0) { out[2 * S] = n; out[2 * S + 1] = 0; }
}

static const int kPartThreads = 256;
static const int kPartItems = 8;

// lanes (among those with ok) whose small value d equals mine, from `bits` ballots
__device__ __forceinline__ u32 same_dest_lanes(int d, int bits, bool ok) {
    u32 peers = __ballot_sync(0xffffffffu, ok);
    for (int b = 0; b < bits; ++b) {
        const bool bit = (d >> b) & 1;
        const u32 bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

// pass 1: how many keys go to each destination (warp-aggregated: one shared atomic per distinct
// destination per warp instruction -- consecutive instances are not sorted, so a warp typically
// sees every destination; the match is on a value < 32, done with a short ballot loop over ranks
// only when there are few ranks, else by per-lane shared atomics)
template <typename K>
__global__ void __launch_bounds__(kPartThreads) dest_count_kernel(const K* __restrict__ keys, u64 n, Splitters sp, u64* __restrict__ totals) {
    __shared__ u32 cnt_s[kMaxRanks];
    if (threadIdx.x < kMaxRanks) cnt_s[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int bits = 0; while ((1 << bits) <= sp.n) ++bits;
    for (u64 base = (u64)blockIdx.x * kPartThreads * kPartItems; base < n; base += (u64)gridDim.x * kPartThreads * kPartItems) {
#pragma unroll
        for (int i = 0; i < kPartItems; ++i) {
            const u64 idx = base + (u64)i * kPartThreads + threadIdx.x;
            const bool ok = idx < n;
            const int d = ok ? dest_of(keys[idx], sp) : 0;
            const u32 peers = same_dest_lanes(d, bits, ok);
            if (ok && (peers & ((1u << lane) - 1)) == 0) atomicAdd(&cnt_s[d], (u32)__popc(peers));
        }
    }
    __syncthreads();
    if (threadIdx.x <= sp.n && cnt_s[threadIdx.x]) atomicAdd(&totals[threadIdx.x], (u64)cnt_s[threadIdx.x]);
}

// pass 2: scatter into per-destination regions (cursor[r] starts at the region offset); order inside
// a region is arbitrary -- the receiver sorts.  Keys are staged through shared memory grouped by
// destination so that the global (or peer) writes of one destination are contiguous.
template <typename K>
__global__ void __launch_bounds__(kPartThreads) dest_scatter_kernel(const K* __restrict__ keys, u64 n, Splitters sp, u64* __restrict__ cursor,
                                                                   K* __restrict__ out) {
    constexpr int TILE = kPartThreads * kPartItems;
    __shared__ u32 cnt_s[kMaxRanks], start_s[kMaxRanks + 1];
    __shared__ u64 base_s[kMaxRanks];
    __shared__ K stage[TILE];
    __shared__ u8 stage_dest[TILE];
    const int lane = threadIdx.x & 31;
    int bits = 0; while ((1 << bits) <= sp.n) ++bits;
    for (u64 base = (u64)blockIdx.x * TILE; base < n; base += (u64)gridDim.x * TILE) {
        if (threadIdx.x < kMaxRanks) cnt_s[threadIdx.x] = 0;
        __syncthreads();
        K k[kPartItems]; int d[kPartItems]; u32 slot[kPartItems];
#pragma unroll
        for (int i = 0; i < kPartItems; ++i) {
            const u64 idx = base + (u64)i * kPartThreads + threadIdx.x;
            d[i] = -1; slot[i] = 0;
            const bool ok = idx < n;
            if (ok) { k[i] = keys[idx]; d[i] = dest_of(k[i], sp); }
            const u32 peers = same_dest_lanes(ok ? d[i] : 0, bits, ok);
            const int leader = __ffs(peers) - 1;
            u32 before = 0;
            if (ok && lane == leader) before = atomicAdd(&cnt_s[d[i]], (u32)__popc(peers));
            before = __shfl_sync(0xffffffffu, before, leader < 0 ? lane : leader);
            slot[i] = before + __popc(peers & ((1u << lane) - 1));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 run = 0;
            for (int r = 0; r <= sp.n; ++r) { start_s[r] = run; run += cnt_s[r]; }
            start_s[sp.n + 1] = run;
        }
        if (threadIdx.x <= sp.n) base_s[threadIdx.x] = cnt_s[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], (u64)cnt_s[threadIdx.x]) : 0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kPartItems; ++i)
            if (d[i] >= 0) { const u32 p = start_s[d[i]] + slot[i]; stage[p] = k[i]; stage_dest[p] = (u8)d[i]; }
        __syncthreads();
        const u32 total = start_s[sp.n + 1];
        for (u32 j = threadIdx.x; j < total; j += kPartThreads) {
            const int r = stage_dest[j];
            out[base_s[r] + (j - start_s[r])] = stage[j];
        }
        __syncthreads();
    }
}

// Fused partition + transfer: the same tile-local grouping by destination, but each destination's
// run is stored straight into that rank's receive window over NVLink (peer-mapped memory), at the
// offset reserved for this source rank.  No staging copy, no collective on the data path.
struct PeerWindows { void* base[kMaxRanks]; u64* vbase[kMaxRanks]; };   // window of rank r, already offset to this source's region

template <typename K, int ITEMS, bool HAS_VALUES>
__global__ void __launch_bounds__(kPartThreads) dest_scatter_p2p_kernel(const K* __restrict__ keys, const u64* __restrict__ values, u64 n, Splitters sp,
                                                                       u64* __restrict__ cursor, PeerWindows win) {
    constexpr int TILE = kPartThreads * ITEMS;
    __shared__ u32 cnt_s[kMaxRanks], start_s[kMaxRanks + 1];
    __shared__ u64 base_s[kMaxRanks];
    __shared__ K stage[TILE];
    __shared__ u64 stage_v[HAS_VALUES ? TILE : 1];
    __shared__ u8 stage_dest[TILE];
    const int lane = threadIdx.x & 31;
    int bits = 0; while ((1 << bits) <= sp.n) ++bits;
    for (u64 base = (u64)blockIdx.x * TILE; base < n; base += (u64)gridDim.x * TILE) {
        if (threadIdx.x < kMaxRanks) cnt_s[threadIdx.x] = 0;
        __syncthreads();
        K k[ITEMS]; u64 v[HAS_VALUES ? ITEMS : 1]; int d[ITEMS]; u32 slot[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const u64 idx = base + (u64)i * kPartThreads + threadIdx.x;
            d[i] = -1; slot[i] = 0;
            const bool ok = idx < n;
            if (ok) { k[i] = keys[idx]; d[i] = dest_of(k[i], sp); if (HAS_VALUES) v[i] = values[idx]; }
            const u32 peers = same_dest_lanes(ok ? d[i] : 0, bits, ok);
            const int leader = __ffs(peers) - 1;
            u32 before = 0;
            if (ok && lane == leader) before = atomicAdd(&cnt_s[d[i]], (u32)__popc(peers));
            before = __shfl_sync(0xffffffffu, before, leader < 0 ? lane : leader);
            slot[i] = before + __popc(peers & ((1u << lane) - 1));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 run = 0;
            for (int r = 0; r <= sp.n; ++r) { start_s[r] = run; run += cnt_s[r]; }
            start_s[sp.n + 1] = run;
        }
        if (threadIdx.x <= sp.n) base_s[threadIdx.x] = cnt_s[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], (u64)cnt_s[threadIdx.x]) : 0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (d[i] >= 0) { const u32 p = start_s[d[i]] + slot[i]; stage[p] = k[i]; stage_dest[p] = (u8)d[i]; if (HAS_VALUES) stage_v[p] = v[i]; }
        __syncthreads();
        const u32 total = start_s[sp.n + 1];
        for (u32 j = threadIdx.x; j < total; j += kPartThreads) {
            const int r = stage_dest[j];
            const u64 o = base_s[r] + (j - start_s[r]);
            static_cast<K*>(win.base[r])[o] = stage[j];
            if (HAS_VALUES) win.vbase[r][o] = stage_v[j];
        }
        __syncthreads();
    }
}

}  // namespace

// n_ranks - 1 splitters at equal quantiles of the pooled sample; rank r owns keys in
// [splitter[r-1], splitter[r]).  Pure host code (also exported for the CPU multi-process tests).
void plan_splitters(const u64* samples /* (lo,hi) pairs */, u64 n_samples, int n_ranks, u64* splitters_out) {
    struct HK { u64 hi, lo; bool operator<(const HK& o) const { return hi < o.hi || (hi == o.hi && lo < o.lo); } };
    std::vector<HK> v(n_samples);
    for (u64 i = 0; i < n_samples; ++i) v[i] = HK{samples[2 * i + 1], samples[2 * i]};
    std::sort(v.begin(), v.end());
    for (int j = 0; j < n_ranks - 1; ++j) {
        const HK& k = v[(size_t)((u64)(j + 1) * n_samples / n_ranks)];
        splitters_out[2 * j] = k.lo; splitters_out[2 * j + 1] = k.hi;
    }
}

struct Exchange {
    ncclComm_t comm = nullptr;
    int n = 1, rank = 0;
    // Peer-memory receive window (the fused partition + transfer path): a plain cudaMalloc allocation,
    // exported with CUDA IPC, mapped by every peer; grows only.
    u8* recv_buf = nullptr;
    u64 recv_cap_bytes = 0;
    std::vector<u8*> peer_ptr;                     // [rank] mapped base of that rank's window (own = recv_buf)
    std::vector<cudaIpcMemHandle_t> peer_handle;   // handle currently mapped for each peer
    std::vector<char> peer_open;
    bool p2p_usable = true;                        // cleared if IPC mapping fails: NCCL send/recv is used instead
    std::vector<u64> peer_cap;                     // [rank] capacity of that rank's window as of the last handshake (same on every rank)
    // partition exchange: results the host reads after its next synchronisation (pinned), events around the scatter
    u64* host_slots = nullptr;                     // [0] abort flag, [1] keys this rank pulled from OTHER ranks' windows, [2] keys received
    u64 slot_bytes_scale = 1;
    cudaEvent_t scatter_e0 = nullptr, scatter_e1 = nullptr, l0_e0 = nullptr, l0_e1 = nullptr;
};

void exchange_make_id(void* id_out) {
    static_assert(sizeof(ncclUniqueId) == GSB_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id_out, &id, sizeof(id));
}

Exchange* exchange_create(const void* id_bytes, int n_ranks, int rank, Workspace& ws) {
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    Exchange* x = new Exchange();
    x->n = n_ranks; x->rank = rank;
    GSB_CUDA_TRY(cudaSetDevice(ws.device));
    ncclResult_t r = nccl().CommInitRank(&x->comm, n_ranks, id, rank);
    if (r != ncclSuccess) { delete x; check(r, "ncclCommInitRank"); }
    return x;
}

void exchange_destroy(Exchange* x) {
    if (!x) return;
    for (size_t r = 0; r < x->peer_ptr.size(); ++r)
        if ((int)r != x->rank && x->peer_open[r]) cudaIpcCloseMemHandle(x->peer_ptr[r]);
    if (x->recv_buf) cudaFree(x->recv_buf);
    if (x->host_slots) cudaFreeHost(x->host_slots);
    if (x->scatter_e0) cudaEventDestroy(x->scatter_e0);
    if (x->scatter_e1) cudaEventDestroy(x->scatter_e1);
    if (x->l0_e0) cudaEventDestroy(x->l0_e0);
    if (x->l0_e1) cudaEventDestroy(x->l0_e1);
    if (x->comm) nccl().CommDestroy(x->comm);
    delete x;
}

u64 exchange_sum(Exchange* x, Workspace& ws, u64 v) {
    DevBuf<u64> d(&ws, 2);
    GSB_CUDA_TRY(cudaMemcpyAsync(d.p, &v, 8, cudaMemcpyHostToDevice, ws.stream));
    check(nccl().AllReduce(d.p, d.p + 1, 1, ncclUint64, ncclSum, x->comm, ws.stream), "ncclAllReduce");
    u64 out = 0;
    GSB_CUDA_TRY(cudaMemcpyAsync(&out, d.p + 1, 8, cudaMemcpyDeviceToHost, ws.stream));
    ws.sync();
    return out;
}

// Collective: afterwards every rank's receive window holds at least the bytes that rank asked for and is
// mapped (CUDA IPC) by every other rank.  Returns false -- on every rank alike -- if that is not possible
// here (IPC not permitted, no peer access): the callers then use NCCL send/recv.
// need_all (optional): the bytes EVERY rank needs, known to every rank (from a count matrix): if all the windows
// mapped at the last handshake are large enough, nothing has to be agreed and no collective runs.
static bool ensure_windows(Exchange* x, Workspace& ws, u64 need_bytes, const std::vector<u64>* need_all = nullptr) {
    const int n = x->n;
    cudaStream_t s = ws.stream;
    NcclApi& api = nccl();
    if (need_all && (int)x->peer_cap.size() == n) {
        bool enough = true;
        for (int r = 0; r < n; ++r) enough = enough && (*need_all)[r] <= x->peer_cap[r];
        if (enough) return true;
    }
    if ((int)x->peer_ptr.size() == n) {
        // A window that is about to be replaced must not be freed while a peer still maps it (CUDA IPC: undefined
        // behaviour).  Whether ANY window may grow is something every rank sees alike: the needs (need_all) and the capacities
        // of the last handshake are common knowledge; a caller without need_all may grow its own window, which only it
        // knows, so then every mapping is dropped.  Peers close first, a barrier, and only then the owners free.
        bool any = false;
        for (int r = 0; r < n; ++r) {
            const bool replace = !need_all || (int)x->peer_cap.size() != n || (*need_all)[r] > x->peer_cap[r];
            if (!replace) continue;
            any = true;
            if (r != x->rank && x->peer_open[r]) { cudaIpcCloseMemHandle(x->peer_ptr[r]); x->peer_open[r] = 0; x->peer_ptr[r] = nullptr; }
        }
        if (any) (void)exchange_sum(x, ws, 0);                     // every peer has closed what will be freed
    }
    if (need_bytes > x->recv_cap_bytes || !x->recv_buf) {
        ws.sync();
        if (x->recv_buf) GSB_CUDA_TRY(cudaFree(x->recv_buf));
        x->recv_buf = nullptr;
        x->recv_cap_bytes = need_bytes + need_bytes / 8 + (1u << 20);
        GSB_CUDA_TRY(cudaMalloc((void**)&x->recv_buf, x->recv_cap_bytes));
        ++ws.device_allocs;
    }
    struct Slot { cudaIpcMemHandle_t h; u64 cap; u64 ok; };
    Slot mine_h;
    memset(&mine_h, 0, sizeof(mine_h));
    mine_h.ok = cudaIpcGetMemHandle(&mine_h.h, x->recv_buf) == cudaSuccess ? 1 : 0;
    if (!mine_h.ok) cudaGetLastError();
    mine_h.cap = x->recv_cap_bytes;
    DevBuf<u8> h_mine(&ws, sizeof(Slot)), h_all(&ws, sizeof(Slot) * n);
    GSB_CUDA_TRY(cudaMemcpyAsync(h_mine.p, &mine_h, sizeof(Slot), cudaMemcpyHostToDevice, s));
    check(api.AllGather(h_mine.p, h_all.p, sizeof(Slot), ncclUint8, x->comm, s), "ncclAllGather(ipc handles)");
    std::vector<Slot> slots(n);
    GSB_CUDA_TRY(cudaMemcpyAsync(slots.data(), h_all.p, sizeof(Slot) * n, cudaMemcpyDeviceToHost, s));
    ws.sync();
    bool all_ok = true;
    for (int r = 0; r < n; ++r) all_ok = all_ok && slots[r].ok;
    if (x->peer_ptr.empty()) { x->peer_ptr.assign(n, nullptr); x->peer_handle.resize(n); x->peer_open.assign(n, 0); }
    u64 local_ok = all_ok ? 1 : 0;
    if (all_ok) {
        for (int r = 0; r < n && local_ok; ++r) {
            if (r == x->rank) { x->peer_ptr[r] = x->recv_buf; continue; }
            if (x->peer_open[r] && memcmp(&x->peer_handle[r], &slots[r].h, sizeof(cudaIpcMemHandle_t)) == 0) continue;
            if (x->peer_open[r]) { cudaIpcCloseMemHandle(x->peer_ptr[r]); x->peer_open[r] = 0; }
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, slots[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); local_ok = 0; break; }
            x->peer_ptr[r] = (u8*)p; x->peer_handle[r] = slots[r].h; x->peer_open[r] = 1;
        }
    }
    const u64 n_ok = exchange_sum(x, ws, local_ok);                 // every rank must take the same path
    x->peer_cap.clear();
    if (n_ok == (u64)n) for (int r = 0; r < n; ++r) x->peer_cap.push_back(slots[r].cap);
    return n_ok == (u64)n;
}

static inline u64 align256(u64 v) { return (v + 255) & ~255ull; }
void exchange_view(const Exchange* x, int key_bytes, const std::vector<u64>& totals, DistRun* out);

int exchange_rank(const Exchange* x) { return x->rank; }
int exchange_size(const Exchange* x) { return x->n; }

void exchange_allgather_u64(Exchange* x, Workspace& ws, const u64* mine_host, size_t n_words, std::vector<u64>& all_host) {
    const int n = x->n;
    cudaStream_t s = ws.stream;
    DevBuf<u64> mine(&ws, n_words), all(&ws, n_words * n);
    GSB_CUDA_TRY(cudaMemcpyAsync(mine.p, mine_host, n_words * 8, cudaMemcpyHostToDevice, s));
    check(nccl().AllGather(mine.p, all.p, n_words, ncclUint64, x->comm, s), "ncclAllGather");
    all_host.resize(n_words * n);
    GSB_CUDA_TRY(cudaMemcpyAsync(all_host.data(), all.p, all_host.size() * 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
}

void exchange_barrier(Exchange* x, Workspace& ws) { (void)exchange_sum(x, ws, 0); }

void exchange_gatherv_root(Exchange* x, Workspace& ws, const void* mine, const std::vector<u64>& bytes, DevBuf<u8>& out_root) {
    const int n = x->n;
    cudaStream_t s = ws.stream;
    NcclApi& api = nccl();
    if (x->rank == 0) {
        u64 total = 0;
        for (int r = 0; r < n; ++r) total += bytes[r];
        out_root.reset(&ws, total);
        if (bytes[0]) GSB_CUDA_TRY(cudaMemcpyAsync(out_root.p, mine, bytes[0], cudaMemcpyDeviceToDevice, s));
        check(api.GroupStart(), "ncclGroupStart");
        u64 off = bytes[0];
        for (int r = 1; r < n; ++r) {
            if (bytes[r]) check(api.Recv(out_root.p + off, bytes[r], ncclUint8, r, x->comm, s), "ncclRecv(gatherv)");
            off += bytes[r];
        }
        check(api.GroupEnd(), "ncclGroupEnd");
    } else {
        check(api.GroupStart(), "ncclGroupStart");
        if (bytes[x->rank]) check(api.Send(mine, bytes[x->rank], ncclUint8, 0, x->comm, s), "ncclSend(gatherv)");
        check(api.GroupEnd(), "ncclGroupEnd");
    }
    ws.sync();
}


bool exchange_publish(Exchange* x, Workspace& ws, int key_bytes, const ReducedRun& run, DistRun* out) {
    if (!x->p2p_usable) return false;
    const int n = x->n;
    cudaStream_t s = ws.stream;
    const u64 need = align256(run.m * key_bytes) + run.m * 8;
    if (!ensure_windows(x, ws, need)) { x->p2p_usable = false; return false; }
    if (run.m) {
        GSB_CUDA_TRY(cudaMemcpyAsync(x->recv_buf, run.keys.p, run.m * key_bytes, cudaMemcpyDeviceToDevice, s));
        GSB_CUDA_TRY(cudaMemcpyAsync(x->recv_buf + align256(run.m * key_bytes), run.counts.p, run.m * 8, cudaMemcpyDeviceToDevice, s));
    }
    // the all-gather of the slice sizes doubles as the barrier: it completes here only after every rank
    // has enqueued it, i.e. after that rank's copies above
    std::vector<u64> m_all;
    exchange_allgather_u64(x, ws, &run.m, 1, m_all);
    exchange_view(x, key_bytes, m_all, out);
    return true;
}

void exchange_runs(Exchange* x, Workspace& ws, int key_bytes, int key_bits, ReducedRun& run) {
    const int n = x->n;
    if (n == 1) return;
    cudaStream_t s = ws.stream;
    NcclApi& api = nccl();
    const u32 S = kExchangeSamplesPerRank;
    const size_t slot = 2 * (size_t)S + 2;                        // u64 words per rank
    // 1. sample + allgather
    DevBuf<u64> mine(&ws, slot), all(&ws, slot * n);
    GSB_CUDA_TRY(cudaMemsetAsync(mine.p, 0, slot * 8, s));
    if (key_bytes == 8) sample_kernel<u64><<<(S + 255) / 256, 256, 0, s>>>((const u64*)run.keys.p, run.m, S, mine.p);
    else sample_kernel<Key128><<<(S + 255) / 256, 256, 0, s>>>((const Key128*)run.keys.p, run.m, S, mine.p);
    ++ws.launches;
    check(api.AllGather(mine.p, all.p, slot, ncclUint64, x->comm, s), "ncclAllGather(samples)");
    std::vector<u64> h(slot * n);
    GSB_CUDA_TRY(cudaMemcpyAsync(h.data(), all.p, h.size() * 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    std::vector<u64> samples;                                     // (lo, hi) pairs of every rank that has keys
    for (int r = 0; r < n; ++r) {
        const u64* p = h.data() + slot * r;
        if (p[2 * S] == 0) continue;
        samples.insert(samples.end(), p, p + 2 * (size_t)S);
    }
    if (samples.empty()) return;                                  // nothing anywhere
    std::vector<u64> split(2 * (size_t)(n - 1));
    plan_splitters(samples.data(), samples.size() / 2, n, split.data());
    // 2. local partition bounds
    DevBuf<u64> split_d(&ws, split.size()), bounds_d(&ws, (size_t)n + 1);
    GSB_CUDA_TRY(cudaMemcpyAsync(split_d.p, split.data(), split.size() * 8, cudaMemcpyHostToDevice, s));
    GSB_CUDA_TRY(cudaMemsetAsync(bounds_d.p, 0, 8, s));
    GSB_CUDA_TRY(cudaMemcpyAsync(bounds_d.p + n, &run.m, 8, cudaMemcpyHostToDevice, s));
    if (key_bytes == 8) bounds_kernel<u64><<<1, 64, 0, s>>>((const u64*)run.keys.p, run.m, split_d.p, (u32)(n - 1), bounds_d.p);
    else bounds_kernel<Key128><<<1, 64, 0, s>>>((const Key128*)run.keys.p, run.m, split_d.p, (u32)(n - 1), bounds_d.p);
    ++ws.launches;
    std::vector<u64> bounds((size_t)n + 1);
    GSB_CUDA_TRY(cudaMemcpyAsync(bounds.data(), bounds_d.p, bounds.size() * 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    // 3. counts matrix
    std::vector<u64> send_cnt(n);
    for (int r = 0; r < n; ++r) send_cnt[r] = bounds[r + 1] - bounds[r];
    DevBuf<u64> cnt_mine(&ws, n), cnt_all(&ws, (size_t)n * n);
    GSB_CUDA_TRY(cudaMemcpyAsync(cnt_mine.p, send_cnt.data(), (size_t)n * 8, cudaMemcpyHostToDevice, s));
    check(api.AllGather(cnt_mine.p, cnt_all.p, n, ncclUint64, x->comm, s), "ncclAllGather(counts)");
    std::vector<u64> cnt((size_t)n * n);
    GSB_CUDA_TRY(cudaMemcpyAsync(cnt.data(), cnt_all.p, cnt.size() * 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    std::vector<u64> recv_cnt(n), recv_off(n + 1, 0);
    for (int r = 0; r < n; ++r) { recv_cnt[r] = cnt[(size_t)r * n + x->rank]; recv_off[r + 1] = recv_off[r] + recv_cnt[r]; }
    const u64 total = recv_off[n];
    // 4. all-to-all of keys and counts
    DevBuf<u8> rkeys(&ws, total * key_bytes), rkeys_alt(&ws, total * key_bytes);
    DevBuf<u64> rcounts(&ws, total), rcounts_alt(&ws, total);
    check(api.GroupStart(), "ncclGroupStart");
    for (int r = 0; r < n; ++r) {
        if (send_cnt[r]) {
            check(api.Send(run.keys.p + bounds[r] * key_bytes, send_cnt[r] * key_bytes, ncclUint8, r, x->comm, s), "ncclSend(keys)");
            check(api.Send(run.counts.p + bounds[r], send_cnt[r], ncclUint64, r, x->comm, s), "ncclSend(counts)");
        }
        if (recv_cnt[r]) {
            check(api.Recv(rkeys.p + recv_off[r] * key_bytes, recv_cnt[r] * key_bytes, ncclUint8, r, x->comm, s), "ncclRecv(keys)");
            check(api.Recv(rcounts.p + recv_off[r], recv_cnt[r], ncclUint64, r, x->comm, s), "ncclRecv(counts)");
        }
    }
    check(api.GroupEnd(), "ncclGroupEnd");
    ws.sync();
    run.keys.free(); run.counts.free(); run.m = 0;
    // 5. merge the n sorted runs that arrived: sort by key carrying counts, sum equal keys
    int where = sort_keys(ws, key_bytes, key_bits, rkeys.p, rkeys_alt.p, rcounts.p, rcounts_alt.p, total, nullptr, nullptr);
    ReducedRun merged; u64 distinct = 0;
    reduce_sorted(ws, key_bytes, where ? rkeys_alt.p : rkeys.p, where ? rcounts_alt.p : rcounts.p, total, 1, merged, &distinct);
    run = std::move(merged);
}

// Range-partition the raw instance keys by sampled splitters and exchange them (one all-to-all):
// afterwards `recv` holds every instance, from all ranks, whose key lies in this rank's range.
static double host_now_ms() {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

void exchange_instances(Exchange* x, Workspace& ws, int key_bytes, const void* keys, u64 n_keys, u8* parted_buf /* n_keys keys */,
                        DevBuf<u8>& recv, u64* recv_cap, u8** recv_ptr_out, u64* n_recv, ExchangeTiming* timing, u64* piggyback_sum) {
#ifdef GSB_PROFILING
    const bool trace = getenv("GSB_TRACE_EXCHANGE") != nullptr;
#else
    const bool trace = false;
#endif
    double t_prev = host_now_ms();
    auto lap = [&](const char* what) {
        if (!trace) return;
        ws.sync();
        double t = host_now_ms();
        fprintf(stderr, "[exchange rank %d] %-28s %8.3f ms\n", x->rank, what, t - t_prev);
        t_prev = t;
    };
    const int n = x->n;
    cudaStream_t s = ws.stream;
    NcclApi& api = nccl();
    if (n > kMaxRanks) throw StatusError{GSB_EINVAL, "at most 32 ranks are supported"};
    const u32 S = kExchangeSamplesPerRank;
    const size_t slot = 2 * (size_t)S + 2;
    DevBuf<u64> mine(&ws, slot), all(&ws, slot * n);
    GSB_CUDA_TRY(cudaMemsetAsync(mine.p, 0, slot * 8, s));
    if (key_bytes == 8) sample_strided_kernel<u64><<<(S + 255) / 256, 256, 0, s>>>((const u64*)keys, n_keys, S, mine.p);
    else sample_strided_kernel<Key128><<<(S + 255) / 256, 256, 0, s>>>((const Key128*)keys, n_keys, S, mine.p);
    ++ws.launches;
    // the spare word of the sample slot carries one caller-supplied number that is summed over the ranks
    if (piggyback_sum) GSB_CUDA_TRY(cudaMemcpyAsync(mine.p + 2 * S + 1, piggyback_sum, 8, cudaMemcpyHostToDevice, s));
    check(api.AllGather(mine.p, all.p, slot, ncclUint64, x->comm, s), "ncclAllGather(samples)");
    std::vector<u64> h(slot * n);
    GSB_CUDA_TRY(cudaMemcpyAsync(h.data(), all.p, h.size() * 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    std::vector<u64> samples;
    u64 piggy = 0;
    for (int r = 0; r < n; ++r) {
        const u64* p = h.data() + slot * r;
        piggy += p[2 * S + 1];
        if (p[2 * S] == 0) continue;
        samples.insert(samples.end(), p, p + 2 * (size_t)S);
    }
    if (piggyback_sum) *piggyback_sum = piggy;
    Splitters sp;
    memset(&sp, 0, sizeof(sp));
    sp.n = n - 1;
    if (!samples.empty()) {
        std::vector<u64> split(2 * (size_t)(n - 1));
        plan_splitters(samples.data(), samples.size() / 2, n, split.data());
        for (int j = 0; j < n - 1; ++j) { sp.lo[j] = split[2 * j]; sp.hi[j] = split[2 * j + 1]; }
    }
    lap("sample + splitters");
    // local partition: count, offsets, scatter
    DevBuf<u64> totals(&ws, 2 * (size_t)kMaxRanks);
    GSB_CUDA_TRY(cudaMemsetAsync(totals.p, 0, totals.bytes(), s));
    const u64 tiles = (n_keys + kPartThreads * kPartItems - 1) / (kPartThreads * kPartItems);
    const int grid = (int)std::max<u64>(1, std::min<u64>(tiles, (u64)ws.sm_count * 8));
    if (n_keys) {
        if (key_bytes == 8) dest_count_kernel<u64><<<grid, kPartThreads, 0, s>>>((const u64*)keys, n_keys, sp, totals.p);
        else dest_count_kernel<Key128><<<grid, kPartThreads, 0, s>>>((const Key128*)keys, n_keys, sp, totals.p);
        ++ws.launches;
    }
    std::vector<u64> send_cnt(kMaxRanks, 0), send_off(n + 1, 0);
    GSB_CUDA_TRY(cudaMemcpyAsync(send_cnt.data(), totals.p, kMaxRanks * 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    for (int r = 0; r < n; ++r) send_off[r + 1] = send_off[r] + send_cnt[r];
    GSB_CUDA_TRY(cudaMemcpyAsync(totals.p + kMaxRanks, send_off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, s));
    lap("count per destination");
    // counts matrix: cnt[src][dst]
    DevBuf<u64> cnt_mine(&ws, n), cnt_all(&ws, (size_t)n * n);
    GSB_CUDA_TRY(cudaMemcpyAsync(cnt_mine.p, send_cnt.data(), (size_t)n * 8, cudaMemcpyHostToDevice, s));
    check(api.AllGather(cnt_mine.p, cnt_all.p, n, ncclUint64, x->comm, s), "ncclAllGather(counts)");
    std::vector<u64> cnt((size_t)n * n);
    GSB_CUDA_TRY(cudaMemcpyAsync(cnt.data(), cnt_all.p, cnt.size() * 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    std::vector<u64> recv_cnt(n), recv_off(n + 1, 0);
    for (int r = 0; r < n; ++r) { recv_cnt[r] = cnt[(size_t)r * n + x->rank]; recv_off[r + 1] = recv_off[r] + recv_cnt[r]; }
    const u64 total = recv_off[n];
    u64 sent_remote = 0;
    for (int r = 0; r < n; ++r) if (r != x->rank) sent_remote += send_cnt[r] * key_bytes;

    lap("counts matrix allgather");
    cudaEvent_t e0, e1;
    GSB_CUDA_TRY(cudaEventCreate(&e0)); GSB_CUDA_TRY(cudaEventCreate(&e1));
    bool done = false;
    if (x->p2p_usable) {
        // ---- peer-memory path: make sure every rank's window is big enough and mapped everywhere ----
        std::vector<u64> need_all(n, 0);
        for (int src = 0; src < n; ++src) for (int dst = 0; dst < n; ++dst) need_all[dst] += cnt[(size_t)src * n + dst] * key_bytes;
        const bool windows_ok = ensure_windows(x, ws, total * key_bytes, &need_all);
        lap("windows: (re)allocation, ipc handles, agree on path");
        if (windows_ok) {
            PeerWindows win;
            memset(&win, 0, sizeof(win));
            for (int r = 0; r < n; ++r) {
                u64 before_me = 0;                                  // keys that lower-ranked sources send to r
                for (int src = 0; src < x->rank; ++src) before_me += cnt[(size_t)src * n + r];
                win.base[r] = x->peer_ptr[r] + before_me * key_bytes;
            }
            GSB_CUDA_TRY(cudaMemsetAsync(totals.p + kMaxRanks, 0, kMaxRanks * 8, s));
            GSB_CUDA_TRY(cudaEventRecord(e0, s));
            if (n_keys) {
                if (key_bytes == 8) dest_scatter_p2p_kernel<u64, kPartItems, false><<<grid, kPartThreads, 0, s>>>((const u64*)keys, nullptr, n_keys, sp, totals.p + kMaxRanks, win);
                else dest_scatter_p2p_kernel<Key128, kPartItems, false><<<grid, kPartThreads, 0, s>>>((const Key128*)keys, nullptr, n_keys, sp, totals.p + kMaxRanks, win);
                ++ws.launches;
            }
            // barrier: when this all-reduce completes here, every rank's scatter kernel has finished
            DevBuf<u64> flag(&ws, 2);
            GSB_CUDA_TRY(cudaMemsetAsync(flag.p, 0, 16, s));
            check(api.AllReduce(flag.p, flag.p + 1, 1, ncclUint64, ncclSum, x->comm, s), "ncclAllReduce(barrier)");
            GSB_CUDA_TRY(cudaEventRecord(e1, s));
            GSB_CUDA_TRY(cudaEventSynchronize(e1));
            *recv_ptr_out = x->recv_buf;
            done = true;
            lap("fused scatter + barrier");
        } else {
            x->p2p_usable = false;                                  // e.g. IPC not permitted in this container
        }
    }
    if (!done) {
        // ---- NCCL path: partition into a staging buffer, then grouped send/recv ----
        GSB_CUDA_TRY(cudaMemcpyAsync(totals.p + kMaxRanks, send_off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, s));
        if (n_keys) {
            if (key_bytes == 8) dest_scatter_kernel<u64><<<grid, kPartThreads, 0, s>>>((const u64*)keys, n_keys, sp, totals.p + kMaxRanks, (u64*)parted_buf);
            else dest_scatter_kernel<Key128><<<grid, kPartThreads, 0, s>>>((const Key128*)keys, n_keys, sp, totals.p + kMaxRanks, (Key128*)parted_buf);
            ++ws.launches;
        }
        if (total > *recv_cap) {                                      // grow-only: steady-state steps allocate nothing
            recv.free();
            *recv_cap = total + total / 16 + 1024;
            recv.reset(&ws, *recv_cap * key_bytes + 64);
        }
        GSB_CUDA_TRY(cudaEventRecord(e0, s));
        check(api.GroupStart(), "ncclGroupStart");
        for (int r = 0; r < n; ++r) {
            if (send_cnt[r]) check(api.Send(parted_buf + send_off[r] * key_bytes, send_cnt[r] * key_bytes, ncclUint8, r, x->comm, s), "ncclSend(instances)");
            if (recv_cnt[r]) check(api.Recv(recv.p + recv_off[r] * key_bytes, recv_cnt[r] * key_bytes, ncclUint8, r, x->comm, s), "ncclRecv(instances)");
        }
        check(api.GroupEnd(), "ncclGroupEnd");
        GSB_CUDA_TRY(cudaEventRecord(e1, s));
        GSB_CUDA_TRY(cudaEventSynchronize(e1));
        *recv_ptr_out = recv.p;
    }
    float ms = 0;
    GSB_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (timing) { timing->ms_all_to_all += ms; timing->bytes_sent_remote += sent_remote; timing->used_peer_memory = done; }
    *n_recv = total;
}

// ---- instance exchange fused into the partition counting (PULL) --------------------------------------------------
// The instances are bit-mixed, so equal shares of the mixed key space are equal shares of the instances whatever the
// genome looks like: child c of the first partition pass (the top bits0 bits) belongs to rank (c * n) >> bits0 -- no
// sampling, no splitters.
//   1. every rank runs the first pass LOCALLY, into its peer-mapped window (HBM speed, long runs);
//   2. it histograms the next bits1 bits of every child of its own output; one all-reduce of those histograms and one
//      all-gather of the child starts (both device-side, no host round trip) tell every rank what it owns and where it lies;
//   3. the SECOND pass of the owner reads its children tile by tile from wherever they are: the partition kernel's bulk
//      copies (cp.async.bulk -> shared memory, one tile ahead) fetch 7/8 of them from the peers' windows over NVLink.
// There is no separate transfer step and no staging copy: the exchange IS the read side of a pass that has to happen
// anyway, and NVLink sees 32 KB bulk reads instead of short scattered stores.
bool exchange_partition_pull(Exchange* x, Workspace& ws, int key_bytes, void* keys, u64 n_keys, const u64* hist_top, const std::vector<u64>& n_keys_all,
                             int bits0, int bits1, void* out_local, u64 out_cap_keys, PartitionedInstances* out) {
    if (!x->p2p_usable) return false;
    const int n = x->n;
    cudaStream_t s = ws.stream;
    NcclApi& api = nccl();
    if (bits0 < 1 || bits0 > kTopHistBits || (1 << bits0) < n || bits1 < 0 || bits1 > 10) throw StatusError{GSB_EINVAL, "internal: bad pass widths for the exchange"};
    const u32 C = 1u << bits0;
    std::vector<u64> need_all(n);
    for (int r = 0; r < n; ++r) need_all[r] = n_keys_all[r] * key_bytes + 64;
    if (!ensure_windows(x, ws, need_all[x->rank], &need_all)) { x->p2p_usable = false; return false; }
    if (!x->host_slots) {
        GSB_CUDA_TRY(cudaMallocHost((void**)&x->host_slots, 64));
        GSB_CUDA_TRY(cudaEventCreate(&x->scatter_e0));
        GSB_CUDA_TRY(cudaEventCreate(&x->scatter_e1));
        GSB_CUDA_TRY(cudaEventCreate(&x->l0_e0));
        GSB_CUDA_TRY(cudaEventCreate(&x->l0_e1));
    }
    // 1. first pass, local, into the window
    DevBuf<u64> cstart0;
    GSB_CUDA_TRY(cudaEventRecord(x->l0_e0, s));
    partition_local_level0(ws, key_bytes, keys, x->recv_buf, n_keys, bits0, hist_top, cstart0);
    GSB_CUDA_TRY(cudaEventRecord(x->l0_e1, s));
    // 2. histograms of the next bits; all-reduce + all-gather
    const size_t n_hist = (size_t)C << bits1;
    DevBuf<u64> hist_all(&ws, n_hist), gathered(&ws, (size_t)(C + 1) * n), scalars(&ws, 4);
    partition_next_hist(ws, key_bytes, x->recv_buf, cstart0.p, bits0, n_keys, bits1, hist_all.p);
    check(api.AllGather(cstart0.p, gathered.p, C + 1, ncclUint64, x->comm, s), "ncclAllGather(child starts)");
    check(api.AllReduce(hist_all.p, hist_all.p, n_hist, ncclUint64, ncclSum, x->comm, s), "ncclAllReduce(histograms)");
    // (the collectives also order this rank's pull behind every peer's first pass)
    GSB_CUDA_TRY(cudaMemsetAsync(scalars.p, 0, 32, s));
    partition_pull_check(ws, hist_all.p, gathered.p, bits0, bits1, n, x->rank, out_cap_keys, (u32*)scalars.p, scalars.p + 2, scalars.p + 1);
    GSB_CUDA_TRY(cudaMemcpyAsync(x->host_slots, scalars.p, 24, cudaMemcpyDeviceToHost, s));   // read by the host after its next synchronisation
    // 3. second pass, pulling
    const u32 lo = (u32)(((u64)x->rank * C + n - 1) / n), hi = (u32)(((u64)(x->rank + 1) * C + n - 1) / n);
    const void* bases[kMaxRanks];
    for (int r = 0; r < n; ++r) bases[r] = x->peer_ptr[r];
    partition_pull_level(ws, key_bytes, bases, n, gathered.p, bits0, lo, hi - lo, out_cap_keys, bits1, hist_all.p + ((size_t)lo << bits1), out_local,
                         (const u32*)scalars.p, out->cstart, x->scatter_e0, x->scatter_e1);
    out->recv = (u8*)out_local;
    out->n_parents = (u64)(hi - lo) << bits1;
    out->n_cap = out_cap_keys;
    out->bits = bits0 + bits1;
    x->slot_bytes_scale = (u64)key_bytes;
    return true;
}

// valid after the host has synchronised with the stream
bool exchange_partition_aborted(const Exchange* x) { return x->host_slots && (u32)x->host_slots[0] != 0; }
u64 exchange_partition_bytes_sent(const Exchange* x) { return x->host_slots ? x->host_slots[1] * x->slot_bytes_scale : 0; }
u64 exchange_partition_received(const Exchange* x) { return x->host_slots ? x->host_slots[2] : 0; }
double exchange_partition_scatter_ms(const Exchange* x) {
    float ms = 0;
    if (x->scatter_e0 && cudaEventElapsedTime(&ms, x->scatter_e0, x->scatter_e1) != cudaSuccess) { cudaGetLastError(); ms = 0; }
    return ms;
}

double exchange_partition_level0_ms(const Exchange* x) {
    float ms = 0;
    if (x->l0_e0 && cudaEventElapsedTime(&ms, x->l0_e0, x->l0_e1) != cudaSuccess) { cudaGetLastError(); ms = 0; }
    return ms;
}

bool exchange_peer_memory_usable(const Exchange* x) { return x->p2p_usable; }

// Range-partition UNSORTED (key,count) pairs by splitters sampled from all ranks and store each pair
// straight into its owner's window over NVLink (keys at the start of the window, counts behind them at a
// 256-byte boundary).  Afterwards this rank's window holds every pair of its key range, in arbitrary
// order; totals[r] = number of pairs rank r received.  Collective; false (on every rank) if peer memory
// cannot be used.
bool exchange_pairs_p2p(Exchange* x, Workspace& ws, int key_bytes, const void* keys, const u64* counts, u64 n_pairs,
                        u8** recv_keys, u64** recv_counts, std::vector<u64>* totals_out) {
    if (!x->p2p_usable) return false;
    const int n = x->n;
    cudaStream_t s = ws.stream;
    NcclApi& api = nccl();
    constexpr int ITEMS = 4;
    const u32 S = kExchangeSamplesPerRank;
    const size_t slot = 2 * (size_t)S + 2;
    DevBuf<u64> mine(&ws, slot), all(&ws, slot * n);
    GSB_CUDA_TRY(cudaMemsetAsync(mine.p, 0, slot * 8, s));
    if (key_bytes == 8) sample_strided_kernel<u64><<<(S + 255) / 256, 256, 0, s>>>((const u64*)keys, n_pairs, S, mine.p);
    else sample_strided_kernel<Key128><<<(S + 255) / 256, 256, 0, s>>>((const Key128*)keys, n_pairs, S, mine.p);
    ++ws.launches;
    check(api.AllGather(mine.p, all.p, slot, ncclUint64, x->comm, s), "ncclAllGather(samples)");
    std::vector<u64> h(slot * n);
    GSB_CUDA_TRY(cudaMemcpyAsync(h.data(), all.p, h.size() * 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    std::vector<u64> samples;
    for (int r = 0; r < n; ++r) {
        const u64* p = h.data() + slot * r;
        if (p[2 * S] == 0) continue;
        samples.insert(samples.end(), p, p + 2 * (size_t)S);
    }
    Splitters sp;
    memset(&sp, 0, sizeof(sp));
    sp.n = n - 1;
    if (!samples.empty()) {
        std::vector<u64> split(2 * (size_t)(n - 1));
        plan_splitters(samples.data(), samples.size() / 2, n, split.data());
        for (int j = 0; j < n - 1; ++j) { sp.lo[j] = split[2 * j]; sp.hi[j] = split[2 * j + 1]; }
    }
    DevBuf<u64> totals(&ws, 2 * (size_t)kMaxRanks);
    GSB_CUDA_TRY(cudaMemsetAsync(totals.p, 0, totals.bytes(), s));
    const u64 tiles = (n_pairs + kPartThreads * ITEMS - 1) / (kPartThreads * ITEMS);
    const int grid = (int)std::max<u64>(1, std::min<u64>(tiles, (u64)ws.sm_count * 8));
    if (n_pairs) {
        if (key_bytes == 8) dest_count_kernel<u64><<<grid, kPartThreads, 0, s>>>((const u64*)keys, n_pairs, sp, totals.p);
        else dest_count_kernel<Key128><<<grid, kPartThreads, 0, s>>>((const Key128*)keys, n_pairs, sp, totals.p);
        ++ws.launches;
    }
    std::vector<u64> send_cnt(kMaxRanks, 0);
    GSB_CUDA_TRY(cudaMemcpyAsync(send_cnt.data(), totals.p, kMaxRanks * 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    std::vector<u64> cnt;                                           // cnt[src][dst]
    exchange_allgather_u64(x, ws, send_cnt.data(), n, cnt);
    std::vector<u64> total(n, 0);
    for (int src = 0; src < n; ++src) for (int dst = 0; dst < n; ++dst) total[dst] += cnt[(size_t)src * n + dst];
    const u64 mine_total = total[x->rank];
    std::vector<u64> need_all(n);
    for (int r = 0; r < n; ++r) need_all[r] = align256(total[r] * key_bytes) + total[r] * 8;
    if (!ensure_windows(x, ws, need_all[x->rank], &need_all)) { x->p2p_usable = false; return false; }
    PeerWindows win;
    memset(&win, 0, sizeof(win));
    for (int r = 0; r < n; ++r) {
        u64 before_me = 0;
        for (int src = 0; src < x->rank; ++src) before_me += cnt[(size_t)src * n + r];
        win.base[r] = x->peer_ptr[r] + before_me * key_bytes;
        win.vbase[r] = (u64*)(x->peer_ptr[r] + align256(total[r] * key_bytes)) + before_me;
    }
    GSB_CUDA_TRY(cudaMemsetAsync(totals.p + kMaxRanks, 0, kMaxRanks * 8, s));
    if (n_pairs) {
        if (key_bytes == 8) dest_scatter_p2p_kernel<u64, ITEMS, true><<<grid, kPartThreads, 0, s>>>((const u64*)keys, counts, n_pairs, sp, totals.p + kMaxRanks, win);
        else dest_scatter_p2p_kernel<Key128, ITEMS, true><<<grid, kPartThreads, 0, s>>>((const Key128*)keys, counts, n_pairs, sp, totals.p + kMaxRanks, win);
        ++ws.launches;
    }
    exchange_barrier(x, ws);                                        // every rank's stores have landed
    *recv_keys = x->recv_buf;
    *recv_counts = (u64*)(x->recv_buf + align256(mine_total * key_bytes));
    *totals_out = total;
    return true;
}

// The same job as exchange_pairs_p2p -- every rank ends up with the r-th contiguous slice of the global order, sorted, in
// its window -- as the FIRST PASS OF THE PAIR SORT (partition.cu): the pairs are split by the top 10 bits of the real key,
// the 1024 children are dealt to the ranks as contiguous ranges of nearly equal size (one all-gather of the per-rank
// histograms; every rank derives the same plan on the device), each child's run is stored straight into its owner's
// window, and the owner finishes the sort locally (remaining passes + shared-memory sort per bucket).  No samples, no
// splitters, one host synchronisation (the totals) instead of three.
//   keys / counts [m]: this rank's survivors (folded when fold_w > 0: the reverse complements are added here).
//   On success: sorted holds this rank's slice (also copied into the window, layout of exchange_view), totals_out[r] the
//   slice sizes.  Collective; false (on every rank alike, nothing changed) if peer memory cannot be used or a slice would not
//   fit -- the caller then takes exchange_pairs_p2p.
bool exchange_pairs_msd(Exchange* x, Workspace& ws, int key_bytes, int key_bits, const void* keys, const u64* counts, u64 m, int fold_w,
                        ReducedRun& sorted, std::vector<u64>* totals_out) {
    if (!x->p2p_usable || x->n > kMaxRanks) return false;
    const int n = x->n;
    cudaStream_t s = ws.stream;
    NcclApi& api = nccl();
    const int bits0 = key_bits < 10 ? key_bits : 10;
    const u32 C = 1u << bits0;
    const u32 eb = pairsort_elem_bytes(key_bytes);
    const u64 n_cap_local = fold_w ? 2 * m : m;
    const u32 cstride = 32;
    DevBuf<u8> elems(&ws, n_cap_local * eb + 64);
    DevBuf<u64> n_dev(&ws, 1), hist_mine(&ws, C), hist_all(&ws, (size_t)C * n), cursor(&ws, (size_t)C * cstride), totals(&ws, kMaxRanks + 2), cstart_local(&ws, (size_t)C + 1);
    DevBuf<u8> owner(&ws, C);
    GSB_CUDA_TRY(cudaMemsetAsync(hist_mine.p, 0, hist_mine.bytes(), s));
    pairsort_pack(ws, key_bytes, key_bits, keys, counts, m, fold_w, elems.p, n_dev.p, hist_mine.p, bits0);
    // (the all-gather also orders every peer's stores into this rank's window behind this rank's last use of it)
    check(api.AllGather(hist_mine.p, hist_all.p, C, ncclUint64, x->comm, s), "ncclAllGather(pair histograms)");
    u32* range = (u32*)(totals.p + kMaxRanks);
    pairsort_plan_owners(ws, hist_all.p, n, x->rank, bits0, cursor.p, cstride, owner.p, totals.p, range, cstart_local.p);
    u64 h[kMaxRanks + 2];
    GSB_CUDA_TRY(cudaMemcpyAsync(h, totals.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    ws.sync();
    std::vector<u64> total(h, h + n), need_all(n);
    const u32 clo = (u32)(h[kMaxRanks] & 0xffffffffu), chi = (u32)(h[kMaxRanks] >> 32);
    for (int r = 0; r < n; ++r) need_all[r] = std::max<u64>(total[r] * eb, align256(total[r] * key_bytes) + total[r] * 8) + 256;
    if (!ensure_windows(x, ws, need_all[x->rank], &need_all)) { x->p2p_usable = false; return false; }
    void* bases[kMaxRanks];
    for (int r = 0; r < n; ++r) bases[r] = x->peer_ptr[r];
    pairsort_scatter_to_peers(ws, key_bytes, key_bits, elems.p, n_cap_local, n_dev.p, bits0, cursor.p, cstride, owner.p, bases, n);
    // barrier on the stream: what follows runs after every rank's stores have landed
    DevBuf<u64> flag(&ws, 2);
    GSB_CUDA_TRY(cudaMemsetAsync(flag.p, 0, 16, s));
    check(api.AllReduce(flag.p, flag.p + 1, 1, ncclUint64, ncclSum, x->comm, s), "ncclAllReduce(barrier)");
    elems.free();
    const u64 mine = total[x->rank];
    if (mine == 0) {
        sorted.keys.reset(&ws, 0); sorted.counts.reset(&ws, 0); sorted.m = 0;
    } else {
        const u64 n_parents = chi - clo;
        // the remaining passes: as if all 2^bits0 children were as full as this rank's (no collective below this point)
        PairSortPlan plan = pairsort_plan(key_bytes, key_bits, mine * C / std::max<u64>(1, n_parents), bits0);
        DevBuf<u8> other(&ws, plan.levels > 1 ? mine * eb + 64 : 1);
        if (!pairsort_finish(ws, key_bytes, key_bits, x->recv_buf, other.p, mine, nullptr, cstart_local, n_parents, bits0, plan, 1, nullptr, sorted)) {
            // the data defeat the bucket geometry: radix sort of what arrived (converted in place by the bucket kernel)
            DevBuf<u8> kb2(&ws, mine * key_bytes);
            DevBuf<u64> cb2(&ws, mine);
            const int where = sort_keys(ws, key_bytes, key_bits, sorted.keys.p, kb2.p, sorted.counts.p, cb2.p, mine, nullptr, nullptr);
            if (where) { sorted.keys = std::move(kb2); sorted.counts = std::move(cb2); }
            sorted.m = mine;
        }
        // the slice, published in the window for the emitters of the other ranks
        GSB_CUDA_TRY(cudaMemcpyAsync(x->recv_buf, sorted.keys.p, mine * key_bytes, cudaMemcpyDeviceToDevice, s));
        GSB_CUDA_TRY(cudaMemcpyAsync(x->recv_buf + align256(mine * key_bytes), sorted.counts.p, mine * 8, cudaMemcpyDeviceToDevice, s));
    }
    *totals_out = total;
    return true;
}

// The global view of slices that already sit in the windows (layout of exchange_pairs_p2p / exchange_publish).
void exchange_view(const Exchange* x, int key_bytes, const std::vector<u64>& totals, DistRun* out) {
    const int n = x->n;
    out->n = n; out->rank = x->rank;
    out->off[0] = 0;
    for (int r = 0; r < n; ++r) {
        out->off[r + 1] = out->off[r] + totals[r];
        out->keys[r] = x->peer_ptr[r];
        out->counts[r] = (const u64*)(x->peer_ptr[r] + align256(totals[r] * key_bytes));
    }
    for (int r = n; r < kMaxRanks; ++r) { out->keys[r] = nullptr; out->counts[r] = nullptr; out->off[r + 1] = out->off[n]; }
}

void exchange_gather(Exchange* x, Workspace& ws, int key_bytes, ReducedRun& run) {
    const int n = x->n;
    if (n == 1) return;
    cudaStream_t s = ws.stream;
    NcclApi& api = nccl();
    DevBuf<u64> mine(&ws, 1), all(&ws, n);
    GSB_CUDA_TRY(cudaMemcpyAsync(mine.p, &run.m, 8, cudaMemcpyHostToDevice, s));
    check(api.AllGather(mine.p, all.p, 1, ncclUint64, x->comm, s), "ncclAllGather(sizes)");
    std::vector<u64> m(n), off(n + 1, 0);
    GSB_CUDA_TRY(cudaMemcpyAsync(m.data(), all.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    for (int r = 0; r < n; ++r) off[r + 1] = off[r] + m[r];
    if (x->rank == 0) {
        DevBuf<u8> keys(&ws, off[n] * key_bytes);
        DevBuf<u64> counts(&ws, off[n]);
        if (run.m) {
            GSB_CUDA_TRY(cudaMemcpyAsync(keys.p, run.keys.p, run.m * key_bytes, cudaMemcpyDeviceToDevice, s));
            GSB_CUDA_TRY(cudaMemcpyAsync(counts.p, run.counts.p, run.m * 8, cudaMemcpyDeviceToDevice, s));
        }
        check(api.GroupStart(), "ncclGroupStart");
        for (int r = 1; r < n; ++r) {
            if (!m[r]) continue;
            check(api.Recv(keys.p + off[r] * key_bytes, m[r] * key_bytes, ncclUint8, r, x->comm, s), "ncclRecv(gather keys)");
            check(api.Recv(counts.p + off[r], m[r], ncclUint64, r, x->comm, s), "ncclRecv(gather counts)");
        }
        check(api.GroupEnd(), "ncclGroupEnd");
        ws.sync();
        run.keys = std::move(keys); run.counts = std::move(counts); run.m = off[n];
    } else {
        check(api.GroupStart(), "ncclGroupStart");
        if (run.m) {
            check(api.Send(run.keys.p, run.m * key_bytes, ncclUint8, 0, x->comm, s), "ncclSend(gather keys)");
            check(api.Send(run.counts.p, run.m, ncclUint64, 0, x->comm, s), "ncclSend(gather counts)");
        }
        check(api.GroupEnd(), "ncclGroupEnd");
        ws.sync();
        run.keys.reset(&ws, 0); run.counts.reset(&ws, 0); run.m = 0;
    }
}

}  // namespace gsb
