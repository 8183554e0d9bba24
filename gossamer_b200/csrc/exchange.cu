// exchange.cu -- multi-GPU step of the counting phase (one process per GPU).
//
// The reference has no distributed mode; its scale-out advice is "build
// parts, then merge-graphs" (docs/goss.md:315-321).  Here every rank counts its own share of
// the reads, the locally reduced (key,count) runs are range-partitioned by splitters taken
// from a sample of all ranks' keys, exchanged with ONE all-to-all (ncclSend/ncclRecv pairs in
// a group, NVLink 5 / NVSwitch underneath) and merged, so that rank r ends up owning the r-th
// contiguous slice of the global sorted edge set.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch already loaded when the
// host program is a torchrun rank), so single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>

#include "exchange.h"

namespace gsb {

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

static const int kMaxRanks = 32;
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

// pass 1: how many keys go to each destination
template <typename K>
__global__ void __launch_bounds__(kPartThreads) dest_count_kernel(const K* __restrict__ keys, u64 n, Splitters sp, u64* __restrict__ totals) {
    __shared__ u32 cnt_s[kMaxRanks];
    if (threadIdx.x < kMaxRanks) cnt_s[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (u64 base = (u64)blockIdx.x * kPartThreads * kPartItems; base < n; base += (u64)gridDim.x * kPartThreads * kPartItems) {
#pragma unroll
        for (int i = 0; i < kPartItems; ++i) {
            const u64 idx = base + (u64)i * kPartThreads + threadIdx.x;
            const int d = idx < n ? dest_of(keys[idx], sp) : -1;
            for (int r = 0; r <= sp.n; ++r) {
                const u32 b = __ballot_sync(0xffffffffu, d == r);
                if (lane == 0 && b) atomicAdd(&cnt_s[r], (u32)__popc(b));
            }
        }
    }
    __syncthreads();
    if (threadIdx.x <= sp.n && cnt_s[threadIdx.x]) atomicAdd(&totals[threadIdx.x], (u64)cnt_s[threadIdx.x]);
}

// pass 2: scatter into per-destination regions (cursor[r] starts at the region offset); order inside
// a region is arbitrary -- the receiver sorts
template <typename K>
__global__ void __launch_bounds__(kPartThreads) dest_scatter_kernel(const K* __restrict__ keys, u64 n, Splitters sp, u64* __restrict__ cursor,
                                                                   K* __restrict__ out) {
    __shared__ u32 cnt_s[kMaxRanks];
    __shared__ u64 base_s[kMaxRanks];
    const int lane = threadIdx.x & 31;
    for (u64 base = (u64)blockIdx.x * kPartThreads * kPartItems; base < n; base += (u64)gridDim.x * kPartThreads * kPartItems) {
        if (threadIdx.x < kMaxRanks) cnt_s[threadIdx.x] = 0;
        __syncthreads();
        K k[kPartItems]; int d[kPartItems]; u32 slot[kPartItems];
#pragma unroll
        for (int i = 0; i < kPartItems; ++i) {
            const u64 idx = base + (u64)i * kPartThreads + threadIdx.x;
            d[i] = -1; slot[i] = 0;
            if (idx < n) { k[i] = keys[idx]; d[i] = dest_of(k[i], sp); }
            for (int r = 0; r <= sp.n; ++r) {
                const u32 b = __ballot_sync(0xffffffffu, d[i] == r);
                u32 w = 0;
                if (lane == 0 && b) w = atomicAdd(&cnt_s[r], (u32)__popc(b));
                w = __shfl_sync(0xffffffffu, w, 0);
                if (d[i] == r) slot[i] = w + __popc(b & ((1u << lane) - 1));
            }
        }
        __syncthreads();
        if (threadIdx.x <= sp.n) base_s[threadIdx.x] = cnt_s[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], (u64)cnt_s[threadIdx.x]) : 0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kPartItems; ++i)
            if (d[i] >= 0) out[base_s[d[i]] + slot[i]] = k[i];
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
    reduce_sorted(ws, key_bytes, where ? rkeys_alt.p : rkeys.p, where ? rcounts_alt.p : rcounts.p, total, 1, merged, &distinct, nullptr);
    run = std::move(merged);
}

// Range-partition the raw instance keys by sampled splitters and exchange them (one all-to-all):
// afterwards `recv` holds every instance, from all ranks, whose key lies in this rank's range.
void exchange_instances(Exchange* x, Workspace& ws, int key_bytes, const void* keys, u64 n_keys, u8* parted_buf /* n_keys keys */,
                        DevBuf<u8>& recv, u64* recv_cap, u64* n_recv, ExchangeTiming* timing) {
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
    struct { u8* p; } parted{parted_buf};
    if (n_keys) {
        if (key_bytes == 8) dest_scatter_kernel<u64><<<grid, kPartThreads, 0, s>>>((const u64*)keys, n_keys, sp, totals.p + kMaxRanks, (u64*)parted.p);
        else dest_scatter_kernel<Key128><<<grid, kPartThreads, 0, s>>>((const Key128*)keys, n_keys, sp, totals.p + kMaxRanks, (Key128*)parted.p);
        ++ws.launches;
    }
    // counts matrix
    DevBuf<u64> cnt_mine(&ws, n), cnt_all(&ws, (size_t)n * n);
    GSB_CUDA_TRY(cudaMemcpyAsync(cnt_mine.p, send_cnt.data(), (size_t)n * 8, cudaMemcpyHostToDevice, s));
    check(api.AllGather(cnt_mine.p, cnt_all.p, n, ncclUint64, x->comm, s), "ncclAllGather(counts)");
    std::vector<u64> cnt((size_t)n * n);
    GSB_CUDA_TRY(cudaMemcpyAsync(cnt.data(), cnt_all.p, cnt.size() * 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    std::vector<u64> recv_cnt(n), recv_off(n + 1, 0);
    for (int r = 0; r < n; ++r) { recv_cnt[r] = cnt[(size_t)r * n + x->rank]; recv_off[r + 1] = recv_off[r] + recv_cnt[r]; }
    const u64 total = recv_off[n];
    if (total > *recv_cap) {                                      // grow-only: steady-state steps allocate nothing
        recv.free();
        *recv_cap = total + total / 16 + 1024;
        recv.reset(&ws, *recv_cap * key_bytes);
    }
    cudaEvent_t e0, e1;
    GSB_CUDA_TRY(cudaEventCreate(&e0)); GSB_CUDA_TRY(cudaEventCreate(&e1));
    GSB_CUDA_TRY(cudaEventRecord(e0, s));
    check(api.GroupStart(), "ncclGroupStart");
    u64 sent_remote = 0;
    for (int r = 0; r < n; ++r) {
        if (send_cnt[r]) check(api.Send(parted.p + send_off[r] * key_bytes, send_cnt[r] * key_bytes, ncclUint8, r, x->comm, s), "ncclSend(instances)");
        if (recv_cnt[r]) check(api.Recv(recv.p + recv_off[r] * key_bytes, recv_cnt[r] * key_bytes, ncclUint8, r, x->comm, s), "ncclRecv(instances)");
        if (r != x->rank) sent_remote += send_cnt[r] * key_bytes;
    }
    check(api.GroupEnd(), "ncclGroupEnd");
    GSB_CUDA_TRY(cudaEventRecord(e1, s));
    GSB_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    GSB_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (timing) { timing->ms_all_to_all += ms; timing->bytes_sent_remote += sent_remote; }
    *n_recv = total;
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
