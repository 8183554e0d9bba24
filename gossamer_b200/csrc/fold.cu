// fold.cu -- strand folding for build-graph: restore both strands after counting.
//
// The reference feeds every (k+1)-mer window x AND its reverse complement rc(x) to the counter
// (src/ReverseComplementAdapter.hh:34-55, src/GossCmdBuildGraph.cc:318-343), so in its multiset
//     count(y) = count(rc y) = #windows whose canonical key min(x, rc x) equals min(y, rc y),
// doubled when y == rc(y) (one window then contributes the same key twice).  The extraction kernel
// therefore emits only the canonical key of each window; sort and run-length reduce run over HALF the
// instances, the min-count filter is applied to the (doubled where needed) counts, and this file
// produces the reference's full edge set from the folded, filtered run C:
//     R = { (rc y, count) : (y, count) in C, y != rc y }     -- disjoint from C: y < rc y is not canonical
//     sort R by key (radix sort carrying the counts)
//     merge C and R (merge path: one binary search per output tile, shared-memory merge inside the tile)
#include "kernels.h"

namespace gsb {

namespace {

static const int kFoldThreads = 256;
static const int kFoldItems = 4;

// counts of self-complementary keys *= 2; their number is added to *n_self
template <typename K>
__global__ void __launch_bounds__(kFoldThreads) double_self_rc_kernel(const K* __restrict__ keys, u64* __restrict__ counts, u64 m, int w, u64* __restrict__ n_self) {
    u32 mine = 0;
    for (u64 i = (u64)blockIdx.x * kFoldThreads + threadIdx.x; i < m; i += (u64)gridDim.x * kFoldThreads) {
        const K y = keys[i];
        if (KeyOps<K>::eq(key_rc(y, w), y)) { counts[i] <<= 1; ++mine; }
    }
    const u32 any = __ballot_sync(0xffffffffu, mine != 0);
    if (any) {
#pragma unroll
        for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(n_self, (u64)mine);
    }
}

// R = reverse complements of the keys that are not self-complementary, with their counts; order is
// arbitrary (warp-aggregated append) -- R is sorted next and its keys are distinct, so the result is
// deterministic.
template <typename K>
__global__ void __launch_bounds__(kFoldThreads) unfold_rc_kernel(const K* __restrict__ keys, const u64* __restrict__ counts, u64 m, int w,
                                                                K* __restrict__ out_keys, u64* __restrict__ out_counts, u64* __restrict__ cursor) {
    const int lane = threadIdx.x & 31;
    const u32 lt = (1u << lane) - 1;
    const u64 tile = (u64)blockIdx.x * kFoldThreads * kFoldItems;
    K r[kFoldItems]; u64 c[kFoldItems]; u32 bal[kFoldItems]; u32 total = 0;
#pragma unroll
    for (int it = 0; it < kFoldItems; ++it) {
        const u64 i = tile + (u64)it * kFoldThreads + threadIdx.x;
        bool take = false;
        r[it] = KeyOps<K>::make(0, 0); c[it] = 0;
        if (i < m) {
            const K y = keys[i];
            r[it] = key_rc(y, w);
            take = !KeyOps<K>::eq(r[it], y);
            c[it] = counts[i];
        }
        bal[it] = __ballot_sync(0xffffffffu, take);
        total += __popc(bal[it]);
    }
    u64 base = 0;
    if (lane == 0 && total) base = atomicAdd(cursor, (u64)total);
    base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
    for (int it = 0; it < kFoldItems; ++it) {
        if ((bal[it] >> lane) & 1u) {
            const u64 o = base + __popc(bal[it] & lt);
            out_keys[o] = r[it];
            out_counts[o] = c[it];
        }
        base += __popc(bal[it]);
    }
}

// ---- merge path ------------------------------------------------------------------------------
static const int kMgThreads = 256;
template <typename K> struct MergeShape { static const int kItems = 8; };
template <> struct MergeShape<Key128> { static const int kItems = 4; };

// split[t] = how many of the first t * TILE outputs come from run a (on equal keys b goes first;
// the runs merged here have disjoint key sets)
template <typename K>
__global__ void merge_partition_kernel(const K* __restrict__ a, u64 na, const K* __restrict__ b, u64 nb, u64* __restrict__ split, u64 tiles) {
    constexpr u64 TILE = (u64)kMgThreads * MergeShape<K>::kItems;
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > tiles) return;
    const u64 total = na + nb;
    const u64 d = t * TILE < total ? t * TILE : total;
    u64 lo = d > nb ? d - nb : 0, hi = d < na ? d : na;
    while (lo < hi) {
        const u64 mid = lo + ((hi - lo) >> 1);
        if (KeyOps<K>::lt(a[mid], b[d - 1 - mid])) lo = mid + 1; else hi = mid;
    }
    split[t] = lo;
}

template <typename K>
__global__ void __launch_bounds__(kMgThreads) merge_kernel(const K* __restrict__ a, const u64* __restrict__ ca, u64 na,
                                                          const K* __restrict__ b, const u64* __restrict__ cb, u64 nb,
                                                          const u64* __restrict__ split, K* __restrict__ out_keys, u64* __restrict__ out_counts) {
    typedef KeyOps<K> KO;
    constexpr int ITEMS = MergeShape<K>::kItems;
    constexpr u32 TILE = kMgThreads * ITEMS;
    __shared__ K sk[TILE];
    __shared__ u64 sc[TILE];
    const u64 total = na + nb;
    const u64 d0 = (u64)blockIdx.x * TILE;
    const u64 d1 = d0 + TILE < total ? d0 + TILE : total;
    const u64 a0 = split[blockIdx.x], a1 = split[blockIdx.x + 1];
    const u64 b0 = d0 - a0, b1 = d1 - a1;
    const u32 la = (u32)(a1 - a0), lb = (u32)(b1 - b0);
    for (u32 i = threadIdx.x; i < la; i += kMgThreads) { sk[i] = a[a0 + i]; sc[i] = ca[a0 + i]; }
    for (u32 i = threadIdx.x; i < lb; i += kMgThreads) { sk[la + i] = b[b0 + i]; sc[la + i] = cb[b0 + i]; }
    __syncthreads();
    const u32 n = la + lb;
    const u32 diag = threadIdx.x * ITEMS < n ? threadIdx.x * ITEMS : n;
    u32 lo = diag > lb ? diag - lb : 0, hi = diag < la ? diag : la;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (KO::lt(sk[mid], sk[la + diag - 1 - mid])) lo = mid + 1; else hi = mid;
    }
    u32 i = lo, j = diag - lo;
#pragma unroll
    for (int s = 0; s < ITEMS; ++s) {
        const u32 o = diag + s;
        if (o < n) {
            const bool from_a = j >= lb || (i < la && KO::lt(sk[i], sk[la + j]));
            const u32 src = from_a ? i : la + j;
            out_keys[d0 + o] = sk[src];
            out_counts[d0 + o] = sc[src];
            if (from_a) ++i; else ++j;
        }
    }
}

template <typename K>
void merge_typed(Workspace& ws, const void* ka, const u64* ca, u64 na, const void* kb, const u64* cb, u64 nb, void* out_keys, u64* out_counts) {
    constexpr u64 TILE = (u64)kMgThreads * MergeShape<K>::kItems;
    const u64 total = na + nb;
    if (!total) return;
    const u64 tiles = (total + TILE - 1) / TILE;
    DevBuf<u64> split(&ws, tiles + 1);
    merge_partition_kernel<K><<<(unsigned)((tiles + 1 + 127) / 128), 128, 0, ws.stream>>>((const K*)ka, na, (const K*)kb, nb, split.p, tiles);
    merge_kernel<K><<<(unsigned)tiles, kMgThreads, 0, ws.stream>>>((const K*)ka, ca, na, (const K*)kb, cb, nb, split.p, (K*)out_keys, out_counts);
    ws.launches += 2;
}

}  // namespace

void merge_disjoint_runs(Workspace& ws, int key_bytes, const void* ka, const u64* ca, u64 na, const void* kb, const u64* cb, u64 nb,
                         void* out_keys, u64* out_counts) {
    if (key_bytes == 8) merge_typed<u64>(ws, ka, ca, na, kb, cb, nb, out_keys, out_counts);
    else merge_typed<Key128>(ws, ka, ca, na, kb, cb, nb, out_keys, out_counts);
}

u64 fold_double_self_rc(Workspace& ws, int key_bytes, int w, const void* keys, u64* counts, u64 m) {
    if (!m) return 0;
    cudaStream_t s = ws.stream;
    DevBuf<u64> n_self(&ws, 1);
    GSB_CUDA_TRY(cudaMemsetAsync(n_self.p, 0, 8, s));
    const u64 blocks = (m + kFoldThreads - 1) / kFoldThreads;
    const int grid = (int)(blocks < (u64)ws.sm_count * 8 ? blocks : (u64)ws.sm_count * 8);
    if (key_bytes == 8) double_self_rc_kernel<u64><<<grid, kFoldThreads, 0, s>>>((const u64*)keys, counts, m, w, n_self.p);
    else double_self_rc_kernel<Key128><<<grid, kFoldThreads, 0, s>>>((const Key128*)keys, counts, m, w, n_self.p);
    ++ws.launches;
    u64 host = 0;
    GSB_CUDA_TRY(cudaMemcpyAsync(&host, n_self.p, 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    return host;
}

u64 unfold_append_rc(Workspace& ws, int key_bytes, int w, const void* keys, const u64* counts, u64 m, void* out_keys, u64* out_counts) {
    if (!m) return 0;
    cudaStream_t s = ws.stream;
    DevBuf<u64> cursor(&ws, 1);
    GSB_CUDA_TRY(cudaMemsetAsync(cursor.p, 0, 8, s));
    const unsigned tiles = (unsigned)((m + kFoldThreads * kFoldItems - 1) / (kFoldThreads * kFoldItems));
    if (key_bytes == 8) unfold_rc_kernel<u64><<<tiles, kFoldThreads, 0, s>>>((const u64*)keys, counts, m, w, (u64*)out_keys, out_counts, cursor.p);
    else unfold_rc_kernel<Key128><<<tiles, kFoldThreads, 0, s>>>((const Key128*)keys, counts, m, w, (Key128*)out_keys, out_counts, cursor.p);
    ++ws.launches;
    u64 n_rc = 0;
    GSB_CUDA_TRY(cudaMemcpyAsync(&n_rc, cursor.p, 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    return n_rc;
}

void unfold_run(Workspace& ws, int key_bytes, int key_bits, int w, ReducedRun& run, bool sorted_input) {
    const u64 m = run.m;
    if (!m) return;
    cudaStream_t s = ws.stream;
    if (!sorted_input) {
        // U = run ++ rc(run) ordered by the full key: most-significant-digit passes + a shared-memory sort per bucket
        // (partition.cu); only data that defeat the bucket geometry take the radix sort below
        {
            ReducedRun sorted;
            if (sort_pairs_msd(ws, key_bytes, key_bits, run.keys.p, run.counts.p, m, w, sorted)) { run = std::move(sorted); return; }
        }
        // U = run ++ rc(run), then one radix sort of the pairs by the full key
        DevBuf<u8> uk(&ws, 2 * m * key_bytes), uk_alt(&ws, 2 * m * key_bytes);
        DevBuf<u64> uc(&ws, 2 * m), uc_alt(&ws, 2 * m);
        GSB_CUDA_TRY(cudaMemcpyAsync(uk.p, run.keys.p, m * key_bytes, cudaMemcpyDeviceToDevice, s));
        GSB_CUDA_TRY(cudaMemcpyAsync(uc.p, run.counts.p, m * 8, cudaMemcpyDeviceToDevice, s));
        const u64 n_rc = unfold_append_rc(ws, key_bytes, w, run.keys.p, run.counts.p, m, uk.p + m * key_bytes, uc.p + m);
        const int where = sort_keys(ws, key_bytes, key_bits, uk.p, uk_alt.p, uc.p, uc_alt.p, m + n_rc, nullptr, nullptr);
        run.keys = std::move(where ? uk_alt : uk);
        run.counts = std::move(where ? uc_alt : uc);
        run.m = m + n_rc;
        return;
    }
    DevBuf<u8> rk(&ws, m * key_bytes), rk_alt(&ws, m * key_bytes);
    DevBuf<u64> rc(&ws, m), rc_alt(&ws, m);
    const u64 n_rc = unfold_append_rc(ws, key_bytes, w, run.keys.p, run.counts.p, m, rk.p, rc.p);
    if (!n_rc) return;                                           // every key is its own reverse complement
    const int where = sort_keys(ws, key_bytes, key_bits, rk.p, rk_alt.p, rc.p, rc_alt.p, n_rc, nullptr, nullptr);
    DevBuf<u8> out_keys(&ws, (m + n_rc) * key_bytes);
    DevBuf<u64> out_counts(&ws, m + n_rc);
    merge_disjoint_runs(ws, key_bytes, run.keys.p, run.counts.p, m, where ? rk_alt.p : rk.p, where ? rc_alt.p : rc.p, n_rc, out_keys.p, out_counts.p);
    run.keys = std::move(out_keys);
    run.counts = std::move(out_counts);
    run.m = m + n_rc;
}

}  // namespace gsb
