// sort.cu -- counting by sorting: LSD radix sort (8-bit digits, one sweep per digit with
// decoupled look-back), run-length reduce, min-count filter.
//
// Replaces (result-wise) BackyardHash::insert + BackyardHash::sort/BlendedSort + the
// duplicate-merging emit loop (src/BackyardHash.cc:115-271, src/BlendedSort.hh:58-167,
// src/GossCmdBuildGraph.cc:239-258) and trim-graph's `count > C` predicate
// (src/GossCmdTrimGraph.cc:119).  A cuckoo hash is a random-access structure; on a GPU with
// 8 TB/s of streaming bandwidth the same multiset is counted faster by sorting the instances
// and measuring run lengths, and the sorted order is what the succinct writers need anyway.
#include "kernels.h"
#include "scan.cuh"

namespace gsb {

// ------------------------------------------------------------------------------------------
// digit histograms (standalone; the extraction kernel fuses the same thing for fresh keys)
// ------------------------------------------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(256) digit_hist_kernel(const K* __restrict__ keys, u64 n, int passes, u64* __restrict__ hist) {
    extern __shared__ u32 hist_s[];
    for (int i = threadIdx.x; i < passes * 256; i += 256) hist_s[i] = 0;
    __syncthreads();
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
        K k = keys[i];
        for (int d = 0; d < passes; ++d) atomicAdd(&hist_s[d * 256 + KeyOps<K>::digit(k, 8 * d)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += 256) {
        u32 c = hist_s[i];
        if (c) atomicAdd(&hist[i], (u64)c);
    }
}

// base[p][d] = number of keys whose digit p is < d
__global__ void __launch_bounds__(256) digit_base_kernel(const u64* __restrict__ hist, u64* __restrict__ base) {
    __shared__ u64 sm[256 / 32 + 1];
    u64 v = hist[blockIdx.x * 256 + threadIdx.x];
    u64 ex = block_exclusive_scan<u64, 256>(v, (u64*)nullptr, sm);
    base[blockIdx.x * 256 + threadIdx.x] = ex;
}

// ------------------------------------------------------------------------------------------
// one radix pass
// ------------------------------------------------------------------------------------------
// Look back over the predecessors of `tile` for one digit: W states are fetched per round trip
// (independent loads), so a chain of tiles that have only published aggregates costs one L2
// latency per W tiles instead of one per tile.
template <typename LB, int W>
__device__ __forceinline__ LB lookback_window(const LB* __restrict__ bin_states /* &lookback[digit] */, u32 tile) {
    const int S = LookbackWord<LB>::kShift;
    const LB M = LookbackWord<LB>::kMask;
    LB excl = 0;
    long long t = (long long)tile - 1;
    for (;;) {
        LB v[W];
#pragma unroll
        for (int i = 0; i < W; ++i) v[i] = (t - i >= 0) ? ld_volatile(bin_states + (size_t)(t - i) * 256) : (LB)((LB)2 << S);
        int used = W;
#pragma unroll
        for (int i = 0; i < W; ++i) {
            if (i < used) {
                const LB st = v[i] >> S;
                if (st == 0) used = i;                          // not published yet: poll again from here
                else {
                    excl += v[i] & M;
                    if (st == 2) return excl;
                }
            }
        }
        t -= used;
    }
}

// One 8-bit radix sweep over a tile.  Two schedules, both kept because ncu favours one or the
// other depending on the tile shape (see profiles/):
//   MODE 0  count digits first (shared atomics) -> publish the tile aggregate -> rank -> look back -> write
//   MODE 1  rank and count in one go (ballots + plain LDS/STS on the warp's counters, no atomics)
//           -> publish -> scatter -> look back -> write
template <typename K, typename LB, int THREADS, int ITEMS, bool HAS_VALUES, int MINB, int MODE>
__global__ void __launch_bounds__(THREADS, MINB) onesweep_kernel(const K* __restrict__ in, K* __restrict__ out,
                                                                 const u64* __restrict__ vin, u64* __restrict__ vout,
                                                                 u64 n, int shift, const u64* __restrict__ digit_base,
                                                                 LB* lookback, u32* ticket, int ablate) {
    typedef KeyOps<K> KO;
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * ITEMS;
    const int S = LookbackWord<LB>::kShift;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K* keys_s = reinterpret_cast<K*>(smem_raw);                                   // [TILE]
    u32* warp_ofs = reinterpret_cast<u32*>(smem_raw + (size_t)TILE * sizeof(K));   // [WARPS][256] counts, then running offsets
    u64* gofs = reinterpret_cast<u64*>(warp_ofs + WARPS * 256);                    // [256]
    u32* vpos_s = reinterpret_cast<u32*>(gofs + 256);                              // [TILE] only with values: source slot of each sorted key
    __shared__ u32 tile_s;
    __shared__ u32 scan_s[THREADS / 32 + 1];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < WARPS * 256; i += THREADS) warp_ofs[i] = 0;
    __syncthreads();
    const u32 tile = tile_s;
    const u64 base = (u64)tile * TILE;
    const u32 tile_n = (u32)((n - base) < (u64)TILE ? (n - base) : (u64)TILE);
    const u32 lt_mask = (1u << lane) - 1;

    K key[ITEMS];
    u16 rank[MODE == 1 ? ITEMS : 1];
    const u32 wbase = warp * 32 * ITEMS + lane;
    u32* my_ofs = warp_ofs + warp * 256;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const u32 idx = wbase + i * 32;
        key[i] = idx < tile_n ? in[base + idx] : KO::make(0, 0);
    }

    // peer set of this lane for item i: lanes (with a key) whose digit equals mine, from 8 ballots
    auto peer_set = [&](u32 d, bool ok) -> u32 {
        u32 peers = __ballot_sync(0xffffffffu, ok);
        if (ablate & 2) return ok ? (1u << lane) : 0u;          // profiling only: wrong ranks, no ballots
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const bool bit = (d >> b) & 1u;
            const u32 bal = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? bal : ~bal;
        }
        return peers;
    };

    if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (wbase + i * 32 < tile_n) atomicAdd(&my_ofs[KO::digit(key[i], shift)], 1u);
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const bool ok = wbase + i * 32 < tile_n;
            const u32 d = KO::digit(key[i], shift);
            const u32 peers = peer_set(d, ok);
            u32 before = 0;
            if (ok) before = my_ofs[d];
            __syncwarp();
            if (ok && (peers & lt_mask) == 0) my_ofs[d] = before + __popc(peers);
            __syncwarp();
            rank[i] = (u16)(before + __popc(peers & lt_mask));
        }
    }
    __syncthreads();

    // per digit: exclusive offsets over warps, tile count -> publish the aggregate
    u32 count = 0;
    if (threadIdx.x < 256) {
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const u32 c = warp_ofs[w * 256 + threadIdx.x];
            warp_ofs[w * 256 + threadIdx.x] = run;
            run += c;
        }
        count = run;
    }
    const u32 bstart = block_exclusive_scan<u32, THREADS>(count, (u32*)nullptr, scan_s);
    LB* my_state = lookback + (size_t)tile * 256 + threadIdx.x;
    if (threadIdx.x < 256) {
        st_volatile(my_state, (LB)(((LB)(tile == 0 ? 2 : 1) << S) | (LB)count));
#pragma unroll
        for (int w = 0; w < WARPS; ++w) warp_ofs[w * 256 + threadIdx.x] += bstart;   // slot in the exchange buffer
    }
    __syncthreads();

    // keys into the exchange buffer, grouped by digit, input order kept inside a digit
    if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const bool ok = wbase + i * 32 < tile_n;
            const u32 d = KO::digit(key[i], shift);
            const u32 peers = peer_set(d, ok);
            const int leader = __ffs(peers) - 1;                // lowest lane of the group (ok lanes only)
            u32 before = 0;
            if (ok && lane == leader) before = atomicAdd(&my_ofs[d], (u32)__popc(peers));
            before = __shfl_sync(0xffffffffu, before, leader < 0 ? lane : leader);
            if (ok) {
                const u32 pos = before + __popc(peers & lt_mask);
                keys_s[pos] = key[i];
                if (HAS_VALUES) vpos_s[pos] = wbase + i * 32;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            if (wbase + i * 32 < tile_n) {
                const u32 pos = my_ofs[KO::digit(key[i], shift)] + rank[i];
                keys_s[pos] = key[i];
                if (HAS_VALUES) vpos_s[pos] = wbase + i * 32;
            }
        }
    }

    // exclusive prefix of this tile's digits over all earlier tiles
    if (threadIdx.x < 256) {
        LB excl = 0;
        if (tile != 0 && !(ablate & 1)) {
            excl = lookback_window<LB, 4>(lookback + threadIdx.x, tile);
            st_volatile(my_state, (LB)(((LB)2 << S) | (excl + (LB)count)));
        }
        gofs[threadIdx.x] = digit_base[threadIdx.x] + (u64)excl - (u64)bstart;
    }
    __syncthreads();

    // coalesced write-out, one contiguous run per digit
    if (ablate & 4) return;
    for (u32 j = threadIdx.x; j < tile_n; j += THREADS) {
        const K k = keys_s[j];
        const u64 dst = (ablate & 1) ? (base + j) : gofs[KO::digit(k, shift)] + j;
        out[dst] = k;
        if (HAS_VALUES) vout[dst] = vin[base + vpos_s[j]];
    }
}

extern int g_sort_ablate_fwd;
template <typename K, typename LB, int THREADS, int ITEMS, bool HAS_VALUES, int MINB, int MODE>
static void launch_onesweep(const void* in, void* out, const u64* vin, u64* vout, u64 n, int shift, const u64* digit_base,
                            void* lookback, u32* ticket, cudaStream_t s) {
    constexpr int TILE = THREADS * ITEMS;
    const size_t smem = (size_t)TILE * sizeof(K) + (size_t)(THREADS / 32) * 256 * 4 + 256 * 8 + (HAS_VALUES ? (size_t)TILE * 4 : 0);
    static bool configured = false;
    auto kern = onesweep_kernel<K, LB, THREADS, ITEMS, HAS_VALUES, MINB, MODE>;
    if (!configured) {
        GSB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GSB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        configured = true;
    }
    const u64 tiles = (n + TILE - 1) / TILE;
    kern<<<(unsigned)tiles, THREADS, smem, s>>>((const K*)in, (K*)out, vin, vout, n, shift, digit_base, (LB*)lookback, ticket, g_sort_ablate_fwd);
}

// Tile shapes.  g_sort_tuning picks among the 64-bit-key variants (set through gsb_debug_set_tuning
// while profiling; the default is the one ncu favoured).
struct SortShape { int threads, items; };
static int g_sort_tuning = 0;
int g_sort_ablate_fwd = 0;
#define g_sort_ablate g_sort_ablate_fwd
// profiling only (gsb_debug_set_tuning(id | ablate << 8)): results are wrong when non-zero
static const SortShape kShapes64[] = {{256, 16}, {256, 16}, {256, 16}, {384, 16}, {512, 8}, {256, 12}, {512, 12}, {256, 12}};
static const int kNumShapes64 = (int)(sizeof(kShapes64) / sizeof(kShapes64[0]));
static const SortShape kShape128 = {256, 8};

void sort_set_tuning(int id) {
    g_sort_ablate = (id >> 8) & 0xFF;
    id &= 0xFF;
    g_sort_tuning = (id >= 0 && id < kNumShapes64) ? id : 0;
}

u64 sort_tile_keys(int key_bytes) {
    const SortShape sh = key_bytes == 8 ? kShapes64[g_sort_tuning] : kShape128;
    return (u64)sh.threads * sh.items;
}

// bytes of look-back state needed for n keys (+ the ticket word at the end)
u64 sort_lookback_bytes(int key_bytes, u64 n) {
    u64 tiles = (n + sort_tile_keys(key_bytes) - 1) / sort_tile_keys(key_bytes);
    u64 word = n < (1ull << 30) ? 4 : 8;
    return tiles * 256 * word + 256;
}

void sort_digit_hist(int key_bytes, const void* keys, u64 n, int passes, u64* hist, int sm_count, cudaStream_t s, u64* launches) {
    if (!n) return;
    u64 blocks = (n + 255) / 256;
    int grid = (int)(blocks < (u64)sm_count * 8 ? blocks : (u64)sm_count * 8);
    size_t smem = (size_t)passes * 256 * 4;
    if (key_bytes == 8) digit_hist_kernel<u64><<<grid, 256, smem, s>>>((const u64*)keys, n, passes, hist);
    else digit_hist_kernel<Key128><<<grid, 256, smem, s>>>((const Key128*)keys, n, passes, hist);
    ++*launches;
}

void sort_digit_base(const u64* hist, u64* base, int passes, cudaStream_t s, u64* launches) {
    digit_base_kernel<<<passes, 256, 0, s>>>(hist, base);
    ++*launches;
}

// One pass over digit `pass` (bits [8*pass, 8*pass+8)).  lookback must hold sort_lookback_bytes().
void sort_pass(int key_bytes, const void* in, void* out, const u64* vin, u64* vout, u64 n, int pass, const u64* digit_base_all,
               void* lookback, cudaStream_t s, u64* launches, cudaEvent_t ev_begin, cudaEvent_t ev_end) {
    if (!n) return;
    const u64 lb_bytes = sort_lookback_bytes(key_bytes, n);
    GSB_CUDA_TRY(cudaMemsetAsync(lookback, 0, lb_bytes, s));
    if (ev_begin) GSB_CUDA_TRY(cudaEventRecord(ev_begin, s));
    u32* ticket = (u32*)((char*)lookback + lb_bytes - 256);
    const u64* db = digit_base_all + (size_t)pass * 256;
    const int shift = 8 * pass;
    const bool small = n < (1ull << 30);
    const bool hv = vin != nullptr;
#define GSB_LAUNCH(K, T, I, MB, MD)                                                                               \
    do {                                                                                                          \
        if (small && !hv) launch_onesweep<K, u32, T, I, false, MB, MD>(in, out, vin, vout, n, shift, db, lookback, ticket, s); \
        else if (small && hv) launch_onesweep<K, u32, T, I, true, (MB > 2 ? 2 : MB), MD>(in, out, vin, vout, n, shift, db, lookback, ticket, s); \
        else if (!hv) launch_onesweep<K, u64, T, I, false, MB, MD>(in, out, vin, vout, n, shift, db, lookback, ticket, s); \
        else launch_onesweep<K, u64, T, I, true, (MB > 2 ? 2 : MB), MD>(in, out, vin, vout, n, shift, db, lookback, ticket, s); \
    } while (0)
    if (key_bytes == 8) {
        switch (g_sort_tuning) {
            default: GSB_LAUNCH(u64, 256, 16, 4, 0); break;
            case 1: GSB_LAUNCH(u64, 256, 16, 3, 1); break;
            case 2: GSB_LAUNCH(u64, 256, 16, 4, 1); break;
            case 3: GSB_LAUNCH(u64, 384, 16, 2, 1); break;
            case 4: GSB_LAUNCH(u64, 512, 8, 2, 1); break;
            case 5: GSB_LAUNCH(u64, 256, 12, 4, 1); break;
            case 6: GSB_LAUNCH(u64, 512, 12, 2, 0); break;
            case 7: GSB_LAUNCH(u64, 256, 12, 4, 0); break;
        }
    } else {
        GSB_LAUNCH(Key128, 256, 8, 3, 0);
    }
#undef GSB_LAUNCH
    if (ev_end) GSB_CUDA_TRY(cudaEventRecord(ev_end, s));
    ++*launches;
}

// ------------------------------------------------------------------------------------------
// run-length reduce: sorted keys -> distinct keys + index of the first instance of each
// ------------------------------------------------------------------------------------------
static const int kRleThreads = 256;
static const int kRleItems = 8;

template <typename K>
__global__ void __launch_bounds__(kRleThreads) rle_kernel(const K* __restrict__ keys, u64 n, K* __restrict__ out_keys, u64* __restrict__ out_pos,
                                                          u64* lookback, u32* ticket, u64* __restrict__ total_out) {
    typedef KeyOps<K> KO;
    __shared__ u32 tile_s;
    __shared__ u64 prefix_s;
    __shared__ u32 scan_s[kRleThreads / 32 + 1];
    if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 tile = tile_s;
    const u64 base = (u64)tile * (kRleThreads * kRleItems) + (u64)threadIdx.x * kRleItems;
    K k[kRleItems];
    bool head[kRleItems];
    u32 cnt = 0;
    K prev = KO::make(0, 0);
    if (base > 0 && base < n) prev = keys[base - 1];
#pragma unroll
    for (int i = 0; i < kRleItems; ++i) {
        const u64 idx = base + i;
        head[i] = false;
        if (idx < n) {
            k[i] = keys[idx];
            head[i] = idx == 0 || !KO::eq(k[i], prev);
            prev = k[i];
            cnt += head[i];
        }
    }
    u32 tile_total;
    u32 ex = block_exclusive_scan<u32, kRleThreads>(cnt, &tile_total, scan_s);
    if (threadIdx.x == 0) {
        u64 p = lookback_exclusive<u64>(lookback, 1u, tile, 0u, (u64)tile_total);
        prefix_s = p;
        if ((u64)(tile + 1) * (kRleThreads * kRleItems) >= n) {           // last tile
            *total_out = p + tile_total;
            out_pos[p + tile_total] = n;                                     // sentinel: count_j = pos[j+1] - pos[j]
        }
    }
    __syncthreads();
    u64 j = prefix_s + ex;
#pragma unroll
    for (int i = 0; i < kRleItems; ++i)
        if (head[i]) { out_keys[j] = k[i]; out_pos[j] = base + i; ++j; }
}

u64 rle_lookback_bytes(u64 n) { return ((n + kRleThreads * kRleItems - 1) / (kRleThreads * kRleItems)) * 8 + 256; }

void sort_rle(int key_bytes, const void* keys, u64 n, void* out_keys, u64* out_pos, void* lookback, u64* total_dev, cudaStream_t s, u64* launches) {
    if (!n) {
        GSB_CUDA_TRY(cudaMemsetAsync(total_dev, 0, 8, s));
        GSB_CUDA_TRY(cudaMemsetAsync(out_pos, 0, 8, s));
        return;
    }
    const u64 lb = rle_lookback_bytes(n);
    GSB_CUDA_TRY(cudaMemsetAsync(lookback, 0, lb, s));
    u32* ticket = (u32*)((char*)lookback + lb - 256);
    const u64 tiles = (n + kRleThreads * kRleItems - 1) / (kRleThreads * kRleItems);
    if (key_bytes == 8) rle_kernel<u64><<<(unsigned)tiles, kRleThreads, 0, s>>>((const u64*)keys, n, (u64*)out_keys, out_pos, (u64*)lookback, ticket, total_dev);
    else rle_kernel<Key128><<<(unsigned)tiles, kRleThreads, 0, s>>>((const Key128*)keys, n, (Key128*)out_keys, out_pos, (u64*)lookback, ticket, total_dev);
    ++*launches;
}

// counts[j] = pos[j+1] - pos[j]              (fresh instances: every instance weighs 1)
// counts[j] = csum[pos[j+1]] - csum[pos[j]]  (merging reduced runs: csum = exclusive scan of the weights, n+1 entries)
__global__ void counts_from_pos_kernel(const u64* __restrict__ pos, const u64* __restrict__ csum, u64 m, u64* __restrict__ counts) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += (u64)gridDim.x * blockDim.x) {
        u64 a = pos[j], b = pos[j + 1];
        counts[j] = csum ? (csum[b] - csum[a]) : (b - a);
    }
}

void sort_counts_from_pos(const u64* pos, const u64* csum, u64 m, u64* counts, cudaStream_t s, u64* launches) {
    if (!m) return;
    u64 blocks = (m + 255) / 256;
    counts_from_pos_kernel<<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, s>>>(pos, csum, m, counts);
    ++*launches;
}

// ------------------------------------------------------------------------------------------
// min-count filter: keep (key,count) with count >= min_count, order preserved
// ------------------------------------------------------------------------------------------
template <typename K>
__global__ void __launch_bounds__(kRleThreads) filter_kernel(const K* __restrict__ keys, const u64* __restrict__ counts, u64 m, u64 min_count,
                                                             K* __restrict__ out_keys, u64* __restrict__ out_counts,
                                                             u64* lookback, u32* ticket, u64* __restrict__ total_out) {
    __shared__ u32 tile_s;
    __shared__ u64 prefix_s;
    __shared__ u32 scan_s[kRleThreads / 32 + 1];
    if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 tile = tile_s;
    const u64 base = (u64)tile * (kRleThreads * kRleItems) + (u64)threadIdx.x * kRleItems;
    u64 c[kRleItems];
    u32 cnt = 0;
#pragma unroll
    for (int i = 0; i < kRleItems; ++i) {
        const u64 idx = base + i;
        c[i] = idx < m ? counts[idx] : 0;
        if (idx < m && c[i] >= min_count) ++cnt;
    }
    u32 tile_total;
    u32 ex = block_exclusive_scan<u32, kRleThreads>(cnt, &tile_total, scan_s);
    if (threadIdx.x == 0) {
        u64 p = lookback_exclusive<u64>(lookback, 1u, tile, 0u, (u64)tile_total);
        prefix_s = p;
        if ((u64)(tile + 1) * (kRleThreads * kRleItems) >= m) *total_out = p + tile_total;
    }
    __syncthreads();
    u64 j = prefix_s + ex;
#pragma unroll
    for (int i = 0; i < kRleItems; ++i) {
        const u64 idx = base + i;
        if (idx < m && c[i] >= min_count) { out_keys[j] = keys[idx]; out_counts[j] = c[i]; ++j; }
    }
}

void sort_filter(int key_bytes, const void* keys, const u64* counts, u64 m, u64 min_count, void* out_keys, u64* out_counts,
                 void* lookback, u64* total_dev, cudaStream_t s, u64* launches) {
    if (!m) { GSB_CUDA_TRY(cudaMemsetAsync(total_dev, 0, 8, s)); return; }
    const u64 lb = rle_lookback_bytes(m);
    GSB_CUDA_TRY(cudaMemsetAsync(lookback, 0, lb, s));
    u32* ticket = (u32*)((char*)lookback + lb - 256);
    const u64 tiles = (m + kRleThreads * kRleItems - 1) / (kRleThreads * kRleItems);
    if (key_bytes == 8) filter_kernel<u64><<<(unsigned)tiles, kRleThreads, 0, s>>>((const u64*)keys, counts, m, min_count, (u64*)out_keys, out_counts, (u64*)lookback, ticket, total_dev);
    else filter_kernel<Key128><<<(unsigned)tiles, kRleThreads, 0, s>>>((const Key128*)keys, counts, m, min_count, (Key128*)out_keys, out_counts, (u64*)lookback, ticket, total_dev);
    ++*launches;
}

// device-wide exclusive scan of u64 weights (n+1 outputs: the last one is the grand total)
void sort_scan_weights(const u64* w, u64* csum, u64 n, u64* tmp, cudaStream_t s, u64* launches) {
    exclusive_scan<u64, u64>(w, csum, n, 0ull, csum + n, tmp, s, launches);
}
u64 sort_scan_tmp_elems(u64 n) { return scan_tmp_elems(n); }

}  // namespace gsb

// ------------------------------------------------------------------------------------------
// orchestration
// ------------------------------------------------------------------------------------------
namespace gsb {

template <typename K>
__global__ void random_keys_kernel(K* keys, u64 n, int key_bits, u64 seed) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        u64 z = (i + seed) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        u64 y = (z + 0x632BE59BD9B4E019ull) * 0xD6E8FEB86659FD93ull; y ^= y >> 32;
        u64 lo = z, hi = y;
        if (key_bits < 64) { lo &= (1ull << key_bits) - 1; hi = 0; }
        else if (key_bits == 64) hi = 0;
        else if (key_bits < 128) hi &= (1ull << (key_bits - 64)) - 1;
        keys[i] = KeyOps<K>::make(lo, hi);
    }
}

void sort_fill_random(int key_bytes, void* keys, u64 n, int key_bits, u64 seed, cudaStream_t s) {
    if (!n) return;
    if (key_bytes == 8) random_keys_kernel<u64><<<148 * 8, 256, 0, s>>>((u64*)keys, n, key_bits, seed);
    else random_keys_kernel<Key128><<<148 * 8, 256, 0, s>>>((Key128*)keys, n, key_bits, seed);
}

int sort_keys(Workspace& ws, int key_bytes, int key_bits, void* a, void* b, u64* va, u64* vb, u64 n,
              const u64* hist_dev, int* passes_run, double* sweep_ms) {
    const int passes = (key_bits + 7) / 8;
    if (passes_run) *passes_run = 0;
    if (n == 0 || passes == 0) return 0;
    cudaStream_t s = ws.stream;
    DevBuf<u64> hist_own;
    const u64* hist = hist_dev;
    if (!hist) {
        hist_own.reset(&ws, (size_t)passes * 256);
        GSB_CUDA_TRY(cudaMemsetAsync(hist_own.p, 0, hist_own.bytes(), s));
        sort_digit_hist(key_bytes, a, n, passes, hist_own.p, ws.sm_count, s, &ws.launches);
        hist = hist_own.p;
    }
    DevBuf<u64> base(&ws, (size_t)passes * 256);
    sort_digit_base(hist, base.p, passes, s, &ws.launches);
    std::vector<u64> h((size_t)passes * 256);
    GSB_CUDA_TRY(cudaMemcpyAsync(h.data(), hist, h.size() * 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    DevBuf<u8> lookback(&ws, sort_lookback_bytes(key_bytes, n));
    int cur = 0, run = 0;
    std::vector<cudaEvent_t> ev;
    if (sweep_ms) { ev.resize(2 * (size_t)passes); for (auto& e : ev) GSB_CUDA_TRY(cudaEventCreate(&e)); }
    for (int p = 0; p < passes; ++p) {
        bool constant = false;
        for (int d = 0; d < 256; ++d) if (h[(size_t)p * 256 + d] == n) { constant = true; break; }
        if (constant) continue;                                // every key has the same digit here: the pass is the identity
        sort_pass(key_bytes, cur ? b : a, cur ? a : b, cur ? vb : va, cur ? va : vb, n, p, base.p, lookback.p, s, &ws.launches,
                  sweep_ms ? ev[2 * run] : nullptr, sweep_ms ? ev[2 * run + 1] : nullptr);
        cur ^= 1; ++run;
    }
    if (sweep_ms) {
        ws.sync();
        for (int i = 0; i < run; ++i) { float ms = 0; GSB_CUDA_TRY(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1])); *sweep_ms += ms; }
        for (auto& e : ev) cudaEventDestroy(e);
    }
    if (passes_run) *passes_run = run;
    return cur;
}

void reduce_sorted(Workspace& ws, int key_bytes, const void* sorted, const u64* weights, u64 n, u64 min_count,
                   ReducedRun& out, u64* m_distinct, void* dkeys_scratch) {
    cudaStream_t s = ws.stream;
    out.m = 0;
    if (m_distinct) *m_distinct = 0;
    if (n == 0) { out.keys.reset(&ws, 0); out.counts.reset(&ws, 0); return; }
    DevBuf<u8> dkeys_own;
    if (!dkeys_scratch) dkeys_own.reset(&ws, (size_t)n * key_bytes);
    u8* const dkeys_p = dkeys_scratch ? (u8*)dkeys_scratch : dkeys_own.p;
    DevBuf<u64> pos(&ws, (size_t)n + 1);
    DevBuf<u64> total(&ws, 1);
    u64 m = 0;
    {
        DevBuf<u8> lookback(&ws, rle_lookback_bytes(n));
        sort_rle(key_bytes, sorted, n, dkeys_p, pos.p, lookback.p, total.p, s, &ws.launches);
        GSB_CUDA_TRY(cudaMemcpyAsync(&m, total.p, 8, cudaMemcpyDeviceToHost, s));
        ws.sync();
    }
    if (m_distinct) *m_distinct = m;
    DevBuf<u64> counts(&ws, (size_t)m);
    if (weights) {
        DevBuf<u64> csum(&ws, (size_t)n + 1);
        DevBuf<u64> tmp(&ws, sort_scan_tmp_elems(n));
        sort_scan_weights(weights, csum.p, n, tmp.p, s, &ws.launches);
        sort_counts_from_pos(pos.p, csum.p, m, counts.p, s, &ws.launches);
        ws.sync();
    } else {
        sort_counts_from_pos(pos.p, nullptr, m, counts.p, s, &ws.launches);
    }
    pos.free();
    if (min_count > 1) {
        DevBuf<u8> fkeys(&ws, (size_t)m * key_bytes);
        DevBuf<u64> fcounts(&ws, (size_t)m);
        DevBuf<u8> lookback(&ws, rle_lookback_bytes(m));
        sort_filter(key_bytes, dkeys_p, counts.p, m, min_count, fkeys.p, fcounts.p, lookback.p, total.p, s, &ws.launches);
        u64 kept = 0;
        GSB_CUDA_TRY(cudaMemcpyAsync(&kept, total.p, 8, cudaMemcpyDeviceToHost, s));
        ws.sync();
        out.keys = std::move(fkeys);
        out.counts = std::move(fcounts);
        out.m = kept;
    } else {
        DevBuf<u8> fkeys(&ws, (size_t)m * key_bytes);
        GSB_CUDA_TRY(cudaMemcpyAsync(fkeys.p, dkeys_p, (size_t)m * key_bytes, cudaMemcpyDeviceToDevice, s));
        out.keys = std::move(fkeys);
        out.counts = std::move(counts);
        out.m = m;
        ws.sync();
    }
}

}  // namespace gsb
