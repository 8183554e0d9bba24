// sort.cu -- counting by sorting: LSD radix sort (8-bit digits, one sweep per digit with
// decoupled look-back), run-length reduce of fully sorted keys, group reduce of partially sorted
// (bit-mixed) keys, min-count filter.
//
// Replaces (result-wise) BackyardHash::insert + BackyardHash::sort/BlendedSort + the
// duplicate-merging emit loop (src/BackyardHash.cc:115-271, src/BlendedSort.hh:58-167,
// src/GossCmdBuildGraph.cc:239-258) and trim-graph's `count > C` predicate
// (src/GossCmdTrimGraph.cc:119).  A cuckoo hash is a random-access structure; on a GPU with
// 8 TB/s of streaming bandwidth the same multiset is counted faster by sorting the instances
// and measuring run lengths, and the sorted order is what the succinct writers need anyway.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <ctime>

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
template <typename K, typename LB, int THREADS, int ITEMS, bool HAS_VALUES, int MINB, int MODE, int LBW = 4>
__global__ void __launch_bounds__(THREADS, MINB) onesweep_kernel(const K* __restrict__ in, K* __restrict__ out,
                                                                 const u64* __restrict__ vin, u64* __restrict__ vout,
                                                                 u64 n, int shift, const u64* __restrict__ digit_base,
                                                                 LB* lookback, u32* ticket, int ablate_arg) {
#ifdef GSB_PROFILING
    const int ablate = ablate_arg;                              // profiling switches (wrong results by construction): -DGSB_PROFILING builds only
#else
    constexpr int ablate = 0;
    (void)ablate_arg;
#endif
    typedef KeyOps<K> KO;
    constexpr int WARPS = THREADS / 32;
    constexpr int TILE = THREADS * ITEMS;
    const int S = LookbackWord<LB>::kShift;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K* keys_s = reinterpret_cast<K*>(smem_raw);                                   // [TILE]
    u32* warp_ofs = reinterpret_cast<u32*>(smem_raw + (size_t)TILE * sizeof(K));   // [WARPS][256] counts, then running offsets
    u64* gofs = reinterpret_cast<u64*>(warp_ofs + WARPS * 256);                    // [256]
    u64* vals_s = gofs + 256;                                                     // [TILE] only with values: the payloads, reordered like the keys
    __shared__ u32 tile_s;
    __shared__ u32 scan_s[THREADS / 32 + 1];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) tile_s = (ablate & 8) ? blockIdx.x : atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < WARPS * 256; i += THREADS) warp_ofs[i] = 0;
    __syncthreads();
    const u32 tile = tile_s;
    const u64 base = (u64)tile * TILE;
    const u32 tile_n = (u32)((n - base) < (u64)TILE ? (n - base) : (u64)TILE);
    const u32 lt_mask = (1u << lane) - 1;
    if (ablate & 16) in += (u64)(tile & 63) * TILE - base;      // profiling only: every load hits the L2

    K key[ITEMS];
    u16 rank[MODE != 0 ? ITEMS : 1];
    const u32 wbase = warp * 32 * ITEMS + lane;
    u32* my_ofs = warp_ofs + warp * 256;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const u32 idx = wbase + i * 32;
        key[i] = idx < tile_n ? in[base + idx] : KO::make(0, 0);
    }

    // peer set of this lane for item i: lanes (with a key) whose digit equals mine, from 8 ballots.
    // Hand-scheduled: per bit one predicate test, one vote, one select, one LOP3 that accumulates the
    // lanes that DIFFER from me in that bit (the compiler's version of the obvious loop spent 36 % of
    // all executed instructions here, ncu source view).
    auto peer_set = [&](u32 d, bool ok) -> u32 {
        const u32 okmask = __ballot_sync(0xffffffffu, ok);
        if (ablate & 2) return ok ? (1u << lane) : 0u;          // profiling only: wrong ranks, no ballots
        u32 differ = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            u32 bal, mine;
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                         "and.b32 t, %2, %3;\n\t"
                         "setp.ne.b32 p, t, 0;\n\t"
                         "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
                         "selp.b32 %1, 0xffffffff, 0, p;\n\t}"
                         : "=r"(bal), "=r"(mine) : "r"(d), "r"(1u << b));
            differ |= bal ^ mine;                               // my bit set: lanes with it clear differ; clear: lanes with it set
        }
        return ~differ & okmask;
    };

    if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (wbase + i * 32 < tile_n) atomicAdd(&my_ofs[KO::digit(key[i], shift)], 1u);
    } else if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const bool ok = wbase + i * 32 < tile_n;
            const u32 d = KO::digit(key[i], shift);
            const u32 peers = peer_set(d, ok);
            const int leader = __ffs(peers) - 1;
            u32 before = 0;
            if (ok && lane == leader) before = atomicAdd(&my_ofs[d], (u32)__popc(peers));
            before = __shfl_sync(0xffffffffu, before, leader < 0 ? lane : leader);
            rank[i] = (u16)(before + __popc(peers & lt_mask));
        }
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
                if (HAS_VALUES) vals_s[pos] = vin[base + wbase + i * 32];   // coalesced load, reordered in shared memory
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            if (wbase + i * 32 < tile_n) {
                const u32 pos = my_ofs[KO::digit(key[i], shift)] + rank[i];
                keys_s[pos] = key[i];
                if (HAS_VALUES) vals_s[pos] = vin[base + wbase + i * 32];
            }
        }
    }

    // exclusive prefix of this tile's digits over all earlier tiles
    if (threadIdx.x < 256) {
        LB excl = 0;
        if (tile != 0 && !(ablate & 1)) {
            excl = lookback_window<LB, LBW>(lookback + threadIdx.x, tile);
            st_volatile(my_state, (LB)(((LB)2 << S) | (excl + (LB)count)));
        }
        gofs[threadIdx.x] = digit_base[threadIdx.x] + (u64)excl - (u64)bstart;
    }
    __syncthreads();

    // coalesced write-out, one contiguous run per digit
    if (ablate & 4) return;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const u32 j = i * THREADS + threadIdx.x;
        if (j < tile_n) {
            const K k = keys_s[j];
            const u64 dst = (ablate & 1) ? (base + j) : gofs[KO::digit(k, shift)] + j;
            out[dst] = k;
            if (HAS_VALUES) vout[dst] = vals_s[j];
        }
    }
}

extern int g_sort_ablate_fwd;
template <typename K, typename LB, int THREADS, int ITEMS, bool HAS_VALUES, int MINB, int MODE, int LBW = 4>
static void launch_onesweep(const void* in, void* out, const u64* vin, u64* vout, u64 n, int shift, const u64* digit_base,
                            void* lookback, u32* ticket, cudaStream_t s) {
    constexpr int TILE = THREADS * ITEMS;
    const size_t smem = (size_t)TILE * sizeof(K) + (size_t)(THREADS / 32) * 256 * 4 + 256 * 8 + (HAS_VALUES ? (size_t)TILE * 8 : 0);
    static bool configured[64] = {false};                   // function attributes are per device
    auto kern = onesweep_kernel<K, LB, THREADS, ITEMS, HAS_VALUES, MINB, MODE, LBW>;
    int dev = 0;
    GSB_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        GSB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GSB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        if (dev >= 0 && dev < 64) configured[dev] = true;
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
static const SortShape kShapes64[] = {{256, 16}, {256, 16}, {256, 16}, {256, 16}, {256, 16}, {1024, 8}, {384, 16}, {256, 20}};
static const int kNumShapes64 = (int)(sizeof(kShapes64) / sizeof(kShapes64[0]));
static const SortShape kShape128 = {256, 8};

void sort_set_tuning(int id) {
    g_sort_ablate = (id >> 8) & 0xFF;
    id &= 0xFF;
    g_sort_tuning = (id >= 0 && id < kNumShapes64) ? id : 0;
}

// tile shape of the (key, payload) sweeps on 64-bit keys: 0 = 256 x 16 at 2 CTAs/SM, 1 = 256 x 8 at 4, 2 = 256 x 12 at 3
static int pair_shape() { return 1; }

u64 sort_tile_keys(int key_bytes, bool with_values) {
    if (key_bytes == 8 && with_values) { static const int items[3] = {16, 8, 12}; return 256ull * items[pair_shape()]; }
    const SortShape sh = key_bytes == 8 ? kShapes64[g_sort_tuning] : kShape128;
    return (u64)sh.threads * sh.items;
}

// bytes of look-back state needed for n keys (+ the ticket word at the end)
u64 sort_lookback_bytes(int key_bytes, u64 n, bool with_values) {
    u64 tiles = (n + sort_tile_keys(key_bytes, with_values) - 1) / sort_tile_keys(key_bytes, with_values);
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
    const u64 lb_bytes = sort_lookback_bytes(key_bytes, n, vin != nullptr);
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
    if (key_bytes == 8 && hv) {
        const int shape = pair_shape();
        if (small) {
            if (shape == 0) launch_onesweep<u64, u32, 256, 16, true, 2, 0>(in, out, vin, vout, n, shift, db, lookback, ticket, s);
            else if (shape == 1) launch_onesweep<u64, u32, 256, 8, true, 4, 0>(in, out, vin, vout, n, shift, db, lookback, ticket, s);
            else launch_onesweep<u64, u32, 256, 12, true, 3, 0>(in, out, vin, vout, n, shift, db, lookback, ticket, s);
        } else {
            if (shape == 0) launch_onesweep<u64, u64, 256, 16, true, 2, 0>(in, out, vin, vout, n, shift, db, lookback, ticket, s);
            else if (shape == 1) launch_onesweep<u64, u64, 256, 8, true, 4, 0>(in, out, vin, vout, n, shift, db, lookback, ticket, s);
            else launch_onesweep<u64, u64, 256, 12, true, 3, 0>(in, out, vin, vout, n, shift, db, lookback, ticket, s);
        }
    } else if (key_bytes == 8) {
        switch (g_sort_tuning) {
            default: GSB_LAUNCH(u64, 256, 16, 4, 0); break;
            case 1: if (small && !hv) { launch_onesweep<u64, u32, 256, 16, false, 4, 0, 8>(in, out, vin, vout, n, shift, db, lookback, ticket, s); break; }
                    GSB_LAUNCH(u64, 256, 16, 4, 0); break;
            case 2: if (small && !hv) { launch_onesweep<u64, u32, 256, 16, false, 4, 0, 16>(in, out, vin, vout, n, shift, db, lookback, ticket, s); break; }
                    GSB_LAUNCH(u64, 256, 16, 4, 0); break;
            case 3: if (small && !hv) { launch_onesweep<u64, u32, 256, 16, false, 4, 0, 32>(in, out, vin, vout, n, shift, db, lookback, ticket, s); break; }
                    GSB_LAUNCH(u64, 256, 16, 4, 0); break;
            case 4: GSB_LAUNCH(u64, 256, 16, 4, 2); break;
            case 5: GSB_LAUNCH(u64, 1024, 8, 1, 0); break;
            case 6: GSB_LAUNCH(u64, 384, 16, 2, 0); break;
            case 7: GSB_LAUNCH(u64, 256, 20, 3, 0); break;
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

// Run-length reduce WITHOUT a cross-tile dependency chain.  ncu showed the single-pass look-back
// version latency bound (58 % of stall samples at the barrier that waits for the look-back, 2.8 ms
// for 5 GB): a tile must hold its keys while it waits, and the register file bounds the bytes in
// flight.  Two streaming passes over the sorted keys instead:
//   pass 1  count, per tile, the run heads that survive the min-count filter (+ all heads, for the stats)
//   scan    tile counts -> tile offsets
//   pass 2  recompute the heads, write (key, count) of the survivors at their final place
// A head at i survives iff keys[i + m - 1] == keys[i] (the keys are sorted), and its count is the
// distance to the end of its run, found by a galloping search that almost always stays inside the
// cache lines the tile has just read.  The second read of the keys comes mostly from the L2.
template <typename K>
struct RleTile {
    // head / survivor ballots of one warp-striped tile; returns the warp's survivor count
    // fold_w > 0: the keys are strand-folded (w symbols each).  A key that is its own reverse complement
    // stands for two instances per occurrence, so it needs only ceil(min_count / 2) occurrences to survive.
    __device__ static __forceinline__ u32 scan(const K* __restrict__ keys, u64 n, u64 wbase, int lane, u64 min_count, int fold_w,
                                               K (&k)[kRleItems], u32 (&kept)[kRleItems], u32 (&headb)[kRleItems], u32& heads, u32& pals) {
        typedef KeyOps<K> KO;
        K carry = KO::make(0, 0);
        if (wbase > 0 && wbase < n) carry = keys[wbase - 1];
#pragma unroll
        for (int i = 0; i < kRleItems; ++i) {
            const u64 idx = wbase + (u64)i * 32 + lane;
            k[i] = idx < n ? keys[idx] : KO::make(0, 0);
        }
        u32 wcount = 0;
        heads = 0; pals = 0;
#pragma unroll
        for (int i = 0; i < kRleItems; ++i) {
            const u64 idx = wbase + (u64)i * 32 + lane;
            const bool ok = idx < n;
            const K& last_src = i ? k[i ? i - 1 : 0] : carry;
            u64 up_lo = __shfl_up_sync(0xffffffffu, KO::lo(k[i]), 1), up_hi = 0;
            u64 last_lo = __shfl_sync(0xffffffffu, KO::lo(last_src), i ? 31 : 0), last_hi = 0;
            if (sizeof(K) == 16) {
                up_hi = __shfl_up_sync(0xffffffffu, KO::hi(k[i]), 1);
                last_hi = __shfl_sync(0xffffffffu, KO::hi(last_src), i ? 31 : 0);
            }
            const K prev = KO::make(lane ? up_lo : last_lo, lane ? up_hi : last_hi);
            const bool head = ok && (idx == 0 || !KO::eq(k[i], prev));
            bool keep = head;
            u64 need = min_count;
            if (fold_w) {
                const bool pal = head && KO::eq(key_rc(k[i], fold_w), k[i]);
                if (pal) need = (min_count + 1) >> 1;
                pals += __popc(__ballot_sync(0xffffffffu, pal));
            }
            if (head && need > 1) keep = idx + need - 1 < n && KO::eq(keys[idx + need - 1], k[i]);
            headb[i] = __ballot_sync(0xffffffffu, head);
            heads += __popc(headb[i]);
            kept[i] = __ballot_sync(0xffffffffu, keep);
            wcount += __popc(kept[i]);
        }
        return wcount;
    }
};

template <typename K>
__global__ void __launch_bounds__(kRleThreads, 4) rle_count_kernel(const K* __restrict__ keys, u64 n, u64 min_count, int fold_w,
                                                                   u32* __restrict__ tile_kept, u64* __restrict__ total_heads /* [0] heads, [2] self-complementary heads */) {
    __shared__ u32 warp_tot[kRleThreads / 32], warp_heads[kRleThreads / 32], warp_pals[kRleThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 wbase = (u64)blockIdx.x * (kRleThreads * kRleItems) + (u64)warp * 32 * kRleItems;
    K k[kRleItems]; u32 kept[kRleItems], headb[kRleItems]; u32 heads, pals;
    const u32 wcount = RleTile<K>::scan(keys, n, wbase, lane, min_count, fold_w, k, kept, headb, heads, pals);
    if (lane == 0) { warp_tot[warp] = wcount; warp_heads[warp] = heads; warp_pals[warp] = pals; }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0, h = 0, p = 0;
#pragma unroll
        for (int w = 0; w < kRleThreads / 32; ++w) { t += warp_tot[w]; h += warp_heads[w]; p += warp_pals[w]; }
        tile_kept[blockIdx.x] = t;
        if (h) atomicAdd(total_heads, (u64)h);
        if (p) atomicAdd(total_heads + 2, (u64)p);
    }
}

// count of the run that starts at idx: distance to the first different key (gallop, then bisect)
template <typename K>
__device__ __forceinline__ u64 run_end(const K* __restrict__ keys, u64 n, u64 known_equal, const K& key) {
    typedef KeyOps<K> KO;
    u64 a = known_equal, step = 1;                              // keys[a] == key
    while (a + step < n && KO::eq(keys[a + step], key)) { a += step; step <<= 1; }
    u64 lo = a + 1, hi = a + step < n ? a + step : n;           // keys[hi] != key or hi == n
    while (lo < hi) { const u64 mid = lo + ((hi - lo) >> 1); if (KO::eq(keys[mid], key)) lo = mid + 1; else hi = mid; }
    return lo;
}

template <typename K>
__global__ void __launch_bounds__(kRleThreads, 4) rle_emit_kernel(const K* __restrict__ keys, const u64* __restrict__ csum, u64 n, u64 min_count, int fold_w,
                                                                  const u64* __restrict__ tile_off, K* __restrict__ out_keys,
                                                                  u64* __restrict__ out_counts) {
    constexpr int WORDS = kRleThreads * kRleItems / 32;          // the tile's run-head bit vector (warp-striped layout
    __shared__ u32 warp_tot[kRleThreads / 32];                  // makes word w*ITEMS+i hold keys [32(w*ITEMS+i), +32))
    __shared__ u32 head_bits[WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 tbase = (u64)blockIdx.x * (kRleThreads * kRleItems);
    const u64 wbase = tbase + (u64)warp * 32 * kRleItems;
    K k[kRleItems]; u32 kept[kRleItems], headb[kRleItems]; u32 heads, pals;
    const u32 wcount = RleTile<K>::scan(keys, n, wbase, lane, min_count, fold_w, k, kept, headb, heads, pals);
    if (lane == 0) {
        warp_tot[warp] = wcount;
#pragma unroll
        for (int i = 0; i < kRleItems; ++i) head_bits[warp * kRleItems + i] = headb[i];
    }
    __syncthreads();
    u64 j = tile_off[blockIdx.x];
    for (int w = 0; w < warp; ++w) j += warp_tot[w];
    const u32 lt = (1u << lane) - 1;
    const u64 tile_end = tbase + kRleThreads * kRleItems < n ? tbase + kRleThreads * kRleItems : n;
#pragma unroll
    for (int i = 0; i < kRleItems; ++i) {
        if ((kept[i] >> lane) & 1u) {
            const u64 idx = wbase + (u64)i * 32 + lane;
            // end of the run = next run head: first from the tile's bit vector, else gallop past the tile
            int w = warp * kRleItems + i;
            u32 m = lane == 31 ? 0u : (head_bits[w] & ~((2u << lane) - 1));
            while (!m && ++w < WORDS) m = head_bits[w];
            const u64 end = m ? tbase + (u64)w * 32 + (__ffs(m) - 1)
                              : (tile_end < n ? run_end<K>(keys, n, tile_end - 1, k[i]) : n);
            const u64 o = j + __popc(kept[i] & lt);
            out_keys[o] = k[i];
            u64 cnt = csum ? (csum[end] - csum[idx]) : (end - idx);
            if (fold_w && KeyOps<K>::eq(key_rc(k[i], fold_w), k[i])) cnt <<= 1;   // both strands of a self-complementary key are this key
            out_counts[o] = cnt;
        }
        j += __popc(kept[i]);
    }
}

__global__ void desc_total_kernel(const ulonglong2* __restrict__ desc, u64 n_desc, u64* __restrict__ total) {
    u64 mine = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_desc; i += (u64)gridDim.x * blockDim.x) mine += desc[i].y;
#pragma unroll
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total, mine);
}

void sort_desc_total(const ulonglong2* desc, u64 n_desc, u64* total_dev, cudaStream_t s, u64* launches) {
    GSB_CUDA_TRY(cudaMemsetAsync(total_dev, 0, 8, s));
    if (!n_desc) return;
    const unsigned blocks = (unsigned)std::min<u64>((n_desc + 255) / 256, 1024);
    desc_total_kernel<<<blocks, 256, 0, s>>>(desc, n_desc, total_dev);
    ++*launches;
}

u64 rle_tiles(u64 n) { return (n + kRleThreads * kRleItems - 1) / (kRleThreads * kRleItems); }
u64 rle_lookback_bytes(u64 n) { return rle_tiles(n) * 8 + 256; }

void sort_rle_count(int key_bytes, const void* keys, u64 n, u64 min_count, int fold_w, u32* tile_kept, u64* total_heads, cudaStream_t s, u64* launches) {
    if (!n) return;
    const unsigned tiles = (unsigned)rle_tiles(n);
    if (key_bytes == 8) rle_count_kernel<u64><<<tiles, kRleThreads, 0, s>>>((const u64*)keys, n, min_count, fold_w, tile_kept, total_heads);
    else rle_count_kernel<Key128><<<tiles, kRleThreads, 0, s>>>((const Key128*)keys, n, min_count, fold_w, tile_kept, total_heads);
    ++*launches;
}

void sort_rle_emit(int key_bytes, const void* keys, const u64* csum, u64 n, u64 min_count, int fold_w, const u64* tile_off, void* out_keys, u64* out_counts,
                   cudaStream_t s, u64* launches) {
    if (!n) return;
    const unsigned tiles = (unsigned)rle_tiles(n);
    if (key_bytes == 8) rle_emit_kernel<u64><<<tiles, kRleThreads, 0, s>>>((const u64*)keys, csum, n, min_count, fold_w, tile_off, (u64*)out_keys, out_counts);
    else rle_emit_kernel<Key128><<<tiles, kRleThreads, 0, s>>>((const Key128*)keys, csum, n, min_count, fold_w, tile_off, (Key128*)out_keys, out_counts);
    ++*launches;
}

// ------------------------------------------------------------------------------------------
// min-count filter: keep (key,count) with count >= min_count, order preserved
// ------------------------------------------------------------------------------------------
// Order-preserving compaction of (key, count >= min_count).  Counts come either from a counts
// array or directly from the run-head positions (count_j = pos[j+1] - pos[j]), which saves writing
// and re-reading the full counts array when a min-count filter follows the run-length reduce.
template <typename K>
__global__ void __launch_bounds__(kRleThreads, 4) filter_kernel(const K* __restrict__ keys, const u64* __restrict__ counts, const u64* __restrict__ pos,
                                                             u64 m, u64 min_count, K* __restrict__ out_keys, u64* __restrict__ out_counts,
                                                             u64* lookback, u32* ticket, u64* __restrict__ total_out) {
    constexpr int WARPS = kRleThreads / 32;
    __shared__ u32 tile_s;
    __shared__ u64 prefix_s;
    __shared__ u32 warp_tot[WARPS];
    if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 tile = tile_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 wbase = (u64)tile * (kRleThreads * kRleItems) + (u64)warp * 32 * kRleItems;
    u64 c[kRleItems];
    u32 ballots[kRleItems];
    u32 wcount = 0;
#pragma unroll
    for (int i = 0; i < kRleItems; ++i) {
        const u64 idx = wbase + (u64)i * 32 + lane;
        c[i] = 0;
        if (idx < m) c[i] = counts ? counts[idx] : (pos[idx + 1] - pos[idx]);
        ballots[i] = __ballot_sync(0xffffffffu, idx < m && c[i] >= min_count);
        wcount += __popc(ballots[i]);
    }
    if (lane == 0) warp_tot[warp] = wcount;
    __syncthreads();
    if (warp == 0) {
        u32 mine = lane < WARPS ? warp_tot[lane] : 0u, inc = mine;
#pragma unroll
        for (int o = 1; o < WARPS; o <<= 1) { const u32 v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
        const u32 run = __shfl_sync(0xffffffffu, inc, WARPS - 1);
        if (lane < WARPS) warp_tot[lane] = inc - mine;
        const u64 p = lookback_exclusive_warp<u64>(lookback, tile, (u64)run);
        if (lane == 0) {
            prefix_s = p;
            if ((u64)(tile + 1) * (kRleThreads * kRleItems) >= m) *total_out = p + run;
        }
    }
    __syncthreads();
    u64 j = prefix_s + warp_tot[warp];
    const u32 lt = (1u << lane) - 1;
#pragma unroll
    for (int i = 0; i < kRleItems; ++i) {
        if ((ballots[i] >> lane) & 1u) {
            const u64 o = j + __popc(ballots[i] & lt);
            out_keys[o] = keys[wbase + (u64)i * 32 + lane];
            out_counts[o] = c[i];
        }
        j += __popc(ballots[i]);
    }
}

void sort_filter(int key_bytes, const void* keys, const u64* counts, const u64* pos, u64 m, u64 min_count, void* out_keys, u64* out_counts,
                 void* lookback, u64* total_dev, cudaStream_t s, u64* launches) {
    if (!m) { GSB_CUDA_TRY(cudaMemsetAsync(total_dev, 0, 8, s)); return; }
    const u64 lb = rle_lookback_bytes(m);
    GSB_CUDA_TRY(cudaMemsetAsync(lookback, 0, lb, s));
    u32* ticket = (u32*)((char*)lookback + lb - 256);
    const u64 tiles = (m + kRleThreads * kRleItems - 1) / (kRleThreads * kRleItems);
    if (key_bytes == 8) filter_kernel<u64><<<(unsigned)tiles, kRleThreads, 0, s>>>((const u64*)keys, counts, pos, m, min_count, (u64*)out_keys, out_counts, (u64*)lookback, ticket, total_dev);
    else filter_kernel<Key128><<<(unsigned)tiles, kRleThreads, 0, s>>>((const Key128*)keys, counts, pos, m, min_count, (Key128*)out_keys, out_counts, (u64*)lookback, ticket, total_dev);
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
              const u64* hist_dev, int* passes_run, double* sweep_ms, int digit_begin, int digit_end) {
    const int passes = (key_bits + 7) / 8;
    if (digit_end < 0 || digit_end > passes) digit_end = passes;
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
    DevBuf<u8> lookback(&ws, sort_lookback_bytes(key_bytes, n, va != nullptr));
    int cur = 0, run = 0;
    std::vector<cudaEvent_t> ev;
    if (sweep_ms) { ev.resize(2 * (size_t)passes); for (auto& e : ev) GSB_CUDA_TRY(cudaEventCreate(&e)); }
    for (int p = digit_begin; p < digit_end; ++p) {
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
                   ReducedRun& out, u64* m_distinct, int fold_w, u64* n_self_rc) {
    cudaStream_t s = ws.stream;
    out.m = 0;
    if (m_distinct) *m_distinct = 0;
    if (n_self_rc) *n_self_rc = 0;
    if (weights && fold_w) throw StatusError{GSB_EINVAL, "internal: strand folding is finished by fold_finalize for merged runs"};
    if (n == 0) { out.keys.reset(&ws, 0); out.counts.reset(&ws, 0); return; }
    if (min_count < 1) min_count = 1;
#ifdef GSB_PROFILING
    const bool trace = getenv("GSB_TRACE_REDUCE") != nullptr;
#else
    const bool trace = false;
#endif
    struct timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
    auto lap = [&](const char* what) {
        if (!trace) return;
        cudaStreamSynchronize(s);
        struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
        fprintf(stderr, "[reduce] %-22s %8.3f ms (reserved %.2f GB)\n", what, (t.tv_sec - ts0.tv_sec) * 1e3 + (t.tv_nsec - ts0.tv_nsec) * 1e-6, ws.reserved_bytes / 1e9);
        ts0 = t;
    };
    // With weights (merging reduced runs) the filter applies to summed weights, not run lengths: emit
    // every run here and let the caller filter.
    const u64 local_min = weights ? 1 : min_count;
    const u64 tiles = rle_tiles(n);
    DevBuf<u32> tile_kept(&ws, tiles);
    DevBuf<u64> tile_off(&ws, tiles), tmp(&ws, scan_tmp_elems(tiles)), scalars(&ws, 4);
    GSB_CUDA_TRY(cudaMemsetAsync(scalars.p, 0, 32, s));
    lap("alloc small");
    sort_rle_count(key_bytes, sorted, n, local_min, fold_w, tile_kept.p, scalars.p, s, &ws.launches);
    lap("rle_count");
    exclusive_scan<u32, u64>(tile_kept.p, tile_off.p, tiles, 0ull, scalars.p + 1, tmp.p, s, &ws.launches);
    u64 host[4] = {0, 0, 0, 0};
    GSB_CUDA_TRY(cudaMemcpyAsync(host, scalars.p, 32, cudaMemcpyDeviceToHost, s));
    ws.sync();
    const u64 heads = host[0], kept = host[1];
    if (m_distinct) *m_distinct = heads;
    if (n_self_rc) *n_self_rc = host[2];
    DevBuf<u64> csum;
    if (weights) {
        csum.reset(&ws, (size_t)n + 1);
        DevBuf<u64> tmp2(&ws, sort_scan_tmp_elems(n));
        sort_scan_weights(weights, csum.p, n, tmp2.p, s, &ws.launches);
    }
    lap("scan + readback");
    out.keys.reset(&ws, (size_t)kept * key_bytes);
    out.counts.reset(&ws, (size_t)kept);
    lap("alloc outputs");
    sort_rle_emit(key_bytes, sorted, weights ? csum.p : nullptr, n, local_min, fold_w, tile_off.p, out.keys.p, out.counts.p, s, &ws.launches);
    out.m = kept;
    lap("rle_emit");
    if (weights && min_count > 1 && kept) {
        DevBuf<u8> fkeys(&ws, (size_t)kept * key_bytes);
        DevBuf<u64> fcounts(&ws, (size_t)kept), total(&ws, 1);
        DevBuf<u8> lookback(&ws, rle_lookback_bytes(kept));
        sort_filter(key_bytes, out.keys.p, out.counts.p, nullptr, kept, min_count, fkeys.p, fcounts.p, lookback.p, total.p, s, &ws.launches);
        u64 k2 = 0;
        GSB_CUDA_TRY(cudaMemcpyAsync(&k2, total.p, 8, cudaMemcpyDeviceToHost, s));
        ws.sync();
        out.keys = std::move(fkeys); out.counts = std::move(fcounts); out.m = k2;
    }
    ws.sync();
}

template <typename K>
__global__ void unmix_kernel(K* keys, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) keys[i] = key_unmix(keys[i]);
}

void sort_unmix_inplace(int key_bytes, void* keys, u64 n, int sm_count, cudaStream_t s, u64* launches) {
    if (!n) return;
    const u64 blocks = (n + 255) / 256;
    const int grid = (int)(blocks < (u64)sm_count * 16 ? blocks : (u64)sm_count * 16);
    if (key_bytes == 8) unmix_kernel<u64><<<grid, 256, 0, s>>>((u64*)keys, n);
    else unmix_kernel<Key128><<<grid, 256, 0, s>>>((Key128*)keys, n);
    ++*launches;
}

}  // namespace gsb
