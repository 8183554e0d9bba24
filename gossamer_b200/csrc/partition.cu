// partition.cu -- counting by PARTITIONING: most-significant-digit passes over the bit-mixed instance keys until a
// bucket fits into one SM's shared memory, then an exact hash count of every bucket in shared memory.
//
// Replaces (result-wise) BackyardHash::insert + the duplicate-merging emit loop + trim-graph's predicate
// (src/BackyardHash.cc:115-242, src/GossCmdBuildGraph.cc:239-258, src/GossCmdTrimGraph.cc:119): the same multiset of
// keys, the same count per distinct key.
//
// Why not the LSD sort of sort.cu for the instances: ncu showed the 8-bit LSD sweep bound by instruction issue, not by
// HBM -- 120 SASS instructions per key, more than half of them the 8 warp ballots per key that a STABLE ranking needs.
// Counting needs no order at all, only that equal keys meet.  So:
//   * the keys are bit-mixed (key_mix, keys.h): uniform, whatever the genome looks like;
//   * a pass splits every parent bucket by the next <= 8 high bits of the mixed key.  Nothing has to be stable, so the
//     rank of a key inside its tile is simply what ONE shared-memory atomicAdd on its digit's counter returns, and the
//     place of the tile's run inside the child bucket is what ONE global atomicAdd per (tile, digit) on the child's
//     cursor returns -- no ballots, no look-back chain between tiles.  ~25 instructions per key: the pass is HBM bound.
//     Tiles are fetched with cp.async.bulk into shared memory behind an mbarrier, one tile ahead of the ranking;
//   * after ceil(log2(n / 3000)) bits (two passes at 200 M keys) a bucket holds ~3000 instances.  One CTA counts a
//     bucket in an open-addressing table in shared memory (exact 64/128-bit key compare), applies the self-complement
//     doubling and the min-count filter, un-mixes the survivors and appends (key, count) to the output.
// Buckets that do not fit (a k-mer repeated a million times) are copied out and go through the full LSD sort +
// run-length reduce of sort.cu, so the result never depends on the bucket geometry.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "kernels.h"
#include "scan.cuh"

namespace gsb {

namespace {

static const int kPtThreads = 256;
static const int kPtTileBytes = 32768;
static const int kPtMaxBins = 1024;                                  // children per parent and pass: 2^10 at most

// ---- mbarrier / bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk) ------------------------------------------
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "WAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra WAIT_DONE;\n\t"
                 "bra WAIT_LOOP;\n\t"
                 "WAIT_DONE:\n\t}" :: "r"(smem_addr(bar)), "r"(parity) : "memory");
}

struct TileRef { u64 begin; u32 count; u32 parent; u32 src; };

// descs == nullptr: one parent covering [0, n).  A descriptor is {begin bits 0..31, begin bits 32..55 | source rank << 24,
// count, parent}: the tile's keys are src_base[source rank][begin, begin + count).
__device__ __forceinline__ TileRef tile_ref(const uint4* __restrict__ descs, u32 tile, u64 n, u32 tile_keys, u64 begin0) {
    TileRef r;
    if (descs) {
        const uint4 d = descs[tile];
        r.begin = (u64)d.x | ((u64)(d.y & 0xFFFFFFu) << 32); r.src = d.y >> 24; r.count = d.z; r.parent = d.w;
    } else {
        const u64 done = (u64)tile * tile_keys;
        r.begin = begin0 + done;                                      // begin0: where the single parent starts in its buffer
        const u64 rem = n - done;
        r.count = rem < (u64)tile_keys ? (u32)rem : tile_keys;
        r.parent = 0; r.src = 0;
    }
    return r;
}

// Consecutive tiles belong to the same parent and would hammer the same 2^bits cursors (one or two L2 slices):
// CTAs take their tiles in a strided permutation so that the tiles in flight spread over all parents.
__device__ __forceinline__ u32 perm_mul(u32 n_tiles) {
    if (n_tiles < 64) return 1;
    if (n_tiles % 7919u) return 7919u;
    if (n_tiles % 10007u) return 10007u;
    return 1;
}
__device__ __forceinline__ u32 perm_tile(u32 i, u32 mul, u32 n_tiles) { return (u32)(((u64)i * mul) % n_tiles); }

struct PartArgs {
    const void* src_base[kMaxRanks];   // where the tiles are read from: [0] = this GPU's buffer; others = peers' buffers (NVLink), pull exchange
    void* out;
    const uint4* descs;        // per-tile {begin lo, begin hi, count, parent}; nullptr = single parent [0, n)
    const u32* n_tiles_dev;    // nullptr = n_tiles
    u32 n_tiles;
    u64 n;
    u64 begin0;                // single-parent passes: element offset of the parent in src_base[0]
    const u64* n_dev;          // optional: the key count of a single-parent pass, known only on the device (n_tiles is then an upper bound)
    u64* cursor;               // [(parent << bits | digit) * cstride]: next free slot of the child (element index from its owner's base)
    u64* hist;                 // [parent << bits | digit]
    u32 cstride;
    int shift, bits;           // digit = (word >> shift) & (2^bits - 1), word = low 64 bits of the mixed key
    // where the children live: child c of a single-parent pass belongs to peer (c * n_peers) >> bits (the multi-GPU
    // exchange stores straight into the owners' windows over NVLink); one GPU: n_peers = 1, peer[0] = out
    void* peer[kMaxRanks];
    int n_peers;
    const u32* abort;          // optional: non-zero = do nothing (the exchange found a window too small)
    const u8* owner_tab;       // optional [2^bits]: the peer that owns each child (balanced ranges of the real key space)
};

// What a pass moves and which bits it splits by.
//   MixedDigit: bare instance keys, bit-mixed; the digit comes from the low (mixed) 64-bit word.
//   RealDigit : (key, count) pairs ordered by the REAL key (pairsort below); shift counts from bit 0 of the whole key.
struct __align__(16) Pair64 { u64 key; u64 count; };
struct __align__(16) Pair128 { u64 lo, hi; u64 count; u64 pad; };
struct MixedDigit {
    template <typename K> static __device__ __forceinline__ u32 digit(const K& k, int shift, u32 mask) { return (u32)(KeyOps<K>::lo(k) >> shift) & mask; }
};
struct RealDigit {
    static __device__ __forceinline__ u32 digit(const Pair64& e, int shift, u32 mask) { return (u32)(e.key >> shift) & mask; }
    static __device__ __forceinline__ u32 digit(const Pair128& e, int shift, u32 mask) {
        Key128 k; k.lo = e.lo; k.hi = e.hi;
        return (u32)KeyOps<Key128>::shr64(k, shift) & mask;
    }
};

// ---- digit histogram of one level ----------------------------------------------------------------------------------
template <typename K, typename D>
__global__ void __launch_bounds__(kPtThreads) part_hist_kernel(PartArgs a) {
    constexpr int ITEMS = kPtTileBytes / (int)sizeof(K) / kPtThreads;
    constexpr u32 TILE = kPtThreads * ITEMS;
    __shared__ u32 cnt_s[kPtMaxBins];
    const int t = threadIdx.x;
    const K* __restrict__ in = (const K*)a.src_base[0];
    const u64 n_keys = a.n_dev ? *a.n_dev : a.n;
    const u32 n_tiles = a.n_tiles_dev ? *a.n_tiles_dev : (a.n_dev ? (u32)((n_keys + TILE - 1) / TILE) : a.n_tiles);
    const u32 mul = perm_mul(n_tiles);
    const u32 nbins = 1u << a.bits, mask = nbins - 1;
    for (u32 i = t; i < nbins; i += kPtThreads) cnt_s[i] = 0;
    __syncthreads();
    u32 cur_parent = 0xffffffffu;
    auto flush = [&]() {
        __syncthreads();
        for (u32 i = t; i < nbins; i += kPtThreads) {
            const u32 c = cnt_s[i];
            if (c) atomicAdd(&a.hist[((u64)cur_parent << a.bits) | i], (u64)c);
            cnt_s[i] = 0;
        }
        __syncthreads();
    };
    for (u32 it = blockIdx.x; it < n_tiles; it += gridDim.x) {
        const TileRef tr = tile_ref(a.descs, perm_tile(it, mul, n_tiles), n_keys, TILE, a.begin0);
        if (tr.parent != cur_parent) {                               // uniform over the CTA
            if (cur_parent != 0xffffffffu) flush();
            cur_parent = tr.parent;
        }
        K key[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const u32 j = (u32)t + (u32)i * kPtThreads;
            if (j < tr.count) key[i] = in[tr.begin + j];
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const u32 j = (u32)t + (u32)i * kPtThreads;
            if (j < tr.count) atomicAdd(&cnt_s[D::digit(key[i], a.shift, mask)], 1u);
        }
    }
    if (cur_parent != 0xffffffffu) flush();
}

// ---- one partition pass ------------------------------------------------------------------------------------------------
// Persistent CTAs; per tile:
//   wait for the tile in shared memory (bulk copy issued one tile ahead) -> keys to registers, rank = what the atomicAdd on
//   the digit's counter returns -> [barrier] -> issue the bulk copy of the NEXT tile into the staging buffer, scan the 2^bits
//   counts, reserve the tile's run in every child with one global atomicAdd per digit -> [barrier] -> keys into the exchange
//   buffer grouped by digit -> [barrier] -> write-out, one contiguous run per digit.
// BPT = bins per thread: 1 (up to 256 children per parent, 3 CTAs per SM) or 4 (up to 1024, 2 CTAs per SM)
template <typename K, typename D, int BPT>
__global__ void __launch_bounds__(kPtThreads, BPT == 1 ? 3 : 2) part_scatter_kernel(PartArgs a) {
    constexpr int ITEMS = kPtTileBytes / (int)sizeof(K) / kPtThreads;
    constexpr u32 TILE = kPtThreads * ITEMS;
    constexpr int MAXBINS = kPtThreads * BPT;
    extern __shared__ __align__(128) unsigned char part_smem[];
    K* stage = reinterpret_cast<K*>(part_smem);                               // TILE keys + 16 bytes (a 64-bit tile may start at an odd index)
    K* out_s = reinterpret_cast<K*>(part_smem + kPtTileBytes + 128);          // TILE keys
    __shared__ __align__(16) u32 cnt_s[MAXBINS];
    __shared__ u32 start_s[MAXBINS];
    __shared__ u64 gaddr_s[MAXBINS];                                          // per digit: address of the tile's run minus its place in out_s
    __shared__ u32 scan_s[kPtThreads / 32 + 1];
    __shared__ __align__(8) u64 full_bar;

    if (a.abort && *a.abort) return;
    const int t = threadIdx.x;
    const u64 n_keys = a.n_dev ? *a.n_dev : a.n;
    const u32 n_tiles = a.n_tiles_dev ? *a.n_tiles_dev : (a.n_dev ? (u32)((n_keys + TILE - 1) / TILE) : a.n_tiles);
    const u32 mul = perm_mul(n_tiles);
    const u32 mask = (1u << a.bits) - 1;
#pragma unroll
    for (int j = 0; j < BPT; ++j) cnt_s[t * BPT + j] = 0;
    if (t == 0) mbar_init(&full_bar, 1);
    __syncthreads();

    auto issue = [&](const TileRef& tr) {                                   // thread 0 only
        const u32 skip = sizeof(K) == 8 ? (u32)(tr.begin & 1) : 0u;
        const u32 bytes = (((skip + tr.count) * (u32)sizeof(K)) + 15u) & ~15u;
        mbar_expect_tx(&full_bar, bytes);
        bulk_load(stage, (const K*)a.src_base[tr.src] + (tr.begin - skip), bytes, &full_bar);   // local HBM, or a peer's buffer over NVLink
    };

    u32 it = blockIdx.x;
    TileRef cur{0, 0, 0, 0};
    if (it < n_tiles) {
        cur = tile_ref(a.descs, perm_tile(it, mul, n_tiles), n_keys, TILE, a.begin0);
        if (t == 0) issue(cur);
    }
    u32 parity = 0;
    for (; it < n_tiles; it += gridDim.x) {
        const u32 nit = it + gridDim.x;
        TileRef nxt{0, 0, 0, 0};
        if (nit < n_tiles) nxt = tile_ref(a.descs, perm_tile(nit, mul, n_tiles), n_keys, TILE, a.begin0);
        mbar_wait(&full_bar, parity);
        parity ^= 1;
        const u32 skip = sizeof(K) == 8 ? (u32)(cur.begin & 1) : 0u;
        K key[ITEMS];
        u32 rank[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const u32 j = (u32)t + (u32)i * kPtThreads;
            if (j < cur.count) key[i] = stage[skip + j];
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const u32 j = (u32)t + (u32)i * kPtThreads;
            rank[i] = 0;
            if (j < cur.count) rank[i] = atomicAdd(&cnt_s[D::digit(key[i], a.shift, mask)], 1u);
        }
        __syncthreads();                                                  // every key is in registers, every count is final
        if (t == 0 && nit < n_tiles) issue(nxt);
        u32 c[BPT];
        u32 sum = 0;
        if (BPT == 4) {
            const uint4 v = *reinterpret_cast<const uint4*>(&cnt_s[t * 4]);
            c[0] = v.x; c[BPT > 1 ? 1 : 0] = v.y; c[BPT > 2 ? 2 : 0] = v.z; c[BPT > 3 ? 3 : 0] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < BPT; ++j) c[j] = cnt_s[t * BPT + j];
        }
#pragma unroll
        for (int j = 0; j < BPT; ++j) sum += c[j];
        u64 g[BPT];
#pragma unroll
        for (int j = 0; j < BPT; ++j) {                                   // issued before the scan, consumed after the scatter below:
            g[j] = 0;                                                     // the round trip to L2 overlaps both
            if (c[j]) g[j] = atomicAdd(&a.cursor[(((u64)cur.parent << a.bits) | (u32)(t * BPT + j)) * a.cstride], (u64)c[j]);
        }
        const u32 ex = block_exclusive_scan<u32, kPtThreads>(sum, (u32*)nullptr, scan_s);
        {
            u32 run = ex;
#pragma unroll
            for (int j = 0; j < BPT; ++j) { start_s[t * BPT + j] = run; run += c[j]; cnt_s[t * BPT + j] = 0; }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const u32 j = (u32)t + (u32)i * kPtThreads;
            if (j < cur.count) out_s[start_s[D::digit(key[i], a.shift, mask)] + rank[i]] = key[i];
        }
        {
            u32 run = ex;
#pragma unroll
            for (int j = 0; j < BPT; ++j) {
                const u32 bin = (u32)(t * BPT + j);
                const u32 owner = a.n_peers > 1 ? (a.owner_tab ? (u32)a.owner_tab[bin & mask] : (u32)(((u64)(bin & mask) * (u32)a.n_peers) >> a.bits)) : 0u;
                gaddr_s[bin] = (u64)a.peer[owner] + (g[j] - (u64)run) * sizeof(K);
                run += c[j];
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const u32 j = (u32)t + (u32)i * kPtThreads;
            if (j < cur.count) {
                const K k = out_s[j];
                *reinterpret_cast<K*>(gaddr_s[D::digit(k, a.shift, mask)] + (u64)j * sizeof(K)) = k;
            }
        }
        cur = nxt;
    }
}

// ---- tiles of the next level: every parent bucket is cut into tiles of its own ------------------------------------------
__global__ void tiles_per_parent_kernel(const u64* __restrict__ cstart, u32 n_parents, u32 tile_keys, u32* __restrict__ tp) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_parents) tp[p] = (u32)((cstart[p + 1] - cstart[p] + tile_keys - 1) / tile_keys);
}

__global__ void fill_descs_kernel(const u64* __restrict__ cstart, const u32* __restrict__ tile_first, u32 n_parents,
                                  const u32* __restrict__ n_tiles_dev, u32 tile_keys, uint4* __restrict__ descs) {
    const u32 tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= *n_tiles_dev) return;
    u32 lo = 0, hi = n_parents;                                     // last parent whose first tile is <= tile (empty parents share the value)
    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (tile_first[mid] <= tile) lo = mid + 1; else hi = mid; }
    const u32 p = lo - 1;
    const u64 begin = cstart[p] + (u64)(tile - tile_first[p]) * tile_keys;
    const u64 rem = cstart[p + 1] - begin;
    descs[tile] = make_uint4((u32)begin, (u32)(begin >> 32), rem < (u64)tile_keys ? (u32)rem : tile_keys, p);
}

__global__ void init_cursor_kernel(const u64* __restrict__ cstart, u64* __restrict__ cursor, u64 n_children, u32 cstride, u64 out_off) {
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_children) cursor[c * cstride] = cstart[c] + out_off;
}

// ---- exact count of every bucket in shared memory -------------------------------------------------------------------------
// Open addressing, linear probing, no deletions.  The home slot comes from the bits of the mixed key just below the
// partition bits.  A slot is claimed with a 64-bit compare-and-swap on the (low) key word; the plain read in front of it
// makes the common case -- the key is there already -- one shared-memory load and one atomicAdd on the count.
// ctr[0] survivors appended, ctr[1] overflow descriptors {first index, length} appended (buckets that do not fit, keys
// that equal the empty marker), ctr[2] distinct keys counted here, ctr[3] of those, self-complementary ones.
static const u64 kEmptyWord = ~0ull;

static const int kBkThreads = 512;

template <typename K> struct CountTable;
template <> struct CountTable<u64> {
    static const int kSlotBytes = 12;
    static const int kPrefetch = 6;                                   // keys per thread held in registers for the NEXT bucket
    volatile u64* key; u32* cnt;
    __device__ __forceinline__ void bind(unsigned char* smem, u32 max_slots) { key = (volatile u64*)smem; cnt = (u32*)(smem + (size_t)max_slots * 8); }
    __device__ __forceinline__ void clear(u32 slots, int t) {
        ulonglong2* k2 = (ulonglong2*)key;
        for (u32 s = t; s < slots / 2; s += kBkThreads) k2[s] = make_ulonglong2(kEmptyWord, kEmptyWord);
        uint4* c4 = (uint4*)cnt;
        for (u32 s = t; s < slots / 4; s += kBkThreads) c4[s] = make_uint4(0, 0, 0, 0);
    }
    __device__ __forceinline__ static bool is_marker(u64 k) { return k == kEmptyWord; }
    __device__ __forceinline__ void insert(u64 k, u32 h, u32 mask) {
        u64 cur = key[h];
        while (cur != k) {
            if (cur == kEmptyWord) {
                const u64 old = atomicCAS((u64*)&key[h], kEmptyWord, k);
                if (old == kEmptyWord || old == k) break;
            }
            h = (h + 1) & mask;
            cur = key[h];
        }
        atomicAdd(&cnt[h], 1u);
    }
    __device__ __forceinline__ u64 load(u32 s) const { return key[s]; }
};
template <> struct CountTable<Key128> {
    static const int kSlotBytes = 20;
    static const int kPrefetch = 4;
    volatile u64* lo; volatile u64* hi; u32* cnt;
    __device__ __forceinline__ void bind(unsigned char* smem, u32 max_slots) {
        lo = (volatile u64*)smem; hi = (volatile u64*)(smem + (size_t)max_slots * 8); cnt = (u32*)(smem + (size_t)max_slots * 16);
    }
    __device__ __forceinline__ void clear(u32 slots, int t) {
        ulonglong2* a = (ulonglong2*)lo; ulonglong2* b = (ulonglong2*)hi;
        for (u32 s = t; s < slots / 2; s += kBkThreads) { a[s] = make_ulonglong2(kEmptyWord, kEmptyWord); b[s] = make_ulonglong2(kEmptyWord, kEmptyWord); }
        uint4* c4 = (uint4*)cnt;
        for (u32 s = t; s < slots / 4; s += kBkThreads) c4[s] = make_uint4(0, 0, 0, 0);
    }
    // the high word of a real key has at most 62 significant bits, so all-ones marks an unset high word; a key whose LOW
    // word equals the marker cannot claim a slot and takes the overflow path
    __device__ __forceinline__ static bool is_marker(const Key128& k) { return k.lo == kEmptyWord; }
    __device__ __forceinline__ void insert(const Key128& k, u32 h, u32 mask) {
        for (;;) {
            u64 cur = lo[h];
            if (cur == kEmptyWord) {
                const u64 old = atomicCAS((u64*)&lo[h], kEmptyWord, k.lo);
                cur = old == kEmptyWord ? k.lo : old;
            }
            if (cur == k.lo) {
                u64 hc = hi[h];
                if (hc == kEmptyWord) {
                    const u64 old = atomicCAS((u64*)&hi[h], kEmptyWord, k.hi);
                    hc = old == kEmptyWord ? k.hi : old;
                }
                if (hc == k.hi) { atomicAdd(&cnt[h], 1u); return; }
            }
            h = (h + 1) & mask;
        }
    }
    __device__ __forceinline__ Key128 load(u32 s) const { Key128 k; k.lo = lo[s]; k.hi = hi[s]; return k; }
};

// a bucket of up to this many keys always finds a table (slots = pow2 >= n + n/8 + 1 <= max_slots)
__host__ __device__ __forceinline__ u32 bucket_cap(u32 max_slots) { return max_slots - max_slots / 8 - 2; }

// FOLD: the keys are strand-folded and self-complementary keys exist (their counts are doubled before the filter)
template <typename K, bool FOLD>
__global__ void __launch_bounds__(kBkThreads, 2) bucket_count_kernel(const K* __restrict__ keys, const u64* __restrict__ cstart, u32 n_buckets,
                                                                     int part_bits, u32 max_slots, u64 min_count, int fold_w,
                                                                     K* __restrict__ out_keys, u64* __restrict__ out_counts, u64 out_cap,
                                                                     ulonglong2* __restrict__ ovf_desc, u64 ovf_cap, u64* __restrict__ ctr) {
    typedef KeyOps<K> KO;
    constexpr int PRE = CountTable<K>::kPrefetch;
    extern __shared__ __align__(16) unsigned char tab_smem[];
    __shared__ u32 scan_s[kBkThreads / 32 + 1];
    __shared__ u64 base_s;
    __shared__ u32 stat_s[2];
    CountTable<K> tab;
    tab.bind(tab_smem, max_slots);
    const int t = threadIdx.x;
    const u32 cap = bucket_cap(max_slots);
    if (t < 2) stat_s[t] = 0;
    __syncthreads();
    u32 my_distinct = 0, my_self = 0;

    // software pipeline: the first PRE * threads keys of the NEXT bucket are fetched into registers while the table of the
    // current one is scanned
    K pre[PRE];
    u32 b = blockIdx.x;
    u64 start = 0, n_b64 = 0;
    auto fetch = [&](u32 bucket, u64& st, u64& nb) {
        st = 0; nb = 0;
        if (bucket < n_buckets) {
            st = cstart[bucket];
            nb = cstart[bucket + 1] - st;
            if (nb <= (u64)cap) {
#pragma unroll
                for (int j = 0; j < PRE; ++j) {
                    const u32 idx = (u32)j * kBkThreads + (u32)t;
                    if (idx < (u32)nb) pre[j] = keys[st + idx];
                }
            }
        }
    };
    fetch(b, start, n_b64);
    for (; b < n_buckets; b += gridDim.x) {
        const u32 nb_next = b + gridDim.x;
        const u64 cur_start = start, cur_n64 = n_b64;
        if (cur_n64 == 0 || cur_n64 > (u64)cap) {
            if (cur_n64 && t == 0) {                                   // does not fit: the whole bucket goes the slow way
                const u64 o = atomicAdd(&ctr[1], 1ull);
                if (o < ovf_cap) ovf_desc[o] = make_ulonglong2(cur_start, cur_n64);
            }
            fetch(nb_next, start, n_b64);
            continue;
        }
        const u32 n_b = (u32)cur_n64;
        u32 slots = 64;
        while (slots < n_b + n_b / 8 + 1) slots <<= 1;
        const u32 mask = slots - 1;
        const int hb = 31 - __clz(slots);
        tab.clear(slots, t);
        __syncthreads();
        auto put = [&](const K& k, u32 idx) {
            if (CountTable<K>::is_marker(k)) {
                const u64 o = atomicAdd(&ctr[1], 1ull);
                if (o < ovf_cap) ovf_desc[o] = make_ulonglong2(cur_start + idx, 1ull);
                return;
            }
            tab.insert(k, (u32)((KO::lo(k) << part_bits) >> (64 - hb)), mask);
        };
#pragma unroll
        for (int j = 0; j < PRE; ++j) {
            const u32 idx = (u32)j * kBkThreads + (u32)t;
            if (idx < n_b) put(pre[j], idx);
        }
        for (u32 idx = (u32)PRE * kBkThreads + (u32)t; idx < n_b; idx += kBkThreads) put(keys[cur_start + idx], idx);   // a rare long bucket
        fetch(nb_next, start, n_b64);                                 // in flight during the table passes below
        __syncthreads();
        // pass 1 over the table: final count of every key (self-complementary keys stand for both strands), survivors noted
        u32 keptmask = 0, my_keep = 0;
        {
            u32 i = 0;
            for (u32 s = t; s < slots; s += kBkThreads, ++i) {
                const u32 c = tab.cnt[s];
                u64 c2 = c;
                if (FOLD) {
                    if (c) {
                        const K real = key_unmix(tab.load(s));
                        if (KO::eq(key_rc(real, fold_w), real)) { c2 <<= 1; ++my_self; tab.cnt[s] = (u32)c2; }
                    }
                }
                my_distinct += c != 0 ? 1u : 0u;
                const bool keep = c != 0 && c2 >= min_count;
                keptmask |= (keep ? 1u : 0u) << i;
                my_keep += keep ? 1u : 0u;
            }
        }
        u32 total = 0;
        const u32 ex = block_exclusive_scan<u32, kBkThreads>(my_keep, &total, scan_s);
        if (t == 0) base_s = total ? atomicAdd(&ctr[0], (u64)total) : 0ull;
        __syncthreads();
        // pass 2: the survivors, un-mixed
        u64 o = base_s + ex;
        while (keptmask) {
            const u32 s = (u32)t + (u32)(__ffs(keptmask) - 1) * kBkThreads;
            keptmask &= keptmask - 1;
            if (o < out_cap) { out_keys[o] = key_unmix(tab.load(s)); out_counts[o] = tab.cnt[s]; }
            ++o;
        }
        __syncthreads();                                              // the table is cleared again at the top of the loop
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        my_distinct += __shfl_xor_sync(0xffffffffu, my_distinct, o);
        my_self += __shfl_xor_sync(0xffffffffu, my_self, o);
    }
    if ((t & 31) == 0) {
        if (my_distinct) atomicAdd(&stat_s[0], my_distinct);
        if (my_self) atomicAdd(&stat_s[1], my_self);
    }
    __syncthreads();
    if (t == 0) {
        if (stat_s[0]) atomicAdd(&ctr[2], (u64)stat_s[0]);
        if (stat_s[1]) atomicAdd(&ctr[3], (u64)stat_s[1]);
    }
}

// overflow descriptors {first index, length}: one warp copies one range, un-mixed
template <typename K>
__global__ void __launch_bounds__(256) copy_overflow_kernel(const K* __restrict__ keys, const ulonglong2* __restrict__ desc, u64 n_desc,
                                                            K* __restrict__ out, u64 out_cap, u64* __restrict__ cursor) {
    const int lane = threadIdx.x & 31;
    for (u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_desc; w += ((u64)gridDim.x * blockDim.x) >> 5) {
        const ulonglong2 d = desc[w];
        u64 base = 0;
        if (lane == 0) base = atomicAdd(cursor, d.y);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (u64 q = lane; q < d.y; q += 32)
            if (base + q < out_cap) out[base + q] = key_unmix(keys[d.x + q]);
    }
}

// out[d] = sum of the 2^(from_bits - bits) bins of `in` that share the top `bits` bits d
__global__ void fold_hist_kernel(const u64* __restrict__ in, int from_bits, int bits, u64* __restrict__ out) {
    const u32 d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= (1u << bits)) return;
    const u32 w = 1u << (from_bits - bits);
    u64 sum = 0;
    for (u32 j = 0; j < w; ++j) sum += in[(d << (from_bits - bits)) + j];
    out[d] = sum;
}

// what a pass moves (element size; which digit functor)
enum ElemKind { EK_KEY64, EK_KEY128, EK_PAIR64, EK_PAIR128 };
static inline ElemKind mixed_kind(int key_bytes) { return key_bytes == 8 ? EK_KEY64 : EK_KEY128; }
static inline ElemKind pair_kind(int key_bytes) { return key_bytes == 8 ? EK_PAIR64 : EK_PAIR128; }
static inline u32 elem_bytes(ElemKind ek) { return ek == EK_KEY64 ? 8u : ek == EK_PAIR128 ? 32u : 16u; }
static inline u32 tile_elems(ElemKind ek) { return (u32)kPtTileBytes / elem_bytes(ek); }

static void launch_hist(ElemKind ek, const PartArgs& a, int grid, cudaStream_t s) {
    switch (ek) {
        case EK_KEY64: part_hist_kernel<u64, MixedDigit><<<grid, kPtThreads, 0, s>>>(a); break;
        case EK_KEY128: part_hist_kernel<Key128, MixedDigit><<<grid, kPtThreads, 0, s>>>(a); break;
        case EK_PAIR64: part_hist_kernel<Pair64, RealDigit><<<grid, kPtThreads, 0, s>>>(a); break;
        case EK_PAIR128: part_hist_kernel<Pair128, RealDigit><<<grid, kPtThreads, 0, s>>>(a); break;
    }
}

template <typename K, typename D, int BPT>
static void launch_scatter_b(const PartArgs& a, int grid, int device, cudaStream_t s) {
    const size_t smem = 2 * (size_t)kPtTileBytes + 256;
    static bool configured[64] = {false};
    if (device < 0 || device >= 64 || !configured[device]) {
        GSB_CUDA_TRY(cudaFuncSetAttribute(part_scatter_kernel<K, D, BPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GSB_CUDA_TRY(cudaFuncSetAttribute(part_scatter_kernel<K, D, BPT>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        if (device >= 0 && device < 64) configured[device] = true;
    }
    part_scatter_kernel<K, D, BPT><<<grid, kPtThreads, smem, s>>>(a);
}

template <typename K, typename D>
static void launch_scatter_w(bool wide, const PartArgs& a, int grid, int device, cudaStream_t s) {
    if (wide) launch_scatter_b<K, D, 4>(a, grid, device, s); else launch_scatter_b<K, D, 1>(a, grid, device, s);
}

static void launch_scatter(ElemKind ek, const PartArgs& a, u64 tiles_ub, Workspace& ws) {
    const bool wide = a.bits > 8;
    const int grid = (int)std::min<u64>(tiles_ub, (u64)ws.sm_count * (wide ? 2 : 3));
    switch (ek) {
        case EK_KEY64: launch_scatter_w<u64, MixedDigit>(wide, a, grid, ws.device, ws.stream); break;
        case EK_KEY128: launch_scatter_w<Key128, MixedDigit>(wide, a, grid, ws.device, ws.stream); break;
        case EK_PAIR64: launch_scatter_w<Pair64, RealDigit>(wide, a, grid, ws.device, ws.stream); break;
        case EK_PAIR128: launch_scatter_w<Pair128, RealDigit>(wide, a, grid, ws.device, ws.stream); break;
    }
    ++ws.launches;
}

template <typename K>
static void launch_bucket_count(const K* keys, const u64* cstart, u32 n_buckets, int part_bits, u32 max_slots, u64 min_count, int fold_w,
                                K* out_keys, u64* out_counts, u64 out_cap, ulonglong2* ovf, u64 ovf_cap, u64* ctr, int grid, int device, cudaStream_t s) {
    const size_t smem = (size_t)max_slots * CountTable<K>::kSlotBytes;
    static size_t configured[64][2] = {{0, 0}};
    const int f = fold_w ? 1 : 0;
    if (device < 0 || device >= 64 || configured[device][f] < smem) {
        if (f) {
            GSB_CUDA_TRY(cudaFuncSetAttribute(bucket_count_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            GSB_CUDA_TRY(cudaFuncSetAttribute(bucket_count_kernel<K, true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        } else {
            GSB_CUDA_TRY(cudaFuncSetAttribute(bucket_count_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            GSB_CUDA_TRY(cudaFuncSetAttribute(bucket_count_kernel<K, false>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        }
        if (device >= 0 && device < 64) configured[device][f] = smem;
    }
    if (f)
        bucket_count_kernel<K, true><<<grid, kBkThreads, smem, s>>>(keys, cstart, n_buckets, part_bits, max_slots, min_count, fold_w, out_keys, out_counts, out_cap,
                                                                    ovf, ovf_cap, ctr);
    else
        bucket_count_kernel<K, false><<<grid, kBkThreads, smem, s>>>(keys, cstart, n_buckets, part_bits, max_slots, min_count, fold_w, out_keys, out_counts, out_cap,
                                                                     ovf, ovf_cap, ctr);
}

}  // namespace

u32 partition_tile_keys(int key_bytes) { return (u32)(kPtTileBytes / key_bytes); }

// test-only overrides of the bucket geometry (gsb_debug_set_partition): results must not depend on them
static u32 g_force_max_slots = 0;
static int g_force_total_bits = 0;
void partition_set_debug(u32 max_slots, int total_bits) { g_force_max_slots = max_slots; g_force_total_bits = total_bits; }
// the same for the pair sort: bucket capacity (0 = default) and partition bits (< 0 = default)
static u32 g_force_pair_cap = 0;
static int g_force_pair_bits = -1;
void pairsort_set_debug(u32 cap, int bits) { g_force_pair_cap = cap; g_force_pair_bits = bits; }

// Bits the partition passes consume for n keys (ALL keys that share the key space: the sum over the ranks of a multi-GPU
// build), and how they are split over the passes: as few passes as 10 bits each allow, the earlier passes taking the
// smaller share (the first pass of a multi-GPU build stores over NVLink, where longer runs per child matter most).
PartitionPlan partition_plan(int key_bytes, u64 n) {
    PartitionPlan p;
    p.max_slots = key_bytes == 8 ? 8192u : 4096u;
    if (g_force_max_slots >= 64 && g_force_max_slots <= p.max_slots && (g_force_max_slots & (g_force_max_slots - 1)) == 0) p.max_slots = g_force_max_slots;
    const u64 cap = bucket_cap(p.max_slots);
    const u64 mean_target = cap / 2;                                  // mean bucket = half of what fits: >= 10 sigma of head room at 50x coverage
    int bits = 0;
    while (bits < 40 && (n >> bits) > mean_target) ++bits;
    if (g_force_total_bits > 0) bits = std::min(g_force_total_bits, 40);
    while (bits > 0 && (1ull << bits) > 64 * n + 256) --bits;          // never (many) more buckets than keys
    p.total_bits = bits;
    p.levels = (bits + 9) / 10;
    if (bits > 10 && bits <= 16) p.levels = 2;
    for (int l = 0; l < p.levels; ++l) p.bits[l] = bits / p.levels + (l >= p.levels - bits % p.levels ? 1 : 0);
    return p;
}

namespace {

struct LevelTiming { cudaEvent_t e0 = nullptr, e1 = nullptr; };

struct LevelTiles {
    DevBuf<uint4> descs;
    DevBuf<u32> n_tiles_dev;
    u64 tiles_ub = 0;
    void apply(PartArgs& pa) const { pa.descs = descs.p; pa.n_tiles_dev = descs.p ? n_tiles_dev.p : nullptr; pa.n_tiles = (u32)tiles_ub; }
};

// tiles over local parents given by cstart [n_parents + 1] (or one parent [0, n) when cstart is null)
void build_tiles(Workspace& ws, ElemKind ek, const u64* cstart, u64 n_parents, u64 n, u64 n_cap, LevelTiles& tiles) {
    cudaStream_t s = ws.stream;
    const u32 tile_keys = tile_elems(ek);
    tiles.tiles_ub = cstart ? (n_cap + tile_keys - 1) / tile_keys + n_parents : (n + tile_keys - 1) / tile_keys;
    if (tiles.tiles_ub > 0xffffffffull) throw StatusError{GSB_EINVAL, "internal: too many partition tiles"};
    if (!cstart) return;
    tiles.descs.reset(&ws, tiles.tiles_ub);
    tiles.n_tiles_dev.reset(&ws, 1);
    DevBuf<u32> tile_first(&ws, n_parents + 1), scan_tmp32(&ws, scan_tmp_elems(n_parents));
    tiles_per_parent_kernel<<<(unsigned)((n_parents + 255) / 256), 256, 0, s>>>(cstart, (u32)n_parents, tile_keys, tile_first.p);
    exclusive_scan<u32, u32>(tile_first.p, tile_first.p, n_parents, 0u, tiles.n_tiles_dev.p, scan_tmp32.p, s, &ws.launches);
    fill_descs_kernel<<<(unsigned)((tiles.tiles_ub + 255) / 256), 256, 0, s>>>(cstart, tile_first.p, (u32)n_parents, tiles.n_tiles_dev.p, tile_keys, tiles.descs.p);
    ws.launches += 2;
}

// hist[parent << bits | digit] += number of such keys (hist must be zeroed by the caller)
void level_hist(Workspace& ws, ElemKind ek, PartArgs pa, const LevelTiles& tiles, u64* hist) {
    tiles.apply(pa);
    pa.hist = hist;
    const int grid = (int)std::min<u64>(tiles.tiles_ub, (u64)ws.sm_count * 8);
    if (grid == 0) return;
    launch_hist(ek, pa, grid, ws.stream);
    ++ws.launches;
}

// child starts from the histogram, cursors, the scatter pass itself
void level_scatter(Workspace& ws, ElemKind ek, PartArgs pa, const LevelTiles& tiles, const u64* hist, u64 n_children, DevBuf<u64>& cstart_out,
                   PartitionTiming* timing, std::vector<LevelTiming>& events, u64 out_off = 0) {
    cudaStream_t s = ws.stream;
    DevBuf<u64> cnext(&ws, n_children + 1), scan_tmp(&ws, scan_tmp_elems(n_children));
    exclusive_scan<u64, u64>(hist, cnext.p, n_children, 0ull, cnext.p + n_children, scan_tmp.p, s, &ws.launches);
    // cursors: spread over distinct cache lines when there are few of them (every tile in flight hits all of them)
    const u32 cstride = n_children <= 4096 ? 32u : 1u;
    DevBuf<u64> cursor(&ws, n_children * cstride);
    init_cursor_kernel<<<(unsigned)((n_children + 255) / 256), 256, 0, s>>>(cnext.p, cursor.p, n_children, cstride, out_off);
    ++ws.launches;
    tiles.apply(pa);
    pa.cursor = cursor.p; pa.cstride = cstride; pa.hist = nullptr;
    LevelTiming lt;
    if (timing) { GSB_CUDA_TRY(cudaEventCreate(&lt.e0)); GSB_CUDA_TRY(cudaEventCreate(&lt.e1)); GSB_CUDA_TRY(cudaEventRecord(lt.e0, s)); }
    if (tiles.tiles_ub) launch_scatter(ek, pa, tiles.tiles_ub, ws);
    if (timing) { GSB_CUDA_TRY(cudaEventRecord(lt.e1, s)); events.push_back(lt); }
    cstart_out = std::move(cnext);
}

// top_bits: where the digits start -- 64 for the mixed low word of an instance key, the key width for pairs ordered by the real key
PartArgs local_args(const void* cur, void* other, u64 n, int consumed, int bits, int top_bits = 64) {
    PartArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.src_base[0] = cur; pa.out = other; pa.n = n;
    pa.shift = top_bits - consumed - bits; pa.bits = bits;
    pa.peer[0] = other; pa.n_peers = 1; pa.abort = nullptr;
    return pa;
}

// One pass over `cur` (parents given by cstart, or one parent [0, n) when cstart is empty) by `bits` more bits into `other`.
// Produces the child starts.  hist_ready: optional histogram [n_parents << bits] that is already known.
// n_dev (single-parent passes only): the key count lives on the device; n is then its upper bound.
void run_level(Workspace& ws, ElemKind ek, void* cur, void* other, u64 n, u64 n_cap, DevBuf<u64>& cstart, u64 n_parents, int consumed, int bits,
               const u64* hist_ready, PartitionTiming* timing, std::vector<LevelTiming>& events, int top_bits = 64, const u64* n_dev = nullptr,
               u64 in_off = 0, u64 out_off = 0) {
    // in_off / out_off (single-parent passes): the parent is cur[in_off, in_off + n), its children go to other[out_off + ...)
    const u64 n_children = n_parents << bits;
    PartArgs pa = local_args(cur, other, n, consumed, bits, top_bits);
    pa.n_dev = cstart.p ? nullptr : n_dev;
    pa.begin0 = cstart.p ? 0 : in_off;
    LevelTiles tiles;
    build_tiles(ws, ek, cstart.p, n_parents, n, n_cap, tiles);
    DevBuf<u64> hist;
    if (!hist_ready) {
        hist.reset(&ws, n_children);
        GSB_CUDA_TRY(cudaMemsetAsync(hist.p, 0, n_children * 8, ws.stream));
        level_hist(ws, ek, pa, tiles, hist.p);
        hist_ready = hist.p;
    }
    level_scatter(ws, ek, pa, tiles, hist_ready, n_children, cstart, timing, events, out_off);
}

// ---- pull exchange: tiles over the runs that every source rank holds of this rank's children ------------------------------
// gathered[s * (C + 1) + c] = start of child c in source s's level-0 output; pair q = p * n_src + s (p: this rank's p-th child)
__global__ void pull_tiles_per_pair_kernel(const u64* __restrict__ gathered, u32 C, u32 n_src, u32 c_lo, u32 n_pairs, u32 tile_keys, u32* __restrict__ tp) {
    const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_pairs) return;
    const u32 p = q / n_src, sr = q % n_src;
    const u64* g = gathered + (size_t)sr * (C + 1) + c_lo + p;
    tp[q] = (u32)((g[1] - g[0] + tile_keys - 1) / tile_keys);
}

// same_base: all sources lie in ONE buffer and their starts are absolute (the blocks of a streamed single-GPU build)
__global__ void pull_fill_descs_kernel(const u64* __restrict__ gathered, u32 C, u32 n_src, u32 c_lo, u32 n_pairs, const u32* __restrict__ tile_first,
                                       const u32* __restrict__ n_tiles_dev, u32 tile_keys, uint4* __restrict__ descs, int same_base) {
    const u32 tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= *n_tiles_dev) return;
    u32 lo = 0, hi = n_pairs;
    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (tile_first[mid] <= tile) lo = mid + 1; else hi = mid; }
    const u32 q = lo - 1, p = q / n_src, sr = q % n_src;
    const u64* g = gathered + (size_t)sr * (C + 1) + c_lo + p;
    const u64 begin = g[0] + (u64)(tile - tile_first[q]) * tile_keys;
    const u64 rem = g[1] - begin;
    descs[tile] = make_uint4((u32)begin, (u32)((begin >> 32) & 0xFFFFFFu) | ((same_base ? 0u : sr) << 24), rem < (u64)tile_keys ? (u32)rem : tile_keys, p);
}

// after the all-reduce of the next level's histograms: does this rank's share fit its buffer?  If not (on any rank: every
// rank evaluates every rank), the flag is raised and this rank's slice of the histogram is zeroed, so that everything
// downstream sees empty buckets.  Also: how many keys this rank will pull, and how many of those from other ranks.
__global__ void __launch_bounds__(1024) pull_check_kernel(u64* __restrict__ hist_all /* [C << b1] */, const u64* __restrict__ gathered, u32 C, int b1, int n, int rank,
                                                          u64 cap_keys, u32* __restrict__ abort_flag, u64* __restrict__ n_recv, u64* __restrict__ n_remote) {
    __shared__ u64 tot_s[kMaxRanks];
    __shared__ u32 abort_s;
    const u32 t = threadIdx.x;
    if (t < kMaxRanks) tot_s[t] = 0;
    if (t == 0) abort_s = 0;
    __syncthreads();
    // per-rank totals from the level-0 child sizes (sum over sources)
    for (u32 c = t; c < C; c += blockDim.x) {
        u64 v = 0;
        for (int sr = 0; sr < n; ++sr) { const u64* g = gathered + (size_t)sr * (C + 1) + c; v += g[1] - g[0]; }
        atomicAdd(&tot_s[(u32)(((u64)c * (u32)n) / C)], v);
    }
    __syncthreads();
    if (t < (u32)n && tot_s[t] > cap_keys) atomicOr(&abort_s, 1u);
    __syncthreads();
    const bool abort = abort_s != 0;
    const u32 lo = (u32)(((u64)rank * C + n - 1) / n), hi = (u32)(((u64)(rank + 1) * C + n - 1) / n);
    if (abort) {
        const u64 a = (u64)lo << b1, b = (u64)hi << b1;
        for (u64 i = a + t; i < b; i += blockDim.x) hist_all[i] = 0;
    }
    if (t == 0) {
        *abort_flag = abort ? 1u : 0u;
        u64 mine_local = 0;
        for (u32 c = lo; c < hi; ++c) { const u64* g = gathered + (size_t)rank * (C + 1) + c; mine_local += g[1] - g[0]; }
        *n_recv = abort ? 0ull : tot_s[rank];
        *n_remote = abort ? 0ull : tot_s[rank] - mine_local;
    }
}

}  // namespace

// level 0 of a multi-GPU build: this rank's n keys are split by their top `bits` bits and every child's run is stored
// straight into its owner's window (peers.base[(child * n_ranks) >> bits], element offsets from cursor[child * cstride])
void partition_scatter_to_peers(Workspace& ws, int key_bytes, const void* in, u64 n, int bits, u64* cursor, u32 cstride,
                                void* const* peer_base, int n_peers, const u32* abort_flag) {
    if (!n) return;
    const u32 tile_keys = partition_tile_keys(key_bytes);
    PartArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.src_base[0] = in; pa.out = nullptr; pa.n = n;
    pa.shift = 64 - bits; pa.bits = bits;
    for (int r = 0; r < n_peers; ++r) pa.peer[r] = peer_base[r];
    pa.n_peers = n_peers; pa.abort = abort_flag;
    pa.cursor = cursor; pa.cstride = cstride;
    const u64 tiles = (n + tile_keys - 1) / tile_keys;
    if (tiles > 0xffffffffull) throw StatusError{GSB_EINVAL, "internal: too many partition tiles"};
    pa.n_tiles = (u32)tiles;
    launch_scatter(mixed_kind(key_bytes), pa, tiles, ws);
}

void partition_fold_hist(Workspace& ws, const u64* hist_top, int bits, u64* out) {
    fold_hist_kernel<<<((1u << bits) + 255) / 256, 256, 0, ws.stream>>>(hist_top, kTopHistBits, bits, out);
    ++ws.launches;
}

// ---- pieces of the multi-GPU pull exchange (orchestrated by exchange.cu) -------------------------------------------------
// level 0, local: this rank's n keys by their top `bits` bits into `out` (its peer-mapped window); cstart_out [2^bits + 1]
void partition_local_level0(Workspace& ws, int key_bytes, void* in, void* out, u64 n, int bits, const u64* hist_top, DevBuf<u64>& cstart_out,
                            cudaEvent_t* e0_out, cudaEvent_t* e1_out, u64 in_off, u64 out_off) {
    std::vector<LevelTiming> ev;
    PartitionTiming want_events;
    DevBuf<u64> folded;
    const u64* hist_ready = nullptr;
    if (hist_top && bits <= kTopHistBits) {
        folded.reset(&ws, (size_t)1 << bits);
        partition_fold_hist(ws, hist_top, bits, folded.p);
        hist_ready = folded.p;
    }
    DevBuf<u64> none;
    run_level(ws, mixed_kind(key_bytes), in, out, n, n, none, 1, 0, bits, hist_ready, e0_out ? &want_events : nullptr, ev, 64, nullptr, in_off, out_off);
    cstart_out = std::move(none);
    if (e0_out) {                                                       // events around the scatter launch: the caller owns them now
        *e0_out = ev.empty() ? nullptr : ev[0].e0;
        *e1_out = ev.empty() ? nullptr : ev[0].e1;
    }
}

// histogram of the NEXT `bits` bits of every level-0 child of this rank's own output: hist [2^(bits0 + bits)] (zeroed here)
void partition_next_hist(Workspace& ws, int key_bytes, const void* keys, const u64* cstart, int bits0, u64 n, int bits, u64* hist, bool accumulate) {
    const u64 n_parents = 1ull << bits0;
    if (!accumulate) GSB_CUDA_TRY(cudaMemsetAsync(hist, 0, (n_parents << bits) * 8, ws.stream));
    PartArgs pa = local_args(keys, nullptr, n, bits0, bits);
    LevelTiles tiles;
    build_tiles(ws, mixed_kind(key_bytes), cstart, n_parents, n, n, tiles);
    level_hist(ws, mixed_kind(key_bytes), pa, tiles, hist);
}

void partition_pull_check(Workspace& ws, u64* hist_all, const u64* gathered, int bits0, int bits1, int n_ranks, int rank, u64 cap_keys,
                          u32* abort_flag, u64* n_recv, u64* n_remote) {
    pull_check_kernel<<<1, 1024, 0, ws.stream>>>(hist_all, gathered, 1u << bits0, bits1, n_ranks, rank, cap_keys, abort_flag, n_recv, n_remote);
    ++ws.launches;
}

// level 1 of a multi-GPU build, fused with the exchange: this rank's children [c_lo, c_lo + n_parents) are PULLED tile by
// tile from wherever they lie -- src_base[s] is rank s's level-0 output, peer-mapped; the partition kernel's bulk copies
// fetch them over NVLink -- and split by the next `bits` bits into `out` (local).  hist_slice [n_parents << bits]: the
// all-reduced histogram of exactly these children.  cstart_out [(n_parents << bits) + 1].
void partition_pull_level(Workspace& ws, int key_bytes, const void* const* src_base, int n_src, const u64* gathered, int bits0, u32 c_lo, u32 n_parents,
                          u64 n_cap, int bits, const u64* hist_slice, void* out, const u32* abort_flag, DevBuf<u64>& cstart_out, cudaEvent_t e0, cudaEvent_t e1,
                          bool same_base) {
    cudaStream_t s = ws.stream;
    const u32 tile_keys = partition_tile_keys(key_bytes);
    const u32 C = 1u << bits0;
    const u32 n_pairs = n_parents * (u32)n_src;
    LevelTiles tiles;
    tiles.tiles_ub = (n_cap + tile_keys - 1) / tile_keys + n_pairs;
    if (tiles.tiles_ub > 0xffffffffull) throw StatusError{GSB_EINVAL, "internal: too many partition tiles"};
    tiles.descs.reset(&ws, tiles.tiles_ub);
    tiles.n_tiles_dev.reset(&ws, 1);
    DevBuf<u32> tile_first(&ws, (size_t)n_pairs + 1), scan_tmp32(&ws, scan_tmp_elems(n_pairs));
    pull_tiles_per_pair_kernel<<<(n_pairs + 255) / 256, 256, 0, s>>>(gathered, C, (u32)n_src, c_lo, n_pairs, tile_keys, tile_first.p);
    exclusive_scan<u32, u32>(tile_first.p, tile_first.p, n_pairs, 0u, tiles.n_tiles_dev.p, scan_tmp32.p, s, &ws.launches);
    pull_fill_descs_kernel<<<(unsigned)((tiles.tiles_ub + 255) / 256), 256, 0, s>>>(gathered, C, (u32)n_src, c_lo, n_pairs, tile_first.p, tiles.n_tiles_dev.p, tile_keys, tiles.descs.p,
                                                                                    same_base ? 1 : 0);
    ws.launches += 2;
    PartArgs pa = local_args(nullptr, out, 0, bits0, bits);
    if (same_base) pa.src_base[0] = src_base[0];
    else for (int r = 0; r < n_src; ++r) pa.src_base[r] = src_base[r];
    pa.abort = abort_flag;
    std::vector<LevelTiming> ev;
    if (e0) GSB_CUDA_TRY(cudaEventRecord(e0, s));
    level_scatter(ws, mixed_kind(key_bytes), pa, tiles, hist_slice, (u64)n_parents << bits, cstart_out, nullptr, ev);
    if (e1) GSB_CUDA_TRY(cudaEventRecord(e1, s));
}

// ---- streamed single-GPU builds: the first pass runs per input block, while the next block is still crossing PCIe ----------
// (context.cu).  Every block is split by the same top bits0 bits into its own region of `alt` (same offsets as in `keys`);
// child c of the batch is then the union of every block's run c -- exactly the situation of the multi-GPU pull exchange
// with the blocks as sources, all in one buffer -- and the second pass gathers the runs while it splits them further.
namespace {
__global__ void add_offset_kernel(const u64* __restrict__ in, u64 n, u64 off, u64* __restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] + off;
}
}  // namespace

int partition_stream_bits0() { return g_force_total_bits > 0 ? std::min(g_force_total_bits, 8) : 8; }

// First pass over ONE block: keys[block_off, block_off + n_block) by their top bits0 bits into alt at the same offsets (the
// buffers are addressed from their aligned bases: a block may start at an odd key); runs_out [2^bits0 + 1] = ABSOLUTE
// starts (block_off added) of the block's children; hist_next
// [2^(bits0 + kTopHistBits)] += histogram of the next kTopHistBits bits of every child.
// hist_block_top: the block's own histogram of the top kTopHistBits bits (fused into the extraction).
void partition_block_level0(Workspace& ws, int key_bytes, void* keys, void* alt, u64 block_off, u64 n_block, int bits0, const u64* hist_block_top,
                            u64* runs_out, u64* hist_next, cudaEvent_t* e0_out, cudaEvent_t* e1_out) {
    const u64 C = 1ull << bits0;
    DevBuf<u64> cstart_local;
    partition_local_level0(ws, key_bytes, keys, alt, n_block, bits0, hist_block_top, cstart_local, e0_out, e1_out, block_off, block_off);
    add_offset_kernel<<<(unsigned)((C + 1 + 255) / 256), 256, 0, ws.stream>>>(cstart_local.p, C + 1, block_off, runs_out);
    ++ws.launches;
    partition_next_hist(ws, key_bytes, alt, runs_out, bits0, n_block, kTopHistBits, hist_next, true);
}

// pass widths when the first pass (bits0) has been run already: the rest as partition_plan would split it; at least one
// more pass (possibly of 0 bits: it gathers the blocks' runs into contiguous buckets)
PartitionPlan partition_plan_streamed(int key_bytes, u64 n, int bits0) {
    PartitionPlan full = partition_plan(key_bytes, n);
    PartitionPlan p;
    p.max_slots = full.max_slots;
    const int rest = full.total_bits > bits0 ? full.total_bits - bits0 : 0;
    const int rl = std::max(1, (rest + 9) / 10);
    p.levels = 1 + rl;
    p.bits[0] = bits0;
    for (int l = 0; l < rl; ++l) p.bits[1 + l] = rest / rl + (l >= rl - rest % rl ? 1 : 0);
    p.total_bits = bits0 + rest;
    return p;
}

// Count bit-mixed keys (see the file header).  in.keys holds them; in.scratch is a buffer of the same capacity; both are
// overwritten and need 16 bytes of slack behind their last key.  Either one parent of in.n keys (host-known), or
// in.cstart / in.n_parents / in.consumed_bits describe buckets that earlier passes (the multi-GPU exchange) have made, the
// key count then being known only on the device (in.n_cap bounds it for allocations).
// Produces every distinct key (un-mixed) whose final count is >= min_count, in ARBITRARY order, with its count; fold_w > 0
// doubles the count of self-complementary keys before the filter (fold.cu).
// Returns false (nothing produced) only when min_count > 1 and more keys survive than the output buffers hold -- the
// caller then sorts by the full key instead; *where_keys: the (still mixed, permuted) keys are in keys (0) or scratch (1).
bool count_partitioned(Workspace& ws, int key_bytes, int key_bits, PartitionInput& in, const PartitionPlan& plan, u64 min_count, int fold_w,
                       ReducedRun& out, u64* m_distinct, u64* n_self_rc, int* where_keys, PartitionTiming* timing) {
    cudaStream_t s = ws.stream;
    out.m = 0;
    *m_distinct = 0; *n_self_rc = 0;
    if (where_keys) *where_keys = 0;
    const bool streamed = in.runs != nullptr;                            // per-block first pass done (partition_block_level0)
    const bool plain = in.cstart.p == nullptr && !streamed;
    const u64 n_cap = (plain || streamed) ? in.n : in.n_cap;
    if (n_cap == 0) { out.keys.reset(&ws, 0); out.counts.reset(&ws, 0); return true; }
    if (min_count < 1) min_count = 1;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    auto mark = [&](int i) { if (timing) { GSB_CUDA_TRY(cudaEventCreate(&ev[i])); GSB_CUDA_TRY(cudaEventRecord(ev[i], s)); } };
    std::vector<LevelTiming> levels_ev;

    // ---- partition passes ----
    void* cur = in.keys; void* other = in.scratch;
    DevBuf<u64> cstart = std::move(in.cstart);                          // child starts of the last pass run: [children + 1]
    u64 n_parents = plain ? 1 : in.n_parents;
    int consumed = plain ? 0 : in.consumed_bits;
    int level = 0, done_bits = 0;
    while (level < plan.levels && done_bits < consumed) done_bits += plan.bits[level++];   // passes the caller has run already
    mark(0);
    DevBuf<u64> folded;
    if (streamed) {
        // second pass: gathers every block's run of each child while splitting it by the next bits
        const int bits0 = in.runs_bits0, bits1 = plan.bits[1];
        const u32 C = 1u << bits0;
        folded.reset(&ws, (size_t)C << bits1);
        fold_hist_kernel<<<(unsigned)((((size_t)C << bits1) + 255) / 256), 256, 0, s>>>(in.runs_hist_next, bits0 + kTopHistBits, bits0 + bits1, folded.p);
        ++ws.launches;
        const void* base[1] = {cur};
        LevelTiming lt;
        if (timing) { GSB_CUDA_TRY(cudaEventCreate(&lt.e0)); GSB_CUDA_TRY(cudaEventCreate(&lt.e1)); }
        partition_pull_level(ws, key_bytes, base, in.runs_n_src, in.runs, bits0, 0, C, in.n, bits1, folded.p, other, nullptr, cstart, lt.e0, lt.e1, true);
        if (timing) levels_ev.push_back(lt);
        std::swap(cur, other);
        n_parents = (u64)C << bits1;
        consumed = bits0 + bits1;
        level = 2;
    }
    for (; level < plan.levels; ++level) {
        const int bits = plan.bits[level];
        const u64* hist_ready = nullptr;
        if (plain && level == 0 && in.hist_top && bits <= kTopHistBits) {      // the extraction kernel counted the top bits already
            folded.reset(&ws, (size_t)1 << bits);
            partition_fold_hist(ws, in.hist_top, bits, folded.p);
            hist_ready = folded.p;
        }
        run_level(ws, mixed_kind(key_bytes), cur, other, in.n, n_cap, cstart, n_parents, consumed, bits, hist_ready, timing, levels_ev);
        std::swap(cur, other);
        n_parents <<= bits;
        consumed += bits;
    }
    mark(1);
    if (cstart.p == nullptr) {                                          // no pass at all: one bucket
        cstart.reset(&ws, 2);
        const u64 h[2] = {0, in.n};
        GSB_CUDA_TRY(cudaMemcpyAsync(cstart.p, h, 16, cudaMemcpyHostToDevice, s));
        ws.sync();                                                     // h is on the stack
    }

    // ---- count every bucket ----
    const u64 n_buckets = n_parents;
    const u64 out_cap = min_count > 1 ? n_cap / 2 + 65536 : n_cap;
    const u64 ovf_cap = std::max<u64>(n_buckets, 1u << 16) + n_cap / 1024;
    DevBuf<u64> out_counts(&ws, out_cap), ctr(&ws, 8);
    DevBuf<ulonglong2> ovf(&ws, ovf_cap);
    GSB_CUDA_TRY(cudaMemsetAsync(ctr.p, 0, 64, s));
    const int grid = (int)std::min<u64>(n_buckets, (u64)ws.sm_count * 2 * 4);       // 2 resident CTAs per SM, 4 interleaved rounds
    if (key_bytes == 8)
        launch_bucket_count<u64>((const u64*)cur, cstart.p, (u32)n_buckets, consumed, plan.max_slots, min_count, fold_w, (u64*)other, out_counts.p, out_cap,
                                 ovf.p, ovf_cap, ctr.p, grid, ws.device, s);
    else
        launch_bucket_count<Key128>((const Key128*)cur, cstart.p, (u32)n_buckets, consumed, plan.max_slots, min_count, fold_w, (Key128*)other, out_counts.p, out_cap,
                                    ovf.p, ovf_cap, ctr.p, grid, ws.device, s);
    ++ws.launches;
    mark(2);
    u64 h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    GSB_CUDA_TRY(cudaMemcpyAsync(h, ctr.p, 64, cudaMemcpyDeviceToHost, s));
    ws.sync();
    if (where_keys) *where_keys = cur == in.keys ? 0 : 1;
    if (timing) {
        float ms = 0;
        GSB_CUDA_TRY(cudaEventElapsedTime(&ms, ev[0], ev[1])); timing->ms_partition += ms;
        GSB_CUDA_TRY(cudaEventElapsedTime(&ms, ev[1], ev[2])); timing->ms_count += ms;
        for (auto& lt : levels_ev) {
            GSB_CUDA_TRY(cudaEventElapsedTime(&ms, lt.e0, lt.e1)); timing->ms_scatter += ms;
            cudaEventDestroy(lt.e0); cudaEventDestroy(lt.e1);
        }
        timing->levels = plan.levels; timing->total_bits = plan.total_bits;
        timing->scatter_launches += levels_ev.size();
        for (int i = 0; i < 3; ++i) cudaEventDestroy(ev[i]);
    }
    if (h[0] > out_cap) return false;
    if (h[1] > ovf_cap) throw StatusError{GSB_EINVAL, "internal: overflow list of the bucket count exceeded"};

    // ---- buckets that did not fit: copied out un-mixed, full sort + run-length reduce ----
    ReducedRun slow; u64 d2 = 0, self2 = 0;
    if (h[1]) {
        DevBuf<u64> total(&ws, 1);
        sort_desc_total(ovf.p, h[1], total.p, s, &ws.launches);
        u64 n_slow = 0;
        GSB_CUDA_TRY(cudaMemcpyAsync(&n_slow, total.p, 8, cudaMemcpyDeviceToHost, s));
        ws.sync();
        DevBuf<u8> imp(&ws, n_slow * key_bytes), alt(&ws, n_slow * key_bytes);
        GSB_CUDA_TRY(cudaMemsetAsync(total.p, 0, 8, s));
        {
            const unsigned blocks = (unsigned)std::min<u64>((h[1] + 7) / 8, (u64)ws.sm_count * 16);
            if (key_bytes == 8) copy_overflow_kernel<u64><<<blocks, 256, 0, s>>>((const u64*)cur, ovf.p, h[1], (u64*)imp.p, n_slow, total.p);
            else copy_overflow_kernel<Key128><<<blocks, 256, 0, s>>>((const Key128*)cur, ovf.p, h[1], (Key128*)imp.p, n_slow, total.p);
            ++ws.launches;
        }
        const int where = sort_keys(ws, key_bytes, key_bits, imp.p, alt.p, nullptr, nullptr, n_slow, nullptr, nullptr);
        reduce_sorted(ws, key_bytes, where ? alt.p : imp.p, nullptr, n_slow, min_count, slow, &d2, fold_w, &self2);
        if (timing) timing->n_overflow_keys += n_slow;
    }
    const u64 m = h[0] + slow.m;
    out.keys.reset(&ws, m * key_bytes);
    out.counts.reset(&ws, m);
    if (h[0]) {
        GSB_CUDA_TRY(cudaMemcpyAsync(out.keys.p, other, h[0] * key_bytes, cudaMemcpyDeviceToDevice, s));
        GSB_CUDA_TRY(cudaMemcpyAsync(out.counts.p, out_counts.p, h[0] * 8, cudaMemcpyDeviceToDevice, s));
    }
    if (slow.m) {
        GSB_CUDA_TRY(cudaMemcpyAsync(out.keys.p + h[0] * key_bytes, slow.keys.p, slow.m * key_bytes, cudaMemcpyDeviceToDevice, s));
        GSB_CUDA_TRY(cudaMemcpyAsync(out.counts.p + h[0], slow.counts.p, slow.m * 8, cudaMemcpyDeviceToDevice, s));
    }
    out.m = m;
    *m_distinct = h[2] + d2;
    *n_self_rc = h[3] + self2;
    ws.sync();
    return true;
}

// ======================================================================================================================
// Ordering (key, count) pairs with DISTINCT keys by the real key: the survivors of the count come back in arbitrary order
// and the writers need them sorted (src/BackyardHash.cc:244-271 is the order the reference's emit loop sees).  An LSD
// radix sort of (key, count) pairs reads and writes every pair once per key byte -- 8 sweeps at k = 31, 14 at k = 55.
// Distinct keys need no stable sort, so the same most-significant-digit passes as above are used, on 16/32-byte
// {key, count} elements and the top bits of the REAL key, until a bucket holds ~2000 pairs; one CTA then orders a bucket
// in shared memory: counting sort on the next 12 bits, rank among the (few) pairs that share a bin by direct comparison.
// A bucket that does not fit (real genomes are not uniform: low-complexity prefixes) is written out unordered and
// radix-sorted on its own afterwards, so the result never depends on the data; if there are many such buckets the caller
// falls back to the radix sort of everything.
// ======================================================================================================================
namespace {

static const int kPkThreads = 256;
static const int kPkItems = 4;
static const int kBsThreads = 512;
static const int kBsBinBits = 12;

template <typename K> struct PairOf;
template <> struct PairOf<u64> {
    typedef Pair64 type;
    static __device__ __forceinline__ Pair64 make(u64 k, u64 c) { Pair64 e; e.key = k; e.count = c; return e; }
    static __device__ __forceinline__ u64 key(const Pair64& e) { return e.key; }
};
template <> struct PairOf<Key128> {
    typedef Pair128 type;
    static __device__ __forceinline__ Pair128 make(const Key128& k, u64 c) { Pair128 e; e.lo = k.lo; e.hi = k.hi; e.count = c; e.pad = 0; return e; }
    static __device__ __forceinline__ Key128 key(const Pair128& e) { Key128 k; k.lo = e.lo; k.hi = e.hi; return k; }
};

// (keys[i], counts[i]) -> out[i]; fold_w > 0: additionally (rc keys[i], counts[i]) for every key that is not its own reverse
// complement, appended behind the first m elements (cursor starts at m; arbitrary order).  Fuses the histogram of the top
// `bits` bits of every element written.
template <typename K>
__global__ void __launch_bounds__(kPkThreads) pairs_pack_kernel(const K* __restrict__ keys, const u64* __restrict__ counts, u64 m, int fold_w,
                                                                typename PairOf<K>::type* __restrict__ out, u64* __restrict__ cursor,
                                                                u64* __restrict__ hist, int shift, int bits) {
    typedef typename PairOf<K>::type E;
    __shared__ u32 cnt_s[kPtMaxBins];
    const int t = threadIdx.x, lane = t & 31;
    const u32 lt = (1u << lane) - 1;
    const u32 nbins = 1u << bits, mask = nbins - 1;
    for (u32 i = t; i < nbins; i += kPkThreads) cnt_s[i] = 0;
    __syncthreads();
    const u64 tiles = (m + (u64)kPkThreads * kPkItems - 1) / ((u64)kPkThreads * kPkItems);
    for (u64 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const u64 base0 = tile * kPkThreads * kPkItems;
        E r[kPkItems]; u32 bal[kPkItems]; u32 total = 0;
#pragma unroll
        for (int it = 0; it < kPkItems; ++it) {
            const u64 i = base0 + (u64)it * kPkThreads + t;
            bool take = false;
            if (i < m) {
                const K y = keys[i];
                const u64 c = counts[i];
                const E e = PairOf<K>::make(y, c);
                out[i] = e;
                if (bits) atomicAdd(&cnt_s[RealDigit::digit(e, shift, mask)], 1u);
                if (fold_w) {
                    const K ry = key_rc(y, fold_w);
                    take = !KeyOps<K>::eq(ry, y);
                    r[it] = PairOf<K>::make(ry, c);
                }
            }
            bal[it] = __ballot_sync(0xffffffffu, take);
            total += __popc(bal[it]);
        }
        if (fold_w) {
            u64 base = 0;
            if (lane == 0 && total) base = atomicAdd(cursor, (u64)total);
            base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
            for (int it = 0; it < kPkItems; ++it) {
                if ((bal[it] >> lane) & 1u) {
                    out[base + __popc(bal[it] & lt)] = r[it];
                    if (bits) atomicAdd(&cnt_s[RealDigit::digit(r[it], shift, mask)], 1u);
                }
                base += __popc(bal[it]);
            }
        }
    }
    __syncthreads();
    for (u32 i = t; i < nbins; i += kPkThreads) { const u32 c = cnt_s[i]; if (c) atomicAdd(&hist[i], (u64)c); }
}

template <typename K> struct BucketSmem;
template <> struct BucketSmem<u64> {
    static const u32 kCap = 4096;
    u64* k; u64* c;
    __device__ __forceinline__ void bind(unsigned char* p) { k = (u64*)p; c = k + kCap; }
    static size_t bytes() { return (size_t)kCap * 16; }
    __device__ __forceinline__ void put(u32 i, const Pair64& e) { k[i] = e.key; c[i] = e.count; }
    __device__ __forceinline__ u64 key(u32 i) const { return k[i]; }
};
template <> struct BucketSmem<Key128> {
    static const u32 kCap = 2048;
    u64* lo; u64* hi; u64* c;
    __device__ __forceinline__ void bind(unsigned char* p) { lo = (u64*)p; hi = lo + kCap; c = hi + kCap; }
    static size_t bytes() { return (size_t)kCap * 24; }
    __device__ __forceinline__ void put(u32 i, const Pair128& e) { lo[i] = e.lo; hi[i] = e.hi; c[i] = e.count; }
    __device__ __forceinline__ Key128 key(u32 i) const { Key128 k; k.lo = lo[i]; k.hi = hi[i]; return k; }
};

// One CTA per bucket [cstart[b], cstart[b+1]) of elems (all pairs of a bucket share the top `consumed` bits of the key):
// the pairs come out ordered by key into out_keys / out_counts at the same positions.  Buckets above the capacity are
// copied unordered and listed in ovf ({first, length}; ctr[0] counts them).
template <typename K>
__global__ void __launch_bounds__(kBsThreads) bucket_sort_kernel(const typename PairOf<K>::type* __restrict__ elems, const u64* __restrict__ cstart, u32 n_buckets,
                                                                 int key_bits, int consumed, u32 cap, K* __restrict__ out_keys, u64* __restrict__ out_counts,
                                                                 ulonglong2* __restrict__ ovf, u64 ovf_cap, u64* __restrict__ ctr) {
    typedef typename PairOf<K>::type E;
    typedef KeyOps<K> KO;
    constexpr u32 CAP = BucketSmem<K>::kCap;
    constexpr u32 NB = 1u << kBsBinBits;
    constexpr int BPT = NB / kBsThreads;
    extern __shared__ __align__(16) unsigned char bs_smem[];
    __shared__ u32 bin_s[NB + 1];
    __shared__ u16 pos_s[CAP];          // place inside the bin, later: perm[rank] = element
    __shared__ u16 ord_s[CAP];          // elements in bin order
    __shared__ u32 scan_s[kBsThreads / 32 + 1];
    BucketSmem<K> sm;
    sm.bind(bs_smem);
    const int t = threadIdx.x;
    const int rem = key_bits - consumed;
    const int bb = rem < kBsBinBits ? (rem > 0 ? rem : 0) : kBsBinBits;
    const int shift = rem - bb;
    const u32 bmask = (1u << bb) - 1;
    for (u32 b = blockIdx.x; b < n_buckets; b += gridDim.x) {
        const u64 st = cstart[b];
        const u64 nb64 = cstart[b + 1] - st;
        if (nb64 == 0) continue;
        if (nb64 > cap) {                                              // cap <= CAP
            for (u64 i = t; i < nb64; i += kBsThreads) {
                const E e = elems[st + i];
                out_keys[st + i] = PairOf<K>::key(e);
                out_counts[st + i] = e.count;
            }
            if (t == 0) {
                const u64 o = atomicAdd(&ctr[0], 1ull);
                if (o < ovf_cap) ovf[o] = make_ulonglong2(st, nb64);
            }
            continue;
        }
        const u32 nb = (u32)nb64;
#pragma unroll
        for (int j = 0; j < BPT; ++j) bin_s[t * BPT + j] = 0;
        __syncthreads();
        for (u32 i = t; i < nb; i += kBsThreads) {
            const E e = elems[st + i];
            sm.put(i, e);
            pos_s[i] = (u16)atomicAdd(&bin_s[(u32)KO::shr64(PairOf<K>::key(e), shift) & bmask], 1u);
        }
        __syncthreads();
        {
            u32 c[BPT]; u32 sum = 0;
#pragma unroll
            for (int j = 0; j < BPT; ++j) { c[j] = bin_s[t * BPT + j]; sum += c[j]; }
            u32 run = block_exclusive_scan<u32, kBsThreads>(sum, (u32*)nullptr, scan_s);
#pragma unroll
            for (int j = 0; j < BPT; ++j) { bin_s[t * BPT + j] = run; run += c[j]; }
            if (t == kBsThreads - 1) bin_s[NB] = run;
        }
        __syncthreads();
        for (u32 i = t; i < nb; i += kBsThreads) ord_s[bin_s[(u32)KO::shr64(sm.key(i), shift) & bmask] + pos_s[i]] = (u16)i;
        __syncthreads();
        for (u32 i = t; i < nb; i += kBsThreads) {
            const K mine = sm.key(i);
            const u32 bin = (u32)KO::shr64(mine, shift) & bmask;
            const u32 s0 = bin_s[bin], s1 = bin_s[bin + 1];
            u32 r = s0;
            if (s1 - s0 > 1) {
                for (u32 j = s0; j < s1; ++j) {
                    const u32 o = ord_s[j];
                    const K other = sm.key(o);
                    r += (KO::lt(other, mine) || (KO::eq(other, mine) && o < i)) ? 1u : 0u;
                }
            }
            pos_s[r] = (u16)i;             // pos_s[i] was last read before the barrier above
        }
        __syncthreads();
        for (u32 r = t; r < nb; r += kBsThreads) {
            const u32 i = pos_s[r];
            out_keys[st + r] = sm.key(i);
            out_counts[st + r] = sm.c[i];
        }
        __syncthreads();
    }
}

template <typename K>
static void launch_bucket_sort(const void* elems, const u64* cstart, u32 n_buckets, int key_bits, int consumed, u32 cap, void* out_keys, u64* out_counts,
                               ulonglong2* ovf, u64 ovf_cap, u64* ctr, int sm_count, int device, cudaStream_t s) {
    typedef typename PairOf<K>::type E;
    const size_t smem = BucketSmem<K>::bytes();
    static bool configured[64] = {false};
    if (device < 0 || device >= 64 || !configured[device]) {
        GSB_CUDA_TRY(cudaFuncSetAttribute(bucket_sort_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GSB_CUDA_TRY(cudaFuncSetAttribute(bucket_sort_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        if (device >= 0 && device < 64) configured[device] = true;
    }
    const int grid = (int)std::min<u64>(n_buckets, (u64)sm_count * 2 * 8);
    bucket_sort_kernel<K><<<grid, kBsThreads, smem, s>>>((const E*)elems, cstart, n_buckets, key_bits, consumed, cap, (K*)out_keys, out_counts, ovf, ovf_cap, ctr);
}

}  // namespace

// ---- pieces of the pair sort (also used by the multi-GPU re-partition of the survivors, exchange.cu) -------------------------
u32 pairsort_elem_bytes(int key_bytes) { return elem_bytes(pair_kind(key_bytes)); }

// Pass widths for n_cap pairs: buckets of half the capacity on average, as few passes as 10 bits each allow.
// first_bits > 0 fixes the width of the first pass (the multi-GPU exchange splits by a fixed 2^first_bits children).
PairSortPlan pairsort_plan(int key_bytes, int key_bits, u64 n_cap, int first_bits) {
    PairSortPlan p;
    p.cap = key_bytes == 8 ? BucketSmem<u64>::kCap : BucketSmem<Key128>::kCap;
    if (g_force_pair_cap >= 2 && g_force_pair_cap < p.cap) p.cap = g_force_pair_cap;
    int bits = 0;
    while (bits < 40 && (n_cap >> bits) > p.cap / 2) ++bits;
    if (g_force_pair_bits >= 0) bits = g_force_pair_bits;
    if (bits > key_bits) bits = key_bits;
    if (first_bits > 0) {
        if (first_bits > key_bits) first_bits = key_bits;
        const int rest = bits > first_bits ? bits - first_bits : 0;
        const int rl = (rest + 9) / 10;
        p.levels = 1 + rl;
        p.lb[0] = first_bits;
        for (int l = 0; l < rl; ++l) p.lb[1 + l] = rest / rl + (l < rest % rl ? 1 : 0);
        p.bits = first_bits + rest;
        return p;
    }
    p.bits = bits;
    p.levels = (bits + 9) / 10;
    for (int l = 0; l < p.levels; ++l) p.lb[l] = bits / p.levels + (l < bits % p.levels ? 1 : 0);
    return p;
}

// (keys, counts)[0, m) -> {key, count} elements in `elems` (room for 2 m when fold_w > 0: the reverse complements are
// appended); *n_dev = elements written; hist0 [2^bits0] += histogram of their top bits0 bits (zeroed by the caller)
void pairsort_pack(Workspace& ws, int key_bytes, int key_bits, const void* keys, const u64* counts, u64 m, int fold_w, void* elems, u64* n_dev, u64* hist0, int bits0) {
    cudaStream_t s = ws.stream;
    GSB_CUDA_TRY(cudaMemcpyAsync(n_dev, &m, 8, cudaMemcpyHostToDevice, s));              // cursor of the appended reverse complements (pageable source: staged before the call returns)
    const u64 tiles = (m + (u64)kPkThreads * kPkItems - 1) / ((u64)kPkThreads * kPkItems);
    const int grid = (int)std::max<u64>(1, std::min<u64>(tiles, (u64)ws.sm_count * 8));
    const int shift0 = key_bits - bits0;
    if (key_bytes == 8) pairs_pack_kernel<u64><<<grid, kPkThreads, 0, s>>>((const u64*)keys, counts, m, fold_w, (Pair64*)elems, n_dev, hist0, shift0, bits0);
    else pairs_pack_kernel<Key128><<<grid, kPkThreads, 0, s>>>((const Key128*)keys, counts, m, fold_w, (Pair128*)elems, n_dev, hist0, shift0, bits0);
    ++ws.launches;
}

// The remaining passes plan.lb[first_level ..] and the shared-memory sort of every bucket.  The elements are in `cur`
// (`other`: scratch of the same capacity, unused when no pass remains); either one parent whose size is *n_dev (cstart
// empty), or parents made by earlier passes (cstart [n_parents + 1], `consumed` bits).  n_cap bounds the element count.
// hist_first: the histogram of the first remaining pass if it is known already.  false = the data defeat the geometry.
bool pairsort_finish(Workspace& ws, int key_bytes, int key_bits, void* cur, void* other, u64 n_cap, const u64* n_dev, DevBuf<u64>& cstart, u64 n_parents,
                     int consumed, const PairSortPlan& plan, int first_level, const u64* hist_first, ReducedRun& out) {
    cudaStream_t s = ws.stream;
    const ElemKind ek = pair_kind(key_bytes);
    std::vector<LevelTiming> ev;
    for (int l = first_level; l < plan.levels; ++l) {
        if (plan.lb[l] == 0) continue;
        run_level(ws, ek, cur, other, n_cap, n_cap, cstart, n_parents, consumed, plan.lb[l], l == first_level ? hist_first : nullptr, nullptr, ev, key_bits, n_dev);
        std::swap(cur, other);
        n_parents <<= plan.lb[l];
        consumed += plan.lb[l];
    }
    if (cstart.p == nullptr) {                                          // no pass at all: one bucket [0, n)
        cstart.reset(&ws, 2);
        GSB_CUDA_TRY(cudaMemsetAsync(cstart.p, 0, 8, s));
        GSB_CUDA_TRY(cudaMemcpyAsync(cstart.p + 1, n_dev, 8, cudaMemcpyDeviceToDevice, s));
    }
    const u64 ovf_cap = 64;
    DevBuf<ulonglong2> ovf(&ws, ovf_cap);
    DevBuf<u64> ctr(&ws, 1);
    GSB_CUDA_TRY(cudaMemsetAsync(ctr.p, 0, 8, s));
    out.keys.reset(&ws, n_cap * key_bytes);
    out.counts.reset(&ws, n_cap);
    if (key_bytes == 8) launch_bucket_sort<u64>(cur, cstart.p, (u32)n_parents, key_bits, consumed, plan.cap, out.keys.p, out.counts.p, ovf.p, ovf_cap, ctr.p, ws.sm_count, ws.device, s);
    else launch_bucket_sort<Key128>(cur, cstart.p, (u32)n_parents, key_bits, consumed, plan.cap, out.keys.p, out.counts.p, ovf.p, ovf_cap, ctr.p, ws.sm_count, ws.device, s);
    ++ws.launches;
    u64 h[2] = {0, 0};
    GSB_CUDA_TRY(cudaMemcpyAsync(&h[0], cstart.p + n_parents, 8, cudaMemcpyDeviceToHost, s));   // the element count
    GSB_CUDA_TRY(cudaMemcpyAsync(&h[1], ctr.p, 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    if (h[1] > ovf_cap) return false;
    if (h[1]) {
        // buckets that did not fit: every one is a contiguous range of the final order, radix-sorted on its own
        std::vector<ulonglong2> d(h[1]);
        GSB_CUDA_TRY(cudaMemcpyAsync(d.data(), ovf.p, h[1] * sizeof(ulonglong2), cudaMemcpyDeviceToHost, s));
        ws.sync();
        for (const ulonglong2& r : d) {
            DevBuf<u8> ka(&ws, r.y * key_bytes);
            DevBuf<u64> ca(&ws, r.y);
            u8* kp = out.keys.p + r.x * key_bytes;
            u64* cp = out.counts.p + r.x;
            const int where = sort_keys(ws, key_bytes, key_bits, kp, ka.p, cp, ca.p, r.y, nullptr, nullptr);
            if (where) {
                GSB_CUDA_TRY(cudaMemcpyAsync(kp, ka.p, r.y * key_bytes, cudaMemcpyDeviceToDevice, s));
                GSB_CUDA_TRY(cudaMemcpyAsync(cp, ca.p, r.y * 8, cudaMemcpyDeviceToDevice, s));
            }
        }
    }
    out.m = h[0];
    return true;
}

// m (key, count) pairs with distinct keys, arbitrary order -> ordered by key; fold_w > 0: the reverse complement of every
// key that is not self-complementary joins the set first (fold.cu: the folded, filtered run becomes the reference's edge
// set).  out is (re)allocated; out.m = pairs produced.  Returns false (nothing produced) when the data defeat the bucket
// geometry -- the caller radix-sorts instead.
bool sort_pairs_msd(Workspace& ws, int key_bytes, int key_bits, const void* keys, const u64* counts, u64 m, int fold_w, ReducedRun& out) {
    cudaStream_t s = ws.stream;
    const u64 n_cap = fold_w ? 2 * m : m;
    if (n_cap == 0) { out.keys.reset(&ws, 0); out.counts.reset(&ws, 0); out.m = 0; return true; }
    const u32 eb = pairsort_elem_bytes(key_bytes);
    const PairSortPlan plan = pairsort_plan(key_bytes, key_bits, n_cap, 0);
    const int bits0 = plan.levels ? plan.lb[0] : 0;
    DevBuf<u8> ea(&ws, n_cap * eb + 64), eb2(&ws, plan.levels ? n_cap * eb + 64 : 1);
    DevBuf<u64> n_dev(&ws, 1), hist0(&ws, (size_t)1 << bits0);
    GSB_CUDA_TRY(cudaMemsetAsync(hist0.p, 0, hist0.bytes(), s));
    pairsort_pack(ws, key_bytes, key_bits, keys, counts, m, fold_w, ea.p, n_dev.p, hist0.p, bits0);
    DevBuf<u64> cstart;
    return pairsort_finish(ws, key_bytes, key_bits, ea.p, eb2.p, n_cap, n_dev.p, cstart, 1, 0, plan, 0, hist0.p, out);
}

// ---- multi-GPU: level 0 of the pair sort crosses NVLink (orchestrated by exchange.cu) ---------------------------------------
namespace {

// One CTA.  hist_all [n][C]: every rank's histogram of the top bits of its elements.  Children are dealt to the ranks as
// contiguous ranges of (nearly) equal element counts -- rank o owns [clo_o, chi_o), so the concatenation of the ranks'
// slices is the global order -- and inside its owner's window a child lies at (elements of the owner's earlier children)
// + (the same child's elements from lower ranks).
//   cursor [C * cstride]: where THIS rank's elements of each child go (element index in the owner's window)
//   owner [C]; totals [n]: elements every rank receives; range [2]: this rank's children [clo, chi)
//   cstart_local [C + 1]: starts of this rank's children in its own window (entries 0 .. chi - clo)
__global__ void __launch_bounds__(1024) pair_owner_plan_kernel(const u64* __restrict__ hist_all, int n, int rank, u32 C, u64* __restrict__ cursor, u32 cstride,
                                                               u8* __restrict__ owner, u64* __restrict__ totals, u32* __restrict__ range, u64* __restrict__ cstart_local) {
    __shared__ u64 scan_s[1024 / 32 + 1];
    __shared__ u64 first_s[kMaxRanks], tot_s[kMaxRanks];
    __shared__ u32 lo_s, hi_s;
    const u32 c = threadIdx.x;
    if (c < (u32)kMaxRanks) { first_s[c] = ~0ull; tot_s[c] = 0; }
    if (c == 0) { lo_s = 0xffffffffu; hi_s = 0; }
    u64 tot = 0, before_me = 0;
    if (c < C) {
        for (int r = 0; r < n; ++r) {
            const u64 v = hist_all[(size_t)r * C + c];
            tot += v;
            if (r < rank) before_me += v;
        }
    }
    u64 grand = 0;
    const u64 P = block_exclusive_scan<u64, 1024>(tot, &grand, scan_s);     // barriers inside: the initialisations above are visible
    u32 own = 0;
    if (c < C) {
        own = grand ? (u32)((P * (u64)n) / grand) : 0u;
        if (own >= (u32)n) own = (u32)n - 1;
        atomicMin(&first_s[own], P);
        atomicAdd(&tot_s[own], tot);
        if (own == (u32)rank) { atomicMin(&lo_s, c); atomicMax(&hi_s, c + 1); }
    }
    __syncthreads();
    if (c < C) {
        owner[c] = (u8)own;
        cursor[(size_t)c * cstride] = P - first_s[own] + before_me;
        if (own == (u32)rank) cstart_local[c - lo_s] = P - first_s[own];
    }
    if (c < (u32)n) totals[c] = tot_s[c];
    if (c == 0) {
        const u32 lo = lo_s == 0xffffffffu ? 0u : lo_s, hi = lo_s == 0xffffffffu ? 0u : hi_s;
        range[0] = lo; range[1] = hi;
        cstart_local[hi - lo] = tot_s[rank];
    }
}

}  // namespace

void pairsort_plan_owners(Workspace& ws, const u64* hist_all, int n_ranks, int rank, int bits0, u64* cursor, u32 cstride, u8* owner, u64* totals, u32* range,
                          u64* cstart_local) {
    pair_owner_plan_kernel<<<1, 1024, 0, ws.stream>>>(hist_all, n_ranks, rank, 1u << bits0, cursor, cstride, owner, totals, range, cstart_local);
    ++ws.launches;
}

// first pass of the pair sort, every child stored straight into its owner's window
void pairsort_scatter_to_peers(Workspace& ws, int key_bytes, int key_bits, const void* elems, u64 n_cap, const u64* n_dev, int bits0, u64* cursor, u32 cstride,
                               const u8* owner, void* const* peer_base, int n_peers) {
    if (!n_cap) return;
    const ElemKind ek = pair_kind(key_bytes);
    PartArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.src_base[0] = elems; pa.out = nullptr; pa.n = n_cap; pa.n_dev = n_dev;
    pa.shift = key_bits - bits0; pa.bits = bits0;
    for (int r = 0; r < n_peers; ++r) pa.peer[r] = peer_base[r];
    pa.n_peers = n_peers; pa.owner_tab = owner;
    pa.cursor = cursor; pa.cstride = cstride;
    const u64 tiles = (n_cap + tile_elems(ek) - 1) / tile_elems(ek);
    if (tiles > 0xffffffffull) throw StatusError{GSB_EINVAL, "internal: too many partition tiles"};
    pa.n_tiles = (u32)tiles;
    launch_scatter(ek, pa, tiles, ws);
}

}  // namespace gsb
