// xeno.cu -- the kernels behind the last two steps of `xenome index` (src/XenoApp.cc:62-76), SURVEY section 8(f) N3:
//
//   merge-and-annotate-kmer-sets   src/GossCmdMergeAndAnnotateKmerSets.cc:27-207
//       the union of two kmer sets plus one membership bit per element and side (<out>.lhs-bits / .rhs-bits,
//       WordyBitVector::Builder::push_backX, src/WordyBitVector.hh:90-116).  The union itself is the merge of reader.cu /
//       fold.cu / sort.cu with the left set weighted 1 and the right set weighted 2: the summed weight of an element IS its
//       pair of membership bits.
//   compute-near-kmers             src/GossCmdComputeNearKmers.cc:57-118,158-225
//       a k-mer that belongs to exactly one side turns "gray" (both bits cleared) when one of its variants
//       y = x ^ (b << j), 0 <= j < K, 1 <= b < 4, is a member that belongs to exactly one side -- the OTHER side.
//       The reference walks the set with one thread per block of elements and a select()/accessAndRank() per probe; here
//       one thread owns one element and binary-searches the decoded, sorted key array for its 3K variants.
//       Restated as the reference BEHAVES (the files must match): the variant mask is shifted by j
//       bits, not bases, and the variant is looked up without normalisation (the reference discards the result of its
//       normalize call).
#include "kernels.h"

namespace gsb {

namespace {

// bit i of lhs = weight[i] & 1, of rhs = weight[i] >> 1; 32 elements per warp ballot, written as u32 halves of the
// little-endian u64 words
__global__ void __launch_bounds__(256) annotate_bits_kernel(const u64* __restrict__ weight, u64 m, u32* __restrict__ lhs, u32* __restrict__ rhs) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 w = i < m ? weight[i] : 0;
    const u32 lb = __ballot_sync(0xffffffffu, (w & 1) != 0), rb = __ballot_sync(0xffffffffu, (w & 2) != 0);
    if ((threadIdx.x & 31) == 0 && i < m) { lhs[i >> 5] = lb; rhs[i >> 5] = rb; }
}

__device__ __forceinline__ bool bit_at(const u32* __restrict__ words, u64 i) { return (words[i >> 5] >> (i & 31)) & 1u; }

// index of y in keys[0, m), or m
template <typename K>
__device__ __forceinline__ u64 find_key(const K* __restrict__ keys, u64 m, const K& y) {
    u64 lo = 0, hi = m;
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (KeyOps<K>::lt(keys[mid], y)) lo = mid + 1; else hi = mid;
    }
    return (lo < m && KeyOps<K>::eq(keys[lo], y)) ? lo : m;
}

template <typename K>
__global__ void __launch_bounds__(256) near_kmers_kernel(const K* __restrict__ keys, u64 m, int k, const u32* __restrict__ lhs, const u32* __restrict__ rhs,
                                                         u32* __restrict__ new_lhs, u32* __restrict__ new_rhs, u64* __restrict__ n_gray) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool li = false, ri = false, gray = false;
    if (i < m) {
        li = bit_at(lhs, i); ri = bit_at(rhs, i);
        if (li != ri) {
            const K x = keys[i];
            for (int j = 0; j < k && !gray; ++j) {
                for (u64 b = 1; b < 4 && !gray; ++b) {                 // b = 0 leaves x unchanged (skipped by the reference, :93)
                    const K y = KeyOps<K>::make(KeyOps<K>::lo(x) ^ (b << j), KeyOps<K>::hi(x));   // j <= 62: the mask stays in the low word
                    const u64 r = find_key<K>(keys, m, y);
                    if (r < m) {
                        const bool lr = bit_at(lhs, r), rr = bit_at(rhs, r);
                        if (lr != rr && li != lr) gray = true;
                    }
                }
            }
        }
    }
    const u32 lb = __ballot_sync(0xffffffffu, li && !gray), rb = __ballot_sync(0xffffffffu, ri && !gray);
    const u32 gb = __ballot_sync(0xffffffffu, gray);
    if ((threadIdx.x & 31) == 0 && i < m) {
        new_lhs[i >> 5] = lb; new_rhs[i >> 5] = rb;
        if (gb) atomicAdd(n_gray, (u64)__popc(gb));
    }
}

}  // namespace

// number of u64 words of a bit vector of m pushed bits (one zero word when nothing was pushed)
u64 bit_vector_words(u64 m) { return m ? (m + 63) / 64 : 1; }

// weight[i] in {1, 2, 3} -> the two membership bit vectors (bit_vector_words(m) words each, allocated here)
void xeno_annotate_bits(Workspace& ws, const u64* weight, u64 m, DevBuf<u64>& lhs, DevBuf<u64>& rhs) {
    const u64 words = bit_vector_words(m);
    lhs.reset(&ws, words); rhs.reset(&ws, words);
    GSB_CUDA_TRY(cudaMemsetAsync(lhs.p, 0, words * 8, ws.stream));
    GSB_CUDA_TRY(cudaMemsetAsync(rhs.p, 0, words * 8, ws.stream));
    if (!m) return;
    annotate_bits_kernel<<<(unsigned)((m + 255) / 256), 256, 0, ws.stream>>>(weight, m, (u32*)lhs.p, (u32*)rhs.p);
    ++ws.launches;
}

// the near-k-mer pass over a decoded kmer set and its membership bits; returns the number of gray k-mers
u64 xeno_near_kmers(Workspace& ws, int key_bytes, int k, const void* keys, u64 m, const u64* lhs, const u64* rhs, DevBuf<u64>& new_lhs, DevBuf<u64>& new_rhs) {
    const u64 words = bit_vector_words(m);
    new_lhs.reset(&ws, words); new_rhs.reset(&ws, words);
    GSB_CUDA_TRY(cudaMemsetAsync(new_lhs.p, 0, words * 8, ws.stream));
    GSB_CUDA_TRY(cudaMemsetAsync(new_rhs.p, 0, words * 8, ws.stream));
    if (!m) return 0;
    DevBuf<u64> gray(&ws, 1);
    GSB_CUDA_TRY(cudaMemsetAsync(gray.p, 0, 8, ws.stream));
    const unsigned grid = (unsigned)((m + 255) / 256);
    if (key_bytes == 8) near_kmers_kernel<u64><<<grid, 256, 0, ws.stream>>>((const u64*)keys, m, k, (const u32*)lhs, (const u32*)rhs, (u32*)new_lhs.p, (u32*)new_rhs.p, gray.p);
    else near_kmers_kernel<Key128><<<grid, 256, 0, ws.stream>>>((const Key128*)keys, m, k, (const u32*)lhs, (const u32*)rhs, (u32*)new_lhs.p, (u32*)new_rhs.p, gray.p);
    ++ws.launches;
    u64 h = 0;
    GSB_CUDA_TRY(cudaMemcpyAsync(&h, gray.p, 8, cudaMemcpyDeviceToHost, ws.stream));
    ws.sync();
    return h;
}

}  // namespace gsb
