// scan.cuh -- device-wide exclusive scan (reduce / scan tile sums / rescan) used off the hot
// path: line tables, block offsets of the select directories, count-plane compaction.
#pragma once
#include "common.cuh"

namespace gsb {

static const int kScanThreads = 256;
static const int kScanItems = 8;
static const int kScanTile = kScanThreads * kScanItems;

template <typename In, typename Out>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const In* __restrict__ in, u64 n, Out* __restrict__ tile_sums) {
    __shared__ Out sm[kScanThreads / 32 + 1];
    const u64 base = (u64)blockIdx.x * kScanTile;
    Out v = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        u64 idx = base + (u64)i * kScanThreads + threadIdx.x;
        if (idx < n) v += (Out)in[idx];
    }
    Out total;
    block_exclusive_scan<Out, kScanThreads>(v, &total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one CTA: in-place exclusive scan of `n` values, grand total to *total
template <typename T>
__global__ void __launch_bounds__(1024) scan_single_cta_kernel(T* data, u64 n, T* total) {
    __shared__ T sm[1024 / 32 + 1];
    __shared__ T carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (u64 base = 0; base < n; base += 1024) {
        u64 idx = base + threadIdx.x;
        T v = idx < n ? data[idx] : (T)0;
        T tot;
        T ex = block_exclusive_scan<T, 1024>(v, &tot, sm);
        T carry = carry_s;
        if (idx < n) data[idx] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry_s;
}

template <typename In, typename Out>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const In* __restrict__ in, u64 n, const Out* __restrict__ tile_offsets,
                                                                 Out* __restrict__ out, Out bias) {
    __shared__ Out sm[kScanThreads / 32 + 1];
    const u64 base = (u64)blockIdx.x * kScanTile + (u64)threadIdx.x * kScanItems;   // blocked arrangement keeps order
    Out v[kScanItems];
    Out sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        u64 idx = base + i;
        v[i] = idx < n ? (Out)in[idx] : (Out)0;
        sum += v[i];
    }
    Out ex = block_exclusive_scan<Out, kScanThreads>(sum, (Out*)nullptr, sm) + tile_offsets[blockIdx.x] + bias;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        u64 idx = base + i;
        if (idx < n) out[idx] = ex;
        ex += v[i];
    }
}

// out[i] = bias + sum(in[0..i)), *total = sum(in[0..n)) (device pointer, may be null).
// tmp must hold ceil(n / kScanTile) + 1 Out values.  `in` and `out` may alias only if In == Out.
template <typename In, typename Out>
static inline void exclusive_scan(const In* in, Out* out, u64 n, Out bias, Out* total, Out* tmp, cudaStream_t s, u64* launches) {
    if (n == 0) {
        if (total) cudaMemsetAsync(total, 0, sizeof(Out), s);
        return;
    }
    u64 tiles = (n + kScanTile - 1) / kScanTile;
    scan_tile_sums_kernel<In, Out><<<(unsigned)tiles, kScanThreads, 0, s>>>(in, n, tmp);
    scan_single_cta_kernel<Out><<<1, 1024, 0, s>>>(tmp, tiles, total);
    scan_apply_kernel<In, Out><<<(unsigned)tiles, kScanThreads, 0, s>>>(in, n, tmp, out, bias);
    if (launches) *launches += 3;
}

static inline u64 scan_tmp_elems(u64 n) { return (n + kScanTile - 1) / kScanTile + 1; }

}  // namespace gsb
