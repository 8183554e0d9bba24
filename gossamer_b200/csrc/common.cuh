// common.cuh -- shared device/host helpers for the gossamer_b200 CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "keys.h"

namespace gsb {

// Base-4 digit reversal (what Gossamer::rev computes, src/Utils.hh:377-396), via the
// hardware bit reverse followed by a swap inside each 2-bit group.
__device__ __forceinline__ u64 rev_base4(u64 x) {
    u64 r = __brevll(x);
    return ((r & 0x5555555555555555ull) << 1) | ((r >> 1) & 0x5555555555555555ull);
}

// Reverse complement of a key of w symbols (BigInteger<2>::reverseComplement, src/BigInteger.hh:204-217:
// complement, base-4 reverse each word and swap the words, shift right by the unused bits).
__device__ __forceinline__ u64 key_rc(u64 x, int w) { return rev_base4(~x) >> (64 - 2 * w); }       // 1 <= w <= 32
__device__ __forceinline__ Key128 key_rc(const Key128& x, int w) {                                  // 32 < w <= 63
    const u64 hi = rev_base4(~x.lo), lo = rev_base4(~x.hi);
    const int sh = 128 - 2 * w;                                 // 2 <= sh < 64
    Key128 r;
    r.lo = (lo >> sh) | (hi << (64 - sh));
    r.hi = hi >> sh;
    return r;
}

// ---- decoupled look-back over tiles ---------------------------------------------------------
// state word = status (top 2 bits: 0 empty, 1 tile aggregate, 2 inclusive prefix) | value.
template <typename T> struct LookbackWord;
template <> struct LookbackWord<u32> { static const int kShift = 30; static const u32 kMask = 0x3FFFFFFFu; };
template <> struct LookbackWord<u64> { static const int kShift = 62; static const u64 kMask = 0x3FFFFFFFFFFFFFFFull; };

// Look-back state is read and written with relaxed GPU-scope accesses (status and value live in
// ONE word, so no ordering against other data is needed).  `volatile` compiles to
// LDG/STG.E.STRONG.SYS -- system scope -- which ncu showed as the dominant stall of the
// run-length kernel; .gpu scope stops at the L2.
__device__ __forceinline__ u32 ld_volatile(const u32* p) { u32 v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ u64 ld_volatile(const u64* p) { u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_volatile(u32* p, u32 v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_volatile(u64* p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }

// Called by ONE thread per (tile, bin).  `states` is indexed [tile * stride + bin]; tiles must be
// numbered by an atomic ticket so that every lower tile is already running.
template <typename T>
__device__ __forceinline__ T lookback_exclusive(T* states, u32 stride, u32 tile, u32 bin, T aggregate) {
    const int S = LookbackWord<T>::kShift;
    const T M = LookbackWord<T>::kMask;
    T* mine = states + (size_t)tile * stride + bin;
    if (tile == 0) {
        st_volatile(mine, (T)(((T)2 << S) | aggregate));
        return 0;
    }
    st_volatile(mine, (T)(((T)1 << S) | aggregate));
    T excl = 0;
    for (long long t = (long long)tile - 1;; --t) {
        const T* p = states + (size_t)t * stride + bin;
        T s;
        do { s = ld_volatile(p); } while ((s >> S) == 0);
        excl += s & M;
        if ((s >> S) == 2) break;
    }
    st_volatile(mine, (T)(((T)2 << S) | (excl + aggregate)));
    return excl;
}

// Warp-wide variant for one scalar per tile (stride 1): called by ALL 32 lanes of one warp; 32
// predecessor states are examined per round trip.  Returns the exclusive prefix in every lane.
template <typename T>
__device__ __forceinline__ T lookback_exclusive_warp(T* states, u32 tile, T aggregate) {
    const int S = LookbackWord<T>::kShift;
    const T M = LookbackWord<T>::kMask;
    const int lane = threadIdx.x & 31;
    T* mine = states + tile;
    if (lane == 0) st_volatile(mine, (T)(((T)(tile == 0 ? 2 : 1) << S) | aggregate));
    if (tile == 0) return 0;
    T excl = 0;
    long long t = (long long)tile - 1;
    for (;;) {
        const long long idx = t - lane;
        const T sv = idx >= 0 ? ld_volatile(states + idx) : (T)((T)2 << S);
        const u32 status = (u32)(sv >> S);
        const u32 zero_mask = __ballot_sync(0xffffffffu, status == 0);
        const u32 pref_mask = __ballot_sync(0xffffffffu, status == 2);
        const int first_zero = zero_mask ? __ffs(zero_mask) - 1 : 32;
        const int first_pref = pref_mask ? __ffs(pref_mask) - 1 : 32;
        const int take = first_pref < first_zero ? first_pref + 1 : first_zero;   // lanes [0, take) contribute
        T v = lane < take ? (sv & M) : (T)0;
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (first_pref < first_zero) break;
        t -= take;                                              // re-poll from the first unpublished tile (or go 32 further back)
    }
    if (lane == 0) st_volatile(mine, (T)(((T)2 << S) | (excl + aggregate)));
    return excl;
}

// block-wide exclusive scan of one value per thread (THREADS multiple of 32, <= 1024)
template <typename T, int THREADS>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* smem /* THREADS/32 + 1 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < THREADS / 32 ? smem[lane] : (T)0;
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        if (lane < THREADS / 32) smem[lane] = winc - w;
        if (lane == 31) smem[THREADS / 32] = winc;
    }
    __syncthreads();
    T res = smem[warp] + inc - v;
    if (total) *total = smem[THREADS / 32];
    __syncthreads();
    return res;
}

static inline int ceil_div_i(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace gsb

#define GSB_CUDA_TRY(expr)                                                                     \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) throw gsb::CudaError(_e, #expr, __FILE__, __LINE__);            \
    } while (0)

namespace gsb {
struct CudaError {
    cudaError_t code; const char* expr; const char* file; int line;
    CudaError(cudaError_t c, const char* e, const char* f, int l) : code(c), expr(e), file(f), line(l) {}
};
}  // namespace gsb
