// keys.h -- key types and pure key arithmetic shared by device and host code (no device intrinsics in
// here: tests/cpp/mix_check.cc compiles it with the host compiler alone).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

namespace gsb {

// 128-bit key for k+1 > 32 (the reference always uses BigInteger<2>, src/RankSelect.hh:72;
// a 64-bit key is the specialisation for 2*rho <= 64).
struct __align__(16) Key128 {
    u64 lo, hi;
};

template <typename K> struct KeyOps;

template <> struct KeyOps<u64> {
    static const int kBytes = 8;
    __host__ __device__ static __forceinline__ u32 digit(u64 k, int shift) { return (u32)(k >> shift) & 0xFFu; }
    __host__ __device__ static __forceinline__ bool eq(u64 a, u64 b) { return a == b; }
    __host__ __device__ static __forceinline__ bool lt(u64 a, u64 b) { return a < b; }
    __host__ __device__ static __forceinline__ u64 shr64(u64 k, int d) { return d >= 64 ? 0ull : (k >> d); }  // (k >> d) low 64 bits
    __host__ __device__ static __forceinline__ u64 lo(u64 k) { return k; }
    __host__ __device__ static __forceinline__ u64 hi(u64) { return 0; }
    __host__ __device__ static __forceinline__ u64 make(u64 lo, u64) { return lo; }
};

template <> struct KeyOps<Key128> {
    static const int kBytes = 16;
    __host__ __device__ static __forceinline__ u32 digit(const Key128& k, int shift) {
        return shift < 64 ? (u32)(k.lo >> shift) & 0xFFu : (u32)(k.hi >> (shift - 64)) & 0xFFu;   // digits are byte aligned
    }
    __host__ __device__ static __forceinline__ bool eq(const Key128& a, const Key128& b) { return a.lo == b.lo && a.hi == b.hi; }
    __host__ __device__ static __forceinline__ bool lt(const Key128& a, const Key128& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
    __host__ __device__ static __forceinline__ u64 shr64(const Key128& k, int d) {
        if (d == 0) return k.lo;
        if (d < 64) return (k.lo >> d) | (k.hi << (64 - d));
        if (d < 128) return k.hi >> (d - 64);
        return 0;
    }
    __host__ __device__ static __forceinline__ u64 lo(const Key128& k) { return k.lo; }
    __host__ __device__ static __forceinline__ u64 hi(const Key128& k) { return k.hi; }
    __host__ __device__ static __forceinline__ Key128 make(u64 lo, u64 hi) { Key128 k; k.lo = lo; k.hi = hi; return k; }
};

// Bijective bit mixing of a key, used when instances are only GROUPED (counting from a partial sort,
// sort.cu): every bit of the mixed key depends on every base of the window, so that keys which differ by
// one or two substitutions -- a true k-mer and its sequencing-error variants share most of their bases --
// fall into the same group of equal low bits only by chance.  The splitmix64 finaliser (two odd
// multiplications, three xor-shifts) and its exact inverse; the high word of a 128-bit key is folded into
// the low word and left unchanged itself.  (The first version, x ^ (x >> 32), is linear: two variants of one
// k-mer with the same substitution pattern 16 bases apart collide in the low 40 bits -- 1.5 M groups with
// several keys at config 2 instead of the ~1.5 k that chance allows.)
__host__ __device__ __forceinline__ u64 key_mix(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ u64 key_unmix(u64 z) {
    z = z ^ (z >> 31) ^ (z >> 62);
    z *= 0x319642B2D24D8EC3ull;                                 // inverse of 0x94D049BB133111EB mod 2^64
    z = z ^ (z >> 27) ^ (z >> 54);
    z *= 0x96DE1B173F119089ull;                                 // inverse of 0xBF58476D1CE4E5B9 mod 2^64
    return z ^ (z >> 30) ^ (z >> 60);
}
__host__ __device__ __forceinline__ Key128 key_mix(const Key128& x) { Key128 r; r.hi = x.hi; r.lo = key_mix(x.lo ^ key_mix(x.hi)); return r; }
__host__ __device__ __forceinline__ Key128 key_unmix(const Key128& x) { Key128 r; r.hi = x.hi; r.lo = key_unmix(x.lo) ^ key_mix(x.hi); return r; }

}  // namespace gsb
