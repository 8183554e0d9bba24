// emit.cu -- sorted distinct (key,count) -> the reference's succinct on-disk structures,
// byte for byte, built with data-parallel kernels instead of sequential push_back writers.
//
//   K7  SparseArray: high-bits bitmap + low-bits byte planes      (src/SparseArray.hh:87-118, src/SparseArray.cc:75-103,
//                                                                   src/IntegerArray.cc:259-357, src/WordyBitVector.hh:54-134)
//   K8  DenseSelect directories over the ones (d1) and zeros (d0) (src/DenseArray.cc:446-694, src/DenseArray.hh:82-136)
//   K9  VariableByteArray planes + presence sets, count histogram (src/VariableByteArray.hh:81-103, src/Graph.cc:127-133)
//
// Closed forms used: i-th one at h_i = (e_i >> D) + i; j-th zero at
// z_j = j + #{i : (e_i >> D) <= j}; bitmap length floor((nd + M + 3)/64) + 1 words; a select
// block's class and size depend only on the positions it indexes, so sizes -> exclusive scan ->
// file offsets -> scatter.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "kernels.h"
#include "scan.cuh"

namespace gsb {

typedef unsigned __int128 u128_t;
static inline u128_t to128(U128 v) { return ((u128_t)v.hi << 64) | v.lo; }

u64 sparse_array_d(U128 universe, u64 m_est) {
    // SparseArray::Builder::d, src/SparseArray.cc:47-72; BigInteger::asDouble, src/BigInteger.hh:181-190
    double n = (double)universe.hi * 18446744073709551616.0 + (double)universe.lo;
    double m = (double)m_est;
    double d0 = log2(n / ((1 + m) * 1.4426950408889634));
    u64 d = (u64)ceil(d0);
    if (d < 8) d = 8; else if (d > 128) d = 128;
    return d;
}

// ------------------------------------------------------------------------------------------
// Emitter: device buffer -> pinned staging -> sink
// ------------------------------------------------------------------------------------------
static void sink_fail(const std::string& what, const std::string& name) { throw StatusError{GSB_EIO, what + " failed for " + name}; }

void Emitter::put_host(const std::string& name, const void* data, u64 len) {
    bytes_out += len;
    if (!sink) return;
    void* h = nullptr;
    if (sink->open(sink->user, name.c_str(), len, &h) != 0) sink_fail("open", name);
    if (len && sink->pwrite(sink->user, h, 0, data, len) != 0) sink_fail("pwrite", name);
    if (sink->close(sink->user, h) != 0) sink_fail("close", name);
}

void Emitter::put_device(const std::string& name, const void* dev, u64 len, const void* host_prefix, u64 prefix_len) {
    bytes_out += len;
    if (!sink) return;
    void* h = nullptr;
    if (sink->open(sink->user, name.c_str(), len, &h) != 0) sink_fail("open", name);
    for (u64 off = 0; off < len; off += pinned_bytes) {
        u64 chunk = std::min<u64>(pinned_bytes, len - off);
        GSB_CUDA_TRY(cudaMemcpyAsync(pinned, (const u8*)dev + off, chunk, cudaMemcpyDeviceToHost, ws->stream));
        ws->sync();
        if (off < prefix_len) memcpy(pinned, (const u8*)host_prefix + off, std::min<u64>(prefix_len - off, chunk));
        if (sink->pwrite(sink->user, h, off, pinned, chunk) != 0) sink_fail("pwrite", name);
    }
    if (sink->close(sink->user, h) != 0) sink_fail("close", name);
}

// ------------------------------------------------------------------------------------------
// K8: DenseSelect
// ------------------------------------------------------------------------------------------
enum { T_SMALL = 0, T_SPILL64 = 1, T_SPILL32 = 2, T_SPILL16 = 3, T_SPILL8 = 4, T_INTERMEDIATE = 5 };
static const u64 kBlock = 8192;
static const u64 kSample = 64;
static const u32 kSamplesPerBlock = 128;

template <typename K> struct OnesPos {
    const K* keys; int D;
    __device__ __forceinline__ u64 operator()(u64 j) const { return KeyOps<K>::shr64(keys[j], D) + j; }
};
template <typename K> struct ZerosPos {
    const K* keys; u64 m; int D;
    __device__ __forceinline__ u64 operator()(u64 j) const {
        u64 lo = 0, hi = m;                                     // #{i : (e_i >> D) <= j}
        while (lo < hi) { u64 mid = lo + ((hi - lo) >> 1); if (KeyOps<K>::shr64(keys[mid], D) <= j) lo = mid + 1; else hi = mid; }
        return j + lo;
    }
};

__device__ __forceinline__ u32 sub_block_bytes(u64 sub_span) {
    if (sub_span <= (kBlock >> 6)) return 0;                    // bit-scan
    if (sub_span < (1ull << 8)) return 64;
    if (sub_span < (1ull << 16)) return 128;
    return 256;
}
__device__ __forceinline__ u32 sub_block_type(u64 sub_span) {
    if (sub_span < (1ull << 8)) return T_SPILL8;
    if (sub_span < (1ull << 16)) return T_SPILL16;
    return T_SPILL32;
}

template <typename F>
__global__ void ds_classify_kernel(F f, u64 count, u64 n_blocks, u8* __restrict__ type, u64* __restrict__ bytes,
                                   u64* __restrict__ padded, u64* __restrict__ first_out) {
    for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (u64)gridDim.x * blockDim.x) {
        const u64 j0 = b * kBlock;
        const u64 nb = count - j0 < kBlock ? count - j0 : kBlock;
        const u64 first = f(j0), last = f(j0 + nb - 1), span = last - first;
        u8 t; u64 sz;
        if (span >= (1ull << 24) || nb < kBlock) {
            if (span < (1ull << 32)) { t = T_SPILL32; sz = nb * 4; } else { t = T_SPILL64; sz = nb * 8; }
        } else if (span >= (1ull << 16)) {
            t = T_INTERMEDIATE; sz = kSamplesPerBlock * 6;
            for (u32 s = 0; s < kSamplesPerBlock; ++s) sz += sub_block_bytes(f(j0 + s * kSample + kSample - 1) - f(j0 + s * kSample));
        } else { t = T_SMALL; sz = kSamplesPerBlock * 2; }
        type[b] = t; bytes[b] = sz; padded[b] = (sz + 7) & ~7ull; first_out[b] = first;
    }
}

// stats[0..5] = smallBlocks, smallBlocksSize, intermediateBlocks, intermediateBlocksSize, largeBlocks, largeBlocksSize
__global__ void ds_stats_kernel(const u8* __restrict__ type, const u64* __restrict__ bytes, u64 n_blocks, u64* __restrict__ stats) {
    u64 v[6] = {0, 0, 0, 0, 0, 0};
    for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (u64)gridDim.x * blockDim.x) {
        int c = type[b] == T_SMALL ? 0 : (type[b] == T_INTERMEDIATE ? 2 : 4);
        v[c] += 1; v[c + 1] += bytes[b];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        u64 x = v[i];
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(&stats[i], x);
    }
}

// one CTA of 128 threads per select block
template <typename F>
__global__ void __launch_bounds__(128) ds_write_kernel(F f, u64 count, const u8* __restrict__ type, const u64* __restrict__ off,
                                                       const u64* __restrict__ first_in, u8* __restrict__ file,
                                                       u64 index_off, u64 rank_off) {
    __shared__ u32 scan_s[128 / 32 + 1];
    const u64 b = blockIdx.x;
    const u64 j0 = b * kBlock;
    const u64 nb = count - j0 < kBlock ? count - j0 : kBlock;
    const u64 first = first_in[b];
    const u8 t = type[b];
    u8* out = file + off[b];
    if (threadIdx.x == 0) {
        reinterpret_cast<u64*>(file + index_off)[b] = off[b] | t;
        reinterpret_cast<u64*>(file + rank_off)[b] = first;
    }
    if (t == T_SMALL) {
        reinterpret_cast<u16*>(out)[threadIdx.x] = (u16)(f(j0 + threadIdx.x * kSample) - first);
    } else if (t == T_SPILL32) {
        for (u64 j = threadIdx.x; j < nb; j += 128) reinterpret_cast<u32*>(out)[j] = (u32)(f(j0 + j) - first);
    } else if (t == T_SPILL64) {
        for (u64 j = threadIdx.x; j < nb; j += 128) reinterpret_cast<u64*>(out)[j] = f(j0 + j);       // absolute (src/DenseArray.cc:484-492)
    } else {                                                                                         // T_INTERMEDIATE
        const u32 s = threadIdx.x;
        const u64 a = f(j0 + s * kSample);
        const u64 sub_span = f(j0 + s * kSample + kSample - 1) - a;
        const u32 sz = sub_block_bytes(sub_span);
        const u32 ex = block_exclusive_scan<u32, 128>(sz, (u32*)nullptr, scan_s);
        const u32 sub_base = kSamplesPerBlock * 6 + ex;
        reinterpret_cast<u32*>(out)[s] = (u32)(a - first);
        reinterpret_cast<u16*>(out + kSamplesPerBlock * 4)[s] = sz ? (u16)(sub_base | sub_block_type(sub_span)) : (u16)0;
        if (sz == 64) { for (u32 j = 0; j < kSample; ++j) out[sub_base + j] = (u8)(f(j0 + s * kSample + j) - a); }
        else if (sz == 128) { for (u32 j = 0; j < kSample; ++j) reinterpret_cast<u16*>(out + sub_base)[j] = (u16)(f(j0 + s * kSample + j) - a); }
        else if (sz == 256) { for (u32 j = 0; j < kSample; ++j) reinterpret_cast<u32*>(out + sub_base)[j] = (u32)(f(j0 + s * kSample + j) - a); }
    }
}

struct DsHeader {
    u64 version, flags, indexArrayOffset, rankArrayOffset, logBlockSize, blockSize, logSampleRate, sampleRate;
    u64 numBlocks, indexSize, smallBlocks, smallBlocksSize, intermediateBlocks, intermediateBlocksSize, largeBlocks, largeBlocksSize;
};

template <typename F>
static void build_dense_select(Emitter& em, F f, u64 count, bool invert, const std::string& name) {
    Workspace& ws = *em.ws;
    cudaStream_t s = ws.stream;
    DsHeader h;
    memset(&h, 0, sizeof(h));
    h.version = 2012092701ull; h.flags = invert ? 1 : 0;
    h.logBlockSize = 13; h.blockSize = 8192; h.logSampleRate = 6; h.sampleRate = 64;
    const u64 nb = (count + kBlock - 1) / kBlock;
    if (nb == 0) {
        h.indexArrayOffset = h.rankArrayOffset = 4096;
        std::vector<u8> page(4096, 0);
        memcpy(page.data(), &h, sizeof(h));
        em.put_host(name, page.data(), page.size());
        return;
    }
    DevBuf<u8> type(&ws, nb);
    DevBuf<u64> bytes(&ws, nb), padded(&ws, nb), first(&ws, nb), off(&ws, nb), tmp(&ws, scan_tmp_elems(nb)), scalars(&ws, 8);
    GSB_CUDA_TRY(cudaMemsetAsync(scalars.p, 0, 64, s));
    const int g = (int)std::min<u64>((nb + 127) / 128, (u64)ws.sm_count * 8);
    ds_classify_kernel<F><<<g, 128, 0, s>>>(f, count, nb, type.p, bytes.p, padded.p, first.p);
    ++ws.launches;
    exclusive_scan<u64, u64>(padded.p, off.p, nb, 4096ull, scalars.p + 6, tmp.p, s, &ws.launches);
    ds_stats_kernel<<<g, 128, 0, s>>>(type.p, bytes.p, nb, scalars.p);
    ++ws.launches;
    u64 host[8];
    GSB_CUDA_TRY(cudaMemcpyAsync(host, scalars.p, 64, cudaMemcpyDeviceToHost, s));
    ws.sync();
    const u64 body_end = 4096 + host[6];
    h.indexArrayOffset = (body_end + 15) & ~15ull;
    h.rankArrayOffset = h.indexArrayOffset + 8 * nb;
    const u64 file_size = h.rankArrayOffset + 8 * nb;
    h.numBlocks = nb; h.indexSize = 16 * nb;
    h.smallBlocks = host[0]; h.smallBlocksSize = host[1];
    h.intermediateBlocks = host[2]; h.intermediateBlocksSize = host[3];
    h.largeBlocks = host[4]; h.largeBlocksSize = host[5];
    DevBuf<u8> file(&ws, file_size);
    GSB_CUDA_TRY(cudaMemsetAsync(file.p, 0, file_size, s));
    ds_write_kernel<F><<<(unsigned)nb, 128, 0, s>>>(f, count, type.p, off.p, first.p, file.p, h.indexArrayOffset, h.rankArrayOffset);
    ++ws.launches;
    em.put_device(name, file.p, file_size, &h, sizeof(h));
}

// ------------------------------------------------------------------------------------------
// K7: SparseArray
// ------------------------------------------------------------------------------------------
template <typename K>
__global__ void high_bits_kernel(const K* __restrict__ keys, u64 m, int D, u64* __restrict__ bitmap) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        const u64 h = KeyOps<K>::shr64(keys[i], D) + i;
        atomicOr(&bitmap[h >> 6], 1ull << (h & 63));
    }
}

// one plane of the IntegerArray: `bytes`-wide little-endian pieces of ((key & DMask) >> shift)
template <typename K, typename T>
__global__ void low_plane_kernel(const K* __restrict__ keys, u64 m, int D, int shift, T* __restrict__ out) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        K k = keys[i];
        u64 lo = KeyOps<K>::lo(k), hi = KeyOps<K>::hi(k);
        if (D < 64) { lo &= (1ull << D) - 1; hi = 0; }
        else if (D < 128) { hi = D == 64 ? 0 : (hi & ((1ull << (D - 64)) - 1)); }
        u64 piece;
        if (shift == 0) piece = lo;
        else if (shift < 64) piece = (lo >> shift) | (hi << (64 - shift));
        else piece = hi >> (shift - 64);
        out[i] = (T)piece;
    }
}

struct PlaneSpec { const char* suffix; int shift; int bytes; };

// IntegerArray::builder's nesting of StackedArrays, src/IntegerArray.cc:259-357
static std::vector<PlaneSpec> integer_array_planes(u64 bits) {
    switch (bits) {
        case 8:   return {{"", 0, 1}};
        case 16:  return {{"", 0, 2}};
        case 24:  return {{".upr", 16, 1}, {".lwr", 0, 2}};
        case 32:  return {{"", 0, 4}};
        case 40:  return {{".upr", 32, 1}, {".lwr", 0, 4}};
        case 48:  return {{".upr", 32, 2}, {".lwr", 0, 4}};
        case 56:  return {{".upr", 48, 1}, {".lwr.upr", 32, 2}, {".lwr.lwr", 0, 4}};
        case 64:  return {{"", 0, 8}};
        case 72:  return {{".upr", 64, 1}, {".lwr", 0, 8}};
        case 80:  return {{".upr", 64, 2}, {".lwr", 0, 8}};
        case 88:  return {{".upr", 80, 1}, {".lwr.upr", 64, 2}, {".lwr.lwr", 0, 8}};
        case 96:  return {{".upr", 64, 4}, {".lwr", 0, 8}};
        case 104: return {{".upr", 96, 1}, {".lwr.upr", 64, 4}, {".lwr.lwr", 0, 8}};
        case 112: return {{".upr", 96, 2}, {".lwr.upr", 64, 4}, {".lwr.lwr", 0, 8}};
        case 120: return {{".upr.upr", 112, 1}, {".upr.lwr", 96, 2}, {".lwr.upr", 64, 4}, {".lwr.lwr", 0, 8}};
        case 128: return {{".upr", 64, 8}, {".lwr", 0, 8}};
        default: throw StatusError{GSB_EINVAL, "IntegerArray::builder: unsupported integer width " + std::to_string(bits)};
    }
}

template <typename K>
static void emit_sparse_array_t(Emitter& em, const K* keys, u64 m, U128 universe_ctor, u64 m_est, U128 universe_end, const std::string& base) {
    Workspace& ws = *em.ws;
    cudaStream_t s = ws.stream;
    const u64 D = sparse_array_d(universe_ctor, m_est);
    const u64 qD = 8 * ((D + 7) / 8);
    const u128_t nd128 = D >= 128 ? (u128_t)0 : (to128(universe_end) >> D);
    if ((u64)(nd128 >> 64)) throw StatusError{GSB_EINVAL, "Internal error in SparseArray; nd does not fit 64 bits"};
    const u64 nd = (u64)nd128;
    const int g = (int)std::max<u64>(1, std::min<u64>((m + 255) / 256, (u64)ws.sm_count * 16));

    // high bits
    const u64 words = (nd + m + 3) / 64 + 1;
    {
        DevBuf<u64> bitmap(&ws, words);
        GSB_CUDA_TRY(cudaMemsetAsync(bitmap.p, 0, words * 8, s));
        if (m) { high_bits_kernel<K><<<g, 256, 0, s>>>(keys, m, (int)D, bitmap.p); ++ws.launches; }
        em.put_device(base + ".high-bits", bitmap.p, words * 8);
    }
    // select directories
    build_dense_select(em, ZerosPos<K>{keys, m, (int)D}, nd + 2, true, base + "-d0");
    build_dense_select(em, OnesPos<K>{keys, (int)D}, m, false, base + "-d1");
    // low bits
    for (const PlaneSpec& p : integer_array_planes(qD)) {
        DevBuf<u8> plane(&ws, m * p.bytes);
        if (m) {
            switch (p.bytes) {
                case 1: low_plane_kernel<K, u8><<<g, 256, 0, s>>>(keys, m, (int)D, p.shift, (u8*)plane.p); break;
                case 2: low_plane_kernel<K, u16><<<g, 256, 0, s>>>(keys, m, (int)D, p.shift, (u16*)plane.p); break;
                case 4: low_plane_kernel<K, u32><<<g, 256, 0, s>>>(keys, m, (int)D, p.shift, (u32*)plane.p); break;
                default: low_plane_kernel<K, u64><<<g, 256, 0, s>>>(keys, m, (int)D, p.shift, (u64*)plane.p); break;
            }
            ++ws.launches;
        }
        em.put_device(base + ".low-bits" + p.suffix, plane.p, m * p.bytes);
    }
    // header (src/SparseArray.hh:60-72)
    struct { u64 version, D, quantizedD, dmask[2], size[2], count; } hd;
    hd.version = 2012030501ull; hd.D = D; hd.quantizedD = qD;
    u128_t mask = D >= 128 ? ~(u128_t)0 : ((((u128_t)1) << D) - 1);
    hd.dmask[0] = (u64)mask; hd.dmask[1] = (u64)(mask >> 64);
    hd.size[0] = universe_end.lo; hd.size[1] = universe_end.hi;
    hd.count = m;
    em.put_host(base + ".header", &hd, sizeof(hd));
}

void emit_sparse_array(Emitter& em, int key_bytes, const void* keys, u64 m, U128 universe_ctor, u64 m_est, U128 universe_end,
                       const std::string& base) {
    if (key_bytes == 8) emit_sparse_array_t<u64>(em, (const u64*)keys, m, universe_ctor, m_est, universe_end, base);
    else emit_sparse_array_t<Key128>(em, (const Key128*)keys, m, universe_ctor, m_est, universe_end, base);
}

// ------------------------------------------------------------------------------------------
// K9: VariableByteArray + histogram
// ------------------------------------------------------------------------------------------
__global__ void vba_flags_kernel(const u64* __restrict__ counts, u64 m, u8* __restrict__ ord0, u8* __restrict__ f1, u8* __restrict__ f2) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        const u32 c = (u32)counts[i];                           // value_type is uint32_t (src/VariableByteArray.hh:72)
        ord0[i] = (u8)(c & 0xFF);
        f1[i] = (c >> 8) ? 1 : 0;
        f2[i] = (c >> 16) ? 1 : 0;
    }
}

__global__ void vba_scatter_kernel(const u64* __restrict__ counts, u64 m, const u64* __restrict__ r1, const u64* __restrict__ r2,
                                   u64* __restrict__ ord1_pos, u8* __restrict__ ord1, u64* __restrict__ ord2_pos, u16* __restrict__ ord2) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        const u32 c = (u32)counts[i];
        if (c >> 8) {
            const u64 a = r1[i];
            ord1_pos[a] = i; ord1[a] = (u8)((c >> 8) & 0xFF);
            if (c >> 16) { const u64 b = r2[i]; ord2_pos[b] = a; ord2[b] = (u16)(c >> 16); }
        }
    }
}

void emit_counts(Emitter& em, const u64* counts, u64 m, u64 m_est, const std::string& base) {
    Workspace& ws = *em.ws;
    cudaStream_t s = ws.stream;
    const int g = (int)std::max<u64>(1, std::min<u64>((m + 255) / 256, (u64)ws.sm_count * 16));
    DevBuf<u8> ord0(&ws, m), f1(&ws, m), f2(&ws, m);
    DevBuf<u64> r1(&ws, m), r2(&ws, m), tmp(&ws, scan_tmp_elems(m)), totals(&ws, 2);
    u64 n1 = 0, n2 = 0;
    if (m) {
        vba_flags_kernel<<<g, 256, 0, s>>>(counts, m, ord0.p, f1.p, f2.p);
        ++ws.launches;
        exclusive_scan<u8, u64>(f1.p, r1.p, m, 0ull, totals.p, tmp.p, s, &ws.launches);
        exclusive_scan<u8, u64>(f2.p, r2.p, m, 0ull, totals.p + 1, tmp.p, s, &ws.launches);
        u64 t[2];
        GSB_CUDA_TRY(cudaMemcpyAsync(t, totals.p, 16, cudaMemcpyDeviceToHost, s));
        ws.sync();
        n1 = t[0]; n2 = t[1];
    }
    DevBuf<u64> ord1_pos(&ws, n1), ord2_pos(&ws, n2);
    DevBuf<u8> ord1(&ws, n1);
    DevBuf<u16> ord2(&ws, n2);
    if (n1) { vba_scatter_kernel<<<g, 256, 0, s>>>(counts, m, r1.p, r2.p, ord1_pos.p, ord1.p, ord2_pos.p, ord2.p); ++ws.launches; }
    em.put_device(base + ".ord0", ord0.p, m);
    em.put_device(base + ".ord1", ord1.p, n1);
    em.put_device(base + ".ord2", ord2.p, n2 * 2);
    f1.free(); f2.free(); r1.free(); r2.free();
    // presence sets: both constructed for N = numItems, M = floor(0.001 * numItems); ended with the number of
    // values pushed at that level (src/VariableByteArray.cc:21-43)
    const U128 n_ctor{m_est, 0};
    const u64 m_frac = (u64)(m_est * 0.001);
    emit_sparse_array(em, 8, ord1_pos.p, n1, n_ctor, m_frac, U128{m, 0}, base + ".ord1p");
    emit_sparse_array(em, 8, ord2_pos.p, n2, n_ctor, m_frac, U128{n1, 0}, base + ".ord2p");
}

void emit_count_histogram(Emitter& em, const u64* counts, u64 m, const std::string& name) {
    // histogram = run-length reduce of the sorted counts (uses the 64-bit count: src/Graph.hh:101-106)
    Workspace& ws = *em.ws;
    cudaStream_t s = ws.stream;
    std::string text;
    if (m) {
        DevBuf<u64> a(&ws, m), b(&ws, m);
        GSB_CUDA_TRY(cudaMemcpyAsync(a.p, counts, m * 8, cudaMemcpyDeviceToDevice, s));
        int passes = 0;
        int where = sort_keys(ws, 8, 64, a.p, b.p, nullptr, nullptr, m, nullptr, &passes);
        ReducedRun run; u64 distinct = 0;
        reduce_sorted(ws, 8, where ? b.p : a.p, nullptr, m, 1, run, &distinct);
        std::vector<u64> vals(run.m), freq(run.m);
        GSB_CUDA_TRY(cudaMemcpyAsync(vals.data(), run.keys.p, run.m * 8, cudaMemcpyDeviceToHost, s));
        GSB_CUDA_TRY(cudaMemcpyAsync(freq.data(), run.counts.p, run.m * 8, cudaMemcpyDeviceToHost, s));
        ws.sync();
        for (u64 i = 0; i < run.m; ++i) text += std::to_string(vals[i]) + "\t" + std::to_string(freq[i]) + "\n";
    }
    em.put_host(name, text.data(), text.size());
}

}  // namespace gsb
