// reader.cu -- the reference's succinct Graph / KmerSet file sets back into sorted (key, count) runs on the device.
//
// Replaces (result-wise) the READ side that trim-graph, merge-graphs, merge-kmer-sets and dump-graph are built on:
//   Graph::LazyIterator                      src/Graph.cc:195-216, src/Graph.hh
//   SparseArray::LazyIterator (Elias-Fano)   src/SparseArray.hh:185-224, src/SparseArray.cc  -- key_i = ((select1(i) - i) << D) | low_i
//   WordyBitVector::LazyIterator1            src/WordyBitVector.tcc:17-54                     -- positions of the one bits
//   IntegerArray planes                      src/IntegerArray.cc:259-357, src/StackedArray.hh:217-245
//   VariableByteArray::operator[]            src/VariableByteArray.hh:227-247                 -- ord0 | ord1 << 8 | ord2 << 16
// and the text form of dump-graph            src/GossCmdDumpGraph.cc:31-60.
//
// The reference walks these structures with sequential iterators (one select / rank at a time).  Here a file set is
// decoded in bulk: one popcount per bitmap word + one exclusive scan gives the rank of every one bit, so the i-th key is
// assembled by the thread that owns its bitmap word; the select directories (-d0 / -d1) are not even read.
#include <algorithm>
#include <cstring>

#include "kernels.h"
#include "scan.cuh"

namespace gsb {

namespace {

__global__ void popc_words_kernel(const u64* __restrict__ words, u64 n_words, u32* __restrict__ pc) {
    for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (u64)gridDim.x * blockDim.x) pc[w] = (u32)__popcll(words[w]);
}

struct Planes {
    const u8* p[4];
    int shift[4];
    int bytes[4];
    int n;
};

__device__ __forceinline__ void load_low(const Planes& pl, u64 i, u64& lo, u64& hi) {
    lo = 0; hi = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (q >= pl.n) break;
        u64 v;
        switch (pl.bytes[q]) {
            case 1: v = pl.p[q][i]; break;
            case 2: v = reinterpret_cast<const u16*>(pl.p[q])[i]; break;
            case 4: v = reinterpret_cast<const u32*>(pl.p[q])[i]; break;
            default: v = reinterpret_cast<const u64*>(pl.p[q])[i]; break;
        }
        const int sh = pl.shift[q];
        if (sh < 64) { lo |= v << sh; if (sh && pl.bytes[q] * 8 + sh > 64) hi |= v >> (64 - sh); }
        else hi |= v << (sh - 64);
    }
}

// one thread per bitmap word: every one bit at position h with rank i (ranks from the scanned popcounts) is element i,
// whose high part is h - i
template <typename K>
__global__ void ef_decode_kernel(const u64* __restrict__ words, u64 n_words, const u64* __restrict__ rank_before, Planes pl, int D, u64 m,
                                 K* __restrict__ keys) {
    for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (u64)gridDim.x * blockDim.x) {
        u64 bits = words[w];
        u64 i = rank_before[w];
        while (bits && i < m) {
            const int b = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            const u64 hpart = w * 64 + (u64)b - i;
            u64 lo, hi;
            load_low(pl, i, lo, hi);
            // key = (hpart << D) | low
            if (D < 64) { hi |= D ? (hpart >> (64 - D)) : 0ull; lo |= hpart << D; }
            else if (D < 128) hi |= hpart << (D - 64);
            keys[i] = KeyOps<K>::make(lo, hi);
            ++i;
        }
    }
}

__global__ void vba_base_kernel(const u8* __restrict__ ord0, u64 m, u64* __restrict__ counts) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) counts[i] = ord0[i];
}
__global__ void vba_ord1_kernel(const u64* __restrict__ pos1, const u8* __restrict__ ord1, u64 n1, u64 m, u64* __restrict__ counts) {
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n1; r += (u64)gridDim.x * blockDim.x) {
        const u64 i = pos1[r];
        if (i < m) counts[i] |= (u64)ord1[r] << 8;
    }
}
__global__ void vba_ord2_kernel(const u64* __restrict__ pos1, const u64* __restrict__ pos2, const u16* __restrict__ ord2, u64 n2, u64 n1, u64 m,
                                u64* __restrict__ counts) {
    for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < n2; q += (u64)gridDim.x * blockDim.x) {
        const u64 r = pos2[q];
        if (r < n1) { const u64 i = pos1[r]; if (i < m) counts[i] |= (u64)ord2[q] << 16; }
    }
}
__global__ void fill_value_kernel(u64* __restrict__ counts, u64 m, u64 v) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) counts[i] = v;
}

// ---- dump-graph text: "<k+1 bases>\t<count>\n" per edge -----------------------------------------------------------
__device__ __forceinline__ u32 dec_digits(u64 v) { u32 d = 1; while (v >= 10) { v /= 10; ++d; } return d; }

__global__ void dump_len_kernel(const u64* __restrict__ counts, u64 m, u32 w, u32* __restrict__ len) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x)
        len[i] = w + 2 + dec_digits((u32)counts[i]);               // Graph::Iterator hands out the stored 32-bit multiplicity
}

template <typename K>
__global__ void dump_write_kernel(const K* __restrict__ keys, const u64* __restrict__ counts, const u64* __restrict__ off, u64 m, u32 w, u8* __restrict__ out) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        u8* p = out + off[i];
        const K k = keys[i];
        for (u32 j = 0; j < w; ++j) {                               // first base most significant
            const int sh = 2 * (int)(w - 1 - j);
            const u32 c = (u32)(KeyOps<K>::shr64(k, sh) & 3u);
            p[j] = "ACGT"[c];
        }
        p[w] = '\t';
        u64 v = (u32)counts[i];
        const u32 nd = dec_digits(v);
        for (u32 j = 0; j < nd; ++j) { p[w + nd - j] = (u8)('0' + v % 10); v /= 10; }
        p[w + 1 + nd] = '\n';
    }
}

// ---- host side: files through the gsb_source callbacks ------------------------------------------------------------------
struct FileOnDevice {
    DevBuf<u8> data;
    u64 size = 0;
    bool present = false;
};

void fetch_file(Workspace& ws, const gsb_source* src, const std::string& name, u8* pinned, size_t pinned_bytes, FileOnDevice& f, bool required) {
    uint64_t size = 0;
    f.present = src->size(src->user, name.c_str(), &size) == 0;
    if (!f.present) {
        if (required) throw StatusError{GSB_EIO, "cannot open " + name};
        f.size = 0; f.data.reset(&ws, 16);
        return;
    }
    f.size = size;
    f.data.reset(&ws, size + 16);
    for (u64 off = 0; off < size; off += pinned_bytes) {
        const u64 chunk = std::min<u64>(pinned_bytes, size - off);
        if (src->pread(src->user, name.c_str(), off, pinned, chunk) != 0) throw StatusError{GSB_EIO, "read failed for " + name};
        GSB_CUDA_TRY(cudaMemcpyAsync(f.data.p + off, pinned, chunk, cudaMemcpyHostToDevice, ws.stream));
        ws.sync();                                                  // the staging buffer is reused
    }
}

void host_file(const gsb_source* src, const std::string& name, std::vector<u8>& out) {
    uint64_t size = 0;
    if (src->size(src->user, name.c_str(), &size) != 0) throw StatusError{GSB_EIO, "cannot open " + name};
    out.resize(size);
    if (size && src->pread(src->user, name.c_str(), 0, out.data(), size) != 0) throw StatusError{GSB_EIO, "read failed for " + name};
}

struct PlaneSpec { const char* suffix; int shift; int bytes; };
// IntegerArray::builder's nesting of StackedArrays, src/IntegerArray.cc:259-357 (same table as the writer in emit.cu)
std::vector<PlaneSpec> planes_for(u64 bits) {
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
        default: throw StatusError{GSB_EINVAL, "IntegerArray: unsupported integer width " + std::to_string(bits)};
    }
}

struct SaHeader { u64 version, D, quantizedD, dmask[2], size[2], count; };

// SparseArray `base` -> m sorted keys of type K on the device
template <typename K>
void read_sparse_array_t(Workspace& ws, const gsb_source* src, const std::string& base, u8* pinned, size_t pinned_bytes, DevBuf<u8>& keys_out, u64* m_out) {
    cudaStream_t s = ws.stream;
    std::vector<u8> hb;
    host_file(src, base + ".header", hb);
    if (hb.size() < sizeof(SaHeader)) throw StatusError{GSB_EIO, base + ".header is truncated"};
    SaHeader h;
    memcpy(&h, hb.data(), sizeof(h));
    if (h.version != 2012030501ull) throw StatusError{GSB_EINVAL, base + ": SparseArray version mismatch " + std::to_string(h.version) + " vs 2012030501"};
    if (h.D > 128 || h.quantizedD != 8 * ((h.D + 7) / 8)) throw StatusError{GSB_EINVAL, base + ".header is inconsistent"};
    const u64 m = h.count;
    *m_out = m;
    keys_out.reset(&ws, m * sizeof(K));
    if (m == 0) return;
    if (sizeof(K) == 8 && h.D > 64) throw StatusError{GSB_EINVAL, base + ": positions do not fit 64-bit keys"};
    FileOnDevice bitmap;
    fetch_file(ws, src, base + ".high-bits", pinned, pinned_bytes, bitmap, true);
    const u64 n_words = bitmap.size / 8;
    std::vector<FileOnDevice> pf;
    Planes pl;
    memset(&pl, 0, sizeof(pl));
    const std::vector<PlaneSpec> specs = planes_for(h.quantizedD);
    pf.resize(specs.size());
    pl.n = (int)specs.size();
    for (size_t q = 0; q < specs.size(); ++q) {
        fetch_file(ws, src, base + ".low-bits" + specs[q].suffix, pinned, pinned_bytes, pf[q], true);
        if (pf[q].size < m * (u64)specs[q].bytes) throw StatusError{GSB_EIO, base + ".low-bits" + specs[q].suffix + " is truncated"};
        pl.p[q] = pf[q].data.p; pl.shift[q] = specs[q].shift; pl.bytes[q] = specs[q].bytes;
    }
    DevBuf<u32> pc(&ws, n_words);
    DevBuf<u64> rank_before(&ws, n_words + 1), tmp(&ws, scan_tmp_elems(n_words));
    const int g = (int)std::max<u64>(1, std::min<u64>((n_words + 255) / 256, (u64)ws.sm_count * 16));
    popc_words_kernel<<<g, 256, 0, s>>>((const u64*)bitmap.data.p, n_words, pc.p);
    ++ws.launches;
    exclusive_scan<u32, u64>(pc.p, rank_before.p, n_words, 0ull, rank_before.p + n_words, tmp.p, s, &ws.launches);
    u64 ones = 0;
    GSB_CUDA_TRY(cudaMemcpyAsync(&ones, rank_before.p + n_words, 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    if (ones < m) throw StatusError{GSB_EIO, base + ".high-bits holds fewer one bits than the header's count"};
    ef_decode_kernel<K><<<g, 256, 0, s>>>((const u64*)bitmap.data.p, n_words, rank_before.p, pl, (int)h.D, m, (K*)keys_out.p);
    ++ws.launches;
    ws.sync();                                                      // the file buffers go out of scope
}

}  // namespace

void read_sparse_array(Workspace& ws, const gsb_source* src, const std::string& base, int key_bytes, u8* pinned, size_t pinned_bytes,
                       DevBuf<u8>& keys_out, u64* m_out) {
    if (key_bytes == 8) read_sparse_array_t<u64>(ws, src, base, pinned, pinned_bytes, keys_out, m_out);
    else read_sparse_array_t<Key128>(ws, src, base, pinned, pinned_bytes, keys_out, m_out);
}

// VariableByteArray `base` with m items -> u64 counts (the stored 32-bit multiplicities)
void read_counts(Workspace& ws, const gsb_source* src, const std::string& base, u64 m, u8* pinned, size_t pinned_bytes, DevBuf<u64>& counts_out) {
    cudaStream_t s = ws.stream;
    counts_out.reset(&ws, m);
    if (!m) return;
    FileOnDevice ord0, ord1, ord2;
    fetch_file(ws, src, base + ".ord0", pinned, pinned_bytes, ord0, true);
    if (ord0.size < m) throw StatusError{GSB_EIO, base + ".ord0 is truncated"};
    const int g = (int)std::max<u64>(1, std::min<u64>((m + 255) / 256, (u64)ws.sm_count * 16));
    vba_base_kernel<<<g, 256, 0, s>>>(ord0.data.p, m, counts_out.p);
    ++ws.launches;
    DevBuf<u8> pos1, pos2;
    u64 n1 = 0, n2 = 0;
    read_sparse_array(ws, src, base + ".ord1p", 8, pinned, pinned_bytes, pos1, &n1);
    if (n1) {
        fetch_file(ws, src, base + ".ord1", pinned, pinned_bytes, ord1, true);
        if (ord1.size < n1) throw StatusError{GSB_EIO, base + ".ord1 is truncated"};
        vba_ord1_kernel<<<(unsigned)std::min<u64>((n1 + 255) / 256, 4096), 256, 0, s>>>((const u64*)pos1.p, ord1.data.p, n1, m, counts_out.p);
        ++ws.launches;
        read_sparse_array(ws, src, base + ".ord2p", 8, pinned, pinned_bytes, pos2, &n2);
        if (n2) {
            fetch_file(ws, src, base + ".ord2", pinned, pinned_bytes, ord2, true);
            if (ord2.size < 2 * n2) throw StatusError{GSB_EIO, base + ".ord2 is truncated"};
            vba_ord2_kernel<<<(unsigned)std::min<u64>((n2 + 255) / 256, 4096), 256, 0, s>>>((const u64*)pos1.p, (const u64*)pos2.p, (const u16*)ord2.data.p, n2, n1, m,
                                                                                                counts_out.p);
            ++ws.launches;
        }
    }
    ws.sync();
}

void fill_value(Workspace& ws, u64* counts, u64 m, u64 v) {
    if (!m) return;
    fill_value_kernel<<<(unsigned)std::min<u64>((m + 255) / 256, (u64)ws.sm_count * 16), 256, 0, ws.stream>>>(counts, m, v);
    ++ws.launches;
}
void fill_ones(Workspace& ws, u64* counts, u64 m) { fill_value(ws, counts, m, 1); }

// dump-graph's body lines for the run, as one device buffer of text
void dump_text(Workspace& ws, int key_bytes, const void* keys, const u64* counts, u64 m, int w, DevBuf<u8>& text_out, u64* bytes_out) {
    cudaStream_t s = ws.stream;
    *bytes_out = 0;
    text_out.reset(&ws, 16);
    if (!m) return;
    DevBuf<u32> len(&ws, m);
    DevBuf<u64> off(&ws, m + 1), tmp(&ws, scan_tmp_elems(m));
    const int g = (int)std::max<u64>(1, std::min<u64>((m + 255) / 256, (u64)ws.sm_count * 16));
    dump_len_kernel<<<g, 256, 0, s>>>(counts, m, (u32)w, len.p);
    ++ws.launches;
    exclusive_scan<u32, u64>(len.p, off.p, m, 0ull, off.p + m, tmp.p, s, &ws.launches);
    u64 total = 0;
    GSB_CUDA_TRY(cudaMemcpyAsync(&total, off.p + m, 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    text_out.reset(&ws, total + 16);
    if (key_bytes == 8) dump_write_kernel<u64><<<g, 256, 0, s>>>((const u64*)keys, counts, off.p, m, (u32)w, text_out.p);
    else dump_write_kernel<Key128><<<g, 256, 0, s>>>((const Key128*)keys, counts, off.p, m, (u32)w, text_out.p);
    ++ws.launches;
    *bytes_out = total;
}

}  // namespace gsb
