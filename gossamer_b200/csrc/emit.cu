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

#include "exchange.h"
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
    if (ring) { ring_put_host(name, len, 0, data, len); return; }
    void* h = nullptr;
    if (sink->open(sink->user, name.c_str(), len, &h) != 0) sink_fail("open", name);
    if (len && sink->pwrite(sink->user, h, 0, data, len) != 0) sink_fail("pwrite", name);
    if (sink->close(sink->user, h) != 0) sink_fail("close", name);
}

// ---- the ring and its writer thread ----------------------------------------------------------------------------------
void EmitRing::create(Workspace& ws) {
    if (created) return;
    device = ws.device;
    GSB_CUDA_TRY(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
    for (int i = 0; i < kSlots; ++i) {
        dev[i] = (u8*)ws.alloc(kChunk);
        GSB_CUDA_TRY(cudaMallocHost((void**)&host[i], kChunk));
        GSB_CUDA_TRY(cudaEventCreateWithFlags(&ready[i], cudaEventDisableTiming));
        GSB_CUDA_TRY(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        slot_busy[i] = false;
    }
    stop = false;
    worker = std::thread([this] { run(); });
    created = true;
}

void EmitRing::wait_idle() {
    std::unique_lock<std::mutex> lk(mu);
    cv_idle.wait(lk, [this] { return queue.empty(); });
}

void EmitRing::drop() {
    if (!created) return;
    wait_idle();
    std::lock_guard<std::mutex> lk(mu);
    error.clear();
    handles.clear();
}

void EmitRing::destroy(Workspace& ws) {
    if (!created) return;
    wait_idle();
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    cv_work.notify_all();
    if (worker.joinable()) worker.join();
    for (int i = 0; i < kSlots; ++i) {
        if (dev[i]) ws.release(dev[i], kChunk);
        if (host[i]) cudaFreeHost(host[i]);
        cudaEventDestroy(ready[i]); cudaEventDestroy(done[i]);
        dev[i] = nullptr; host[i] = nullptr;
    }
    cudaStreamDestroy(copy);
    created = false;
}

int EmitRing::acquire_slot() {
    std::unique_lock<std::mutex> lk(mu);
    int slot = -1;
    cv_idle.wait(lk, [&] {
        for (int i = 0; i < kSlots; ++i) if (!slot_busy[i]) { slot = i; return true; }
        return false;
    });
    slot_busy[slot] = true;
    return slot;
}

void EmitRing::push(Job&& j) {
    { std::lock_guard<std::mutex> lk(mu); queue.push_back(std::move(j)); }
    cv_work.notify_one();
}

// Writer thread.  Queues the device -> host copy of every job that has none yet (so PCIe stays busy while a callback
// runs), then completes the oldest job: waits for its copy, makes the sink calls, frees its slot.
void EmitRing::run() {
    cudaSetDevice(device);
    for (;;) {
        Job* job = nullptr;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_work.wait(lk, [this] { return stop || !queue.empty(); });
            if (queue.empty()) return;                            // stop requested and nothing left
            for (Job& j : queue) {
                if (j.slot >= 0 && !j.issued) {
                    if (j.len) {
                        cudaStreamWaitEvent(copy, ready[j.slot], 0);
                        cudaMemcpyAsync(host[j.slot], dev[j.slot], j.len, cudaMemcpyDeviceToHost, copy);
                    }
                    cudaEventRecord(done[j.slot], copy);
                    j.issued = true;
                }
            }
            job = &queue.front();                                 // deque: stays valid while the producer appends
        }
        std::string fail;
        bool skip;
        { std::lock_guard<std::mutex> lk(mu); skip = !error.empty(); }
        if (job->slot >= 0 && cudaEventSynchronize(done[job->slot]) != cudaSuccess) { fail = "device to host copy failed for " + job->name; cudaGetLastError(); }
        if (!skip && fail.empty()) {
            const u8* data = job->slot >= 0 ? host[job->slot] : job->bytes.data();
            if (job->slot >= 0 && !job->bytes.empty()) memcpy(host[job->slot], job->bytes.data(), job->bytes.size());   // header bytes over the payload
            void* h = nullptr;
            if (job->open) {
                if (sink->open(sink->user, job->name.c_str(), job->size_hint, &h) != 0) fail = "open failed for " + job->name;
                else handles[job->name] = h;
            } else {
                h = handles[job->name];
            }
            if (fail.empty() && job->len && sink->pwrite(sink->user, h, job->file_off, data, job->len) != 0) fail = "pwrite failed for " + job->name;
            if (fail.empty() && job->close) {
                if (sink->close(sink->user, h) != 0) fail = "close failed for " + job->name;
                handles.erase(job->name);
            }
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            if (!fail.empty() && error.empty()) error = fail;
            if (job->slot >= 0) slot_busy[job->slot] = false;
            queue.pop_front();
        }
        cv_idle.notify_all();
    }
}

void Emitter::flush() {
    if (!ring || !sink) return;
    ring->wait_idle();
    std::string err;
    { std::lock_guard<std::mutex> lk(ring->mu); err = ring->error; }
    if (!err.empty()) throw StatusError{GSB_EIO, err};
}

// [file_off, file_off + len) of file `name` from device memory, through the ring: opened before the first chunk, closed after
// the last one
void Emitter::ring_put(const std::string& name, u64 size_hint, u64 file_off, const void* dev, u64 len, const void* host_prefix, u64 prefix_len) {
    for (u64 off = 0; off < len || off == 0; off += EmitRing::kChunk) {
        const u64 chunk = std::min<u64>(EmitRing::kChunk, len - off);
        EmitRing::Job j;
        j.name = name; j.size_hint = size_hint; j.file_off = file_off + off; j.len = chunk;
        j.open = off == 0; j.close = off + chunk >= len;
        j.slot = ring->acquire_slot();
        if (chunk) {
            cudaError_t e = cudaMemcpyAsync(ring->dev[j.slot], (const u8*)dev + off, chunk, cudaMemcpyDeviceToDevice, ws->stream);
            if (e == cudaSuccess) e = cudaEventRecord(ring->ready[j.slot], ws->stream);
            if (e != cudaSuccess) {                                 // the slot must not stay taken
                { std::lock_guard<std::mutex> lk(ring->mu); ring->slot_busy[j.slot] = false; }
                ring->cv_idle.notify_all();
                GSB_CUDA_TRY(e);
            }
        }
        if (off < prefix_len) j.bytes.assign((const u8*)host_prefix + off, (const u8*)host_prefix + std::min<u64>(prefix_len, off + chunk));
        ring->push(std::move(j));
        if (len == 0) break;
    }
}

void Emitter::ring_put_host(const std::string& name, u64 size_hint, u64 file_off, const void* data, u64 len) {
    EmitRing::Job j;
    j.name = name; j.size_hint = size_hint; j.file_off = file_off; j.len = len;
    j.open = true; j.close = true;
    j.bytes.assign((const u8*)data, (const u8*)data + len);
    ring->push(std::move(j));
}

void Emitter::put_device(const std::string& name, const void* dev, u64 len, const void* host_prefix, u64 prefix_len) {
    bytes_out += len;
    if (!sink) return;
    if (ring) { ring_put(name, len, 0, dev, len, host_prefix, prefix_len); return; }
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

// one contiguous piece [offset, offset + len) of a file whose total size is `total` (multi-GPU emission:
// every rank hands over its own pieces; bytes nobody writes are zero)
void Emitter::put_device_at(const std::string& name, u64 total, u64 offset, const void* dev, u64 len) {
    bytes_out += len;
    if (!sink || !len) return;
    if (ring) { ring_put(name, total, offset, dev, len, nullptr, 0); return; }
    void* h = nullptr;
    if (sink->open(sink->user, name.c_str(), total, &h) != 0) sink_fail("open", name);
    for (u64 off = 0; off < len; off += pinned_bytes) {
        u64 chunk = std::min<u64>(pinned_bytes, len - off);
        GSB_CUDA_TRY(cudaMemcpyAsync(pinned, (const u8*)dev + off, chunk, cudaMemcpyDeviceToHost, ws->stream));
        ws->sync();
        if (sink->pwrite(sink->user, h, offset + off, pinned, chunk) != 0) sink_fail("pwrite", name);
    }
    if (sink->close(sink->user, h) != 0) sink_fail("close", name);
}

void Emitter::put_host_at(const std::string& name, u64 total, u64 offset, const void* data, u64 len) {
    bytes_out += len;
    if (!sink || !len) return;
    if (ring) { ring_put_host(name, total, offset, data, len); return; }
    void* h = nullptr;
    if (sink->open(sink->user, name.c_str(), total, &h) != 0) sink_fail("open", name);
    if (sink->pwrite(sink->user, h, offset, data, len) != 0) sink_fail("pwrite", name);
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

// blocks [b0, b0 + n_blocks) of the directory (b0 > 0: this rank's share in a multi-GPU emission);
// outputs are indexed by b - b0
template <typename F>
__global__ void ds_classify_kernel(F f, u64 count, u64 b0, u64 n_blocks, u8* __restrict__ type, u64* __restrict__ bytes,
                                   u64* __restrict__ padded, u64* __restrict__ first_out) {
    for (u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += (u64)gridDim.x * blockDim.x) {
        const u64 j0 = (b0 + b) * kBlock;
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
// `body` holds this launch's blocks; off[] are FILE offsets and body_file_off is the file offset of body[0]
template <typename F>
__global__ void __launch_bounds__(128) ds_write_kernel(F f, u64 count, u64 b0, const u8* __restrict__ type, const u64* __restrict__ off,
                                                       const u64* __restrict__ first_in, u8* __restrict__ body, u64 body_file_off,
                                                       u64* __restrict__ index_out, u64* __restrict__ rank_out) {
    __shared__ u32 scan_s[128 / 32 + 1];
    const u64 b = blockIdx.x;
    const u64 j0 = (b0 + b) * kBlock;
    const u64 nb = count - j0 < kBlock ? count - j0 : kBlock;
    const u64 first = first_in[b];
    const u8 t = type[b];
    u8* out = body + (off[b] - body_file_off);
    if (threadIdx.x == 0) {
        index_out[b] = off[b] | t;
        rank_out[b] = first;
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
    ds_classify_kernel<F><<<g, 128, 0, s>>>(f, count, 0, nb, type.p, bytes.p, padded.p, first.p);
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
    ds_write_kernel<F><<<(unsigned)nb, 128, 0, s>>>(f, count, 0, type.p, off.p, first.p, file.p, 0, (u64*)(file.p + h.indexArrayOffset), (u64*)(file.p + h.rankArrayOffset));
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

// i_bias / a_bias: global index of this rank's first item / first ord1 entry (0 on one GPU)
__global__ void vba_scatter_kernel(const u64* __restrict__ counts, u64 m, const u64* __restrict__ r1, const u64* __restrict__ r2,
                                   u64* __restrict__ ord1_pos, u8* __restrict__ ord1, u64* __restrict__ ord2_pos, u16* __restrict__ ord2,
                                   u64 i_bias, u64 a_bias) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
        const u32 c = (u32)counts[i];
        if (c >> 8) {
            const u64 a = r1[i];
            ord1_pos[a] = i + i_bias; ord1[a] = (u8)((c >> 8) & 0xFF);
            if (c >> 16) { const u64 b = r2[i]; ord2_pos[b] = a + a_bias; ord2[b] = (u16)(c >> 16); }
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
    if (n1) { vba_scatter_kernel<<<g, 256, 0, s>>>(counts, m, r1.p, r2.p, ord1_pos.p, ord1.p, ord2_pos.p, ord2.p, 0, 0); ++ws.launches; }
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

// ------------------------------------------------------------------------------------------
// Multi-GPU emission (one process per GPU)
// ------------------------------------------------------------------------------------------
// After counting, rank r holds the r-th contiguous slice of the globally sorted run in its peer-mapped
// window (DistRun, exchange.h).  Every byte of the reference's files is a function of the global index of
// an element, so each rank writes the byte ranges that belong to its slice:
//   .low-bits planes, .ord0        element i -> byte i * width: the rank's own index range
//   .high-bits                     the words from the one holding the rank's first one-bit up to (not
//                                  including) the next rank's; bits that earlier ranks' last elements put
//                                  into the rank's first word are recomputed from those elements
//   -d1 / -d0                      the select blocks that START in the rank's slice (ones) or in the range
//                                  of high parts it covers (zeros); block sizes -> one all-gather -> file offsets
// Elements of neighbouring slices (the tail of a block, the word shared with the previous rank, binary
// searches that leave the slice) are loaded straight from the owner's memory over NVLink.
// The wide-count planes (.ord1/.ord2 and their presence sets: counts >= 256) and the count histogram are
// small; they are gathered and written by rank 0.
template <typename K>
struct GlobalKeys {
    const K* base[kMaxRanks];
    u64 off[kMaxRanks + 1];
    const K* mine; u64 g0, g1;                                  // this rank's slice (fast path)
    int n;
    __device__ __forceinline__ K operator[](u64 i) const {
        if (i >= g0 && i < g1) return mine[i - g0];
        int r = 0;
        while (r + 1 < n && i >= off[r + 1]) ++r;
        return base[r][i - off[r]];
    }
};

template <typename K> struct OnesPosG {
    GlobalKeys<K> G; int D;
    __device__ __forceinline__ u64 operator()(u64 j) const { return KeyOps<K>::shr64(G[j], D) + j; }
};
template <typename K> struct ZerosPosG {
    GlobalKeys<K> G; u64 m; int D;
    u64 jlo, jhi;                                               // high parts [jlo, jhi) resolve inside this rank's slice
    __device__ __forceinline__ u64 operator()(u64 j) const {
        u64 lo = 0, hi = m;                                     // #{i : (e_i >> D) <= j}
        if (j >= jlo && j < jhi) { lo = G.g0; hi = G.g1; }
        while (lo < hi) { u64 mid = lo + ((hi - lo) >> 1); if (KeyOps<K>::shr64(G[mid], D) <= j) lo = mid + 1; else hi = mid; }
        return j + lo;
    }
};

// out[2r] = high part of rank r's first element, out[2r+1] = position of its one-bit (non-empty ranks only)
template <typename K>
__global__ void first_elements_kernel(GlobalKeys<K> G, int D, u64* __restrict__ out) {
    const int r = threadIdx.x;
    if (r < G.n && G.off[r] < G.off[r + 1]) {
        const u64 hi = KeyOps<K>::shr64(G.base[r][0], D);
        out[2 * r] = hi; out[2 * r + 1] = hi + G.off[r];
    }
}

template <typename K>
__global__ void high_bits_dist_kernel(GlobalKeys<K> G, u64 i0, u64 i1, int D, u64 w0, u64 w1, u64* __restrict__ bitmap /* words [w0, w1) */) {
    for (u64 i = i0 + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (u64)gridDim.x * blockDim.x) {
        const u64 h = KeyOps<K>::shr64(G[i], D) + i;
        const u64 w = h >> 6;
        if (w >= w0 && w < w1) atomicOr(&bitmap[w - w0], 1ull << (h & 63));
    }
}

// Every rank classifies ALL blocks of the directory (a block's class and size need its first and last position, plus
// the 128 sample spans for the rare "intermediate" class; positions outside the rank's slice are read from the owner
// over NVLink), so each rank knows every file offset and the header statistics without a collective; it then writes
// only the blocks [b0, b1) it owns.
template <typename F>
static void build_dense_select_dist(Emitter& em, Exchange* x, F f, u64 count, bool invert, const std::string& name, u64 b0, u64 b1) {
    Workspace& ws = *em.ws;
    cudaStream_t s = ws.stream;
    const int rank = exchange_rank(x);
    DsHeader h;
    memset(&h, 0, sizeof(h));
    h.version = 2012092701ull; h.flags = invert ? 1 : 0;
    h.logBlockSize = 13; h.blockSize = 8192; h.logSampleRate = 6; h.sampleRate = 64;
    const u64 nb = (count + kBlock - 1) / kBlock;
    std::vector<u8> page(4096, 0);
    if (nb == 0) {
        h.indexArrayOffset = h.rankArrayOffset = 4096;
        memcpy(page.data(), &h, sizeof(h));
        if (rank == 0) em.put_host(name, page.data(), page.size());
        return;
    }
    DevBuf<u8> type(&ws, nb);
    DevBuf<u64> bytes(&ws, nb), padded(&ws, nb), first(&ws, nb), off(&ws, nb + 1), tmp(&ws, scan_tmp_elems(nb)), scalars(&ws, 8);
    GSB_CUDA_TRY(cudaMemsetAsync(scalars.p, 0, 64, s));
    const int g = (int)std::min<u64>((nb + 127) / 128, (u64)ws.sm_count * 8);
    ds_classify_kernel<F><<<g, 128, 0, s>>>(f, count, 0, nb, type.p, bytes.p, padded.p, first.p);
    ++ws.launches;
    exclusive_scan<u64, u64>(padded.p, off.p, nb, 4096ull, scalars.p + 6, tmp.p, s, &ws.launches);
    ds_stats_kernel<<<g, 128, 0, s>>>(type.p, bytes.p, nb, scalars.p);
    ++ws.launches;
    u64 host[8], seg[2] = {0, 0};
    GSB_CUDA_TRY(cudaMemcpyAsync(host, scalars.p, 64, cudaMemcpyDeviceToHost, s));
    if (b0 < nb) GSB_CUDA_TRY(cudaMemcpyAsync(&seg[0], off.p + b0, 8, cudaMemcpyDeviceToHost, s));
    if (b1 < nb) GSB_CUDA_TRY(cudaMemcpyAsync(&seg[1], off.p + b1, 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    const u64 body_end = 4096 + host[6];
    const u64 seg_begin = b0 < nb ? seg[0] : body_end, seg_end = b1 < nb ? seg[1] : body_end;
    h.indexArrayOffset = (body_end + 15) & ~15ull;
    h.rankArrayOffset = h.indexArrayOffset + 8 * nb;
    const u64 file_size = h.rankArrayOffset + 8 * nb;
    h.numBlocks = nb; h.indexSize = 16 * nb;
    h.smallBlocks = host[0]; h.smallBlocksSize = host[1];
    h.intermediateBlocks = host[2]; h.intermediateBlocksSize = host[3];
    h.largeBlocks = host[4]; h.largeBlocksSize = host[5];
    const u64 nbl = b1 - b0;
    if (nbl) {
        const u64 body_bytes = seg_end - seg_begin;
        DevBuf<u8> body(&ws, body_bytes);
        DevBuf<u64> index(&ws, nbl), ranks(&ws, nbl);
        GSB_CUDA_TRY(cudaMemsetAsync(body.p, 0, body_bytes ? body_bytes : 1, s));
        ds_write_kernel<F><<<(unsigned)nbl, 128, 0, s>>>(f, count, b0, type.p + b0, off.p + b0, first.p + b0, body.p, seg_begin, index.p, ranks.p);
        ++ws.launches;
        em.put_device_at(name, file_size, seg_begin, body.p, body_bytes);
        em.put_device_at(name, file_size, h.indexArrayOffset + 8 * b0, index.p, 8 * nbl);
        em.put_device_at(name, file_size, h.rankArrayOffset + 8 * b0, ranks.p, 8 * nbl);
    }
    if (rank == 0) {
        memcpy(page.data(), &h, sizeof(h));
        em.put_host_at(name, file_size, 0, page.data(), page.size());
        const u8 zeros[16] = {0};
        if (h.indexArrayOffset > body_end) em.put_host_at(name, file_size, body_end, zeros, h.indexArrayOffset - body_end);   // alignment gap
    }
}

template <typename K>
static void emit_sparse_array_dist_t(Emitter& em, Exchange* x, const DistRun& run, U128 universe_ctor, u64 m_est, U128 universe_end, const std::string& base) {
    Workspace& ws = *em.ws;
    cudaStream_t s = ws.stream;
    const int n = run.n, rank = run.rank;
    const u64 m = run.off[n], g0 = run.off[rank], g1 = run.off[rank + 1], ml = g1 - g0;
    if (m == 0) {                                               // nothing anywhere: rank 0 writes the empty structure
        if (rank == 0) emit_sparse_array_t<K>(em, (const K*)nullptr, 0, universe_ctor, m_est, universe_end, base);
        return;
    }
    const u64 D = sparse_array_d(universe_ctor, m_est);
    const u64 qD = 8 * ((D + 7) / 8);
    const u128_t nd128 = D >= 128 ? (u128_t)0 : (to128(universe_end) >> D);
    if ((u64)(nd128 >> 64)) throw StatusError{GSB_EINVAL, "Internal error in SparseArray; nd does not fit 64 bits"};
    const u64 nd = (u64)nd128;

    GlobalKeys<K> G;
    for (int r = 0; r < kMaxRanks; ++r) { G.base[r] = (const K*)run.keys[r]; G.off[r] = run.off[r]; }
    G.off[kMaxRanks] = run.off[kMaxRanks];
    G.mine = (const K*)run.keys[rank]; G.g0 = g0; G.g1 = g1; G.n = n;

    // high part / one-bit position of every non-empty rank's first element
    DevBuf<u64> firsts_d(&ws, 2 * (size_t)kMaxRanks);
    std::vector<u64> firsts(2 * (size_t)kMaxRanks, 0);
    first_elements_kernel<K><<<1, 32, 0, s>>>(G, (int)D, firsts_d.p);
    ++ws.launches;
    GSB_CUDA_TRY(cudaMemcpyAsync(firsts.data(), firsts_d.p, firsts.size() * 8, cudaMemcpyDeviceToHost, s));
    ws.sync();
    auto nonempty = [&](int r) { return run.off[r] < run.off[r + 1]; };
    int first_nonempty = 0;
    while (!nonempty(first_nonempty)) ++first_nonempty;
    int next_nonempty = rank + 1;
    while (next_nonempty < n && !nonempty(next_nonempty)) ++next_nonempty;

    // ---- high bits ----
    const u64 words = (nd + m + 3) / 64 + 1;
    std::vector<u64> W(n + 1);
    W[n] = words;
    for (int r = n - 1; r >= 0; --r) W[r] = nonempty(r) ? (firsts[2 * r + 1] >> 6) : W[r + 1];
    for (int r = 0; r <= first_nonempty; ++r) W[r] = 0;
    {
        const u64 wl = W[rank + 1] - W[rank];
        if (wl) {
            DevBuf<u64> bitmap(&ws, wl);
            GSB_CUDA_TRY(cudaMemsetAsync(bitmap.p, 0, wl * 8, s));
            const u64 i0 = g0 >= 64 ? g0 - 64 : 0;              // one-bit positions grow by >= 1 per element: older ones are in earlier words
            if (g1 > i0) {
                const int g = (int)std::max<u64>(1, std::min<u64>((g1 - i0 + 255) / 256, (u64)ws.sm_count * 16));
                high_bits_dist_kernel<K><<<g, 256, 0, s>>>(G, i0, g1, (int)D, W[rank], W[rank + 1], bitmap.p);
                ++ws.launches;
            }
            em.put_device_at(base + ".high-bits", words * 8, W[rank] * 8, bitmap.p, wl * 8);
        }
    }
    // ---- select directory over the zeros: rank r takes the blocks that start inside the high parts it covers ----
    {
        const u64 count0 = nd + 2, nbz = (count0 + kBlock - 1) / kBlock;
        std::vector<u64> B(n + 1);
        B[n] = nbz;
        for (int r = n - 1; r >= 0; --r) B[r] = nonempty(r) ? std::min<u64>(nbz, (firsts[2 * r] + kBlock - 1) / kBlock) : B[r + 1];
        for (int r = 0; r <= first_nonempty; ++r) B[r] = 0;
        ZerosPosG<K> f{G, m, (int)D, 0, 0};
        if (ml) { f.jlo = firsts[2 * rank]; f.jhi = next_nonempty < n ? firsts[2 * next_nonempty] : ~0ull; }
        build_dense_select_dist(em, x, f, count0, true, base + "-d0", B[rank], B[rank + 1]);
    }
    // ---- select directory over the ones: the blocks whose first element is in the rank's slice ----
    build_dense_select_dist(em, x, OnesPosG<K>{G, (int)D}, m, false, base + "-d1", (g0 + kBlock - 1) / kBlock, (g1 + kBlock - 1) / kBlock);
    // ---- low bits: the rank's own elements ----
    const int g = (int)std::max<u64>(1, std::min<u64>((ml + 255) / 256, (u64)ws.sm_count * 16));
    for (const PlaneSpec& p : integer_array_planes(qD)) {
        if (!ml) continue;
        DevBuf<u8> plane(&ws, ml * p.bytes);
        switch (p.bytes) {
            case 1: low_plane_kernel<K, u8><<<g, 256, 0, s>>>(G.mine, ml, (int)D, p.shift, (u8*)plane.p); break;
            case 2: low_plane_kernel<K, u16><<<g, 256, 0, s>>>(G.mine, ml, (int)D, p.shift, (u16*)plane.p); break;
            case 4: low_plane_kernel<K, u32><<<g, 256, 0, s>>>(G.mine, ml, (int)D, p.shift, (u32*)plane.p); break;
            default: low_plane_kernel<K, u64><<<g, 256, 0, s>>>(G.mine, ml, (int)D, p.shift, (u64*)plane.p); break;
        }
        ++ws.launches;
        em.put_device_at(base + ".low-bits" + p.suffix, m * p.bytes, g0 * p.bytes, plane.p, ml * p.bytes);
    }
    if (rank == 0) {
        struct { u64 version, D, quantizedD, dmask[2], size[2], count; } hd;
        hd.version = 2012030501ull; hd.D = D; hd.quantizedD = qD;
        u128_t mask = D >= 128 ? ~(u128_t)0 : ((((u128_t)1) << D) - 1);
        hd.dmask[0] = (u64)mask; hd.dmask[1] = (u64)(mask >> 64);
        hd.size[0] = universe_end.lo; hd.size[1] = universe_end.hi;
        hd.count = m;
        em.put_host(base + ".header", &hd, sizeof(hd));
    }
}

void emit_sparse_array_dist(Emitter& em, Exchange* x, int key_bytes, const DistRun& run, U128 universe_ctor, u64 m_est, U128 universe_end,
                            const std::string& base) {
    if (key_bytes == 8) emit_sparse_array_dist_t<u64>(em, x, run, universe_ctor, m_est, universe_end, base);
    else emit_sparse_array_dist_t<Key128>(em, x, run, universe_ctor, m_est, universe_end, base);
}

void emit_counts_dist(Emitter& em, Exchange* x, const DistRun& run, u64 m_est, const std::string& base) {
    Workspace& ws = *em.ws;
    cudaStream_t s = ws.stream;
    const int n = run.n, rank = run.rank;
    const u64 m = run.off[n], g0 = run.off[rank], ml = run.off[rank + 1] - g0;
    const u64* counts = run.counts[rank];
    const int g = (int)std::max<u64>(1, std::min<u64>((ml + 255) / 256, (u64)ws.sm_count * 16));
    DevBuf<u8> ord0(&ws, ml), f1(&ws, ml), f2(&ws, ml);
    DevBuf<u64> r1(&ws, ml), r2(&ws, ml), tmp(&ws, scan_tmp_elems(ml)), totals(&ws, 2);
    u64 mine[2] = {0, 0};
    if (ml) {
        vba_flags_kernel<<<g, 256, 0, s>>>(counts, ml, ord0.p, f1.p, f2.p);
        ++ws.launches;
        exclusive_scan<u8, u64>(f1.p, r1.p, ml, 0ull, totals.p, tmp.p, s, &ws.launches);
        exclusive_scan<u8, u64>(f2.p, r2.p, ml, 0ull, totals.p + 1, tmp.p, s, &ws.launches);
        GSB_CUDA_TRY(cudaMemcpyAsync(mine, totals.p, 16, cudaMemcpyDeviceToHost, s));
        ws.sync();
    }
    std::vector<u64> all;
    exchange_allgather_u64(x, ws, mine, 2, all);
    u64 n1 = 0, n2 = 0, a_bias = 0;
    std::vector<u64> b_pos1(n), b_ord1(n), b_pos2(n), b_ord2(n);
    for (int r = 0; r < n; ++r) {
        if (r < rank) a_bias += all[2 * r];
        n1 += all[2 * r]; n2 += all[2 * r + 1];
        b_pos1[r] = all[2 * r] * 8; b_ord1[r] = all[2 * r]; b_pos2[r] = all[2 * r + 1] * 8; b_ord2[r] = all[2 * r + 1] * 2;
    }
    em.put_device_at(base + ".ord0", m, g0, ord0.p, ml);
    // counts >= 256 are the exception (repeats): their planes and presence sets go to rank 0
    DevBuf<u8> pos1_all, ord1_all, pos2_all, ord2_all;
    if (n1) {
        DevBuf<u64> ord1_pos(&ws, mine[0]), ord2_pos(&ws, mine[1]);
        DevBuf<u8> ord1(&ws, mine[0]);
        DevBuf<u16> ord2(&ws, mine[1]);
        if (mine[0]) { vba_scatter_kernel<<<g, 256, 0, s>>>(counts, ml, r1.p, r2.p, ord1_pos.p, ord1.p, ord2_pos.p, ord2.p, g0, a_bias); ++ws.launches; }
        exchange_gatherv_root(x, ws, ord1_pos.p, b_pos1, pos1_all);
        exchange_gatherv_root(x, ws, ord1.p, b_ord1, ord1_all);
        exchange_gatherv_root(x, ws, ord2_pos.p, b_pos2, pos2_all);
        exchange_gatherv_root(x, ws, ord2.p, b_ord2, ord2_all);
    }
    if (rank == 0) {
        em.put_device(base + ".ord1", ord1_all.p, n1);
        em.put_device(base + ".ord2", ord2_all.p, n2 * 2);
        const U128 n_ctor{m_est, 0};
        const u64 m_frac = (u64)(m_est * 0.001);
        emit_sparse_array(em, 8, pos1_all.p, n1, n_ctor, m_frac, U128{m, 0}, base + ".ord1p");
        emit_sparse_array(em, 8, pos2_all.p, n2, n_ctor, m_frac, U128{n1, 0}, base + ".ord2p");
    }
}

void emit_count_histogram_dist(Emitter& em, Exchange* x, const DistRun& run, const std::string& name) {
    Workspace& ws = *em.ws;
    cudaStream_t s = ws.stream;
    const int n = run.n, rank = run.rank;
    const u64 ml = run.off[rank + 1] - run.off[rank];
    std::vector<u64> pairs;                                       // (count value, frequency) of this rank's slice
    if (ml) {
        DevBuf<u64> a(&ws, ml), b(&ws, ml);
        GSB_CUDA_TRY(cudaMemcpyAsync(a.p, run.counts[rank], ml * 8, cudaMemcpyDeviceToDevice, s));
        int passes = 0;
        int where = sort_keys(ws, 8, 64, a.p, b.p, nullptr, nullptr, ml, nullptr, &passes);
        ReducedRun rr; u64 distinct = 0;
        reduce_sorted(ws, 8, where ? b.p : a.p, nullptr, ml, 1, rr, &distinct);
        std::vector<u64> vals(rr.m), freq(rr.m);
        GSB_CUDA_TRY(cudaMemcpyAsync(vals.data(), rr.keys.p, rr.m * 8, cudaMemcpyDeviceToHost, s));
        GSB_CUDA_TRY(cudaMemcpyAsync(freq.data(), rr.counts.p, rr.m * 8, cudaMemcpyDeviceToHost, s));
        ws.sync();
        pairs.resize(2 * rr.m);
        for (u64 i = 0; i < rr.m; ++i) { pairs[2 * i] = vals[i]; pairs[2 * i + 1] = freq[i]; }
    }
    // one all-gather of [number of pairs, the first kInline pairs]; a second one only if some rank has more
    const u64 kInline = 510;
    const u64 mine_n = pairs.size() / 2;
    std::vector<u64> msg(1 + 2 * kInline, 0), got;
    msg[0] = mine_n;
    for (u64 i = 0; i < std::min(mine_n, kInline) * 2; ++i) msg[1 + i] = pairs[i];
    exchange_allgather_u64(x, ws, msg.data(), msg.size(), got);
    std::vector<u64> sizes(n);
    u64 cap = 1;
    for (int r = 0; r < n; ++r) { sizes[r] = got[(size_t)r * msg.size()]; cap = std::max(cap, sizes[r]); }
    std::vector<u64> all;
    if (cap > kInline) {
        pairs.resize(2 * cap, 0);
        exchange_allgather_u64(x, ws, pairs.data(), 2 * cap, all);
    } else {
        cap = kInline;
        all.resize((size_t)n * 2 * cap);
        for (int r = 0; r < n; ++r)
            for (u64 i = 0; i < 2 * cap; ++i) all[(size_t)r * 2 * cap + i] = got[(size_t)r * msg.size() + 1 + i];
    }
    if (rank == 0) {
        std::map<u64, u64> hist;
        for (int r = 0; r < n; ++r)
            for (u64 i = 0; i < sizes[r]; ++i) hist[all[2 * cap * r + 2 * i]] += all[2 * cap * r + 2 * i + 1];
        std::string text;
        for (const auto& kv : hist) text += std::to_string(kv.first) + "\t" + std::to_string(kv.second) + "\n";
        em.put_host(name, text.data(), text.size());
    }
}

}  // namespace gsb
