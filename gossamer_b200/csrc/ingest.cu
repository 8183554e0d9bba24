// ingest.cu -- raw FASTA / FASTQ / line text -> packed symbol stream -> window keys.
//
//   K1  newline scan + line table                    (PlainLineSource, src/LineSource.cc:17-48)
//   K1b record framing per line                      (FastaParser::next src/FastaParser.hh:51-87,
//                                                     FastqParser::next src/FastqParser.hh:78-176,
//                                                     LineParser::next  src/LineParser.hh:71-82)
//   K2  2-bit packing + validity mask                (GossReadBaseString::getBase, src/GossReadBaseString.hh:133-170)
//   K3  window extraction, reverse complement / FNV normalisation, fused radix-digit histograms
//                                                    (firstKmer/nextKmer src/GossReadBaseString.hh:52-103,
//                                                     ReverseComplementAdapter.hh:34-55, RankSelect.hh:126-140)
//
// Design: the reference walks each read with a rolling cursor on one CPU thread.  Here every
// record is first flattened into ONE symbol stream (2-bit code + valid bit per symbol, record
// boundaries and non-ACGT bytes become invalid symbols, FASTA/FASTQ line breaks inside a record
// disappear), so that a window is valid iff its rho valid bits are all set and its key is a
// plain bit-field of the stream.  One thread per stream position extracts its window directly
// (two funnel shifts), which keeps all loads and stores coalesced and needs no per-read
// scheduling; rolling is a CPU idiom that would serialise a warp.
#include "kernels.h"
#include "scan.cuh"

namespace gsb {

// ------------------------------------------------------------------------------------------
// K1: newline scan
// ------------------------------------------------------------------------------------------
static const int kNlThreads = 256;
static const int kNlChunksPerThread = 4;
static const int kNlTileBytes = kNlThreads * kNlChunksPerThread * 16;   // 16 KiB

__device__ __forceinline__ uint4 load_chunk16(const u8* __restrict__ text, u64 off, u64 n) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (off + 16 <= n) {
        v = *reinterpret_cast<const uint4*>(text + off);     // text is 16-byte aligned (checked on the host)
    } else if (off < n) {
        u32 w[4] = {0, 0, 0, 0};
        for (u64 i = off; i < n; ++i) w[(i - off) >> 2] |= (u32)text[i] << (8 * ((i - off) & 3));
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return v;
}

__device__ __forceinline__ u32 newline_mask16(const uint4& v) {
    // bit i set iff byte i of the chunk is '\n'
    u32 m = 0;
    const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        u32 eq = __vcmpeq4(w[j], 0x0A0A0A0Au);               // 0xFF per matching byte
        m |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * j);
    }
    return m;
}

__global__ void __launch_bounds__(kNlThreads) count_newlines_kernel(const u8* __restrict__ text, u64 n, u32* __restrict__ tile_counts) {
    __shared__ u32 sm[kNlThreads / 32 + 1];
    const u64 tile_base = (u64)blockIdx.x * kNlTileBytes;
    u32 c = 0;
#pragma unroll
    for (int j = 0; j < kNlChunksPerThread; ++j) {
        u64 off = tile_base + ((u64)j * kNlThreads + threadIdx.x) * 16;
        c += __popc(newline_mask16(load_chunk16(text, off, n)));
    }
    u32 total;
    block_exclusive_scan<u32, kNlThreads>(c, &total, sm);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

// line_start[L] = offset of the first byte of line L; line_start[n_lines] = one past the
// (possibly virtual) terminator of the last line, so len(L) = line_start[L+1] - line_start[L] - 1.
__global__ void __launch_bounds__(kNlThreads) fill_line_starts_kernel(const u8* __restrict__ text, u64 n, const u32* __restrict__ tile_offsets,
                                                                     u32* __restrict__ line_start, u32 n_lines) {
    __shared__ u32 sm[kNlThreads / 32 + 1];
    const u64 tile_base = (u64)blockIdx.x * kNlTileBytes;
    u32 base = tile_offsets[blockIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        line_start[0] = 0;
        if (n > 0 && text[n - 1] != '\n') line_start[n_lines] = (u32)n + 1;   // unterminated last line
    }
    for (int j = 0; j < kNlChunksPerThread; ++j) {
        u64 off = tile_base + ((u64)j * kNlThreads + threadIdx.x) * 16;
        u32 m = newline_mask16(load_chunk16(text, off, n));
        u32 total;
        u32 ex = block_exclusive_scan<u32, kNlThreads>(__popc(m), &total, sm) + base;
        while (m) {
            int b = __ffs(m) - 1;
            m &= m - 1;
            line_start[++ex] = (u32)(off + b + 1);            // the line after newline #ex starts here
        }
        base += total;
    }
}

// ------------------------------------------------------------------------------------------
// K1b: per-line framing
// ------------------------------------------------------------------------------------------
enum LineKind : u8 { LK_SKIP = 0, LK_SEP = 1, LK_SEQ = 2, LK_LINE = 3 };

__device__ __forceinline__ u32 line_len(const u32* __restrict__ line_start, u32 L) { return line_start[L + 1] - line_start[L] - 1; }

__global__ void classify_fasta_kernel(const u8* __restrict__ text, const u32* __restrict__ line_start, u32 n_lines, int file_start,
                                      u8* __restrict__ kind, u32* __restrict__ nsym, IngestStatus* st) {
    u32 reads = 0;
    for (u32 L = blockIdx.x * blockDim.x + threadIdx.x; L < n_lines; L += gridDim.x * blockDim.x) {
        u32 len = line_len(line_start, L);
        bool header = len > 0 && text[line_start[L]] == '>';
        if (header) { kind[L] = LK_SEP; nsym[L] = 1; ++reads; }
        else {
            kind[L] = LK_SEQ; nsym[L] = len;                  // '\r' is kept and becomes an invalid symbol (src/FastaParser.hh:85)
            if (L == 0 && file_start) { st->error = GSB_PE_FASTA_EXPECT_GT; st->error_line = 0; }
        }
    }
    reads = __reduce_add_sync(0xffffffffu, reads);
    if ((threadIdx.x & 31) == 0 && reads) atomicAdd(&st->n_reads, (u64)reads);
}

__global__ void classify_line_kernel(const u32* __restrict__ line_start, u32 n_lines, u8* __restrict__ kind, u32* __restrict__ nsym, IngestStatus* st) {
    for (u32 L = blockIdx.x * blockDim.x + threadIdx.x; L < n_lines; L += gridDim.x * blockDim.x) {
        kind[L] = LK_LINE; nsym[L] = 1 + line_len(line_start, L);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&st->n_reads, (u64)n_lines);
}

// FASTQ pass 1 (parallel): per line, class of the first byte and length without one trailing '\r'.
//   kind[L] <- 0 other / 1 '@' / 2 '+'   (only meaningful when the stripped length is > 0)
//   nsym[L] <- stripped length
__global__ void fastq_line_info_kernel(const u8* __restrict__ text, const u32* __restrict__ line_start, u32 n_lines,
                                       u8* __restrict__ kind, u32* __restrict__ nsym) {
    for (u32 L = blockIdx.x * blockDim.x + threadIdx.x; L < n_lines; L += gridDim.x * blockDim.x) {
        u32 s = line_start[L], len = line_len(line_start, L);
        if (len > 0 && text[s + len - 1] == '\r') --len;
        u8 c = 0;
        if (len > 0) { u8 ch = text[s]; c = ch == '@' ? 1 : (ch == '+' ? 2 : 0); }
        kind[L] = c; nsym[L] = len;
    }
}

__device__ bool fastq_label_matches(const u8* __restrict__ text, const u32* __restrict__ line_start, const u32* __restrict__ slen, u32 hdr, u32 plus) {
    u32 n = slen[plus] - 1;
    if (n == 0) return true;                                   // bare '+'
    if (slen[hdr] - 1 != n) return false;
    const u8* a = text + line_start[hdr] + 1; const u8* b = text + line_start[plus] + 1;
    for (u32 i = 0; i < n; ++i) if (a[i] != b[i]) return false;
    return true;
}

// FASTQ pass 2 (parallel, the common 4-line layout): verifies that the reference's state
// machine would frame every record as header/sequence/plus/quality; any doubt -> st->fastq_irregular.
__global__ void fastq_check_regular_kernel(const u8* __restrict__ text, const u32* __restrict__ line_start, u32 n_lines,
                                           const u8* __restrict__ c0, const u32* __restrict__ slen, IngestStatus* st) {
    u32 bad = 0;
    const u32 n_rec = n_lines / 4;
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        u32 L = 4 * r;
        bool ok = slen[L] > 0 && c0[L] == 1;                                   // '@' header
        ok = ok && !(slen[L + 1] > 0 && c0[L + 1] != 0);                       // sequence line is not '@'/'+'
        ok = ok && slen[L + 2] > 0 && c0[L + 2] == 2;                          // '+' line
        ok = ok && slen[L + 3] == slen[L + 1];                                 // quality as long as sequence
        ok = ok && (slen[L + 1] > 0);                                          // empty reads go the careful way
        if (ok && slen[L + 2] > 1) ok = fastq_label_matches(text, line_start, slen, L, L + 2);
        if (!ok) bad = 1;
    }
    if (bad) st->fastq_irregular = 1;
}

__global__ void fastq_apply_regular_kernel(u32 n_lines, u8* __restrict__ kind, u32* __restrict__ nsym, IngestStatus* st) {
    for (u32 L = blockIdx.x * blockDim.x + threadIdx.x; L < n_lines; L += gridDim.x * blockDim.x) {
        u32 ph = L & 3;
        if (ph == 0) { kind[L] = LK_SEP; nsym[L] = 1; }
        else if (ph == 1) { kind[L] = LK_SEQ; /* nsym already = stripped length */ }
        else { kind[L] = LK_SKIP; nsym[L] = 0; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&st->n_reads, (u64)(n_lines / 4));
}

// FASTQ pass 2' (one thread): the reference's exact state machine over the line table, for
// wrapped records, '@'/'+' leading quality lines, stray empty lines and every error case.
__global__ void fastq_sequential_kernel(const u8* __restrict__ text, const u32* __restrict__ line_start, u32 n_lines,
                                        u8* __restrict__ kind, u32* __restrict__ nsym, u64 line_base, IngestStatus* st) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u32 L = 0;
    u64 line_num = 1 + line_base;
    u64 reads = 0;
    auto fail = [&](int code) { st->error = code; st->error_line = line_num; st->n_reads += reads; };
    while (L < n_lines) {
        if (!(nsym[L] > 0 && kind[L] == 1)) { fail(GSB_PE_FASTQ_EXPECT_AT); return; }
        const u32 hdr = L;
        const u32 hdr_len = nsym[L];
        u64 seq_len = 0;
        for (;;) {
            ++L; ++line_num;
            if (L >= n_lines) { fail(GSB_PE_FASTQ_EXPECT_SEQ); return; }
            if (nsym[L] > 0 && kind[L] != 0) break;
            kind[L] = LK_SEQ; seq_len += nsym[L];
        }
        if (kind[L] != 2) { fail(GSB_PE_FASTQ_EXPECT_PLUS); return; }
        if (nsym[L] > 1) {
            bool same = (hdr_len == nsym[L]);
            if (same) {
                const u8* a = text + line_start[hdr] + 1; const u8* b = text + line_start[L] + 1;
                for (u32 i = 0; i + 1 < hdr_len; ++i) if (a[i] != b[i]) { same = false; break; }
            }
            if (!same) { fail(GSB_PE_FASTQ_TITLE_MISMATCH); return; }
        }
        kind[L] = LK_SKIP; nsym[L] = 0;
        kind[hdr] = LK_SEP; nsym[hdr] = 1;
        u64 qual_len = 0;
        for (;;) {
            ++L; ++line_num;
            if (L >= n_lines) break;
            if (nsym[L] > 0 && kind[L] != 0 && qual_len >= seq_len) break;
            qual_len += nsym[L];
            kind[L] = LK_SKIP; nsym[L] = 0;
        }
        if (seq_len != qual_len) { fail(GSB_PE_FASTQ_LEN_MISMATCH); return; }
        ++reads;
    }
    st->n_reads += reads;
}

// FASTQ pass 2'' (parallel, irregular layouts): SPECULATE, THEN VERIFY.  The state machine above is sequential because a
// line's role depends on everything before it (a quality line may start with '@' or '+').  The line table is cut into
// segments of kFqSegLines lines;
//   fastq_anchor_kernel : one thread per segment looks for the first '@' line at or behind the segment's start from which
//                         the machine frames kFqLookahead well-formed records in a row -- a quality line posing as a header
//                         fails that within a record or two;
//   fastq_segment_kernel: one thread per segment runs the machine from its anchor to the first header at or behind the
//                         segment's end and checks that this is exactly the next segment's anchor.
// Segment 0 starts at line 0, a record boundary by contract, so if every hand-over matches, the framing IS the sequential
// one, by induction.  Any mismatch or parse error raises a flag and the block is framed again by the one-thread kernel,
// which also produces the reference's exact error text.  Results go to separate arrays: the inputs stay intact for it.
static const u32 kFqSegLines = 2048;
static const int kFqLookahead = 3;
static const u32 kFqNone = 0xffffffffu;

// One record starting at header line L.  Returns the line behind the record (the next header, or n_lines), or kFqNone when
// the record is malformed or runs into the end of the table before it is complete (need_complete) .  Optionally writes the
// framing.  seq/qual lengths as in fastq_sequential_kernel.
__device__ __forceinline__ u32 fastq_one_record(const u8* __restrict__ text, const u32* __restrict__ line_start, u32 n_lines, const u8* __restrict__ c0,
                                                const u32* __restrict__ slen, u32 L, u8* __restrict__ kind_out, u32* __restrict__ nsym_out) {
    if (!(slen[L] > 0 && c0[L] == 1)) return kFqNone;
    const u32 hdr = L, hdr_len = slen[L];
    u64 seq_len = 0;
    for (;;) {
        ++L;
        if (L >= n_lines) return kFqNone;
        if (slen[L] > 0 && c0[L] != 0) break;
        if (kind_out) { kind_out[L] = LK_SEQ; nsym_out[L] = slen[L]; }
        seq_len += slen[L];
    }
    if (c0[L] != 2) return kFqNone;
    if (slen[L] > 1) {
        if (hdr_len != slen[L]) return kFqNone;
        const u8* a = text + line_start[hdr] + 1; const u8* b = text + line_start[L] + 1;
        for (u32 i = 0; i + 1 < hdr_len; ++i) if (a[i] != b[i]) return kFqNone;
    }
    if (kind_out) { kind_out[L] = LK_SKIP; nsym_out[L] = 0; kind_out[hdr] = LK_SEP; nsym_out[hdr] = 1; }
    u64 qual_len = 0;
    for (;;) {
        ++L;
        if (L >= n_lines) break;
        if (slen[L] > 0 && c0[L] != 0 && qual_len >= seq_len) break;
        qual_len += slen[L];
        if (kind_out) { kind_out[L] = LK_SKIP; nsym_out[L] = 0; }
    }
    if (seq_len != qual_len) return kFqNone;
    return L;
}

__global__ void fastq_anchor_kernel(const u8* __restrict__ text, const u32* __restrict__ line_start, u32 n_lines, const u8* __restrict__ c0,
                                    const u32* __restrict__ slen, u32 n_seg, u32* __restrict__ anchor) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_seg) return;
    if (i == 0) { anchor[0] = 0; return; }
    const u32 lo = i * kFqSegLines, hi = min(n_lines, lo + kFqSegLines);
    u32 found = kFqNone;
    for (u32 L = lo; L < hi && found == kFqNone; ++L) {
        if (!(slen[L] > 0 && c0[L] == 1)) continue;
        u32 p = L;
        bool ok = true;
        for (int r = 0; r < kFqLookahead && ok && p < n_lines; ++r) {
            p = fastq_one_record(text, line_start, n_lines, c0, slen, p, nullptr, nullptr);
            ok = p != kFqNone;
        }
        if (ok) found = L;
    }
    anchor[i] = found;
}

__global__ void fastq_segment_kernel(const u8* __restrict__ text, const u32* __restrict__ line_start, u32 n_lines, const u8* __restrict__ c0,
                                     const u32* __restrict__ slen, u32 n_seg, const u32* __restrict__ anchor, u8* __restrict__ kind_out,
                                     u32* __restrict__ nsym_out, u32* __restrict__ failed, u64* __restrict__ n_reads) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_seg) return;
    u32 L = anchor[i];
    if (L == kFqNone) return;                                  // no record starts in this segment: an earlier one runs through it
    u32 next = n_lines;                                        // where the next segment that has an anchor takes over
    for (u32 j = i + 1; j < n_seg; ++j) if (anchor[j] != kFqNone) { next = anchor[j]; break; }
    u32 reads = 0;
    while (L < next) {
        L = fastq_one_record(text, line_start, n_lines, c0, slen, L, kind_out, nsym_out);
        if (L == kFqNone) { atomicOr(failed, 1u); return; }    // malformed (or truncated) record: the one-thread kernel renders the error
        ++reads;
    }
    if (L != next) { atomicOr(failed, 1u); return; }           // the hand-over does not match: speculation failed
    if (reads) atomicAdd(n_reads, (u64)reads);
}

// ------------------------------------------------------------------------------------------
// K2: pack.  One thread per 32 output symbols -> one u64 of 2-bit codes (symbol j of the word
// at bits [62-2j, 63-2j]) and one u32 of valid bits (symbol j at bit 31-j).
// Stream layout: [64 invalid pad symbols][n_carry carried symbols][this block's symbols].
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 base_code(u8 c) {
    // A/a 0, C/c 1, G/g 2, T/t 3, anything else 4
    switch (c | 0x20) {
        case 'a': return 0;
        case 'c': return 1;
        case 'g': return 2;
        case 't': return 3;
        default: return 4;
    }
}

// 32 sequence bytes starting at text + off (any alignment) -> packed codes + valid bits, four bytes per
// step with SIMD-in-a-word arithmetic: upper-case, code = t ^ (t >> 1) with t = (c >> 1) & 3
// (A 0, C 1, G 2, T 3), valid = byte is one of A C G T; a multiply gathers the four 2-bit codes
// (or the four valid bits) of a word into one byte (nibble).
__device__ __forceinline__ void pack32_fast(const u8* __restrict__ text, u64 off, u64& cw, u32& vw) {
    const u32* base = reinterpret_cast<const u32*>(text + (off & ~3ull));
    const u32 sh = (u32)(off & 3) * 8;
    u32 w[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) w[i] = base[i];
    cw = 0; vw = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const u32 x = __funnelshift_r(w[i], w[i + 1], sh);      // bytes off+4i .. off+4i+3, first byte lowest
        const u32 u = x & 0xDFDFDFDFu;
        const u32 t = (u >> 1) & 0x03030303u;
        const u32 c = t ^ ((t >> 1) & 0x01010101u);
        const u32 ok = (__vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u)) & 0x01010101u;
        const u32 cm = c & (ok * 3u);                           // invalid symbols carry code 0
        const u32 c8 = (cm * 0x40100401u) >> 24;                // c0<<6 | c1<<4 | c2<<2 | c3
        const u32 v4 = ((ok * 0x08040201u) >> 24) & 0xFu;       // v0<<3 | v1<<2 | v2<<1 | v3
        cw |= (u64)c8 << (56 - 8 * i);
        vw |= v4 << (28 - 4 * i);
    }
}

// word_line[w] = the line that holds the first symbol of stream word w, for every word that starts inside this block's
// symbols: one thread per line marks the (typically five) words that begin in it.  The packing kernel used to find its
// line with a binary search over sym_off -- 23 dependent loads per word, 60 % of its stall samples (ncu).
__global__ void __launch_bounds__(256) word_lines_kernel(const u32* __restrict__ sym_off, u32 n_lines, u64 first_block_sym, u64 n_words,
                                                         u32* __restrict__ word_line) {
    const u32 L = blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= n_lines) return;
    const u64 a = first_block_sym + sym_off[L], b = first_block_sym + sym_off[L + 1];
    for (u64 w = (a + 31) >> 5; (w << 5) < b && w < n_words; ++w) word_line[w] = L;
}

__global__ void __launch_bounds__(256) pack_symbols_kernel(const u8* __restrict__ text, u64 text_bytes, const u32* __restrict__ line_start,
                                                           const u8* __restrict__ kind, const u32* __restrict__ sym_off /* n_lines+1 */,
                                                           u32 n_lines, const u8* __restrict__ carry, u32 n_carry, u64 n_sym_total,
                                                           u64* __restrict__ codes, u32* __restrict__ valid, u64 n_words,
                                                           const u32* __restrict__ word_line) {
    const u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const u64 s0 = w * 32;
    const u64 first_block_sym = 64 + n_carry;
    u64 cw = 0; u32 vw = 0;
    // locate the line of the first in-block symbol this thread touches
    long long L = -1; u32 line_lo = 0, line_hi = 0, src = 0; u8 lk = LK_SKIP;
    auto seek = [&](u32 x) {                                   // x = symbol index relative to this block's first symbol
        u32 lo = 0, hi = n_lines;                              // upper_bound over sym_off[0..n_lines]
        while (lo < hi) { u32 mid = (lo + hi) >> 1; if (sym_off[mid] <= x) lo = mid + 1; else hi = mid; }
        L = (long long)lo - 1;
        line_lo = sym_off[L]; line_hi = sym_off[L + 1]; src = line_start[L]; lk = kind[L];
    };
    // Words made only of this block's symbols: walk the (one to three, typically) line segments that
    // overlap the word and convert each with the vectorised routine, masked to the segment and shifted
    // into place.  (A per-symbol loop here cost 105 warp instructions per symbol, ncu.)
    if (s0 >= first_block_sym && s0 + 32 <= n_sym_total) {
        u32 x = (u32)(s0 - first_block_sym);
        L = (long long)word_line[w];
        line_lo = sym_off[L]; line_hi = sym_off[L + 1]; src = line_start[L]; lk = kind[L];
        int j = 0;
        while (j < 32) {
            while (x >= line_hi) { ++L; line_lo = line_hi; line_hi = sym_off[L + 1]; src = line_start[L]; lk = kind[L]; }
            const u32 r = x - line_lo;
            u32 seg = line_hi - x;
            if (seg > (u32)(32 - j)) seg = 32 - j;
            if (lk == LK_SEP || (lk == LK_LINE && r == 0)) { ++j; ++x; continue; }   // one invalid symbol
            const u64 off = (u64)src + (lk == LK_LINE ? r - 1 : r);
            u64 c = 0; u32 v = 0;
            if (off + 36 <= text_bytes) {
                pack32_fast(text, off, c, v);
            } else {
                for (u32 i = 0; i < seg; ++i) {
                    const u32 code = base_code(text[off + i]);
                    if (code < 4) { c |= (u64)code << (2 * (31 - i)); v |= 1u << (31 - i); }
                }
            }
            if (seg < 32) { c &= ~0ull << (2 * (32 - seg)); v &= ~0u << (32 - seg); }
            cw |= c >> (2 * j); vw |= v >> j;
            j += seg; x += seg;
        }
        codes[w] = cw; valid[w] = vw;
        return;
    }
    // the few words that touch the pad, the carried symbols or the end of the stream: symbol by symbol
    for (int j = 0; j < 32; ++j) {
        const u64 s = s0 + j;
        u32 code = 4;
        if (s < 64 || s >= n_sym_total) code = 4;
        else if (s < first_block_sym) code = carry[s - 64];
        else {
            const u32 x = (u32)(s - first_block_sym);
            if (L < 0) seek(x);
            while (x >= line_hi) { ++L; line_lo = line_hi; line_hi = sym_off[L + 1]; src = line_start[L]; lk = kind[L]; }
            const u32 r = x - line_lo;
            if (lk == LK_SEQ) code = base_code(text[src + r]);
            else if (lk == LK_LINE) code = r == 0 ? 4 : base_code(text[src + r - 1]);
        }
        if (code < 4) { cw |= (u64)code << (2 * (31 - j)); vw |= 1u << (31 - j); }
    }
    codes[w] = cw; valid[w] = vw;
}

// the last min(want, available) symbols of the stream, oldest first, as one byte each (4 = invalid)
__global__ void save_carry_kernel(const u64* __restrict__ codes, const u32* __restrict__ valid, u64 n_sym_total, u32 want, u8* __restrict__ carry_out) {
    u32 i = threadIdx.x;
    if (i >= want) return;
    u64 s = n_sym_total - want + i;                            // caller guarantees n_sym_total - 64 >= want
    u64 w = s >> 5; u32 j = (u32)(s & 31);
    bool ok = (valid[w] >> (31 - j)) & 1u;
    carry_out[i] = ok ? (u8)((codes[w] >> (2 * (31 - j))) & 3u) : (u8)4;
}

// ------------------------------------------------------------------------------------------
// K3: window extraction
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 fnv16(u64 lo, u64 hi) {
    u64 h = 14695981039346656037ull;
#pragma unroll
    for (int i = 0; i < 8; ++i) { h ^= lo & 0xFFull; lo >>= 8; h *= 1099511628211ull; }
#pragma unroll
    for (int i = 0; i < 8; ++i) { h ^= hi & 0xFFull; hi >>= 8; h *= 1099511628211ull; }
    return h;
}

template <typename K> struct WindowOps;

template <> struct WindowOps<u64> {
    // window of w symbols ending at stream position p (w <= 32)
    __device__ static __forceinline__ u64 load(const u64* __restrict__ codes, u64 b, int sh, int w) {
        u64 x = codes[b] >> sh;
        if (sh) x |= codes[b - 1] << (64 - sh);
        return w >= 32 ? x : (x & ((1ull << (2 * w)) - 1));
    }
    __device__ static __forceinline__ u64 rc(u64 x, int w) { return key_rc(x, w); }
    __device__ static __forceinline__ u64 hash(u64 x) { return fnv16(x, 0); }
};

template <> struct WindowOps<Key128> {
    __device__ static __forceinline__ Key128 load(const u64* __restrict__ codes, u64 b, int sh, int w) {
        u64 c0 = codes[b], c1 = codes[b - 1], c2 = codes[b - 2];
        Key128 k;
        k.lo = sh ? ((c0 >> sh) | (c1 << (64 - sh))) : c0;
        k.hi = sh ? ((c1 >> sh) | (c2 << (64 - sh))) : c1;
        int hb = 2 * w - 64;                                   // significant bits of the high word (w > 32 here)
        if (hb < 64) k.hi &= (1ull << hb) - 1;
        return k;
    }
    __device__ static __forceinline__ Key128 rc(const Key128& x, int w) { return key_rc(x, w); }
    __device__ static __forceinline__ u64 hash(const Key128& x) { return fnv16(x.lo, x.hi); }
};

static const int kExThreads = 256;

// Graph mode emits ONE key per window: the smaller of the window and its reverse complement
// ("strand folding").  The reference pushes x and rc(x) (src/ReverseComplementAdapter.hh:34-55), so
// in its multiset count(y) == count(rc y) == #windows whose canonical key is min(y, rc y), doubled
// when y is its own reverse complement.  Counting the folded keys halves every later pass over the
// instances; fold.cu restores both strands after the run-length reduce.
static const int kExItems = 4;                                  // stream positions per thread per iteration

// PASSES > 0 unrolls the histogram update (7 = k 25, 8 = k 31, 14 = k 55); 0 = run-time count; -1 = ONE histogram, of
// the top kTopHistBits bits of the low key word (the first partition pass of partition.cu splits by those or fewer bits).
template <typename K, int MODE, int PASSES>
__global__ void __launch_bounds__(kExThreads) extract_kernel(const u64* __restrict__ codes, const u32* __restrict__ valid,
                                                             u64 p_begin, u64 p_end, int w, int passes_rt, int mix,
                                                             K* __restrict__ out, u64* __restrict__ cursor, u64 capacity,
                                                             u64* __restrict__ digit_hist /* [passes][256] */, IngestStatus* st) {
    typedef KeyOps<K> KO;
    typedef WindowOps<K> WO;
    const int passes = PASSES > 0 ? PASSES : (PASSES < 0 ? (1 << kTopHistBits) / 256 : passes_rt);
    extern __shared__ u32 hist_s[];                            // [passes][256] (PASSES < 0: [2^kTopHistBits])
    // The stream words a tile needs (its own 32-33 words and the two in front) are staged in shared memory ONE TILE AHEAD:
    // ncu showed every warp waiting on its own first-touch loads of codes / valid (long scoreboard 38 % of the stall
    // samples).  Everything that is reused across iterations is double-buffered, which leaves two barriers per tile.
    constexpr int kStage = kExThreads * kExItems / 32 + 4;     // words per tile incl. the two leading ones and misalignment
    __shared__ u64 codes_s[2][kStage];
    __shared__ u32 valid_s[2][kStage];
    __shared__ u32 warp_cnt[2][kExThreads / 32];
    __shared__ u64 base_s[2];
    for (int i = threadIdx.x; i < passes * 256; i += kExThreads) hist_s[i] = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 wmask = w >= 64 ? ~0ull : ((1ull << w) - 1);
    const u32 lt = (1u << lane) - 1;
    constexpr u64 kSpan = (u64)kExThreads * kExItems;
    const int t = threadIdx.x;
    // thread t < kStage fetches code word t of a tile, thread 64 + t valid word t (two different warps)
    const bool f_code = t < kStage, f_valid = t >= 64 && t < 64 + kStage;
    const int f_idx = f_code ? t : t - 64;
    auto fetch = [&](u64 tile, u64& c, u32& v) {
        c = 0; v = 0;
        if (tile >= p_end) return;
        const u64 b = (tile >> 5) - 2 + (u64)f_idx;            // tile >= 64: the stream starts with 64 pad symbols
        if ((b << 5) >= p_end) return;
        if (f_code) c = codes[b];
        if (f_valid) v = valid[b];
    };
    u64 tile = p_begin + (u64)blockIdx.x * kSpan;
    {
        u64 c; u32 v;
        fetch(tile, c, v);
        if (f_code) codes_s[0][f_idx] = c;
        if (f_valid) valid_s[0][f_idx] = v;
    }
    __syncthreads();
    int cur = 0;
    for (; tile < p_end; tile += (u64)gridDim.x * kSpan, cur ^= 1) {
        u64 pre_c; u32 pre_v;
        fetch(tile + (u64)gridDim.x * kSpan, pre_c, pre_v);    // in flight during this tile's arithmetic
        const u64* cs = codes_s[cur];
        const u32* vs_ = valid_s[cur];
        const u64 b0 = (tile >> 5) - 2;
        K x[kExItems];
        u32 ballot[kExItems];
        u32 wtot = 0;
#pragma unroll
        for (int it = 0; it < kExItems; ++it) {
            const u64 p = tile + (u64)it * kExThreads + threadIdx.x;
            bool ok = false;
            x[it] = KO::make(0, 0);
            if (p < p_end) {
                const int b = (int)((p >> 5) - b0); const int j = (int)(p & 31);
                const int vs = 31 - j;
                u64 v = ((u64)vs_[b] >> vs) | ((u64)vs_[b - 1] << (32 - vs));
                if (vs) v |= (u64)vs_[b - 2] << (64 - vs);
                ok = (v & wmask) == wmask;
                if (ok) {
                    x[it] = WO::load(cs, (u64)b, 2 * vs, w);
                    const K r = WO::rc(x[it], w);
                    if (MODE == GSB_KIND_GRAPH) {               // fold the two strands
                        if (KO::lt(r, x[it])) x[it] = r;
                        else if (KO::eq(r, x[it])) atomicAdd(&st->n_self_rc, 1ull);   // rare: lets the reduce skip its self-complement test
                    } else {                                    // position_type::normalize, src/RankSelect.hh:126-140
                        const u64 h0 = WO::hash(x[it]), h1 = WO::hash(r);
                        if (h0 > h1 || (h0 == h1 && KO::lt(r, x[it]))) x[it] = r;
                    }
                    if (mix) x[it] = key_mix(x[it]);            // instances will only be grouped, not ordered (partition.cu)
                }
            }
            ballot[it] = __ballot_sync(0xffffffffu, ok);
            wtot += __popc(ballot[it]);
        }
        if (lane == 0) warp_cnt[cur][warp] = wtot;
        if (f_code) codes_s[cur ^ 1][f_idx] = pre_c;
        if (f_valid) valid_s[cur ^ 1][f_idx] = pre_v;
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 tot = 0;
#pragma unroll
            for (int i = 0; i < kExThreads / 32; ++i) { u32 c = warp_cnt[cur][i]; warp_cnt[cur][i] = tot; tot += c; }
            base_s[cur] = tot ? atomicAdd(cursor, (u64)tot) : 0;
        }
        __syncthreads();
        u32 slot = warp_cnt[cur][warp];
        const u64 base = base_s[cur];
#pragma unroll
        for (int it = 0; it < kExItems; ++it) {
            if ((ballot[it] >> lane) & 1u) {
                const u64 idx = base + (u64)(slot + __popc(ballot[it] & lt));
                if (idx < capacity) out[idx] = x[it];
                else st->error = GSB_PE_KEY_OVERFLOW;
                if (PASSES < 0) {
                    atomicAdd(&hist_s[(u32)(KO::lo(x[it]) >> (64 - kTopHistBits))], 1u);
                } else if (PASSES > 0) {
#pragma unroll
                    for (int d = 0; d < (PASSES > 0 ? PASSES : 1); ++d) atomicAdd(&hist_s[d * 256 + KO::digit(x[it], 8 * d)], 1u);
                } else {
                    for (int d = 0; d < passes; ++d) atomicAdd(&hist_s[d * 256 + KO::digit(x[it], 8 * d)], 1u);
                }
            }
            slot += __popc(ballot[it]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += kExThreads) {
        u32 c = hist_s[i];
        if (c) atomicAdd(&digit_hist[i], (u64)c);
    }
}

// ------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------
u32 ingest_newline_tiles(u64 nbytes) { return (u32)((nbytes + kNlTileBytes - 1) / kNlTileBytes); }

void ingest_count_newlines(const u8* text, u64 n, u32* tile_counts, cudaStream_t s, u64* launches) {
    u32 tiles = ingest_newline_tiles(n);
    if (!tiles) return;
    count_newlines_kernel<<<tiles, kNlThreads, 0, s>>>(text, n, tile_counts);
    ++*launches;
}

void ingest_scan_tiles(u32* tile_counts, u32 tiles, u32* total, cudaStream_t s, u64* launches) {
    scan_single_cta_kernel<u32><<<1, 1024, 0, s>>>(tile_counts, tiles, total);
    ++*launches;
}

void ingest_fill_line_starts(const u8* text, u64 n, const u32* tile_offsets, u32* line_start, u32 n_lines, cudaStream_t s, u64* launches) {
    u32 tiles = ingest_newline_tiles(n);
    if (!tiles) return;
    fill_line_starts_kernel<<<tiles, kNlThreads, 0, s>>>(text, n, tile_offsets, line_start, n_lines);
    ++*launches;
}

static inline int line_grid(u32 n) { int g = (int)((n + 255) / 256); return g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g); }

__global__ void fastq_add_reads_kernel(IngestStatus* st, u64 reads) { st->n_reads += reads; }

u64 ingest_fastq_scratch_words(u32 n_lines) {                  // u32 words of scratch for the speculative FASTQ framing
    const u64 n_seg = (n_lines + kFqSegLines - 1) / kFqSegLines;
    return ((n_seg + 2 + 1) & ~1ull) + 2 + (u64)n_lines + ((u64)n_lines + 3) / 4 + 4;
}

void ingest_classify(const u8* text, const u32* line_start, u32 n_lines, int format, int file_start, u64 line_base,
                     u8* kind, u32* nsym, IngestStatus* st_dev, cudaStream_t s, u64* launches, u32* fq_scratch) {
    if (n_lines == 0) return;
    const int g = line_grid(n_lines);
    if (format == GSB_FMT_FASTA) {
        classify_fasta_kernel<<<g, 256, 0, s>>>(text, line_start, n_lines, file_start, kind, nsym, st_dev);
        ++*launches;
    } else if (format == GSB_FMT_LINE) {
        classify_line_kernel<<<g, 256, 0, s>>>(line_start, n_lines, kind, nsym, st_dev);
        ++*launches;
    } else {
        fastq_line_info_kernel<<<g, 256, 0, s>>>(text, line_start, n_lines, kind, nsym);
        ++*launches;
        bool regular = (n_lines % 4) == 0;
        IngestStatus h;
        if (regular) {
            fastq_check_regular_kernel<<<line_grid(n_lines / 4), 256, 0, s>>>(text, line_start, n_lines, kind, nsym, st_dev);
            ++*launches;
            GSB_CUDA_TRY(cudaMemcpyAsync(&h, st_dev, sizeof(h), cudaMemcpyDeviceToHost, s));
            GSB_CUDA_TRY(cudaStreamSynchronize(s));
            regular = !h.fastq_irregular;
        }
        if (regular) {
            fastq_apply_regular_kernel<<<g, 256, 0, s>>>(n_lines, kind, nsym, st_dev);
            ++*launches;
        } else {
            // irregular layout: speculative parallel framing, verified; the one-thread machine only if that fails
            bool framed = false;
            if (fq_scratch && n_lines > kFqSegLines) {
                const u32 n_seg = (n_lines + kFqSegLines - 1) / kFqSegLines;
                u32* anchor = fq_scratch;                                       // [n_seg]
                u32* failed = fq_scratch + n_seg;                               // [1] (+1 pad)
                u64* reads = reinterpret_cast<u64*>(fq_scratch + ((n_seg + 2 + 1) & ~1u));   // [1], 8-byte aligned
                u32* nsym_out = reinterpret_cast<u32*>(reads + 1);              // [n_lines]
                u8* kind_out = reinterpret_cast<u8*>(nsym_out + n_lines);       // [n_lines]
                GSB_CUDA_TRY(cudaMemsetAsync(failed, 0, 8, s));
                GSB_CUDA_TRY(cudaMemsetAsync(reads, 0, 8, s));
                fastq_anchor_kernel<<<(n_seg + 127) / 128, 128, 0, s>>>(text, line_start, n_lines, kind, nsym, n_seg, anchor);
                fastq_segment_kernel<<<(n_seg + 127) / 128, 128, 0, s>>>(text, line_start, n_lines, kind, nsym, n_seg, anchor, kind_out, nsym_out, failed, reads);
                *launches += 2;
                u32 h_failed = 1; u64 h_reads = 0;
                GSB_CUDA_TRY(cudaMemcpyAsync(&h_failed, failed, 4, cudaMemcpyDeviceToHost, s));
                GSB_CUDA_TRY(cudaMemcpyAsync(&h_reads, reads, 8, cudaMemcpyDeviceToHost, s));
                GSB_CUDA_TRY(cudaStreamSynchronize(s));
                if (!h_failed) {
                    GSB_CUDA_TRY(cudaMemcpyAsync(kind, kind_out, n_lines, cudaMemcpyDeviceToDevice, s));
                    GSB_CUDA_TRY(cudaMemcpyAsync(nsym, nsym_out, (size_t)n_lines * 4, cudaMemcpyDeviceToDevice, s));
                    fastq_add_reads_kernel<<<1, 1, 0, s>>>(st_dev, h_reads);
                    ++*launches;
                    framed = true;
                }
            }
            if (!framed) {
                fastq_sequential_kernel<<<1, 32, 0, s>>>(text, line_start, n_lines, kind, nsym, line_base, st_dev);
                ++*launches;
            }
        }
    }
}

void ingest_symbol_offsets(const u32* nsym, u32* sym_off, u32 n_lines, u32* total_dev, u32* tmp, cudaStream_t s, u64* launches) {
    // sym_off[L] = symbols before line L; sym_off[n_lines] = total (written by the caller from *total_dev)
    exclusive_scan<u32, u32>(nsym, sym_off, n_lines, 0u, total_dev, tmp, s, launches);
}

void ingest_pack(const u8* text, u64 text_bytes, const u32* line_start, const u8* kind, const u32* sym_off, u32 n_lines,
                 const u8* carry, u32 n_carry, u64 n_sym_total, u64* codes, u32* valid, u64 n_words, u32* word_line /* scratch [n_words] */,
                 cudaStream_t s, u64* launches) {
    if (!n_words) return;
    if (n_lines) {
        word_lines_kernel<<<(n_lines + 255) / 256, 256, 0, s>>>(sym_off, n_lines, 64 + (u64)n_carry, n_words, word_line);
        ++*launches;
    }
    pack_symbols_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, s>>>(text, text_bytes, line_start, kind, sym_off, n_lines, carry, n_carry,
                                                                        n_sym_total, codes, valid, n_words, word_line);
    ++*launches;
}

void ingest_save_carry(const u64* codes, const u32* valid, u64 n_sym_total, u32 want, u8* carry_out, cudaStream_t s, u64* launches) {
    if (!want) return;
    save_carry_kernel<<<1, 64, 0, s>>>(codes, valid, n_sym_total, want, carry_out);
    ++*launches;
}

template <typename K, int PASSES>
static void launch_extract_p(int kind, const u64* codes, const u32* valid, u64 p_begin, u64 p_end, int w, int passes, int mix,
                             K* out, u64* cursor, u64 capacity, u64* digit_hist, IngestStatus* st, int sm_count, cudaStream_t s) {
    const u64 span = (u64)kExThreads * kExItems;
    u64 tiles = (p_end - p_begin + span - 1) / span;
    int grid = (int)(tiles < (u64)sm_count * 8 ? tiles : (u64)sm_count * 8);
    size_t smem = (size_t)(passes < 0 ? (1 << kTopHistBits) / 256 : passes) * 256 * sizeof(u32);
    if (kind == GSB_KIND_GRAPH)
        extract_kernel<K, GSB_KIND_GRAPH, PASSES><<<grid, kExThreads, smem, s>>>(codes, valid, p_begin, p_end, w, passes, mix, out, cursor, capacity, digit_hist, st);
    else
        extract_kernel<K, GSB_KIND_KMERSET, PASSES><<<grid, kExThreads, smem, s>>>(codes, valid, p_begin, p_end, w, passes, mix, out, cursor, capacity, digit_hist, st);
}

template <typename K>
static void launch_extract(int kind, const u64* codes, const u32* valid, u64 p_begin, u64 p_end, int w, int passes, int mix,
                           K* out, u64* cursor, u64 capacity, u64* digit_hist, IngestStatus* st, int sm_count, cudaStream_t s) {
    switch (passes) {
        case -1: launch_extract_p<K, -1>(kind, codes, valid, p_begin, p_end, w, passes, mix, out, cursor, capacity, digit_hist, st, sm_count, s); break;
        case 7: launch_extract_p<K, 7>(kind, codes, valid, p_begin, p_end, w, passes, mix, out, cursor, capacity, digit_hist, st, sm_count, s); break;
        case 8: launch_extract_p<K, 8>(kind, codes, valid, p_begin, p_end, w, passes, mix, out, cursor, capacity, digit_hist, st, sm_count, s); break;
        case 14: launch_extract_p<K, 14>(kind, codes, valid, p_begin, p_end, w, passes, mix, out, cursor, capacity, digit_hist, st, sm_count, s); break;
        default: launch_extract_p<K, 0>(kind, codes, valid, p_begin, p_end, w, passes, mix, out, cursor, capacity, digit_hist, st, sm_count, s); break;
    }
}

void ingest_extract(int kind, int key_bytes, const u64* codes, const u32* valid, u64 p_begin, u64 p_end, int w, int passes, int mix,
                    void* out, u64* cursor, u64 capacity, u64* digit_hist, IngestStatus* st, int sm_count, cudaStream_t s, u64* launches) {
    if (p_end <= p_begin) return;
    if (key_bytes == 8) launch_extract<u64>(kind, codes, valid, p_begin, p_end, w, passes, mix, (u64*)out, cursor, capacity, digit_hist, st, sm_count, s);
    else launch_extract<Key128>(kind, codes, valid, p_begin, p_end, w, passes, mix, (Key128*)out, cursor, capacity, digit_hist, st, sm_count, s);
    ++*launches;
}

}  // namespace gsb
