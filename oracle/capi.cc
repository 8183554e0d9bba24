// capi.cc -- C entry points over goss_oracle.hh so that Python tests / bench.py can drive the
// CPU restatement through ctypes.  TEST INFRASTRUCTURE ONLY (see goss_oracle.hh header).
#include "goss_oracle.hh"

using namespace goss_oracle;

namespace {
struct FsHandle {
    MemFS fs;
    std::vector<std::string> names;   // stable listing, refreshed on demand
    void refresh() { names.clear(); for (auto& kv : fs) names.push_back(kv.first); }
};
void set_err(char* err, int cap, const std::string& m) {
    if (!err || cap <= 0) return;
    size_t n = std::min<size_t>(m.size(), (size_t)cap - 1);
    memcpy(err, m.data(), n); err[n] = 0;
}
std::vector<Input> to_inputs(const void* p, int n) {
    struct Raw { const void* data; uint64_t size; int format; };
    const Raw* r = (const Raw*)p;
    std::vector<Input> v;
    for (int i = 0; i < n; ++i) v.push_back(Input{(const char*)r[i].data, (size_t)r[i].size, r[i].format});
    return v;
}
}  // namespace

#define ORC_TRY try {
#define ORC_CATCH(rv) } catch (const ParseError& e) { set_err(err, errcap, e.what()); return (rv) - 1; } \
                        catch (const std::exception& e) { set_err(err, errcap, e.what()); return (rv); }

extern "C" {

struct orc_input { const void* data; uint64_t size; int format; };
struct orc_stats { uint64_t n_reads, n_instances, n_distinct, n_kept; double t_extract, t_sort, t_emit; };

void* orc_fs_new() { return new FsHandle(); }
void orc_fs_free(void* h) { delete (FsHandle*)h; }
int orc_fs_count(void* h) { FsHandle* f = (FsHandle*)h; f->refresh(); return (int)f->names.size(); }
const char* orc_fs_name(void* h, int i) { return ((FsHandle*)h)->names[i].c_str(); }
uint64_t orc_fs_size(void* h, int i) { FsHandle* f = (FsHandle*)h; return f->fs[f->names[i]].size(); }
const void* orc_fs_data(void* h, int i) { FsHandle* f = (FsHandle*)h; return f->fs[f->names[i]].data(); }
void orc_fs_put(void* h, const char* name, const void* data, uint64_t n) { ((FsHandle*)h)->fs[name].assign((const char*)data, n); }

// error codes: 0 ok, -1 general error, -2 parse error
int orc_build_graph(const orc_input* in, int n_in, int k, uint64_t min_count, int threads, const char* base, void* fs,
                    orc_stats* st, char* err, int errcap) {
    ORC_TRY
        BuildStats s = build_graph(to_inputs(in, n_in), (unsigned)k, min_count, threads, base, ((FsHandle*)fs)->fs);
        if (st) { st->n_reads = s.n_reads; st->n_instances = s.n_instances; st->n_distinct = s.n_distinct; st->n_kept = s.n_kept;
                  st->t_extract = s.t_extract; st->t_sort = s.t_sort; st->t_emit = s.t_emit; }
        return 0;
    ORC_CATCH(-1)
}

int orc_build_kmer_set(const orc_input* in, int n_in, int k, int threads, const char* base, void* fs,
                       orc_stats* st, char* err, int errcap) {
    ORC_TRY
        BuildStats s = build_kmer_set(to_inputs(in, n_in), (unsigned)k, threads, base, ((FsHandle*)fs)->fs);
        if (st) { st->n_reads = s.n_reads; st->n_instances = s.n_instances; st->n_distinct = s.n_distinct; st->n_kept = s.n_kept;
                  st->t_extract = s.t_extract; st->t_sort = s.t_sort; st->t_emit = s.t_emit; }
        return 0;
    ORC_CATCH(-1)
}

// Window keys in stream order.  Returns the number of keys (which may exceed cap; only cap are stored).
int64_t orc_extract(const orc_input* in, int n_in, int w, int mode, uint64_t* lo, uint64_t* hi, uint64_t cap,
                    uint64_t* n_reads, char* err, int errcap) {
    ORC_TRY
        std::vector<u128> keys;
        extract_keys(to_inputs(in, n_in), (unsigned)w, mode, keys, n_reads);
        for (uint64_t i = 0; i < keys.size() && i < cap; ++i) { if (lo) lo[i] = lo64(keys[i]); if (hi) hi[i] = hi64(keys[i]); }
        return (int64_t)keys.size();
    ORC_CATCH(-1)
}

// The reads themselves (concatenated, '\n'-separated) -- lets tests check framing on its own.
int64_t orc_frame(const orc_input* in, int n_in, char* out, uint64_t cap, uint64_t* n_reads, char* err, int errcap) {
    ORC_TRY
        std::string all; uint64_t reads = 0;
        for (const Input& i : to_inputs(in, n_in))
            for_each_read(i.data, i.size, i.format, [&](const char* s, size_t len) { all.append(s ? s : "", len); all.push_back('\n'); ++reads; });
        if (out) memcpy(out, all.data(), std::min<uint64_t>(cap, all.size()));
        if (n_reads) *n_reads = reads;
        return (int64_t)all.size();
    ORC_CATCH(-1)
}

// multiset -> sorted distinct (key,count) with count >= min_count.  Outputs sized n by the caller.
int64_t orc_count(const uint64_t* lo, const uint64_t* hi, uint64_t n, int key_bits, uint64_t min_count, int threads,
                  uint64_t* out_lo, uint64_t* out_hi, uint64_t* out_counts) {
    std::vector<u128> raw(n);
    for (uint64_t i = 0; i < n; ++i) raw[i] = mk128(hi ? hi[i] : 0, lo[i]);
    std::vector<u128> keys; std::vector<uint64_t> counts;
    count_keys(raw, (unsigned)key_bits, min_count, threads, keys, counts);
    for (uint64_t i = 0; i < keys.size(); ++i) { out_lo[i] = lo64(keys[i]); if (out_hi) out_hi[i] = hi64(keys[i]); out_counts[i] = counts[i]; }
    return (int64_t)keys.size();
}

int orc_write_graph(const uint64_t* lo, const uint64_t* hi, const uint64_t* counts, uint64_t m, int k, uint64_t m_est,
                    const char* base, void* fs, char* err, int errcap) {
    ORC_TRY
        GraphWriter g(((FsHandle*)fs)->fs, base, (uint64_t)k, m_est);
        for (uint64_t i = 0; i < m; ++i) g.push_back(mk128(hi ? hi[i] : 0, lo[i]), counts[i]);
        g.finish();
        return 0;
    ORC_CATCH(-1)
}

int orc_write_kmer_set(const uint64_t* lo, const uint64_t* hi, uint64_t m, int k, uint64_t m_est,
                       const char* base, void* fs, char* err, int errcap) {
    ORC_TRY
        KmerSetWriter s(((FsHandle*)fs)->fs, base, (uint64_t)k, m_est);
        for (uint64_t i = 0; i < m; ++i) s.push_back(mk128(hi ? hi[i] : 0, lo[i]));
        s.finish();
        return 0;
    ORC_CATCH(-1)
}

int orc_write_sparse_array(const uint64_t* lo, const uint64_t* hi, uint64_t m, uint64_t n_lo, uint64_t n_hi, uint64_t m_est,
                           const char* base, void* fs, char* err, int errcap) {
    ORC_TRY
        SparseArrayWriter w(((FsHandle*)fs)->fs, base, mk128(n_hi, n_lo), m_est);
        for (uint64_t i = 0; i < m; ++i) w.push_back(mk128(hi ? hi[i] : 0, lo[i]));
        w.finish(mk128(n_hi, n_lo));
        return 0;
    ORC_CATCH(-1)
}

int orc_write_dense_select(const uint64_t* pos, uint64_t n, int invert, const char* name, void* fs, char* err, int errcap) {
    ORC_TRY
        DenseSelectWriter w(((FsHandle*)fs)->fs[name], invert != 0);
        for (uint64_t i = 0; i < n; ++i) w.push_back(pos[i]);
        w.finish();
        return 0;
    ORC_CATCH(-1)
}

int orc_write_vba(const uint32_t* counts, uint64_t n, uint64_t m_est, const char* base, void* fs, char* err, int errcap) {
    ORC_TRY
        VariableByteArrayWriter w(((FsHandle*)fs)->fs, base, m_est);
        for (uint64_t i = 0; i < n; ++i) w.push_back(counts[i]);
        w.finish();
        return 0;
    ORC_CATCH(-1)
}

int64_t orc_read_graph(void* fs, const char* base, uint64_t* lo, uint64_t* hi, uint32_t* counts, uint64_t cap,
                       uint64_t* k_out, uint64_t* hist_total, int exercise_select, char* err, int errcap) {
    ORC_TRY
        GraphContents g = read_graph(((FsHandle*)fs)->fs, base, exercise_select != 0);
        for (uint64_t i = 0; i < g.edges.size() && i < cap; ++i) {
            if (lo) lo[i] = lo64(g.edges[i]);
            if (hi) hi[i] = hi64(g.edges[i]);
            if (counts) counts[i] = g.counts[i];
        }
        if (k_out) *k_out = g.k;
        if (hist_total) *hist_total = g.hist_total;
        return (int64_t)g.edges.size();
    ORC_CATCH(-1)
}

int64_t orc_read_kmer_set(void* fs, const char* base, uint64_t* lo, uint64_t* hi, uint64_t cap,
                          uint64_t* k_out, uint64_t* count_out, int exercise_select, char* err, int errcap) {
    ORC_TRY
        KmerSetContents s = read_kmer_set(((FsHandle*)fs)->fs, base, exercise_select != 0);
        for (uint64_t i = 0; i < s.kmers.size() && i < cap; ++i) { if (lo) lo[i] = lo64(s.kmers[i]); if (hi) hi[i] = hi64(s.kmers[i]); }
        if (k_out) *k_out = s.k;
        if (count_out) *count_out = s.count;
        return (int64_t)s.kmers.size();
    ORC_CATCH(-1)
}

// select(i) for i in [0,n) through the restated DenseSelect reader
int orc_dense_select_eval(void* fs, const char* bitmap_name, const char* ds_name, int invert, uint64_t n, uint64_t* out,
                          char* err, int errcap) {
    ORC_TRY
        MemFS& m = ((FsHandle*)fs)->fs;
        BitmapReader bits(fs_get(m, bitmap_name));
        DenseSelectReader ds(bits, fs_get(m, ds_name), invert != 0);
        for (uint64_t i = 0; i < n; ++i) out[i] = ds.select(i);
        return 0;
    ORC_CATCH(-1)
}

void orc_reverse_complement(uint64_t lo, uint64_t hi, int k, uint64_t* out_lo, uint64_t* out_hi) {
    u128 r = reverse_complement(mk128(hi, lo), (unsigned)k); *out_lo = lo64(r); *out_hi = hi64(r);
}
void orc_normalize(uint64_t lo, uint64_t hi, int k, uint64_t* out_lo, uint64_t* out_hi) {
    u128 r = normalize(mk128(hi, lo), (unsigned)k); *out_lo = lo64(r); *out_hi = hi64(r);
}
uint64_t orc_fnv_hash(uint64_t lo, uint64_t hi) { return fnv_hash(mk128(hi, lo)); }
uint64_t orc_sparse_d(uint64_t n_lo, uint64_t n_hi, uint64_t m) { return sparse_array_d(mk128(n_hi, n_lo), m); }
void orc_kmer_to_string(uint64_t lo, uint64_t hi, int k, char* out) {
    std::string s = kmer_to_string((unsigned)k, mk128(hi, lo)); memcpy(out, s.data(), s.size()); out[s.size()] = 0;
}

// xenome index steps 3 and 4 on file sets already in `fs`; stats = {n_lhs, n_rhs, n_common, n_out}
int orc_merge_and_annotate(void* fs, const char* lhs, const char* rhs, const char* out, uint64_t* stats, char* err, int errcap) {
    ORC_TRY
        AnnotateStats a = merge_and_annotate(((FsHandle*)fs)->fs, lhs, rhs, out);
        if (stats) { stats[0] = a.n_lhs; stats[1] = a.n_rhs; stats[2] = a.n_common; stats[3] = a.n_out; }
        return 0;
    ORC_CATCH(-1)
}

int64_t orc_compute_near_kmers(void* fs, const char* base, char* err, int errcap) {
    ORC_TRY
        return (int64_t)compute_near_kmers(((FsHandle*)fs)->fs, base);
    ORC_CATCH(-1)
}

}  // extern "C"
