// ref_driver.cc -- C entry points over the REAL reference writers/readers, compiled unmodified from
// /root/reference/src against the std::-only Boost shim in ./shim (see ../README.md).
// TEST INFRASTRUCTURE ONLY.  Used to check that the restated oracle (../goss_oracle.hh) writes the
// same bytes as the reference's own Graph::Builder / KmerSet::Builder / SparseArray::Builder /
// DenseSelect::Builder / VariableByteArray::Builder, and that the reference's own readers open
// the files this repository produces.
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "GossCmdBuildGraph.hh"
#include "GossCmdBuildKmerSet.hh"
#include "GossCmdTrimGraph.hh"
#include "GossCmdMergeGraphs.hh"
#include "GossCmdMergeKmerSets.hh"
#include "GossCmdDumpGraph.hh"
#include "GossCmdRestoreGraph.hh"
#include "GossCmdMergeAndAnnotateKmerSets.hh"
#include "GossCmdComputeNearKmers.hh"
#include "Graph.hh"
#include "KmerSet.hh"
#include "Logger.hh"
#include "SparseArray.hh"
#include "StringFileFactory.hh"
#include "VariableByteArray.hh"

using Gossamer::position_type;

namespace {
struct Store {
    StringFileFactory fac;
    std::vector<std::string> names;
    std::map<std::string, std::string> data;
};
position_type make_pos(uint64_t lo, uint64_t hi) {
    position_type p(hi);
    p <<= 64;
    p |= position_type(lo);
    return p;
}
void set_err(char* err, int cap, const std::string& m) {
    if (!err || cap <= 0) return;
    size_t n = std::min<size_t>(m.size(), (size_t)cap - 1);
    memcpy(err, m.data(), n); err[n] = 0;
}
std::string describe(const std::exception& e) {
    const boost::exception* be = dynamic_cast<const boost::exception*>(&e);
    if (be) {
        if (const std::string* m = boost::get_error_info<Gossamer::general_error_info>(e)) return *m;
        if (const std::string* m = boost::get_error_info<Gossamer::parse_error_info>(e)) return *m;
        if (const std::pair<uint64_t, uint64_t>* v = boost::get_error_info<Gossamer::version_mismatch_info>(e))
            return "version mismatch " + std::to_string(v->first) + " vs " + std::to_string(v->second);
    }
    return e.what();
}
}  // namespace

#define REF_TRY try {
#define REF_CATCH } catch (const std::exception& e) { set_err(err, errcap, describe(e)); return -1; } \
                    catch (const char* m) { set_err(err, errcap, m); return -1; } \
                    catch (...) { set_err(err, errcap, "unknown exception"); return -1; }

extern "C" {

void* ref_store_new() { return new Store(); }
void ref_store_free(void* s) { delete (Store*)s; }
void ref_store_put(void* s, const char* name, const void* data, uint64_t n) {
    ((Store*)s)->fac.addFile(name, std::string((const char*)data, n));
}
// snapshot the files the factory holds that start with `prefix`
int ref_store_list(void* sv, const char** candidates, int n_candidates) {
    Store* s = (Store*)sv;
    s->names.clear(); s->data.clear();
    for (int i = 0; i < n_candidates; ++i)
        if (s->fac.fileExists(candidates[i])) { s->names.push_back(candidates[i]); s->data[candidates[i]] = s->fac.readFile(candidates[i]); }
    return (int)s->names.size();
}
const char* ref_store_name(void* s, int i) { return ((Store*)s)->names[i].c_str(); }
uint64_t ref_store_size(void* s, int i) { Store* st = (Store*)s; return st->data[st->names[i]].size(); }
const void* ref_store_data(void* s, int i) { Store* st = (Store*)s; return st->data[st->names[i]].data(); }

// Graph::Builder(K, base, fac, numEdges).push_back(edge,count)... end()   src/Graph.cc:115-167
int ref_write_graph(void* sv, const uint64_t* lo, const uint64_t* hi, const uint64_t* counts, uint64_t m, int k, uint64_t m_est,
                    const char* base, char* err, int errcap) {
    REF_TRY
        Store* s = (Store*)sv;
        Graph::Builder b((uint64_t)k, base, s->fac, m_est);
        for (uint64_t i = 0; i < m; ++i) b.push_back(make_pos(lo[i], hi ? hi[i] : 0), counts[i]);
        b.end();
        return 0;
    REF_CATCH
}

// KmerSet::Builder   src/KmerSet.hh:61-103
int ref_write_kmer_set(void* sv, const uint64_t* lo, const uint64_t* hi, uint64_t m, int k, uint64_t m_est, const char* base,
                       char* err, int errcap) {
    REF_TRY
        Store* s = (Store*)sv;
        KmerSet::Builder b((uint64_t)k, base, s->fac, m_est);
        for (uint64_t i = 0; i < m; ++i) b.push_back(make_pos(lo[i], hi ? hi[i] : 0));
        b.end();
        return 0;
    REF_CATCH
}

// SparseArray::Builder(base, fac, N, M) ... end(N)   src/SparseArray.cc:75-117
int ref_write_sparse_array(void* sv, const uint64_t* lo, const uint64_t* hi, uint64_t m, uint64_t n_lo, uint64_t n_hi, uint64_t m_est,
                           const char* base, char* err, int errcap) {
    REF_TRY
        Store* s = (Store*)sv;
        SparseArray::Builder b(base, s->fac, make_pos(n_lo, n_hi), m_est);
        for (uint64_t i = 0; i < m; ++i) b.push_back(make_pos(lo[i], hi ? hi[i] : 0));
        b.end(make_pos(n_lo, n_hi));
        return 0;
    REF_CATCH
}

// Graph::open + select/multiplicity/rank through the reference's own readers   src/Graph.cc:366-396
int64_t ref_read_graph(void* sv, const char* base, uint64_t* lo, uint64_t* hi, uint32_t* counts, uint64_t cap, uint64_t* k_out,
                       char* err, int errcap) {
    REF_TRY
        Store* s = (Store*)sv;
        GraphPtr gp = Graph::open(base, s->fac);
        Graph& g(*gp);
        const uint64_t n = g.count();
        if (k_out) *k_out = g.K();
        for (uint64_t i = 0; i < n && i < cap; ++i) {
            Graph::Edge e = g.select(i);
            position_type v = e.value();
            if (g.rank(e) != i) throw std::runtime_error("rank(select(i)) != i");
            const position_type::value_type big = v.value();            // keep the words alive
            std::pair<const uint64_t*, const uint64_t*> w = big.words();
            lo[i] = w.first[0]; hi[i] = w.first[1];
            counts[i] = g.multiplicity(i);
        }
        return (int64_t)n;
    REF_CATCH
}

// The whole command, exactly as the reference's own test drives it (src/testGossCmdBuildGraph.cc:120-137):
// GossCmdBuildGraph(K, S, N, T, out, fastas, fastqs, lines)(GossCmdContext(fac, log, "build-graph", opts)).
// S = log2 hash slots, N = slots (src/GossCmdBuildGraph.cc:445-447), T = threads.  Input files are put into
// the StringFileFactory under the given names first.
int ref_build_graph(void* sv, int k, uint64_t S, uint64_t N, uint64_t T, const char* out,
                    const char** fastas, int n_fastas, const char** fastqs, int n_fastqs, const char** lines, int n_lines,
                    char* err, int errcap) {
    REF_TRY
        Store* s = (Store*)sv;
        Logger log("log.txt", s->fac);
        std::vector<std::string> fa(fastas, fastas + n_fastas), fq(fastqs, fastqs + n_fastqs), ln(lines, lines + n_lines);
        GossCmdBuildGraph cmd((uint64_t)k, S, N, T, out, fa, fq, ln);
        boost::program_options::variables_map opts;
        GossCmdContext cxt(s->fac, log, "build-graph", opts);
        cmd(cxt);
        return 0;
    REF_CATCH
}

int ref_build_kmer_set(void* sv, int k, uint64_t S, uint64_t N, uint64_t T, const char* out,
                       const char** fastas, int n_fastas, const char** fastqs, int n_fastqs, const char** lines, int n_lines,
                       char* err, int errcap) {
    REF_TRY
        Store* s = (Store*)sv;
        Logger log("log.txt", s->fac);
        std::vector<std::string> fa(fastas, fastas + n_fastas), fq(fastqs, fastqs + n_fastqs), ln(lines, lines + n_lines);
        GossCmdBuildKmerSet cmd((uint64_t)k, S, N, T, out, fa, fq, ln);
        boost::program_options::variables_map opts;
        GossCmdContext cxt(s->fac, log, "build-kmer-set", opts);
        cmd(cxt);
        return 0;
    REF_CATCH
}

// trim-graph -C c   (src/GossCmdTrimGraph.cc:27-127)
int ref_trim_graph(void* sv, const char* in, const char* out, uint64_t c, char* err, int errcap) {
    REF_TRY
        Store* s = (Store*)sv;
        Logger log("log.txt", s->fac);
        GossCmdTrimGraph cmd(in, out, c, false, false, boost::optional<uint64_t>());
        boost::program_options::variables_map opts;
        GossCmdContext cxt(s->fac, log, "trim-graph", opts);
        cmd(cxt);
        return 0;
    REF_CATCH
}

// merge-graphs / merge-kmer-sets (src/GossCmdMerge.tcc:148-296)
int ref_merge_graphs(void* sv, const char** ins, int n_ins, uint64_t max_merge, const char* out, int kmer_sets, char* err, int errcap) {
    REF_TRY
        Store* s = (Store*)sv;
        Logger log("log.txt", s->fac);
        std::vector<std::string> in(ins, ins + n_ins);
        boost::program_options::variables_map opts;
        if (kmer_sets) {
            GossCmdMergeKmerSets cmd(in, max_merge, out);
            GossCmdContext cxt(s->fac, log, "merge-kmer-sets", opts);
            cmd(cxt);
        } else {
            GossCmdMergeGraphs cmd(in, max_merge, out);
            GossCmdContext cxt(s->fac, log, "merge-graphs", opts);
            cmd(cxt);
        }
        return 0;
    REF_CATCH
}

// dump-graph (src/GossCmdDumpGraph.cc:31-60) and restore-graph (src/GossCmdRestoreGraph.cc:70-128)
int ref_dump_graph(void* sv, const char* in, const char* out_file, char* err, int errcap) {
    REF_TRY
        Store* s = (Store*)sv;
        Logger log("log.txt", s->fac);
        GossCmdDumpGraph cmd(in, out_file);
        boost::program_options::variables_map opts;
        GossCmdContext cxt(s->fac, log, "dump-graph", opts);
        cmd(cxt);
        return 0;
    REF_CATCH
}

int ref_restore_graph(void* sv, const char* in_file, const char* out, char* err, int errcap) {
    REF_TRY
        Store* s = (Store*)sv;
        Logger log("log.txt", s->fac);
        GossCmdRestoreGraph cmd(in_file, out);
        boost::program_options::variables_map opts;
        GossCmdContext cxt(s->fac, log, "restore-graph", opts);
        cmd(cxt);
        return 0;
    REF_CATCH
}

// xenome index, steps 3 and 4 (src/XenoApp.cc:62-76): merge-and-annotate-kmer-sets (src/GossCmdMergeAndAnnotateKmerSets.cc:27-207)
// and compute-near-kmers (src/GossCmdComputeNearKmers.cc:158-225)
int ref_merge_and_annotate(void* sv, const char* lhs, const char* rhs, const char* out, char* err, int errcap) {
    REF_TRY
        Store* s = (Store*)sv;
        Logger log("log.txt", s->fac);
        GossCmdMergeAndAnnotateKmerSets cmd(lhs, rhs, out);
        boost::program_options::variables_map opts;
        GossCmdContext cxt(s->fac, log, "merge-and-annotate-kmer-sets", opts);
        cmd(cxt);
        return 0;
    REF_CATCH
}

int ref_compute_near_kmers(void* sv, const char* in, uint64_t threads, char* err, int errcap) {
    REF_TRY
        Store* s = (Store*)sv;
        Logger log("log.txt", s->fac);
        GossCmdComputeNearKmers cmd(in, threads);
        boost::program_options::variables_map opts;
        GossCmdContext cxt(s->fac, log, "compute-near-kmers", opts);
        cmd(cxt);
        return 0;
    REF_CATCH
}

}  // extern "C"
