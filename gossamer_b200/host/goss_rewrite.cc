// goss_rewrite.cc -- the commands that read existing file sets:
//   trim-graph      src/GossCmdTrimGraph.cc:27-185      (explicit -C cutoff; the inferred cutoff is not part of this path)
//   merge-graphs / merge-kmer-sets   src/GossCmdMerge.tcc:148-383   (including the grouping into temporary file sets
//                                    when there are more than --max-merge inputs: the last merge's size estimate depends on it)
//   dump-graph      src/GossCmdDumpGraph.cc:31-60
//   restore-graph   src/GossCmdRestoreGraph.cc:70-128
//   merge-and-annotate-kmer-sets   src/GossCmdMergeAndAnnotateKmerSets.cc:27-229   (xenome index, step 3)
//   compute-near-kmers             src/GossCmdComputeNearKmers.cc:158-247          (xenome index, step 4)
// All decoding, merging, filtering and writing happens on the GPU through the C ABI (gsb_graph_*); this file is option
// handling, the reference's messages, and the text parser of restore-graph.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iostream>
#include <sstream>

#include <unistd.h>

#include "goss_cmd.hh"

namespace goss {

namespace {

void library_log(void* user, int sev, const char* msg) { (*(Logger*)user)((Severity)sev, msg); }

struct Ctx {
    gsb_ctx* c = nullptr;
    ~Ctx() { if (c) gsb_destroy(c); }
    void create(int kind, uint64_t k, int device, Logger& log) {
        gsb_config cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.abi_version = GSB_ABI_VERSION;
        cfg.kind = kind;
        cfg.k = (int32_t)k;
        cfg.device = device;
        cfg.min_count = 1;
        cfg.log = &library_log;
        cfg.log_user = &log;
        if (gsb_create(&cfg, &c) != GSB_OK) throw Error{std::string(gsb_last_error(nullptr)) + "\n"};
    }
    void check(int rc) { if (rc != GSB_OK) throw Error{std::string(gsb_last_error(c)) + "\n"}; }
};

gsb_graph_info peek(const std::string& prefix, InputFiles& in, int kind) {
    gsb_graph_info info;
    char err[512] = {0};
    if (gsb_graph_peek(prefix.c_str(), in.source(), kind, &info, err, sizeof(err)) != GSB_OK) throw Error{std::string(err) + "\n"};
    return info;
}

uint64_t parse_u64(const std::string& opt, const std::string& v) {
    char* end = nullptr;
    errno = 0;
    unsigned long long x = strtoull(v.c_str(), &end, 10);
    if (errno || !end || *end || v.empty() || v[0] == '-')
        throw Error{"the argument ('" + v + "') for option '" + opt + "' is invalid\n"};
    return x;
}

// one merge of file sets `ins` into `out` (GossCmdMerge<T>::merge, src/GossCmdMerge.tcc:208-296)
void merge_once(const std::vector<std::string>& ins, const std::string& out, int kind, int device, Logger& log) {
    InputFiles in;
    uint64_t k = 0, tot = 0;
    bool asymmetric = false;
    for (size_t i = 0; i < ins.size(); ++i) {
        const gsb_graph_info gi = peek(ins[i], in, kind);
        const bool asym = kind == GSB_KIND_GRAPH && (gi.flags & 1);
        if (i == 0) { k = gi.k; asymmetric = asym; }
        else {
            if (gi.k != k)
                throw Error{"all graphs involved in a merge must have the same kmer-size.\n" + ins[0] + " has k=" + std::to_string(k) + ".\n" + ins[i] +
                            " has k=" + std::to_string(gi.k) + ".\n\n"};
            if (asym != asymmetric)
                throw Error{"graphs involved in a merge must either all preserve sense or not.\n" + ins[0] + (asymmetric ? " preserves sense" : " does not preserve sense") +
                            ".\n" + ins[i] + (asym ? " preserves sense" : " does not preserve sense") + ".\n\n"};
        }
        log(info, " " + ins[i] + " " + std::to_string(gi.n_items));
        tot += gi.n_items;
    }
    Ctx h;
    h.create(kind, k, device, log);
    log(info, "starting graph merge");
    auto t0 = std::chrono::steady_clock::now();
    for (const std::string& p : ins) h.check(gsb_graph_load(h.c, p.c_str(), in.source()));
    gsb_counts counts;
    h.check(gsb_graph_finish(h.c, 0, tot, &counts));          // Builder(k, out, fac, tot): the SUM of the inputs' sizes
    OutputFiles files;
    int rc = gsb_emit(h.c, out.c_str(), files.sink());
    if (rc == GSB_EIO) throw Error{"\tcannot write to '" + out + "'\n"};
    h.check(rc);
    log(info, "finishing graph merge");
    log(info, "total build time: " + std::to_string(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()));
}

}  // namespace

void run_trim_graph(const RewriteOptions& o, const GossCmdContext& cxt) {
    Logger& log = cxt.log;
    auto t0 = std::chrono::steady_clock::now();
    InputFiles in;
    log(info, "scanning histogram");
    const gsb_graph_info gi = peek(o.ins[0], in, GSB_KIND_GRAPH);
    if (gi.flags & 1) throw Error{"Asymmetric graphs not yet handled\n"};
    Ctx h;
    h.create(GSB_KIND_GRAPH, gi.k, o.device, log);
    h.check(gsb_graph_load(h.c, o.ins[0].c_str(), in.source()));
    gsb_counts counts;
    h.check(gsb_graph_finish(h.c, o.cutoff, 0, &counts));        // Builder(k, out, fac, n): n = the exact number kept
    log(info, "scanning to trim edges");
    log(info, o.ins[0] + " had " + std::to_string(gi.n_items));
    log(info, o.out + " will have " + std::to_string(counts.n_kept));
    OutputFiles files;
    int rc = gsb_emit(h.c, o.out.c_str(), files.sink());
    if (rc == GSB_EIO) throw Error{"\tcannot write to '" + o.out + "'\n"};
    h.check(rc);
    log(info, "total elapsed time: " + std::to_string(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()));
}

void run_merge(const RewriteOptions& o, const GossCmdContext& cxt, bool kmer_sets) {
    const int kind = kmer_sets ? GSB_KIND_KMERSET : GSB_KIND_GRAPH;
    // GossCmdMerge<T>::operator(), src/GossCmdMerge.tcc:148-205
    std::deque<std::pair<std::string, bool>> todo;
    for (const std::string& s : o.ins) todo.push_back({s, false});
    uint64_t tmp_n = 0;
    const std::string tmp_base = o.out + ".merge-tmp-" + std::to_string((long)getpid()) + "-";
    auto flush_group = [&](size_t n, const std::string& out) {
        std::vector<std::string> ins, to_remove;
        for (size_t i = 0; i < n; ++i) {
            ins.push_back(todo.front().first);
            if (todo.front().second) to_remove.push_back(todo.front().first);
            todo.pop_front();
        }
        merge_once(ins, out, kind, o.device, cxt.log);
        for (const std::string& r : to_remove) remove_file_set(r);
    };
    while (todo.size() > o.max_merge) {
        const std::string out = tmp_base + std::to_string(tmp_n++);
        flush_group(o.max_merge, out);
        todo.push_back({out, true});
    }
    flush_group(todo.size(), o.out);
}

void run_dump_graph(const RewriteOptions& o, const GossCmdContext& cxt) {
    InputFiles in;
    const gsb_graph_info gi = peek(o.ins[0], in, GSB_KIND_GRAPH);
    Ctx h;
    h.create(GSB_KIND_GRAPH, gi.k, o.device, cxt.log);
    h.check(gsb_graph_load(h.c, o.ins[0].c_str(), in.source()));
    gsb_counts counts;
    h.check(gsb_graph_finish(h.c, 0, 0, &counts));
    if (o.text_file == "-") {
        // the sink hands the text over in order; stdout is not seekable
        struct Out {
            static int open(void*, const char*, uint64_t, void** hd) { *hd = (void*)1; return 0; }
            static int pwrite(void*, void*, uint64_t, const void* data, uint64_t len) { return fwrite(data, 1, len, stdout) == len ? 0 : -1; }
            static int close(void*, void*) { return fflush(stdout) == 0 ? 0 : -1; }
        };
        gsb_sink s{nullptr, &Out::open, &Out::pwrite, &Out::close};
        h.check(gsb_graph_dump(h.c, "-", &s));
    } else {
        OutputFiles files;
        int rc = gsb_graph_dump(h.c, o.text_file.c_str(), files.sink());
        if (rc == GSB_EIO) throw Error{"\tcannot write to '" + o.text_file + "'\n"};
        h.check(rc);
    }
}

void run_restore_graph(const RewriteOptions& o, const GossCmdContext& cxt) {
    std::ifstream file;
    std::istream* in = &std::cin;
    if (o.text_file != "-") {
        file.open(o.text_file);
        if (!file) throw Error{"\t'" + o.text_file + "': " + strerror(errno) + "\n"};
        in = &file;
    }
    std::string x;
    std::getline(*in, x);
    if (!in->good()) throw Error{"\t'" + o.text_file + "': unexpected end of file\n"};
    uint64_t k = 0, n = 0, flags = 0;
    *in >> k >> n >> flags;
    if (!in->good()) throw Error{"\t'" + o.text_file + "': unexpected end of file\n"};
    if (flags & 1) throw Error{"Asymmetric graphs not yet handled\n"};
    if (k < 1 || k > 62) throw Error{"unable to build a graph with k=" + std::to_string(k) + "\n"};
    std::vector<uint64_t> lo, hi, cn;
    lo.reserve(n); hi.reserve(n); cn.reserve(n);
    while (in->good()) {
        x.clear();
        uint32_t c = 0;
        *in >> x >> c;
        if (!in->good()) break;
        if (x.size() != k + 1) throw Error{"\t'" + o.text_file + "': sequence " + x + " has wrong length\n"};
        unsigned __int128 v = 0;
        for (char ch : x) {
            unsigned code;
            switch (ch) {
                case 'A': case 'a': code = 0; break;
                case 'C': case 'c': code = 1; break;
                case 'G': case 'g': code = 2; break;
                case 'T': case 't': code = 3; break;
                default: throw Error{"invalid sequence " + x + "\n"};
            }
            v = (v << 2) | code;
        }
        lo.push_back((uint64_t)v); hi.push_back((uint64_t)(v >> 64)); cn.push_back(c);
    }
    Ctx h;
    h.create(GSB_KIND_GRAPH, k, o.device, cxt.log);
    h.check(gsb_graph_load_pairs(h.c, lo.data(), hi.data(), cn.data(), lo.size()));
    gsb_counts counts;
    h.check(gsb_graph_finish(h.c, 0, n, &counts));               // Builder(k, out, fac, n): the header's n
    OutputFiles files;
    int rc = gsb_emit(h.c, o.out.c_str(), files.sink());
    if (rc == GSB_EIO) throw Error{"\tcannot write to '" + o.out + "'\n"};
    h.check(rc);
}

void run_merge_and_annotate(const RewriteOptions& o, const GossCmdContext& cxt) {
    Logger& log = cxt.log;
    auto t0 = std::chrono::steady_clock::now();
    InputFiles in;
    const gsb_graph_info lhs = peek(o.ins[0], in, GSB_KIND_KMERSET), rhs = peek(o.ins[1], in, GSB_KIND_KMERSET);
    if (lhs.n_items == 0 || rhs.n_items == 0 || lhs.k != rhs.k) throw Error{"nonsense\n"};     // the reference's own words (:41-49)
    log(info, "counting kmers.");
    Ctx h;
    h.create(GSB_KIND_KMERSET, lhs.k, o.device, log);
    OutputFiles files;
    uint64_t st[4] = {0, 0, 0, 0};
    int rc = gsb_kmerset_merge_annotate(h.c, o.ins[0].c_str(), o.ins[1].c_str(), in.source(), o.out.c_str(), files.sink(), st);
    if (rc == GSB_EIO) throw Error{"\tcannot write to '" + o.out + "'\n"};
    h.check(rc);
    log(info, "writing out " + std::to_string(st[3]) + " kmers.");
    log(info, "of which " + std::to_string(st[2]) + " are common.");
    std::cout << st[0] << '\t' << st[1] << '\t' << st[2] << '\n';                            // :204
    log(info, "total elapsed time: " + std::to_string(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()));
}

void run_compute_near_kmers(const RewriteOptions& o, const GossCmdContext& cxt) {
    Logger& log = cxt.log;
    auto t0 = std::chrono::steady_clock::now();
    InputFiles in;
    const gsb_graph_info gi = peek(o.ins[0], in, GSB_KIND_KMERSET);
    Ctx h;
    h.create(GSB_KIND_KMERSET, gi.k, o.device, log);
    log(info, "calculating grey set");
    // the bit vectors are rewritten in place, as the reference does (:218-226): collected first, written after the pass
    OutputFiles files;
    uint64_t gray = 0;
    int rc = gsb_kmerset_near_kmers(h.c, o.ins[0].c_str(), in.source(), files.sink(), &gray);
    if (rc == GSB_EIO) throw Error{"\tcannot write to '" + o.ins[0] + "'\n"};
    h.check(rc);
    log(info, "found " + std::to_string(gray) + " gray bits (out of " + std::to_string(gi.n_items) + ").");
    log(info, "total elapsed time: " + std::to_string(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()));
}

std::string rewrite_usage_text(const std::string& cmd) {
    std::ostringstream os;
    if (cmd == "trim-graph") os << "usage: goss trim-graph -G <graph> -O <graph> -C <cutoff>\n";
    else if (cmd == "dump-graph") os << "usage: goss dump-graph -G <graph> [-o <file>]\n";
    else if (cmd == "restore-graph") os << "usage: goss restore-graph [-f <file>] -O <graph>\n";
    else if (cmd == "merge-and-annotate-kmer-sets") os << "usage: goss merge-and-annotate-kmer-sets -G <kmer set> -G <kmer set> -O <kmer set>\n";
    else if (cmd == "compute-near-kmers") os << "usage: goss compute-near-kmers -G <annotated kmer set>\n";
    else os << "usage: goss " << cmd << " {-G <in>}+ [--graphs-in <list>] [--max-merge n] -O <out>\n";
    os << "  -G, --graph-in arg         name of the input graph object (repeatable for merges)\n"
       << "  -O, --graph-out arg        name of the output graph object\n"
       << "  -C, --cutoff arg           coverage cutoff: keep edges with count > cutoff (trim-graph)\n"
       << "      --graphs-in arg        file with one input name per line (merges)\n"
       << "      --max-merge arg        maximum number of graphs to merge at once (default 8)\n"
       << "  -o, --output-file arg      output file name ('-' for standard output; dump-graph)\n"
       << "  -f, --input-file arg       input file name ('-' for standard input; restore-graph)\n"
       << "      --device arg           CUDA device ordinal (default 0)\n"
       << "  -v, --verbose              show progress messages\n"
       << "  -l, --log-file arg         place to write messages\n"
       << "  -h, --help                 show this message\n";
    return os.str();
}

RewriteOptions parse_rewrite_args(const std::string& cmd, int argc, char** argv) {
    RewriteOptions o;
    std::string list;
    auto usage = [&](const std::string& what) -> Error { return Error{what + "use\n\tgoss " + cmd + " -h\nfor more usage information.\n"}; };
    for (int i = 0; i < argc; ++i) {
        std::string a = argv[i], val;
        bool has_inline = false;
        if (a.rfind("--", 0) == 0) {
            size_t eq = a.find('=');
            if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_inline = true; }
        } else if (a.size() > 2 && a[0] == '-' && a[1] != '-') { val = a.substr(2); a = a.substr(0, 2); has_inline = true; }
        auto need = [&]() -> std::string {
            if (has_inline) return val;
            if (i + 1 >= argc) throw usage("the required argument for option '" + a + "' is missing\n");
            return argv[++i];
        };
        if (a == "-h" || a == "--help") o.help = true;
        else if (a == "-v" || a == "--verbose") o.verbose = true;
        else if (a == "-G" || a == "--graph-in") o.ins.push_back(need());
        else if (a == "-O" || a == "--graph-out") o.out = need();
        else if ((a == "-C" || a == "--cutoff") && cmd == "trim-graph") { o.cutoff = parse_u64(a, need()); o.have_cutoff = true; }
        else if (a == "--graphs-in" && (cmd == "merge-graphs" || cmd == "merge-kmer-sets")) list = need();
        else if (a == "--max-merge" && (cmd == "merge-graphs" || cmd == "merge-kmer-sets")) o.max_merge = parse_u64(a, need());
        else if ((a == "-o" || a == "--output-file") && cmd == "dump-graph") o.text_file = need();
        else if ((a == "-f" || a == "--input-file") && cmd == "restore-graph") o.text_file = need();
        else if (a == "--device") o.device = (int)parse_u64(a, need());
        else if (a == "-l" || a == "--log-file") o.log_file = need();
        else if (a == "-T" || a == "--num-threads" || a == "--tmp-dir" || a == "-D" || a == "--debug") (void)need();
        else throw usage("unrecognised option '" + std::string(argv[i]) + "'\n");
    }
    if (o.help) return o;
    if (!list.empty()) for (const std::string& n : expand_file_list(list)) o.ins.push_back(n);
    const bool merge = cmd == "merge-graphs" || cmd == "merge-kmer-sets";
    if (cmd == "merge-and-annotate-kmer-sets") {
        if (o.ins.size() != 2) throw usage("the option '--graph-in' must be given exactly twice\n");      // GossOptionChecker::getRepeatingTwice
        if (o.out.empty()) throw usage("the option '--graph-out' is required but missing\n");
        check_output_prefix(o.out);
        return o;
    }
    if (cmd == "compute-near-kmers") {
        if (o.ins.size() != 1) throw usage(o.ins.empty() ? "the option '--graph-in' is required but missing\n" : "the option '--graph-in' may only be given once\n");
        return o;
    }
    if (cmd != "restore-graph") {
        if (o.ins.empty()) {
            if (merge) throw usage("At least one input graph must be supplied either using --graph-in or --graphs-in.\n\n");
            throw usage("the option '--graph-in' is required but missing\n");
        }
        if (!merge && o.ins.size() != 1) throw usage("the option '--graph-in' may only be given once\n");
    }
    if (cmd != "dump-graph") {
        if (o.out.empty()) throw usage("the option '--graph-out' is required but missing\n");
        check_output_prefix(o.out);
    }
    if (cmd == "trim-graph" && !o.have_cutoff)
        throw usage("this build wants an explicit --cutoff (inferring the cutoff from the histogram is outside the accelerated path)\n");
    if (merge && o.max_merge < 2) o.max_merge = 2;
    return o;
}

}  // namespace goss
