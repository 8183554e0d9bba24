// goss_cmd.cc -- see goss_cmd.hh
#include "goss_cmd.hh"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <sstream>

#include <signal.h>
#include <sys/wait.h>
#include <unistd.h>

namespace goss {

void Logger::operator()(Severity s, const std::string& msg) {
    if (sev_ > s) return;
    time_t rt; time(&rt);
    std::string now = asctime(localtime(&rt));
    now.erase(now.size() - 1);
    static const char* names[] = {"info", "warning", "error"};
    out_ << now << '\t' << names[s] << '\t' << msg << std::endl;
}

namespace {

void library_log(void* user, int sev, const char* msg) { (*(Logger*)user)((Severity)sev, msg); }

struct Ctx {
    gsb_ctx* c = nullptr;
    ~Ctx() { if (c) gsb_destroy(c); }
};

[[noreturn]] void fail_from_library(gsb_ctx* c, int rc, const std::string& file) {
    std::string m = gsb_last_error(c);
    if (rc == GSB_EPARSE) throw Error{"\t'" + file + "': " + m + "\n"};       // src/App.cc:352-358
    throw Error{m + "\n"};
}

// ---- several GPUs: one worker PROCESS per device -----------------------------------------------------------------------------
// The seam is the body of GossCmdBuildGraph::operator() (src/GossCmdBuildGraph.cc:270-426); the reference has no distributed
// mode.  The launcher forks one worker per device before anything touches CUDA; worker 0 makes the NCCL id and the launcher
// relays it over pipes; every worker attaches, pushes its share of the input (FASTQ / line files: every n-th block, cut at
// record boundaries; FASTA files: whole files, because a record may straddle blocks), and after the collective finish /
// emit writes ITS byte ranges of every output file with pwrite into the shared files.
bool read_all(int fd, void* dst, size_t n) {
    size_t done = 0;
    while (done < n) {
        ssize_t r = ::read(fd, (char*)dst + done, n - done);
        if (r < 0) { if (errno == EINTR) continue; return false; }
        if (r == 0) return false;
        done += (size_t)r;
    }
    return true;
}
bool write_all(int fd, const void* src, size_t n) {
    size_t done = 0;
    while (done < n) {
        ssize_t w = ::write(fd, (const char*)src + done, n - done);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        done += (size_t)w;
    }
    return true;
}

int build_worker(const BuildOptions& opt, const GossCmdContext& cxt, int kind, int rank, int n, int id_in_fd, int id_out_fd) {
    Logger& log = cxt.log;
    try {
        auto t0 = std::chrono::steady_clock::now();
        unsigned char id[GSB_NCCL_ID_BYTES];
        if (rank == 0) {
            if (gsb_comm_make_id(id) != GSB_OK) throw Error{std::string(gsb_last_error(nullptr)) + "\n"};
            if (!write_all(id_out_fd, id, sizeof(id))) throw Error{"cannot hand the communicator id to the launcher\n"};
        } else if (!read_all(id_in_fd, id, sizeof(id))) {
            throw Error{"no communicator id from the launcher\n"};
        }
        gsb_host_bind_near_device(opt.devices[rank]);             // pinned block buffers on the GPU's own NUMA node
        gsb_config cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.abi_version = GSB_ABI_VERSION;
        cfg.kind = kind;
        cfg.k = (int32_t)opt.K;
        cfg.device = opt.devices[rank];
        cfg.min_count = kind == GSB_KIND_GRAPH ? opt.min_count : 1;
        cfg.log = &library_log;
        cfg.log_user = &log;
        Ctx h;
        int rc = gsb_create(&cfg, &h.c);
        if (rc != GSB_OK) throw Error{std::string(gsb_last_error(nullptr)) + "\n"};
        rc = gsb_comm_attach(h.c, id, n, rank);
        if (rc != GSB_OK) fail_from_library(h.c, rc, opt.out);
        struct Item { const std::vector<std::string>* names; int format; };
        const Item items[3] = {{&opt.lines, GSB_FMT_LINE}, {&opt.fastas, GSB_FMT_FASTA}, {&opt.fastqs, GSB_FMT_FASTQ}};
        uint64_t n_files = 0, block_no = 0, fasta_no = 0;
        for (const Item& it : items)
            for (const std::string& name : *it.names) {
                ++n_files;
                if (it.format == GSB_FMT_FASTA) {                       // whole FASTA files, round robin
                    if ((int)(fasta_no++ % (uint64_t)n) != rank) continue;
                    if (rank == 0 || opt.fastas.size() > 1) log(info, "reading " + name);
                    BlockReader rd(name, it.format, (size_t)opt.block_mb << 20);
                    const uint8_t* data; size_t size; bool last;
                    while (rd.next(data, size, last)) {
                        rc = gsb_push_block(h.c, data, size, it.format, last ? GSB_BLOCK_LAST_OF_FILE : GSB_BLOCK_ASYNC);
                        if (rc != GSB_OK) fail_from_library(h.c, rc, name);
                    }
                    continue;
                }
                if (rank == 0) log(info, "reading " + name);
                BlockReader rd(name, it.format, (size_t)opt.block_mb << 20);
                const uint8_t* data; size_t size; bool last;
                while (rd.next(data, size, last)) {                     // every worker scans the file; block b belongs to worker b mod n
                    if ((int)(block_no++ % (uint64_t)n) != rank) continue;
                    rc = gsb_push_block(h.c, data, size, it.format, GSB_BLOCK_LAST_OF_FILE | (last ? 0u : GSB_BLOCK_ASYNC));
                    if (rc != GSB_OK) fail_from_library(h.c, rc, name);
                }
            }
        if (n_files == 0) throw Error{"No valid reads.\n"};
        gsb_counts counts;
        if (rank == 0) log(info, "sorting and counting...");
        rc = gsb_finish_counting(h.c, &counts);
        if (rc != GSB_OK) fail_from_library(h.c, rc, opt.out);
        if (counts.n_instances == 0 && kind == GSB_KIND_GRAPH) throw Error{"No valid reads.\n"};
        OutputFiles files(true);
        rc = gsb_emit(h.c, opt.out.c_str(), files.sink());
        if (rc != GSB_OK) {
            if (rc == GSB_EIO) throw Error{"\tcannot write to '" + opt.out + "'\n"};
            fail_from_library(h.c, rc, opt.out);
        }
        if (rank == 0) {
            gsb_stats st;
            gsb_get_stats(h.c, &st);
            std::ostringstream os;
            os << n << " GPUs: " << counts.n_instances << " instances, " << counts.n_distinct << " distinct, " << counts.n_kept
               << " kept; rank 0 device ms: scan " << st.ms_scan << " extract " << st.ms_extract << " exchange " << st.ms_exchange << " count " << st.ms_reduce
               << " emit " << st.ms_emit;
            log(info, os.str());
            log(info, "total build time: " + std::to_string(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()));
        }
        return 0;
    } catch (const Error& e) {
        std::cerr << "error performing " << cxt.cmdName << " (worker " << rank << ", device " << opt.devices[rank] << "):\n" << e.text;
        return 1;
    } catch (const std::exception& e) {
        std::cerr << "caught unexpected exception in worker " << rank << ": " << e.what() << std::endl;
        return 1;
    }
}

void run_build_multi(const BuildOptions& opt, const GossCmdContext& cxt, int kind) {
    const int n = (int)opt.devices.size();
    remove_file_set(opt.out);                                          // shared files are opened without truncation
    int up[2];
    if (pipe(up) != 0) throw Error{"pipe failed\n"};
    std::vector<int> down_r(n, -1), down_w(n, -1);
    std::vector<pid_t> pids(n, -1);
    for (int r = 1; r < n; ++r) { int p[2]; if (pipe(p) != 0) throw Error{"pipe failed\n"}; down_r[r] = p[0]; down_w[r] = p[1]; }
    fflush(nullptr);
    for (int r = 0; r < n; ++r) {
        pid_t pid = fork();
        if (pid < 0) throw Error{"fork failed\n"};
        if (pid == 0) {
            close(up[0]);
            for (int q = 1; q < n; ++q) { close(down_w[q]); if (q != r) close(down_r[q]); }
            const int rc = build_worker(opt, cxt, kind, r, n, down_r[r], up[1]);
            fflush(nullptr);
            _exit(rc);
        }
        pids[r] = pid;
    }
    close(up[1]);
    for (int r = 1; r < n; ++r) close(down_r[r]);
    unsigned char id[GSB_NCCL_ID_BYTES];
    const bool have_id = read_all(up[0], id, sizeof(id));
    close(up[0]);
    for (int r = 1; r < n; ++r) { if (have_id) write_all(down_w[r], id, sizeof(id)); close(down_w[r]); }
    bool failed = !have_id;
    int left = n;
    while (left > 0) {
        int status = 0;
        pid_t done = wait(&status);
        if (done < 0) { if (errno == EINTR) continue; break; }
        --left;
        const bool ok = WIFEXITED(status) && WEXITSTATUS(status) == 0;
        if (!ok && !failed) {                                          // one worker failed: the others would wait in a collective for ever
            failed = true;
            for (int r = 0; r < n; ++r) if (pids[r] != done) kill(pids[r], SIGTERM);
        }
    }
    if (failed) throw Error{"a worker process failed (see above)\n"};
}

void run_build(const BuildOptions& opt, const GossCmdContext& cxt, int kind) {
    if (opt.devices.size() > 1) { run_build_multi(opt, cxt, kind); return; }
    Logger& log = cxt.log;
    auto t0 = std::chrono::steady_clock::now();
    gsb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = GSB_ABI_VERSION;
    cfg.kind = kind;
    cfg.k = (int32_t)opt.K;
    cfg.device = opt.devices.size() == 1 ? opt.devices[0] : opt.device;
    cfg.min_count = kind == GSB_KIND_GRAPH ? opt.min_count : 1;
    cfg.max_batch_keys = 0;
    cfg.log = &library_log;
    cfg.log_user = &log;
    Ctx h;
    int rc = gsb_create(&cfg, &h.c);
    if (rc != GSB_OK) throw Error{std::string(gsb_last_error(nullptr)) + "\n"};

    // The reference's progress lines (src/GossCmdBuildGraph.cc:310-419, src/GossCmdBuildKmerSet.tcc:221-331) are kept where
    // they have a meaning here; the hash-table lines (slot bits, table bits, load, spills) have none: nothing is hashed.
    log(info, "accumulating edges.");
    // line files first, then FASTA, then FASTQ (src/GossCmdBuildGraph.cc:284-300)
    struct Item { const std::vector<std::string>* names; int format; };
    const Item items[3] = {{&opt.lines, GSB_FMT_LINE}, {&opt.fastas, GSB_FMT_FASTA}, {&opt.fastqs, GSB_FMT_FASTQ}};
    uint64_t n_files = 0, n_blocks = 0;
    for (const Item& it : items)
        for (const std::string& name : *it.names) {
            ++n_files;
            log(info, "reading " + name);
            BlockReader rd(name, it.format, (size_t)opt.block_mb << 20);
            const uint8_t* data; size_t size; bool last;
            while (rd.next(data, size, last)) {
                // all but the last block of a file are pushed asynchronously: the copy of block i+1 (and the file read
                // behind it -- BlockReader double-buffers in pinned memory) overlaps the device work of block i; the last
                // block is synchronous, so that a parse error is always reported while its file is the current one
                rc = gsb_push_block(h.c, data, size, it.format, last ? GSB_BLOCK_LAST_OF_FILE : GSB_BLOCK_ASYNC);
                if (rc != GSB_OK) fail_from_library(h.c, rc, name);
                if ((++n_blocks & 15) == 0) {                                    // the reference ticks on its hash-table load (:345-358)
                    gsb_stats pst;
                    gsb_get_stats(h.c, &pst);
                    log(info, "processed " + std::to_string(pst.n_symbols) + " symbols in " + std::to_string(pst.bytes_in) + " input bytes.");
                }
            }
        }
    if (n_files == 0) throw Error{"No valid reads.\n"};                          // src/ReverseComplementAdapter.hh:77-86
    gsb_counts counts;
    log(info, "sorting and counting...");
    rc = gsb_finish_counting(h.c, &counts);
    if (rc != GSB_OK) fail_from_library(h.c, rc, opt.out);
    if (counts.n_reads == 0) throw Error{"No valid reads.\n"};
    log(info, "sorting done.");
    log(info, std::string("estimated number of ") + (kind == GSB_KIND_GRAPH ? "edges" : "kmers") + " is " + std::to_string(counts.n_kept));
    {
        gsb_stats bst;
        gsb_get_stats(h.c, &bst);
        log(info, bst.n_batches > 1 ? "merging temporary graphs" : "writing out graph (no merging necessary).");   // :390 / :407
    }
    OutputFiles files;
    rc = gsb_emit(h.c, opt.out.c_str(), files.sink());
    if (rc != GSB_OK) {
        if (rc == GSB_EIO) throw Error{"\tcannot write to '" + opt.out + "'\n"};   // src/App.cc:360-364
        fail_from_library(h.c, rc, opt.out);
    }
    gsb_stats st;
    gsb_get_stats(h.c, &st);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::ostringstream os;
    os << counts.n_reads << " reads, " << counts.n_instances << " instances, " << counts.n_distinct << " distinct, " << counts.n_kept
       << " kept; device ms: scan " << st.ms_scan << " extract " << st.ms_extract << " sort " << st.ms_sort << " reduce " << st.ms_reduce
       << " merge " << st.ms_merge << " emit " << st.ms_emit << "; " << st.n_batches << " batch(es), " << st.kernel_launches
       << " kernel launches, " << files.bytes_written() << " bytes written";
    log(info, os.str());
    log(info, "finish graph build");                                             // :418-419
    log(info, "total build time: " + std::to_string(secs));
}

uint64_t parse_u64(const std::string& opt, const char* v) {
    char* end = nullptr;
    errno = 0;
    unsigned long long x = strtoull(v, &end, 10);
    if (errno || !end || *end || v[0] == '-' || v[0] == 0)
        throw Error{"the argument ('" + std::string(v) + "') for option '" + opt + "' is invalid\n"};
    return x;
}

}  // namespace

void GossCmdBuildGraph::operator()(const GossCmdContext& cxt) { run_build(opt_, cxt, GSB_KIND_GRAPH); }
void GossCmdBuildKmerSet::operator()(const GossCmdContext& cxt) { run_build(opt_, cxt, GSB_KIND_KMERSET); }

std::string usage_text(const std::string& cmd) {
    std::ostringstream os;
    os << "usage: goss " << cmd << " -k <int> {-I <fasta> | -i <fastq> | --line-in <file> | -F <list> | -f <list>}+ -O <prefix>\n"
       << "  -k, --kmer-size arg        kmer size to use\n"
       << "  -I, --fasta-in arg         input file in FASTA format ('.gz' by suffix, '-' for stdin)\n"
       << "  -i, --fastq-in arg         input file in FASTQ format\n"
       << "      --line-in arg          input file with one sequence per line\n"
       << "  -F, --fastas-in arg        input file containing filenames in FASTA format\n"
       << "  -f, --fastqs-in arg        input file containing filenames in FASTQ format\n"
       << "  -O, --graph-out arg        name of the output " << (cmd == "build-graph" ? "graph" : "kmer set") << " object\n"
       << "  -B, --buffer-size arg      accepted for compatibility (the GPU sizes its own batches)\n"
       << "  -T, --num-threads arg      accepted for compatibility (the work runs on the GPU)\n";
    if (cmd == "build-graph") os << "  -m, --min-count arg        keep edges seen at least this often (== trim-graph -C m-1)\n";
    if (cmd == "build-kmer-set") os << "  -S, --log-hash-slots arg   accepted for compatibility\n";
    os << "      --device arg           CUDA device ordinal (default 0)\n"
       << "      --devices arg          several GPUs, one worker process each: 0,1,2,3 or 0-7\n"
       << "      --block-mb arg         input block size in MiB (default 256, at most 1024)\n"
       << "  -v, --verbose              show progress messages\n"
       << "  -l, --log-file arg         place to write messages\n"
       << "      --tmp-dir arg          accepted for compatibility (no temporary files are used)\n"
       << "  -D, --debug arg            accepted for compatibility\n"
       << "  -h, --help                 show this message\n";
    return os.str();
}

ParsedArgs parse_build_args(const std::string& cmd, int argc, char** argv, uint64_t max_k) {
    ParsedArgs pa;
    BuildOptions& o = pa.opt;
    bool have_k = false, have_o = false;
    std::vector<std::string> fasta_lists, fastq_lists;
    auto usage = [&](const std::string& what) -> Error {
        return Error{what + "use\n\tgoss " + cmd + " -h\nfor more usage information.\n"};
    };
    for (int i = 0; i < argc; ++i) {
        std::string a = argv[i];
        std::string val;
        bool has_inline = false;
        if (a.rfind("--", 0) == 0) {
            size_t eq = a.find('=');
            if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_inline = true; }
        } else if (a.size() > 2 && a[0] == '-' && a[1] != '-') {
            val = a.substr(2); a = a.substr(0, 2); has_inline = true;                 // -k27
        }
        auto need = [&]() -> std::string {
            if (has_inline) return val;
            if (i + 1 >= argc) throw usage("the required argument for option '" + a + "' is missing\n");
            return argv[++i];
        };
        if (a == "-h" || a == "--help") pa.help = true;
        else if (a == "-v" || a == "--verbose") pa.verbose = true;
        else if (a == "-k" || a == "--kmer-size") { o.K = parse_u64(a, need().c_str()); have_k = true; }
        else if (a == "-B" || a == "--buffer-size") o.B = parse_u64(a, need().c_str());
        else if (a == "-T" || a == "--num-threads") o.T = parse_u64(a, need().c_str());
        else if (a == "-O" || a == "--graph-out" || a == "--kmer-set-out") { o.out = need(); have_o = true; }
        else if (a == "-I" || a == "--fasta-in") o.fastas.push_back(need());
        else if (a == "-i" || a == "--fastq-in") o.fastqs.push_back(need());
        else if (a == "--line-in") o.lines.push_back(need());
        else if (a == "-F" || a == "--fastas-in") fasta_lists.push_back(need());
        else if (a == "-f" || a == "--fastqs-in") fastq_lists.push_back(need());
        else if ((a == "-m" || a == "--min-count") && cmd == "build-graph") o.min_count = parse_u64(a, need().c_str());
        else if ((a == "-S" || a == "--log-hash-slots") && cmd == "build-kmer-set") (void)parse_u64(a, need().c_str());
        else if (a == "--device") o.device = (int)parse_u64(a, need().c_str());
        else if (a == "--devices") {                                                   // 0,1,2 or 0-7
            const std::string v = need();
            const size_t dash = v.find('-');
            if (dash != std::string::npos && v.find(',') == std::string::npos) {
                const int lo = (int)parse_u64(a, v.substr(0, dash).c_str()), hi = (int)parse_u64(a, v.substr(dash + 1).c_str());
                if (hi < lo) throw usage("the argument ('" + v + "') for option '--devices' is invalid\n");
                for (int d = lo; d <= hi; ++d) o.devices.push_back(d);
            } else {
                std::stringstream ss(v);
                std::string tok;
                while (std::getline(ss, tok, ',')) o.devices.push_back((int)parse_u64(a, tok.c_str()));
            }
            if (o.devices.empty() || o.devices.size() > 32) throw usage("the argument ('" + v + "') for option '--devices' is invalid\n");
        }
        else if (a == "--block-mb") o.block_mb = parse_u64(a, need().c_str());
        else if (a == "-l" || a == "--log-file") pa.log_file = need();
        else if (a == "--tmp-dir" || a == "-D" || a == "--debug") (void)need();
        else throw usage("unrecognised option '" + std::string(argv[i]) + "'\n");
    }
    if (pa.help) return pa;
    if (!have_k) throw usage("the option '--kmer-size' is required but missing\n");
    if (!have_o) throw usage("the option '--graph-out' is required but missing\n");
    if (o.K > max_k) throw Error{"the value for --kmer-size must be at most " + std::to_string(max_k) + "\n"};   // src/GossCmdBuildGraph.cc:433-434
    if (o.K < 1) throw Error{"the value for --kmer-size must be at least 1\n"};
    if (o.B > 24 && cmd == "build-graph") o.B = 24;                                     // capped silently like :439-443
    if (o.block_mb < 1) o.block_mb = 1;
    if (o.block_mb > 1024) o.block_mb = 1024;
    if (o.min_count < 1) o.min_count = 1;
    for (const std::string& l : fasta_lists) for (const std::string& n : expand_file_list(l)) o.fastas.push_back(n);
    for (const std::string& l : fastq_lists) for (const std::string& n : expand_file_list(l)) o.fastqs.push_back(n);
    check_output_prefix(o.out);
    for (const auto* v : {&o.lines, &o.fastas, &o.fastqs}) for (const std::string& n : *v) check_readable(n);
    return pa;
}

}  // namespace goss
