// goss_cmd.hh -- the two commands of the hot path, shaped like the reference's command objects:
//   GossCmdBuildGraph   (src/GossCmdBuildGraph.{hh,cc}:  ctor(K, B, T, out, fastas, fastqs, lines), operator()(cxt))
//   GossCmdBuildKmerSet (src/GossCmdBuildKmerSet.{hh,cc,tcc})
// and their factories' option handling (src/GossCmdBuildGraph.cc:428-490, src/GossApp.cc:145-202).
#pragma once
#include <cstdint>
#include <iostream>
#include <string>
#include <vector>

#include "file_io.hh"

namespace goss {

enum Severity { info = 0, warning = 1, error = 2 };

// timestamp \t severity \t message (src/Logger.hh:30-133)
class Logger {
public:
    Logger(std::ostream& out, Severity min_sev) : out_(out), sev_(min_sev) {}
    void operator()(Severity s, const std::string& msg);
    Severity sev() const { return sev_; }
private:
    std::ostream& out_;
    Severity sev_;
};

struct GossCmdContext {
    Logger& log;
    std::string cmdName;
};

struct BuildOptions {
    uint64_t K = 0;
    uint64_t B = 2;               // accepted for compatibility: caps the batch like the -B hash budget (GB)
    uint64_t T = 4;               // accepted and ignored: the work runs on the GPU
    uint64_t min_count = 1;       // -m / --min-count (addition; == build-graph then trim-graph -C m-1)
    int device = 0;
    std::vector<int> devices;     // --devices a,b,c | a-b: one worker process per GPU (src/GossCmdBuildGraph.cc:270-426 is the seam;
                                  // the reference itself has no distributed mode)
    uint64_t block_mb = 256;
    std::string out;
    std::vector<std::string> fastas, fastqs, lines;
};

class GossCmdBuildGraph {
public:
    explicit GossCmdBuildGraph(const BuildOptions& o) : opt_(o) {}
    void operator()(const GossCmdContext& cxt);
private:
    BuildOptions opt_;
};

class GossCmdBuildKmerSet {
public:
    explicit GossCmdBuildKmerSet(const BuildOptions& o) : opt_(o) {}
    void operator()(const GossCmdContext& cxt);
private:
    BuildOptions opt_;
};

// trim-graph (src/GossCmdTrimGraph.cc), merge-graphs / merge-kmer-sets (src/GossCmdMerge.tcc), dump-graph
// (src/GossCmdDumpGraph.cc), restore-graph (src/GossCmdRestoreGraph.cc): existing file sets in, a new file set (or text) out
struct RewriteOptions {
    std::vector<std::string> ins;     // -G / --graph-in (repeatable), --graphs-in <list>
    std::string out;                  // -O / --graph-out
    std::string text_file = "-";      // dump-graph -o / restore-graph -f
    uint64_t cutoff = 0;              // trim-graph -C
    bool have_cutoff = false;
    uint64_t max_merge = 8;           // merge: --max-merge
    int device = 0;
    bool verbose = false, help = false;
    std::string log_file;
};
RewriteOptions parse_rewrite_args(const std::string& cmd, int argc, char** argv);
std::string rewrite_usage_text(const std::string& cmd);
void run_trim_graph(const RewriteOptions& o, const GossCmdContext& cxt);
void run_merge(const RewriteOptions& o, const GossCmdContext& cxt, bool kmer_sets);
void run_dump_graph(const RewriteOptions& o, const GossCmdContext& cxt);
void run_restore_graph(const RewriteOptions& o, const GossCmdContext& cxt);
void run_merge_and_annotate(const RewriteOptions& o, const GossCmdContext& cxt);   // xenome index step 3
void run_compute_near_kmers(const RewriteOptions& o, const GossCmdContext& cxt);   // xenome index step 4

struct ParsedArgs {
    BuildOptions opt;
    bool verbose = false;
    bool help = false;
    std::string log_file;
};

// argv (after the command name) -> options; throws Error with usage text on unknown / missing options
ParsedArgs parse_build_args(const std::string& cmd, int argc, char** argv, uint64_t max_k);
std::string usage_text(const std::string& cmd);

}  // namespace goss
