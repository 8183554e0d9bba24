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
