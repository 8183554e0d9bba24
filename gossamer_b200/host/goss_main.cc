// goss_main.cc -- `goss` host executable restricted to the graph-construction hot path:
//   goss build-graph ...      goss build-kmer-set ...      goss help
// Command lookup, option errors, "error performing <cmd>:" rendering and exit codes follow
// src/App.cc:175-419 and src/GossApp.cc:98-203.  Everything compute-related happens in
// libgossamer_b200.so through the C ABI (include/gossamer_b200.h); there is no CPU path.
#include <fstream>
#include <iostream>
#include <memory>

#include "goss_cmd.hh"

using namespace goss;

static int help(std::ostream& os) {
    os << "goss (gossamer_b200) -- B200-native graph construction\n"
       << "commands:\n"
       << "  build-graph       build a de Bruijn graph from FASTA / FASTQ / line input\n"
       << "  build-kmer-set    build a k-mer set from FASTA / FASTQ / line input\n"
       << "  help              show this message\n"
       << "use `goss <command> -h` for the options of a command.\n";
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { help(std::cerr); return 1; }
    const std::string cmd = argv[1];
    if (cmd == "help" || cmd == "-h" || cmd == "--help") return help(std::cout);
    if (cmd != "build-graph" && cmd != "build-kmer-set") {
        std::cerr << "unrecognised command '" << cmd << "'\n";
        help(std::cerr);
        return 1;
    }
    try {
        ParsedArgs pa = parse_build_args(cmd, argc - 2, argv + 2, cmd == "build-graph" ? 62 : 63);
        if (pa.help) { std::cout << usage_text(cmd); return 0; }
        std::unique_ptr<std::ofstream> log_file;
        std::ostream* log_out = &std::cerr;
        if (!pa.log_file.empty()) {
            log_file.reset(new std::ofstream(pa.log_file));
            if (!*log_file) throw Error{"\tcannot write to '" + pa.log_file + "'\n"};
            log_out = log_file.get();
        }
        Logger log(*log_out, pa.verbose ? info : warning);
        GossCmdContext cxt{log, cmd};
        try {
            if (cmd == "build-graph") GossCmdBuildGraph(pa.opt)(cxt);
            else GossCmdBuildKmerSet(pa.opt)(cxt);
        } catch (Error& e) {
            e.text = "error performing " + cmd + ":\n" + e.text;
            throw;
        }
    } catch (const Error& e) {
        std::cerr << e.text;
        if (!e.text.empty() && e.text.back() != '\n') std::cerr << std::endl;
        return 1;
    } catch (const std::exception& e) {
        std::cerr << "caught unexpected exception: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
