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
       << "  trim-graph        create a new graph by trimming low frequency edges\n"
       << "  merge-graphs      create a new graph by merging existing graphs\n"
       << "  merge-kmer-sets   create a new k-mer set by merging existing k-mer sets\n"
       << "  dump-graph        write out the graph in a robust text representation\n"
       << "  restore-graph     read in a graph from the text representation\n"
       << "  merge-and-annotate-kmer-sets   union of two k-mer sets + membership bit vectors (xenome index)\n"
       << "  compute-near-kmers             mark k-mers of one set that are one variant away from the other (xenome index)\n"
       << "  help              show this message\n"
       << "use `goss <command> -h` for the options of a command.\n";
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { help(std::cerr); return 1; }
    const std::string cmd = argv[1];
    if (cmd == "help" || cmd == "-h" || cmd == "--help") return help(std::cout);
    const bool rewrite = cmd == "trim-graph" || cmd == "merge-graphs" || cmd == "merge-kmer-sets" || cmd == "dump-graph" || cmd == "restore-graph" ||
                         cmd == "merge-and-annotate-kmer-sets" || cmd == "compute-near-kmers";
    if (cmd != "build-graph" && cmd != "build-kmer-set" && !rewrite) {
        std::cerr << "unrecognised command '" << cmd << "'\n";
        help(std::cerr);
        return 1;
    }
    if (rewrite) {
        try {
            RewriteOptions o = parse_rewrite_args(cmd, argc - 2, argv + 2);
            if (o.help) { std::cout << rewrite_usage_text(cmd); return 0; }
            std::unique_ptr<std::ofstream> log_file;
            std::ostream* log_out = &std::cerr;
            if (!o.log_file.empty()) {
                log_file.reset(new std::ofstream(o.log_file));
                if (!*log_file) throw Error{"\tcannot write to '" + o.log_file + "'\n"};
                log_out = log_file.get();
            }
            Logger log(*log_out, o.verbose ? info : warning);
            GossCmdContext cxt{log, cmd};
            try {
                if (cmd == "trim-graph") run_trim_graph(o, cxt);
                else if (cmd == "merge-graphs") run_merge(o, cxt, false);
                else if (cmd == "merge-kmer-sets") run_merge(o, cxt, true);
                else if (cmd == "dump-graph") run_dump_graph(o, cxt);
                else if (cmd == "merge-and-annotate-kmer-sets") run_merge_and_annotate(o, cxt);
                else if (cmd == "compute-near-kmers") run_compute_near_kmers(o, cxt);
                else run_restore_graph(o, cxt);
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
