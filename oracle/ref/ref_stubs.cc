// ref_stubs.cc -- link-time stand-ins for the few reference symbols that the command translation
// units reference but that are never reached when the commands are driven the way the reference's
// own tests drive them (command object + StringFileFactory, no argv parsing):
//   * App::{fileFactory,logger,help} and GossOptions::addOpt<bool> -- only used by the
//     GossCmdFactory*::create / option-registration code paths (src/App.cc needs all 70 commands);
//   * EstimateGraphStatistics -- trim-graph's cutoff *inference* (needs boost::numeric::ublas);
//     parity only uses the explicit `-C c` path (src/GossCmdTrimGraph.cc:97-124).
// Every stub throws, so an unexpected call is loud.  TEST INFRASTRUCTURE ONLY.
#include <stdexcept>

#include "App.hh"
#include "EstimateGraphStatistics.hh"
#include "GossOption.hh"

namespace {
[[noreturn]] void unavailable(const char* what) { throw std::runtime_error(std::string(what) + " is not part of the oracle/ref build"); }
}

FileFactory& App::fileFactory() { unavailable("App::fileFactory"); }
Logger& App::logger() { unavailable("App::logger"); }
void App::help(bool) { unavailable("App::help"); }

template <> void GossOptions::addOpt<bool>(const std::string&, const std::string&, const std::string&) {}

EstimateGraphStatistics::EstimateGraphStatistics(Logger&, const std::map<uint64_t, uint64_t>&, double, double) { unavailable("EstimateGraphStatistics"); }
void EstimateGraphStatistics::report(uint64_t) const { unavailable("EstimateGraphStatistics::report"); }
bool EstimateGraphStatistics::modelFits() const { unavailable("EstimateGraphStatistics::modelFits"); }
uint64_t EstimateGraphStatistics::estimateTrimPoint() const { unavailable("EstimateGraphStatistics::estimateTrimPoint"); }
