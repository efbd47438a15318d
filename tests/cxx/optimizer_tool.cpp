// Test tool for host/optimizer.hpp (no GPU): prints the search grids and runs the Pareto helpers on
// statistics read from stdin ("models bf avg_log2 max_log2 size" per line).
//   optimizer_tool first                      -> first-phase configurations, one "models bf" per line
//   optimizer_tool second  < stats            -> second-phase configurations given first-phase results
//   optimizer_tool front <restrict> < stats   -> pareto_front, narrow_front(restrict), sorted by avg log2 error
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../host/optimizer.hpp"

using namespace rmihost;

static std::vector<RMIStatistics> read_stats() {
  std::vector<RMIStatistics> v;
  std::string line;
  while (std::getline(std::cin, line)) {
    if (line.empty()) continue;
    std::istringstream ss(line);
    RMIStatistics s;
    ss >> s.models >> s.branching_factor >> s.average_log2_error >> s.max_log2_error >> s.size;
    v.push_back(s);
  }
  return v;
}

int main(int argc, char** argv) {
  std::string mode = argc > 1 ? argv[1] : "";
  try {
    if (mode == "first") {
      for (auto& c : first_phase_configs()) std::cout << c.first << " " << c.second << "\n";
    } else if (mode == "second") {
      for (auto& c : second_phase_configs(read_stats())) std::cout << c.first << " " << c.second << "\n";
    } else if (mode == "front") {
      auto front = narrow_front(pareto_front(read_stats()), (size_t)std::stoul(argv[2]));
      std::stable_sort(front.begin(), front.end(),
                       [](const RMIStatistics& a, const RMIStatistics& b) { return a.average_log2_error < b.average_log2_error; });
      for (auto& s : front) std::cout << s.models << " " << s.branching_factor << " " << s.size << "\n";
    } else return 2;
  } catch (std::exception& e) { std::cerr << e.what() << "\n"; return 1; }
  return 0;
}
