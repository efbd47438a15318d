// Test tool for host/param_grid.hpp (no GPU).
//   param_grid_tool parse <grid.json>             -> one line per entry: layers \t bf \t namespace-or-<none>
//   param_grid_tool results <num_rows> < lines    -> the <file>_results JSON for entries given as
//        "layers bf ns|- avg_error avg_l2 avg_log2 max_log2 max_error size" per line
#include <fstream>
#include <iostream>
#include <sstream>

#include "../../host/param_grid.hpp"

using namespace rmihost;

int main(int argc, char** argv) {
  std::string mode = argc > 1 ? argv[1] : "";
  try {
    if (mode == "parse" && argc > 2) {
      std::ifstream in(argv[2]);
      std::stringstream ss; ss << in.rdbuf();
      for (auto& e : parse_param_grid(ss.str()))
        std::cout << e.layers << "\t" << e.branching_factor << "\t" << (e.has_namespace ? e.ns : std::string("<none>")) << "\n";
    } else if (mode == "results" && argc > 2) {
      std::vector<GridResult> rs;
      std::string line;
      while (std::getline(std::cin, line)) {
        if (line.empty()) continue;
        std::istringstream ss(line);
        GridResult r; std::string ns;
        ss >> r.entry.layers >> r.entry.branching_factor >> ns >> r.avg_error >> r.avg_l2 >> r.avg_log2 >> r.max_log2 >> r.max_error >> r.size_bs;
        if (ns != "-") { r.entry.has_namespace = true; r.entry.ns = ns; }
        rs.push_back(r);
      }
      std::cout << grid_results_json(rs, std::strtoull(argv[2], nullptr, 10));
    } else return 2;
  } catch (std::exception& e) { std::cerr << e.what() << "\n"; return 1; }
  return 0;
}
