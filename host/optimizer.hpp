// optimizer.hpp — the reference's two-phase configuration search (rmi_lib/src/optimizer.rs),
// CPU-side: the search loop stays on the host, every candidate is one rmi_train call with
// RMI_FLAG_STATS_ONLY (only avg/max log2 error and the model size are consumed,
// optimizer.rs:163-171), so no leaf table ever leaves the GPU during a sweep.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../include/rmi_b200.h"
#include "codegen.hpp"

namespace rmihost {

struct RMIStatistics {   // optimizer.rs:153-160
  std::string models;
  uint64_t branching_factor = 0;
  double average_log2_error = 0, max_log2_error = 0;
  uint64_t size = 0;

  bool dominated_by(const RMIStatistics& o) const {   // :173-187
    if (size < o.size) return false;
    if (average_log2_error < o.average_log2_error) return false;
    if (size == o.size && average_log2_error <= o.average_log2_error) return false;
    double d = std::fabs(average_log2_error - o.average_log2_error);
    if (size <= o.size && d < 2.220446049250313e-16) return false;
    return true;
  }
  bool has_config(const std::string& m, uint64_t bf) const { return models == m && branching_factor == bf; }
};

inline std::string optimizer_profile() {
  const char* p = std::getenv("RMI_OPTIMIZER_PROFILE");
  std::string s = p ? p : "";
  if (!s.empty() && s != "fast" && s != "memory" && s != "disk") throw std::runtime_error("Invalid optimizer profile " + s);
  return s;
}
inline std::vector<std::string> top_only_layers() {   // :15-28
  std::string p = optimizer_profile();
  if (p == "fast") return {"robust_linear"};
  if (p == "disk") return {"radix", "radix18", "radix22", "robust_linear", "normal", "lognormal", "loglinear"};
  return {"radix", "radix18", "radix22", "robust_linear"};
}
inline std::vector<std::string> anywhere_layers() {   // :30-41
  if (optimizer_profile() == "fast") return {"linear", "cubic"};
  return {"linear", "cubic", "linear_spline"};
}
inline std::vector<uint64_t> branching_factors() {   // :43-57
  std::string p = optimizer_profile();
  int hi = p == "disk" ? 28 : 25, step = p == "fast" ? 2 : 1;
  std::vector<uint64_t> v;
  for (int i = 6; i < hi; i += step) v.push_back((uint64_t)1 << i);
  return v;
}

inline std::vector<RMIStatistics> pareto_front(const std::vector<RMIStatistics>& r) {   // :59-72
  std::vector<RMIStatistics> front;
  for (auto& x : r) {
    bool dominated = false;
    for (auto& v : r) if (x.dominated_by(v)) { dominated = true; break; }
    if (!dominated) front.push_back(x);
  }
  return front;
}

inline std::vector<RMIStatistics> narrow_front(const std::vector<RMIStatistics>& results, size_t desired) {   // :74-108
  if (desired < 2) throw std::runtime_error("assertion failed: desired_size >= 2");
  if (results.size() <= desired) return results;
  std::vector<RMIStatistics> tmp = results;
  std::stable_sort(tmp.begin(), tmp.end(), [](const RMIStatistics& a, const RMIStatistics& b) { return a.size < b.size; });
  RMIStatistics best = tmp.front();
  tmp.erase(tmp.begin());
  while (tmp.size() > desired - 1) {
    size_t gi = 0;
    double gv = 0;
    bool have = false;
    for (size_t i = 0; i + 1 < tmp.size(); ++i) {   // min_by keeps the first minimum
      double v = (double)tmp[i + 1].size / (double)tmp[i].size;
      if (!have || v < gv) { gv = v; gi = i; have = true; }
    }
    double e1 = tmp[gi].average_log2_error, e2 = tmp[gi + 1].average_log2_error;
    if (e1 > e2) tmp.erase(tmp.begin() + gi); else tmp.erase(tmp.begin() + gi + 1);
  }
  tmp.insert(tmp.begin(), best);
  return tmp;
}

typedef std::pair<std::string, uint64_t> Config;

inline std::vector<Config> first_phase_configs() {   // :110-125
  std::vector<Config> out;
  std::vector<std::string> tops = top_only_layers();
  for (auto& a : anywhere_layers()) tops.push_back(a);
  auto bfs = branching_factors();
  for (auto& t : tops)
    for (auto& b : anywhere_layers())
      for (size_t i = 0; i < bfs.size(); i += 5) out.push_back({t + "," + b, bfs[i]});
  return out;
}
inline std::vector<Config> second_phase_configs(const std::vector<RMIStatistics>& first) {   // :127-151
  std::set<std::string> qualifying;   // BTreeSet: sorted iteration
  for (auto& r : pareto_front(first)) qualifying.insert(r.models);
  std::vector<Config> out;
  for (auto& m : qualifying)
    for (uint64_t bf : branching_factors()) {
      bool seen = false;
      for (auto& v : first) if (v.has_config(m, bf)) { seen = true; break; }
      if (!seen) out.push_back({m, bf});
    }
  return out;
}

// measure_rmis (:220-231).  The reference maps the configurations over a rayon pool on one shared
// data set (optimizer.rs:224-229); here `replicas` holds the SAME key set on one or more devices
// (rmi_dataset_replicate) and one host thread per replica pulls the next configuration from a
// shared counter — independent builds, no communication, results in configuration order.
// A panicking configuration aborts the sweep, as in the reference.
inline std::vector<RMIStatistics> measure_rmis(const std::vector<const rmi_dataset*>& replicas, const std::vector<Config>& configs,
                                               uint32_t flags, bool verbose) {
  std::vector<RMIStatistics> out(configs.size());
  // Unit of work = the configurations that share (top model, branching factor): rmi_train_stats_batch fits the top
  // model and derives the leaf boundaries once for the whole group (several configurations
  // per key pass).  Groups keep the order of their first member; results land at the configurations' own positions.
  struct Group { std::string top; uint64_t bf; std::vector<std::string> leaves; std::vector<size_t> index; bool batched; };
  std::vector<Group> groups;
  static const bool batching = [] { const char* e = std::getenv("RMI_OPTIMIZER_NO_BATCH"); return !(e && e[0] == '1'); }();
  for (size_t i = 0; i < configs.size(); ++i) {
    const std::string& m = configs[i].first;
    const size_t comma = m.find(',');
    const std::string top = m.substr(0, comma), leaf = comma == std::string::npos ? "" : m.substr(comma + 1);
    // Batching pays where the shared passes (top fit, boundary search) are a visible part of a configuration's cost:
    // at small branching factors a configuration's time is its leaves' serial recurrences (n / bf keys per lane), and
    // keeping such configurations apart lets the replicas balance them.
    const bool batchable = batching && comma != std::string::npos && leaf.find(',') == std::string::npos &&
                           configs[i].second >= 4096;
    Group* g = nullptr;
    if (batchable)
      for (auto& c : groups) if (c.batched && c.top == top && c.bf == configs[i].second) { g = &c; break; }
    if (!g) {
      groups.push_back(Group{batchable ? top : m, configs[i].second, {}, {}, batchable});
      g = &groups.back();
    }
    g->leaves.push_back(leaf);
    g->index.push_back(i);
  }
  // longest first: the smaller the branching factor, the longer the per-lane chains (results keep their own positions)
  std::vector<size_t> order(groups.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return groups[a].bf < groups[b].bf; });
  std::atomic<size_t> next{0};
  std::atomic<bool> failed{false};
  std::vector<std::string> errors(replicas.size());
  auto worker = [&](size_t w) {
    const rmi_dataset* ds = replicas[w];
    for (;;) {
      size_t oi = next.fetch_add(1);
      if (oi >= groups.size() || failed.load()) return;
      const Group& g = groups[order[oi]];
      const size_t K = g.index.size();
      std::vector<rmi_result*> res(K, nullptr);
      int rc;
      if (g.batched) {   // g.top holds the bare top-model name
        std::vector<const char*> names;
        for (auto& l : g.leaves) names.push_back(l.c_str());
        rc = rmi_train_stats_batch(ds, g.top.c_str(), names.data(), (int)K, g.bf, flags, res.data());
      } else {
        rc = rmi_train(ds, g.top.c_str(), g.bf, flags | RMI_FLAG_STATS_ONLY, &res[0]);
      }
      if (rc != RMI_OK) {
        errors[w] = "training " + configs[g.index[0]].first + " " + std::to_string(g.bf) + ": " + rmi_last_error();
        failed.store(true);
        return;
      }
      for (size_t k = 0; k < K; ++k) {
        const Config& c = configs[g.index[k]];
        rmi_result* r = res[k];
        RMIStatistics& s = out[g.index[k]];
        s.models = c.first; s.branching_factor = c.second;
        s.average_log2_error = r->model_avg_log2_error; s.max_log2_error = r->model_max_log2_error;
        s.size = rmi_size(*r, true);
        if (verbose) std::fprintf(stderr, "  [replica %zu] %-28s %10llu  avg_log2 %.5f  size %llu  (%.2f ms)\n", w, c.first.c_str(),
                                  (unsigned long long)c.second, s.average_log2_error, (unsigned long long)s.size, r->device_time_ns / 1e6);
        rmi_result_free(r);
      }
    }
  };
  if (replicas.size() == 1) worker(0);
  else {
    std::vector<std::thread> th;
    for (size_t w = 0; w < replicas.size(); ++w) th.emplace_back([&worker, w] { worker(w); rmi_thread_release(); });
    for (auto& t : th) t.join();
  }
  for (auto& e : errors) if (!e.empty()) throw std::runtime_error(e);
  return out;
}

inline std::vector<RMIStatistics> find_pareto_efficient_configs(const std::vector<const rmi_dataset*>& replicas, size_t restrict_to,
                                                                uint32_t flags, bool verbose) {   // :233-249
  auto first = measure_rmis(replicas, first_phase_configs(), flags, verbose);
  auto second = measure_rmis(replicas, second_phase_configs(first), flags, verbose);
  auto front = narrow_front(pareto_front(second), restrict_to);
  std::stable_sort(front.begin(), front.end(),
                   [](const RMIStatistics& a, const RMIStatistics& b) { return a.average_log2_error < b.average_log2_error; });
  return front;
}

inline std::vector<RMIStatistics> find_pareto_efficient_configs(const rmi_dataset* ds, size_t restrict_to, uint32_t flags,
                                                                bool verbose) {
  return find_pareto_efficient_configs(std::vector<const rmi_dataset*>{ds}, restrict_to, flags, verbose);
}

inline void display_table(const std::vector<RMIStatistics>& items) {   // :193-206
  std::vector<std::vector<std::string>> rows;
  rows.push_back({"Models", "Branch", "   AvgLg2", "   MaxLg2", "   Size (b)"});
  char buf[64];
  for (auto& it : items) {
    std::vector<std::string> r;
    r.push_back(it.models);
    std::snprintf(buf, sizeof buf, "%10llu", (unsigned long long)it.branching_factor); r.push_back(buf);
    std::snprintf(buf, sizeof buf, "     %.5f", it.average_log2_error); r.push_back(buf);
    std::snprintf(buf, sizeof buf, "     %.5f", it.max_log2_error); r.push_back(buf);
    r.push_back("     " + std::to_string(it.size));
    rows.push_back(r);
  }
  size_t w[5] = {0, 0, 0, 0, 0};
  for (auto& r : rows) for (int c = 0; c < 5; ++c) w[c] = std::max(w[c], r[c].size());
  for (auto& r : rows) {
    std::string line = r[0] + std::string(w[0] - r[0].size(), ' ');
    for (int c = 1; c < 5; ++c) line += " " + std::string(w[c] - r[c].size(), ' ') + r[c];
    std::printf("%s\n", line.c_str());
  }
}

}  // namespace rmihost
