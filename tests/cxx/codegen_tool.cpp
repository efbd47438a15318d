// Test tool: rebuild an rmi_result from a flat dump (written by tests/test_codegen.py from an
// oracle- or GPU-trained model) and run the product's code generator (host/codegen.hpp) on it.
//   codegen_tool <dump.bin> <namespace> <data_dir> <out_dir> <include_errors 0|1> <key_type>
//                [<spline.bin> <line_size> <num_data_rows>]      (a --bounded RMI, codegen.rs cache_fix)
//   codegen_tool cachefix <keyfile (reference format, uint64)> <line_size> <spline.bin>
//                runs host/cache_fix.hpp and writes the knots as (u64 key, u64 offset) pairs
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../host/codegen.hpp"

static uint64_t rd(std::ifstream& in) { uint64_t v = 0; in.read((char*)&v, 8); return v; }

static int run_cachefix(char** argv) {
  std::ifstream in(argv[2], std::ios::binary);
  uint64_t n = rd(in);
  std::vector<uint64_t> keys(n);
  in.read((char*)keys.data(), (std::streamsize)(n * 8));
  try {
    auto sp = rmihost::cache_fix(keys.data(), n, std::strtoull(argv[3], nullptr, 10));
    std::ofstream out(argv[4], std::ios::binary);
    for (auto& p : sp) { out.write((const char*)&p.first, 8); out.write((const char*)&p.second, 8); }
  } catch (std::exception& e) { std::fprintf(stderr, "cache_fix: %s\n", e.what()); return 1; }
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 5 && std::string(argv[1]) == "cachefix") return run_cachefix(argv);
  if (argc < 7) { std::fprintf(stderr, "usage\n"); return 2; }
  std::ifstream in(argv[1], std::ios::binary);
  rmi_result r{};
  r.num_rmi_rows = r.num_data_rows = rd(in);
  r.branching_factor = rd(in);
  r.l0_model_id = (uint32_t)rd(in);
  r.l0_bradix_high = (uint32_t)rd(in);
  r.l0_table_bits = (uint32_t)rd(in);
  r.l0_num_fparams = (uint32_t)rd(in);
  for (int q = 0; q < 4; ++q) { uint64_t b = rd(in); memcpy(&r.l0_fparams[q], &b, 8); }
  r.l0_num_iparams = (uint32_t)rd(in);
  for (int q = 0; q < 4; ++q) r.l0_iparams[q] = rd(in);
  std::vector<uint32_t> t32(rd(in));
  for (auto& v : t32) v = (uint32_t)rd(in);
  std::vector<uint64_t> a1(rd(in));
  for (auto& v : a1) v = rd(in);
  std::vector<uint64_t> a2(rd(in));
  for (auto& v : a2) v = rd(in);
  r.l0_table32_len = t32.size(); r.l0_table32 = t32.empty() ? nullptr : t32.data();
  r.l0_array1_len = a1.size(); r.l0_array1 = a1.empty() ? nullptr : a1.data();
  r.l0_array2_len = a2.size(); r.l0_array2 = a2.empty() ? nullptr : a2.data();
  r.l1_model_id = (uint32_t)rd(in);
  r.l1_params_per_model = (uint32_t)rd(in);
  std::vector<double> params(r.branching_factor * r.l1_params_per_model);
  for (auto& v : params) { uint64_t b = rd(in); memcpy(&v, &b, 8); }
  std::vector<uint64_t> errs(r.branching_factor);
  for (auto& v : errs) v = rd(in);
  r.l1_params = params.data();
  r.l1_errors = errs.data();
  std::vector<rmihost::SplinePoint> spline;
  rmihost::CacheFixInfo cf;
  if (argc >= 10) {
    std::ifstream sp(argv[7], std::ios::binary);
    for (;;) { uint64_t k = 0, v = 0; sp.read((char*)&k, 8); sp.read((char*)&v, 8); if (!sp) break; spline.emplace_back(k, v); }
    cf.line_size = std::strtoull(argv[8], nullptr, 10);
    cf.spline = &spline;
    cf.num_data_rows = std::strtoull(argv[9], nullptr, 10);
  }
  try {
    rmihost::output_rmi(argv[2], r, argv[3], std::atoi(argv[6]), std::atoi(argv[5]) != 0, 0, argv[4], argc >= 10 ? &cf : nullptr);
  } catch (std::exception& e) { std::fprintf(stderr, "codegen: %s\n", e.what()); return 1; }
  return 0;
}
