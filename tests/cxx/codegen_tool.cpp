// Test tool: rebuild an rmi_result from a flat dump (written by tests/test_codegen.py from an
// oracle- or GPU-trained model) and run the product's code generator (host/codegen.hpp) on it.
//   codegen_tool <dump.bin> <namespace> <data_dir> <out_dir> <include_errors 0|1> <key_type>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "../../host/codegen.hpp"

static uint64_t rd(std::ifstream& in) { uint64_t v = 0; in.read((char*)&v, 8); return v; }

int main(int argc, char** argv) {
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
  try {
    rmihost::output_rmi(argv[2], r, argv[3], std::atoi(argv[6]), std::atoi(argv[5]) != 0, 0, argv[4]);
  } catch (std::exception& e) { std::fprintf(stderr, "codegen: %s\n", e.what()); return 1; }
  return 0;
}
