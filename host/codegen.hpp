// codegen.hpp — emits the reference's output artefacts from an rmi_result:
//   <ns>.cpp, <ns>.h, <ns>_data.h and <data_dir>/<ns>_L{i}_PARAMETERS
// CPU-side restatement of rmi_lib/src/codegen.rs (generate_code :450-754, output_rmi :757-788,
// LayerParams :24-315, rmi_size :375-394) and of the text / binary forms of ModelParam
// (models/mod.rs:565-651).  The C function bodies below are the reference models' code()
// strings (models/*.rs) and stdlib snippets (models/stdlib.rs): they are the FORMAT of the
// emitted artefact and therefore reproduced character for character.
//
// Where the reference iterates a HashSet (decls / sigs / needed_vars, codegen.rs:564-610,
// :634-650) its own output order varies from run to run; this writer uses first-insertion
// order, which is one of the orders the reference can produce.
#pragma once
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/rmi_b200.h"
#include "cache_fix.hpp"

namespace rmihost {

// TrainedRMI.cache_fix (train/mod.rs:31) of a `--bounded` build: the spline's knots, the line
// size, and the length of the ORIGINAL data set (train_bounded, train/mod.rs:175-176).
struct CacheFixInfo {
  uint64_t line_size = 0;
  const std::vector<SplinePoint>* spline = nullptr;
  uint64_t num_data_rows = 0;
};

// ---- ModelParam (models/mod.rs:509-674) ------------------------------------------------------
struct Param {
  enum Kind { Int, Float, IntArray, Int32Array } kind;
  uint64_t i = 0;
  double f = 0.0;
  const uint64_t* a64 = nullptr;
  const uint32_t* a32 = nullptr;
  size_t len = 1;

  static Param make_int(uint64_t v) { Param p; p.kind = Int; p.i = v; return p; }
  static Param make_float(double v) { Param p; p.kind = Float; p.f = v; return p; }
  static Param make_u64_array(const uint64_t* a, size_t n) { Param p; p.kind = IntArray; p.a64 = a; p.len = n; return p; }
  static Param make_u32_array(const uint32_t* a, size_t n) { Param p; p.kind = Int32Array; p.a32 = a; p.len = n; return p; }

  size_t size() const {   // :521-530
    switch (kind) { case Int: case Float: return 8; case IntArray: return 8 * len; default: return 4 * len; }
  }
  const char* c_type() const {   // :532-541
    switch (kind) { case Int: case IntArray: return "uint64_t"; case Float: return "double"; default: return "uint32_t"; }
  }
  bool is_array() const { return kind == IntArray || kind == Int32Array; }
  const char* c_type_mod() const { return is_array() ? "[]" : ""; }
  bool same_type(const Param& o) const { return kind == o.kind; }
};

// Rust `format!("{:.}", v)` for f64 = Display: shortest digits that round-trip, never an
// exponent; c_val() then appends ".0" when no '.' is present (models/mod.rs:568-574).
inline std::string rust_f64_display(double v) {
  if (v != v) return "NaN";
  if (v == 1.0 / 0.0) return "inf";
  if (v == -1.0 / 0.0) return "-inf";
  // shortest round-trip DIGITS (as Rust's Grisu/Ryu output), then laid out positionally with
  // zero padding — e.g. 4607535066590541824.0 prints as 4607535066590542000
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
  std::string sci(buf, r.ptr);
  std::string out;
  size_t pos = 0;
  if (sci[0] == '-') { out = "-"; pos = 1; }
  size_t epos = sci.find('e');
  std::string mant = sci.substr(pos, epos - pos);
  int exp10 = std::atoi(sci.c_str() + epos + 1);
  std::string digits;
  for (char c : mant) if (c != '.') digits += c;
  int len = (int)digits.size();
  if (exp10 >= len - 1) out += digits + std::string((size_t)(exp10 - (len - 1)), '0');
  else if (exp10 >= 0) out += digits.substr(0, (size_t)exp10 + 1) + "." + digits.substr((size_t)exp10 + 1);
  else out += "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
  return out;
}
inline std::string c_val(const Param& p) {   // models/mod.rs:565-596
  switch (p.kind) {
    case Param::Int: return std::to_string(p.i) + "UL";
    case Param::Float: {
      std::string s = rust_f64_display(p.f);
      if (s.find('.') == std::string::npos) s += ".0";
      return s;
    }
    case Param::IntArray: {
      std::string s = "{ ";
      for (size_t k = 0; k < p.len; ++k) { if (k) s += ", "; s += std::to_string(p.a64[k]) + "UL"; }
      return s + " }";
    }
    default: {
      std::string s = "{ ";
      for (size_t k = 0; k < p.len; ++k) { if (k) s += ", "; s += std::to_string(p.a32[k]) + "UL"; }
      return s + " }";
    }
  }
}
inline void write_param(std::ostream& out, const Param& p) {   // models/mod.rs:613-651 (little endian)
  switch (p.kind) {
    case Param::Int: out.write(reinterpret_cast<const char*>(&p.i), 8); break;
    case Param::Float: out.write(reinterpret_cast<const char*>(&p.f), 8); break;
    case Param::IntArray: out.write(reinterpret_cast<const char*>(p.a64), 8 * p.len); break;
    default: out.write(reinterpret_cast<const char*>(p.a32), 4 * p.len); break;
  }
}

// ---- what codegen needs to know about a model (Model trait, models/mod.rs:730-764) -----------
struct ModelInfo {
  std::string function_name, code;
  bool input_float = true, output_float = true, needs_bounds_check = true;
  std::vector<int> stdlib;   // 0 = EXP1, 1 = PHI, 2 = BinarySearch  (models/stdlib.rs)
  size_t params_per_model = 2;
};

inline const char* stdlib_decl(int f) {
  switch (f) {
    case 0: return "inline double exp1(double x);";
    case 1: return "inline double phi(double x);";
    default: return "uint64_t bs_lower_bound(const uint64_t a[], uint64_t n, uint64_t x);";
  }
}
inline const char* stdlib_code(int f) {
  switch (f) {
    case 0: return "\ninline double exp1(double x) {\n  x = 1.0 + x / 64.0;\n  x *= x; x *= x; x *= x; x *= x;\n  x *= x; x *= x;\n  return x;\n}\n";
    case 1: return "\ninline double phi(double x) {\n  return 1.0 / (1.0 + exp1(- 1.65451 * x));\n}\n";
    default:
      return "\nuint64_t bs_upper_bound(const uint64_t a[], uint64_t n, uint64_t x) {\n    int l = 0;\n    int h = n; // Not n - 1\n"
             "    while (l < h) {\n        int mid = (l + h) / 2;\n        if (x >= a[mid]) {\n            l = mid + 1;\n"
             "        } else {\n            h = mid;\n        }\n    }\n    return l;\n}\n\n";
  }
}

inline ModelInfo model_info(uint32_t id, bool bradix_high, unsigned prefix_bits, unsigned table_bits) {
  ModelInfo m;
  switch (id) {
    case RMI_MODEL_LINEAR: case RMI_MODEL_ROBUST_LINEAR: case RMI_MODEL_LINEAR_SPLINE:   // linear.rs:103-114, :280-291; linear_spline.rs:64-75
      m.function_name = "linear";
      m.code = "\ninline double linear(double alpha, double beta, double inp) {\n    return std::fma(beta, inp, alpha);\n}";
      break;
    case RMI_MODEL_CUBIC:   // cubic_spline.rs:169-186
      m.function_name = "cubic";
      m.code = "\ninline double cubic(double a, double b, double c, double d, double x) {\n    auto v1 = std::fma(a, x, b);\n"
               "    auto v2 = std::fma(v1, x, c);\n    auto v3 = std::fma(v2, x, d);\n    return v3;\n}";
      m.needs_bounds_check = false; m.params_per_model = 4;
      break;
    case RMI_MODEL_LOGLINEAR:   // linear.rs:193-209
      m.function_name = "loglinear";
      m.code = "\ninline double loglinear(double alpha, double beta, double inp) {\n    return exp1(std::fma(beta, inp, alpha));\n}";
      m.stdlib = {0};
      break;
    case RMI_MODEL_NORMAL:   // normal.rs:108-125
      m.function_name = "ncdf";
      m.code = "\ninline double ncdf(double mean, double stdev, double scale, double inp) {\n    return phi((inp - mean) / stdev) * scale;\n}";
      m.stdlib = {0, 1}; m.params_per_model = 3;
      break;
    case RMI_MODEL_LOGNORMAL:   // normal.rs:183-200
      m.function_name = "lncdf";
      m.code = "\ninline double lncdf(double mean, double stdev, double scale, double inp) {\n    return phi((fmax(0.0, log(inp)) - mean) / stdev) * scale;\n}";
      m.stdlib = {0, 1}; m.params_per_model = 3;
      break;
    case RMI_MODEL_RADIX:   // radix.rs:63-77
      m.function_name = "radix";
      m.code = "\ninline uint64_t radix(uint64_t prefix_length, uint64_t bits, uint64_t inp) {\n    return (inp << prefix_length) >> (64 - bits);\n}";
      m.input_float = m.output_float = false; m.needs_bounds_check = false;
      break;
    case RMI_MODEL_RADIX_TABLE: {   // radix.rs:147-166
      unsigned nb = (prefix_bits + table_bits > 64) ? 0 : 64 - (prefix_bits + table_bits);
      std::ostringstream o;
      o << "\ninline uint64_t radix_table(const uint32_t* table, const uint64_t inp) {\n    return table[((inp << " << prefix_bits
        << ") >> " << prefix_bits << ") >> " << nb << "];\n}";
      m.function_name = "radix_table"; m.code = o.str();
      m.input_float = m.output_float = false; m.needs_bounds_check = false; m.params_per_model = 1;
      break;
    }
    case RMI_MODEL_BRADIX:   // balanced_radix.rs:130-166
      if (bradix_high) {
        m.function_name = "bradix_clamp_high";
        m.code = "\ninline uint64_t bradix_clamp_high(uint64_t prefix_length, \n                                  uint64_t bits, uint64_t clamp, uint64_t inp) {\n"
                 "    uint64_t tmp = (inp << prefix_length) >> (64 - bits);\n    return (tmp > clamp ? clamp : tmp);\n    \n}\n";
      } else {
        m.function_name = "bradix_clamp_low";
        m.code = "\ninline uint64_t bradix_clamp_low(uint64_t prefix_length,\n                                 uint64_t bits, uint64_t clamp, uint64_t inp) {\n"
                 "    uint64_t tmp = (inp << prefix_length) >> (64 - bits);\n    return (tmp < clamp ? 0 : tmp - clamp);\n}\n";
      }
      m.input_float = m.output_float = false; m.needs_bounds_check = false; m.params_per_model = 3;
      break;
    case RMI_MODEL_HISTOGRAM:   // histogram.rs:80-103
      m.function_name = "ed_histogram";
      m.code = "\ninline uint64_t ed_histogram(const uint64_t length,\n                             const uint64_t radix[], \n"
               "                             const uint64_t pivots[], \n                             uint64_t key) {\n"
               "    uint64_t key_radix = key >> (64 - 20);\n    unsigned int radix_lb = radix[key_radix];\n"
               "    unsigned int radix_ub = radix[key_radix+1];\n"
               "    uint64_t li = bs_upper_bound(pivots + radix_lb, radix_ub - radix_lb, key) + radix_lb - 1;\n    return li;\n}\n";
      m.input_float = m.output_float = false; m.needs_bounds_check = false; m.stdlib = {2}; m.params_per_model = 3;
      break;
    default: throw std::runtime_error("unknown model id");
  }
  return m;
}

// ---- LayerParams (codegen.rs:24-315) ----------------------------------------------------------
struct LayerParams {
  enum Kind { Constant, Array, MixedArray } kind;
  size_t idx = 0, ppm = 0;
  std::vector<Param> params;

  static LayerParams make(size_t idx, bool array_access, size_t ppm, std::vector<Param> params) {   // :44-62
    LayerParams lp; lp.idx = idx; lp.ppm = ppm;
    bool mixed = false;
    for (auto& p : params) if (!params[0].same_type(p)) mixed = true;
    size_t bytes = 0;
    for (auto& p : params) bytes += p.size();
    if (mixed) lp.kind = MixedArray;
    else if (array_access || bytes > 4096) lp.kind = Array;
    else { lp.kind = Constant; lp.ppm = params.size(); }
    lp.params = std::move(params);
    return lp;
  }
  std::string array_name() const { return "L" + std::to_string(idx) + "_PARAMETERS"; }
  std::string constant_name(size_t p) const { return "L" + std::to_string(idx) + "_PARAMETER" + std::to_string(p); }
  size_t size() const { size_t b = 0; for (auto& p : params) b += p.size(); return b; }
  bool requires_malloc() const {   // :104-113
    if (kind == Array) return size() >= 4 * 1024;
    return kind == MixedArray;
  }
  const char* pointer_type() const { return kind == Array ? params[0].c_type() : "char"; }
  size_t params_per_model() const { return kind == Constant ? params.size() : ppm; }

  void to_code(std::ostream& t) const {   // :64-102
    if (kind == Constant) {
      for (size_t p = 0; p < params.size(); ++p)
        t << "const " << params[p].c_type() << " " << constant_name(p) << params[p].c_type_mod() << " = " << c_val(params[p]) << ";\n";
    } else if (kind == Array) {
      t << "const " << params[0].c_type() << " " << array_name() << "[] = {";
      for (size_t p = 0; p + 1 < params.size(); ++p) t << c_val(params[p]) << ",";
      t << c_val(params.back()) << "};\n";
    } else throw std::runtime_error("Cannot hardcode mixed array.");
  }
  void to_decl(std::ostream& t) const {   // :124-160
    if (kind == Array) {
      if (!requires_malloc()) {
        size_t items = 0;
        for (auto& p : params) items += p.len;
        t << params[0].c_type() << " " << array_name() << "[" << items << "];\n";
      } else t << params[0].c_type() << "* " << array_name() << ";\n";
    } else if (kind == MixedArray) t << "char* " << array_name() << ";\n";
    else throw std::runtime_error("Cannot forward-declare constants");
  }
  void write_to(std::ostream& t) const { for (auto& p : params) write_param(t, p); }   // :163-182
  std::string access_by_ref(const std::string& model_index, size_t pidx) const {   // :228-286
    if (params[0].is_array()) return array_name();
    if (kind == Array) return array_name() + "[" + std::to_string(ppm) + "*" + model_index + " + " + std::to_string(pidx) + "]";
    if (kind == MixedArray) {
      size_t bpm = 0, off = 0;
      for (size_t k = 0; k < ppm && k < params.size(); ++k) bpm += params[k].size();
      for (size_t k = 0; k < pidx && k < params.size(); ++k) off += params[k].size();
      return std::string("*((") + params[pidx].c_type() + "*) (" + array_name() + " + (" + model_index + " * " + std::to_string(bpm) + ") + " +
             std::to_string(off) + "))";
    }
    throw std::runtime_error("Cannot access constant parameters by reference");
  }
  std::string access_by_const(size_t pidx) const {   // :216-226
    if (kind == Constant) return constant_name(pidx);
    return access_by_ref("0", pidx);
  }
};

inline std::string model_index_from_output(bool from_float, uint64_t bound, bool needs_check) {   // :343-373
  std::string b = std::to_string(bound);
  if (from_float) return needs_check ? "FCLAMP(fpred, " + b + ".0 - 1.0)" : "(uint64_t) fpred";
  return needs_check ? "(ipred > " + b + " - 1 ? " + b + " - 1 : ipred)" : "ipred";
}

inline std::vector<Param> top_params(const rmi_result& r) {   // Model::params() of the top model
  std::vector<Param> p;
  switch (r.l0_model_id) {
    case RMI_MODEL_RADIX: p = {Param::make_int(r.l0_iparams[0]), Param::make_int(r.l0_iparams[1])}; break;
    case RMI_MODEL_BRADIX: p = {Param::make_int(r.l0_iparams[0]), Param::make_int(r.l0_iparams[1]), Param::make_int(r.l0_iparams[2])}; break;
    case RMI_MODEL_RADIX_TABLE: p = {Param::make_u32_array(r.l0_table32, r.l0_table32_len)}; break;
    case RMI_MODEL_HISTOGRAM:
      p = {Param::make_int(r.l0_array2_len), Param::make_u64_array(r.l0_array1, r.l0_array1_len),
           Param::make_u64_array(r.l0_array2, r.l0_array2_len)};
      break;
    default:
      for (uint32_t q = 0; q < r.l0_num_fparams; ++q) p.push_back(Param::make_float(r.l0_fparams[q]));
  }
  return p;
}

// codegen.rs:375-394
inline uint64_t rmi_size(const rmi_result& r, bool with_errors, const CacheFixInfo* cf = nullptr) {
  uint64_t total = 0;
  for (auto& p : top_params(r)) total += p.size();
  total += (uint64_t)r.l1_params_per_model * 8 * r.branching_factor;
  if (with_errors) total += r.branching_factor * 8;
  if (cf) total += (uint64_t)cf->spline->size() * 16;
  return total;
}

// codegen.rs:396-448 generate_cache_fix_code: the public lookup() of a `--bounded` RMI — the RMI
// finds the spline segment, the segment interpolates the position, rounded down to its line.
inline void generate_cache_fix_code(std::ostream& t, const CacheFixInfo& cf, const std::string& array_name) {
  const std::string ls = std::to_string(cf.line_size);
  t << "\nstruct __attribute__((packed)) SplinePoint {\n  uint64_t key;\n  uint64_t value;\n};\n\n"
       "uint64_t lookup(uint64_t key, size_t* err) {\n"
       "  const uint64_t num_spline_pts = " << cf.spline->size() << ";\n"
       "  const uint64_t total_keys = " << cf.num_data_rows << ";\n"
       "  size_t error_on_spline_search;\n\n"
       "  struct SplinePoint* begin = (struct SplinePoint*) " << array_name << ";\n\n"
       "  *err = " << ls << ";\n"
       "  uint64_t start = _rmi_lookup_pre_cachefix(key, &error_on_spline_search);\n\n"
       "  size_t upper = (start + error_on_spline_search > num_spline_pts\n"
       "                  ? num_spline_pts : start + error_on_spline_search);\n"
       "  size_t lower = (error_on_spline_search > start\n"
       "                  ? 0 : start - error_on_spline_search);\n"
       "                  \n"
       "  \n"
       "  struct SplinePoint* res = std::lower_bound(begin + lower,\n"
       "                                             begin + upper,\n"
       "                                             key,\n"
       "                                             [](const auto& lhs, const auto rhs) { return lhs.key < rhs; });\n\n"
       "  if (res == begin + num_spline_pts)\n"
       "    // we've searched for something past the last point\n"
       "    return total_keys - 1;\n\n"
       "  auto pt1 = *(res - 1);\n"
       "  auto pt2 = *res;\n\n"
       "  auto v0 = (double)pt1.value;\n"
       "  auto v1 = (double)pt2.value;\n"
       "  auto t = ((double)(key - pt1.key)) / (double)(pt2.key - pt1.key);\n"
       "  return (((uint64_t) std::fma(1.0 - t, v0, t * v1)) / " << ls << ") * " << ls << ";\n"
       "}\n";
}

struct KeyTypeInfo { const char* c_type; bool is_float; };
// src/main.rs:122-132: uint32 files keep KeyType::U64; f64 files use KeyType::F64
inline KeyTypeInfo key_type_info(int key_type) {
  return key_type == RMI_KEY_F64 ? KeyTypeInfo{"double", true} : KeyTypeInfo{"uint64_t", false};
}

// codegen.rs:757-788 output_rmi + :450-754 generate_code.  Files are written relative to `out_dir`.
inline void output_rmi(const std::string& ns, const rmi_result& r, const std::string& data_dir, int key_type,
                       bool include_errors, uint64_t build_time_ns, const std::string& out_dir = ".",
                       const CacheFixInfo* cache_fix = nullptr) {
  const uint64_t N = r.branching_factor;
  if (!r.l1_params || (include_errors && !r.l1_errors)) throw std::runtime_error("result was trained with STATS_ONLY");
  ModelInfo top = model_info(r.l0_model_id, r.l0_bradix_high != 0, (unsigned)r.l0_iparams[0], r.l0_table_bits);
  ModelInfo leaf = model_info(r.l1_model_id, true, 0, 0);
  const size_t ppm = r.l1_params_per_model;

  std::vector<LayerParams> layers;
  {
    auto tp = top_params(r);
    layers.push_back(LayerParams::make(0, false, tp.size(), tp));                       // params_for_layer :330-341
    std::vector<Param> lp;
    lp.reserve(N * ppm);
    for (uint64_t j = 0; j < N; ++j) for (size_t q = 0; q < ppm; ++q) lp.push_back(Param::make_float(r.l1_params[j * ppm + q]));
    layers.push_back(LayerParams::make(1, N > 1, ppm, lp));
  }
  const bool report_lle = include_errors;
  std::string report_line;
  if (report_lle) {
    if (N > 1) {   // with_zipped_errors :288-315
      LayerParams old = std::move(layers.back());
      layers.pop_back();
      std::vector<Param> z;
      z.reserve(N * (ppm + 1));
      size_t opp = old.params_per_model();
      for (uint64_t j = 0; j < N; ++j) {
        for (size_t q = 0; q < opp; ++q) z.push_back(old.params[j * opp + q]);
        z.push_back(Param::make_int(r.l1_errors[j]));
      }
      LayerParams nl = LayerParams::make(old.idx, old.kind == LayerParams::Constant, opp + 1, z);
      report_line = "  *err = " + nl.access_by_ref("modelIndex", nl.params_per_model() - 1) + ";\n";
      layers.push_back(std::move(nl));
    } else {
      report_line = "  *err = " + std::to_string(r.l1_errors[0]) + ";";
    }
  }

  if (cache_fix) {   // codegen.rs:487-496: the knots become one more parameter array, {key, offset} per knot
    std::vector<Param> cfv;
    cfv.reserve(cache_fix->spline->size() * 2);
    for (const auto& pt : *cache_fix->spline) { cfv.push_back(Param::make_int(pt.first)); cfv.push_back(Param::make_int(pt.second)); }
    layers.push_back(LayerParams::make(layers.size(), true, 2, cfv));
  }

  std::ofstream code(out_dir + "/" + ns + ".cpp"), data(out_dir + "/" + ns + "_data.h"), header(out_dir + "/" + ns + ".h");
  if (!code || !data || !header) throw std::runtime_error("Could not write RMI source files");

  data << "namespace " << ns << " {\n";
  std::vector<std::string> read_code{"bool load(char const* dataPath) {"};
  for (auto& lp : layers) {
    if (lp.kind == LayerParams::Constant) { lp.to_code(data); continue; }
    std::string path = data_dir + "/" + ns + "_" + lp.array_name();
    std::ofstream bw(path, std::ios::binary);
    if (!bw) throw std::runtime_error("Could not write data file to RMI directory");
    lp.write_to(bw);
    lp.to_decl(data);
    read_code.push_back("  {");
    read_code.push_back("    std::ifstream infile(std::filesystem::path(dataPath) / \"" + ns + "_" + lp.array_name() +
                        "\", std::ios::in | std::ios::binary);");
    read_code.push_back("    if (!infile.good()) return false;");
    if (lp.requires_malloc()) {
      read_code.push_back("    " + lp.array_name() + " = (" + lp.pointer_type() + "*) malloc(" + std::to_string(lp.size()) + ");");
      read_code.push_back("    if (" + lp.array_name() + " == NULL) return false;");
    }
    read_code.push_back("    infile.read((char*)" + lp.array_name() + ", " + std::to_string(lp.size()) + ");");
    read_code.push_back("    if (!infile.good()) return false;");
    read_code.push_back("  }");
  }
  read_code.push_back("  return true;");
  read_code.push_back("}");
  std::vector<std::string> free_code{"void cleanup() {"};
  for (auto& lp : layers) if (lp.requires_malloc()) free_code.push_back("    free(" + lp.array_name() + ");");
  free_code.push_back("}");
  data << "} // namespace\n";

  // stdlib declarations / definitions, then the model functions (sets: first-insertion order)
  std::vector<std::string> decls, sigs;
  auto add_unique = [](std::vector<std::string>& v, const std::string& s) { for (auto& e : v) if (e == s) return; v.push_back(s); };
  for (const ModelInfo* m : {&top, &leaf}) for (int f : m->stdlib) { add_unique(decls, stdlib_decl(f)); add_unique(sigs, stdlib_code(f)); }

  code << "#include \"" << ns << ".h\"\n" << "#include \"" << ns << "_data.h\"\n" << "#include <math.h>\n#include <cmath>\n#include <fstream>\n"
       << "#include <filesystem>\n#include <iostream>\n";
  if (cache_fix) code << "#include <algorithm>\n";   // :580-582
  code << "namespace " << ns << " {\n";
  for (auto& l : read_code) code << l << "\n";
  for (auto& l : free_code) code << l << "\n";
  for (auto& d : decls) code << d << "\n";
  for (auto& s : sigs) code << s << "\n";
  std::vector<std::string> msigs;
  add_unique(msigs, top.code);
  add_unique(msigs, leaf.code);
  for (auto& s : msigs) code << s << "\n";
  code << "\ninline size_t FCLAMP(double inp, double bound) {\n  if (inp < 0.0) return 0;\n  return (inp > bound ? bound : (size_t)inp);\n}\n\n";

  KeyTypeInfo kt = key_type_info(key_type);
  const std::string lookup_name = cache_fix ? "_rmi_lookup_pre_cachefix" : "lookup";   // :621-625
  std::string lookup_sig = report_lle ? "uint64_t " + lookup_name + "(" + kt.c_type + " key, size_t* err)"
                                      : "uint64_t " + lookup_name + "(" + kt.c_type + " key)";
  code << lookup_sig << " {\n";
  std::vector<std::string> vars;
  add_unique(vars, "size_t modelIndex;");
  for (const ModelInfo* m : {&top, &leaf}) add_unique(vars, m->output_float ? "double fpred;" : "uint64_t ipred;");
  for (auto& v : vars) code << "  " << v << "\n";

  // layer 0: single model, constant indexing (:664-677)
  code << "  " << (top.output_float ? "fpred" : "ipred") << " = " << top.function_name << "(";
  {
    size_t np = top_params(r).size();
    for (size_t p = 0; p < np; ++p) code << layers[0].access_by_const(p) << ", ";
  }
  code << "(" << (top.input_float ? "double" : "uint64_t") << ")key);\n";
  // layer 1 (:678-701)
  if (N > 1) {
    code << "  modelIndex = " << model_index_from_output(top.output_float, N, top.needs_bounds_check) << ";\n";
    code << "  " << (leaf.output_float ? "fpred" : "ipred") << " = " << leaf.function_name << "(";
    for (size_t p = 0; p < ppm; ++p) code << layers[1].access_by_ref("modelIndex", p) << ", ";
  } else {
    code << "  " << (leaf.output_float ? "fpred" : "ipred") << " = " << leaf.function_name << "(";
    for (size_t p = 0; p < ppm; ++p) code << layers[1].access_by_const(p) << ", ";
  }
  code << "(" << (leaf.input_float ? "double" : "uint64_t") << ")key);\n";
  code << report_line << "\n";
  code << "  return " << model_index_from_output(leaf.output_float, r.num_rmi_rows, true) << ";\n";
  code << "}\n";
  if (cache_fix) generate_cache_fix_code(code, *cache_fix, layers.back().array_name());   // :720-722
  code << "} // namespace\n";

  header << "#include <cstddef>\n#include <cstdint>\n";
  header << "namespace " << ns << " {\n";
  header << "bool load(char const* dataPath);\nvoid cleanup();\n";
  header << "const size_t RMI_SIZE = " << rmi_size(r, include_errors, cache_fix) << ";\n";
  header << "const uint64_t BUILD_TIME_NS = " << build_time_ns << ";\n";
  header << "const char NAME[] = \"" << ns << "\";\n";
  if (cache_fix) header << "uint64_t lookup(uint64_t key, size_t* err);\n";   // :746-750
  else header << lookup_sig << ";\n";
  header << "}\n";
}

}  // namespace rmihost
