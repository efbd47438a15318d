// param_grid.hpp — the `--param-grid <file>` input and `<file>_results` output of the reference CLI
// (src/main.rs:171-261): {"configs": [{"layers": "...", "branching factor": N, "namespace": "..."?}, ...]}
// in, {"results": [{...statistics...}, ...]} out.  A JSON reader just large enough for these files
// (objects, arrays, strings, numbers, bools, null) and the result writer; no GPU involved.
#pragma once
#include <charconv>
#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace rmihost {

struct JVal {
  enum T { Null, Bool, Num, Str, Arr, Obj } t = Null;
  bool b = false;
  double num = 0;
  std::string s;
  std::vector<JVal> a;
  std::vector<std::pair<std::string, JVal>> o;
  const JVal* get(const std::string& k) const {
    for (auto& kv : o) if (kv.first == k) return &kv.second;
    return nullptr;
  }
};

struct JParser {
  const std::string& src;
  size_t i = 0;
  explicit JParser(const std::string& s) : src(s) {}
  [[noreturn]] static void bad(const std::string& what) { throw std::runtime_error("param grid: " + what); }
  char peek() const { return i < src.size() ? src[i] : '\0'; }
  void ws() { while (i < src.size() && std::isspace((unsigned char)src[i])) ++i; }
  JVal parse_document() {
    JVal v = parse();
    ws();
    if (i != src.size()) bad("trailing characters after the JSON value");
    return v;
  }
  JVal parse() {
    ws();
    if (i >= src.size()) bad("unexpected end of JSON");
    JVal v;
    char c = src[i];
    if (c == '{') {
      v.t = JVal::Obj; ++i; ws();
      if (peek() == '}') { ++i; return v; }
      for (;;) {
        ws();
        JVal k = parse();
        if (k.t != JVal::Str) bad("object key must be a string");
        ws();
        if (peek() != ':') bad("':' expected");
        ++i;
        v.o.push_back({k.s, parse()});
        ws();
        if (peek() == ',') { ++i; continue; }
        if (peek() == '}') { ++i; break; }
        bad("',' or '}' expected");
      }
    } else if (c == '[') {
      v.t = JVal::Arr; ++i; ws();
      if (peek() == ']') { ++i; return v; }
      for (;;) {
        v.a.push_back(parse());
        ws();
        if (peek() == ',') { ++i; continue; }
        if (peek() == ']') { ++i; break; }
        bad("',' or ']' expected");
      }
    } else if (c == '"') {
      v.t = JVal::Str; ++i;
      for (;;) {
        if (i >= src.size()) bad("unterminated string");
        char ch = src[i++];
        if (ch == '"') break;
        if (ch == '\\') {
          if (i >= src.size()) bad("unterminated escape");
          char e = src[i++];
          switch (e) {
            case 'n': v.s += '\n'; break;
            case 't': v.s += '\t'; break;
            case 'r': v.s += '\r'; break;
            case 'b': v.s += '\b'; break;
            case 'f': v.s += '\f'; break;
            case 'u': {   // \uXXXX: Basic Latin only is needed for namespaces / model lists; others pass through as '?'
              if (i + 4 > src.size()) bad("bad \\u escape");
              unsigned code = (unsigned)std::strtoul(src.substr(i, 4).c_str(), nullptr, 16);
              i += 4;
              v.s += code < 0x80 ? (char)code : '?';
              break;
            }
            default: v.s += e;   // \" \\ \/
          }
        } else v.s += ch;
      }
    } else if (!src.compare(i, 4, "true")) { v.t = JVal::Bool; v.b = true; i += 4; }
    else if (!src.compare(i, 5, "false")) { v.t = JVal::Bool; i += 5; }
    else if (!src.compare(i, 4, "null")) { i += 4; }
    else {
      v.t = JVal::Num;
      const char* b0 = src.c_str() + i;
      char* e = nullptr;
      v.num = std::strtod(b0, &e);
      if (e == b0) bad(std::string("unexpected character '") + c + "'");
      i += (size_t)(e - b0);
    }
    return v;
  }
};

inline std::string json_num(double v) {   // shortest round-trip form; non-finite values have no JSON spelling (json crate: null)
  if (!(v == v) || v == 1.0 / 0.0 || v == -1.0 / 0.0) return "null";
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof buf, v);
  return std::string(buf, r.ptr);
}
inline std::string json_str(const std::string& s) {
  std::string o = "\"";
  for (char c : s) {
    if (c == '"' || c == '\\') { o += '\\'; o += c; }
    else if (c == '\n') o += "\\n";
    else if (c == '\t') o += "\\t";
    else o += c;
  }
  return o + "\"";
}

struct GridEntry {   // main.rs:181-191
  std::string layers;
  uint64_t branching_factor = 0;
  bool has_namespace = false;
  std::string ns;
};

// main.rs:172-192, :257-259: "configs" must be an array; "layers" a string, "branching factor" an
// unsigned integer (as_u64), "namespace" optional.
inline std::vector<GridEntry> parse_param_grid(const std::string& text) {
  JVal root = JParser(text).parse_document();
  const JVal* configs = root.get("configs");
  if (!configs || configs->t != JVal::Arr) throw std::runtime_error("Configs must have an array as its value");
  std::vector<GridEntry> out;
  for (auto& el : configs->a) {
    const JVal* l = el.get("layers");
    const JVal* b = el.get("branching factor");
    const JVal* n = el.get("namespace");
    if (!l || l->t != JVal::Str) throw std::runtime_error("called `Option::unwrap()` on a `None` value (param grid entry: layers)");
    if (!b || b->t != JVal::Num || !(b->num >= 0.0) || b->num != (double)(uint64_t)b->num)
      throw std::runtime_error("called `Option::unwrap()` on a `None` value (param grid entry: branching factor)");
    GridEntry e;
    e.layers = l->s;
    e.branching_factor = (uint64_t)b->num;
    if (n && n->t == JVal::Str) { e.has_namespace = true; e.ns = n->s; }
    out.push_back(e);
  }
  return out;
}

struct GridResult {   // main.rs:205-220
  GridEntry entry;
  double avg_error = 0, avg_l2 = 0, avg_log2 = 0, max_log2 = 0;
  uint64_t max_error = 0, size_bs = 0;
};

// NB "average error %" is computed from the MAX error in the reference (main.rs:210-211).
inline std::string grid_results_json(const std::vector<GridResult>& rs, uint64_t num_rows) {
  std::string out = "{\"results\":[";
  for (size_t i = 0; i < rs.size(); ++i) {
    const GridResult& r = rs[i];
    const double pct = (double)r.max_error / (double)num_rows * 100.0;
    if (i) out += ",";
    out += "{\"layers\":" + json_str(r.entry.layers) + ",\"branching factor\":" + std::to_string(r.entry.branching_factor) +
           ",\"average error\":" + json_num(r.avg_error) + ",\"average error %\":" + json_num(pct) + ",\"average l2 error\":" +
           json_num(r.avg_l2) + ",\"average log2 error\":" + json_num(r.avg_log2) + ",\"max error\":" + std::to_string(r.max_error) +
           ",\"max error %\":" + json_num(pct) + ",\"max log2 error\":" + json_num(r.max_log2) + ",\"size binary search\":" +
           std::to_string(r.size_bs) + ",\"namespace\":" + (r.entry.has_namespace ? json_str(r.entry.ns) : std::string("null")) + "}";
  }
  return out + "]}";
}

}  // namespace rmihost
