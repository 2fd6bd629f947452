// Minimal stand-in for <yaml-cpp/yaml.h>, TEST INFRASTRUCTURE ONLY.
//
// The reference (tspooner/rl_markets) git-clones yaml-cpp at build time
// (ext/CMakeLists.txt:5-60), which cannot run here (no network).  Its sources
// only use a small slice of the yaml-cpp API:
//   YAML::LoadFile, Node::operator[](string|int), as<T>(), as<T>(fallback),
//   explicit operator bool, operator=.
// and config/example.yaml only contains block maps, scalars, comments and flow
// sequences.  This header implements exactly that slice so that the UNMODIFIED
// reference translation units compile into oracle/_ref.  None of the
// hot-path arithmetic lives in yaml-cpp, so the shim cannot change results.
#ifndef RLM_ORACLE_YAML_SHIM_H
#define RLM_ORACLE_YAML_SHIM_H

#include <cstdlib>
#include <fstream>
#include <list>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace YAML {

class Exception : public std::runtime_error {
 public:
  explicit Exception(const std::string& m) : std::runtime_error(m) {}
};
class BadConversion : public Exception {
 public:
  explicit BadConversion(const std::string& m) : Exception("yaml shim: bad conversion: " + m) {}
};
class BadFile : public Exception {
 public:
  explicit BadFile(const std::string& m) : Exception("yaml shim: cannot open " + m) {}
};

namespace detail {

struct Data {
  enum Kind { Undefined, Scalar, Map, Seq } kind = Undefined;
  std::string scalar;
  std::vector<std::pair<std::string, std::shared_ptr<Data>>> map;
  std::vector<std::shared_ptr<Data>> seq;

  std::shared_ptr<Data> find(const std::string& k) const {
    for (auto& kv : map)
      if (kv.first == k) return kv.second;
    return nullptr;
  }
};

inline std::string trim(const std::string& s) {
  size_t a = s.find_first_not_of(" \t\r\n");
  if (a == std::string::npos) return "";
  size_t b = s.find_last_not_of(" \t\r\n");
  return s.substr(a, b - a + 1);
}

inline std::string unquote(const std::string& s) {
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\'')))
    return s.substr(1, s.size() - 2);
  return s;
}

// Strip a trailing comment that is not inside quotes.
inline std::string strip_comment(const std::string& line) {
  bool in_s = false, in_d = false;
  for (size_t i = 0; i < line.size(); ++i) {
    char c = line[i];
    if (c == '"' && !in_s) in_d = !in_d;
    else if (c == '\'' && !in_d) in_s = !in_s;
    else if (c == '#' && !in_s && !in_d && (i == 0 || line[i - 1] == ' ' || line[i - 1] == '\t'))
      return line.substr(0, i);
  }
  return line;
}

inline std::shared_ptr<Data> parse_value(const std::string& raw) {
  auto d = std::make_shared<Data>();
  std::string v = trim(raw);
  if (!v.empty() && v.front() == '[') {
    if (v.back() != ']') throw Exception("yaml shim: unterminated flow sequence: " + v);
    d->kind = Data::Seq;
    std::string body = v.substr(1, v.size() - 2);
    std::string cur;
    bool in_s = false, in_d = false;
    auto flush = [&]() {
      std::string t = trim(cur);
      if (!t.empty()) {
        auto e = std::make_shared<Data>();
        e->kind = Data::Scalar;
        e->scalar = unquote(t);
        d->seq.push_back(e);
      }
      cur.clear();
    };
    for (char c : body) {
      if (c == '"' && !in_s) in_d = !in_d;
      if (c == '\'' && !in_d) in_s = !in_s;
      if (c == ',' && !in_s && !in_d) flush();
      else cur.push_back(c);
    }
    flush();
  } else {
    d->kind = Data::Scalar;
    d->scalar = unquote(v);
  }
  return d;
}

struct Line {
  int indent;
  std::string key;
  std::string value;  // empty => nested map follows
};

inline std::shared_ptr<Data> parse_block(const std::vector<Line>& lines, size_t& i, int indent) {
  auto d = std::make_shared<Data>();
  d->kind = Data::Map;
  while (i < lines.size() && lines[i].indent == indent) {
    const Line& ln = lines[i];
    ++i;
    if (!ln.value.empty()) {
      d->map.emplace_back(ln.key, parse_value(ln.value));
    } else if (i < lines.size() && lines[i].indent > indent) {
      d->map.emplace_back(ln.key, parse_block(lines, i, lines[i].indent));
    } else {
      auto e = std::make_shared<Data>();  // "key:" with nothing => null
      d->map.emplace_back(ln.key, e);
    }
  }
  if (i < lines.size() && lines[i].indent > indent)
    throw Exception("yaml shim: bad indentation near key " + lines[i].key);
  return d;
}

inline std::shared_ptr<Data> parse_stream(std::istream& in) {
  std::vector<Line> lines;
  std::string raw;
  while (std::getline(in, raw)) {
    std::string s = strip_comment(raw);
    if (trim(s).empty()) continue;
    if (trim(s) == "---") continue;
    int indent = 0;
    while (indent < (int)s.size() && s[indent] == ' ') ++indent;
    std::string body = trim(s);
    size_t colon = std::string::npos;
    bool in_s = false, in_d = false;
    for (size_t k = 0; k < body.size(); ++k) {
      char c = body[k];
      if (c == '"' && !in_s) in_d = !in_d;
      if (c == '\'' && !in_d) in_s = !in_s;
      if (c == ':' && !in_s && !in_d && (k + 1 == body.size() || body[k + 1] == ' ')) {
        colon = k;
        break;
      }
    }
    if (colon == std::string::npos) throw Exception("yaml shim: unsupported line: " + raw);
    Line ln;
    ln.indent = indent;
    ln.key = unquote(trim(body.substr(0, colon)));
    ln.value = trim(body.substr(colon + 1));
    lines.push_back(ln);
  }
  size_t i = 0;
  if (lines.empty()) {
    auto d = std::make_shared<Data>();
    d->kind = Data::Map;
    return d;
  }
  return parse_block(lines, i, lines[0].indent);
}

template <typename T>
struct conv;

template <>
struct conv<std::string> {
  static bool get(const Data& d, std::string& out) {
    if (d.kind != Data::Scalar) return false;
    out = d.scalar;
    return true;
  }
};
template <>
struct conv<bool> {
  static bool get(const Data& d, bool& out) {
    if (d.kind != Data::Scalar) return false;
    const std::string& s = d.scalar;
    if (s == "true" || s == "True" || s == "TRUE" || s == "yes" || s == "on") { out = true; return true; }
    if (s == "false" || s == "False" || s == "FALSE" || s == "no" || s == "off") { out = false; return true; }
    return false;
  }
};

template <typename T>
struct conv {
  // arithmetic scalars, parsed the way yaml-cpp does (stream extraction)
  static bool get(const Data& d, T& out) {
    static_assert(std::is_arithmetic<T>::value, "yaml shim: unsupported as<T>");
    if (d.kind != Data::Scalar) return false;
    std::istringstream ss(d.scalar);
    ss.unsetf(std::ios::dec);
    if (std::is_unsigned<T>::value && !d.scalar.empty() && d.scalar[0] == '-') return false;
    ss >> std::noskipws >> out;
    if (ss.fail()) return false;
    ss >> std::ws;
    return ss.eof();
  }
};

template <typename E>
struct conv<std::vector<E>> {
  static bool get(const Data& d, std::vector<E>& out) {
    if (d.kind != Data::Seq) return false;
    out.clear();
    for (auto& e : d.seq) {
      E v;
      if (!conv<E>::get(*e, v)) return false;
      out.push_back(v);
    }
    return true;
  }
};
template <typename E>
struct conv<std::list<E>> {
  static bool get(const Data& d, std::list<E>& out) {
    std::vector<E> v;
    if (!conv<std::vector<E>>::get(d, v)) return false;
    out.assign(v.begin(), v.end());
    return true;
  }
};

}  // namespace detail

class Node {
 public:
  Node() : d_(std::make_shared<detail::Data>()), valid_(true) {}

  // yaml-cpp: operator[] on a missing key yields a node that converts to false
  // and whose as<T>(fallback) returns the fallback.
  Node operator[](const std::string& key) const {
    if (d_ && d_->kind == detail::Data::Map) {
      auto c = d_->find(key);
      if (c) return Node(c, true, d_, key);
    }
    return Node(nullptr, false, d_, key);
  }
  Node operator[](const char* key) const { return (*this)[std::string(key)]; }
  Node operator[](int idx) const {
    if (d_ && d_->kind == detail::Data::Seq && idx >= 0 && (size_t)idx < d_->seq.size())
      return Node(d_->seq[idx], true, nullptr, "");
    return Node(nullptr, false, nullptr, "");
  }

  bool IsDefined() const { return valid_ && d_ && d_->kind != detail::Data::Undefined; }
  explicit operator bool() const { return IsDefined(); }
  bool operator!() const { return !IsDefined(); }

  size_t size() const {
    if (!d_) return 0;
    if (d_->kind == detail::Data::Seq) return d_->seq.size();
    if (d_->kind == detail::Data::Map) return d_->map.size();
    return 0;
  }

  template <typename T>
  T as() const {
    T out;
    if (!IsDefined() || !detail::conv<T>::get(*d_, out)) throw BadConversion(key_);
    return out;
  }
  template <typename T, typename S>
  T as(const S& fallback) const {
    T out;
    if (!IsDefined() || !detail::conv<T>::get(*d_, out)) return T(fallback);
    return out;
  }

  // c["a"]["b"] = value;  (src/main.cpp:342-347)
  template <typename T>
  Node& operator=(const T& v) {
    std::ostringstream ss;
    ss << v;
    assign_scalar(ss.str());
    return *this;
  }
  Node& operator=(const Node& o) = default;
  Node(const Node&) = default;

 private:
  friend Node LoadFile(const std::string&);
  friend Node Load(const std::string&);
  Node(std::shared_ptr<detail::Data> d, bool valid, std::shared_ptr<detail::Data> parent, std::string key)
      : d_(std::move(d)), valid_(valid), parent_(std::move(parent)), key_(std::move(key)) {}

  void assign_scalar(const std::string& s) {
    if (!d_) {
      d_ = std::make_shared<detail::Data>();
      if (parent_) {
        if (parent_->kind == detail::Data::Undefined) parent_->kind = detail::Data::Map;
        if (parent_->kind == detail::Data::Map) parent_->map.emplace_back(key_, d_);
      }
    }
    d_->kind = detail::Data::Scalar;
    d_->scalar = s;
    valid_ = true;
  }

  std::shared_ptr<detail::Data> d_;
  bool valid_;
  std::shared_ptr<detail::Data> parent_;
  std::string key_;
};

inline Node Load(const std::string& text) {
  std::istringstream ss(text);
  return Node(detail::parse_stream(ss), true, nullptr, "");
}

inline Node LoadFile(const std::string& path) {
  std::ifstream f(path);
  if (!f.is_open()) throw BadFile(path);
  return Node(detail::parse_stream(f), true, nullptr, "");
}

}  // namespace YAML

#endif
