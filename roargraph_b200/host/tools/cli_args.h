// Minimal command-line parser covering what the reference's CLIs rely on from boost::program_options:
// "--name value", "--name v1 v2 v3" (multitoken), "-T value" short aliases, required options (missing -> message on
// stderr and exit code -1, tests/test_search_roargraph.cpp:114-117), defaults, and --help/-h printing the option table.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace cli {

struct Option {
  std::string name, help;
  char short_name = 0;
  bool required = false, multitoken = false, seen = false, has_default = false, numeric = false;
  std::vector<std::string> values;
};

class Parser {
 public:
  explicit Parser(std::string title) : title_(std::move(title)) {}

  Parser &add(const std::string &name, const std::string &help, bool required, bool multitoken = false,
              const std::string &dflt = "", char short_name = 0) {
    Option o;
    o.name = name;
    o.help = help;
    o.required = required;
    o.multitoken = multitoken;
    o.short_name = short_name;
    if (!dflt.empty()) {
      o.values = {dflt};
      o.has_default = true;
    }
    opts_.push_back(o);
    return *this;
  }

  // marks the option added last as an unsigned integer: parse() then rejects other values, where
  // boost::program_options rejects them (inside the CLI's try block: message on stderr, exit code -1)
  Parser &numeric() {
    if (!opts_.empty()) opts_.back().numeric = true;
    return *this;
  }

  // returns 0 to continue, 1 when --help was printed; throws std::runtime_error on bad input
  int parse(int argc, char **argv) {
    Option *cur = nullptr;
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i];
      Option *o = nullptr;
      if (a == "--help" || a == "-h") {
        print(std::cout);
        return 1;
      }
      if (a.rfind("--", 0) == 0) o = find(a.substr(2));
      else if (a.size() == 2 && a[0] == '-' && !(a[1] >= '0' && a[1] <= '9')) o = find_short(a[1]);
      if (a.rfind("--", 0) == 0 && !o) throw std::runtime_error("unrecognised option '" + a + "'");
      if (o) {
        cur = o;
        cur->seen = true;
        cur->values.clear();
        continue;
      }
      if (!cur) throw std::runtime_error("too many positional options have been specified on the command line");
      cur->values.push_back(a);
      if (!cur->multitoken) cur = nullptr;
    }
    for (auto &o : opts_) {
      if (o.seen && o.values.empty()) throw std::runtime_error("the required argument for option '--" + o.name + "' is missing");
      if (o.required && !o.seen && !o.has_default) throw std::runtime_error("the option '--" + o.name + "' is required but missing");
      if (o.numeric)
        for (auto &v : o.values) (void)to_u64(o.name, v);
    }
    return 0;
  }

  bool has(const std::string &name) const {
    for (auto &o : opts_)
      if (o.name == name) return o.seen || o.has_default;
    return false;
  }
  std::string str(const std::string &name) const {
    for (auto &o : opts_)
      if (o.name == name) return o.values.empty() ? std::string() : o.values[0];
    return std::string();
  }
  uint64_t u64(const std::string &name) const { return to_u64(name, str(name)); }
  std::vector<uint32_t> u32_list(const std::string &name) const {
    std::vector<uint32_t> v;
    for (auto &o : opts_)
      if (o.name == name)
        for (auto &s : o.values) v.push_back((uint32_t)to_u64(name, s));
    return v;
  }
  void print(std::ostream &os) const {
    os << title_ << ":\n";
    for (auto &o : opts_) {
      os << "  ";
      if (o.short_name) os << "-" << o.short_name << " [ --" << o.name << " ]";
      else os << "--" << o.name;
      if (o.has_default) os << " arg (=" << o.values[0] << ")";
      else os << " arg";
      os << "\n        " << o.help << "\n";
    }
  }

 private:
  static uint64_t to_u64(const std::string &name, const std::string &s) {
    char *end = nullptr;
    const unsigned long long v = strtoull(s.c_str(), &end, 10);
    if (s.empty() || *end) throw std::runtime_error("the argument ('" + s + "') for option '--" + name + "' is invalid");
    return v;
  }
  Option *find(const std::string &n) {
    for (auto &o : opts_)
      if (o.name == n) return &o;
    return nullptr;
  }
  Option *find_short(char c) {
    for (auto &o : opts_)
      if (o.short_name == c) return &o;
    return nullptr;
  }
  std::string title_;
  std::vector<Option> opts_;
};

}  // namespace cli
