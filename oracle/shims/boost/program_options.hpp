// Minimal stand-in for <boost/program_options.hpp> (Boost is not installed in this image).
// TEST INFRASTRUCTURE ONLY: covers exactly the subset the reference CLIs use
// (tests/test_search_roargraph.cpp:79-117, tests/test_build_roargraph.cpp:34-80,
//  thirdparty/DiskANN/tests/utils/compute_groundtruth.cpp:462-498):
//   options_description{"caption"}; desc.add_options()(name[,value][,help]);
//   value<T>(&v)->required()/default_value(x)/multitoken(); "long,S" short aliases;
//   variables_map; store(parse_command_line(argc, argv, desc), vm); notify(vm); vm.count(name);
//   operator<<(ostream&, options_description).
#pragma once
#include <cstdint>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost {
namespace program_options {

struct value_base {
    bool is_required = false, has_default = false, is_multitoken = false;
    virtual ~value_base() {}
    virtual void parse(const std::vector<std::string> &toks) = 0;
    virtual void apply_default() = 0;
};

template <class T>
struct scalar_parser {
    static void run(T *dst, const std::vector<std::string> &toks) {
        if (toks.empty()) throw std::runtime_error("option requires a value");
        std::stringstream ss(toks.back());
        T v;
        if (!(ss >> v)) throw std::runtime_error("bad option value '" + toks.back() + "'");
        *dst = v;
    }
};
template <>
struct scalar_parser<std::string> {
    static void run(std::string *dst, const std::vector<std::string> &toks) {
        if (toks.empty()) throw std::runtime_error("option requires a value");
        *dst = toks.back();
    }
};
template <class E>
struct scalar_parser<std::vector<E>> {
    static void run(std::vector<E> *dst, const std::vector<std::string> &toks) {
        dst->clear();
        for (auto &t : toks) {
            E e;
            std::vector<std::string> one{t};
            scalar_parser<E>::run(&e, one);
            dst->push_back(e);
        }
    }
};

template <class T>
struct typed_value : value_base {
    T *dst;
    T dflt{};
    explicit typed_value(T *d) : dst(d) {}
    typed_value *required() { is_required = true; return this; }
    typed_value *multitoken() { is_multitoken = true; return this; }
    typed_value *default_value(const T &v) { dflt = v; has_default = true; return this; }
    void parse(const std::vector<std::string> &toks) override { scalar_parser<T>::run(dst, toks); }
    void apply_default() override { *dst = dflt; }
};

template <class T>
typed_value<T> *value(T *dst) { return new typed_value<T>(dst); }

struct option_entry {
    std::string long_name, short_name, help;
    std::shared_ptr<value_base> val;  // null => flag
};

class options_description;
class options_adder {
    options_description *owner_;
   public:
    explicit options_adder(options_description *o) : owner_(o) {}
    options_adder &operator()(const char *name, const char *help);
    options_adder &operator()(const char *name, value_base *v, const char *help = "");
};

class options_description {
   public:
    std::string caption;
    std::vector<option_entry> opts;
    options_description() {}
    options_description(const std::string &c) : caption(c) {}
    options_adder add_options() { return options_adder(this); }
    const option_entry *find(const std::string &tok) const {
        for (auto &o : opts) {
            if (tok == "--" + o.long_name) return &o;
            if (!o.short_name.empty() && tok == "-" + o.short_name) return &o;
        }
        return nullptr;
    }
};

inline void split_name(const char *name, option_entry &e) {
    std::string s(name);
    auto c = s.find(',');
    if (c == std::string::npos) e.long_name = s;
    else { e.long_name = s.substr(0, c); e.short_name = s.substr(c + 1); }
}
inline options_adder &options_adder::operator()(const char *name, const char *help) {
    option_entry e; split_name(name, e); e.help = help; owner_->opts.push_back(e); return *this;
}
inline options_adder &options_adder::operator()(const char *name, value_base *v, const char *help) {
    option_entry e; split_name(name, e); e.help = help; e.val.reset(v); owner_->opts.push_back(e); return *this;
}

inline std::ostream &operator<<(std::ostream &os, const options_description &d) {
    os << d.caption << ":\n";
    for (auto &o : d.opts) {
        os << "  ";
        if (!o.short_name.empty()) os << "-" << o.short_name << " [ --" << o.long_name << " ]";
        else os << "--" << o.long_name;
        if (o.val) os << " arg";
        os << "\t" << o.help << "\n";
    }
    return os;
}

struct parsed_options {
    const options_description *desc;
    std::map<std::string, std::vector<std::string>> seen;
};

inline parsed_options parse_command_line(int argc, const char *const *argv, const options_description &desc) {
    parsed_options p; p.desc = &desc;
    int i = 1;
    while (i < argc) {
        std::string tok(argv[i]);
        std::string inline_val; bool has_inline = false;
        auto eq = tok.find('=');
        if (tok.rfind("--", 0) == 0 && eq != std::string::npos) {
            inline_val = tok.substr(eq + 1); tok = tok.substr(0, eq); has_inline = true;
        }
        const option_entry *o = desc.find(tok);
        if (!o) throw std::runtime_error("unrecognised option '" + tok + "'");
        ++i;
        std::vector<std::string> &vals = p.seen[o->long_name];
        if (!o->val) continue;
        if (has_inline) { vals.push_back(inline_val); continue; }
        if (i >= argc) throw std::runtime_error("the required argument for option '--" + o->long_name + "' is missing");
        vals.push_back(argv[i++]);
        if (o->val->is_multitoken) {
            while (i < argc) {
                std::string nx(argv[i]);
                bool looks_opt = nx.size() > 1 && nx[0] == '-' && !(nx[1] >= '0' && nx[1] <= '9') && nx[1] != '.';
                if (looks_opt) break;
                vals.push_back(nx); ++i;
            }
        }
    }
    return p;
}

class variables_map {
   public:
    std::map<std::string, std::vector<std::string>> seen;
    const options_description *desc = nullptr;
    size_t count(const std::string &name) const { return seen.count(name); }
};

inline void store(const parsed_options &p, variables_map &vm) { vm.seen = p.seen; vm.desc = p.desc; }

inline void notify(variables_map &vm) {
    if (!vm.desc) return;
    for (auto &o : vm.desc->opts) {
        if (!o.val) continue;
        auto it = vm.seen.find(o.long_name);
        if (it != vm.seen.end()) o.val->parse(it->second);
        else if (o.val->has_default) o.val->apply_default();
        else if (o.val->is_required)
            throw std::runtime_error("the option '--" + o.long_name + "' is required but missing");
    }
}

}  // namespace program_options
}  // namespace boost
