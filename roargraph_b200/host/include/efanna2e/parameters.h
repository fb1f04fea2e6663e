// efanna2e::Parameters — string-keyed parameter bag with the reference's Set/Get contract
// (include/efanna2e/parameters.h:18-41 of the reference: Get on a missing key throws std::invalid_argument
// "Invalid parameter name."; values round-trip through a stringstream).
#pragma once
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>

namespace efanna2e {

class Parameters {
 public:
  template <typename T>
  void Set(const std::string &name, const T &value) {
    std::ostringstream os;
    os << value;
    kv_[name] = os.str();
  }

  template <typename T>
  T Get(const std::string &name) const {
    auto it = kv_.find(name);
    if (it == kv_.end()) throw std::invalid_argument("Invalid parameter name.");
    std::istringstream is(it->second);
    T v;
    if (!(is >> v) || !is.eof()) {
      throw std::runtime_error("Failed to convert value '" + it->second + "' to type: " + typeid(T).name());
    }
    return v;
  }

  template <typename T>
  T Get(const std::string &name, const T &fallback) const {
    return kv_.count(name) ? Get<T>(name) : fallback;
  }

  bool Has(const std::string &name) const { return kv_.count(name) != 0; }

 private:
  std::map<std::string, std::string> kv_;
};

}  // namespace efanna2e
