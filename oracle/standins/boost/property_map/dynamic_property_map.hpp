// Boost API stand-in (oracle/standins/README.md): dynamic_properties as a registry of named, string-valued
// vertex property maps (what read_graphml needs for the reference's "state" property).
#pragma once
#include <functional>
#include <map>
#include <string>
#include "boost/graph/properties.hpp"
namespace boost {
class dynamic_properties {
public:
  template <class PropertyMap> dynamic_properties& property(const std::string& name, PropertyMap pm) {
    setters_[name] = [pm](std::size_t key, const std::string& value) { put(pm, key, value); };
    return *this;
  }
  bool has(const std::string& name) const { return setters_.count(name) != 0; }
  void set(const std::string& name, std::size_t key, const std::string& value) const { setters_.at(name)(key, value); }
private:
  std::map<std::string, std::function<void(std::size_t, const std::string&)>> setters_;
};
}  // namespace boost
