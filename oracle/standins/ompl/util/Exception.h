// OMPL API stand-in (oracle/standins/README.md)
#pragma once
#include <stdexcept>
#include <string>
namespace ompl {
class Exception : public std::runtime_error {
public:
  explicit Exception(const std::string& what) : std::runtime_error(what) {}
  Exception(const std::string& prefix, const std::string& what) : std::runtime_error(prefix + ": " + what) {}
};
}  // namespace ompl
