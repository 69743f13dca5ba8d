// OMPL API stand-in (oracle/standins/README.md)
#pragma once
#include <memory>
#include "ompl/base/SpaceInformation.h"
namespace ompl {
namespace base {
class Path {
public:
  explicit Path(const SpaceInformationPtr& si) : si_(si) {}
  virtual ~Path() = default;
  template <class T> T* as() { return static_cast<T*>(this); }
protected:
  SpaceInformationPtr si_;
};
typedef std::shared_ptr<Path> PathPtr;
}  // namespace base
}  // namespace ompl
