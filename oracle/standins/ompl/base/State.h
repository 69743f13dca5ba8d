// OMPL API stand-in (oracle/standins/README.md)
#pragma once
namespace ompl {
namespace base {
class State {
public:
  template <class T> T* as() { return static_cast<T*>(this); }
  template <class T> const T* as() const { return static_cast<const T*>(this); }
protected:
  State() = default;
  ~State() = default;
};
}  // namespace base
}  // namespace ompl
