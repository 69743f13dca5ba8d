// OMPL API stand-in (tests/stubs/README.md)
#pragma once
namespace ompl {
namespace base {
class State {
public:
  template <class T> const T* as() const { return static_cast<const T*>(this); }
  template <class T> T* as() { return static_cast<T*>(this); }
protected:
  State() = default;
  virtual ~State() = default;
};
}  // namespace base
}  // namespace ompl
