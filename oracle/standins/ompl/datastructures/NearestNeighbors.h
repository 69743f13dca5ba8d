// OMPL API stand-in (oracle/standins/README.md)
#pragma once
#include <functional>
#include <vector>
namespace ompl {
template <typename _T>
class NearestNeighbors {
public:
  typedef std::function<double(const _T&, const _T&)> DistanceFunction;
  virtual ~NearestNeighbors() = default;
  virtual void setDistanceFunction(const DistanceFunction& f) { distFun_ = f; }
  const DistanceFunction& getDistanceFunction() const { return distFun_; }
  virtual void clear() = 0;
  virtual void add(const _T& data) = 0;
  virtual void add(const std::vector<_T>& data) {
    for (const auto& e : data) add(e);
  }
  virtual bool remove(const _T& data) = 0;
  virtual _T nearest(const _T& data) const = 0;
  virtual void nearestK(const _T& data, std::size_t k, std::vector<_T>& nbh) const = 0;
  virtual void nearestR(const _T& data, double radius, std::vector<_T>& nbh) const = 0;
  virtual std::size_t size() const = 0;
  virtual void list(std::vector<_T>& data) const = 0;
protected:
  DistanceFunction distFun_;
};
}  // namespace ompl
