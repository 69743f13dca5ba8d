// OMPL API stand-in (oracle/standins/README.md).  Exact nearest-neighbour queries by linear scan; results
// ascending by distance, equal distances in insertion order.
#pragma once
#include <algorithm>
#include <stdexcept>
#include "ompl/datastructures/NearestNeighbors.h"
namespace ompl {
template <typename _T>
class NearestNeighborsLinear : public NearestNeighbors<_T> {
public:
  void clear() override { data_.clear(); }
  void add(const _T& data) override { data_.push_back(data); }
  void add(const std::vector<_T>& data) override { data_.insert(data_.end(), data.begin(), data.end()); }
  bool remove(const _T& data) override {
    for (std::size_t i = 0; i < data_.size(); ++i)
      if (data_[i] == data) {
        data_.erase(data_.begin() + (std::ptrdiff_t)i);
        return true;
      }
    return false;
  }
  _T nearest(const _T& data) const override {
    std::vector<_T> nbh;
    nearestK(data, 1, nbh);
    if (nbh.empty()) throw std::runtime_error("No elements found in nearest neighbors data structure");
    return nbh[0];
  }
  void nearestK(const _T& data, std::size_t k, std::vector<_T>& nbh) const override {
    std::vector<std::pair<double, std::size_t>> d = scan(data);
    if (d.size() > k) d.resize(k);
    nbh.clear();
    for (auto& p : d) nbh.push_back(data_[p.second]);
  }
  void nearestR(const _T& data, double radius, std::vector<_T>& nbh) const override {
    std::vector<std::pair<double, std::size_t>> d = scan(data);
    nbh.clear();
    for (auto& p : d)
      if (p.first <= radius) nbh.push_back(data_[p.second]);
  }
  std::size_t size() const override { return data_.size(); }
  void list(std::vector<_T>& data) const override { data = data_; }
protected:
  std::vector<std::pair<double, std::size_t>> scan(const _T& q) const {
    std::vector<std::pair<double, std::size_t>> d(data_.size());
    for (std::size_t i = 0; i < data_.size(); ++i) d[i] = std::make_pair(this->distFun_(q, data_[i]), i);
    std::stable_sort(d.begin(), d.end(),
                     [](const std::pair<double, std::size_t>& a, const std::pair<double, std::size_t>& b) { return a.first < b.first; });
    return d;
  }
  std::vector<_T> data_;
};
}  // namespace ompl
