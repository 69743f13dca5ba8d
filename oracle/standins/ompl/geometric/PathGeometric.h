// OMPL API stand-in (oracle/standins/README.md)
#pragma once
#include <vector>
#include "ompl/base/Path.h"
namespace ompl {
namespace geometric {
class PathGeometric : public base::Path {
public:
  explicit PathGeometric(const base::SpaceInformationPtr& si) : base::Path(si) {}
  ~PathGeometric() override { for (auto* s : states_) si_->freeState(s); }
  void append(const base::State* s) {
    base::State* c = si_->allocState();
    si_->getStateSpace()->copyState(c, s);
    states_.push_back(c);
  }
  std::size_t getStateCount() const { return states_.size(); }
  const base::State* getState(std::size_t i) const { return states_[i]; }
  std::vector<base::State*>& getStates() { return states_; }
private:
  std::vector<base::State*> states_;
};
}  // namespace geometric
}  // namespace ompl
