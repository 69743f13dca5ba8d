// OMPL API stand-in (oracle/standins/README.md)
#pragma once
#include <functional>
#include <map>
#include <sstream>
#include <string>
#include "ompl/base/ProblemDefinition.h"
#include "ompl/base/ScopedState.h"
namespace ompl {
namespace base {
struct PlannerStatus {
  enum StatusType { UNKNOWN = 0, INVALID_START, INVALID_GOAL, UNRECOGNIZED_GOAL_TYPE, TIMEOUT, APPROXIMATE_SOLUTION,
                    EXACT_SOLUTION, CRASH, ABORT, TYPE_COUNT };
  PlannerStatus(StatusType s = UNKNOWN) : status_(s) {}
  operator StatusType() const { return status_; }
  StatusType status_;
};
class PlannerTerminationCondition {
public:
  explicit PlannerTerminationCondition(std::function<bool()> fn) : fn_(std::move(fn)) {}
  bool eval() const { return fn_(); }
  bool operator()() const { return eval(); }
  operator bool() const { return eval(); }
private:
  std::function<bool()> fn_;
};
class ParamSet {
public:
  struct Param {
    std::function<bool(const std::string&)> set;
    std::function<std::string()> get;
  };
  bool setParam(const std::string& key, const std::string& value) {
    auto it = params_.find(key);
    return it != params_.end() && it->second.set(value);
  }
  bool getParam(const std::string& key, std::string& value) const {
    auto it = params_.find(key);
    if (it == params_.end()) return false;
    value = it->second.get();
    return true;
  }
  bool hasParam(const std::string& key) const { return params_.count(key) != 0; }
  std::size_t size() const { return params_.size(); }
  std::map<std::string, Param> params_;
};
class Planner {
public:
  Planner(SpaceInformationPtr si, std::string name) : si_(std::move(si)), name_(std::move(name)) {}
  virtual ~Planner() = default;
  virtual void setup() { setup_ = true; }
  virtual void setProblemDefinition(const ProblemDefinitionPtr& pdef) { pdef_ = pdef; }
  virtual PlannerStatus solve(const PlannerTerminationCondition& ptc) = 0;
  const std::string& getName() const { return name_; }
  ParamSet& params() { return params_; }
  template <typename T, typename PlannerType, typename SetterType, typename GetterType>
  void declareParam(const std::string& name, const PlannerType& planner, const SetterType& setter,
                    const GetterType& getter, const std::string& = "") {
    ParamSet::Param p;
    p.set = [planner, setter](const std::string& v) {
      std::istringstream ss(v);
      T t;
      ss >> t;
      if (ss.fail()) return false;
      (planner->*setter)(t);
      return true;
    };
    p.get = [planner, getter]() {
      std::ostringstream ss;
      ss << (planner->*getter)();
      return ss.str();
    };
    params_.params_[name] = p;
  }
protected:
  SpaceInformationPtr si_;
  ProblemDefinitionPtr pdef_;
  std::string name_;
  ParamSet params_;
  bool setup_ = false;
};
}  // namespace base
}  // namespace ompl
