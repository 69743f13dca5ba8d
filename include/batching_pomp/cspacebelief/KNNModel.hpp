// batching_pomp/cspacebelief/KNNModel.hpp — the k-nearest-neighbour belief model of
// /root/reference/include/batching_pomp/cspacebelief/KNNModel.hpp:56-134 with its constructor signature
// (k, prior, priorWeight, distance function).  The model itself lives on the GPU: points are appended to the
// device-resident model of a bpomp context (bpomp_belief_append) and estimates are the exact kNN estimates of
// csrc/belief.cu (bpomp_belief_estimate; same k nearest, same weights, same order of summation as
// KNNModel.hpp:108-130).  Before the planner binds the model to its device context, addPoint() buffers on the
// host and estimate() returns the prior of an empty model or throws if points are pending — there is no CPU
// fallback.  The distance function argument is accepted for source compatibility; the device model uses the
// state space's Euclidean distance, which is what the planner's beliefDistanceFunction (BP.cpp:213-219) binds.
#pragma once

#include <cstddef>
#include <functional>
#include <stdexcept>
#include <vector>

#include "bpomp_cuda.h"
#include "batching_pomp/cspacebelief/BeliefPoint.hpp"
#include "batching_pomp/cspacebelief/Model.hpp"

namespace batching_pomp {
namespace cspacebelief {

class KNNModel : public Model<BeliefPoint> {
public:
  KNNModel(std::size_t knn, double prior, double priorWeight,
           const std::function<double(const BeliefPoint&, const BeliefPoint&)>& distanceFunction = {})
      : mKNN(knn), mPrior(prior), mPriorWeight(priorWeight), mDistanceFunction(distanceFunction) {
    if (knn < 1 || knn > 32) throw std::invalid_argument("KNNModel (B200): k must be in 1..32");
  }

  std::size_t getK() const { return mKNN; }
  double getPrior() const { return mPrior; }
  double getPriorWeight() const { return mPriorWeight; }

  /// Called by the planner once its device context exists: pending points go to the device model.
  void bind(bpomp_ctx* ctx, unsigned int dim) {
    mCtx = ctx;
    mDim = dim;
    if (!mPendingVals.empty()) {
      check(bpomp_belief_append(mCtx, mPendingStates.data(), mPendingVals.data(), (int64_t)mPendingVals.size()));
      mPendingStates.clear();
      mPendingVals.clear();
    }
  }
  bool bound() const { return mCtx != nullptr; }

  void addPoint(const BeliefPoint& data) override {  // KNNModel.hpp:89-93
    if (mCtx) {
      if ((unsigned int)data.stateValues.size() != mDim) throw std::invalid_argument("KNNModel: wrong dimension");
      check(bpomp_belief_append(mCtx, &data.stateValues[0], &data.value, 1));
    } else {
      for (long j = 0; j < (long)data.stateValues.size(); ++j) mPendingStates.push_back(data.stateValues[j]);
      mPendingVals.push_back(data.value);
    }
  }

  void removePoint(const BeliefPoint&) override {  // KNNModel.hpp:95-99; the device model is append-only
    throw std::logic_error("KNNModel (B200): removePoint is not supported by the device-resident model");
  }

  double estimate(const BeliefPoint& query) const override {  // KNNModel.hpp:101-134
    if (!mCtx) {
      if (mPendingVals.empty()) return mPrior;  // :104-106
      throw std::logic_error("KNNModel (B200): estimate() needs the planner's device context (no CPU fallback)");
    }
    if ((unsigned int)query.stateValues.size() != mDim) throw std::invalid_argument("KNNModel: wrong dimension");
    double out = 0.0;
    check(bpomp_belief_estimate(mCtx, &query.stateValues[0], 1, &out));
    return out;
  }

  /// Number of points in the model.
  std::size_t size() const {
    if (!mCtx) return mPendingVals.size();
    uint64_t n = 0;
    check(bpomp_belief_size(mCtx, &n));
    return (std::size_t)n;
  }

private:
  void check(int rc) const {
    if (rc != BPOMP_OK) throw std::runtime_error(std::string("KNNModel (B200): ") + bpomp_last_error(mCtx));
  }
  std::size_t mKNN;
  double mPrior, mPriorWeight;
  std::function<double(const BeliefPoint&, const BeliefPoint&)> mDistanceFunction;
  bpomp_ctx* mCtx = nullptr;
  unsigned int mDim = 0;
  std::vector<double> mPendingStates, mPendingVals;
};

}  // namespace cspacebelief
}  // namespace batching_pomp
