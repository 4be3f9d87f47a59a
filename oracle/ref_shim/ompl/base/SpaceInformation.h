// TEST INFRASTRUCTURE — not part of the product.
//
// The few ompl::base declarations that plainmp's ompl/motion_validator.{hpp,cpp} touch, so that
// file can be compiled UNMODIFIED from /root/reference into oracle/_ref/ (OMPL is a system library
// that is absent from this image).  Restated from OMPL's published behaviour:
//   RealVectorStateSpace::interpolate  state[i] = from[i] + (to[i] - from[i]) * t
//   RealVectorStateSpace::distance     sqrt(sum (a[i] - b[i])^2)
//   SpaceInformation::isValid          forwards to the installed StateValidityChecker
#pragma once
#include <cmath>
#include <functional>
#include <memory>
#include <utility>

namespace ompl {
namespace base {

class State {
 public:
  virtual ~State() = default;
  template <class T>
  const T* as() const {
    return static_cast<const T*>(this);
  }
  template <class T>
  T* as() {
    return static_cast<T*>(this);
  }
};

class StateSpace {
 public:
  virtual ~StateSpace() = default;
  virtual unsigned int getDimension() const = 0;
  virtual void interpolate(const State* from, const State* to, double t, State* state) const = 0;
  virtual double distance(const State* a, const State* b) const = 0;
  virtual State* allocState() const = 0;
  virtual void freeState(State* s) const = 0;
};
using StateSpacePtr = std::shared_ptr<StateSpace>;

class RealVectorStateSpace : public StateSpace {
 public:
  class StateType : public State {
   public:
    double* values = nullptr;
    double operator[](unsigned int i) const { return values[i]; }
    double& operator[](unsigned int i) { return values[i]; }
  };
  explicit RealVectorStateSpace(unsigned int dim) : dim_(dim) {}
  unsigned int getDimension() const override { return dim_; }
  void interpolate(const State* from, const State* to, double t, State* state) const override {
    const auto* rfrom = static_cast<const StateType*>(from);
    const auto* rto = static_cast<const StateType*>(to);
    auto* rstate = static_cast<StateType*>(state);
    for (unsigned int i = 0; i < dim_; ++i) rstate->values[i] = rfrom->values[i] + (rto->values[i] - rfrom->values[i]) * t;
  }
  double distance(const State* a, const State* b) const override {
    double dist = 0.0;
    const double* s1 = static_cast<const StateType*>(a)->values;
    const double* s2 = static_cast<const StateType*>(b)->values;
    for (unsigned int i = 0; i < dim_; ++i) {
      double diff = (*s1++) - (*s2++);
      dist += diff * diff;
    }
    return std::sqrt(dist);
  }
  State* allocState() const override {
    auto* s = new StateType();
    s->values = new double[dim_];
    return s;
  }
  void freeState(State* s) const override {
    auto* r = static_cast<StateType*>(s);
    delete[] r->values;
    delete r;
  }

 private:
  unsigned int dim_;
};

class SpaceInformation {
 public:
  explicit SpaceInformation(StateSpacePtr space) : space_(std::move(space)) {}
  const StateSpacePtr& getStateSpace() const { return space_; }
  unsigned int getStateDimension() const { return space_->getDimension(); }
  State* allocState() const { return space_->allocState(); }
  void freeState(State* s) const { space_->freeState(s); }
  double distance(const State* a, const State* b) const { return space_->distance(a, b); }
  bool isValid(const State* s) const { return checker_(s); }
  void setStateValidityChecker(std::function<bool(const State*)> fn) { checker_ = std::move(fn); }

 private:
  StateSpacePtr space_;
  std::function<bool(const State*)> checker_;
};
using SpaceInformationPtr = std::shared_ptr<SpaceInformation>;

class MotionValidator {
 public:
  explicit MotionValidator(const SpaceInformationPtr& si) : si_(si.get()) {}
  virtual ~MotionValidator() = default;
  virtual bool checkMotion(const State* s1, const State* s2) const = 0;
  virtual bool checkMotion(const State* s1, const State* s2, std::pair<State*, double>& lastValid) const = 0;

 protected:
  SpaceInformation* si_;
};

}  // namespace base
}  // namespace ompl
