/*
 * plainmp_b200_ompl.hpp — header-only OMPL adaptor over the validator session of plainmp_b200.h.
 *
 * Drop-in for what the reference installs in CollisionAwareSpaceInformation
 * (cpp/plainmp/ompl/ompl_thin_wrap.hpp:80-91, 114-120):
 *
 *   si_->setMotionValidator(std::make_shared<BoxMotionValidator>(si_, vconfig.box_width));   // reference
 *   si_->setStateValidityChecker([this](const ob::State* s) { return this->is_valid(s); });
 * becomes
 *   auto mv = std::make_shared<plainmp_b200::SessionMotionValidator>(si_, session, &is_valid_call_count_);
 *   si_->setMotionValidator(mv);
 *   si_->setStateValidityChecker(plainmp_b200::state_validity_fn(session, dim, &is_valid_call_count_));
 *
 * where `session` comes from pmp_session_create(robot, scene, &validator, capacity, &session).
 * checkMotion(s1, s2) keeps the reference's contract (cpp/plainmp/ompl/motion_validator.cpp:17-79:
 * s1 is assumed valid, s2 and the interior waypoints are tested, lastValid is ignored as in
 * motion_validator.hpp:30-34) but tests all waypoints of the edge in one GPU wave;
 * checkMotions(...) is the additive batch call for planners that can propose many edges at once
 * (RRTConnect extensions over several trees, shortcutting candidates).
 *
 * Only ompl::base::{State, RealVectorStateSpace::StateType, SpaceInformation, MotionValidator} are
 * used, so the header also compiles against the few-line stand-ins under oracle/ref_shim/ompl/ that
 * this repo's tests use (OMPL is not installed in the build image).
 */
#ifndef PLAINMP_B200_OMPL_HPP_
#define PLAINMP_B200_OMPL_HPP_

#include <ompl/base/MotionValidator.h>
#include <ompl/base/SpaceInformation.h>
#include <ompl/base/spaces/RealVectorStateSpace.h>

#include <cstddef>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "plainmp_b200.h"

namespace plainmp_b200 {

namespace ob = ompl::base;

inline void throw_on_error(int status) {
  if (status != PMP_OK) throw std::runtime_error(std::string("plainmp_b200: ") + pmp_last_error());
}

inline const double* values_of(const ob::State* s) { return s->as<ob::RealVectorStateSpace::StateType>()->values; }

/* isValid(state): what CollisionAwareSpaceInformation::is_valid does (ompl_thin_wrap.hpp:114-120),
 * including the call counter. */
inline std::function<bool(const ob::State*)> state_validity_fn(pmp_session_t* session, unsigned dim,
                                                               std::size_t* call_count = nullptr) {
  return [session, dim, call_count](const ob::State* s) {
    (void)dim;
    std::uint8_t ok = 0;
    throw_on_error(pmp_session_is_valid(session, values_of(s), 1, &ok));
    if (call_count) ++*call_count;
    return ok != 0;
  };
}

class SessionMotionValidator : public ob::MotionValidator {
 public:
  /* call_count (optional): the counter the state validity function bumps.  In the reference every
   * waypoint of checkMotion goes through si_->isValid (motion_validator.cpp:52-71) and therefore through
   * is_valid_call_count_ (ompl_thin_wrap.hpp:114-120), which is what a planner's max_is_valid_call budget
   * looks at; the session reports exactly that number per call and it is added here. */
  SessionMotionValidator(const ob::SpaceInformationPtr& si, pmp_session_t* session, std::size_t* call_count = nullptr)
      : ob::MotionValidator(si), session_(session), dim_(si->getStateDimension()), call_count_(call_count) {}

  bool checkMotion(const ob::State* s1, const ob::State* s2) const override {
    std::uint8_t ok = 0;
    std::uint64_t n = 0;
    throw_on_error(pmp_session_check_motion(session_, values_of(s1), values_of(s2), 1, &ok, &n));
    checks_ += n;
    if (call_count_) *call_count_ += (std::size_t)n;
    return ok != 0;
  }
  bool checkMotion(const ob::State* s1, const ob::State* s2, std::pair<ob::State*, double>& /*lastValid*/) const override {
    return checkMotion(s1, s2);
  }

  /* Batch form: result[i] = checkMotion(edges[i].first, edges[i].second). */
  std::vector<char> checkMotions(const std::vector<std::pair<const ob::State*, const ob::State*>>& edges) const {
    const std::size_t n = edges.size();
    std::vector<double> a(n * dim_), b(n * dim_);
    for (std::size_t i = 0; i < n; i++)
      for (unsigned k = 0; k < dim_; k++) {
        a[i * dim_ + k] = values_of(edges[i].first)[k];
        b[i * dim_ + k] = values_of(edges[i].second)[k];
      }
    std::vector<std::uint8_t> ok(n);
    std::uint64_t cnt = 0;
    throw_on_error(pmp_session_check_motion(session_, a.data(), b.data(), (std::int64_t)n, ok.data(), &cnt));
    checks_ += cnt;
    if (call_count_) *call_count_ += (std::size_t)cnt;
    return std::vector<char>(ok.begin(), ok.end());
  }

  std::uint64_t waypoint_checks() const { return checks_; }

 private:
  pmp_session_t* session_;
  unsigned dim_;
  std::size_t* call_count_;
  mutable std::uint64_t checks_ = 0;
};

}  // namespace plainmp_b200
#endif /* PLAINMP_B200_OMPL_HPP_ */
