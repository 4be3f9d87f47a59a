// Test helper: drives include/plainmp_b200_ompl.hpp the way ompl_thin_wrap.hpp:80-91 installs the
// reference's validators — a SpaceInformation over a RealVectorStateSpace, the session-backed
// state validity function and motion validator — compiled against the ompl::base stand-in of
// oracle/ref_shim/ (OMPL is not installed here).  Built by tests/test_gpu_parity.py with g++.
#include <cstdint>
#include <memory>
#include <vector>

#include "plainmp_b200_ompl.hpp"

namespace ob = ompl::base;

extern "C" int probe_validators(pmp_session_t* session, int dim, const double* s1, const double* s2, int n,
                                uint8_t* out_single, uint8_t* out_batch, uint8_t* out_s2_valid, uint64_t* n_is_valid_calls) {
  try {
    auto space = std::make_shared<ob::RealVectorStateSpace>((unsigned)dim);
    auto si = std::make_shared<ob::SpaceInformation>(space);
    std::size_t calls = 0;
    si->setStateValidityChecker(plainmp_b200::state_validity_fn(session, (unsigned)dim, &calls));
    auto mv = std::make_shared<plainmp_b200::SessionMotionValidator>(si, session, &calls);  // one counter, as in the reference
    std::vector<ob::State*> A(n), B(n);
    std::vector<std::pair<const ob::State*, const ob::State*>> edges;
    for (int i = 0; i < n; i++) {
      A[i] = si->allocState();
      B[i] = si->allocState();
      for (int k = 0; k < dim; k++) {
        A[i]->as<ob::RealVectorStateSpace::StateType>()->values[k] = s1[i * dim + k];
        B[i]->as<ob::RealVectorStateSpace::StateType>()->values[k] = s2[i * dim + k];
      }
      edges.push_back({A[i], B[i]});
    }
    const ob::MotionValidator& base = *mv;  // through the ompl::base interface, as a planner calls it
    for (int i = 0; i < n; i++) {
      out_single[i] = base.checkMotion(A[i], B[i]) ? 1 : 0;
      out_s2_valid[i] = si->isValid(B[i]) ? 1 : 0;
    }
    *n_is_valid_calls = calls;  // n isValid(s2) calls + every waypoint check of the n single checkMotion calls
    auto batch = mv->checkMotions(edges);
    for (int i = 0; i < n; i++) out_batch[i] = batch[i];
    for (int i = 0; i < n; i++) {
      si->freeState(A[i]);
      si->freeState(B[i]);
    }
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}
