// TEST INFRASTRUCTURE — not part of the product.
//
// C entry points around the REFERENCE's own hot-path classes, compiled from the sources where they
// lie under /root/reference (see oracle/Makefile.ref): plainmp::kinematics::utils::_KinematicModel,
// plainmp::collision::{Box,Cylinder,Sphere,Ground,Cloud,Union}SDF and KDTree,
// plainmp::constraint::{SphereCollisionCst,IneqCompositeCst}, plainmp::ompl_wrapper::{Box,Euclidean}MotionValidator.
// This file contains no algorithm: it only constructs those objects and loops over batches, the way
// bindings/{kinematics,primitive,constraint,kdtree}.cpp and ompl/ompl_thin_wrap.hpp:80-120 do.
// Output: oracle/_ref/libplainmp_ref.so (git-ignored).  Used by tests/ to pin the restated oracle
// and by bench.py's reference arm as the CPU baseline.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

// The flatten header (include/plainmp_b200_plainmp.hpp) reads a few private members that a plainmp
// maintainer would expose with a `friend` declaration; the reference headers stay unmodified here, so
// this one translation unit sees them with private/protected spelled public (std headers come first).
#include <algorithm>
#include <array>
#include <fstream>
#include <functional>
#include <iostream>
#include <optional>
#include <stack>
#include <stdexcept>
#include <unordered_map>
#include <utility>
#define private public
#define protected public
#include "plainmp/collision/kdtree.hpp"
#include "plainmp/collision/primitive_sdf.hpp"
#include "plainmp/constraints/composite_constraint.hpp"
#include "plainmp/constraints/primitive_sphere_collision.hpp"
#include "plainmp/kinematics/kinematic_model_wrapper.hpp"
#include "plainmp/ompl/motion_validator.hpp"
#include "plainmp_b200_plainmp.hpp"
#undef private
#undef protected

namespace kin = plainmp::kinematics;
namespace col = plainmp::collision;
namespace cst = plainmp::constraint;
namespace ow = plainmp::ompl_wrapper;
namespace ob = ompl::base;

static thread_local std::string g_err;
static thread_local int g_err_kind = 0;  // 1: the last error was std::invalid_argument
static void set_err(const std::exception& e) {
  g_err = e.what();
  g_err_kind = dynamic_cast<const std::invalid_argument*>(&e) ? 1 : 0;
}
#define REF_TRY try {
#define REF_CATCH(ret)              \
  }                                 \
  catch (const std::exception& e) { \
    set_err(e);                     \
    return ret;                     \
  }

static std::vector<std::string> split_names(const char* joined) {
  std::vector<std::string> out;
  if (!joined || !*joined) return out;
  std::stringstream ss(joined);
  std::string item;
  while (std::getline(ss, item, '\n')) out.push_back(item);
  return out;
}

using KinPtr = std::shared_ptr<kin::utils::_KinematicModel>;
using SdfPtr = col::SDFBase::Ptr;
using CstPtr = std::shared_ptr<cst::SphereCollisionCst>;

static col::Pose make_pose(const double* pos, const double* rot_rowmajor) {
  Eigen::Matrix3d R;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R(i, j) = rot_rowmajor[3 * i + j];
  return col::Pose(Eigen::Vector3d(pos[0], pos[1], pos[2]), R);
}

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }
// 1 if the error of the last failing call on this thread was std::invalid_argument
int ref_last_error_is_invalid_argument() { return g_err_kind; }

// ---- kinematics (bindings/kinematics.cpp:30-50) ---------------------------------------------
void* ref_kin_create(const char* xml) {
  REF_TRY
  return new KinPtr(std::make_shared<kin::utils::_KinematicModel>(std::string(xml)));
  REF_CATCH(nullptr)
}
void ref_kin_destroy(void* h) { delete static_cast<KinPtr*>(h); }

int ref_kin_add_new_link(void* h, const char* name, const char* parent, const double* xyz, const double* rpy,
                         int consider_rotation) {
  REF_TRY
  return (int)(*static_cast<KinPtr*>(h))
      ->add_new_link_py(name, parent, {xyz[0], xyz[1], xyz[2]}, {rpy[0], rpy[1], rpy[2]}, consider_rotation != 0);
  REF_CATCH(-1)
}

int ref_kin_set_joint_positions(void* h, const char* joint_names, const double* q, int accurate) {
  try {
    auto& k = *static_cast<KinPtr*>(h);
    auto ids = k->get_joint_ids(split_names(joint_names));
    Eigen::VectorXd v(ids.size());
    for (size_t i = 0; i < ids.size(); i++) v[i] = q[i];
    k->set_joint_angles(ids, v, accurate != 0);
    return 0;
  } catch (const std::exception& e) {
    set_err(e);
    return -1;
  }
}

void ref_kin_set_base_pose(void* h, const double* pose7) {
  kin::utils::Vector7d v;
  for (int i = 0; i < 7; i++) v[i] = pose7[i];
  (*static_cast<KinPtr*>(h))->set_base_pose(v);
}

int ref_kin_link_pose(void* h, const char* link, double* out7) {
  try {
    auto v = (*static_cast<KinPtr*>(h))->debug_get_link_pose(link);
    for (int i = 0; i < 7; i++) out7[i] = v[i];
    return 0;
  } catch (const std::exception& e) {
    set_err(e);
    return -1;
  }
}

int ref_kin_joint_limits(void* h, const char* joint_names, double* out_n2) {
  REF_TRY
  auto& k = *static_cast<KinPtr*>(h);
  auto lims = k->get_joint_position_limits(k->get_joint_ids(split_names(joint_names)));
  for (size_t i = 0; i < lims.size(); i++) {
    out_n2[2 * i] = lims[i].first;
    out_n2[2 * i + 1] = lims[i].second;
  }
  return (int)lims.size();
  REF_CATCH(-1)
}

// ---- psdf (bindings/primitive.cpp:20-111) -----------------------------------------------------
void* ref_sdf_box(const double* width, const double* pos, const double* rot) {
  return new SdfPtr(std::make_shared<col::BoxSDF>(Eigen::Vector3d(width[0], width[1], width[2]), make_pose(pos, rot)));
}
void* ref_sdf_cylinder(double radius, double height, const double* pos, const double* rot) {
  return new SdfPtr(std::make_shared<col::CylinderSDF>(radius, height, make_pose(pos, rot)));
}
void* ref_sdf_sphere(double radius, const double* pos, const double* rot) {
  return new SdfPtr(std::make_shared<col::SphereSDF>(radius, make_pose(pos, rot)));
}
void* ref_sdf_ground(double height) { return new SdfPtr(std::make_shared<col::GroundSDF>(height)); }
void* ref_sdf_cloud(const double* pts_nx3, int64_t n, double radius) {
  std::vector<Eigen::Vector3d> pts((size_t)n);
  for (int64_t i = 0; i < n; i++) pts[(size_t)i] = Eigen::Vector3d(pts_nx3[3 * i], pts_nx3[3 * i + 1], pts_nx3[3 * i + 2]);
  return new SdfPtr(std::make_shared<col::CloudSDF>(pts, radius));
}
void* ref_sdf_union(void** members, int n) {
  std::vector<SdfPtr> v;
  for (int i = 0; i < n; i++) v.push_back(*static_cast<SdfPtr*>(members[i]));
  return new SdfPtr(std::make_shared<col::UnionSDF>(v));
}
void ref_sdf_destroy(void* h) { delete static_cast<SdfPtr*>(h); }
void* ref_sdf_clone(void* h) {
  REF_TRY
  return new SdfPtr((*static_cast<SdfPtr*>(h))->clone());
  REF_CATCH(nullptr)
}
int ref_sdf_translate(void* h, const double* t) {
  REF_TRY(*static_cast<SdfPtr*>(h))->translate(Eigen::Vector3d(t[0], t[1], t[2]));
  return 0;
  REF_CATCH(-1)
}
int ref_sdf_rotate_z(void* h, double a) {
  REF_TRY(*static_cast<SdfPtr*>(h))->rotate_z(a);
  return 0;
  REF_CATCH(-1)
}
// evaluate_batch takes 3xN column-major points in the binding; here [n][3] rows are repacked.
void ref_sdf_evaluate_batch(void* h, const double* pts_nx3, int64_t n, double* out) {
  col::Points P(3, n);
  for (int64_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) P(k, i) = pts_nx3[3 * i + k];
  col::Values v = (*static_cast<SdfPtr*>(h))->evaluate_batch(P);
  for (int64_t i = 0; i < n; i++) out[i] = v[i];
}
void ref_sdf_evaluate_each(void* h, const double* pts_nx3, int64_t n, double* out) {
  auto& s = *static_cast<SdfPtr*>(h);
  for (int64_t i = 0; i < n; i++) out[i] = s->evaluate(Eigen::Vector3d(pts_nx3[3 * i], pts_nx3[3 * i + 1], pts_nx3[3 * i + 2]));
}
void ref_sdf_is_outside_batch(void* h, const double* pts_nx3, const double* radii, int64_t n, uint8_t* out) {
  auto& s = *static_cast<SdfPtr*>(h);
  for (int64_t i = 0; i < n; i++)
    out[i] = s->is_outside(Eigen::Vector3d(pts_nx3[3 * i], pts_nx3[3 * i + 1], pts_nx3[3 * i + 2]), radii[i]) ? 1 : 0;
}
int ref_sdf_aabb(void* h, double* lb, double* ub) {
  auto& s = *static_cast<SdfPtr*>(h);
  if (s->get_type() == col::SDFType::UNION) return -1;
  auto p = std::static_pointer_cast<col::PrimitiveSDFBase>(s);
  for (int k = 0; k < 3; k++) {
    lb[k] = p->lb(k);
    ub[k] = p->ub(k);
  }
  return 0;
}
int ref_sdf_pose(void* h, double* pos, double* rot_rowmajor, int32_t* flags2) {
  auto& s = *static_cast<SdfPtr*>(h);
  auto t = s->get_type();
  if (t != col::SDFType::BOX && t != col::SDFType::CYLINDER && t != col::SDFType::SPHERE) return -1;
  auto p = std::static_pointer_cast<col::TransformableSDFBase>(s);
  for (int i = 0; i < 3; i++) {
    pos[i] = p->pose.position_(i);
    for (int j = 0; j < 3; j++) rot_rowmajor[3 * i + j] = p->pose.rot_(i, j);
  }
  flags2[0] = p->pose.axis_aligned_;
  flags2[1] = p->pose.z_axis_aligned_;
  return 0;
}

// ---- kdtree (bindings/kdtree.cpp:19-23) -------------------------------------------------------
void* ref_kdtree_create(const double* pts_nx3, int64_t n) {
  std::vector<Eigen::Vector3d> pts((size_t)n);
  for (int64_t i = 0; i < n; i++) pts[(size_t)i] = Eigen::Vector3d(pts_nx3[3 * i], pts_nx3[3 * i + 1], pts_nx3[3 * i + 2]);
  return new col::KDTree(pts);
}
void ref_kdtree_destroy(void* h) { delete static_cast<col::KDTree*>(h); }
void ref_kdtree_query_batch(void* h, const double* pts_nx3, int64_t n, double* sqdist, double* nearest_nx3) {
  auto* t = static_cast<col::KDTree*>(h);
  for (int64_t i = 0; i < n; i++) {
    Eigen::Vector3d p(pts_nx3[3 * i], pts_nx3[3 * i + 1], pts_nx3[3 * i + 2]);
    if (sqdist) sqdist[i] = t->sqdist(p);
    if (nearest_nx3) {
      Eigen::Vector3d q = t->query(p);
      for (int k = 0; k < 3; k++) nearest_nx3[3 * i + k] = q(k);
    }
  }
}

// ---- SphereCollisionCst (bindings/constraint.cpp:65-84) ---------------------------------------
void* ref_cst_create(void* kin_h, const char* control_joint_names, int base_type, int n_groups,
                     const char* group_link_names, const int32_t* group_offsets, const double* spheres_nx4,
                     const uint8_t* only_self, const char* pair_names, void* fixed_sdf, int reorder) {
  try {
    auto links = split_names(group_link_names);
    std::vector<cst::SphereAttachmentSpec> specs;
    for (int g = 0; g < n_groups; g++) {
      const int b = group_offsets[g], e = group_offsets[g + 1];
      Eigen::Matrix3Xd pos(3, e - b);
      Eigen::VectorXd rad(e - b);
      for (int s = b; s < e; s++) {
        for (int k = 0; k < 3; k++) pos(k, s - b) = spheres_nx4[4 * s + k];
        rad[s - b] = spheres_nx4[4 * s + 3];
      }
      specs.push_back({links[(size_t)g], pos, rad, only_self[g] != 0});
    }
    auto pn = split_names(pair_names);
    std::vector<std::pair<std::string, std::string>> pairs;
    for (size_t i = 0; i + 1 < pn.size(); i += 2) pairs.push_back({pn[i], pn[i + 1]});
    std::optional<SdfPtr> fixed = std::nullopt;
    if (fixed_sdf) fixed = *static_cast<SdfPtr*>(fixed_sdf);
    return new CstPtr(std::make_shared<cst::SphereCollisionCst>(
        std::static_pointer_cast<kin::KinematicModel<double>>(*static_cast<KinPtr*>(kin_h)),
        split_names(control_joint_names), static_cast<kin::BaseType>(base_type), specs, pairs, fixed, reorder != 0));
  } catch (const std::exception& e) {
    set_err(e);
    return nullptr;
  }
}
void ref_cst_destroy(void* h) { delete static_cast<CstPtr*>(h); }
void ref_cst_set_sdf(void* h, void* sdf) { (*static_cast<CstPtr*>(h))->set_sdf(sdf ? *static_cast<SdfPtr*>(sdf) : nullptr); }
int ref_cst_q_dim(void* h) { return (int)(*static_cast<CstPtr*>(h))->q_dim(); }
int ref_cst_dim(void* h) { return (int)(*static_cast<CstPtr*>(h))->cst_dim(); }

static Eigen::VectorXd row_of(const double* Q, int64_t i, int d) {
  Eigen::VectorXd q(d);
  for (int k = 0; k < d; k++) q[k] = Q[i * d + k];
  return q;
}

void ref_cst_is_valid_batch(void* h, const double* Q, int64_t n, uint8_t* out) {
  auto& c = *static_cast<CstPtr*>(h);
  const int d = (int)c->q_dim();
  Eigen::VectorXd q(d);
  for (int64_t i = 0; i < n; i++) {
    for (int k = 0; k < d; k++) q[k] = Q[i * d + k];
    out[i] = c->is_valid(q) ? 1 : 0;
  }
}

// vals [n][2] = (ext, self), NaN where the term is disabled; jac [n][2][q_dim] optional.
void ref_cst_evaluate_batch(void* h, const double* Q, int64_t n, double* vals, double* jac) {
  auto& c = *static_cast<CstPtr*>(h);
  const int d = (int)c->q_dim();
  const double nan = std::numeric_limits<double>::quiet_NaN();
  for (int64_t i = 0; i < n; i++) {
    auto r = c->evaluate(row_of(Q, i, d));
    int head = 0;
    for (int term = 0; term < 2; term++) {
      const bool on = term == 0 ? c->ext_colliision_enabled() : c->self_collision_enabled();
      vals[2 * i + term] = on ? r.first[head] : nan;
      if (jac)
        for (int k = 0; k < d; k++) jac[(2 * i + term) * d + k] = on ? r.second(head, k) : nan;
      if (on) head++;
    }
  }
}

int ref_cst_n_spheres(void* h) { return (int)(*static_cast<CstPtr*>(h))->get_all_spheres().size(); }
int ref_cst_n_groups(void* h) { return (int)(*static_cast<CstPtr*>(h))->get_group_spheres().size(); }

// out [n][n_spheres][4] after update_kintree(q, accurate)
void ref_cst_spheres_batch(void* h, const double* Q, int64_t n, int accurate, int group_level, double* out) {
  auto& c = *static_cast<CstPtr*>(h);
  const int d = (int)c->q_dim();
  int64_t w = 0;
  for (int64_t i = 0; i < n; i++) {
    c->update_kintree(row_of(Q, i, d), accurate != 0);
    c->post_update_kintree();
    auto sp = group_level ? c->get_group_spheres() : c->get_all_spheres();
    for (auto& s : sp) {
      out[w++] = s.first(0);
      out[w++] = s.first(1);
      out[w++] = s.first(2);
      out[w++] = s.second;
    }
  }
}

// ---- reference objects -> C-ABI tables (include/plainmp_b200_plainmp.hpp) ------------------------
void* ref_cst_flatten(void* h) {
  REF_TRY
  return new plainmp_b200::Flat(plainmp_b200::Flatten::run(**static_cast<CstPtr*>(h)));
  REF_CATCH(nullptr)
}
void ref_flat_destroy(void* f) { delete static_cast<plainmp_b200::Flat*>(f); }
// out[8]: n_ops, n_groups, n_spheres, n_pairs, n_prims, n_cloud_points, dof, base_type
void ref_flat_sizes(void* f, int64_t* out) {
  auto* F = static_cast<plainmp_b200::Flat*>(f);
  out[0] = (int64_t)F->ops.size();
  out[1] = (int64_t)F->groups.size();
  out[2] = (int64_t)F->spheres.size() / 4;
  out[3] = (int64_t)F->pairs.size() / 2;
  out[4] = (int64_t)F->prims.size();
  out[5] = (int64_t)F->cloud_points.size() / 3;
  out[6] = F->dof;
  out[7] = F->base_type;
}
// cloud_point_radii [n_cloud_points]: filled (and *has_point_radii set) only when the CloudSDF leaves have different radii
void ref_flat_copy(void* f, void* ops, void* groups, float* spheres, int32_t* pairs, void* prims, float* cloud,
                   float* base_pose7, float* cloud_radius, float* cloud_point_radii, int32_t* has_point_radii) {
  auto* F = static_cast<plainmp_b200::Flat*>(f);
  std::memcpy(ops, F->ops.data(), F->ops.size() * sizeof(pmp_link_op_t));
  std::memcpy(groups, F->groups.data(), F->groups.size() * sizeof(pmp_group_t));
  std::memcpy(spheres, F->spheres.data(), F->spheres.size() * sizeof(float));
  std::memcpy(pairs, F->pairs.data(), F->pairs.size() * sizeof(int32_t));
  std::memcpy(prims, F->prims.data(), F->prims.size() * sizeof(pmp_prim_t));
  std::memcpy(cloud, F->cloud_points.data(), F->cloud_points.size() * sizeof(float));
  std::memcpy(base_pose7, F->base_pose, sizeof(F->base_pose));
  *cloud_radius = F->cloud_radius;
  *has_point_radii = F->cloud_point_radii.empty() ? 0 : 1;
  if (!F->cloud_point_radii.empty()) std::memcpy(cloud_point_radii, F->cloud_point_radii.data(), F->cloud_point_radii.size() * sizeof(float));
}

// ---- motion validators (ompl/motion_validator.cpp, installed as in ompl_thin_wrap.hpp:80-120) ----
// calls_per_edge (optional): how many times checkMotion called isValid for edge i — the position of the
// first failing waypoint in the reference's visiting order, else the waypoint count
void ref_check_motion_batch_ex(void* cst_h, int type, int dim, const double* box_width, double resolution,
                               const double* S1, const double* S2, int64_t n, uint8_t* out, uint32_t* calls_per_edge,
                               uint64_t* n_checked);
void ref_check_motion_batch(void* cst_h, int type, int dim, const double* box_width, double resolution,
                            const double* S1, const double* S2, int64_t n, uint8_t* out, uint64_t* n_checked) {
  ref_check_motion_batch_ex(cst_h, type, dim, box_width, resolution, S1, S2, n, out, nullptr, n_checked);
}
void ref_check_motion_batch_ex(void* cst_h, int type, int dim, const double* box_width, double resolution,
                               const double* S1, const double* S2, int64_t n, uint8_t* out, uint32_t* calls_per_edge,
                               uint64_t* n_checked) {
  auto& c = *static_cast<CstPtr*>(cst_h);
  auto space = std::make_shared<ob::RealVectorStateSpace>((unsigned)dim);
  auto si = std::make_shared<ob::SpaceInformation>(space);
  uint64_t count = 0;
  Eigen::VectorXd tmp(dim);
  si->setStateValidityChecker([&](const ob::State* s) {
    auto rs = s->as<ob::RealVectorStateSpace::StateType>();
    for (int k = 0; k < dim; k++) tmp[k] = rs->values[k];
    count++;
    return c->is_valid(tmp);
  });
  std::unique_ptr<ow::CustomValidatorBase> mv;
  if (type == 0)
    mv.reset(new ow::BoxMotionValidator(si, std::vector<double>(box_width, box_width + dim)));
  else
    mv.reset(new ow::EuclideanMotionValidator(si, resolution));
  auto* a = si->allocState()->as<ob::RealVectorStateSpace::StateType>();
  auto* b = si->allocState()->as<ob::RealVectorStateSpace::StateType>();
  for (int64_t i = 0; i < n; i++) {
    for (int k = 0; k < dim; k++) {
      a->values[k] = S1[i * dim + k];
      b->values[k] = S2[i * dim + k];
    }
    const uint64_t before = count;
    out[i] = mv->checkMotion(a, b) ? 1 : 0;
    if (calls_per_edge) calls_per_edge[i] = (uint32_t)(count - before);
  }
  si->freeState(a);
  si->freeState(b);
  if (n_checked) *n_checked = count;
}

// ---- IneqCompositeCst (constraints/composite_constraint.hpp:88-105, bindings/constraint.cpp:108-113) ----
using CompPtr = std::shared_ptr<cst::IneqCompositeCst>;
void* ref_composite_create(void** members, int n) {
  REF_TRY
  std::vector<cst::IneqConstraintBase::Ptr> v;
  for (int i = 0; i < n; i++) v.push_back(*static_cast<CstPtr*>(members[i]));
  return new CompPtr(std::make_shared<cst::IneqCompositeCst>(v));
  REF_CATCH(nullptr)
}
void ref_composite_destroy(void* h) { delete static_cast<CompPtr*>(h); }
int ref_composite_dim(void* h) { return (int)(*static_cast<CompPtr*>(h))->CompositeConstraintBase::cst_dim(); }
void ref_composite_is_valid_batch(void* h, const double* Q, int64_t n, uint8_t* out) {
  auto& c = *static_cast<CompPtr*>(h);
  const int d = (int)c->q_dim();
  for (int64_t i = 0; i < n; i++) out[i] = c->is_valid(row_of(Q, i, d)) ? 1 : 0;
}
// vals [n][cst_dim], jac [n][cst_dim][q_dim] (optional): members stacked in order, disabled terms absent
void ref_composite_evaluate_batch(void* h, const double* Q, int64_t n, double* vals, double* jac) {
  auto& c = *static_cast<CompPtr*>(h);
  const int d = (int)c->q_dim();
  const int m = (int)c->CompositeConstraintBase::cst_dim();
  for (int64_t i = 0; i < n; i++) {
    auto r = c->evaluate(row_of(Q, i, d));
    for (int t = 0; t < m; t++) {
      vals[i * m + t] = r.first[t];
      if (jac)
        for (int k = 0; k < d; k++) jac[(i * m + t) * d + k] = r.second(t, k);
    }
  }
}

double ref_step_ratio(int type, int dim, const double* box_width, double resolution, const double* s1, const double* s2) {
  auto space = std::make_shared<ob::RealVectorStateSpace>((unsigned)dim);
  auto si = std::make_shared<ob::SpaceInformation>(space);
  std::unique_ptr<ow::CustomValidatorBase> mv;
  if (type == 0)
    mv.reset(new ow::BoxMotionValidator(si, std::vector<double>(box_width, box_width + dim)));
  else
    mv.reset(new ow::EuclideanMotionValidator(si, resolution));
  auto* a = si->allocState()->as<ob::RealVectorStateSpace::StateType>();
  auto* b = si->allocState()->as<ob::RealVectorStateSpace::StateType>();
  for (int k = 0; k < dim; k++) {
    a->values[k] = s1[k];
    b->values[k] = s2[k];
  }
  const double r = mv->determine_step_ratio(a, b);
  si->freeState(a);
  si->freeState(b);
  return r;
}

}  // extern "C"
