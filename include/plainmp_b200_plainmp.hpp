/*
 * plainmp_b200_plainmp.hpp — the REFERENCE-SIDE half of the drop-in: flattens plainmp's own C++
 * objects into the tables of include/plainmp_b200.h.  It is meant to be compiled INSIDE the plainmp
 * tree (it includes plainmp's headers and reads plainmp's members); it contains no CUDA and does not
 * link against libplainmp_b200 — it only fills pmp_robot_desc_t / pmp_scene_desc_t.
 *
 *   plainmp::constraint::SphereCollisionCst  --flatten-->  pmp_robot_desc_t  (+ pmp_scene_desc_t)
 *
 * What is read, and where it lives upstream (paths relative to cpp/plainmp/):
 *   kinematics/kinematics.hpp:62-91      link_parent_link_ids_, joint_{types,axes,positions,orientations}_,
 *                                        joint_orientation_identity_flags_, joint_child_link_ids_,
 *                                        tf_plink_to_hlink_cache_ (all public), base_pose_
 *   constraints/primitive.hpp:96-99      control_joint_ids_, base_type_ (public)
 *   constraints/primitive_sphere_collision.hpp:108-114   sphere_groups_, selcol_group_id_pairs_,
 *                                        all_sdfs_cache_ (PRIVATE)
 *   collision/primitive_sdf.hpp          Pose (public), lb / ub (public); BoxSDF::half_width_,
 *                                        CylinderSDF::r_cylinder_ / half_height_, SphereSDF::r_sphere_,
 *                                        GroundSDF::height_, CloudSDF::kdtree_ / radius_ (PRIVATE)
 *   collision/kdtree.hpp:36              KDTree::nodes_ (PRIVATE; holds every input point once)
 * A maintainer adds `friend struct plainmp_b200::Flatten;` to the six classes with private members
 * (SphereCollisionCst, BoxSDF, CylinderSDF, SphereSDF, GroundSDF, CloudSDF, KDTree); this repo's test
 * build (oracle/ref_driver.cpp) compiles the unmodified reference headers with `-Dprivate=public`
 * instead.
 *
 * Semantics mirrored: the op list is the ancestor-closed, depth-first set of the
 * links that carry sphere groups; a controlled joint becomes a REVOLUTE / PRISMATIC op with the
 * joint origin (a prismatic joint drops the origin rotation, kinematics/algorithm.cpp:149-155);
 * every other link is a FIXED op holding tf_plink_to_hlink_cache_ as it stands, i.e. non-controlled
 * joints are baked at their current angle; identity flags are copied, never recomputed; groups keep
 * the reference's (re-ordered) sphere order; all_sdfs_cache_ is already flattened, anchors first.
 * Re-flatten after set_sdf, after translate / rotate_z on a leaf, and after moving a non-controlled
 * joint or the base pose (or use pmp_scene_update_prims for the scene).
 */
#ifndef PLAINMP_B200_PLAINMP_HPP_
#define PLAINMP_B200_PLAINMP_HPP_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <unordered_map>
#include <vector>

// Include AFTER "plainmp/constraints/primitive_sphere_collision.hpp": that header has no include
// guard upstream, so it cannot be pulled in again from here (it brings primitive_sdf.hpp, kdtree.hpp,
// kinematics.hpp and urdf_model/joint.h with it).
#include "plainmp_b200.h"

namespace plainmp_b200 {

struct Flat {
  std::vector<pmp_link_op_t> ops;
  std::vector<pmp_group_t> groups;
  std::vector<float> spheres;  // [n][4]
  std::vector<int32_t> pairs;  // [n][2]
  std::vector<pmp_prim_t> prims;
  std::vector<float> cloud_points;  // [n][3]: the points of ALL CloudSDF leaves
  std::vector<float> cloud_point_radii;  // [n]: filled only when the leaves have different radii
  float cloud_radius = 0.f;
  int32_t dof = 0, base_type = PMP_BASE_FIXED;
  float base_pose[7] = {0, 0, 0, 0, 0, 0, 1};

  pmp_robot_desc_t robot_desc() const {
    pmp_robot_desc_t d{};
    d.dof = dof;
    d.base_type = base_type;
    std::memcpy(d.base_pose, base_pose, sizeof(base_pose));
    d.n_ops = (int32_t)ops.size();
    d.ops = ops.data();
    d.n_groups = (int32_t)groups.size();
    d.groups = groups.data();
    d.n_spheres = (int32_t)(spheres.size() / 4);
    d.spheres = spheres.data();
    d.n_pairs = (int32_t)(pairs.size() / 2);
    d.pairs = pairs.data();
    return d;
  }
  pmp_scene_desc_t scene_desc() const {
    pmp_scene_desc_t d{};
    d.n_prims = (int32_t)prims.size();
    d.prims = prims.data();
    d.n_cloud_points = (int64_t)(cloud_points.size() / 3);
    d.cloud_points = cloud_points.data();
    d.cloud_radius = cloud_radius;
    d.cloud_cell = 0.f;
    d.cloud_point_radii = cloud_point_radii.empty() ? nullptr : cloud_point_radii.data();
    return d;
  }
  bool has_scene() const { return !prims.empty() || !cloud_points.empty(); }
};

struct Flatten {
  using Cst = plainmp::constraint::SphereCollisionCst;
  using Kin = plainmp::kinematics::KinematicModel<double>;
  using BaseType = plainmp::kinematics::BaseType;

  static Flat run(const Cst& cst) {
    Flat out;
    robot(cst, &out);
    scene(cst, &out);
    return out;
  }

  // ---- robot: kinematic tables + sphere groups + pairs ---------------------------------------
  static void robot(const Cst& cst, Flat* out) {
    const Kin& kin = *cst.kin_;
    const size_t n_links = kin.link_parent_link_ids_.size();
    const size_t none = std::numeric_limits<size_t>::max();
    // ancestor-closed set of the group links (+ the root)
    std::vector<char> needed(n_links, 0);
    for (const auto& g : cst.sphere_groups_)
      for (size_t l = g.parent_link_id; l != none && !needed[l]; l = kin.link_parent_link_ids_[l]) needed[l] = 1;
    needed[kin.root_link_id_] = 1;
    // depth-first order so that chains stay adjacent (parent == index - 1)
    std::vector<size_t> order, stack{kin.root_link_id_};
    while (!stack.empty()) {
      const size_t l = stack.back();
      stack.pop_back();
      order.push_back(l);
      const auto& ch = kin.link_child_link_idss_[l];
      for (size_t k = ch.size(); k-- > 0;)
        if (needed[ch[k]]) stack.push_back(ch[k]);
    }
    std::unordered_map<size_t, int> op_of_link;
    for (size_t k = 0; k < order.size(); k++) op_of_link[order[k]] = (int)k;
    // controlled joints: child link -> (joint id, column of q)
    std::unordered_map<size_t, std::pair<size_t, int>> ctrl;
    for (size_t k = 0; k < cst.control_joint_ids_.size(); k++)
      ctrl[(size_t)kin.joint_child_link_ids_[cst.control_joint_ids_[k]]] = {cst.control_joint_ids_[k], (int)k};

    out->ops.assign(order.size(), pmp_link_op_t{});
    for (size_t k = 0; k < order.size(); k++) {
      const size_t l = order[k];
      if (!kin.link_consider_rotation_[l])
        throw std::runtime_error("plainmp_b200: consider_rotation=false links cannot carry spheres or children");
      pmp_link_op_t& op = out->ops[k];
      op.origin_quat[3] = 1.f;
      op.q_index = -1;
      op.kind = PMP_OP_FIXED;
      const size_t p = kin.link_parent_link_ids_[l];
      op.parent = (p == none) ? -1 : op_of_link.at(p);
      auto it = ctrl.find(l);
      if (p == none) {  // the root: frame = base pose
        op.flags = PMP_OPFLAG_ORIGIN_IDENTITY;
      } else if (it != ctrl.end()) {
        const size_t j = it->second.first;
        op.q_index = it->second.second;
        for (int a = 0; a < 3; a++) {
          op.origin_xyz[a] = (float)kin.joint_positions_[j](a);
          op.axis[a] = (float)kin.joint_axes_[j](a);
        }
        if (kin.joint_types_[j] == urdf::Joint::PRISMATIC) {  // algorithm.cpp:149-155
          op.kind = PMP_OP_PRISMATIC;
          op.flags = PMP_OPFLAG_ORIGIN_IDENTITY;
        } else {  // algorithm.cpp:124-148
          op.kind = PMP_OP_REVOLUTE;
          const auto& oq = kin.joint_orientations_[j];
          op.origin_quat[0] = (float)oq.x();
          op.origin_quat[1] = (float)oq.y();
          op.origin_quat[2] = (float)oq.z();
          op.origin_quat[3] = (float)oq.w();
          if (kin.all_joint_orientation_identity_ || kin.joint_orientation_identity_flags_[j]) op.flags = PMP_OPFLAG_ORIGIN_IDENTITY;
        }
      } else {  // fixed joint, or a movable joint that is not controlled: the cached local transform
        const auto& tf = kin.tf_plink_to_hlink_cache_[l];
        op.origin_quat[0] = (float)tf.quat().x();
        op.origin_quat[1] = (float)tf.quat().y();
        op.origin_quat[2] = (float)tf.quat().z();
        op.origin_quat[3] = (float)tf.quat().w();
        for (int a = 0; a < 3; a++) op.origin_xyz[a] = (float)tf.trans()(a);
        op.flags = tf.is_quat_identity_ ? PMP_OPFLAG_ORIGIN_IDENTITY : 0;
      }
    }
    out->dof = (int32_t)cst.control_joint_ids_.size();
    out->base_type = cst.base_type_ == BaseType::FIXED ? PMP_BASE_FIXED
                     : cst.base_type_ == BaseType::PLANAR ? PMP_BASE_PLANAR : PMP_BASE_FLOATING;
    const auto& bp = kin.base_pose_;
    out->base_pose[0] = (float)bp.trans()(0);
    out->base_pose[1] = (float)bp.trans()(1);
    out->base_pose[2] = (float)bp.trans()(2);
    out->base_pose[3] = (float)bp.quat().x();
    out->base_pose[4] = (float)bp.quat().y();
    out->base_pose[5] = (float)bp.quat().z();
    out->base_pose[6] = (float)bp.quat().w();

    out->groups.clear();
    out->spheres.clear();
    for (const auto& g : cst.sphere_groups_) {
      pmp_group_t G{};
      for (int a = 0; a < 3; a++) G.center[a] = (float)g.group_sphere_relative_position(a);
      G.radius = (float)g.group_radius;
      G.op = op_of_link.at(g.parent_link_id);
      G.sphere_begin = (int32_t)(out->spheres.size() / 4);
      G.sphere_count = (int32_t)g.radii.size();
      G.only_self_collision = g.only_self_collision ? 1 : 0;
      for (int i = 0; i < (int)g.radii.size(); i++) {
        for (int a = 0; a < 3; a++) out->spheres.push_back((float)g.sphere_relative_positions(a, i));
        out->spheres.push_back((float)g.radii[i]);
      }
      out->groups.push_back(G);
    }
    out->pairs.clear();
    for (const auto& pr : cst.selcol_group_id_pairs_) {
      out->pairs.push_back((int32_t)pr.first);
      out->pairs.push_back((int32_t)pr.second);
    }
  }

  // ---- scene: all_sdfs_cache_ (anchors first, unions already expanded) ------------------------
  static void scene(const Cst& cst, Flat* out) {
    namespace col = plainmp::collision;
    out->prims.clear();
    out->cloud_points.clear();
    out->cloud_point_radii.clear();
    out->cloud_radius = 0.f;
    std::vector<float> leaf_radii, per_point;
    for (const auto& s : cst.all_sdfs_cache_) {
      const col::SDFType t = s->get_type();
      if (t == col::SDFType::CLOUD) {
        // every CloudSDF leaf goes into ONE point set (one device index); leaves of different radii are kept
        // apart by the per-point radius (radius classes, pmp_scene_desc_t.cloud_point_radii)
        const auto& c = static_cast<const col::CloudSDF&>(*s);
        leaf_radii.push_back((float)c.radius_);
        for (const auto& node : c.kdtree_->nodes_) {
          for (int a = 0; a < 3; a++) out->cloud_points.push_back((float)node.point(a));
          per_point.push_back((float)c.radius_);
        }
        out->cloud_radius = std::max(out->cloud_radius, (float)c.radius_);
        continue;
      }
      pmp_prim_t P{};
      for (int a = 0; a < 3; a++) {
        P.lb[a] = (float)std::max(s->lb(a), -3.0e38);
        P.ub[a] = (float)std::min(s->ub(a), 3.0e38);
      }
      P.rot_col0[0] = P.rot_col1[1] = P.rot_col2[2] = 1.f;
      if (t == col::SDFType::GROUND) {
        P.type = PMP_PRIM_GROUND;
        P.param0 = (float)static_cast<const col::GroundSDF&>(*s).height_;
        P.flags = PMP_PRIMFLAG_AXIS_ALIGNED | PMP_PRIMFLAG_Z_ALIGNED;
        out->prims.push_back(P);
        continue;
      }
      const col::Pose& pose = static_cast<const col::TransformableSDFBase&>(*s).pose;
      P.flags = (pose.axis_aligned_ ? PMP_PRIMFLAG_AXIS_ALIGNED : 0) | (pose.z_axis_aligned_ ? PMP_PRIMFLAG_Z_ALIGNED : 0);
      for (int a = 0; a < 3; a++) P.position[a] = (float)pose.position_(a);
      bool world_axes = pose.axis_aligned_;  // the reference skips the rotation entirely (primitive_sdf.hpp:328-340)
      if (t == col::SDFType::BOX) {
        const auto& b = static_cast<const col::BoxSDF&>(*s);
        P.type = PMP_PRIM_BOX;
        P.param0 = (float)b.half_width_(0);
        P.param1 = (float)b.half_width_(1);
        P.param2 = (float)b.half_width_(2);
      } else if (t == col::SDFType::CYLINDER) {
        const auto& c = static_cast<const col::CylinderSDF&>(*s);
        P.type = PMP_PRIM_CYLINDER;
        P.param0 = (float)c.r_cylinder_;
        P.param1 = (float)c.half_height_;
        world_axes = world_axes || pose.z_axis_aligned_;  // world axes for all three coordinates (:446-449, 466-472)
      } else if (t == col::SDFType::SPHERE) {
        P.type = PMP_PRIM_SPHERE;
        P.param0 = (float)static_cast<const col::SphereSDF&>(*s).r_sphere_;
      } else {
        throw std::runtime_error("plainmp_b200: unexpected SDF type in all_sdfs_cache_");
      }
      if (!world_axes) {
        for (int a = 0; a < 3; a++) {
          P.rot_col0[a] = (float)pose.rot_(a, 0);
          P.rot_col1[a] = (float)pose.rot_(a, 1);
          P.rot_col2[a] = (float)pose.rot_(a, 2);
        }
      }
      out->prims.push_back(P);
    }
    for (float r : leaf_radii)
      if (r != leaf_radii.front()) {  // different radii: hand the per-point radii over
        out->cloud_point_radii = per_point;
        break;
      }
  }
};

}  // namespace plainmp_b200
#endif /* PLAINMP_B200_PLAINMP_HPP_ */
