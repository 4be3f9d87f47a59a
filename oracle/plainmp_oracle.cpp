// plainmp_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// FP64, scalar, single-file CPU restatement of the plainmp hot path
// (validity checking of configurations and motion edges).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The product (plainmp_b200) never does.
//
// Parity status: the upstream library cannot be built in this container
// (Eigen3/Boost/OMPL absent), and it ships no golden vectors.
// This restatement is pinned against the known-answer facts the reference's own
// tests and comments contain (tests/test_oracle_kat.py) and against the
// reference's own generate_sequence.py run in this container
// (tests/golden/sequence_table.json).  Everything else is "parity unpinned":
// agreement is between this restatement and the CUDA path.
//
// Every function cites the reference file:line it restates
// (paths relative to /root/reference/cpp/plainmp unless noted).
// Eigen is absent: the handful of Eigen 3.4 primitives used by the reference
// (quaternion product, quat*vec, toRotationMatrix, Matrix3d::inverse, isApprox)
// are restated in the "eig" namespace below, in Eigen's operation order.

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

namespace eig {

struct V3 {
  double x = 0, y = 0, z = 0;
  double operator()(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  double& operator()(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(const V3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3& a, const V3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double sqnorm(const V3& a) { return dot(a, a); }
inline double norm(const V3& a) { return std::sqrt(sqnorm(a)); }

// Eigen::Quaternion coefficient order (x, y, z, w).
struct Quat {
  double x = 0, y = 0, z = 0, w = 1;
};
// Eigen/src/Geometry/Quaternion.h: quat_product<..>::run
inline Quat operator*(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
// QuaternionBase::_transformVector: uv = 2 * (q.vec x v); v + w*uv + q.vec x uv
inline V3 operator*(const Quat& q, const V3& v) {
  V3 qv{q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = uv + uv;
  return v + uv * q.w + cross(qv, uv);
}

struct M3 {
  double m[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double operator()(int r, int c) const { return m[r][c]; }
  V3 col(int c) const { return {m[0][c], m[1][c], m[2][c]}; }
};
inline V3 operator*(const M3& a, const V3& v) {
  return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
          a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
          a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
inline M3 operator*(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
// QuaternionBase::toRotationMatrix
inline M3 to_rotation_matrix(const Quat& q) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 r;
  r.m[0][0] = 1 - (tyy + tzz);
  r.m[0][1] = txy - twz;
  r.m[0][2] = txz + twy;
  r.m[1][0] = txy + twz;
  r.m[1][1] = 1 - (txx + tzz);
  r.m[1][2] = tyz - twx;
  r.m[2][0] = txz - twy;
  r.m[2][1] = tyz + twx;
  r.m[2][2] = 1 - (txx + tyy);
  return r;
}
// Eigen/src/LU/InverseImpl.h compute_inverse<.., 3>: cofactors / determinant
inline M3 inverse(const M3& a) {
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return a.m[i1][j1] * a.m[i2][j2] - a.m[i1][j2] * a.m[i2][j1];
  };
  const double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  const double det = c00 * a.m[0][0] + c10 * a.m[1][0] + c20 * a.m[2][0];
  const double invdet = 1.0 / det;
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[j][i] = cof(i, j) * invdet;
  return r;
}
// DenseBase::isApprox(Identity) with the default double precision 1e-12
inline bool is_approx_identity(const M3& a) {
  double diff = 0, na = 0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      const double e = a.m[i][j] - (i == j ? 1.0 : 0.0);
      diff += e * e;
      na += a.m[i][j] * a.m[i][j];
    }
  const double prec = 1e-12;
  return diff <= prec * prec * std::min(na, 3.0);
}

}  // namespace eig

namespace orc {
using eig::M3;
using eig::Quat;
using eig::V3;

// ----------------------------------------------------------------------------
// kinematics/transform.hpp:17-74  QuatTrans
// ----------------------------------------------------------------------------
struct QuatTrans {
  Quat quat;
  V3 trans;
  bool is_quat_identity = false;
  // transform.hpp:25-30: operator= copies quat and trans only (not the flag)
  void assign(const QuatTrans& o) {
    quat = o.quat;
    trans = o.trans;
  }
  // transform.hpp:64-74
  void quat_identity_sensitive_mult_and_assign(const QuatTrans& other, QuatTrans& result) const {
    if (other.is_quat_identity) {
      result.quat = quat;
      result.trans = trans + quat * other.trans;
    } else {
      result.quat = quat * other.quat;
      result.trans = trans + quat * other.trans;
    }
  }
  // transform.hpp:56-62 operator*
  QuatTrans mul(const QuatTrans& other) const {
    QuatTrans r;
    r.quat = other.is_quat_identity ? quat : quat * other.quat;
    r.trans = trans + quat * other.trans;
    return r;
  }
  // transform.hpp:87-90 getInverse (Eigen Quaternion::inverse = conjugate / squaredNorm)
  QuatTrans get_inverse() const {
    const double n2 = quat.x * quat.x + quat.y * quat.y + quat.z * quat.z + quat.w * quat.w;
    QuatTrans r;
    r.quat = {-quat.x / n2, -quat.y / n2, -quat.z / n2, quat.w / n2};
    r.trans = r.quat * (trans * -1.0);
    return r;
  }
  // transform.hpp:92-120 getRPY
  V3 get_rpy() const {
    const double sqx = quat.x * quat.x, sqy = quat.y * quat.y, sqz = quat.z * quat.z, sqw = quat.w * quat.w;
    const double sarg = -2 * (quat.x * quat.z - quat.w * quat.y);
    const double pi_2 = 1.57079632679489661923;
    double roll, pitch, yaw;
    if (sarg <= -0.99999) {
      pitch = -pi_2;
      roll = 0;
      yaw = -2 * atan2(quat.x, quat.y);
    } else if (sarg >= 0.99999) {
      pitch = pi_2;
      roll = 0;
      yaw = 2 * atan2(quat.x, quat.y);
    } else {
      pitch = asin(sarg);
      roll = atan2(2 * (quat.y * quat.z + quat.w * quat.x), sqw - sqx - sqy + sqz);
      yaw = atan2(2 * (quat.x * quat.y + quat.w * quat.z), sqw + sqx - sqy - sqz);
    }
    return {roll, pitch, yaw};
  }
  // transform.hpp:122-130
  void set_quaternion_from_rpy(double roll, double pitch, double yaw) {
    const double phi = roll / 2.0, the = pitch / 2.0, psi = yaw / 2.0;
    quat.x = sin(phi) * cos(the) * cos(psi) - cos(phi) * sin(the) * sin(psi);
    quat.y = cos(phi) * sin(the) * cos(psi) + sin(phi) * cos(the) * sin(psi);
    quat.z = cos(phi) * cos(the) * sin(psi) - sin(phi) * sin(the) * cos(psi);
    quat.w = cos(phi) * cos(the) * cos(psi) + sin(phi) * sin(the) * sin(psi);
  }
};

// kinematics/algorithm.cpp:49-86
static void compute_approx_sin_cos(double x, double& s, double& c) {
  constexpr auto coeff2 = 1.0 / (1.0 * 2.0);
  constexpr auto coeff3 = 1.0 / (1.0 * 2.0 * 3.0);
  constexpr auto coeff4 = 1.0 / (1.0 * 2.0 * 3.0 * 4.0);
  constexpr auto coeff5 = 1.0 / (1.0 * 2.0 * 3.0 * 4.0 * 5.0);
  constexpr auto coeff6 = 1.0 / (1.0 * 2.0 * 3.0 * 4.0 * 5.0 * 6.0);
  constexpr auto coeff7 = 1.0 / (1.0 * 2.0 * 3.0 * 4.0 * 5.0 * 6.0 * 7.0);
  constexpr auto coeff8 = 1.0 / (1.0 * 2.0 * 3.0 * 4.0 * 5.0 * 6.0 * 7.0 * 8.0);
  constexpr auto coeff9 = 1.0 / (1.0 * 2.0 * 3.0 * 4.0 * 5.0 * 6.0 * 7.0 * 8.0 * 9.0);
  constexpr auto half_pi = M_PI * 0.5;
  constexpr auto one_dev_2pi = 1.0 / (2 * M_PI);
  auto cos_sign = 1.0;
  if (x > half_pi || x < -half_pi) {
    if (x > M_PI || x < -M_PI) {
      x = x - 2 * M_PI * std::floor(x * one_dev_2pi + 0.5);
    }
    if (x < -half_pi) {
      x = -x - M_PI;
      cos_sign = -1.0;
    } else if (x > half_pi) {
      x = -x + M_PI;
      cos_sign = -1.0;
    }
  }
  auto xx = x * x;
  auto xxxx = xx * xx;
  auto xxxxxx = xxxx * xx;
  auto xxxxxxxx = xxxx * xxxx;
  s = x * (1 - xx * coeff3 + xxxx * coeff5 - xxxxxx * coeff7 + xxxxxxxx * coeff9);
  c = cos_sign * (1 - xx * coeff2 + xxxx * coeff4 - xxxxxx * coeff6 + xxxxxxxx * coeff8);
}

enum JointType { J_REVOLUTE = 1, J_CONTINUOUS = 2, J_PRISMATIC = 3 };
enum BaseType { BASE_FIXED = 0, BASE_PLANAR = 1, BASE_FLOATING = 2 };

// ----------------------------------------------------------------------------
// kinematics/kinematics.hpp:50-170, kinematics.cpp:29-223 (tables),
// algorithm.cpp:88-210 (set_joint_angles, lazy FK)
// The URDF XML layer (third/urdf_parser) is not restated: the caller hands over
// the link/joint arrays that init_link_info / init_joint_info /
// init_transform_cache would have produced.
// ----------------------------------------------------------------------------
struct KinModel {
  size_t root_link_id = 0;
  std::vector<size_t> link_parent_link_ids;
  std::vector<uint8_t> link_consider_rotation;
  std::vector<int> joint_types;
  std::vector<V3> joint_axes, joint_positions;
  std::vector<Quat> joint_orientations;
  std::vector<uint8_t> joint_orientation_identity_flags;
  bool all_joint_orientation_identity = true;
  std::vector<int> joint_child_link_ids;
  std::vector<double> joint_angles;
  QuatTrans base_pose;
  mutable std::vector<uint8_t> cache_pred;
  mutable std::vector<QuatTrans> transform_cache;
  mutable std::vector<QuatTrans> tf_plink_to_hlink_cache;

  // kinematics.cpp:147-181 init_transform_cache + :66-144 init_joint_info
  KinModel(int n_links, int root, const int* parent, const double* origin_xyz,
           const double* origin_quat_xyzw, int n_joints, const int* jtypes,
           const int* jchild, const double* jaxes) {
    root_link_id = root;
    link_parent_link_ids.resize(n_links);
    link_consider_rotation.assign(n_links, 1);
    cache_pred.assign(n_links, 0);
    transform_cache.resize(n_links);
    tf_plink_to_hlink_cache.resize(n_links);
    for (int i = 0; i < n_links; i++) {
      link_parent_link_ids[i] =
          parent[i] < 0 ? std::numeric_limits<size_t>::max() : (size_t)parent[i];
      if (parent[i] >= 0) {
        QuatTrans tf;
        tf.quat = {origin_quat_xyzw[4 * i], origin_quat_xyzw[4 * i + 1],
                   origin_quat_xyzw[4 * i + 2], origin_quat_xyzw[4 * i + 3]};
        tf.trans = {origin_xyz[3 * i], origin_xyz[3 * i + 1], origin_xyz[3 * i + 2]};
        tf.is_quat_identity = (std::abs(tf.quat.w - 1.0) < 1e-6);  // kinematics.cpp:176-178
        tf_plink_to_hlink_cache[i] = tf;
      }
    }
    for (int j = 0; j < n_joints; j++) {
      const int c = jchild[j];
      joint_types.push_back(jtypes[j]);
      joint_axes.push_back({jaxes[3 * j], jaxes[3 * j + 1], jaxes[3 * j + 2]});
      joint_positions.push_back(tf_plink_to_hlink_cache[c].trans);
      joint_orientations.push_back(tf_plink_to_hlink_cache[c].quat);
      const bool ident = (std::abs(joint_orientations.back().w - 1.0) < 1e-6);  // :110-113
      joint_orientation_identity_flags.push_back(ident);
      joint_child_link_ids.push_back(c);
      all_joint_orientation_identity = all_joint_orientation_identity && ident;  // :139-144
    }
    joint_angles.assign(n_joints, 0.0);
    base_pose = QuatTrans();
    clear_cache();
  }

  // kinematics.hpp:116-119
  void clear_cache() {
    std::fill(cache_pred.begin(), cache_pred.end(), 0);
    cache_pred[root_link_id] = 1;
    transform_cache[root_link_id].assign(base_pose);
  }
  // kinematics.hpp:111-114
  void set_base_pose(const QuatTrans& pose) {
    base_pose.assign(pose);
    clear_cache();
  }

  // algorithm.cpp:88-158 (the <approx, all_quat_identity> template collapsed
  // into run-time flags; arithmetic identical)
  void set_joint_angles(const std::vector<size_t>& joint_ids, const double* angles,
                        bool high_accuracy) {
    const bool approx = !high_accuracy;
    Quat tf_pjoint_to_hlink_quat;
    for (size_t i = 0; i < joint_ids.size(); i++) {
      auto joint_id = joint_ids[i];
      joint_angles[joint_id] = angles[i];
      auto& tf_plink_to_hlink = tf_plink_to_hlink_cache[joint_child_link_ids[joint_id]];
      auto& tf_plink_to_pjoint_trans = joint_positions[joint_id];
      if (joint_types[joint_id] != J_PRISMATIC) {
        auto x = angles[i] * 0.5;
        double s, c;
        if (approx) {
          compute_approx_sin_cos(x, s, c);
        } else {
          s = sin(x);
          c = cos(x);
        }
        const V3& ax = joint_axes[joint_id];
        tf_pjoint_to_hlink_quat = {s * ax.x, s * ax.y, s * ax.z, c};
        const auto& tf_plink_to_pjoint_quat = joint_orientations[joint_id];
        if (all_joint_orientation_identity) {
          tf_plink_to_hlink.quat = tf_pjoint_to_hlink_quat;
        } else {
          tf_plink_to_hlink.quat = joint_orientation_identity_flags[joint_id]
                                       ? tf_pjoint_to_hlink_quat
                                       : tf_plink_to_pjoint_quat * tf_pjoint_to_hlink_quat;
        }
        tf_plink_to_hlink.trans = tf_plink_to_pjoint_trans;
        tf_plink_to_hlink.is_quat_identity = false;
      } else {
        V3 trans = joint_axes[joint_id] * angles[i];
        tf_plink_to_hlink.trans = tf_plink_to_pjoint_trans + trans;
        tf_plink_to_hlink.quat = Quat();
        tf_plink_to_hlink.is_quat_identity = true;
      }
    }
    clear_cache();
  }

  // kinematics.hpp:131-136
  const QuatTrans& get_link_pose(size_t link_id) const {
    if (!cache_pred[link_id]) build_cache_until(link_id);
    return transform_cache[link_id];
  }
  // algorithm.cpp:160-179
  void build_cache_until(size_t link_id) const {
    if (link_consider_rotation[link_id]) {
      build_cache_until_inner(link_id);
    } else {
      auto plink_id = link_parent_link_ids[link_id];
      if (!cache_pred[plink_id]) build_cache_until_inner(plink_id);
      QuatTrans& tf_rlink_to_plink = transform_cache[plink_id];
      M3 rotmat = eig::to_rotation_matrix(tf_rlink_to_plink.quat);
      auto& plink_to_hlink_trans = tf_plink_to_hlink_cache[link_id].trans;
      V3 pos = tf_rlink_to_plink.trans + rotmat * plink_to_hlink_trans;
      cache_pred[link_id] = 1;
      transform_cache[link_id].trans = pos;
    }
  }
  // algorithm.cpp:181-210
  void build_cache_until_inner(size_t hlink_id) const {
    std::array<size_t, 64> id_stack_like;
    size_t idx = 0;
    while (!cache_pred[hlink_id]) {
      id_stack_like[idx++] = hlink_id;
      hlink_id = link_parent_link_ids[hlink_id];
    }
    size_t hid = hlink_id;
    while (idx > 0) {
      size_t hid_prev = hid;
      hid = id_stack_like[--idx];
      transform_cache[hid_prev].quat_identity_sensitive_mult_and_assign(
          tf_plink_to_hlink_cache[hid], transform_cache[hid]);
      cache_pred[hid] = 1;
    }
  }
  // kinematics.hpp:31-46 + kinematics.cpp:340-363: rptable_[link][joint] is true iff the link lies
  // in the subtree under the joint's child link (restated as a climb to the root)
  bool is_relevant(size_t link_id, size_t joint_id) const {
    const size_t clink = (size_t)joint_child_link_ids[joint_id];
    for (size_t l = link_id; l != std::numeric_limits<size_t>::max(); l = link_parent_link_ids[l])
      if (l == clink) return true;
    return false;
  }
  // algorithm.cpp:346-433 get_attached_point_jacobian: 3 x dim_dof, row-major into jac
  void get_attached_point_jacobian(size_t plink_id, const V3& apoint_global_pos, const std::vector<size_t>& joint_ids,
                                   int base_type, std::vector<double>& jac) const {
    const size_t dim_dof = joint_ids.size() + (base_type == BASE_FLOATING) * 6 + (base_type == BASE_PLANAR) * 3;
    jac.assign(3 * dim_dof, 0.0);
    auto set_col = [&](size_t c, const V3& v) {
      jac[c] = v.x;
      jac[dim_dof + c] = v.y;
      jac[2 * dim_dof + c] = v.z;
    };
    for (size_t i = 0; i < joint_ids.size(); i++) {
      const size_t jid = joint_ids[i];
      if (!is_relevant(plink_id, jid)) continue;
      const QuatTrans& tf_rlink_to_clink = get_link_pose((size_t)joint_child_link_ids[jid]);
      const V3 world_axis = tf_rlink_to_clink.quat * joint_axes[jid];
      V3 dpos;
      if (joint_types[jid] == J_PRISMATIC) {
        dpos = world_axis;
      } else {
        dpos = eig::cross(world_axis, apoint_global_pos - tf_rlink_to_clink.trans);
      }
      set_col(i, dpos);
    }
    QuatTrans tf_rlink_to_elink;  // Identity with trans = the point
    tf_rlink_to_elink.trans = apoint_global_pos;
    QuatTrans tf_rlink_to_blink, tf_blink_to_elink;
    V3 rpy_rlink_to_blink{0, 0, 0};
    if (base_type != BASE_FIXED) {
      tf_rlink_to_blink = get_link_pose(root_link_id);
      tf_blink_to_elink = tf_rlink_to_blink.get_inverse().mul(tf_rlink_to_elink);
      rpy_rlink_to_blink = tf_rlink_to_blink.get_rpy();
    }
    constexpr double eps = 1e-7;
    const size_t n_joint = joint_ids.size();
    auto tweak_col = [&](int rpy_idx, size_t idx_col) {
      double rpy[3] = {rpy_rlink_to_blink.x, rpy_rlink_to_blink.y, rpy_rlink_to_blink.z};
      rpy[rpy_idx] += eps;
      QuatTrans tweaked;
      tweaked.assign(tf_rlink_to_blink);
      tweaked.set_quaternion_from_rpy(rpy[0], rpy[1], rpy[2]);
      const QuatTrans pose_out = tweaked.mul(tf_blink_to_elink);
      const V3 d = pose_out.trans - tf_rlink_to_elink.trans;
      set_col(idx_col, {d.x / eps, d.y / eps, d.z / eps});
    };
    if (base_type == BASE_FLOATING) {
      jac[n_joint + 0] = 1.0;
      jac[dim_dof + n_joint + 1] = 1.0;
      jac[2 * dim_dof + n_joint + 2] = 1.0;
      for (int k = 0; k < 3; k++) tweak_col(k, n_joint + 3 + k);
    } else if (base_type == BASE_PLANAR) {
      jac[n_joint + 0] = 1.0;
      jac[dim_dof + n_joint + 1] = 1.0;
      tweak_col(2, n_joint + 2);
    }
  }
  // kinematics.cpp:300-337 add_new_link(parent, pose, consider_rotation)
  size_t add_new_link(size_t parent_id, const QuatTrans& pose, bool consider_rotation) {
    size_t link_id = link_parent_link_ids.size();
    transform_cache.push_back(QuatTrans());
    cache_pred.push_back(0);
    std::fill(cache_pred.begin(), cache_pred.end(), 0);  // SizedCache::extend clears
    tf_plink_to_hlink_cache.push_back(pose);
    link_parent_link_ids.push_back(parent_id);
    link_consider_rotation.push_back(consider_rotation);
    // NOTE: the reference leaves the root un-cached until the next clear_cache()
    // call; every hot-path entry calls set_joint_angles first, which re-caches it.
    clear_cache();
    return link_id;
  }
};

// ----------------------------------------------------------------------------
// collision/kdtree.hpp:19-47, kdtree.cpp:20-96
// ----------------------------------------------------------------------------
struct KDNode {
  V3 point;
  int axis;
  int left = -1, right = -1;
};
struct KDTree {
  std::vector<KDNode> nodes;
  int root_index = -1;
  explicit KDTree(const std::vector<V3>& points) {
    nodes.reserve(points.size());
    std::vector<V3> copy = points;
    root_index = build(copy.begin(), copy.end(), 0);
  }
  // kdtree.cpp:40-66
  int build(std::vector<V3>::iterator begin, std::vector<V3>::iterator end, int depth) {
    if (begin >= end) return -1;
    int axis = depth % 3;
    auto comparator = [axis](const V3& a, const V3& b) { return a(axis) < b(axis); };
    auto n = std::distance(begin, end);
    auto median_it = begin + n / 2;
    std::nth_element(begin, median_it, end, comparator);
    KDNode nd;
    nd.point = *median_it;
    nd.axis = axis;
    nodes.push_back(nd);
    int node_index = (int)nodes.size() - 1;
    int l = build(begin, median_it, depth + 1);
    nodes[node_index].left = l;
    int r = build(median_it + 1, end, depth + 1);
    nodes[node_index].right = r;
    return node_index;
  }
  // kdtree.cpp:68-96
  void nearest(int node_index, const V3& target, double& best_sqdist, V3& best_point) const {
    if (node_index == -1) return;
    const KDNode& node = nodes[node_index];
    double sqdist = eig::sqnorm(node.point - target);
    if (sqdist < best_sqdist) {
      best_sqdist = sqdist;
      best_point = node.point;
    }
    int axis = node.axis;
    double diff = target(axis) - node.point(axis);
    int first_index = (diff < 0) ? node.left : node.right;
    int second_index = (diff < 0) ? node.right : node.left;
    nearest(first_index, target, best_sqdist, best_point);
    if (diff * diff < best_sqdist) nearest(second_index, target, best_sqdist, best_point);
  }
  // kdtree.cpp:26-38
  V3 query(const V3& target) const {
    V3 best_point;
    double best = std::numeric_limits<double>::max();
    nearest(root_index, target, best, best_point);
    return best_point;
  }
  double sqdist(const V3& target) const {
    V3 best_point;
    double best = std::numeric_limits<double>::max();
    nearest(root_index, target, best, best_point);
    return best;
  }
};

// ----------------------------------------------------------------------------
// collision/primitive_sdf.hpp
// ----------------------------------------------------------------------------
struct Pose {  // :30-75
  V3 position;
  M3 rot, rot_inv;
  bool axis_aligned = true, z_axis_aligned = true;
  Pose() { update_alignment_flags(); }
  Pose(const V3& p, const M3& r) : position(p), rot(r), rot_inv(eig::inverse(r)) {
    update_alignment_flags();
  }
  void translate(const V3& t) { position = position + t; }
  void rotate_z(double angle) {  // :45-53
    double c = std::cos(angle), s = std::sin(angle);
    M3 rz;
    rz.m[0][0] = c; rz.m[0][1] = -s; rz.m[0][2] = 0;
    rz.m[1][0] = s; rz.m[1][1] = c;  rz.m[1][2] = 0;
    rz.m[2][0] = 0; rz.m[2][1] = 0;  rz.m[2][2] = 1;
    rot = rz * rot;
    rot_inv = eig::inverse(rot);
    update_alignment_flags();
  }
  void update_alignment_flags() {  // :57-68
    axis_aligned = eig::is_approx_identity(rot);
    if (axis_aligned) {
      z_axis_aligned = true;
    } else {
      double tol = 1e-6;
      z_axis_aligned = (std::abs(rot(2, 2) - 1.0) < tol) && (std::abs(rot(0, 2)) < tol) &&
                       (std::abs(rot(1, 2)) < tol) && (std::abs(rot(2, 0)) < tol) &&
                       (std::abs(rot(2, 1)) < tol);
    }
  }
  V3 transform_point(const V3& p) const { return rot * p + position; }
};

enum SDFType { T_UNION = 0, T_BOX = 1, T_CYLINDER = 2, T_SPHERE = 3, T_GROUND = 4, T_CLOUD = 5 };

struct SDFBase {  // :79-97
  virtual ~SDFBase() = default;
  virtual SDFType get_type() const = 0;
  virtual std::shared_ptr<SDFBase> clone() const = 0;
  virtual void rotate_z(double angle) = 0;
  virtual void translate(const V3& t) = 0;
  virtual double evaluate(const V3& p) const = 0;
  virtual bool is_outside(const V3& p, double radius) const = 0;
};
using SDFPtr = std::shared_ptr<SDFBase>;

struct UnionSDF : SDFBase {  // :99-173
  std::vector<SDFPtr> sdfs;
  explicit UnionSDF(std::vector<SDFPtr> s) : sdfs(std::move(s)) {}
  SDFType get_type() const override { return T_UNION; }
  SDFPtr clone() const override {
    std::vector<SDFPtr> c;
    for (auto& s : sdfs) c.push_back(s->clone());
    return std::make_shared<UnionSDF>(c);
  }
  void rotate_z(double a) override { for (auto& s : sdfs) s->rotate_z(a); }
  void translate(const V3& t) override { for (auto& s : sdfs) s->translate(t); }
  double evaluate(const V3& p) const override {
    double val = std::numeric_limits<double>::max();
    for (auto& s : sdfs) val = std::min(val, s->evaluate(p));
    return val;
  }
  bool is_outside(const V3& p, double r) const override {
    for (auto& s : sdfs) if (!s->is_outside(p, r)) return false;
    return true;
  }
};

struct PrimitiveSDFBase : SDFBase {  // :175-223
  V3 lb, ub;
  inline bool is_outside_aabb(const V3& p, double radius) const {  // :180-184
    return p(0) < lb(0) - radius || p(0) > ub(0) + radius || p(1) < lb(1) - radius ||
           p(1) > ub(1) + radius || p(2) < lb(2) - radius || p(2) > ub(2) + radius;
  }
};
using PrimPtr = std::shared_ptr<PrimitiveSDFBase>;

struct GroundSDF : PrimitiveSDFBase {  // :225-259
  double height;
  explicit GroundSDF(double h) : height(h) {
    const double inf = std::numeric_limits<double>::infinity();
    lb = {-inf, -inf, 0.0};  // :250-255 (AABB pinned at z = 0 for any height)
    ub = {inf, inf, 0.0};
  }
  SDFType get_type() const override { return T_GROUND; }
  SDFPtr clone() const override { return std::make_shared<GroundSDF>(height); }
  void rotate_z(double) override { throw std::runtime_error("GroundSDF does not support rotation"); }
  void translate(const V3&) override { throw std::runtime_error("GroundSDF does not support translation"); }
  double evaluate(const V3& p) const override { return p(2) + height; }
  bool is_outside(const V3& p, double radius) const override { return p(2) + height > radius; }
};

struct TransformableSDF : PrimitiveSDFBase {  // :261-276
  Pose pose;
  explicit TransformableSDF(const Pose& p) : pose(p) {}
  virtual void update_aabb() = 0;
  void rotate_z(double a) override { pose.rotate_z(a); update_aabb(); }
  void translate(const V3& t) override { pose.translate(t); update_aabb(); }
  void aabb_from_local_box(double hx, double hy, double hz) {  // :400-421, :500-521
    const double inf = std::numeric_limits<double>::infinity();
    lb = {inf, inf, inf};
    ub = {-inf, -inf, -inf};
    for (int k = 0; k < 8; k++) {
      V3 v{(k & 1) ? hx : -hx, (k & 2) ? hy : -hy, (k & 4) ? hz : -hz};
      V3 w = pose.transform_point(v);
      for (int a = 0; a < 3; a++) {
        lb(a) = std::min(lb(a), w(a));
        ub(a) = std::max(ub(a), w(a));
      }
    }
  }
};

struct BoxSDF : TransformableSDF {  // :278-426
  V3 width, half_width;
  BoxSDF(const V3& w, const Pose& p) : TransformableSDF(p), width(w), half_width(w * 0.5) {
    update_aabb();
  }
  SDFType get_type() const override { return T_BOX; }
  SDFPtr clone() const override { return std::make_shared<BoxSDF>(width, pose); }
  void update_aabb() override { aabb_from_local_box(width.x * 0.5, width.y * 0.5, width.z * 0.5); }
  double evaluate(const V3& p) const override {  // :299-311
    V3 d = p - pose.position;
    if (!pose.axis_aligned) d = pose.rot_inv * d;
    V3 sd{std::abs(d.x) - half_width.x, std::abs(d.y) - half_width.y, std::abs(d.z) - half_width.z};
    V3 m{std::max(sd.x, 0.0), std::max(sd.y, 0.0), std::max(sd.z, 0.0)};
    double outside_distance = eig::norm(m);
    double inside_distance =
        std::max(std::max(std::min(sd.x, 0.0), std::min(sd.y, 0.0)), std::min(sd.z, 0.0));
    return outside_distance + inside_distance;
  }
  bool is_outside(const V3& p, double radius) const override {  // :313-398
    double x_signed_dist, y_signed_dist, z_signed_dist;
    if (pose.axis_aligned) {
      x_signed_dist = std::abs(p(0) - pose.position(0)) - half_width(0);
      if (x_signed_dist > radius) return true;
      y_signed_dist = std::abs(p(1) - pose.position(1)) - half_width(1);
      if (y_signed_dist > radius) return true;
      z_signed_dist = std::abs(p(2) - pose.position(2)) - half_width(2);
      if (z_signed_dist > radius) return true;
    } else if (pose.z_axis_aligned) {
      z_signed_dist = std::abs(p(2) - pose.position(2)) - half_width(2);
      if (z_signed_dist > radius) return true;
      V3 pc = p - pose.position;
      x_signed_dist = std::abs(eig::dot(pc, pose.rot.col(0))) - half_width(0);
      if (x_signed_dist > radius) return true;
      y_signed_dist = std::abs(eig::dot(pc, pose.rot.col(1))) - half_width(1);
      if (y_signed_dist > radius) return true;
    } else {
      V3 pc = p - pose.position;
      x_signed_dist = std::abs(eig::dot(pc, pose.rot.col(0))) - half_width(0);
      if (x_signed_dist > radius) return true;
      y_signed_dist = std::abs(eig::dot(pc, pose.rot.col(1))) - half_width(1);
      if (y_signed_dist > radius) return true;
      z_signed_dist = std::abs(eig::dot(pc, pose.rot.col(2))) - half_width(2);
      if (z_signed_dist > radius) return true;
    }
    if (radius < 1e-6) return false;
    bool is_x_out = x_signed_dist > 0, is_y_out = y_signed_dist > 0, is_z_out = z_signed_dist > 0;
    std::uint8_t out_count = is_x_out + is_y_out + is_z_out;
    if (out_count < 2) return false;
    if (out_count == 3)
      return x_signed_dist * x_signed_dist + y_signed_dist * y_signed_dist +
                 z_signed_dist * z_signed_dist > radius * radius;
    if (!is_x_out) return y_signed_dist * y_signed_dist + z_signed_dist * z_signed_dist > radius * radius;
    if (!is_y_out) return x_signed_dist * x_signed_dist + z_signed_dist * z_signed_dist > radius * radius;
    return x_signed_dist * x_signed_dist + y_signed_dist * y_signed_dist > radius * radius;
  }
};

struct CylinderSDF : TransformableSDF {  // :428-528
  double r_cylinder, rsq_cylinder, height, half_height;
  CylinderSDF(double radius, double h, const Pose& p)
      : TransformableSDF(p), r_cylinder(radius), rsq_cylinder(radius * radius), height(h),
        half_height(0.5 * h) {
    update_aabb();
  }
  SDFType get_type() const override { return T_CYLINDER; }
  SDFPtr clone() const override { return std::make_shared<CylinderSDF>(r_cylinder, height, pose); }
  void update_aabb() override { aabb_from_local_box(r_cylinder, r_cylinder, half_height); }
  double evaluate(const V3& p) const override {  // :444-462
    double z_signed_dist, xdot_abs, ydot_abs;
    if (pose.z_axis_aligned) {
      z_signed_dist = std::abs(p(2) - pose.position(2)) - half_height;
      xdot_abs = std::abs(p(0) - pose.position(0));
      ydot_abs = std::abs(p(1) - pose.position(1));
    } else {
      V3 pc = p - pose.position;
      z_signed_dist = std::abs(eig::dot(pc, pose.rot.col(2))) - half_height;
      xdot_abs = std::abs(eig::dot(pc, pose.rot.col(0)));
      ydot_abs = std::abs(eig::dot(pc, pose.rot.col(1)));
    }
    double r_signed_dist = sqrt(xdot_abs * xdot_abs + ydot_abs * ydot_abs) - r_cylinder;
    double ox = std::max(r_signed_dist, 0.0), oz = std::max(z_signed_dist, 0.0);
    double outside_distance = std::sqrt(ox * ox + oz * oz);
    double inside_distance = std::max(std::min(r_signed_dist, 0.0), std::min(z_signed_dist, 0.0));
    return outside_distance + inside_distance;
  }
  bool is_outside(const V3& p, double radius) const override {  // :464-498
    double z_signed_dist, xdot_abs, ydot_abs;
    if (pose.z_axis_aligned) {
      z_signed_dist = std::abs(p(2) - pose.position(2)) - half_height;
      if (z_signed_dist > radius) return true;
      xdot_abs = std::abs(p(0) - pose.position(0));
      ydot_abs = std::abs(p(1) - pose.position(1));
    } else {
      V3 pc = p - pose.position;
      z_signed_dist = std::abs(eig::dot(pc, pose.rot.col(2))) - half_height;
      if (z_signed_dist > radius) return true;
      xdot_abs = std::abs(eig::dot(pc, pose.rot.col(0)));
      ydot_abs = std::abs(eig::dot(pc, pose.rot.col(1)));
    }
    double dist_sq = xdot_abs * xdot_abs + ydot_abs * ydot_abs;
    if (radius < 1e-6) return dist_sq > rsq_cylinder;
    if (dist_sq > (r_cylinder + radius) * (r_cylinder + radius)) return true;
    bool h_out = z_signed_dist > 0;
    bool r_out = dist_sq > rsq_cylinder;
    if (h_out && r_out) {
      double r_signed_dist = sqrt(dist_sq) - r_cylinder;
      return z_signed_dist * z_signed_dist + r_signed_dist * r_signed_dist > radius * radius;
    }
    return false;
  }
};

struct SphereSDF : TransformableSDF {  // :530-566
  double r_sphere, rsq_sphere;
  SphereSDF(double radius, const Pose& p) : TransformableSDF(p), r_sphere(radius), rsq_sphere(radius * radius) {
    update_aabb();
  }
  SDFType get_type() const override { return T_SPHERE; }
  SDFPtr clone() const override { return std::make_shared<SphereSDF>(r_sphere, pose); }
  void update_aabb() override {
    lb = pose.position - V3{r_sphere, r_sphere, r_sphere};
    ub = pose.position + V3{r_sphere, r_sphere, r_sphere};
  }
  double evaluate(const V3& p) const override { return eig::norm(p - pose.position) - r_sphere; }
  bool is_outside(const V3& p, double radius) const override {
    if (radius < 1e-6) return eig::sqnorm(p - pose.position) > rsq_sphere;
    return eig::sqnorm(p - pose.position) > (r_sphere + radius) * (r_sphere + radius);
  }
};

struct CloudSDF : PrimitiveSDFBase {  // :568-612
  std::shared_ptr<KDTree> kdtree;
  double radius_;
  CloudSDF(const std::vector<V3>& points, double radius)
      : kdtree(std::make_shared<KDTree>(points)), radius_(radius) {
    const double inf = std::numeric_limits<double>::infinity();
    lb = {inf, inf, inf};
    ub = {-inf, -inf, -inf};
    for (auto& p : points)
      for (int j = 0; j < 3; j++) {
        lb(j) = std::min(lb(j), p(j));
        ub(j) = std::max(ub(j), p(j));
      }
    lb = lb - V3{radius, radius, radius};
    ub = ub + V3{radius, radius, radius};
  }
  SDFType get_type() const override { return T_CLOUD; }
  SDFPtr clone() const override { throw std::runtime_error("TODO: Not implemented yet"); }
  void rotate_z(double) override { throw std::runtime_error("TODO: Not implemented yet"); }
  void translate(const V3&) override { throw std::runtime_error("TODO: Not implemented yet"); }
  double evaluate(const V3& p) const override { return std::sqrt(kdtree->sqdist(p)) - radius_; }
  bool is_outside(const V3& p, double radius) const override { return this->evaluate(p) > radius; }
};

// ----------------------------------------------------------------------------
// constraints/primitive_sphere_collision.{hpp,cpp}, constraints/primitive.hpp
// ----------------------------------------------------------------------------
struct SphereSpec {  // .hpp:16-21
  size_t parent_link_id;
  std::vector<V3> relative_positions;
  std::vector<double> radii;
  bool only_self_collision;
};

struct SphereGroup {  // .hpp:23-53
  size_t parent_link_id;
  std::vector<double> radii;
  double group_radius;
  bool only_self_collision;
  std::vector<V3> sphere_relative_positions;
  V3 group_sphere_relative_position;
  M3 rot_mat_cache;
  V3 group_sphere_position_cache;
  bool is_group_sphere_position_dirty = true;
  std::vector<V3> sphere_positions_cache;
  bool is_sphere_positions_dirty = true;
  std::vector<size_t> order;  // oracle-only: original index of each (re-ordered) sphere

  void clear_cache() {
    is_group_sphere_position_dirty = true;
    is_sphere_positions_dirty = true;
  }
  // .cpp:26-36
  void create_group_sphere_position_cache(const KinModel& kin) {
    if (!is_group_sphere_position_dirty) return;
    const QuatTrans& plink_pose = kin.get_link_pose(parent_link_id);
    rot_mat_cache = eig::to_rotation_matrix(plink_pose.quat);
    group_sphere_position_cache = rot_mat_cache * group_sphere_relative_position + plink_pose.trans;
    is_group_sphere_position_dirty = false;
  }
  // .cpp:38-51
  void create_sphere_position_cache(const KinModel& kin) {
    if (!is_sphere_positions_dirty) return;
    const V3& plink_trans = kin.transform_cache[parent_link_id].trans;
    for (size_t i = 0; i < sphere_positions_cache.size(); i++)
      sphere_positions_cache[i] = rot_mat_cache * sphere_relative_positions[i] + plink_trans;
    is_sphere_positions_dirty = false;
  }
  // .cpp:53-115 (std::unordered_set iteration order is implementation-defined;
  // the same libstdc++ container is used here on purpose)
  void max_distance_reorder() {
    const size_t n = sphere_relative_positions.size();
    std::vector<size_t> visited;
    std::unordered_set<size_t> unvisited;
    unvisited.reserve(n);
    for (size_t i = 0; i < n; i++) unvisited.insert(i);
    auto first_unvisited = *unvisited.begin();
    visited.push_back(first_unvisited);
    unvisited.erase(first_unvisited);
    for (size_t i = 0; i + 1 < n; i++) {
      double max_dist = 0.0;
      auto it_max_idx = unvisited.begin();
      for (auto it = unvisited.begin(); it != unvisited.end(); it++) {
        double min_dist_to_visited = std::numeric_limits<double>::max();
        size_t idx_unvisited = *it;
        for (size_t idx_visited : visited) {
          double dist = eig::norm(sphere_relative_positions[idx_unvisited] -
                                  sphere_relative_positions[idx_visited]);
          if (dist < min_dist_to_visited) min_dist_to_visited = dist;
        }
        if (min_dist_to_visited > max_dist) {
          max_dist = min_dist_to_visited;
          it_max_idx = it;
        }
      }
      visited.push_back(*it_max_idx);
      unvisited.erase(it_max_idx);
    }
    std::vector<V3> rp(n);
    std::vector<double> rr(n);
    std::vector<size_t> ro(n);
    for (size_t i = 0; i < n; i++) {
      rp[i] = sphere_relative_positions[visited[i]];
      rr[i] = radii[visited[i]];
      ro[i] = order[visited[i]];
    }
    sphere_relative_positions = rp;
    radii = rr;
    order = ro;
    clear_cache();
  }
};

struct SphereCollisionCst {
  std::shared_ptr<KinModel> kin;
  std::vector<size_t> control_joint_ids;
  int base_type;
  std::vector<SphereGroup> sphere_groups;
  std::vector<std::pair<size_t, size_t>> selcol_group_id_pairs;
  SDFPtr fixed_sdf, sdf;
  std::vector<PrimPtr> all_sdfs_cache;
  double cutoff_dist = 0.1;  // .hpp:115

  // .cpp:117-179
  SphereCollisionCst(std::shared_ptr<KinModel> k, std::vector<size_t> cj, int bt,
                     const std::vector<SphereSpec>& specs,
                     const std::vector<std::pair<size_t, size_t>>& pair_parent_links,
                     SDFPtr fixed, bool reorder_spheres)
      : kin(std::move(k)), control_joint_ids(std::move(cj)), base_type(bt), fixed_sdf(fixed) {
    for (size_t i = 0; i < specs.size(); i++) {
      auto& spec = specs[i];
      V3 group_center{0, 0, 0};
      for (auto& p : spec.relative_positions) group_center = group_center + p;
      const double n = (double)spec.relative_positions.size();
      group_center = {group_center.x / n, group_center.y / n, group_center.z / n};
      double max_dist = 0.0;
      for (size_t j = 0; j < spec.relative_positions.size(); j++) {
        double dist = eig::norm(spec.relative_positions[j] - group_center) + spec.radii[j];
        if (dist > max_dist) max_dist = dist;
      }
      SphereGroup g;
      g.parent_link_id = spec.parent_link_id;
      g.radii = spec.radii;
      g.group_radius = max_dist;
      g.only_self_collision = spec.only_self_collision;
      g.sphere_relative_positions = spec.relative_positions;
      g.group_sphere_relative_position = group_center;
      g.sphere_positions_cache.resize(spec.radii.size());
      g.order.resize(spec.radii.size());
      for (size_t j = 0; j < g.order.size(); j++) g.order[j] = j;
      sphere_groups.push_back(g);
      if (reorder_spheres)
        for (auto& group : sphere_groups) group.max_distance_reorder();  // .cpp:155-159
    }
    for (auto& pr : pair_parent_links) {  // .cpp:162-177 (names resolved to link ids by caller)
      auto f = [&](size_t link) {
        for (size_t g = 0; g < sphere_groups.size(); g++)
          if (sphere_groups[g].parent_link_id == link) return (long)g;
        return -1L;
      };
      long g1 = f(pr.first), g2 = f(pr.second);
      if (g1 < 0 || g2 < 0) throw std::runtime_error("(cpp) Invalid pair of link names for self collision");
      selcol_group_id_pairs.push_back({(size_t)g1, (size_t)g2});
    }
    set_all_sdfs();
  }

  // .cpp:426-448
  void set_all_sdfs() {
    all_sdfs_cache.clear();
    if (fixed_sdf) set_all_sdfs_inner(fixed_sdf);
    if (sdf) set_all_sdfs_inner(sdf);
  }
  void set_all_sdfs_inner(const SDFPtr& s) {
    if (s->get_type() == T_UNION) {
      for (auto& sub : std::static_pointer_cast<UnionSDF>(s)->sdfs) set_all_sdfs_inner(sub);
    } else {
      all_sdfs_cache.push_back(std::static_pointer_cast<PrimitiveSDFBase>(s));
    }
  }
  void set_sdf(const SDFPtr& s) {  // .hpp:73-76
    sdf = s;
    set_all_sdfs();
  }
  bool ext_collision_enabled() const { return all_sdfs_cache.size() > 0; }
  bool self_collision_enabled() const { return selcol_group_id_pairs.size() > 0; }
  size_t q_dim() const {  // primitive.hpp:68-72
    return control_joint_ids.size() + (base_type == BASE_FLOATING) * 6 + (base_type == BASE_PLANAR) * 3;
  }

  // constraints/primitive.hpp:38-62
  void update_kintree(const double* q, bool high_accuracy) {
    if (base_type != BASE_FIXED) {
      const size_t n_joint = control_joint_ids.size();
      kin->set_joint_angles(control_joint_ids, q, true);  // default high_accuracy=true (:42)
      QuatTrans pose;
      if (base_type == BASE_FLOATING) {
        pose.trans = {q[n_joint], q[n_joint + 1], q[n_joint + 2]};
        pose.set_quaternion_from_rpy(q[n_joint + 3], q[n_joint + 4], q[n_joint + 5]);
      }
      if (base_type == BASE_PLANAR) {
        pose.trans = {q[n_joint], q[n_joint + 1], 0.0};
        pose.set_quaternion_from_rpy(0, 0, q[n_joint + 2]);
      }
      kin->set_base_pose(pose);
    } else {
      kin->set_joint_angles(control_joint_ids, q, high_accuracy);
    }
  }
  void post_update_kintree() {  // .hpp:67-71
    for (auto& g : sphere_groups) g.clear_cache();
  }

  // primitive.hpp:113-117
  bool is_valid(const double* q) {
    update_kintree(q, false);
    post_update_kintree();
    return is_valid_dirty();
  }
  // .cpp:181-189
  bool is_valid_dirty() {
    if (ext_collision_enabled() && !check_ext_collision()) return false;
    if (self_collision_enabled()) return check_self_collision();
    return true;
  }
  // .cpp:191-220
  bool check_ext_collision() {
    for (auto& group : sphere_groups) {
      if (group.only_self_collision) continue;
      group.create_group_sphere_position_cache(*kin);
      for (auto& s : all_sdfs_cache) {
        if (!s->is_outside_aabb(group.group_sphere_position_cache, group.group_radius)) {
          if (!s->is_outside(group.group_sphere_position_cache, group.group_radius)) {
            group.create_sphere_position_cache(*kin);
            for (size_t i = 0; i < group.radii.size(); i++) {
              if (!s->is_outside_aabb(group.sphere_positions_cache[i], group.radii[i])) {
                if (!s->is_outside(group.sphere_positions_cache[i], group.radii[i])) return false;
              }
            }
          }
        }
      }
    }
    return true;
  }
  // .cpp:222-254
  bool check_self_collision() {
    for (auto& pr : selcol_group_id_pairs) {
      auto& group1 = sphere_groups[pr.first];
      auto& group2 = sphere_groups[pr.second];
      group1.create_group_sphere_position_cache(*kin);
      group2.create_group_sphere_position_cache(*kin);
      double outer_sqdist =
          eig::sqnorm(group1.group_sphere_position_cache - group2.group_sphere_position_cache);
      double outer_r_sum = group1.group_radius + group2.group_radius;
      if (outer_sqdist > outer_r_sum * outer_r_sum) continue;
      group1.create_sphere_position_cache(*kin);
      group2.create_sphere_position_cache(*kin);
      for (size_t i = 0; i < group1.radii.size(); i++) {
        for (size_t j = 0; j < group2.radii.size(); j++) {
          double sqdist = eig::sqnorm(group1.sphere_positions_cache[i] - group2.sphere_positions_cache[j]);
          double r_sum = group1.radii[i] + group2.radii[j];
          if (sqdist < r_sum * r_sum) return false;
        }
      }
    }
    return true;
  }

  // ---- instrumented twin of is_valid: the same control flow as check_ext_collision /
  // check_self_collision above with a counter at every unit of work the reference executes, so that
  // bench.py can report executed vs dense arithmetic.  w[0..1] joints set (revolute+continuous,
  // prismatic), w[2..3] link products built lazily (with rotation, identity-flagged), w[4] group
  // centres placed (incl. toRotationMatrix), w[5] spheres placed, w[6] AABB pre-tests,
  // w[7..11] is_outside calls on box / cylinder / sphere / ground / cloud, w[12] group-pair tests,
  // w[13] sphere-pair tests, w[14] valid.  Returns the validity, which must equal is_valid(q).
  bool count_is_valid(const double* q, uint64_t* w) {
    update_kintree(q, false);
    post_update_kintree();
    for (size_t jid : control_joint_ids) w[kin->joint_types[jid] == J_PRISMATIC ? 1 : 0]++;
    auto place_group = [&](SphereGroup& g) {
      if (g.is_group_sphere_position_dirty) w[4]++;
      g.create_group_sphere_position_cache(*kin);
    };
    auto place_spheres = [&](SphereGroup& g) {
      if (g.is_sphere_positions_dirty) w[5] += g.radii.size();
      g.create_sphere_position_cache(*kin);
    };
    auto outside = [&](const std::shared_ptr<PrimitiveSDFBase>& s, const V3& c, double r) {
      w[6]++;
      if (s->is_outside_aabb(c, r)) return true;
      const int t = (int)s->get_type();
      w[t == T_BOX ? 7 : t == T_CYLINDER ? 8 : t == T_SPHERE ? 9 : t == T_GROUND ? 10 : 11]++;
      return s->is_outside(c, r);
    };
    bool valid = true;
    if (ext_collision_enabled()) {
      for (auto& group : sphere_groups) {
        if (!valid) break;
        if (group.only_self_collision) continue;
        place_group(group);
        for (auto& s : all_sdfs_cache) {
          if (!valid) break;
          if (!outside(s, group.group_sphere_position_cache, group.group_radius)) {
            place_spheres(group);
            for (size_t i = 0; i < group.radii.size(); i++)
              if (!outside(s, group.sphere_positions_cache[i], group.radii[i])) {
                valid = false;
                break;
              }
          }
        }
      }
    }
    if (valid && self_collision_enabled()) {
      for (auto& pr : selcol_group_id_pairs) {
        if (!valid) break;
        auto& group1 = sphere_groups[pr.first];
        auto& group2 = sphere_groups[pr.second];
        place_group(group1);
        place_group(group2);
        w[12]++;
        double outer_sqdist = eig::sqnorm(group1.group_sphere_position_cache - group2.group_sphere_position_cache);
        double outer_r_sum = group1.group_radius + group2.group_radius;
        if (outer_sqdist > outer_r_sum * outer_r_sum) continue;
        place_spheres(group1);
        place_spheres(group2);
        for (size_t i = 0; i < group1.radii.size() && valid; i++)
          for (size_t j = 0; j < group2.radii.size(); j++) {
            w[13]++;
            double sqdist = eig::sqnorm(group1.sphere_positions_cache[i] - group2.sphere_positions_cache[j]);
            double r_sum = group1.radii[i] + group2.radii[j];
            if (sqdist < r_sum * r_sum) {
              valid = false;
              break;
            }
          }
      }
    }
    // lazily built links: everything cached except the root
    for (size_t l = 0; l < kin->cache_pred.size(); l++)
      if (kin->cache_pred[l] && l != kin->root_link_id) w[kin->tf_plink_to_hlink_cache[l].is_quat_identity ? 3 : 2]++;
    w[14] += valid ? 1 : 0;
    return valid;
  }
  // the dense counterpart: every group and sphere placed (no culling), links counted the same way
  void count_dense_fk(const double* q, uint64_t* w) {
    update_kintree(q, false);
    post_update_kintree();
    for (size_t jid : control_joint_ids) w[kin->joint_types[jid] == J_PRISMATIC ? 1 : 0]++;
    for (auto& g : sphere_groups) {
      g.create_group_sphere_position_cache(*kin);
      g.create_sphere_position_cache(*kin);
      w[4]++;
      w[5] += g.radii.size();
    }
    for (size_t l = 0; l < kin->cache_pred.size(); l++)
      if (kin->cache_pred[l] && l != kin->root_link_id) w[kin->tf_plink_to_hlink_cache[l].is_quat_identity ? 3 : 2]++;
  }

  // primitive.hpp:74-79 + .cpp:385-399 (values only; Jacobian rows are out of scope)
  int evaluate(const double* q, double* vals) {
    update_kintree(q, true);
    post_update_kintree();
    int head = 0;
    if (ext_collision_enabled()) vals[head++] = evaluate_ext_collision();
    if (self_collision_enabled()) vals[head++] = evaluate_self_collision();
    return head;
  }
  // .cpp:256-295,297-300,320 (value part)
  double evaluate_ext_collision() {
    double min_val_other = cutoff_dist;
    for (size_t i = 0; i < sphere_groups.size(); i++) {
      auto& group = sphere_groups[i];
      if (group.only_self_collision) continue;
      group.create_group_sphere_position_cache(*kin);
      for (size_t j = 0; j < all_sdfs_cache.size(); j++) {
        auto& s = all_sdfs_cache[j];
        if (!s->is_outside_aabb(group.group_sphere_position_cache, group.group_radius + cutoff_dist)) {
          if (!s->is_outside(group.group_sphere_position_cache, group.group_radius + cutoff_dist)) {
            group.create_sphere_position_cache(*kin);
            for (size_t k = 0; k < group.radii.size(); k++) {
              const V3& c = group.sphere_positions_cache[k];
              if (s->is_outside_aabb(c, group.radii[k] + cutoff_dist)) continue;
              double val = s->evaluate(c) - group.radii[k];
              if (val < min_val_other) min_val_other = val;
            }
          }
        }
      }
    }
    return min_val_other;
  }
  // .cpp:324-361,363-365,381 (value part)
  double evaluate_self_collision() {
    double dist_min = cutoff_dist;
    for (auto& pr : selcol_group_id_pairs) {
      auto& group1 = sphere_groups[pr.first];
      auto& group2 = sphere_groups[pr.second];
      group1.create_group_sphere_position_cache(*kin);
      group2.create_group_sphere_position_cache(*kin);
      double outer_sqdist =
          eig::sqnorm(group1.group_sphere_position_cache - group2.group_sphere_position_cache);
      double m = group1.group_radius + group2.group_radius + cutoff_dist;
      if (outer_sqdist > m * m) continue;
      group1.create_sphere_position_cache(*kin);
      group2.create_sphere_position_cache(*kin);
      for (size_t i = 0; i < group1.radii.size(); i++)
        for (size_t j = 0; j < group2.radii.size(); j++) {
          double dist = eig::norm(group1.sphere_positions_cache[i] - group2.sphere_positions_cache[j]) -
                        (group1.radii[i] + group2.radii[j]);
          if (dist < dist_min) dist_min = dist;
        }
    }
    return dist_min;
  }

  // primitive.hpp:74-79 + .cpp:385-399 WITH the Jacobian rows: jac is [cst_dim][q_dim] row-major
  int evaluate_with_grad(const double* q, double* vals, double* jac) {
    update_kintree(q, true);
    post_update_kintree();
    const size_t dim = q_dim();
    int head = 0;
    if (ext_collision_enabled()) {
      vals[head] = evaluate_ext_collision_grad(jac + head * dim);
      head++;
    }
    if (self_collision_enabled()) {
      vals[head] = evaluate_self_collision_grad(jac + head * dim);
      head++;
    }
    return head;
  }
  // .cpp:256-322
  double evaluate_ext_collision_grad(double* grad_out) {
    const size_t dim = q_dim();
    double min_val_other = cutoff_dist;
    long min_sphere_idx = -1, min_group_idx = -1, min_sdf_idx = -1;
    for (size_t i = 0; i < sphere_groups.size(); i++) {
      auto& group = sphere_groups[i];
      if (group.only_self_collision) continue;
      group.create_group_sphere_position_cache(*kin);
      for (size_t j = 0; j < all_sdfs_cache.size(); j++) {
        auto& s = all_sdfs_cache[j];
        if (!s->is_outside_aabb(group.group_sphere_position_cache, group.group_radius + cutoff_dist)) {
          if (!s->is_outside(group.group_sphere_position_cache, group.group_radius + cutoff_dist)) {
            group.create_sphere_position_cache(*kin);
            for (size_t k = 0; k < group.radii.size(); k++) {
              const V3& c = group.sphere_positions_cache[k];
              if (s->is_outside_aabb(c, group.radii[k] + cutoff_dist)) continue;
              double val = s->evaluate(c) - group.radii[k];
              if (val < min_val_other) {
                min_val_other = val;
                min_group_idx = (long)i;
                min_sdf_idx = (long)j;
                min_sphere_idx = (long)k;
              }
            }
          }
        }
      }
    }
    if (min_sphere_idx < 0) {  // cutoff case
      for (size_t c = 0; c < dim; c++) grad_out[c] = 0.0;
      return min_val_other;
    }
    const V3 min_sphere_trans = sphere_groups[min_group_idx].sphere_positions_cache[min_sphere_idx];
    const double r = sphere_groups[min_group_idx].radii[min_sphere_idx];
    double grad[3];
    for (int i = 0; i < 3; i++) {
      V3 perturbed = min_sphere_trans;
      (i == 0 ? perturbed.x : i == 1 ? perturbed.y : perturbed.z) += 1e-6;
      const double val = all_sdfs_cache[min_sdf_idx]->evaluate(perturbed) - r;
      grad[i] = (val - min_val_other) / 1e-6;
    }
    std::vector<double> J;
    kin->get_attached_point_jacobian(sphere_groups[min_group_idx].parent_link_id, min_sphere_trans, control_joint_ids,
                                     base_type, J);
    for (size_t c = 0; c < dim; c++) grad_out[c] = J[c] * grad[0] + J[dim + c] * grad[1] + J[2 * dim + c] * grad[2];
    return min_val_other;
  }
  // .cpp:324-383
  double evaluate_self_collision_grad(double* grad) {
    const size_t dim = q_dim();
    bool found = false;
    size_t mp[4] = {0, 0, 0, 0};
    double dist_min = cutoff_dist;
    for (auto& pr : selcol_group_id_pairs) {
      auto& group1 = sphere_groups[pr.first];
      auto& group2 = sphere_groups[pr.second];
      group1.create_group_sphere_position_cache(*kin);
      group2.create_group_sphere_position_cache(*kin);
      double outer_sqdist =
          eig::sqnorm(group1.group_sphere_position_cache - group2.group_sphere_position_cache);
      double m = group1.group_radius + group2.group_radius + cutoff_dist;
      if (outer_sqdist > m * m) continue;
      group1.create_sphere_position_cache(*kin);
      group2.create_sphere_position_cache(*kin);
      for (size_t i = 0; i < group1.radii.size(); i++)
        for (size_t j = 0; j < group2.radii.size(); j++) {
          double dist = eig::norm(group1.sphere_positions_cache[i] - group2.sphere_positions_cache[j]) -
                        (group1.radii[i] + group2.radii[j]);
          if (dist < dist_min) {
            dist_min = dist;
            found = true;
            mp[0] = pr.first; mp[1] = i; mp[2] = pr.second; mp[3] = j;
          }
        }
    }
    if (!found) {
      for (size_t c = 0; c < dim; c++) grad[c] = 0.0;
      return dist_min;
    }
    auto& group1 = sphere_groups[mp[0]];
    const V3 center1 = group1.sphere_positions_cache[mp[1]];
    auto& group2 = sphere_groups[mp[2]];
    const V3 center2 = group2.sphere_positions_cache[mp[3]];
    const V3 center_diff = center1 - center2;
    std::vector<double> J1, J2;
    kin->get_attached_point_jacobian(group1.parent_link_id, center1, control_joint_ids, base_type, J1);
    kin->get_attached_point_jacobian(group2.parent_link_id, center2, control_joint_ids, base_type, J2);
    const double norminv = 1.0 / eig::norm(center_diff);
    for (size_t c = 0; c < dim; c++)
      grad[c] = norminv * (center_diff.x * (J1[c] - J2[c]) + center_diff.y * (J1[dim + c] - J2[dim + c]) +
                           center_diff.z * (J1[2 * dim + c] - J2[2 * dim + c]));
    return dist_min;
  }

  // .cpp:412-424 get_all_spheres (after an update_kintree)
  size_t n_spheres() const {
    size_t n = 0;
    for (auto& g : sphere_groups) n += g.radii.size();
    return n;
  }
  void get_all_spheres(double* out4) {
    size_t k = 0;
    for (auto& g : sphere_groups) {
      g.create_group_sphere_position_cache(*kin);
      g.create_sphere_position_cache(*kin);
      for (size_t i = 0; i < g.radii.size(); i++, k++) {
        out4[4 * k] = g.sphere_positions_cache[i].x;
        out4[4 * k + 1] = g.sphere_positions_cache[i].y;
        out4[4 * k + 2] = g.sphere_positions_cache[i].z;
        out4[4 * k + 3] = g.radii[i];
      }
    }
  }
  void get_group_spheres(double* out4) {  // .cpp:401-410
    size_t k = 0;
    for (auto& g : sphere_groups) {
      g.create_group_sphere_position_cache(*kin);
      out4[4 * k] = g.group_sphere_position_cache.x;
      out4[4 * k + 1] = g.group_sphere_position_cache.y;
      out4[4 * k + 2] = g.group_sphere_position_cache.z;
      out4[4 * k + 3] = g.group_radius;
      k++;
    }
  }

  // ORACLE-ONLY (no reference counterpart): the smallest |signed distance - contact threshold|
  // over every sphere/primitive and listed sphere/sphere test is_valid(q) can make, with the
  // same approximate FK.  Parity tests use it to decide whether an item lies in the epsilon
  // band where an FP32 implementation may legitimately disagree (BASELINE.json north_star).
  double validity_margin(const double* q) { return validity_margin_dirty_after(*this, q); }
  // `driver` runs the FK update (the composite's first member, composite_constraint.hpp:39-44)
  double validity_margin_dirty_after(SphereCollisionCst& driver, const double* q) {
    driver.update_kintree(q, false);
    post_update_kintree();
    double m = std::numeric_limits<double>::infinity();
    for (auto& g : sphere_groups) {
      g.create_group_sphere_position_cache(*kin);
      g.create_sphere_position_cache(*kin);
    }
    if (ext_collision_enabled()) {
      for (auto& g : sphere_groups) {
        if (g.only_self_collision) continue;
        for (size_t i = 0; i < g.radii.size(); i++) {
          const V3& c = g.sphere_positions_cache[i];
          const double r = g.radii[i];
          for (auto& s : all_sdfs_cache) {
            m = std::min(m, std::abs(s->evaluate(c) - r));
            if (s->get_type() == T_GROUND) m = std::min(m, std::abs(std::abs(c.z) - r));  // AABB face
          }
        }
      }
    }
    for (auto& pr : selcol_group_id_pairs) {
      auto& g1 = sphere_groups[pr.first];
      auto& g2 = sphere_groups[pr.second];
      for (size_t i = 0; i < g1.radii.size(); i++)
        for (size_t j = 0; j < g2.radii.size(); j++)
          m = std::min(m, std::abs(eig::norm(g1.sphere_positions_cache[i] - g2.sphere_positions_cache[j]) -
                                   (g1.radii[i] + g2.radii[j])));
    }
    return m;
  }

  // Deep copy with a private KinModel: the reference objects are stateful caches
  // and not thread-safe, so a multi-core baseline needs one per thread.
  // Scene SDFs are only read on the hot path and stay shared.
  std::unique_ptr<SphereCollisionCst> clone_for_thread() const {
    auto c = std::make_unique<SphereCollisionCst>(*this);
    c->kin = std::make_shared<KinModel>(*kin);
    return c;
  }
};

// ----------------------------------------------------------------------------
// ompl/generate_sequence.py:11-63 (greedy farthest-first permutation, ties -> larger)
// ----------------------------------------------------------------------------
static std::vector<int> find_farthest_sequence(int N) {
  std::vector<int> result(N, 0);
  std::vector<char> used(N, 0);
  used[0] = 1;
  int current_size = 1;
  for (int i = 1; i < N; i++) {
    long max_min_distance = -1;
    int best_num = -1;
    for (int num = 0; num < N; num++) {
      if (used[num]) continue;
      long min_distance = std::numeric_limits<long>::max();
      for (int j = 0; j < current_size; j++) {
        long distance = std::abs(num - result[j]);
        if (distance < min_distance) min_distance = distance;
      }
      if (min_distance > max_min_distance || (min_distance == max_min_distance && num > best_num)) {
        max_min_distance = min_distance;
        best_num = num;
      }
    }
    result[i] = best_num;
    used[best_num] = 1;
    current_size++;
  }
  return result;
}
static const std::vector<std::vector<int>>& sequence_table() {  // generate_sequence.py:76-98
  static std::vector<std::vector<int>> table = [] {
    std::vector<std::vector<int>> t;
    for (int i = 0; i < 128; i++) {
      auto row = find_farthest_sequence(i + 1);
      row.resize(128, 0);
      t.push_back(row);
    }
    return t;
  }();
  return table;
}

// ----------------------------------------------------------------------------
// ompl/motion_validator.{hpp,cpp}.  OMPL's RealVectorStateSpace::interpolate is
// s1 + (s2 - s1) * t per dimension; si_->distance is the L2 norm.
// ----------------------------------------------------------------------------
struct Validator {
  int type;  // 0 = BOX, 1 = EUCLIDEAN
  std::vector<double> width, inv_width;
  double resolution = 0;
  size_t dim = 0;
  // motion_validator.cpp:89-111
  double box_step_ratio(const double* s1, const double* s2) const {
    double diff_longest_axis = 0;
    double max_diff = -std::numeric_limits<double>::infinity();
    size_t longest_idx = 0;
    for (size_t idx = 0; idx < dim; ++idx) {
      const double diff = s2[idx] - s1[idx];
      const double abs_scaled_diff = std::abs(diff) * inv_width[idx];
      if (abs_scaled_diff > max_diff) {
        max_diff = abs_scaled_diff;
        longest_idx = idx;
        diff_longest_axis = diff;
      }
    }
    if (std::abs(diff_longest_axis) < 1e-6) return std::numeric_limits<double>::infinity();
    return width[longest_idx] / std::abs(diff_longest_axis);
  }
  // motion_validator.cpp:113-126
  double euclid_step_ratio(const double* s1, const double* s2) const {
    double d2 = 0;
    for (size_t i = 0; i < dim; i++) d2 += (s1[i] - s2[i]) * (s1[i] - s2[i]);
    double dist = std::sqrt(d2);
    return resolution / dist;
  }
  double determine_step_ratio(const double* s1, const double* s2) const {
    return type == 0 ? box_step_ratio(s1, s2) : euclid_step_ratio(s1, s2);
  }
  // motion_validator.cpp:17-79
  bool check_motion(SphereCollisionCst& cst, const double* s1, const double* s2, double* s_test,
                    uint64_t* n_calls) const {
    const double step_ratio = determine_step_ratio(s1, s2);
    if (step_ratio == std::numeric_limits<double>::infinity()) return true;
    const auto& SEQUENCE_TABLE = sequence_table();
    size_t n_test = std::floor(1 / step_ratio) + 2;
    auto is_valid = [&](const double* s) {
      if (n_calls) (*n_calls)++;
      return cst.is_valid(s);
    };
    auto interpolate = [&](double t) {
      for (size_t i = 0; i < dim; i++) s_test[i] = s1[i] + (s2[i] - s1[i]) * t;
    };
    if (n_test < SEQUENCE_TABLE.size() + 1) {
      auto& sequence = SEQUENCE_TABLE[n_test - 1];
      if (!is_valid(s2)) return false;
      for (size_t i = 2; i < n_test; i++) {
        double travel_rate = sequence[i] * step_ratio;
        interpolate(travel_rate);
        if (!is_valid(s_test)) return false;
      }
      return true;
    } else {
      if (!is_valid(s2)) return false;
      for (size_t i = 1; i < n_test - 1; i++) {
        double travel_rate = i * step_ratio;
        interpolate(travel_rate);
        if (!is_valid(s_test)) return false;
      }
      return true;
    }
  }
};

template <class F>
static void parallel_for(int64_t n, int n_threads, F&& fn) {
  if (n_threads <= 1) {
    fn(0, 0, n);
    return;
  }
  std::vector<std::thread> th;
  const int64_t chunk = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; t++) {
    const int64_t lo = std::min<int64_t>(n, t * chunk), hi = std::min<int64_t>(n, lo + chunk);
    th.emplace_back([&, t, lo, hi] { fn(t, lo, hi); });
  }
  for (auto& x : th) x.join();
}

}  // namespace orc

// ============================================================================
// C ABI for ctypes (tests / bench cpu_baseline only)
// ============================================================================
using namespace orc;

struct SdfHandle {
  SDFPtr p;
};
static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH(ret)                  \
  }                                     \
  catch (const std::exception& e) {     \
    g_err = e.what();                   \
    return ret;                         \
  }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// ---- kinematics -----------------------------------------------------------
void* orc_kin_create(int n_links, int root, const int* parent, const double* origin_xyz,
                     const double* origin_quat_xyzw, int n_joints, const int* jtypes,
                     const int* jchild, const double* jaxes) {
  return new std::shared_ptr<KinModel>(std::make_shared<KinModel>(
      n_links, root, parent, origin_xyz, origin_quat_xyzw, n_joints, jtypes, jchild, jaxes));
}
void orc_kin_destroy(void* h) { delete (std::shared_ptr<KinModel>*)h; }
void orc_kin_set_joint_angles(void* h, int n, const int* joint_ids, const double* angles, int high_accuracy) {
  auto& k = *(std::shared_ptr<KinModel>*)h;
  std::vector<size_t> ids(joint_ids, joint_ids + n);
  k->set_joint_angles(ids, angles, high_accuracy != 0);
}
void orc_kin_get_joint_angles(void* h, int n, const int* joint_ids, double* out) {
  auto& k = *(std::shared_ptr<KinModel>*)h;
  for (int i = 0; i < n; i++) out[i] = k->joint_angles[joint_ids[i]];
}
void orc_kin_set_base_pose(void* h, const double* xyz_qxyzw) {
  auto& k = *(std::shared_ptr<KinModel>*)h;
  QuatTrans p;
  p.trans = {xyz_qxyzw[0], xyz_qxyzw[1], xyz_qxyzw[2]};
  p.quat = {xyz_qxyzw[3], xyz_qxyzw[4], xyz_qxyzw[5], xyz_qxyzw[6]};
  k->set_base_pose(p);
}
void orc_kin_get_link_pose(void* h, int link_id, double* out7) {
  auto& k = *(std::shared_ptr<KinModel>*)h;
  const QuatTrans& p = k->get_link_pose(link_id);
  out7[0] = p.trans.x; out7[1] = p.trans.y; out7[2] = p.trans.z;
  out7[3] = p.quat.x; out7[4] = p.quat.y; out7[5] = p.quat.z; out7[6] = p.quat.w;
}
int orc_kin_add_new_link(void* h, int parent_id, const double* xyz, const double* rpy, int consider_rotation) {
  auto& k = *(std::shared_ptr<KinModel>*)h;
  QuatTrans pose;  // kinematics.cpp:283-297
  pose.trans = {xyz[0], xyz[1], xyz[2]};
  pose.set_quaternion_from_rpy(rpy[0], rpy[1], rpy[2]);
  return (int)k->add_new_link(parent_id, pose, consider_rotation != 0);
}
void orc_approx_sin_cos(double x, double* s, double* c) { compute_approx_sin_cos(x, *s, *c); }

// ---- psdf -------------------------------------------------------------------
static Pose make_pose(const double* pos, const double* rot9_rowmajor) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = rot9_rowmajor[3 * i + j];
  return Pose({pos[0], pos[1], pos[2]}, r);
}
void* orc_sdf_box(const double* width, const double* pos, const double* rot9) {
  return new SdfHandle{std::make_shared<BoxSDF>(V3{width[0], width[1], width[2]}, make_pose(pos, rot9))};
}
void* orc_sdf_cylinder(double radius, double height, const double* pos, const double* rot9) {
  return new SdfHandle{std::make_shared<CylinderSDF>(radius, height, make_pose(pos, rot9))};
}
void* orc_sdf_sphere(double radius, const double* pos, const double* rot9) {
  return new SdfHandle{std::make_shared<SphereSDF>(radius, make_pose(pos, rot9))};
}
void* orc_sdf_ground(double height) { return new SdfHandle{std::make_shared<GroundSDF>(height)}; }
void* orc_sdf_cloud(const double* pts, int64_t n, double radius) {
  std::vector<V3> v(n);
  for (int64_t i = 0; i < n; i++) v[i] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  return new SdfHandle{std::make_shared<CloudSDF>(v, radius)};
}
void* orc_sdf_union(void** members, int n) {
  std::vector<SDFPtr> v;
  for (int i = 0; i < n; i++) v.push_back(((SdfHandle*)members[i])->p);
  return new SdfHandle{std::make_shared<UnionSDF>(v)};
}
void orc_sdf_destroy(void* h) { delete (SdfHandle*)h; }
void* orc_sdf_clone(void* h) {
  ORC_TRY return (void*)new SdfHandle{((SdfHandle*)h)->p->clone()};
  ORC_CATCH(nullptr)
}
int orc_sdf_translate(void* h, const double* t) {
  ORC_TRY((SdfHandle*)h)->p->translate({t[0], t[1], t[2]});
  return 0;
  ORC_CATCH(-1)
}
int orc_sdf_rotate_z(void* h, double a) {
  ORC_TRY((SdfHandle*)h)->p->rotate_z(a);
  return 0;
  ORC_CATCH(-1)
}
double orc_sdf_evaluate(void* h, const double* p) { return ((SdfHandle*)h)->p->evaluate({p[0], p[1], p[2]}); }
int orc_sdf_is_outside(void* h, const double* p, double r) {
  return ((SdfHandle*)h)->p->is_outside({p[0], p[1], p[2]}, r) ? 1 : 0;
}
void orc_sdf_evaluate_batch(void* h, const double* pts, int64_t n, double* out) {
  auto& s = ((SdfHandle*)h)->p;
  for (int64_t i = 0; i < n; i++) out[i] = s->evaluate({pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]});
}
void orc_sdf_is_outside_batch(void* h, const double* pts, const double* radii, int64_t n, uint8_t* out) {
  auto& s = ((SdfHandle*)h)->p;
  for (int64_t i = 0; i < n; i++)
    out[i] = s->is_outside({pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}, radii[i]) ? 1 : 0;
}
int orc_sdf_aabb(void* h, double* lb, double* ub) {
  auto p = std::dynamic_pointer_cast<PrimitiveSDFBase>(((SdfHandle*)h)->p);
  if (!p) return -1;
  lb[0] = p->lb.x; lb[1] = p->lb.y; lb[2] = p->lb.z;
  ub[0] = p->ub.x; ub[1] = p->ub.y; ub[2] = p->ub.z;
  return 0;
}
int orc_sdf_pose(void* h, double* pos, double* rot9, int* flags) {
  auto p = std::dynamic_pointer_cast<TransformableSDF>(((SdfHandle*)h)->p);
  if (!p) return -1;
  pos[0] = p->pose.position.x; pos[1] = p->pose.position.y; pos[2] = p->pose.position.z;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) rot9[3 * i + j] = p->pose.rot.m[i][j];
  flags[0] = p->pose.axis_aligned;
  flags[1] = p->pose.z_axis_aligned;
  return 0;
}

// ---- kdtree -------------------------------------------------------------------
void* orc_kdtree_create(const double* pts, int64_t n) {
  std::vector<V3> v(n);
  for (int64_t i = 0; i < n; i++) v[i] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  return new KDTree(v);
}
void orc_kdtree_destroy(void* h) { delete (KDTree*)h; }
void orc_kdtree_query(void* h, const double* p, double* out3) {
  V3 r = ((KDTree*)h)->query({p[0], p[1], p[2]});
  out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
double orc_kdtree_sqdist(void* h, const double* p) { return ((KDTree*)h)->sqdist({p[0], p[1], p[2]}); }
void orc_kdtree_sqdist_batch(void* h, const double* pts, int64_t n, double* out) {
  for (int64_t i = 0; i < n; i++) out[i] = ((KDTree*)h)->sqdist({pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]});
}

// ---- SphereCollisionCst ---------------------------------------------------------
// specs are passed flattened: group g owns spheres [group_offsets[g], group_offsets[g+1])
void* orc_cst_create(void* kin_h, int n_ctrl, const int* ctrl_joint_ids, int base_type, int n_groups,
                     const int* group_parent_link, const int* group_offsets, const double* spheres_xyzr,
                     const uint8_t* group_only_self, int n_pairs, const int* pair_links, void* fixed_sdf,
                     int reorder_spheres) {
  ORC_TRY
  auto& k = *(std::shared_ptr<KinModel>*)kin_h;
  std::vector<size_t> cj(ctrl_joint_ids, ctrl_joint_ids + n_ctrl);
  std::vector<SphereSpec> specs;
  for (int g = 0; g < n_groups; g++) {
    SphereSpec s;
    s.parent_link_id = group_parent_link[g];
    s.only_self_collision = group_only_self[g] != 0;
    for (int i = group_offsets[g]; i < group_offsets[g + 1]; i++) {
      s.relative_positions.push_back({spheres_xyzr[4 * i], spheres_xyzr[4 * i + 1], spheres_xyzr[4 * i + 2]});
      s.radii.push_back(spheres_xyzr[4 * i + 3]);
    }
    specs.push_back(s);
  }
  std::vector<std::pair<size_t, size_t>> pairs;
  for (int p = 0; p < n_pairs; p++) pairs.push_back({(size_t)pair_links[2 * p], (size_t)pair_links[2 * p + 1]});
  SDFPtr fixed = fixed_sdf ? ((SdfHandle*)fixed_sdf)->p : nullptr;
  return (void*)new SphereCollisionCst(k, cj, base_type, specs, pairs, fixed, reorder_spheres != 0);
  ORC_CATCH(nullptr)
}
void orc_cst_destroy(void* h) { delete (SphereCollisionCst*)h; }
void orc_cst_set_sdf(void* h, void* sdf) {
  ((SphereCollisionCst*)h)->set_sdf(sdf ? ((SdfHandle*)sdf)->p : nullptr);
}
int orc_cst_q_dim(void* h) { return (int)((SphereCollisionCst*)h)->q_dim(); }
int orc_cst_n_spheres(void* h) { return (int)((SphereCollisionCst*)h)->n_spheres(); }
int orc_cst_n_groups(void* h) { return (int)((SphereCollisionCst*)h)->sphere_groups.size(); }
int orc_cst_n_prims(void* h) { return (int)((SphereCollisionCst*)h)->all_sdfs_cache.size(); }
int orc_cst_is_valid(void* h, const double* q) { return ((SphereCollisionCst*)h)->is_valid(q) ? 1 : 0; }
int orc_cst_evaluate(void* h, const double* q, double* vals) { return ((SphereCollisionCst*)h)->evaluate(q, vals); }
void orc_cst_update_kintree(void* h, const double* q, int high_accuracy) {
  auto* c = (SphereCollisionCst*)h;
  c->update_kintree(q, high_accuracy != 0);
  c->post_update_kintree();
}
void orc_cst_get_all_spheres(void* h, double* out4) { ((SphereCollisionCst*)h)->get_all_spheres(out4); }
void orc_cst_get_group_spheres(void* h, double* out4) { ((SphereCollisionCst*)h)->get_group_spheres(out4); }
// original (pre-reorder) index of every sphere, group by group
void orc_cst_get_sphere_order(void* h, int* out) {
  size_t k = 0;
  for (auto& g : ((SphereCollisionCst*)h)->sphere_groups)
    for (auto o : g.order) out[k++] = (int)o;
}

// Batched drivers (the reference has no batch API: this is a loop over is_valid,
// split over threads with one private kin+cst per thread).
void orc_cst_is_valid_batch(void* h, const double* q, int64_t n, int n_threads, uint8_t* out) {
  auto* base = (SphereCollisionCst*)h;
  const size_t d = base->q_dim();
  parallel_for(n, n_threads, [&](int, int64_t lo, int64_t hi) {
    auto c = base->clone_for_thread();
    for (int64_t i = lo; i < hi; i++) out[i] = c->is_valid(q + i * d) ? 1 : 0;
  });
}
// vals: n x 2 (ext, self); missing entries are NaN.  Optionally all sphere centres (exact FK).
void orc_cst_evaluate_batch(void* h, const double* q, int64_t n, int n_threads, double* vals) {
  auto* base = (SphereCollisionCst*)h;
  const size_t d = base->q_dim();
  parallel_for(n, n_threads, [&](int, int64_t lo, int64_t hi) {
    auto c = base->clone_for_thread();
    for (int64_t i = lo; i < hi; i++) {
      double v[2];
      c->evaluate(q + i * d, v);
      int head = 0;
      vals[2 * i] = c->ext_collision_enabled() ? v[head++] : std::nan("");
      vals[2 * i + 1] = c->self_collision_enabled() ? v[head++] : std::nan("");
    }
  });
}
// vals [n][2] (NaN for a disabled term), jac [n][2][q_dim] (NaN rows for a disabled term)
void orc_cst_evaluate_grad_batch(void* h, const double* q, int64_t n, int n_threads, double* vals, double* jac) {
  auto* base = (SphereCollisionCst*)h;
  const size_t dim = base->q_dim();
  const bool ext = base->ext_collision_enabled(), slf = base->self_collision_enabled();
  parallel_for(n, n_threads, [&](int, int64_t lo, int64_t hi) {
    auto c = base->clone_for_thread();
    std::vector<double> v(2), J(2 * dim);
    for (int64_t i = lo; i < hi; i++) {
      c->evaluate_with_grad(q + i * dim, v.data(), J.data());
      int head = 0;
      for (int t = 0; t < 2; t++) {
        const bool on = t == 0 ? ext : slf;
        vals[2 * i + t] = on ? v[head] : std::nan("");
        for (size_t k = 0; k < dim; k++) jac[(2 * i + t) * dim + k] = on ? J[head * dim + k] : std::nan("");
        if (on) head++;
      }
    }
  });
}

void orc_cst_margin_batch(void* h, const double* q, int64_t n, int n_threads, double* out) {
  auto* base = (SphereCollisionCst*)h;
  const size_t d = base->q_dim();
  parallel_for(n, n_threads, [&](int, int64_t lo, int64_t hi) {
    auto c = base->clone_for_thread();
    for (int64_t i = lo; i < hi; i++) out[i] = c->validity_margin(q + i * d);
  });
}
// SURVEY 8d instrumented counters, summed over the batch: dense = 0 -> the reference's executed work
// (count_is_valid), dense = 1 -> FK + every placement without culling (count_dense_fk).  out[15].
// Returns the number of items whose counted validity differs from is_valid (must be 0).
int64_t orc_cst_work_counts(void* h, const double* q, int64_t n, int dense, uint64_t* out) {
  auto* base = (SphereCollisionCst*)h;
  const size_t d = base->q_dim();
  auto c = base->clone_for_thread();
  for (int k = 0; k < 15; k++) out[k] = 0;
  int64_t diff = 0;
  for (int64_t i = 0; i < n; i++) {
    if (dense) {
      c->count_dense_fk(q + i * d, out);
    } else {
      const bool v = c->count_is_valid(q + i * d, out);
      if (v != c->is_valid(q + i * d)) diff++;
    }
  }
  return diff;
}
void orc_cst_spheres_batch(void* h, const double* q, int64_t n, int high_accuracy, double* out4) {
  auto* base = (SphereCollisionCst*)h;
  const size_t d = base->q_dim(), ns = base->n_spheres();
  auto c = base->clone_for_thread();
  for (int64_t i = 0; i < n; i++) {
    c->update_kintree(q + i * d, high_accuracy != 0);
    c->post_update_kintree();
    c->get_all_spheres(out4 + 4 * ns * i);
  }
}

// ---- motion validator --------------------------------------------------------------
void* orc_validator_create(int type, int dim, const double* box_width, double resolution) {
  auto* v = new Validator();
  v->type = type;
  v->dim = dim;
  v->resolution = resolution;
  if (type == 0)
    for (int i = 0; i < dim; i++) {
      v->width.push_back(box_width[i]);
      v->inv_width.push_back(1.0 / box_width[i]);  // motion_validator.cpp:81-87
    }
  return v;
}
void orc_validator_destroy(void* v) { delete (Validator*)v; }
double orc_validator_step_ratio(void* v, const double* s1, const double* s2) {
  return ((Validator*)v)->determine_step_ratio(s1, s2);
}
int orc_check_motion(void* v, void* cst, const double* s1, const double* s2, uint64_t* n_calls) {
  auto* V = (Validator*)v;
  std::vector<double> tmp(V->dim);
  return V->check_motion(*(SphereCollisionCst*)cst, s1, s2, tmp.data(), n_calls) ? 1 : 0;
}
void orc_check_motion_batch(void* v, void* cst, const double* s1, const double* s2, int64_t n, int n_threads,
                            uint8_t* out, uint64_t* n_calls_total) {
  auto* V = (Validator*)v;
  auto* base = (SphereCollisionCst*)cst;
  const size_t d = V->dim;
  std::vector<uint64_t> calls(std::max(1, n_threads), 0);
  parallel_for(n, n_threads, [&](int t, int64_t lo, int64_t hi) {
    auto c = base->clone_for_thread();
    std::vector<double> tmp(d);
    uint64_t nc = 0;
    for (int64_t i = lo; i < hi; i++)
      out[i] = V->check_motion(*c, s1 + i * d, s2 + i * d, tmp.data(), &nc) ? 1 : 0;
    calls[t] = nc;
  });
  if (n_calls_total) {
    *n_calls_total = 0;
    for (auto c : calls) *n_calls_total += c;
  }
}
// The same walk with what the parity tests need per edge: validity, how many is_valid calls the
// reference's visiting order makes (position of the first failing waypoint, else the waypoint count),
// and — ORACLE-ONLY — the smallest validity_margin over the waypoints visited.  An FP32 implementation
// that walks the same order can differ from the reference on an edge (in validity or in call count)
// only if one of those visited waypoints lies in the contact band.
void orc_check_motion_batch_ex(void* v, void* cst, const double* s1, const double* s2, int64_t n, int n_threads,
                               uint8_t* out, uint32_t* calls_per_edge, double* min_margin_per_edge) {
  auto* V = (Validator*)v;
  auto* base = (SphereCollisionCst*)cst;
  const size_t d = V->dim;
  parallel_for(n, n_threads, [&](int, int64_t lo, int64_t hi) {
    auto c = base->clone_for_thread();
    std::vector<double> tmp(d);
    for (int64_t i = lo; i < hi; i++) {
      const double* a = s1 + i * d;
      const double* b = s2 + i * d;
      uint64_t nc = 0;
      out[i] = V->check_motion(*c, a, b, tmp.data(), &nc) ? 1 : 0;
      if (calls_per_edge) calls_per_edge[i] = (uint32_t)nc;
      if (!min_margin_per_edge) continue;
      // replay the first nc waypoints of motion_validator.cpp:45-78
      double m = std::numeric_limits<double>::infinity();
      if (nc > 0) {
        const double step_ratio = V->determine_step_ratio(a, b);
        const size_t n_test = std::floor(1 / step_ratio) + 2;
        const bool table = n_test < sequence_table().size() + 1;
        m = std::min(m, c->validity_margin(b));
        for (uint64_t k = 1; k < nc; k++) {
          const double t = table ? sequence_table()[n_test - 1][k + 1] * step_ratio : k * step_ratio;
          for (size_t j = 0; j < d; j++) tmp[j] = a[j] + (b[j] - a[j]) * t;
          m = std::min(m, c->validity_margin(tmp.data()));
        }
      }
      min_margin_per_edge[i] = m;
    }
  });
}

// ---- IneqCompositeCst (constraints/composite_constraint.hpp:27-44, 98-105) -----------------------
// members share ONE kinematic model: update_kintree through the first member, post_update_kintree on
// all, then AND of is_valid_dirty (is_valid) or the members' values stacked in order (evaluate)
struct Composite {
  std::vector<std::unique_ptr<SphereCollisionCst>> m;
  static Composite clone_of(void** members, int n) {
    Composite c;
    for (int i = 0; i < n; i++) {
      c.m.push_back(((SphereCollisionCst*)members[i])->clone_for_thread());
      if (i > 0) c.m[i]->kin = c.m[0]->kin;
    }
    return c;
  }
  void update_kintree(const double* q, bool high_accuracy) {  // :39-44
    m.front()->update_kintree(q, high_accuracy);
    for (auto& c : m) c->post_update_kintree();
  }
  bool is_valid(const double* q) {  // :98-105
    update_kintree(q, false);
    for (auto& c : m)
      if (!c->is_valid_dirty()) return false;
    return true;
  }
};
int orc_composite_check(void** members, int n) {  // ctor :27-37: same kinematic chain
  for (int i = 1; i < n; i++)
    if (((SphereCollisionCst*)members[i])->kin != ((SphereCollisionCst*)members[0])->kin) return -1;
  return 0;
}
void orc_composite_is_valid_batch(void** members, int n_members, const double* q, int64_t n, int n_threads, uint8_t* out) {
  const size_t d = ((SphereCollisionCst*)members[0])->q_dim();
  parallel_for(n, n_threads, [&](int, int64_t lo, int64_t hi) {
    Composite c = Composite::clone_of(members, n_members);
    for (int64_t i = lo; i < hi; i++) out[i] = c.is_valid(q + i * d) ? 1 : 0;
  });
}
// ORACLE-ONLY: min over the members of validity_margin under the shared (first member's) FK
void orc_composite_margin_batch(void** members, int n_members, const double* q, int64_t n, int n_threads, double* out) {
  const size_t d = ((SphereCollisionCst*)members[0])->q_dim();
  parallel_for(n, n_threads, [&](int, int64_t lo, int64_t hi) {
    Composite c = Composite::clone_of(members, n_members);
    for (int64_t i = lo; i < hi; i++) {
      double mm = std::numeric_limits<double>::infinity();
      for (auto& k : c.m) mm = std::min(mm, k->validity_margin_dirty_after(*c.m.front(), q + i * d));
      out[i] = mm;
    }
  });
}
// vals [n][2 * n_members]: (ext, self) per member, NaN where a term is disabled (:46-61 stacks only the
// enabled ones; the bridge compacts)
void orc_composite_evaluate_batch(void** members, int n_members, const double* q, int64_t n, int n_threads, double* vals) {
  const size_t d = ((SphereCollisionCst*)members[0])->q_dim();
  parallel_for(n, n_threads, [&](int, int64_t lo, int64_t hi) {
    Composite c = Composite::clone_of(members, n_members);
    for (int64_t i = lo; i < hi; i++) {
      c.update_kintree(q + i * d, true);
      for (int k = 0; k < n_members; k++) {
        auto& mk = *c.m[k];
        vals[(i * n_members + k) * 2] = mk.ext_collision_enabled() ? mk.evaluate_ext_collision() : std::nan("");
        vals[(i * n_members + k) * 2 + 1] = mk.self_collision_enabled() ? mk.evaluate_self_collision() : std::nan("");
      }
    }
  });
}

void orc_sequence_row(int n, int* out) {  // row n-1 of SEQUENCE_TABLE (first n entries)
  auto& t = sequence_table();
  for (int i = 0; i < n; i++) out[i] = t[n - 1][i];
}
void orc_farthest_sequence(int n, int* out) {
  auto r = find_farthest_sequence(n);
  for (int i = 0; i < n; i++) out[i] = r[i];
}

}  // extern "C"
