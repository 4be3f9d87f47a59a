/*
 * plainmp_b200.h — C ABI of libplainmp_b200.so (hand-written sm_100a CUDA behind plain C).
 *
 * This is the drop-in boundary for plainmp's data-parallel hot path: validity checking of
 * batches of robot configurations and motion edges.  The reference enters that path one item
 * at a time through pybind11 (paths relative to the plainmp tree, cpp/plainmp/...):
 *
 *   SphereCollisionCst::is_valid(q)        bindings/constraint.cpp:82,
 *                                          constraints/primitive.hpp:113-117,
 *                                          constraints/primitive_sphere_collision.cpp:181-254
 *   ConstraintBase::evaluate(q) (values)   bindings/constraint.cpp:33, primitive.hpp:74-79,
 *                                          primitive_sphere_collision.cpp:256-295,324-361
 *   get_all_spheres / get_group_spheres    bindings/constraint.cpp:83-84, .cpp:401-424
 *   CustomValidatorBase::checkMotion       ompl/motion_validator.cpp:17-79 (Box :89-111,
 *                                          Euclidean :113-126), installed by
 *                                          ompl/ompl_thin_wrap.hpp:80-91
 *   SDFBase::evaluate_batch / is_outside   bindings/primitive.cpp:41-111,
 *                                          collision/primitive_sdf.hpp
 *   KDTree::sqdist / query                 bindings/kdtree.cpp:19-23, collision/kdtree.cpp:26-96
 *   _KinematicModel::debug_get_link_pose   bindings/kinematics.cpp:34,
 *                                          kinematics/kinematic_model_wrapper.cpp:50-54
 *
 * Conventions
 *  - No exceptions cross this boundary: every call returns PMP_OK (0) or a negative status and
 *    pmp_last_error() returns a thread-local message.
 *  - Handles are opaque.  Tables are copied at create time (host pointers in, device-resident
 *    copies kept by the handle).  A handle is immutable after creation and may be used from
 *    several host threads / streams concurrently.
 *  - "_dev" pointers are CUDA device pointers owned by the caller; calls taking them are
 *    stream-ordered on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream)
 *    and do not synchronise.  "_host" entry points take host pointers, perform the
 *    host<->device copies themselves on an internal stream and return after the results are
 *    in the caller's buffers.
 *  - Configurations are row-major float32 [n][q_dim], q_dim = dof (+3 planar: x,y,yaw;
 *    +6 floating: x,y,z,roll,pitch,yaw), exactly the layout of the reference's q vector
 *    (constraints/primitive.hpp:38-72).
 *  - Validity is returned bit-packed: bit (i & 31) of word (i >> 5) is 1 iff item i is valid.
 *  - There is no CPU fallback: if no CUDA device is usable every compute call fails with
 *    PMP_ERR_CUDA.
 */
#ifndef PLAINMP_B200_H_
#define PLAINMP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMP_ABI_VERSION 1

enum {
  PMP_OK = 0,
  PMP_ERR_INVALID = -1, /* bad argument / malformed table */
  PMP_ERR_CUDA = -2,    /* CUDA runtime error, or no usable device */
  PMP_ERR_LIMIT = -3    /* table exceeds a compile-time capacity */
};

/* ---- robot tables -------------------------------------------------------------------- */

enum { PMP_OP_FIXED = 0, PMP_OP_REVOLUTE = 1, PMP_OP_PRISMATIC = 2 };
enum { PMP_BASE_FIXED = 0, PMP_BASE_PLANAR = 1, PMP_BASE_FLOATING = 2 };

/* One frame of the flattened, ancestor-closed link tree, in topological order.
 * frame = parent_frame (x) local, with local = (origin_quat [(x) joint_quat], origin_xyz
 * [+ axis * q]) as in kinematics/algorithm.cpp:109-158 and :181-210. */
typedef struct {
  float origin_quat[4]; /* x, y, z, w of the parent->joint rotation */
  float origin_xyz[3];
  int32_t kind;         /* PMP_OP_* */
  float axis[3];
  int32_t q_index;      /* column of q driving this joint, -1 for PMP_OP_FIXED */
  int32_t parent;       /* index of the parent op, -1 = the base frame */
  int32_t flags;        /* bit0: origin_quat is flagged identity (product skipped,
                           kinematics.cpp:110-113,176-178; transform.hpp:64-74) */
  int32_t reserved[2];
} pmp_link_op_t; /* 64 bytes */

#define PMP_OPFLAG_ORIGIN_IDENTITY 1

/* A sphere group (constraints/primitive_sphere_collision.hpp:23-53). */
typedef struct {
  float center[3]; /* group sphere centre in the link frame (mean of members, .cpp:130-134) */
  float radius;    /* group radius (.cpp:136-144) */
  int32_t op;      /* index into ops: the parent link's frame */
  int32_t sphere_begin, sphere_count;
  int32_t only_self_collision;
} pmp_group_t; /* 32 bytes */

typedef struct {
  int32_t dof;       /* number of joint columns in q */
  int32_t base_type; /* PMP_BASE_* */
  float base_pose[7]; /* x,y,z,qx,qy,qz,qw of the root frame for PMP_BASE_FIXED
                         (KinematicModel::set_base_pose, kinematics.hpp:111-119) */
  int32_t n_ops;
  const pmp_link_op_t* ops;
  int32_t n_groups;
  const pmp_group_t* groups;
  int32_t n_spheres;
  const float* spheres; /* [n_spheres][4]: x, y, z (link frame), radius */
  int32_t n_pairs;
  const int32_t* pairs; /* [n_pairs][2] group indices (.cpp:162-177) */
} pmp_robot_desc_t;

typedef struct pmp_robot pmp_robot_t;
int pmp_robot_create(const pmp_robot_desc_t* desc, pmp_robot_t** out);
/* The same with a collision filter: group_prim_mask[g] (may be NULL = all ones) says which entries of
 * the scene group g is tested against — bit k (k < 30) = primitive k, bit 30 = the scene's point cloud.
 * This is how a composite of several collision constraints on ONE kinematic model, each with its own
 * scene (IneqCompositeCst, constraints/composite_constraint.hpp:27-44,98-105), is checked with a single
 * forward-kinematics pass: the members' groups and pairs are concatenated into one robot table, their
 * primitives into one scene, and every group keeps the bits of its own member's primitives.  Only the
 * validity kernels (check_configs / check_edges / sessions) honour the mask; scenes used with a masked
 * robot must have at most 30 primitives. */
int pmp_robot_create_ex(const pmp_robot_desc_t* desc, const uint32_t* group_prim_mask, pmp_robot_t** out);
void pmp_robot_destroy(pmp_robot_t* robot);
int pmp_robot_q_dim(const pmp_robot_t* robot);
int pmp_robot_n_spheres(const pmp_robot_t* robot);

/* ---- scene tables (collision/primitive_sdf.hpp) ---------------------------------------- */

enum { PMP_PRIM_BOX = 1, PMP_PRIM_CYLINDER = 2, PMP_PRIM_SPHERE = 3, PMP_PRIM_GROUND = 4 };
#define PMP_PRIMFLAG_AXIS_ALIGNED 1   /* Pose::axis_aligned_   (primitive_sdf.hpp:57-68) */
#define PMP_PRIMFLAG_Z_ALIGNED 2      /* Pose::z_axis_aligned_ */

typedef struct {
  float position[3];
  int32_t type;     /* PMP_PRIM_* */
  float rot_col0[3];
  float param0;     /* box: half width x | cylinder, sphere: radius | ground: height */
  float rot_col1[3];
  float param1;     /* box: half width y | cylinder: half height */
  float rot_col2[3];
  float param2;     /* box: half width z */
  float lb[3];      /* AABB (primitive_sdf.hpp:218-219; Ground: z pinned at 0, :250-255) */
  int32_t flags;    /* PMP_PRIMFLAG_* */
  float ub[3];
  int32_t reserved;
} pmp_prim_t; /* 96 bytes */

typedef struct {
  int32_t n_prims;
  const pmp_prim_t* prims; /* flattened union, anchors first (.cpp:426-448) */
  int64_t n_cloud_points;  /* 0 = no CloudSDF in the scene */
  const float* cloud_points; /* [n][3] host pointer */
  float cloud_radius;        /* CloudSDF::radius_ (primitive_sdf.hpp:597-603) */
  float cloud_cell;          /* grid cell edge in metres; <= 0 picks a default */
  /* Several CloudSDF leaves in one scene (the reference keeps every leaf, primitive_sphere_collision.cpp:436-448)
   * are passed as ONE point set.  Leaves of equal radius need nothing else; for leaves of different radii give
   * the radius of the leaf every point came from here ([n] host pointer, at most 4 distinct values; cloud_radius
   * is then ignored).  NULL = every point has cloud_radius. */
  const float* cloud_point_radii;
} pmp_scene_desc_t;

typedef struct pmp_scene pmp_scene_t;
int pmp_scene_create(const pmp_scene_desc_t* desc, pmp_scene_t** out);
void pmp_scene_destroy(pmp_scene_t* scene);

/* Scene mutation.  The reference's constraint holds the SDF objects by pointer, so
 * SDFBase::translate / rotate_z on a leaf (primitive_sdf.hpp:135-145, 265-273) are visible to the next
 * is_valid without re-flattening; these two calls are that path for a device-resident scene.
 * update_prims overwrites the primitive records in place (same count and order as at creation;
 * adding or removing leaves needs a new scene, as adding to a Union after set_sdf needs a new
 * set_sdf upstream, primitive_sphere_collision.cpp:426-448); it is stream-ordered on `stream`.
 * update_cloud replaces the CloudSDF's points (a new depth frame): the Morton sort, box tree and
 * distance field are rebuilt on the device and swapped in after a device synchronise; n = 0 removes
 * the cloud.  Handles that reference the scene (sessions) stay valid. */
int pmp_scene_update_prims(pmp_scene_t* scene, const pmp_prim_t* prims, int32_t n_prims, void* stream);
int pmp_scene_update_cloud(pmp_scene_t* scene, const float* cloud_points, int64_t n_cloud_points,
                           float cloud_radius, float cloud_cell);
/* ... with per-point radii (see pmp_scene_desc_t.cloud_point_radii; NULL = cloud_radius for every point) */
int pmp_scene_update_cloud_ex(pmp_scene_t* scene, const float* cloud_points, int64_t n_cloud_points,
                              float cloud_radius, const float* cloud_point_radii, float cloud_cell);

/* ---- the hot path ------------------------------------------------------------------------ */

/* Batched SphereCollisionCst::is_valid.  scene may be NULL (self-collision only).
 * valid_bits_dev: ceil(n/32) words.  n_valid_dev: optional uint64 counter that the kernel
 * ADDS the number of valid items to (caller zeroes it). */
int pmp_check_configs(const pmp_robot_t* robot, const pmp_scene_t* scene, const float* q_dev,
                      int64_t n, uint32_t* valid_bits_dev, unsigned long long* n_valid_dev,
                      void* stream);

/* Batched ConstraintBase::evaluate, values only: dist_dev[2*i] = ext (min over env spheres x
 * primitives of sdf - r), dist_dev[2*i+1] = self (min over listed sphere pairs), both clipped
 * at the reference's cutoff 0.1 and computed with exact sin/cos FK; a disabled term is NaN. */
int pmp_eval_configs(const pmp_robot_t* robot, const pmp_scene_t* scene, const float* q_dev,
                     int64_t n, float* dist_dev, void* stream);

/* Batched ConstraintBase::evaluate WITH its Jacobian rows (bindings/constraint.cpp:33;
 * primitive_sphere_collision.cpp:297-322, 363-399; point Jacobian kinematics/algorithm.cpp:346-433):
 * dist_dev as above, jac_dev [n][2][q_dim] row-major = d(ext)/dq, d(self)/dq at the arg-min sphere /
 * sphere pair; a row is all zero when the term sits at the 0.1 cutoff and NaN when it is disabled.
 * The SDF gradient and the base-rotation columns are analytic where the reference uses forward
 * differences (steps 1e-6 / 1e-7); the two agree to the reference's own test tolerance (1e-4). */
int pmp_eval_grad_configs(const pmp_robot_t* robot, const pmp_scene_t* scene, const float* q_dev,
                          int64_t n, float* dist_dev, float* jac_dev, void* stream);

/* Batched get_all_spheres after update_kintree(q, accurate): out [n][n_spheres][4]. */
int pmp_fk_spheres(const pmp_robot_t* robot, const float* q_dev, int64_t n, int accurate,
                   float* spheres_dev, void* stream);

/* Link frames after update_kintree(q, accurate): out [n][n_ops][7] (x,y,z,qx,qy,qz,qw). */
int pmp_fk_frames(const pmp_robot_t* robot, const float* q_dev, int64_t n, int accurate,
                  float* frames_dev, void* stream);

/* Motion validator (ompl/motion_validator.cpp). */
enum { PMP_VALIDATOR_BOX = 0, PMP_VALIDATOR_EUCLIDEAN = 1 };
typedef struct {
  int32_t type;           /* PMP_VALIDATOR_* */
  int32_t dim;            /* must equal q_dim (ompl_thin_wrap.hpp:84-86) */
  const double* box_width; /* [dim], BOX only */
  double resolution;       /* EUCLIDEAN only */
} pmp_validator_t;

/* Batched checkMotion(s1, s2): q0_dev/q1_dev are [n][q_dim].  s1 is never checked, s2 and the
 * interior waypoints t = m * step_ratio are (motion_validator.cpp:45-78).
 * n_checked_dev: optional uint64 counter, ADDED the number of waypoint checks executed. */
int pmp_check_edges(const pmp_robot_t* robot, const pmp_scene_t* scene,
                    const pmp_validator_t* validator, const float* q0_dev, const float* q1_dev,
                    int64_t n, uint32_t* valid_bits_dev, unsigned long long* n_valid_dev,
                    unsigned long long* n_checked_dev, void* stream);

/* The same with what the reference's own callers hold and what its tests observe:
 *  - endpoints_f64 != 0: q0_dev / q1_dev are double [n][q_dim] (RealVectorStateSpace states).  The step
 *    ratio and n_test are computed from the doubles exactly as determine_step_ratio does (:89-126), so
 *    the waypoint count never depends on a rounding of the endpoints; the waypoints themselves are
 *    interpolated and checked in FP32.
 *  - calls_per_edge_dev (optional, [n]): how many times is_valid was called for edge i — the position
 *    of the first failing waypoint in the reference's visiting order, else the waypoint count; the sum
 *    is what n_checked_dev receives (si_->isValid calls, ompl_thin_wrap.hpp:114-120). */
int pmp_check_edges_ex(const pmp_robot_t* robot, const pmp_scene_t* scene,
                       const pmp_validator_t* validator, const void* q0_dev, const void* q1_dev,
                       int endpoints_f64, int64_t n, uint32_t* valid_bits_dev,
                       uint32_t* calls_per_edge_dev, unsigned long long* n_valid_dev,
                       unsigned long long* n_checked_dev, void* stream);

/* Host-buffer entry points (what the Python shim / an FFI caller binds): H2D copy of the
 * inputs, kernel, D2H copy of the results, all inside the call. */
int pmp_check_configs_host(const pmp_robot_t* robot, const pmp_scene_t* scene, const float* q_host,
                           int64_t n, uint32_t* valid_bits_host, uint64_t* n_valid);
int pmp_eval_configs_host(const pmp_robot_t* robot, const pmp_scene_t* scene, const float* q_host,
                          int64_t n, float* dist_host);
int pmp_check_edges_host(const pmp_robot_t* robot, const pmp_scene_t* scene,
                         const pmp_validator_t* validator, const float* q0_host,
                         const float* q1_host, int64_t n, uint32_t* valid_bits_host,
                         uint64_t* n_valid, uint64_t* n_checked);

/* FP64 host states, as the reference's bindings hold them (numpy float64 -> Eigen::VectorXd,
 * bindings/bindings.hpp:11-13).  Pageable memory is fine: it is packed to FP32 by a few host threads into pinned
 * staging while the previous chunk is copied and checked, so half the bytes cross PCIe (bound by the host's pack
 * pass, ~1.8e8 panda configurations/s).  A PAGE-LOCKED array (pmp_host_alloc, cudaHostAlloc, cudaHostRegister —
 * e.g. a planner's reused state buffer) is recognised and DMA'd as it is, the doubles narrowed on the device
 * (~7e8 configurations/s). */
int pmp_check_configs_host_f64(const pmp_robot_t* robot, const pmp_scene_t* scene, const double* q_host,
                               int64_t n, uint32_t* valid_bits_host, uint64_t* n_valid);
/* FP64 host endpoints, as bindings hand them over (pybind11 converts numpy to Eigen::VectorXd,
 * bindings/bindings.hpp:11-13): uploaded as doubles, converted on the device. */
int pmp_check_edges_host_f64(const pmp_robot_t* robot, const pmp_scene_t* scene,
                             const pmp_validator_t* validator, const double* q0_host,
                             const double* q1_host, int64_t n, uint32_t* valid_bits_host,
                             uint64_t* n_valid, uint64_t* n_checked);

/* ---- validator sessions: what an OMPL-facing adaptor binds ---------------------
 * The reference installs its validators inside CollisionAwareSpaceInformation
 * (ompl/ompl_thin_wrap.hpp:80-91): isValid(state) -> IneqConstraintBase::is_valid (:114-120) and
 * checkMotion(s1, s2) -> CustomValidatorBase::checkMotion (ompl/motion_validator.cpp:17-79), one
 * state / one edge per call, states held as double[dim] by RealVectorStateSpace.  A session is the
 * batched, low-latency form of that pair: it owns a stream and pinned + device staging sized at
 * creation, takes the caller's FP64 states as they are, and answers with one byte per item.
 * include/plainmp_b200_ompl.hpp wraps it as ompl::base::MotionValidator / StateValidityChecker.
 *  - validator may be NULL for a session that only answers is_valid.
 *  - capacity = items per launch; larger calls are chunked.  Calls on one session are serialised.
 *  - check_motion: the step ratio and n_test are computed on the host in FP64 from the caller's
 *    doubles exactly as determine_step_ratio does (:89-126).  Small batches (up to 2^17 waypoints)
 *    test every waypoint of every edge in one wave (result identical: checkMotion is an AND);
 *    larger ones use the early-exit kernel of pmp_check_edges on the host-computed ratios.  n_checked
 *    is the reference's is_valid call count on both paths (on the first one it is derived from the
 *    per-waypoint results: position of the first failing waypoint in the reference's visiting order). */
typedef struct pmp_session pmp_session_t;
int pmp_session_create(const pmp_robot_t* robot, const pmp_scene_t* scene, const pmp_validator_t* validator,
                       int64_t capacity, pmp_session_t** out);
void pmp_session_destroy(pmp_session_t* session);
/* Swap the scene the session checks against (the handle must outlive the session's use of it). */
int pmp_session_set_scene(pmp_session_t* session, const pmp_scene_t* scene);
int pmp_session_is_valid(pmp_session_t* session, const double* states_host, int64_t n, uint8_t* valid_host);
int pmp_session_check_motion(pmp_session_t* session, const double* s1_host, const double* s2_host, int64_t n,
                             uint8_t* valid_host, uint64_t* n_checked);

/* ---- psdf / kdtree surface ------------------------------------------------------------------ */

/* UnionSDF::evaluate_batch over the scene's primitives and cloud: out_dev[i] = min_P sdf_P(p_i).
 * points_dev is [n][3]. */
int pmp_sdf_evaluate(const pmp_scene_t* scene, const float* points_dev, int64_t n, float* out_dev,
                     void* stream);
/* UnionSDF::is_outside(p_i, r_i) -> out_dev[i] in {0,1}; radii_dev may be NULL (r = 0). */
int pmp_sdf_is_outside(const pmp_scene_t* scene, const float* points_dev, const float* radii_dev,
                       int64_t n, uint8_t* out_dev, void* stream);
/* KDTree::sqdist / query against the scene's cloud: exact nearest neighbour.
 * sqdist_dev [n]; nearest_index_dev optional [n] index into the ORIGINAL point order. */
int pmp_cloud_nearest(const pmp_scene_t* scene, const float* points_dev, int64_t n,
                      float* sqdist_dev, int32_t* nearest_index_dev, void* stream);

/* ---- misc ------------------------------------------------------------------------------------ */
const char* pmp_last_error(void);
int pmp_abi_version(void);
int pmp_device_count(void);
/* Select the CUDA device for the calling host thread (the library carries its own copy of the
 * CUDA runtime, so a host framework's current-device setting is not visible to it). */
int pmp_set_device(int device);
/* FFMA micro-benchmark on the current device: best-of-5 dense FP32 FMA throughput in TFLOP/s
 * (the FP32-SIMT roofline denominator; MEASURED_PEAKS.json has no FP32 figure). */
int pmp_measure_fp32_peak(double* tflops);
/* Number of kernel launches issued by this library since load (all threads). */
uint64_t pmp_launch_count(void);
/* Page-locked host memory for the *_host entry points' inputs (any host pointer works there; pinned
 * buffers let the H2D copies run asynchronously at the link rate).  write_combined != 0 asks for
 * write-combined pages: the copy engine reads them without snooping the CPU caches, which raises the
 * host-side ceiling when several GPUs pull from one socket; the CPU should only WRITE such memory
 * sequentially (reads are uncached).  Free with pmp_host_free. */
int pmp_host_alloc(size_t bytes, int write_combined, void** out);
int pmp_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* PLAINMP_B200_H_ */
