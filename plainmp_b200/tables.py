"""Flattening of the host objects into the packed tables the C ABI takes, and the thin device
plumbing (torch tensors as device buffers, ``data_ptr`` into ctypes).

Reference counterparts of what is flattened here (setup-time code, runs once per robot/scene):
  SphereCollisionCst ctor            cpp/plainmp/constraints/primitive_sphere_collision.cpp:117-179
  set_all_sdfs / set_all_sdfs_inner  .cpp:426-448
  KinematicModel tables              cpp/plainmp/kinematics/kinematics.cpp:29-181
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from .kinematics import J_CONTINUOUS, J_FIXED, J_NONE, J_PRISMATIC, J_REVOLUTE, BaseType, KinematicModel


def _torch():
    import torch

    return torch


def _device():
    _capi.require_cuda()
    torch = _torch()
    if not torch.cuda.is_available():
        raise _capi.PmpError("torch reports no CUDA device; plainmp_b200 has no CPU fallback")
    idx = torch.cuda.current_device()
    _capi.check(_capi.load().pmp_set_device(idx))
    return torch.device("cuda", idx)


def _stream_ptr() -> int:
    return _torch().cuda.current_stream().cuda_stream


# ---------------------------------------------------------------------------------------------
# robot
# ---------------------------------------------------------------------------------------------
class LinkProgram:
    """The ancestor-closed, topologically ordered list of link ops for a set of needed links."""

    def __init__(
        self,
        kin: KinematicModel,
        control_joint_ids: Sequence[int],
        needed_links: Sequence[int],
    ):
        links = kin.links
        needed = set()
        for lid in needed_links:
            while lid >= 0 and lid not in needed:
                needed.add(lid)
                lid = links[lid].parent
        root = next(i for i, L in enumerate(links) if L.parent < 0)
        needed.add(root)
        children: Dict[int, List[int]] = {i: [] for i in needed}
        for i in sorted(needed):
            if links[i].parent >= 0:
                children[links[i].parent].append(i)
        # depth-first so that chains stay adjacent (parent == index - 1)
        order: List[int] = []
        stack = [root]
        while stack:
            i = stack.pop()
            order.append(i)
            stack.extend(reversed(children[i]))
        self.link_of_op = order
        self.op_of_link = {lid: k for k, lid in enumerate(order)}
        q_index_of_joint = {jid: k for k, jid in enumerate(control_joint_ids)}
        all_ident = kin.all_joint_orientation_identity
        ops = np.zeros(len(order), dtype=_capi.LINK_OP_DTYPE)
        for k, lid in enumerate(order):
            L = links[lid]
            if not L.consider_rotation:
                # (sphere groups AT such a link were moved to its parent by RobotTable.build_arrays)
                raise NotImplementedError(
                    f"link {L.name}: a consider_rotation=False link cannot have children or be queried for a full pose on the CUDA path"
                )
            op = ops[k]
            op["parent"] = -1 if L.parent < 0 else self.op_of_link[L.parent]
            op["q_index"] = -1
            if L.kind == J_NONE:
                # the root: frame = base pose (identity-flagged product is skipped, exact)
                op["origin_quat"] = (0, 0, 0, 1)
                op["kind"] = _capi.OP_FIXED
                op["flags"] = _capi.OPFLAG_ORIGIN_IDENTITY
            elif L.joint_id >= 0 and L.joint_id in q_index_of_joint:
                op["origin_quat"] = L.origin_quat
                op["origin_xyz"] = L.origin_xyz
                op["axis"] = L.axis
                op["q_index"] = q_index_of_joint[L.joint_id]
                if L.kind == J_PRISMATIC:
                    # algorithm.cpp:149-155: quat = identity (the joint origin's rotation is
                    # dropped and the product skipped), trans = origin + axis * q
                    op["kind"] = _capi.OP_PRISMATIC
                    op["origin_quat"] = (0, 0, 0, 1)
                    op["flags"] = _capi.OPFLAG_ORIGIN_IDENTITY
                else:
                    op["kind"] = _capi.OP_REVOLUTE
                    # algorithm.cpp:140-147
                    if all_ident or abs(L.origin_quat[3] - 1.0) < 1e-6:
                        op["flags"] = _capi.OPFLAG_ORIGIN_IDENTITY
            else:
                # fixed joint, or a movable joint that is not controlled: bake its current
                # local transform
                quat, trans, ident = kin.local_transform(lid)
                op["origin_quat"] = quat
                op["origin_xyz"] = trans
                op["kind"] = _capi.OP_FIXED
                op["flags"] = _capi.OPFLAG_ORIGIN_IDENTITY if ident else 0
        self.ops = ops


class RobotTable:
    """Device-resident robot tables (``pmp_robot_t``)."""

    def __init__(
        self,
        kin: KinematicModel,
        control_joint_ids: Sequence[int],
        base_type: BaseType,
        group_links: Sequence[int],
        group_spheres: Sequence[np.ndarray],  # each (n_i, 4): x, y, z, r in the link frame
        group_only_self: Sequence[bool],
        pairs: Sequence[Tuple[int, int]],
        extra_links: Sequence[int] = (),
        group_prim_mask: Optional[Sequence[int]] = None,
    ):
        """group_prim_mask (optional, one uint32 per group): which scene entries the group is tested
        against (bit k < 30 = primitive k, bit 30 = the cloud) — pmp_robot_create_ex."""
        arrays = RobotTable.build_arrays(kin, control_joint_ids, base_type, group_links, group_spheres, group_only_self, pairs, extra_links)
        self.program = arrays["program"]
        self.groups = arrays["groups"]
        self.spheres = arrays["spheres"]
        self.pairs = arrays["pairs"]
        self.dof = arrays["dof"]
        self.base_type = arrays["base_type"]
        self.q_dim = arrays["q_dim"]
        self.n_spheres = int(self.spheres.shape[0])
        self.n_ops = int(self.program.ops.shape[0])
        self.base_pose = arrays["base_pose"]
        n_groups = len(group_links)

        lib = _capi.load()
        desc = _capi.RobotDesc()
        desc.dof = self.dof
        desc.base_type = int(self.base_type)
        for i in range(7):
            desc.base_pose[i] = float(self.base_pose[i])
        desc.n_ops = self.n_ops
        desc.ops = self.program.ops.ctypes.data
        desc.n_groups = n_groups
        desc.groups = self.groups.ctypes.data
        desc.n_spheres = self.n_spheres
        desc.spheres = self.spheres.ctypes.data
        desc.n_pairs = int(self.pairs.shape[0])
        desc.pairs = self.pairs.ctypes.data
        h = C.c_void_p()
        _device()  # tables are uploaded to torch's current device
        self.group_prim_mask = None
        if group_prim_mask is not None:
            self.group_prim_mask = np.ascontiguousarray(group_prim_mask, dtype=np.uint32)
            if self.group_prim_mask.shape != (n_groups,):
                raise ValueError("group_prim_mask needs one entry per group")
            _capi.check(lib.pmp_robot_create_ex(C.byref(desc), self.group_prim_mask.ctypes.data, C.byref(h)))
        else:
            _capi.check(lib.pmp_robot_create(C.byref(desc), C.byref(h)))
        self._h = h
        self._lib = lib

    @staticmethod
    def build_arrays(kin, control_joint_ids, base_type, group_links, group_spheres, group_only_self, pairs, extra_links=()):
        """The C-ABI tables of a robot as numpy arrays (no device needed): link program, groups with
        their bounding spheres, spheres, pairs."""
        group_links = list(group_links)
        group_spheres = [np.asarray(sp, dtype=np.float64).reshape(-1, 4) for sp in group_spheres]
        for g, lid in enumerate(group_links):
            L = kin.links[lid]
            if L.consider_rotation:
                continue
            # A link added with consider_rotation=False gets only its POSITION from FK (parent position +
            # parent rotation * local offset, algorithm.cpp:164-178); its quaternion is never written.  Spheres
            # AT the link origin are therefore well defined — and identical to spheres of the parent link at
            # the link's local offset, which is how they are flattened.  Offsets from such a link would be
            # rotated by whatever the reference's cache holds (uninitialised memory): refused.
            if np.any(group_spheres[g][:, :3] != 0.0):
                raise NotImplementedError(
                    f"link {L.name}: spheres offset from a consider_rotation=False link are rotated by an unset quaternion in the reference"
                )
            _, trans, _ = kin.local_transform(lid)
            sp = group_spheres[g].copy()
            sp[:, :3] = np.asarray(trans, dtype=np.float64).reshape(1, 3)
            group_spheres[g] = sp
            group_links[g] = L.parent
        program = LinkProgram(kin, control_joint_ids, list(group_links) + list(extra_links))
        n_groups = len(group_links)
        groups = np.zeros(n_groups, dtype=_capi.GROUP_DTYPE)
        flat: List[np.ndarray] = []
        begin = 0
        for g in range(n_groups):
            sp = np.asarray(group_spheres[g], dtype=np.float64).reshape(-1, 4)
            centre, radius = group_bounding_sphere(sp)
            groups[g]["center"] = centre
            groups[g]["radius"] = radius
            groups[g]["op"] = program.op_of_link[group_links[g]]
            groups[g]["sphere_begin"] = begin
            groups[g]["sphere_count"] = sp.shape[0]
            groups[g]["only_self_collision"] = 1 if group_only_self[g] else 0
            begin += sp.shape[0]
            flat.append(sp)
        spheres = (
            np.ascontiguousarray(np.concatenate(flat, axis=0), dtype=np.float32) if flat else np.zeros((0, 4), dtype=np.float32)
        )
        dof = len(control_joint_ids)
        bt = BaseType(base_type)
        return {
            "program": program,
            "groups": groups,
            "spheres": spheres,
            "pairs": np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2)),
            "dof": dof,
            "base_type": bt,
            "q_dim": dof + (3 if bt == BaseType.PLANAR else 0) + (6 if bt == BaseType.FLOATING else 0),
            "base_pose": np.asarray(kin.get_base_pose(), dtype=np.float32),
        }

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.pmp_robot_destroy(h)


class RawTables:
    """pmp_robot_t / pmp_scene_t built directly from C-ABI tables (numpy arrays in the _capi dtypes),
    e.g. the output of include/plainmp_b200_plainmp.hpp run on plainmp's own objects.  Quacks like
    (RobotTable, SceneTable) for check_configs / eval_configs / check_edges."""

    class _Handle:
        def __init__(self, h, lib, destroy, **attrs):
            self._h, self._lib, self._destroy = h, lib, destroy
            self.__dict__.update(attrs)

        @property
        def handle(self):
            return self._h

        def __del__(self):
            h, self._h = getattr(self, "_h", None), None
            if h:
                getattr(self._lib, self._destroy)(h)

    def __init__(self, flat: dict):
        lib = _capi.load()
        _device()
        self._keep = flat
        d = _capi.RobotDesc()
        d.dof, d.base_type = flat["dof"], flat["base_type"]
        for i in range(7):
            d.base_pose[i] = float(flat["base_pose"][i])
        d.n_ops, d.ops = len(flat["ops"]), flat["ops"].ctypes.data
        d.n_groups, d.groups = len(flat["groups"]), flat["groups"].ctypes.data
        d.n_spheres, d.spheres = len(flat["spheres"]), flat["spheres"].ctypes.data
        d.n_pairs, d.pairs = len(flat["pairs"]), flat["pairs"].ctypes.data
        h = C.c_void_p()
        _capi.check(lib.pmp_robot_create(C.byref(d), C.byref(h)))
        q_dim = lib.pmp_robot_q_dim(h)
        self.robot = RawTables._Handle(h, lib, "pmp_robot_destroy", q_dim=q_dim, n_spheres=len(flat["spheres"]),
                                       n_ops=len(flat["ops"]), dof=flat["dof"])
        self.scene = None
        if len(flat["prims"]) or len(flat["cloud_points"]):
            sd = _capi.SceneDesc()
            sd.n_prims = len(flat["prims"])
            sd.prims = flat["prims"].ctypes.data if len(flat["prims"]) else None
            sd.n_cloud_points = len(flat["cloud_points"])
            sd.cloud_points = flat["cloud_points"].ctypes.data if len(flat["cloud_points"]) else None
            sd.cloud_radius = flat["cloud_radius"]
            sd.cloud_cell = 0.0
            self._cloud_radii = flat.get("cloud_point_radii")
            if self._cloud_radii is not None:
                sd.cloud_point_radii = self._cloud_radii.ctypes.data
            hs = C.c_void_p()
            _capi.check(lib.pmp_scene_create(C.byref(sd), C.byref(hs)))
            self.scene = RawTables._Handle(hs, lib, "pmp_scene_destroy")


def group_bounding_sphere(sp: np.ndarray) -> Tuple[np.ndarray, float]:
    """Group centre = mean of member centres, radius = max(|p_j - centre| + r_j)
    (primitive_sphere_collision.cpp:130-144), float64."""
    centre = np.zeros(3)
    for p in sp[:, :3]:
        centre = centre + p
    centre = centre / sp.shape[0]
    radius = 0.0
    for p, r in zip(sp[:, :3], sp[:, 3]):
        d = float(np.sqrt(((p - centre) ** 2).sum())) + float(r)
        if d > radius:
            radius = d
    return centre, radius


def link_pose_via_cuda(kin: KinematicModel, link_name: str) -> np.ndarray:
    """debug_get_link_pose: every joint baked at its stored angle, one FK on the device."""
    lid = kin.get_link_ids([link_name])[0]
    table = RobotTable(kin, [], BaseType.FIXED, [], [], [], [], extra_links=[lid])
    frames = fk_frames(table, np.zeros((1, 0), dtype=np.float32), accurate=True)
    return frames[0, table.program.op_of_link[lid]].astype(np.float64)


# ---------------------------------------------------------------------------------------------
# scene
# ---------------------------------------------------------------------------------------------
def flatten_primitives(leaves) -> Tuple[np.ndarray, Optional[object]]:
    """[anchor primitives..., user primitives...] -> pmp_prim_t array (+ the scene's cloud, if any).

    The reference keeps every CloudSDF leaf as its own entry (primitive_sphere_collision.cpp:436-448).
    All CloudSDF leaves of a scene are merged here into ONE point set served by one device index.  For
    leaves of equal radius min over the leaves of sqrt(NN_k^2) - rho equals sqrt(min_k NN_k^2) - rho; leaves
    of different radii become radius classes of the index (per-point radius, csrc/pmp_cloud.cuh).  A leaf's
    AABB pre-test (primitive_sdf.hpp:573-582) is implied by its distance test, so validity and values are
    the reference's."""
    from . import psdf

    prims = []
    clouds = []
    for leaf in leaves:
        t = leaf.get_type()
        if t == psdf.T_CLOUD:
            clouds.append(leaf)
            continue
        rec = np.zeros((), dtype=_capi.PRIM_DTYPE)
        rec["lb"] = np.maximum(leaf._lb, -3.0e38)
        rec["ub"] = np.minimum(leaf._ub, 3.0e38)
        if t == psdf.T_GROUND:
            rec["type"] = _capi.PRIM_GROUND
            rec["param0"] = leaf.height
            rec["rot_col0"], rec["rot_col1"], rec["rot_col2"] = (1, 0, 0), (0, 1, 0), (0, 0, 1)
            rec["flags"] = _capi.PRIMFLAG_AXIS_ALIGNED | _capi.PRIMFLAG_Z_ALIGNED
        else:
            pose = leaf._pose
            rot = pose._rot
            if not np.allclose(pose._rot_inv, rot.T, atol=1e-6):
                raise ValueError("the CUDA path requires an orthonormal primitive rotation")
            flags = (_capi.PRIMFLAG_AXIS_ALIGNED if pose.axis_aligned else 0) | (
                _capi.PRIMFLAG_Z_ALIGNED if pose.z_axis_aligned else 0
            )
            cols = rot.copy()
            if pose.axis_aligned:
                cols = np.eye(3)  # the reference skips the rotation entirely (primitive_sdf.hpp:328-340)
            rec["position"] = pose._position
            rec["flags"] = flags
            if t == psdf.T_BOX:
                rec["type"] = _capi.PRIM_BOX
                hw = 0.5 * leaf._width
                rec["param0"], rec["param1"], rec["param2"] = hw
            elif t == psdf.T_CYLINDER:
                rec["type"] = _capi.PRIM_CYLINDER
                rec["param0"] = leaf.radius
                rec["param1"] = 0.5 * leaf.height
                if pose.z_axis_aligned:
                    cols = np.eye(3)  # world axes for all three coordinates (primitive_sdf.hpp:446-449,466-472)
            elif t == psdf.T_SPHERE:
                rec["type"] = _capi.PRIM_SPHERE
                rec["param0"] = leaf.radius
            else:
                raise ValueError(f"unsupported SDF type {t}")
            rec["rot_col0"], rec["rot_col1"], rec["rot_col2"] = cols[:, 0], cols[:, 1], cols[:, 2]
        prims.append(rec)
    arr = np.array(prims, dtype=_capi.PRIM_DTYPE) if prims else np.zeros(0, dtype=_capi.PRIM_DTYPE)
    cloud = None
    if len(clouds) == 1:
        cloud = clouds[0]
    elif clouds:
        if len({c.radius for c in clouds}) > 4:
            raise NotImplementedError("more than 4 distinct CloudSDF radii in one scene are not supported on the CUDA path")
        cloud = _MergedCloud(clouds)
    return arr, cloud


class _MergedCloud:
    """Several CloudSDF leaves as one point set (see flatten_primitives); ``point_radii`` is None when all
    leaves have the same radius, else the radius of the leaf each point came from (radius classes of the
    device index, csrc/pmp_cloud.cuh)."""

    def __init__(self, clouds):
        self._members = list(clouds)  # keeps the leaves (and with them their uids) alive
        radii = [c.radius for c in clouds]
        self.radius = max(radii)
        self._uniform = len(set(radii)) == 1
        self._points = None
        self._point_radii = None

    @property
    def points(self) -> np.ndarray:
        if self._points is None:
            self._points = np.ascontiguousarray(np.vstack([c.points for c in self._members]))
        return self._points

    @property
    def point_radii(self) -> Optional[np.ndarray]:
        if self._uniform:
            return None
        if self._point_radii is None:
            self._point_radii = np.ascontiguousarray(
                np.concatenate([np.full(c.points.shape[0], c.radius, dtype=np.float32) for c in self._members])
            )
        return self._point_radii

    def _state_key(self):
        return tuple(c._state_key() for c in self._members)


class SceneTable:
    """Device-resident scene (``pmp_scene_t``): flattened primitive list + optional cloud grid."""

    def __init__(self, leaves, cloud_cell: float = 0.0):
        self.prims, cloud = flatten_primitives(leaves)
        self.n_prims = int(self.prims.shape[0])
        self.has_cloud = cloud is not None
        lib = _capi.load()
        desc = _capi.SceneDesc()
        desc.n_prims = self.n_prims
        desc.prims = self.prims.ctypes.data if self.n_prims else None
        self._cloud_pts = None
        self._cloud_radii = None
        if cloud is not None:
            self._cloud_pts = np.ascontiguousarray(cloud.points, dtype=np.float32)
            desc.n_cloud_points = self._cloud_pts.shape[0]
            desc.cloud_points = self._cloud_pts.ctypes.data
            desc.cloud_radius = cloud.radius
            desc.cloud_cell = cloud_cell
            self._cloud_radii = getattr(cloud, "point_radii", None)
            if self._cloud_radii is not None:
                desc.cloud_point_radii = self._cloud_radii.ctypes.data
        h = C.c_void_p()
        _device()
        _capi.check(lib.pmp_scene_create(C.byref(desc), C.byref(h)))
        self._h = h
        self._lib = lib
        self._cloud_id = cloud._state_key() if cloud is not None else None
        self._cloud_cell = cloud_cell
        self.n_updates = 0

    def try_update(self, leaves) -> bool:
        """Scene mutation in place (pmp_scene_update_prims / pmp_scene_update_cloud): the same
        primitive list with moved poses, and / or a replaced CloudSDF.  Returns False when the
        primitive list itself changed (count or types) — the caller then builds a new table."""
        prims, cloud = flatten_primitives(leaves)
        if prims.shape[0] != self.n_prims or (self.n_prims and not (prims["type"] == self.prims["type"]).all()):
            return False
        if self.n_prims and prims.tobytes() != self.prims.tobytes():
            _capi.check(self._lib.pmp_scene_update_prims(self._h, prims.ctypes.data, self.n_prims, _stream_ptr()))
            _torch().cuda.current_stream().synchronize()  # other streams (sessions) may read the records next
            self.prims = prims
            self.n_updates += 1
        new_id = cloud._state_key() if cloud is not None else None
        if new_id != self._cloud_id:
            if cloud is None:
                _capi.check(self._lib.pmp_scene_update_cloud(self._h, None, 0, 0.0, 0.0))
                self._cloud_pts = None
            else:
                self._cloud_pts = np.ascontiguousarray(cloud.points, dtype=np.float32)
                self._cloud_radii = getattr(cloud, "point_radii", None)
                _capi.check(
                    self._lib.pmp_scene_update_cloud_ex(
                        self._h, self._cloud_pts.ctypes.data, self._cloud_pts.shape[0], cloud.radius,
                        self._cloud_radii.ctypes.data if self._cloud_radii is not None else None, self._cloud_cell,
                    )
                )
            self._cloud_id = new_id
            self.has_cloud = cloud is not None
            self.n_updates += 1
        return True

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.pmp_scene_destroy(h)

    # -- psdf surface --
    def evaluate(self, points_nx3: np.ndarray) -> np.ndarray:
        torch = _torch()
        dev = _device()
        p = torch.from_numpy(np.ascontiguousarray(points_nx3, dtype=np.float32)).to(dev)
        out = torch.empty(p.shape[0], dtype=torch.float32, device=dev)
        _capi.check(self._lib.pmp_sdf_evaluate(self._h, p.data_ptr(), p.shape[0], out.data_ptr(), _stream_ptr()))
        return out.cpu().numpy().astype(np.float64)

    def is_outside(self, points_nx3: np.ndarray, radii: np.ndarray) -> np.ndarray:
        torch = _torch()
        dev = _device()
        p = torch.from_numpy(np.ascontiguousarray(points_nx3, dtype=np.float32)).to(dev)
        r = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(radii, (p.shape[0],)), dtype=np.float32)).to(dev)
        out = torch.empty(p.shape[0], dtype=torch.uint8, device=dev)
        _capi.check(
            self._lib.pmp_sdf_is_outside(self._h, p.data_ptr(), r.data_ptr(), p.shape[0], out.data_ptr(), _stream_ptr())
        )
        return out.cpu().numpy().astype(bool)

    def cloud_nearest(self, points_nx3: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        torch = _torch()
        dev = _device()
        p = torch.from_numpy(np.ascontiguousarray(points_nx3, dtype=np.float32)).to(dev)
        sq = torch.empty(p.shape[0], dtype=torch.float32, device=dev)
        idx = torch.empty(p.shape[0], dtype=torch.int32, device=dev)
        _capi.check(
            self._lib.pmp_cloud_nearest(self._h, p.data_ptr(), p.shape[0], sq.data_ptr(), idx.data_ptr(), _stream_ptr())
        )
        return sq.cpu().numpy().astype(np.float64), idx.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# device calls on resident buffers
# ---------------------------------------------------------------------------------------------
def _as_q(table: RobotTable, q) -> "object":
    """q as a contiguous float32 CUDA tensor [n, q_dim] (numpy or torch in)."""
    torch = _torch()
    dev = _device()
    if isinstance(q, torch.Tensor):
        t = q
        if t.device.type != "cuda":
            t = t.to(dev)
        t = t.to(torch.float32).contiguous()
    else:
        t = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32)).to(dev)
    if t.ndim != 2 or t.shape[1] != table.q_dim:
        raise ValueError(f"q must have shape [n, {table.q_dim}], got {tuple(t.shape)}")
    return t


def check_configs_device(table: RobotTable, scene: Optional[SceneTable], q_dev, bits_dev=None, count_dev=None):
    """Stream-ordered launch on caller-resident tensors; returns (bits tensor, count tensor)."""
    torch = _torch()
    n = q_dev.shape[0]
    if bits_dev is None:
        bits_dev = torch.empty((n + 31) // 32, dtype=torch.int32, device=q_dev.device)
    if count_dev is None:
        count_dev = torch.zeros(1, dtype=torch.int64, device=q_dev.device)
    _capi.check(
        table._lib.pmp_check_configs(
            table.handle,
            scene.handle if scene is not None else None,
            q_dev.data_ptr(),
            n,
            bits_dev.data_ptr(),
            count_dev.data_ptr(),
            _stream_ptr(),
        )
    )
    return bits_dev, count_dev


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    w = np.ascontiguousarray(words).view(np.uint32)
    return np.unpackbits(w.view(np.uint8), bitorder="little")[:n].astype(bool)


def check_configs(table: RobotTable, scene: Optional[SceneTable], q) -> np.ndarray:
    qd = _as_q(table, q)
    bits, _ = check_configs_device(table, scene, qd)
    return unpack_bits(bits.cpu().numpy(), qd.shape[0])


def eval_configs(table: RobotTable, scene: Optional[SceneTable], q) -> np.ndarray:
    torch = _torch()
    qd = _as_q(table, q)
    out = torch.empty((qd.shape[0], 2), dtype=torch.float32, device=qd.device)
    _capi.check(
        table._lib.pmp_eval_configs(
            table.handle, scene.handle if scene is not None else None, qd.data_ptr(), qd.shape[0], out.data_ptr(), _stream_ptr()
        )
    )
    return out.cpu().numpy().astype(np.float64)


def eval_grad_configs(table: RobotTable, scene: Optional[SceneTable], q):
    """(values [n, 2], jacobian rows [n, 2, q_dim]) of evaluate(); NaN where a term is disabled."""
    torch = _torch()
    qd = _as_q(table, q)
    out = torch.empty((qd.shape[0], 2), dtype=torch.float32, device=qd.device)
    jac = torch.empty((qd.shape[0], 2, table.q_dim), dtype=torch.float32, device=qd.device)
    _capi.check(
        table._lib.pmp_eval_grad_configs(
            table.handle, scene.handle if scene is not None else None, qd.data_ptr(), qd.shape[0], out.data_ptr(), jac.data_ptr(), _stream_ptr()
        )
    )
    return out.cpu().numpy().astype(np.float64), jac.cpu().numpy().astype(np.float64)


def fk_spheres(table: RobotTable, q, accurate: bool) -> np.ndarray:
    torch = _torch()
    qd = _as_q(table, q)
    out = torch.empty((qd.shape[0], table.n_spheres, 4), dtype=torch.float32, device=qd.device)
    _capi.check(
        table._lib.pmp_fk_spheres(table.handle, qd.data_ptr(), qd.shape[0], 1 if accurate else 0, out.data_ptr(), _stream_ptr())
    )
    return out.cpu().numpy()


def fk_frames(table: RobotTable, q, accurate: bool) -> np.ndarray:
    torch = _torch()
    qd = _as_q(table, q)
    out = torch.empty((qd.shape[0], table.n_ops, 7), dtype=torch.float32, device=qd.device)
    _capi.check(
        table._lib.pmp_fk_frames(table.handle, qd.data_ptr(), qd.shape[0], 1 if accurate else 0, out.data_ptr(), _stream_ptr())
    )
    return out.cpu().numpy()


def make_validator(vtype: int, dim: int, box_width: Optional[Sequence[float]], resolution: float):
    v = _capi.Validator()
    v.type = vtype
    v.dim = dim
    keep = None
    if box_width is not None:
        keep = np.ascontiguousarray(box_width, dtype=np.float64)
        v.box_width = keep.ctypes.data
    v.resolution = float(resolution)
    return v, keep


def _as_q64(table: RobotTable, q) -> "object":
    """q as a contiguous float64 CUDA tensor [n, q_dim]: the reference's own state type."""
    torch = _torch()
    dev = _device()
    if isinstance(q, torch.Tensor):
        t = q if q.device.type == "cuda" else q.to(dev)
        t = t.to(torch.float64).contiguous()
    else:
        t = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float64)).to(dev)
    if t.ndim != 2 or t.shape[1] != table.q_dim:
        raise ValueError(f"q must have shape [n, {table.q_dim}], got {tuple(t.shape)}")
    return t


def _is_f64(q) -> bool:
    torch = _torch()
    if isinstance(q, torch.Tensor):
        return q.dtype == torch.float64
    return np.asarray(q).dtype == np.float64


def check_edges(table: RobotTable, scene: Optional[SceneTable], validator, q0, q1, with_calls: bool = False):
    """Returns (valid bool[n], n_waypoint_checks) — and, with ``with_calls``, the number of is_valid
    calls per edge (uint32[n]) as a third element.  float64 endpoints (what the reference's callers
    hold) keep their precision for the step ratio and the waypoint count (determine_step_ratio works on
    the caller's doubles, motion_validator.cpp:89-126); float32 endpoints are exact as they are."""
    torch = _torch()
    f64 = _is_f64(q0) or _is_f64(q1)
    a = _as_q64(table, q0) if f64 else _as_q(table, q0)
    b = _as_q64(table, q1) if f64 else _as_q(table, q1)
    if a.shape != b.shape:
        raise ValueError("q0 and q1 must have the same shape")
    n = a.shape[0]
    bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=a.device)
    counters = torch.zeros(2, dtype=torch.int64, device=a.device)
    calls = torch.empty(n, dtype=torch.int32, device=a.device) if with_calls else None
    _capi.check(
        table._lib.pmp_check_edges_ex(
            table.handle,
            scene.handle if scene is not None else None,
            C.byref(validator),
            a.data_ptr(),
            b.data_ptr(),
            1 if f64 else 0,
            n,
            bits.data_ptr(),
            calls.data_ptr() if with_calls else None,
            counters.data_ptr(),
            counters.data_ptr() + 8,
            _stream_ptr(),
        )
    )
    c = counters.cpu().numpy()
    valid = unpack_bits(bits.cpu().numpy(), n)
    if with_calls:
        return valid, int(c[1]), calls.cpu().numpy().view(np.uint32)
    return valid, int(c[1])
