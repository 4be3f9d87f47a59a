"""Host mirror of ``plainmp.constraint`` for the collision path (reference:
cpp/plainmp/bindings/constraint.cpp:31-35,65-84,108-113).

``SphereCollisionCst`` keeps the reference's constructor and methods; ``is_valid`` / ``evaluate``
run the same device kernels as the additive batch calls (``is_valid_batch``, ``evaluate_batch``,
``get_all_spheres_batch``), so a single configuration is just a batch of one.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import tables
from .kinematics import BaseType, KinematicModel
from .psdf import SDFBase, UnionSDF


class SphereAttachmentSpec:
    """primitive_sphere_collision.hpp:16-21; binding constraint.cpp:65-71."""

    def __init__(self, parent_link_name: str, relative_positions: np.ndarray, radii: np.ndarray, only_self_collision: bool):
        self._parent_link_name = str(parent_link_name)
        self.relative_positions = np.array(relative_positions, dtype=np.float64).reshape(3, -1)
        self.radii = np.array(radii, dtype=np.float64).reshape(-1)
        if self.relative_positions.shape[1] != self.radii.shape[0]:
            raise ValueError("relative_positions must be 3xN with N radii")
        self.only_self_collision = bool(only_self_collision)

    @property
    def parent_link_name(self) -> str:
        return self._parent_link_name


def reorder_spheres_for_informative_collision_check(positions: np.ndarray) -> List[int]:
    """primitive_sphere_collision.cpp:53-96.  The reference starts from
    ``*unordered_set.begin()`` (implementation-defined).  With libstdc++ and consecutive integer
    keys after ``reserve(n)`` that iterates n-1, n-2, ..., 0, so the start element is n-1 and
    ties go to the larger index; this function reproduces that order (it only affects the order
    in which spheres are tested, never a result — SURVEY App. C-9)."""
    n = positions.shape[1]
    visited = [n - 1]
    unvisited = list(range(n - 2, -1, -1))
    for _ in range(n - 1):
        max_dist = 0.0
        best = 0
        for k, idx in enumerate(unvisited):
            d = positions[:, [idx]] - positions[:, visited]
            md = float(np.sqrt((d * d).sum(axis=0)).min())
            if md > max_dist:
                max_dist = md
                best = k
        visited.append(unvisited.pop(best))
    return visited


class IneqConstraintBase:
    pass


class SphereCollisionCst(IneqConstraintBase):
    """``SphereCollisionCst(kin, control_joint_names, base_type, sphere_specs, selfcol_pairs,
    fixed_sdf, reorder_spheres)`` — primitive_sphere_collision.cpp:117-179."""

    def __init__(
        self,
        kin: KinematicModel,
        control_joint_names: Sequence[str],
        base_type: BaseType,
        sphere_specs: Sequence[SphereAttachmentSpec],
        selfcol_pairs: Sequence[Tuple[str, str]],
        fixed_sdf: Optional[SDFBase],
        reorder_spheres: bool = True,
    ):
        self._kin = kin
        self._control_joint_names = list(control_joint_names)
        self._control_joint_ids = kin.get_joint_ids(self._control_joint_names)  # ValueError on unknown names
        self._base_type = BaseType(base_type)
        self._group_link_names: List[str] = []
        self._group_links: List[int] = []
        self._group_spheres: List[np.ndarray] = []
        self._group_only_self: List[bool] = []
        self._group_spheres_unordered: List[np.ndarray] = []  # as given (before the reorder)
        self._reorder_spheres = bool(reorder_spheres)
        for spec in sphere_specs:
            lid = kin.get_link_ids([spec.parent_link_name])[0]
            pos = np.asarray(spec.relative_positions, dtype=np.float64).reshape(3, -1)
            rad = np.asarray(spec.radii, dtype=np.float64).reshape(-1)
            if pos.shape[1] == 0:
                raise RuntimeError("a sphere group needs at least one sphere")
            self._group_spheres_unordered.append(np.concatenate([pos.T, rad[:, None]], axis=1))
            self._group_link_names.append(spec.parent_link_name)
            self._group_links.append(lid)
            self._group_spheres.append(np.concatenate([pos.T, rad[:, None]], axis=1))
            self._group_only_self.append(bool(spec.only_self_collision))
            if reorder_spheres:
                # the reference re-runs the reorder for ALL groups built so far on every
                # iteration of the spec loop (.cpp:155-159); order-only effect, kept for fidelity
                for g in range(len(self._group_spheres)):
                    sp = self._group_spheres[g]
                    order = reorder_spheres_for_informative_collision_check(sp[:, :3].T)
                    self._group_spheres[g] = sp[order]
        self._pairs: List[Tuple[int, int]] = []
        for a, b in selfcol_pairs:  # .cpp:162-177
            if a not in self._group_link_names or b not in self._group_link_names:
                raise RuntimeError("(cpp) Invalid pair of link names for self collision")
            self._pairs.append((self._group_link_names.index(a), self._group_link_names.index(b)))
        self._fixed_sdf = fixed_sdf
        self._sdf: Optional[SDFBase] = None
        self._robot_cache = None  # (kin.state_key, RobotTable)
        self._scene_cache = None  # (state key, SceneTable)

    # ---- tables ----
    def _robot(self) -> tables.RobotTable:
        v = self._kin.state_key(self._control_joint_ids)
        if self._robot_cache is None or self._robot_cache[0] != v:
            self._robot_cache = (
                v,
                tables.RobotTable(
                    self._kin,
                    self._control_joint_ids,
                    self._base_type,
                    self._group_links,
                    self._group_spheres,
                    self._group_only_self,
                    self._pairs,
                ),
            )
        return self._robot_cache[1]

    def _leaves(self):
        # set_all_sdfs: anchors first, then the user SDF, unions expanded (.cpp:426-448)
        out = []
        if self._fixed_sdf is not None:
            out.extend(self._fixed_sdf._leaves())
        if self._sdf is not None:
            out.extend(self._sdf._leaves())
        return out

    def _scene(self) -> Optional[tables.SceneTable]:
        leaves = self._leaves()
        if not leaves:
            return None
        key = tuple(l._state_key() for l in leaves)
        if self._scene_cache is None:
            self._scene_cache = (key, tables.SceneTable(leaves))
        elif self._scene_cache[0] != key:
            # moved primitives / a new cloud frame are written into the resident scene in place, so
            # handles that point at it (validator sessions) follow; a changed primitive list rebuilds
            table = self._scene_cache[1]
            self._scene_cache = (key, table if table.try_update(leaves) else tables.SceneTable(leaves))
        return self._scene_cache[1]

    # ---- binding surface ----
    def set_sdf(self, sdf: Optional[SDFBase]) -> None:
        self._sdf = sdf

    def get_sdf(self) -> Optional[SDFBase]:
        return self._sdf

    def get_kin(self) -> KinematicModel:
        return self._kin

    def get_control_joint_names(self) -> List[str]:
        return list(self._control_joint_names)

    def q_dim(self) -> int:
        return len(self._control_joint_ids) + (6 if self._base_type == BaseType.FLOATING else 0) + (
            3 if self._base_type == BaseType.PLANAR else 0
        )

    def cst_dim(self) -> int:  # .hpp:98-101
        return (1 if self._leaves() else 0) + (1 if self._pairs else 0)

    def update_kintree(self, q: Sequence[float], high_accuracy: bool = True) -> None:
        """Stores q as the model's joint state (the reference also refreshes its FK cache; here
        the cache lives on the device and is rebuilt per call)."""
        q = np.asarray(q, dtype=np.float64).reshape(-1)
        n = len(self._control_joint_ids)
        accurate = True if self._base_type != BaseType.FIXED else bool(high_accuracy)
        self._kin.set_joint_positions(self._control_joint_ids, q[:n], accurate)
        self._last_q = q.copy()
        self._last_accurate = accurate

    def is_valid(self, q: Sequence[float]) -> bool:
        """One configuration, same call as the binding (constraint.cpp:82).  Goes through a cached
        validator session (pinned, mapped staging; one kernel launch whose completion the host sees by
        spinning on the result byte, no torch tensors) so that callers that loop over single states pay
        ~20 us, not a tensor round trip; the session is rebuilt when the robot or scene tables it points at are replaced."""
        from .ompl import ValidatorSession  # late import: ompl imports tables

        robot, scene = self._robot(), self._scene()
        c = getattr(self, "_single_session", None)
        if c is None or c[0] is not robot or c[1] is not scene:
            if c is not None:
                c[2].close()
            c = (robot, scene, ValidatorSession(self, None, capacity=256))
            self._single_session = c
        return bool(c[2].is_valid(np.asarray(q, dtype=np.float64).reshape(-1)))

    def evaluate(self, q: Sequence[float]):
        """``(vals, jac)`` like the binding (constraint.cpp:33): one row per enabled term
        (ext, then self), jac = d vals / d q at the arg-min sphere / sphere pair."""
        d, j = self.evaluate_grad_batch(np.asarray(q, dtype=np.float64).reshape(1, -1))
        on = ~np.isnan(d[0])
        return d[0][on], j[0][on]

    def get_all_spheres(self) -> List[Tuple[np.ndarray, float]]:
        """Sphere centres for the state set by the last ``update_kintree`` (.cpp:412-424)."""
        q = getattr(self, "_last_q", None)
        if q is None:
            raise RuntimeError("call update_kintree(q) first")
        s = self.get_all_spheres_batch(q.reshape(1, -1), accurate=self._last_accurate)[0]
        return [(row[:3].astype(np.float64), float(row[3])) for row in s]

    def get_group_spheres(self) -> List[Tuple[np.ndarray, float]]:
        """Group bounding spheres for the last ``update_kintree`` state (.cpp:401-410)."""
        q = getattr(self, "_last_q", None)
        if q is None:
            raise RuntimeError("call update_kintree(q) first")
        robot = self._robot()
        frames = tables.fk_frames(robot, q.reshape(1, -1), accurate=self._last_accurate)[0].astype(np.float64)
        out = []
        from .kinematics import quat_rotate

        for g in range(len(self._group_links)):
            f = frames[robot.groups[g]["op"]]
            c = quat_rotate(f[3:7], robot.groups[g]["center"].astype(np.float64)) + f[:3]
            out.append((c, float(robot.groups[g]["radius"])))
        return out

    # ---- additive batch surface ----
    def is_valid_batch(self, Q) -> np.ndarray:
        """Validity of every row of ``Q`` ([n, q_dim]); same semantics as ``is_valid`` per row."""
        robot = self._robot()
        scene = self._scene()
        if scene is None and not self._pairs:
            n = Q.shape[0]
            return np.ones(n, dtype=bool)  # App. C-7: nothing to collide with
        return tables.check_configs(robot, scene, Q)

    def evaluate_batch(self, Q) -> np.ndarray:
        """[n, 2] = (ext, self) min signed distances clipped at 0.1; NaN where a term is disabled."""
        return tables.eval_configs(self._robot(), self._scene(), Q)

    def evaluate_grad_batch(self, Q):
        """``evaluate_batch`` plus the Jacobian rows: ([n, 2], [n, 2, q_dim]); a row is zero when the
        term sits at the 0.1 cutoff (primitive_sphere_collision.cpp:297-300, 363-365), NaN when disabled."""
        return tables.eval_grad_configs(self._robot(), self._scene(), Q)

    def get_all_spheres_batch(self, Q, accurate: bool = True) -> np.ndarray:
        """[n, n_spheres, 4] world sphere centres + radii, groups in spec order."""
        return tables.fk_spheres(self._robot(), Q, accurate)


class IneqCompositeCst:
    """composite_constraint.hpp:27-44, 98-105: ONE forward-kinematics update (through the first member),
    then the AND of the members' ``is_valid_dirty``.  Only sphere-collision members are on this repo's
    path; they must share the kinematic model (the reference's constructor check).

    The members are merged into one device table — their sphere groups and listed pairs concatenated
    over a single link program, their primitives concatenated into one scene, and every group masked to
    the primitives of its own member (``pmp_robot_create_ex``) — so that ``is_valid`` /
    ``is_valid_batch`` / edges run ONE kernel with ONE FK pass per configuration, as the reference does
    one ``update_kintree``.  ``evaluate`` stacks the members' rows (composite_constraint.hpp:46-61)."""

    MAX_MERGED_PRIMS = 30  # the direct bits of the kernels' cull mask

    def __init__(self, constraints: Sequence[SphereCollisionCst]):
        self.constraints = list(constraints)
        if not self.constraints:
            raise RuntimeError("IneqCompositeCst needs at least one constraint")
        first = self.constraints[0]
        for c in self.constraints:
            if c.get_kin() is not first.get_kin():
                raise RuntimeError("All constraints must have the same kinematic chain")
        self._merged_cache = None

    # the reference drives the shared model through the FIRST member (update_kintree, :39-44): the other
    # members see whatever joints / base that member controls.  A merged table can only represent
    # members that agree on them, which is how the reference's own callers build composites.
    def _mergeable(self) -> bool:
        first = self.constraints[0]
        same_q = all(c._control_joint_ids == first._control_joint_ids and c._base_type == first._base_type for c in self.constraints)
        n_prims = sum(1 for c in self.constraints for leaf in c._leaves() if not _is_cloud(leaf))
        n_clouds = sum(1 for c in self.constraints for leaf in c._leaves() if _is_cloud(leaf))
        return same_q and n_prims <= self.MAX_MERGED_PRIMS and n_clouds <= 1

    def _merge_parts(self):
        """The members' tables concatenated (host arrays only, no device): group links / spheres /
        only-self flags, pairs re-indexed, the flattened scene leaves in member order and the per-group
        scene mask (bit k < 30 = k-th non-cloud leaf, bit 30 = the cloud)."""
        leaves, links, spheres, only_self, pairs, masks = [], [], [], [], [], []
        for c in self.constraints:
            mask, base_g = 0, len(links)
            for leaf in c._leaves():
                if _is_cloud(leaf):
                    mask |= 1 << 30
                else:
                    mask |= 1 << sum(1 for l in leaves if not _is_cloud(l))
                leaves.append(leaf)
            links += c._group_links
            spheres += c._group_spheres
            only_self += c._group_only_self
            pairs += [(a + base_g, b + base_g) for a, b in c._pairs]
            masks += [mask] * len(c._group_links)
        return {"leaves": leaves, "links": links, "spheres": spheres, "only_self": only_self, "pairs": pairs, "masks": masks}

    def _merged(self):
        """(RobotTable with per-group scene masks, SceneTable or None) for the current member state."""
        first = self.constraints[0]
        m = self._merge_parts()
        leaves, masks = m["leaves"], m["masks"]
        key = (first._kin.state_key(first._control_joint_ids), tuple(id(c) for c in self.constraints), tuple(masks),
               tuple(len(c._group_links) for c in self.constraints))
        skey = tuple(l._state_key() for l in leaves)
        mc = self._merged_cache
        if mc is None or mc["key"] != key:
            robot = tables.RobotTable(first._kin, first._control_joint_ids, first._base_type, m["links"], m["spheres"], m["only_self"],
                                      m["pairs"], group_prim_mask=masks)
            mc = {"key": key, "robot": robot, "skey": None, "scene": None}
        if mc["skey"] != skey:
            if not leaves:
                mc["scene"] = None
            elif mc["scene"] is None or not mc["scene"].try_update(leaves):
                mc["scene"] = tables.SceneTable(leaves)
            mc["skey"] = skey
        self._merged_cache = mc
        return mc["robot"], mc["scene"]

    # the surface the validators / sessions use (same names as SphereCollisionCst)
    def _robot(self):
        return self._merged()[0]

    def _scene(self):
        return self._merged()[1]

    def q_dim(self) -> int:
        return self.constraints[0].q_dim()

    def get_kin(self):
        return self.constraints[0].get_kin()

    def cst_dim(self) -> int:
        return sum(c.cst_dim() for c in self.constraints)

    def is_valid(self, q) -> bool:
        return bool(self.is_valid_batch(np.asarray(q, dtype=np.float64).reshape(1, -1))[0])

    def is_valid_batch(self, Q) -> np.ndarray:
        if self._mergeable():
            robot, scene = self._merged()
            if scene is None and robot.pairs.shape[0] == 0:
                return np.ones(np.asarray(Q).shape[0], dtype=bool)
            return tables.check_configs(robot, scene, Q)
        # members that cannot share one table (different control joints / base, more than 30 primitives or
        # several clouds in total): one launch per member
        out = None
        for c in self.constraints:
            v = c.is_valid_batch(Q)
            out = v if out is None else (out & v)
        return out

    def evaluate(self, q):
        # composite_constraint.hpp:46-61: member values and Jacobian rows stacked in order
        parts = [c.evaluate(q) for c in self.constraints]
        dim = self.constraints[0].q_dim()
        v = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0)
        j = np.concatenate([p[1].reshape(-1, dim) for p in parts], axis=0) if parts else np.zeros((0, dim))
        return v, j

    def evaluate_batch(self, Q) -> np.ndarray:
        """[n, n_members, 2]: (ext, self) of every member, NaN where a term is disabled."""
        return np.stack([c.evaluate_batch(Q) for c in self.constraints], axis=1)


def _is_cloud(leaf) -> bool:
    from . import psdf

    return leaf.get_type() == psdf.T_CLOUD
