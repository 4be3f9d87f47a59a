"""Robot specs for the collision path (reference: src/plainmp/robot_spec.py:93-162,263-325,
345-492 — only the parts that build the kinematic model and the collision constraint).

The sphere / pair / anchor tables are the reference's own data, imported into
``data/conf/*.json`` by ``tools/import_reference_conf.py``.  The panda / fetch / pr2 URDFs are
not in the reference tree (scikit-robot downloads them at run time); ``data/urdf/*.urdf`` are
kinematic skeletons authored by this repo from the public robot descriptions.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .constraint import SphereAttachmentSpec, SphereCollisionCst
from .kinematics import BaseType, KinematicModel
from .psdf import BoxSDF, CylinderSDF, Pose, SDFBase, SphereSDF, UnionSDF

DATA_DIR = Path(__file__).resolve().parent / "data"


def rpy_matrix(yaw: float, pitch: float, roll: float) -> np.ndarray:
    """Rz(yaw) @ Ry(pitch) @ Rx(roll) — the convention of skrobot.coordinates.math.rpy_matrix,
    which robot_spec.py:287 calls as rpy_matrix(yaw, pitch, roll)."""
    cz, sz = np.cos(yaw), np.sin(yaw)
    cy, sy = np.cos(pitch), np.sin(pitch)
    cx, sx = np.cos(roll), np.sin(roll)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
    ry = np.array([[cy, 0, sy], [0, 1.0, 0], [-sy, 0, cy]])
    rx = np.array([[1.0, 0, 0], [0, cx, -sx], [0, sx, cx]])
    return rz @ ry @ rx


def load_conf_yaml(conf_file: Path) -> dict:
    """A robot configuration in the REFERENCE's own YAML schema, loaded the way RobotSpec.__init__
    does (robot_spec.py:93-114: one level of ``include``, key conflicts are errors) and normalised
    to the dict this module consumes: ``self_collision_group_pairs`` expanded and duplicates
    rejected (:387-411), ``collision_spheres.<link>.spheres`` rows [x, y, z, r] (:345-382),
    ``default_joint_positions`` as a mapping (fetch.yaml lists one-key dicts)."""
    import yaml

    conf_file = Path(conf_file)
    with open(conf_file) as f:
        d = yaml.safe_load(f)
    for rel in d.pop("include", []):
        with open(conf_file.parent / rel) as f:
            dep = yaml.safe_load(f)
        if set(d) & set(dep):
            raise AssertionError("conflict")
        if "include" in dep:
            raise AssertionError("recursive include is not supported yet")
        d.update(dep)
    out = {"control_joint_names": list(d["control_joint_names"])}
    pairs = [list(p) for p in d.get("self_collision_pairs", [])]
    for gp in d.get("self_collision_group_pairs", []):
        pairs += [[l1, l2] for l1 in gp["group1"] for l2 in gp["group2"]]
    seen = set()
    for p in pairs:
        key = tuple(sorted(p))
        if key in seen:
            raise AssertionError(f"duplicated pair {key}")
        seen.add(key)
    out["self_collision_pairs"] = pairs
    out["only_self_collision_links"] = list(d.get("only_self_collision_links", []))
    out["collision_spheres"] = {
        link: [[float(x) for x in row] for row in v["spheres"]] for link, v in d["collision_spheres"].items()
    }
    out["self_body_collision_primitives"] = d.get("self_body_collision_primitives", [])
    djp = d.get("default_joint_positions", {})
    if isinstance(djp, list):
        merged = {}
        for e in djp:
            merged.update(e)
        djp = merged
    out["default_joint_positions"] = {k: float(v) for k, v in djp.items()}
    out["custom_links"] = d.get("custom_links", {})
    for k in ("urdf_path", "robot_name"):
        if k in d:
            out[k] = d[k]
    return out


class RobotSpec:
    conf_name: str = ""
    urdf_name: str = ""

    def __init__(self, base_type: BaseType = BaseType.FIXED, conf_file: Optional[Path] = None, urdf_file: Optional[Path] = None):
        """Stock robots load the imported tables under data/; ``conf_file`` (the reference's YAML
        schema, robot_spec.py:93-114) + ``urdf_file`` describe any other robot."""
        if conf_file is not None:
            self.conf_dict = load_conf_yaml(Path(conf_file))
        else:
            with open(DATA_DIR / "conf" / f"{self.conf_name}.json") as f:
                self.conf_dict = json.load(f)
        self._urdf_file = Path(urdf_file) if urdf_file is not None else None
        self.base_type = BaseType(base_type)
        self._kin: Optional[KinematicModel] = None
        self._csts: Dict[bool, SphereCollisionCst] = {}

    @property
    def urdf_path(self) -> Path:
        if self._urdf_file is not None:  # robot_spec.py:119-122 urdf_path_override
            return self._urdf_file
        return DATA_DIR / "urdf" / f"{self.urdf_name}.urdf"

    def get_kin(self) -> KinematicModel:
        # one shared kinematic model per spec (robot_spec.py:132-162)
        if self._kin is None:
            self._kin = KinematicModel(self.urdf_path.read_text())
            for name, d in self.conf_dict.get("custom_links", {}).items():
                try:
                    self._kin.get_link_ids([name])
                except ValueError:
                    self._kin.add_new_link(name, d["parent_link"], np.array(d["position"]), np.array(d["rpy"]), True)
        return self._kin

    @property
    def control_joint_names(self) -> List[str]:
        return self.conf_dict["control_joint_names"]

    def reflect_joint_positions(self, table: Dict[str, float]) -> None:
        """robot_spec.py:164-175: set non-controlled joints on the shared model."""
        kin = self.get_kin()
        kin.set_joint_positions(kin.get_joint_ids(list(table.keys())), list(table.values()))

    def angle_bounds(self) -> Tuple[np.ndarray, np.ndarray]:
        """URDF limits with +-inf clipped to +-2 pi (robot_spec.py:301-325)."""
        kin = self.get_kin()
        limits = kin.get_joint_position_limits(kin.get_joint_ids(self.control_joint_names))
        lb = np.array([l[0] for l in limits])
        ub = np.array([l[1] for l in limits])
        lb[lb == -np.inf] = -np.pi * 2
        ub[ub == np.inf] = np.pi * 2
        return lb, ub

    def self_body_collision_primitives(self) -> List[SDFBase]:
        """Anchor primitives as psdf objects (robot_spec.py:263-299 + utils.py:25-57)."""
        out: List[SDFBase] = []
        for d in self.conf_dict.get("self_body_collision_primitives", []):
            roll, pitch, yaw = d["rotation"]
            pose = Pose(np.array(d["position"], dtype=np.float64), rpy_matrix(yaw, pitch, roll))
            if d["type"] == "box":
                out.append(BoxSDF(d["extents"], pose))
            elif d["type"] == "sphere":
                out.append(SphereSDF(d["radius"], pose))
            elif d["type"] == "cylinder":
                out.append(CylinderSDF(d["radius"], d["height"], pose))
            else:
                raise ValueError("Invalid primitive type")
        if self.base_type != BaseType.FIXED and out:
            raise ValueError("self_body_collision_primitives are only used for the fixed base robot.")
        return out

    def parse_sphere_specs(self) -> List[SphereAttachmentSpec]:
        """robot_spec.py:345-382."""
        only_self = set(self.conf_dict.get("only_self_collision_links", []))
        specs = []
        for link, spheres in self.conf_dict["collision_spheres"].items():
            arr = np.array(spheres, dtype=np.float64)
            flag = (link in only_self) if self.base_type == BaseType.FIXED else False
            specs.append(SphereAttachmentSpec(link, arr[:, :3].T, arr[:, 3], flag))
        return specs

    def get_self_collision_pairs(self) -> List[List[str]]:
        return [list(p) for p in self.conf_dict.get("self_collision_pairs", [])]

    def create_collision_const(
        self,
        self_collision: bool = True,
        attachments: Sequence[SphereAttachmentSpec] = tuple(),
        use_cache: bool = True,
        reorder_spheres: bool = True,
    ) -> SphereCollisionCst:
        """robot_spec.py:413-492."""
        if not attachments and use_cache and self_collision in self._csts:
            return self._csts[self_collision]
        specs = self.parse_sphere_specs()
        for att in attachments:
            for spec in specs:
                if spec.parent_link_name == att.parent_link_name:
                    spec.relative_positions = np.hstack([spec.relative_positions, att.relative_positions])
                    spec.radii = np.hstack([spec.radii, att.radii])
                    break
            else:
                specs.append(att)
        pairs: List[List[str]] = []
        anchor = None
        if self_collision:
            pairs = self.get_self_collision_pairs()
            prims = self.self_body_collision_primitives()
            if prims:
                anchor = UnionSDF(prims)
        cst = SphereCollisionCst(
            self.get_kin(), self.control_joint_names, self.base_type, specs, [tuple(p) for p in pairs], anchor, reorder_spheres
        )
        if use_cache and not attachments:
            self._csts[self_collision] = cst
        return cst

    def create_attached_box_collision_const(
        self,
        box,
        parent_link_name: str,
        relative_position: np.ndarray,
        n_grid: int = 6,
        reorder_spheres: bool = True,
    ) -> SphereCollisionCst:
        """robot_spec.py:561-595 (deprecated there, kept for drop-in use): a grasped box as an n_grid^3
        lattice of ZERO-radius spheres, keeping the lattice points within 1 cm of the box surface,
        attached to ``parent_link_name`` at ``relative_position`` + (point - box centre); one group,
        no self-collision pairs, no anchor.  ``box`` is a psdf BoxSDF (the reference takes a skrobot
        Box and converts it with primitive_to_plainmp_sdf, utils.py:25-57; anything with
        ``_extents``, ``worldpos()`` and ``worldrot()`` is accepted the same way)."""
        if isinstance(box, BoxSDF):
            extent, center, rot = box.width, box.pose.position, box.pose.rotation
        else:
            extent = np.asarray(box._extents, dtype=np.float64)
            center = np.asarray(box.worldpos(), dtype=np.float64)
            rot = np.asarray(box.worldrot(), dtype=np.float64)
        axes = [np.linspace(-0.5 * e, 0.5 * e, n_grid) for e in extent]
        grid = np.meshgrid(*axes)  # default 'xy' indexing, as the reference calls it
        local = np.stack([g.flatten() for g in grid], axis=1)
        # signed distance of a point inside the box: minus the distance to the nearest face
        # (BoxSDF::evaluate, primitive_sdf.hpp:300-312, for points with every |p_k| <= half_k)
        sd = -np.min(0.5 * extent[None, :] - np.abs(local), axis=1)
        local = local[sd > -1e-2]
        points_from_center = local @ rot.T  # transform_vector(p) - worldpos()
        points_from_link = points_from_center + np.asarray(relative_position, dtype=np.float64).reshape(1, 3)
        spec = SphereAttachmentSpec(parent_link_name, points_from_link.T, np.zeros(len(points_from_link)), False)
        return SphereCollisionCst(self.get_kin(), self.control_joint_names, self.base_type, [spec], [], None, reorder_spheres)


class PandaSpec(RobotSpec):
    conf_name = "panda"
    urdf_name = "panda"


class FetchSpec(RobotSpec):
    conf_name = "fetch"
    urdf_name = "fetch"

    def q_default(self) -> np.ndarray:
        d = self.conf_dict["default_joint_positions"]
        return np.array([d[n] for n in self.control_joint_names])


class PR2SpecBase(RobotSpec):
    """robot_spec.py:691-717."""

    urdf_name = "pr2"

    @property
    def default_joint_positions(self) -> Dict[str, float]:
        return dict(self.conf_dict["default_joint_positions"])

    @property
    def q_default(self) -> np.ndarray:
        d = self.conf_dict["default_joint_positions"]
        return np.array([d[n] for n in self.control_joint_names])

    def reflect_default_joint_positions(self) -> None:
        """example/pr2_dualarm_ik_and_plan.py:27-29: non-controlled joints at the aux.yaml values."""
        self.reflect_joint_positions(self.conf_dict["default_joint_positions"])


class PR2BaseOnlySpec(PR2SpecBase):
    """robot_spec.py:720-732: no controlled joint, PLANAR base — q = (x, y, yaw)."""

    conf_name = "pr2_base_only"

    def __init__(self, conf_file: Optional[Path] = None, urdf_file: Optional[Path] = None):
        super().__init__(BaseType.PLANAR, conf_file, urdf_file)


class PR2DualarmSpec(PR2SpecBase):
    conf_name = "pr2_dualarm"


class PR2RarmSpec(PR2SpecBase):
    conf_name = "pr2_rarm"


class PR2LarmSpec(PR2SpecBase):
    conf_name = "pr2_larm"


class JaxonSpec(RobotSpec):
    """robot_spec.py:841-933: FLOATING-base humanoid, 31 controlled joints + 6 base entries; the foot
    links (on the ground) and optionally the grippers carry no spheres.  The URDF under data/ is a
    jaxon-LIKE skeleton (tools/make_jaxon_skeleton.py), the sphere table is the reference's."""

    conf_name = "jaxon"
    urdf_name = "jaxon"

    def __init__(self, gripper_collision: bool = True, conf_file: Optional[Path] = None, urdf_file: Optional[Path] = None):
        super().__init__(BaseType.FLOATING, conf_file, urdf_file)
        self.gripper_collision = bool(gripper_collision)

    def parse_sphere_specs(self) -> List[SphereAttachmentSpec]:
        ignore = ["RLEG_LINK5", "LLEG_LINK5"]
        if not self.gripper_collision:
            ignore += ["RARM_FINGER0", "RARM_FINGER1", "RARM_LINK7", "LARM_FINGER0", "LARM_FINGER1", "LARM_LINK7"]
        return [s for s in super().parse_sphere_specs() if s.parent_link_name not in ignore]

    @property
    def reset_manip_pose_q(self) -> np.ndarray:
        table = {
            "RLEG": [0.0, 0.0, -0.349066, 0.698132, -0.349066, 0.0],
            "LLEG": [0.0, 0.0, -0.349066, 0.698132, -0.349066, 0.0],
            "CHEST": [0.0, 0.0, 0.0],
            "RARM": [0.0, 0.959931, -0.349066, -0.261799, -1.74533, -0.436332, 0.0, -0.785398],
            "LARM": [0.0, 0.959931, 0.349066, 0.261799, -1.74533, 0.436332, 0.0, -0.785398],
        }
        d = {f"{k}_JOINT{i}": a for k, v in table.items() for i, a in enumerate(v)}
        return np.hstack([[d[j] for j in self.control_joint_names], [0, 0, 1.0, 0, 0, 0]])

    def angle_bounds(self) -> Tuple[np.ndarray, np.ndarray]:
        lb, ub = super().angle_bounds()
        return (np.hstack([lb, [-1.0, -1.0, 0.0, -1.0, -1.0, -1.0]]), np.hstack([ub, [2.0, 1.0, 3.0, 1.0, 1.0, 1.0]]))


class OpenArmV10DualSpec(RobotSpec):
    conf_name = "openarm_v10_dual"
    urdf_name = "openarm_v10"


class OpenArmV10RarmSpec(RobotSpec):
    conf_name = "openarm_v10_rarm"
    urdf_name = "openarm_v10"


class OpenArmV10LarmSpec(RobotSpec):
    conf_name = "openarm_v10_larm"
    urdf_name = "openarm_v10"
