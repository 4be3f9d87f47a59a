#!/usr/bin/env python3
"""Import the reference's robot *data tables* (not code) into plainmp_b200/data/.

Run in the authoring container only (it reads /root/reference, which does not exist on
the GPU box).  Outputs are small, committed fixtures:

  plainmp_b200/data/conf/<robot>.json   sphere tables, self-collision pairs, self-only links,
                                        anchor primitives, default joint positions
                                        <- /root/reference/src/plainmp/conf/*.yaml
                                        (schema: src/plainmp/robot_spec.py:93-114,263-299,345-411)
  plainmp_b200/data/urdf/openarm_v10.urdf   kinematic skeleton only (links + joints; visuals,
                                        collisions, inertials and meshes dropped)
                                        <- /root/reference/src/plainmp/assets/openarm_v10.urdf

The panda / fetch / pr2 URDFs are NOT in the reference tree (they are downloaded by
scikit-robot at run time, conf/panda.yaml:1), nor is jaxon_jvrc.urdf (robot_descriptions cache,
conf/jaxon.yaml:1); plainmp_b200/data/urdf/{panda,fetch,pr2,jaxon}.urdf are authored by this repo
("robot tables"; jaxon: tools/make_jaxon_skeleton.py).
"""
import json
import sys
import xml.etree.ElementTree as ET
from pathlib import Path

import yaml

REF = Path("/root/reference/src/plainmp")
OUT = Path(__file__).resolve().parent.parent / "plainmp_b200" / "data"

CONFS = {
    "panda": "conf/panda.yaml",
    "fetch": "conf/fetch.yaml",
    "pr2_dualarm": "conf/pr2_dualarm.yaml",
    "pr2_rarm": "conf/pr2_rarm.yaml",
    "pr2_larm": "conf/pr2_larm.yaml",
    "pr2_base_only": "conf/pr2_base_only.yaml",
    "jaxon": "conf/jaxon.yaml",
    "openarm_v10_dual": "conf/openarm_v10/openarm_v10_dual.yaml",
    "openarm_v10_rarm": "conf/openarm_v10/openarm_v10_rarm.yaml",
    "openarm_v10_larm": "conf/openarm_v10/openarm_v10_larm.yaml",
}


def load_conf(path: Path) -> dict:
    # one-level include, robot_spec.py:100-114
    with open(path) as f:
        d = yaml.safe_load(f)
    for rel in d.pop("include", []):
        with open(path.parent / rel) as f:
            dep = yaml.safe_load(f)
        assert not (set(d) & set(dep)), "conflict"
        assert "include" not in dep
        d.update(dep)
    return d


def normalise(d: dict) -> dict:
    out = {"control_joint_names": list(d["control_joint_names"])}
    # robot_spec.py:387-411 get_self_collision_pairs
    pairs = [list(p) for p in d.get("self_collision_pairs", [])]
    for gp in d.get("self_collision_group_pairs", []):
        for l1 in gp["group1"]:
            for l2 in gp["group2"]:
                pairs.append([l1, l2])
    seen = set()
    for p in pairs:
        key = tuple(sorted(p))
        assert key not in seen, f"duplicated pair {key}"
        seen.add(key)
    out["self_collision_pairs"] = pairs
    out["only_self_collision_links"] = list(d.get("only_self_collision_links", []))
    out["collision_spheres"] = {
        link: [[float(x) for x in s] for s in v["spheres"]] for link, v in d["collision_spheres"].items()
    }
    out["self_body_collision_primitives"] = d.get("self_body_collision_primitives", [])
    djp = d.get("default_joint_positions", {})
    if isinstance(djp, list):  # fetch.yaml uses a list of one-key dicts
        merged = {}
        for e in djp:
            merged.update(e)
        djp = merged
    out["default_joint_positions"] = {k: float(v) for k, v in djp.items()}
    out["custom_links"] = d.get("custom_links", {})
    return out


def strip_urdf(src: Path, dst: Path) -> None:
    root = ET.parse(src).getroot()
    new = ET.Element("robot", {"name": root.get("name", "robot")})
    for el in root:
        if el.tag == "link":
            ET.SubElement(new, "link", {"name": el.get("name")})
        elif el.tag == "joint":
            j = ET.SubElement(new, "joint", {"name": el.get("name"), "type": el.get("type")})
            for tag in ("origin", "parent", "child", "axis", "limit"):
                c = el.find(tag)
                if c is not None:
                    ET.SubElement(j, tag, dict(c.attrib))
    ET.indent(new)
    dst.write_text(ET.tostring(new, encoding="unicode") + "\n")


def main() -> int:
    (OUT / "conf").mkdir(parents=True, exist_ok=True)
    (OUT / "urdf").mkdir(parents=True, exist_ok=True)
    for name, rel in CONFS.items():
        d = normalise(load_conf(REF / rel))
        (OUT / "conf" / f"{name}.json").write_text(json.dumps(d, indent=1) + "\n")
        n_s = sum(len(v) for v in d["collision_spheres"].values())
        print(f"{name}: {len(d['collision_spheres'])} groups, {n_s} spheres, {len(d['self_collision_pairs'])} pairs")
    strip_urdf(REF / "assets" / "openarm_v10.urdf", OUT / "urdf" / "openarm_v10.urdf")
    return 0


if __name__ == "__main__":
    sys.exit(main())
