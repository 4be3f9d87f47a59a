#!/usr/bin/env python3
"""Write plainmp_b200/data/urdf/jaxon.urdf: a *jaxon-like* kinematic skeleton (links + joints only).

The reference reads jaxon_jvrc.urdf from the robot_descriptions cache (conf/jaxon.yaml:1), which is
not in the reference tree and cannot be downloaded here.  This skeleton has the joint / link NAMES
conf/jaxon.yaml refers to (31 controlled joints, the sphere-carrying links, the parents of its
custom_links) and the humanoid's joint-axis pattern (legs yaw-roll-pitch-pitch-pitch-roll, chest
roll-pitch-yaw, arms yaw-pitch-roll-yaw-pitch-yaw-roll-pitch, two finger joints per hand) with link
lengths sized from the sphere tables of that file; the numbers are NOT the real robot's.  Oracle,
compiled reference and CUDA path always consume the same URDF text, so parity is unaffected.
"""
from pathlib import Path

OUT = Path(__file__).resolve().parent.parent / "plainmp_b200" / "data" / "urdf" / "jaxon.urdf"

links, joints = ["BODY"], []


def joint(name, parent, child, xyz, axis, lo, hi, kind="revolute"):
    links.append(child)
    joints.append((name, kind, parent, child, xyz, axis, lo, hi))


for side, sy in (("R", -1.0), ("L", 1.0)):
    p = f"{side}LEG"
    joint(f"{p}_JOINT0", "BODY", f"{p}_LINK0", (0.0, 0.1 * sy, 0.0), (0, 0, 1), -0.55, 0.55)
    joint(f"{p}_JOINT1", f"{p}_LINK0", f"{p}_LINK1", (0.0, 0.0, -0.09), (1, 0, 0), -0.6, 0.6)
    joint(f"{p}_JOINT2", f"{p}_LINK1", f"{p}_LINK2", (0.0, 0.0, 0.0), (0, 1, 0), -2.1, 0.75)
    joint(f"{p}_JOINT3", f"{p}_LINK2", f"{p}_LINK3", (0.0, 0.0, -0.38), (0, 1, 0), 0.0, 2.75)
    joint(f"{p}_JOINT4", f"{p}_LINK3", f"{p}_LINK4", (0.0, 0.0, -0.38), (0, 1, 0), -1.4, 1.4)
    joint(f"{p}_JOINT5", f"{p}_LINK4", f"{p}_LINK5", (0.0, 0.0, -0.04), (1, 0, 0), -1.0, 1.0)
joint("CHEST_JOINT0", "BODY", "CHEST_LINK0", (0.0, 0.0, 0.18), (1, 0, 0), -0.15, 0.15)
joint("CHEST_JOINT1", "CHEST_LINK0", "CHEST_LINK1", (0.0, 0.0, 0.0), (0, 1, 0), -0.2, 0.6)
joint("CHEST_JOINT2", "CHEST_LINK1", "CHEST_LINK2", (0.0, 0.0, 0.32), (0, 0, 1), -0.8, 0.8)
joint("HEAD_JOINT0", "CHEST_LINK2", "HEAD_LINK0", (0.0, 0.0, 0.1), (0, 0, 1), -1.0, 1.0)
joint("HEAD_JOINT1", "HEAD_LINK0", "HEAD_LINK1", (0.0, 0.0, 0.05), (0, 1, 0), -0.5, 0.8)
for side, sy in (("R", -1.0), ("L", 1.0)):
    p = f"{side}ARM"
    joint(f"{p}_JOINT0", "CHEST_LINK2", f"{p}_LINK0", (0.0, 0.1 * sy, 0.0), (0, 0, 1), -0.3, 0.8)
    joint(f"{p}_JOINT1", f"{p}_LINK0", f"{p}_LINK1", (0.0, 0.16 * sy, 0.0), (0, 1, 0), -3.1, 1.0)
    joint(f"{p}_JOINT2", f"{p}_LINK1", f"{p}_LINK2", (0.0, 0.0, 0.0), (1, 0, 0), -2.6 if sy < 0 else -0.5, 0.5 if sy < 0 else 2.6)
    joint(f"{p}_JOINT3", f"{p}_LINK2", f"{p}_LINK3", (0.0, 0.0, -0.2), (0, 0, 1), -1.5, 1.5)
    joint(f"{p}_JOINT4", f"{p}_LINK3", f"{p}_LINK4", (0.0, 0.0, -0.08), (0, 1, 0), -2.4, 0.0)
    joint(f"{p}_JOINT5", f"{p}_LINK4", f"{p}_LINK5", (0.0, 0.0, -0.17), (0, 0, 1), -1.6, 1.6)
    joint(f"{p}_JOINT6", f"{p}_LINK5", f"{p}_LINK6", (0.0, 0.0, -0.17), (1, 0, 0), -1.4, 1.4)
    joint(f"{p}_JOINT7", f"{p}_LINK6", f"{p}_LINK7", (0.0, 0.0, 0.0), (0, 1, 0), -1.4, 1.4)
    joint(f"{p}_F_JOINT0", f"{p}_LINK7", f"{p}_FINGER0", (0.0, 0.03, -0.17), (0, 1, 0), -0.8, 0.8)
    joint(f"{p}_F_JOINT1", f"{p}_LINK7", f"{p}_FINGER1", (0.0, -0.03, -0.17), (0, 1, 0), -0.8, 0.8)

out = ['<?xml version="1.0"?>',
       '<!-- jaxon-LIKE kinematic skeleton (links + joints only) written by tools/make_jaxon_skeleton.py; names follow',
       '     conf/jaxon.yaml, dimensions are NOT the real robot\'s; see DESIGN.md "robot tables" -->',
       '<robot name="jaxon_like">']
out += [f'  <link name="{l}" />' for l in links]
for name, kind, parent, child, xyz, axis, lo, hi in joints:
    out += [f'  <joint name="{name}" type="{kind}">',
            f'    <origin xyz="{xyz[0]} {xyz[1]} {xyz[2]}" rpy="0 0 0" />',
            f'    <parent link="{parent}" />', f'    <child link="{child}" />',
            f'    <axis xyz="{axis[0]} {axis[1]} {axis[2]}" />',
            f'    <limit lower="{lo}" upper="{hi}" effort="100" velocity="4" />', '  </joint>']
out.append("</robot>")
OUT.write_text("\n".join(out) + "\n")
print(OUT, len(links), "links", len(joints), "joints")
