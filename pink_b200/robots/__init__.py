"""Robot descriptions available offline.

The reference's examples fetch URDFs from the network through
``robot_descriptions.loaders.pinocchio.load_robot_description``
(``/root/reference/examples/arm_ur5.py:26``,
``examples/humanoid_draco3.py:55-57``, ``examples/humanoid_g1_com.py:28-30``).
None of these files exists offline, so :func:`load_robot_description` serves

* ``ur5_description`` / ``ur5_official_description``: the hand-authored
  ``ur5.urdf`` next to this file (provenance in its header);
* ``draco3_description``: a SYNTHETIC Draco3-class humanoid - same joint
  count (27 + free-flyer), joint names, topology (7-dof legs with a
  proximal/distal knee pair, 6-dof arms, neck) and task frame names as the
  example uses, with made-up link lengths, limits and masses;
* ``g1_description``: a SYNTHETIC G1-class humanoid - 29 joints (6-dof legs,
  3-dof waist, 7-dof arms) + free-flyer, frame names of the example, made-up
  geometry.

Oracle and GPU read the same generated tables, so parity tests are unaffected
by the geometry being synthetic; benchmark lines name the model explicitly.
"""

from __future__ import annotations

import os
from typing import List

from ..model import RobotWrapper, model_from_urdf_string

_HERE = os.path.dirname(os.path.abspath(__file__))


class _Urdf:
    """Tiny URDF text builder for serial limbs."""

    def __init__(self, name: str, root_link: str, root_mass: float, root_com=(0, 0, 0)):
        self.lines: List[str] = [f'<robot name="{name}">']
        self.link(root_link, root_mass, root_com)

    def link(self, name, mass=0.0, com=(0, 0, 0)):
        if mass > 0:
            self.lines.append(
                f'<link name="{name}"><inertial><mass value="{mass}"/>'
                f'<origin xyz="{com[0]} {com[1]} {com[2]}" rpy="0 0 0"/>'
                '<inertia ixx="0.01" ixy="0" ixz="0" iyy="0.01" iyz="0" izz="0.01"/></inertial></link>'
            )
        else:
            self.lines.append(f'<link name="{name}"/>')

    def joint(self, name, parent, child, xyz, axis, lower, upper, vel, rpy=(0, 0, 0), mass=1.0, com=(0, 0, 0)):
        self.link(child, mass, com)
        self.lines.append(
            f'<joint name="{name}" type="revolute"><parent link="{parent}"/><child link="{child}"/>'
            f'<origin xyz="{xyz[0]} {xyz[1]} {xyz[2]}" rpy="{rpy[0]} {rpy[1]} {rpy[2]}"/>'
            f'<axis xyz="{axis[0]} {axis[1]} {axis[2]}"/>'
            f'<limit lower="{lower}" upper="{upper}" effort="100" velocity="{vel}"/></joint>'
        )

    def fixed(self, name, parent, child, xyz, rpy=(0, 0, 0)):
        self.link(child)
        self.lines.append(
            f'<joint name="{name}" type="fixed"><parent link="{parent}"/><child link="{child}"/>'
            f'<origin xyz="{xyz[0]} {xyz[1]} {xyz[2]}" rpy="{rpy[0]} {rpy[1]} {rpy[2]}"/></joint>'
        )

    def text(self) -> str:
        return "\n".join(self.lines + ["</robot>"])


X, Y, Z = (1, 0, 0), (0, 1, 0), (0, 0, 1)


def draco3_class_urdf() -> str:
    """SYNTHETIC 27-joint humanoid with Draco3's joint and frame names."""
    u = _Urdf("draco3_class_synthetic", "torso_link", 12.0, (0.0, 0.0, 0.18))
    u.fixed("torso_com_joint", "torso_link", "torso_com_link", (0.0, 0.0, 0.18))
    u.joint("neck_pitch", "torso_link", "neck_pitch_link", (0.0, 0.0, 0.46), Y, -0.52, 1.05, 3.0, mass=1.5, com=(0, 0, 0.08))
    for s, sy in (("l", 1.0), ("r", -1.0)):
        # 7-dof leg with proximal / distal knee joints
        leg = [
            ("hip_ie", (0.0, sy * 0.10, -0.04), Z, -0.87, 0.87, 6.0, 1.2, (0, 0, -0.02)),
            ("hip_aa", (0.0, 0.0, -0.06), X, -0.52, 0.52, 6.0, 1.0, (0, 0, -0.02)),
            ("hip_fe", (0.0, 0.0, -0.04), Y, -1.57, 0.52, 8.0, 3.5, (0, 0, -0.10)),
            ("knee_fe_jp", (0.0, 0.0, -0.21), Y, -0.09, 1.57, 8.0, 0.6, (0, 0, -0.02)),
            ("knee_fe_jd", (0.02, 0.0, -0.05), Y, -0.09, 1.57, 8.0, 2.4, (0, 0, -0.12)),
            ("ankle_fe", (-0.02, 0.0, -0.25), Y, -1.57, 1.05, 8.0, 0.3, (0, 0, -0.01)),
            ("ankle_ie", (0.0, 0.0, -0.02), X, -0.52, 0.52, 8.0, 0.8, (0.02, 0, -0.04)),
        ]
        parent = "torso_link"
        for jn, xyz, ax, lo, hi, vel, mass, com in leg:
            child = f"{s}_{jn}_link"
            u.joint(f"{s}_{jn}", parent, child, xyz, ax, lo, hi, vel, mass=mass, com=com)
            parent = child
        u.fixed(f"{s}_foot_contact_frame", parent, f"{s}_foot_contact", (0.0, 0.0, -0.076))
        # 6-dof arm
        arm = [
            ("shoulder_fe", (0.0, sy * 0.16, 0.36), Y, -2.27, 1.05, 8.0, 1.0, (0, sy * 0.02, 0)),
            ("shoulder_aa", (0.0, sy * 0.05, 0.0), X, -1.57 if sy < 0 else 0.0, 0.0 if sy < 0 else 1.57, 8.0, 0.9, (0, 0, -0.03)),
            ("shoulder_ie", (0.0, 0.0, -0.06), Z, -1.5, 1.5, 8.0, 1.1, (0, 0, -0.10)),
            ("elbow_fe", (0.0, 0.0, -0.19), Y, -2.09, 0.09, 8.0, 0.8, (0, 0, -0.06)),
            ("wrist_ps", (0.0, 0.0, -0.12), Z, -1.57, 1.57, 10.0, 0.5, (0, 0, -0.05)),
            ("wrist_pitch", (0.0, 0.0, -0.10), Y, -1.57, 1.57, 10.0, 0.4, (0, 0, -0.03)),
        ]
        parent = "torso_link"
        for jn, xyz, ax, lo, hi, vel, mass, com in arm:
            child = f"{s}_{jn}_link"
            u.joint(f"{s}_{jn}", parent, child, xyz, ax, lo, hi, vel, mass=mass, com=com)
            parent = child
        u.fixed(f"{s}_hand_contact_frame", parent, f"{s}_hand_contact", (0.0, 0.0, -0.09))
    return u.text()


def g1_class_urdf() -> str:
    """SYNTHETIC 29-joint humanoid with Unitree G1's joint and link names."""
    u = _Urdf("g1_class_synthetic", "pelvis", 6.0, (0.0, 0.0, -0.02))
    for side, sy in (("left", 1.0), ("right", -1.0)):
        leg = [
            ("hip_pitch", (0.0, sy * 0.064, -0.10), Y, -2.53, 2.88, 32.0, 1.35, (0, sy * 0.03, -0.02)),
            ("hip_roll", (0.0, sy * 0.052, -0.03), X, -0.52 if sy > 0 else -2.97, 2.97 if sy > 0 else 0.52, 20.0, 1.52, (0.02, 0, -0.06)),
            ("hip_yaw", (0.025, 0.0, -0.12), Z, -2.76, 2.76, 32.0, 1.70, (0, 0, -0.08)),
            ("knee", (0.078, 0.0, -0.18), Y, -0.087, 2.88, 20.0, 1.93, (0.0, 0, -0.12)),
            ("ankle_pitch", (0.0, 0.0, -0.30), Y, -0.87, 0.52, 37.0, 0.07, (0, 0, 0)),
            ("ankle_roll", (0.0, 0.0, -0.017), X, -0.26, 0.26, 37.0, 0.61, (0.03, 0, -0.02)),
        ]
        parent = "pelvis"
        for jn, xyz, ax, lo, hi, vel, mass, com in leg:
            child = f"{side}_{jn}_link"
            u.joint(f"{side}_{jn}_joint", parent, child, xyz, ax, lo, hi, vel, mass=mass, com=com)
            parent = child
    u.joint("waist_yaw_joint", "pelvis", "waist_yaw_link", (0.0, 0.0, 0.0), Z, -2.62, 2.62, 32.0, mass=0.21, com=(0, 0, 0.02))
    u.joint("waist_roll_joint", "waist_yaw_link", "waist_roll_link", (0.0, 0.0, 0.035), X, -0.52, 0.52, 37.0, mass=0.09, com=(0, 0, 0.01))
    u.joint("waist_pitch_joint", "waist_roll_link", "torso_link", (0.0, 0.0, 0.019), Y, -0.52, 0.52, 37.0, mass=7.8, com=(0.0, 0, 0.15))
    for side, sy in (("left", 1.0), ("right", -1.0)):
        arm = [
            ("shoulder_pitch", (0.004, sy * 0.10, 0.24), Y, -3.09, 2.67, 37.0, 0.72, (0, sy * 0.03, 0)),
            ("shoulder_roll", (0.0, sy * 0.038, -0.014), X, -1.59 if sy > 0 else -2.25, 2.25 if sy > 0 else 1.59, 37.0, 0.64, (0, 0, -0.03)),
            ("shoulder_yaw", (0.0, sy * 0.006, -0.10), Z, -2.62, 2.62, 37.0, 0.73, (0, 0, -0.05)),
            ("elbow", (0.016, 0.0, -0.08), Y, -1.05, 2.09, 37.0, 0.60, (0.05, 0, -0.01)),
            ("wrist_roll", (0.10, 0.0, -0.01), X, -1.97, 1.97, 37.0, 0.09, (0.02, 0, 0)),
            ("wrist_pitch", (0.038, 0.0, 0.0), Y, -1.61, 1.61, 22.0, 0.48, (0.02, 0, 0)),
            ("wrist_yaw", (0.046, 0.0, 0.0), Z, -1.61, 1.61, 22.0, 0.25, (0.03, 0, 0)),
        ]
        parent = "torso_link"
        for jn, xyz, ax, lo, hi, vel, mass, com in arm:
            child = f"{side}_{jn}_link"
            u.joint(f"{side}_{jn}_joint", parent, child, xyz, ax, lo, hi, vel, mass=mass, com=com)
            parent = child
    return u.text()


def ur5_urdf() -> str:
    with open(os.path.join(_HERE, "ur5.urdf"), "r", encoding="utf-8") as fh:
        return fh.read()


_DESCRIPTIONS = {
    "ur5_description": ur5_urdf,
    "ur5_official_description": ur5_urdf,
    "draco3_description": draco3_class_urdf,
    "g1_description": g1_class_urdf,
}


def load_robot_description(name: str, root_joint=None) -> RobotWrapper:
    """Offline stand-in for ``robot_descriptions.loaders.pinocchio.load_robot_description``."""
    if name not in _DESCRIPTIONS:
        raise KeyError(f"no offline description named {name!r}; available: {sorted(_DESCRIPTIONS)}")
    return RobotWrapper(model_from_urdf_string(_DESCRIPTIONS[name](), root_joint=root_joint))
