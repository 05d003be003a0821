"""``robot_descriptions.loaders.pinocchio.load_robot_description`` for the reference's own test
files (scripts/run_reference_tests.py).  The real package clones URDFs from the network; here a
name maps to a STAND-IN robot of the same class built from the offline descriptions of
``pink_b200.robots`` (same joint / frame names where the reference's tests use them, different
geometry), or the test is skipped.  Known-answer tests that hold numbers of the real robot
(JVRC-1 Jacobian, Stretch relative Jacobian) cannot pass on a stand-in and are expected to fail;
the model-independent ones (finite differences, identities, shapes, convergence) are the point."""
import unittest

import numpy as np

from pink_b200.model import SE3, model_from_urdf_string, RobotWrapper
from pink_b200.robots import load_robot_description as _offline

_UPKIE = """
<robot name="upkie_standin">
  <link name="base"><inertial><origin xyz="0 0 0.05"/><mass value="4.0"/></inertial></link>
  <link name="imu"/>
  <joint name="imu_joint" type="fixed"><parent link="base"/><child link="imu"/><origin xyz="0 0 0.1"/></joint>
  %s
</robot>
"""
_LEG = """
  <link name="{s}_hip_link"><inertial><origin xyz="0 0 -0.08"/><mass value="0.6"/></inertial></link>
  <joint name="{s}_hip" type="revolute"><parent link="base"/><child link="{s}_hip_link"/><origin xyz="0 {y} 0"/>
    <axis xyz="0 1 0"/><limit lower="-1.26" upper="1.26" velocity="28.8" effort="16"/></joint>
  <link name="{s}_knee_link"><inertial><origin xyz="0 0 -0.08"/><mass value="0.5"/></inertial></link>
  <joint name="{s}_knee" type="revolute"><parent link="{s}_hip_link"/><child link="{s}_knee_link"/><origin xyz="0 0 -0.17"/>
    <axis xyz="0 1 0"/><limit lower="-2.51" upper="2.51" velocity="28.8" effort="16"/></joint>
  <link name="{s}_wheel_link"><inertial><origin xyz="0 0 0"/><mass value="0.3"/></inertial></link>
  <joint name="{s}_wheel" type="continuous"><parent link="{s}_knee_link"/><child link="{s}_wheel_link"/><origin xyz="0 {w} -0.17"/>
    <axis xyz="0 1 0"/><limit velocity="111.0" effort="1.7"/></joint>
  <link name="{s}_contact"/>
  <joint name="{s}_contact_joint" type="fixed"><parent link="{s}_wheel_link"/><child link="{s}_contact"/><origin xyz="0 0 -0.06"/></joint>
"""

_JVRC_ALIASES = {"l_ankle": "left_ankle_roll_link", "r_ankle": "right_ankle_roll_link", "PELVIS_S": "pelvis",
                 "l_wrist": "left_wrist_yaw_link", "r_wrist": "right_wrist_yaw_link"}


def _alias_frames(model, aliases):
    for new, old in aliases.items():
        f = model.frames[model.getFrameId(old)]
        model.add_frame(new, f.parentJoint, SE3(f.placement.rotation, f.placement.translation), "BODY")


def load_robot_description(name, root_joint=None, commit=None):
    import pinocchio as pin  # the stand-in next to this package

    if isinstance(root_joint, pin._Unsupported):
        raise unittest.SkipTest(f"{type(root_joint).__name__} root joints are outside the scope of this repo")
    robot = _load(name, root_joint)
    pin.pinocchio_like_limits(robot.model)
    return pin.RobotWrapper(robot.model)


def _load(name, root_joint):
    if name in ("ur3_official_description", "ur5_description", "ur5_official_description"):
        return _offline("ur5_description", root_joint=root_joint)
    if name == "draco3_description":
        return _offline("draco3_description", root_joint=root_joint)
    if name == "jvrc_description":
        robot = _offline("g1_description", root_joint=root_joint)
        _alias_frames(robot.model, _JVRC_ALIASES)
        return RobotWrapper(robot.model)
    if name == "upkie_description":
        legs = _LEG.format(s="left", y=0.1, w=0.05) + _LEG.format(s="right", y=-0.1, w=-0.05)
        return RobotWrapper(model_from_urdf_string(_UPKIE % legs, root_joint=root_joint))
    if name == "yumi_description":
        return RobotWrapper(model_from_urdf_string(_yumi(), root_joint=root_joint))
    if name == "sigmaban_description":  # a humanoid whose URDF carries velocity limits but no position limits
        robot = _offline("g1_description", root_joint=root_joint)
        m = robot.model
        m.lowerPositionLimit = np.full(m.nq, -np.inf)
        m.upperPositionLimit = np.full(m.nq, np.inf)
        return RobotWrapper(m)
    raise unittest.SkipTest(f"no offline stand-in for {name!r}")


def _yumi():
    """Fixed-base dual arm: 2 x (7 revolute + 2 prismatic fingers) = 18 joints, frames yumi_link_7_{l,r}."""
    out = ['<robot name="yumi_standin"><link name="yumi_body"/>']
    axes = ["0 0 1", "0 1 0", "0 0 1", "0 1 0", "0 0 1", "0 1 0", "0 0 1"]
    for side, y in (("l", 0.1), ("r", -0.1)):
        parent = "yumi_body"
        for k in range(1, 8):
            link = f"yumi_link_{k}_{side}"
            out.append(f'<link name="{link}"><inertial><origin xyz="0 0 0.05"/><mass value="1.0"/></inertial></link>')
            origin = f"0.05 {y} 0.3" if k == 1 else f"{0.03 * (k % 2)} 0 {0.1 + 0.02 * k}"
            out.append(f'<joint name="yumi_joint_{k}_{side}" type="revolute"><parent link="{parent}"/><child link="{link}"/>'
                       f'<origin xyz="{origin}" rpy="{0.3 if k == 1 else 0} {0.5 if k == 2 else 0} 0"/><axis xyz="{axes[k - 1]}"/>'
                       f'<limit lower="-2.9" upper="2.9" velocity="3.14" effort="50"/></joint>')
            parent = link
        for f in ("", "_m"):
            link = f"gripper_{side}_finger{f}"
            out.append(f'<link name="{link}"/>')
            out.append(f'<joint name="gripper_{side}_joint{f}" type="prismatic"><parent link="yumi_link_7_{side}"/><child link="{link}"/>'
                       f'<origin xyz="0 {0.01 if f else -0.01} 0.08"/><axis xyz="1 0 0"/>'
                       f'<limit lower="0" upper="0.025" velocity="2" effort="20"/></joint>')
    out.append("</robot>")
    return "\n".join(out)
