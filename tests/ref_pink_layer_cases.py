"""Seeded cases shared by ``scripts/make_reference_golden.py`` (which runs the REFERENCE's own
Pink-layer Python on them, build container only) and ``tests/test_reference_pink_layer_golden.py``
(which compares the oracle and the kernels with the frozen outputs).  Each case is one of the
scenarios the parity suites already use (tests/helpers.py, tests/extras.py), in product form
(pink_b200 objects) and oracle form (plain records), minus what the reference cannot evaluate
without hpp-fcl / coal: the sphere self-collision barrier."""

import types

from tests import extras, helpers

NAMES = ["ur5_arm", "ur5_unreachable", "draco3_relative", "g1_com_relative", "ur5_limits_barriers_constraint",
         "g1_coupling_floating_base_limit"]


def _from_scenario(sc):
    return types.SimpleNamespace(
        model=sc.model, table=sc.table, B=sc.B, q32=sc.q32, q64=sc.q64, dt=sc.dt, damping=sc.damping,
        tasks=sc.tasks, otasks=sc.oracle_tasks, limits=sc.limits, olimits=sc.oracle_limits,
        barriers=[], obarriers=[], constraints=[], oconstraints=[], collision_model=None, safety_break=False)


def _from_extras(sc):
    keep = [k for k, o in enumerate(sc.obarriers) if o["type"] != "self_collision"]
    return types.SimpleNamespace(
        model=sc.model, table=sc.table, B=sc.B, q32=sc.q32, q64=sc.q64, dt=sc.dt, damping=sc.damping,
        tasks=sc.tasks, otasks=sc.otasks, limits=sc.limits, olimits=sc.olimits,
        barriers=[sc.barriers[k] for k in keep], obarriers=[sc.obarriers[k] for k in keep],
        constraints=sc.constraints, oconstraints=sc.oconstraints, collision_model=None, safety_break=False)


def build(name):
    if name == "ur5_arm":
        return _from_scenario(helpers.ur5_scenario(12, "reachable", seed=501))
    if name == "ur5_unreachable":
        return _from_scenario(helpers.ur5_scenario(12, "unreachable", seed=502))
    if name == "draco3_relative":
        return _from_scenario(helpers.humanoid_scenario("draco3_description", 6, seed=503, with_relative=True))
    if name == "g1_com_relative":
        return _from_scenario(helpers.humanoid_scenario("g1_description", 6, seed=504, with_com=True, with_relative=True))
    if name == "ur5_limits_barriers_constraint":
        return _from_extras(extras.ur5_extras(10, seed=505))
    if name == "g1_coupling_floating_base_limit":
        return _from_extras(extras.g1_extras(6, seed=506))
    raise KeyError(name)
