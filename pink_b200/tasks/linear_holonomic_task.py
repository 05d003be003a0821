"""Linear holonomic task
(``/root/reference/pink/tasks/linear_holonomic_task.py``)."""

from typing import Optional, Sequence, Union

import numpy as np

from .._cabi import PK_TASK_LINEAR
from ..exceptions import TaskDefinitionError, TaskJacobianNotSet
from ..model import neutral
from ..utils import get_root_joint_dim
from .task import Task


class LinearHolonomicTask(Task):
    r"""Linear constraint on the configuration,
    :math:`e(q) = A (q \ominus q_0) - b`, :math:`J = A\,\partial(q \ominus q_0)/\partial q`
    (``linear_holonomic_task.py:148-192``).

    The CUDA engine evaluates it for ``A`` acting on joint coordinates only
    (zero columns on the floating base, as built by :class:`JointCouplingTask`):
    there ``q (-) q_0`` is a subtraction and ``J = A``.
    """

    def __init__(self, A: np.ndarray, b: np.ndarray, q_0: Optional[np.ndarray],
                 cost: Optional[Union[float, Sequence[float], np.ndarray]] = None, lm_damping: float = 0.0,
                 gain: float = 1.0) -> None:
        super().__init__(cost=cost, gain=gain, lm_damping=lm_damping)
        if b.shape[0] != A.shape[0]:
            raise TaskDefinitionError(f"Shape mismatch between {A.shape=} and {b.shape=}")
        self.A = A
        self.b = b
        self.q_0 = q_0

    def _pk_describe(self, model) -> dict:
        A = np.asarray(self.A, dtype=np.float64)
        if A.ndim != 2 or A.shape[1] != model.nv:
            raise TaskJacobianNotSet(f"A has shape {A.shape} but the model has nv={model.nv}")
        p = A.shape[0]
        if p > 6:
            raise NotImplementedError("LinearHolonomicTask supports up to 6 rows on the CUDA engine")
        _, root_nv = get_root_joint_dim(model)
        if root_nv and np.any(A[:, :root_nv] != 0.0):
            raise NotImplementedError(
                "LinearHolonomicTask on floating-base coordinates is not supported by the CUDA engine"
            )
        q_ref = neutral(model) if self.q_0 is None else np.asarray(self.q_0, dtype=np.float64)
        cost = self.cost
        cost6 = np.zeros(6)
        if cost is None:
            cost6[:p] = 1.0
        elif isinstance(cost, (float, int)):
            cost6[:p] = float(cost)
        else:
            c = np.asarray(cost, dtype=np.float64).reshape(-1)
            if c.shape[0] != p:
                raise TaskDefinitionError(f"cost has {c.shape[0]} entries but the task has {p} rows")
            cost6[:p] = c
        data = np.concatenate([A.reshape(-1), np.asarray(self.b, dtype=np.float64).reshape(-1), q_ref.reshape(-1)])
        return {"type": PK_TASK_LINEAR, "frame": 0, "root": 0, "cost6": cost6, "k": p, "target": np.zeros(0),
                "rows": p, "data": data}

    def __repr__(self):
        return f"LinearHolonomicTask(cost={self.cost}, gain={self.gain}, lm_damping={self.lm_damping})"


class JointCouplingTask(LinearHolonomicTask):
    r"""Coupling :math:`\sum_i r_i q_i = 0` between joints
    (``/root/reference/pink/tasks/joint_coupling_task.py:53-102``)."""

    def __init__(self, joint_names: Sequence[str], ratios: Sequence[float], cost: float, configuration,
                 lm_damping: float = 0.0, gain: float = 1.0) -> None:
        assert len(joint_names) == len(ratios)
        model = configuration.model
        A = np.zeros((1, model.nv))
        for joint, ratio in zip(joint_names, ratios):
            joint_obj = model.joints[model.getJointId(joint)]
            A[:, joint_obj.idx_v:joint_obj.idx_v + joint_obj.nv] = ratio
        super().__init__(A, np.zeros(1), neutral(model), cost=cost, gain=gain, lm_damping=lm_damping)
        self.joint_names = joint_names
        self.ratios = ratios

    def __repr__(self):
        return f"JointCouplingTask(cost={self.cost}, gain={self.gain}, lm_damping={self.lm_damping})"
