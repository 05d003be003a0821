"""Barrier base class (``/root/reference/pink/barriers/barrier.py:20-268``).

``compute_barrier`` / ``compute_jacobian`` / ``compute_qp_objective`` /
``compute_qp_inequalities`` are public reference API and are evaluated by the
CUDA library (``pk_constraint_rows_batched`` / ``pk_build_ik_batched``); inside
``solve_ik`` the fused kernel computes the rows on chip.
"""

import abc
from typing import Union

import numpy as np

from .._cabi import PK_GAINFN_IDENTITY


class Barrier(abc.ABC):
    r"""Control barrier function :math:`h(q) \geq 0` with the row
    :math:`-\frac{\partial h}{\partial q} \frac{\Delta q}{dt} \leq \mathrm{gain}\,\alpha(h(q))`
    and the objective term :math:`\frac{r}{2\|J_h\|^2}\|\Delta q\|^2`.

    Attributes:
        dim: Dimension of the barrier.
        gain: linear barrier gain (vector of size ``dim``).
        safe_displacement_gain: gain of the safe backup displacement term.

    The class-K function is one of the two the reference ships: the identity
    (default) and ``h / (1 + |h|)`` (``BodySphericalBarrier``); arbitrary Python
    callables cannot run inside the kernel and are rejected.
    """

    gain_function_id: int = PK_GAINFN_IDENTITY

    def __init__(self, dim: int, gain: Union[float, np.ndarray] = 1.0, gain_function=None,
                 safe_displacement_gain: float = 0.0):
        if gain_function is not None:
            raise NotImplementedError(
                "custom gain functions cannot run inside the CUDA kernel; use the identity "
                "(default) or BodySphericalBarrier's h / (1 + |h|)"
            )
        self.dim = dim
        self.gain = gain if isinstance(gain, np.ndarray) else np.ones(dim) * gain
        self.safe_displacement = np.zeros(self.dim)
        self.safe_displacement_gain = safe_displacement_gain
        self.__q_cache = None

    # -- description for the C-ABI (PkBarrierDesc) --------------------------------
    @abc.abstractmethod
    def _pk_describe(self, configuration_or_model) -> dict:
        """Fields of ``PkBarrierDesc`` plus optional ``pairs`` / ``radii`` arrays."""

    def _raw_rows(self, configuration):
        from ..solve_ik import _barrier_rows

        return _barrier_rows(configuration, self, raw=True)

    def compute_barrier(self, configuration):
        """Value of the barrier function ``h(q)``: ``[dim]`` (``[B, dim]`` batched)."""
        return self._raw_rows(configuration)[1]

    def compute_jacobian(self, configuration):
        """Jacobian ``dh/dq``: ``[dim, nv]`` (``[B, dim, nv]`` batched)."""
        return -self._raw_rows(configuration)[0]

    def compute_safe_displacement(self, configuration):
        """Safe backup displacement: zeros (``barrier.py:131-149``)."""
        return np.zeros(configuration.model.nv)

    def compute_qp_objective(self, configuration):
        r"""``(H, c)`` of the barrier alone (``barrier.py:151-204``):
        ``H = r / |J_h|_F^2 I`` when ``safe_displacement_gain > 1e-6``, ``c = 0``."""
        from ..solve_ik import _barrier_objective

        self.__remember(configuration)
        return _barrier_objective(configuration, self)

    def compute_qp_inequalities(self, configuration, dt: float = 1e-3):
        r"""``(G, h)`` with ``G = -J_h / dt`` and ``h_i = gain_i alpha(h_i(q))``
        (``barrier.py:206-254``)."""
        from ..solve_ik import _barrier_rows

        self.__remember(configuration)
        return _barrier_rows(configuration, self, raw=False, dt=dt)

    def __remember(self, configuration) -> None:
        # the reference caches h(q), J(q) per configuration (``barrier.py:110-129``) and its tests
        # read the cached ``q``; here every evaluation is one kernel launch, only ``q`` is kept
        q = configuration.q
        self.__q_cache = q if hasattr(q, "detach") else np.array(q)  # (no device sync for batched tensors)

    def __repr__(self) -> str:
        return (
            f"{self.__class__.__name__}("
            f"gain={self.gain}, "
            f"safe_displacement={self.safe_displacement}, "
            f"safe_displacement_gain={self.safe_displacement_gain}, "
            f"dim={self.dim})"
        )
