"""Batched robot configuration.

Mirrors ``pink.Configuration`` (``/root/reference/pink/configuration.py:26-293``):
same constructor arguments, attributes (``model``, ``data``, ``q``,
``tangent``) and methods.  Extension: ``q`` may carry a leading batch
dimension ``[B, nq]`` (numpy or torch); frame transforms and Jacobians are then
returned as device tensors with that leading dimension.  A 1-D ``q`` behaves
like the reference (numpy / :class:`SE3` values in, numpy out).

Forward kinematics is evaluated by the CUDA library on demand (the fused
``solve_ik`` kernel recomputes it from ``q`` and never needs these caches).
"""

from __future__ import annotations

import logging
from typing import Optional

import numpy as np
import torch

from .engine import get_engine
from .exceptions import FrameNotFound, NotWithinConfigurationLimits
from .limits import ConfigurationLimit, VelocityLimit
from .spatial import SE3
from .utils import VectorSpace, get_root_joint_dim


class Configuration:
    """Kinematic state of ``B`` independent robot instances of one model."""

    def __init__(
        self,
        model,
        data=None,
        q=None,
        copy_data: bool = True,
        forward_kinematics: bool = True,
        collision_model=None,
        collision_data=None,
        device=None,
    ):
        if collision_model is not None and collision_model.model is not model:
            raise ValueError("the collision model was built for another robot model")
        # per-model defaults cached on the model object (configuration.py:101-108)
        if not hasattr(model, "tangent"):
            model.tangent = VectorSpace(model.nv)
        if not hasattr(model, "configuration_limit"):
            model.configuration_limit = ConfigurationLimit(model)
        if not hasattr(model, "velocity_limit"):
            model.velocity_limit = VelocityLimit(model)
        if not hasattr(model, "floating_base_velocity_limit"):
            model.floating_base_velocity_limit = None
        self.model = model
        if data is None:
            data = model.createData()
        self.data = data.copy() if copy_data else data
        self.tangent = model.tangent
        # sphere collision model (pink_b200.collision.SphereCollisionModel); the kernels
        # evaluate the pair distances themselves, collision_data only serves user code that
        # reads distanceResults (pink_b200.collision.SphereCollisionData)
        self.collision_model = collision_model
        self.collision_data = collision_data
        if collision_data is not None and hasattr(collision_data, "bind"):
            collision_data.bind(self)
        self._device = device
        self._set_q(q)
        # forward kinematics is lazy: nothing to do here

    # -- state ---------------------------------------------------------------
    def _set_q(self, q) -> None:
        if q is None:
            raise ValueError("a configuration vector q is required")
        self._fk = None
        self._com = None
        if isinstance(q, torch.Tensor):
            self.batched = q.dim() == 2
            if q.dim() not in (1, 2) or q.shape[-1] != self.model.nq:
                raise ValueError(f"q has shape {tuple(q.shape)}, expected [..., {self.model.nq}]")
            if q.is_cuda:
                self._device = q.device
                self._q_dev = q.detach().to(torch.float32).reshape(-1, self.model.nq).clone()
                self._q_host = None
                self.q = self._q_dev if self.batched else self._q_dev[0]
            else:
                self._q_dev = None
                self._q_host = q.detach().to(torch.float64).reshape(-1, self.model.nq).numpy().copy()
                self.q = self._q_host if self.batched else self._q_host[0]
                if isinstance(self.q, np.ndarray):
                    self.q.setflags(write=False)
        else:
            arr = np.array(q, dtype=np.float64)  # copy (configuration.py:109)
            if arr.ndim not in (1, 2) or arr.shape[-1] != self.model.nq:
                raise ValueError(f"q has shape {arr.shape}, expected [..., {self.model.nq}]")
            self.batched = arr.ndim == 2
            arr.setflags(write=False)
            # (a model without joints, nq = 0, is one instance with an empty vector)
            self._q_host = arr.reshape(-1, self.model.nq) if self.model.nq else arr.reshape(1, 0)
            self._q_dev = None
            self.q = arr

    @property
    def batch_size(self) -> int:
        return int(self._q_dev.shape[0] if self._q_dev is not None else self._q_host.shape[0])

    @property
    def engine(self):
        return get_engine(self.model, self._device)

    @property
    def q_device(self) -> torch.Tensor:
        """``[B, nq]`` fp32 tensor on the compute device."""
        if self._q_dev is None:
            eng = self.engine
            self._q_dev = torch.tensor(self._q_host, dtype=torch.float32).to(eng.device)
            self._device = eng.device
        return self._q_dev

    def update(self, q=None) -> None:
        """Set a new configuration (``configuration.py:131-164``)."""
        if q is not None:
            self._set_q(q)
        else:
            self._fk = None
            self._com = None

    # -- limits --------------------------------------------------------------
    def check_limits(self, tol: float = 1e-6, safety_break: bool = True) -> None:
        """``configuration.py:166-201``; batched: the first offending instance
        is reported."""
        q_max = self.model.upperPositionLimit
        q_min = self.model.lowerPositionLimit
        root_nq, _ = get_root_joint_dim(self.model)
        q = self._q_host if self._q_host is not None else self._q_dev.detach().cpu().numpy()
        for i in range(root_nq, self.model.nq):
            if q_max[i] <= q_min[i] + tol:  # no limit
                continue
            bad = np.nonzero((q[:, i] < q_min[i] - tol) | (q[:, i] > q_max[i] + tol))[0]
            if bad.size:
                b = int(bad[0])
                if safety_break:
                    raise NotWithinConfigurationLimits(
                        i, float(q[b, i]), q_min[i], q_max[i], instance=b if self.batched else None
                    )
                logging.warning(
                    "Value %f at index %d is out of limits: [%f, %f]",
                    float(q[b, i]), i, q_min[i], q_max[i],
                )

    # -- kinematics ----------------------------------------------------------
    def _ensure_fk(self) -> torch.Tensor:
        if self._fk is None:
            self._fk, _ = self.engine.forward_kinematics(self.q_device)
            if not self.batched:
                host = self._fk[0].cpu().numpy().astype(np.float64)
                self.data.oMf = [SE3(T[:, :3], T[:, 3]) for T in host]
        return self._fk

    def _frame_id(self, frame: str) -> int:
        if not self.model.existFrame(frame):
            raise FrameNotFound(frame, self.model.frames)
        return self.model.getFrameId(frame)

    def get_frame_jacobian(self, frame: str):
        """LOCAL frame Jacobian (``configuration.py:203-236``): ``[6, nv]``
        numpy, or ``[B, 6, nv]`` tensor when batched."""
        fid = self._frame_id(frame)
        J = self.engine.frame_jacobian(fid, self.q_device)
        return J if self.batched else J[0].cpu().numpy().astype(np.float64)

    def get_transform_frame_to_world(self, frame: str):
        """Frame pose (``configuration.py:238-254``): :class:`SE3`, or a
        ``[B, 3, 4]`` tensor ``[R | p]`` when batched."""
        fid = self._frame_id(frame)
        fk = self._ensure_fk()
        if self.batched:
            return fk[:, fid].clone()
        return self.data.oMf[fid].copy()

    def get_transform(self, source: str, dest: str):
        """Pose of ``source`` in ``dest`` (``configuration.py:256-271``)."""
        a = self.get_transform_frame_to_world(source)
        b = self.get_transform_frame_to_world(dest)
        if not self.batched:
            return b.actInv(a)
        Rb, pb = b[:, :, :3], b[:, :, 3]
        Ra, pa = a[:, :, :3], a[:, :, 3]
        R = Rb.transpose(1, 2) @ Ra
        p = (Rb.transpose(1, 2) @ (pa - pb).unsqueeze(-1)).squeeze(-1)
        return torch.cat([R, p.unsqueeze(-1)], dim=-1)

    def get_center_of_mass(self):
        """``pin.centerOfMass`` (used by ``ComTask.set_target_from_configuration``,
        ``pink/tasks/com_task.py:94-106``)."""
        if self._com is None:
            _, self._com = self.engine.forward_kinematics(self.q_device, want_com=True)
        return self._com if self.batched else self._com[0].cpu().numpy().astype(np.float64)

    # -- integration ---------------------------------------------------------
    def integrate(self, velocity, dt):
        """``q (+) velocity * dt`` (``configuration.py:273-283``)."""
        eng = self.engine
        v = eng._f32(velocity, self.model.nv)
        if v.shape[0] == 1 and self.batch_size > 1:
            v = v.expand(self.batch_size, -1).contiguous()
        out = eng.integrate(self.q_device, v, float(dt))
        return out if self.batched else out[0].cpu().numpy().astype(np.float64)

    def integrate_inplace(self, velocity, dt) -> None:
        """``configuration.py:285-293``."""
        self.update(self.integrate(velocity, dt))
