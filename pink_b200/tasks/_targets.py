"""Conversion of user-supplied targets to the C-ABI layouts."""

import numpy as np
import torch

from ..spatial import SE3


def as_se3_target(T):
    """SE3-like input -> :class:`SE3` (shared) or ``[B, 12]`` tensor (per instance)."""
    if isinstance(T, SE3):
        return T.copy()
    if hasattr(T, "rotation") and hasattr(T, "translation"):  # pin.SE3 duck type
        return SE3(np.array(T.rotation), np.array(T.translation))
    if isinstance(T, torch.Tensor):
        t = T.detach()
        if t.dim() == 2 and tuple(t.shape) in ((3, 4), (4, 4)):
            return SE3(t.cpu().numpy())
        if t.dim() == 3 and t.shape[1:] in ((3, 4), (4, 4)):
            return t[:, :3, :].to(torch.float32).reshape(t.shape[0], 12).contiguous().clone()
        if t.dim() == 2 and t.shape[1] == 12:
            return t.to(torch.float32).contiguous().clone()
        raise ValueError(f"cannot interpret a tensor of shape {tuple(t.shape)} as SE(3) targets")
    a = np.asarray(T, dtype=np.float64)
    if a.ndim == 2 and a.shape in ((3, 4), (4, 4)):
        return SE3(a)
    if a.ndim == 2 and a.shape[1] == 12:
        return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32)
    if a.ndim == 3 and a.shape[1:] in ((3, 4), (4, 4)):
        return torch.as_tensor(np.ascontiguousarray(a[:, :3, :].reshape(a.shape[0], 12)), dtype=torch.float32)
    raise ValueError(f"cannot interpret an array of shape {a.shape} as SE(3) targets")


def as_vector_target(x, n):
    """``[n]`` -> numpy copy (shared); ``[B, n]`` -> fp32 tensor copy (per instance)."""
    if isinstance(x, torch.Tensor):
        t = x.detach()
        if t.dim() == 1:
            return t.cpu().numpy().astype(np.float64).copy()
        return t.to(torch.float32).contiguous().clone()
    a = np.array(x, dtype=np.float64)
    if a.ndim == 1:
        if a.shape[0] != n:
            raise ValueError(f"target has {a.shape[0]} entries, expected {n}")
        return a
    if a.ndim == 2 and a.shape[1] == n:
        return torch.as_tensor(a, dtype=torch.float32)
    raise ValueError(f"cannot interpret an array of shape {a.shape} as [.., {n}] targets")
