"""Kinematic limits (``/root/reference/pink/limits/__init__.py``).

On the hot path: :class:`ConfigurationLimit`, :class:`VelocityLimit`.
``FloatingBaseVelocityLimit`` and ``AccelerationLimit`` are SURVEY section 8(f) "next"
rows and are not provided yet.
"""

from .configuration_limit import ConfigurationLimit
from .limit import Limit
from .velocity_limit import VelocityLimit

__all__ = ["ConfigurationLimit", "Limit", "VelocityLimit"]
