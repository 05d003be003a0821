"""Kinematic limits (``/root/reference/pink/limits/__init__.py``).

:class:`ConfigurationLimit`, :class:`VelocityLimit` and :class:`AccelerationLimit`
produce ``+-e_i`` rows (a box on the displacement);
:class:`FloatingBaseVelocityLimit` produces dense rows on the base twist.
"""

from .acceleration_limit import AccelerationLimit
from .configuration_limit import ConfigurationLimit
from .floating_base_velocity_limit import FloatingBaseVelocityLimit
from .limit import Limit
from .velocity_limit import VelocityLimit

__all__ = [cls.__name__ for cls in (
    AccelerationLimit, ConfigurationLimit,
    FloatingBaseVelocityLimit, Limit, VelocityLimit,
)]
