"""Control Barrier Functions (``/root/reference/pink/barriers/__init__.py``)."""

from .barrier import Barrier
from .body_spherical_barrier import BodySphericalBarrier
from .position_barrier import PositionBarrier
from .self_collision_barrier import SelfCollisionBarrier

__all__ = [cls.__name__ for cls in (Barrier, PositionBarrier, BodySphericalBarrier, SelfCollisionBarrier)]
