"""Exceptions with the names of ``/root/reference/pink/exceptions.py:11-120``."""


class PinkError(Exception):
    """Base class for Pink exceptions."""


class ConfigurationError(PinkError):
    """Exception raised when encountering an invalid configuration vector."""


class FrameNotFound(PinkError):
    """Exception raised when a frame is not found in the robot model."""

    def __init__(self, name, frames):
        self.message = f"Frame '{name}' not found in: {[getattr(f, 'name', f) for f in frames]}"
        super().__init__(self.message)


class InvalidCollisionPairs(PinkError):
    """Exception raised when the number of collision pairs is invalid."""


class NegativeMinimumDistance(PinkError):
    """Exception raised when the minimum distance threshold is negative."""


class NoPositionLimitProvided(PinkError):
    """Exception raised when neither a minimum nor a maximum position limit is provided."""


class NoSolutionFound(PinkError):
    """The QP solver did not find a solution to the differential IK problem.

    Batched extension: ``instances`` lists the failing rows of the batch."""

    def __init__(self, problem=None, results=None, instances=None):
        msg = "QP solver did not find a solution to the differential IK problem"
        if instances is not None:
            msg += f" for {len(instances)} instance(s), first: {list(instances[:8])}"
        super().__init__(msg)
        self.problem = problem
        self.results = results
        self.instances = instances


class NotWithinConfigurationLimits(PinkError):
    """Exception thrown when a robot configuration violates its limits."""

    def __init__(self, joint, value, lower, upper, instance=None):
        self.joint = joint
        self.value = value
        self.lower = lower
        self.upper = upper
        self.instance = instance
        where = "" if instance is None else f" (batch instance {instance})"
        self.message = (
            f"Joint {joint} violates configuration limits "
            f"{lower} <= {value} <= {upper}{where}"
        )
        super().__init__(self.message)


class TargetNotSet(PinkError):
    """Exception raised when attempting to compute with an unset target."""


class TaskDefinitionError(PinkError):
    """Exception raised when a task definition is ill-formed."""


class TaskJacobianNotSet(PinkError):
    """Exception raised when attempting to compute without a task Jacobian."""
