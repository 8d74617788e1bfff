from .base import Environment, StepType, TimeStep  # noqa: F401
