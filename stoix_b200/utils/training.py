"""Learning-rate helpers -- stoix/utils/training.py:6-53."""
from __future__ import annotations

from typing import Callable, Optional, Union


def make_learning_rate_schedule(init_lr: float, num_updates: int, num_epochs: int, num_minibatches: int) -> Callable:
    """Linear decay indexed by optimiser-step count with floor division (training.py:24-26).  The
    returned callable also carries the constants so the fused Adam kernel can evaluate the same
    schedule on the device."""

    def linear_scedule(count: int) -> float:
        frac = 1.0 - (count // (num_epochs * num_minibatches)) / num_updates
        return init_lr * frac

    linear_scedule.init_lr = float(init_lr)
    linear_scedule.num_updates = int(num_updates)
    linear_scedule.steps_per_update = int(num_epochs * num_minibatches)
    return linear_scedule


def make_learning_rate(init_lr: float, config, num_epochs: int, num_minibatches: Optional[int] = None) -> Union[float, Callable]:
    if num_minibatches is None:
        num_minibatches = 1
    if config.system.decay_learning_rates:
        return make_learning_rate_schedule(init_lr, config.arch.num_updates, num_epochs, num_minibatches)
    return init_lr
