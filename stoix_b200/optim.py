"""optax-shaped facade over the fused optimiser kernel (stx_clip_adam_step).

`chain(clip_by_global_norm(m), adam(lr, eps=...))` mirrors the construction at
stoix/systems/ppo/anakin/ff_ppo.py:456-463.  The returned object's `.update` is what the reference
passes around as `update_fns`; the learner reads its hyper-parameters and runs BOTH optimisers
(actor and critic) in a single fused launch over the flat arenas."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, NamedTuple, Optional, Union


@dataclass
class _Clip:
    max_norm: float


@dataclass
class _Adam:
    learning_rate: Union[float, Callable]
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8
    eps_root: float = 0.0


def clip_by_global_norm(max_norm: float) -> _Clip:
    return _Clip(float(max_norm))


def adam(learning_rate, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8, eps_root: float = 0.0) -> _Adam:
    if eps_root != 0.0:
        raise NotImplementedError("eps_root != 0 is not used by the reference PPO and is not built")
    return _Adam(learning_rate, b1, b2, eps, eps_root)


class OptState(NamedTuple):
    """(ScaleByAdamState{count, mu, nu}, ScaleByScheduleState{count}) as views into the flat state."""

    count: Any
    mu: Any
    nu: Any
    sched_count: Any


class GradientTransformation:
    def __init__(self, clip: Optional[_Clip], adam_: _Adam):
        self.clip, self.adam = clip, adam_

    @property
    def max_grad_norm(self) -> float:
        return self.clip.max_norm if self.clip is not None else float("inf")

    @property
    def init_lr(self) -> float:
        lr = self.adam.learning_rate
        return float(getattr(lr, "init_lr", lr))

    @property
    def schedule(self) -> Optional[Callable]:
        lr = self.adam.learning_rate
        return lr if callable(lr) else None

    def init(self, params) -> None:
        """Optimiser state lives in the learner's flat mu/nu/count arenas (allocated in learner_setup)."""
        return None

    def update(self, *args, **kwargs):
        raise RuntimeError(
            "stoix_b200 optimisers are applied by the fused stx_clip_adam_step kernel inside the learner; "
            "the `.update` handle only carries the hyper-parameters"
        )


def chain(*parts) -> GradientTransformation:
    clip = next((p for p in parts if isinstance(p, _Clip)), None)
    adam_ = next((p for p in parts if isinstance(p, _Adam)), None)
    if adam_ is None:
        raise ValueError("chain() needs an adam() stage")
    return GradientTransformation(clip, adam_)
