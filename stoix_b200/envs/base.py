"""The env-wrapper interface the learner consumes (stoa's Environment / TimeStep as used at
stoix/systems/ppo/anakin/ff_ppo.py:104-116,432,466,495 and stoix/utils/make_env.py:29-61), batched
natively on the GPU (the reference's VmapWrapper is implicit: every tensor has a leading env axis).

  env.reset(keys)            -> (state, TimeStep)       len(keys) envs
  env.step(state, action)    -> (state, TimeStep)
  TimeStep.last()            -> bool tensor
  TimeStep.extras["next_obs"], extras["episode_metrics"] = {episode_return, episode_length,
                                                            is_terminal_step}
  env.observation_space().generate_value(), env.action_space().num_values

Optional fast path used by the learner when present: `env.step_into(state, action, out, t)` writes
the step's outputs straight into rows of the trajectory buffers (no intermediate copies)."""
from __future__ import annotations

from typing import Any, Dict, NamedTuple, Tuple

import torch


class StepType:
    FIRST = 0
    MID = 1
    LAST = 2


class TimeStep(NamedTuple):
    step_type: torch.Tensor
    reward: torch.Tensor
    discount: torch.Tensor
    observation: torch.Tensor
    extras: Dict[str, Any]

    def first(self) -> torch.Tensor:
        return self.step_type == StepType.FIRST

    def last(self) -> torch.Tensor:
        return self.step_type == StepType.LAST


class ArraySpace:
    def __init__(self, shape: Tuple[int, ...], dtype=torch.float32, device="cpu"):
        self.shape, self.dtype, self.device = tuple(shape), dtype, device

    def generate_value(self) -> torch.Tensor:
        return torch.zeros(self.shape, dtype=self.dtype, device=self.device)


class DiscreteSpace:
    def __init__(self, num_values: int):
        self.num_values = int(num_values)
        self.shape = ()
        self.dtype = torch.int32

    def generate_value(self) -> torch.Tensor:
        return torch.zeros((), dtype=torch.int32)


class StepOut(NamedTuple):
    """Destination rows for Environment.step_into (all (E, ...) views into trajectory buffers)."""

    obs: torch.Tensor          # observation after the step (reset obs where the episode ended)
    next_obs: torch.Tensor     # extras["next_obs"]: true successor
    reward: torch.Tensor
    done: torch.Tensor         # uint8: discount == 0            (ff_ppo.py:107)
    truncated: torch.Tensor    # uint8: last() & discount != 0   (ff_ppo.py:108)
    episode_return: torch.Tensor
    episode_length: torch.Tensor
    is_terminal_step: torch.Tensor


class Environment:
    def reset(self, keys) -> Tuple[Any, TimeStep]:
        raise NotImplementedError

    def step(self, state, action) -> Tuple[Any, TimeStep]:
        raise NotImplementedError

    def observation_space(self) -> ArraySpace:
        raise NotImplementedError

    def action_space(self) -> DiscreteSpace:
        raise NotImplementedError


def timestep_from_out(out: StepOut) -> TimeStep:
    last = out.is_terminal_step.bool()
    step_type = torch.where(last, StepType.LAST, StepType.MID).to(torch.int8)
    discount = 1.0 - out.done.to(torch.float32)
    extras = {
        "next_obs": out.next_obs,
        "episode_metrics": {
            "episode_return": out.episode_return,
            "episode_length": out.episode_length,
            "is_terminal_step": last,
        },
    }
    return TimeStep(step_type, out.reward, discount, out.obs, extras)
