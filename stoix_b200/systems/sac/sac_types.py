"""SACParams / SACOptStates -- same fields and order as stoix/systems/sac/sac_types.py:8-19; OnlineAndTarget and
OffPolicyLearnerState as stoix/base_types.py:131-137,152-154; Transition as stoix/systems/q_learning/dqn_types.py."""
from typing import Any, Dict, NamedTuple

import torch


class OnlineAndTarget(NamedTuple):
    online: Any
    target: Any


class SACParams(NamedTuple):
    actor_params: Any
    q_params: OnlineAndTarget
    log_alpha: torch.Tensor


class SACOptStates(NamedTuple):
    actor_opt_state: Any
    q_opt_state: Any
    alpha_opt_state: Any


class Transition(NamedTuple):
    obs: torch.Tensor
    action: torch.Tensor
    reward: torch.Tensor
    done: torch.Tensor
    next_obs: torch.Tensor
    info: Dict


class OffPolicyLearnerState(NamedTuple):
    params: Any
    opt_states: Any
    buffer_state: Any
    key: Any
    env_state: Any
    timestep: Any
