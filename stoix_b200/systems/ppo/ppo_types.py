"""PPOTransition, RNNPPOTransition, ActorCriticHiddenStates -- same fields and order as stoix/systems/ppo/ppo_types.py."""
from typing import Dict, NamedTuple

import torch


class PPOTransition(NamedTuple):
    done: torch.Tensor
    truncated: torch.Tensor
    action: torch.Tensor
    value: torch.Tensor
    reward: torch.Tensor
    bootstrap_value: torch.Tensor
    log_prob: torch.Tensor
    obs: torch.Tensor
    info: Dict


class ActorCriticHiddenStates(NamedTuple):
    policy_hidden_state: torch.Tensor
    critic_hidden_state: torch.Tensor


class RNNPPOTransition(NamedTuple):
    """done / truncated are the flags BEFORE the step, hstates the hidden states AFTER it (rec_ppo.py:118-129)."""

    done: torch.Tensor
    truncated: torch.Tensor
    action: torch.Tensor
    value: torch.Tensor
    reward: torch.Tensor
    log_prob: torch.Tensor
    obs: torch.Tensor
    hstates: ActorCriticHiddenStates
    info: Dict


class RNNLearnerState(NamedTuple):
    """stoix/base_types.py RNNLearnerState: the on-policy learner state plus the carried flags and hidden states."""

    params: object
    opt_states: object
    key: object
    env_state: object
    timestep: object
    done: torch.Tensor
    truncated: torch.Tensor
    hstates: ActorCriticHiddenStates
