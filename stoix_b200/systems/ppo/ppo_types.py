"""PPOTransition -- same fields and order as stoix/systems/ppo/ppo_types.py:9-20."""
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
