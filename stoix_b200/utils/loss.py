"""Drop-in for the two hot functions of stoix/utils/loss.py (forward values).

The training step does not call these (the fused K3 kernel evaluates the same arithmetic together
with its gradient); they keep `stoix.utils.loss` call sites working on CUDA tensors.
"""
from __future__ import annotations

import torch

from .. import ops


def _flat(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float32).reshape(-1).contiguous()


def ppo_clip_loss(pi_log_prob_t: torch.Tensor, b_pi_log_prob_t: torch.Tensor, gae_t: torch.Tensor, epsilon: float) -> torch.Tensor:
    """stoix/utils/loss.py:17-32: -mean(min(ratio*A, clip(ratio, 1-eps, 1+eps)*A))."""
    return ops.ppo_clip_loss_value(_flat(pi_log_prob_t), _flat(b_pi_log_prob_t), _flat(gae_t), epsilon)


def clipped_value_loss(pred_value_t: torch.Tensor, behavior_value_t: torch.Tensor, targets_t: torch.Tensor, epsilon: float) -> torch.Tensor:
    """stoix/utils/loss.py:68-78: 0.5*mean(max((v-tgt)^2, (v_clip-tgt)^2))."""
    return ops.clipped_value_loss_value(_flat(pred_value_t), _flat(behavior_value_t), _flat(targets_t), epsilon)
