"""CategoricalHead / NormalAffineTanhDistributionHead / ScalarCriticHead -- constructor-compatible with
stoix/networks/heads.py:30-41, 44-65, 129-134, and the distribution objects the actor returns (tfd.Categorical's sample /
log_prob / entropy / mode as used at ff_ppo.py:100-101,199,205; the tanh-Normal's sample / log_prob / mode as used at
ff_sac.py:119-122,165-167)."""
from __future__ import annotations

from typing import Optional, Sequence, Union

import numpy as np
import torch

from .. import ops


class Categorical:
    """Categorical(logits) backed by stx_categorical."""

    def __init__(self, logits: torch.Tensor):
        self.logits = logits
        self._lead = tuple(logits.shape[:-1])                     # any leading shape, e.g. (T, E) from the recurrent networks
        self._flat = logits.reshape(-1, logits.shape[-1]).contiguous()

    def sample(self, seed=None, offset: int = 0, dev_counter: Optional[torch.Tensor] = None) -> torch.Tensor:
        """`seed` is an int or a PRNG key tensor (any integer tensor: its first two words are used)."""
        s = _seed_to_int(seed)
        a, lp, _ = ops.categorical(self._flat, None, s, offset, dev_counter)
        a, lp = a.view(self._lead), lp.view(self._lead)
        self._last = (a, lp)
        return a

    def log_prob(self, action: torch.Tensor) -> torch.Tensor:
        last = getattr(self, "_last", None)
        if last is not None and last[0] is action:
            return last[1]
        _, lp, _ = ops.categorical(self._flat, action.to(torch.int32).reshape(-1).contiguous())
        return lp.view(self._lead)

    def entropy(self) -> torch.Tensor:
        a = torch.zeros(self._flat.shape[0], dtype=torch.int32, device=self.logits.device)
        _, _, ent = ops.categorical(self._flat, a, want_entropy=True)
        return ent.view(self._lead)

    def mode(self) -> torch.Tensor:
        return torch.argmax(self.logits, dim=-1).to(torch.int32)


def _seed_to_int(seed) -> int:
    if seed is None:
        return 0
    if isinstance(seed, torch.Tensor):
        w = seed.reshape(-1).to(torch.int64).cpu().tolist()
        return ((w[0] & 0xFFFFFFFF) << 32 | (w[-1] & 0xFFFFFFFF)) & (2**64 - 1)
    return int(seed) & (2**64 - 1)


class CategoricalHead:
    def __init__(self, action_dim: Union[int, Sequence[int]], kernel_init: Optional[float] = None):
        if not isinstance(action_dim, (int, np.integer)):
            raise NotImplementedError("factorised action_dim is outside the B200 hot path")
        self.action_dim = int(action_dim)
        self.out_dim = self.action_dim
        self.kernel_init_scale = 0.01 if kernel_init is None else float(kernel_init)  # heads.py:32

    def distribution(self, logits: torch.Tensor) -> Categorical:
        return Categorical(logits)


class ScalarCriticHead:
    def __init__(self, kernel_init: Optional[float] = None):
        self.out_dim = 1
        self.kernel_init_scale = 1.0 if kernel_init is None else float(kernel_init)  # heads.py:130


class AffineTanhNormal:
    """Independent(AffineTanhTransformedDistribution(Normal(loc, scale), minimum, maximum)) (stoix/networks/distributions.py:19-79)
    backed by stx_tanh_normal_sample.  `head_out` = (rows, 2A): loc | scale pre-activation."""

    def __init__(self, head_out: torch.Tensor, minimum: float, maximum: float, min_scale: float):
        self.head_out, self.minimum, self.maximum, self.min_scale = head_out, float(minimum), float(maximum), float(min_scale)
        self._last = None

    def sample(self, seed=None, offset: int = 0, dev_counter: Optional[torch.Tensor] = None) -> torch.Tensor:
        a, lp, eps = ops.tanh_normal_sample(self.head_out, self.minimum, self.maximum, self.min_scale, seed=_seed_to_int(seed), offset=offset,
                                            dev_counter=dev_counter)
        self._last = (a, lp, eps)
        return a

    def log_prob(self, action: torch.Tensor) -> torch.Tensor:
        if self._last is None or self._last[0] is not action:
            raise NotImplementedError("log_prob is available for the action this distribution object has just sampled "
                                      "(the only use in the SAC systems, ff_sac.py:165-167, 187-188, 217-218)")
        return self._last[1]

    def mode(self) -> torch.Tensor:
        A = self.head_out.shape[-1] // 2
        s, sh = (self.maximum - self.minimum) / 2.0, (self.minimum + self.maximum) / 2.0
        return sh + s * torch.tanh(self.head_out[..., :A])


class NormalAffineTanhDistributionHead:
    """heads.py:44-65: loc = Dense(A), scale = softplus(Dense(A)) + min_scale, both orthogonal(0.01); the two Dense layers are
    ONE (in x 2A) matrix in the arena: columns [0, A) = loc (flax `Dense_0`), [A, 2A) = scale (`Dense_1`)."""

    def __init__(self, action_dim: int, minimum: float, maximum: float, min_scale: float = 1e-3, kernel_init: Optional[float] = None):
        self.action_dim, self.minimum, self.maximum, self.min_scale = int(action_dim), float(minimum), float(maximum), float(min_scale)
        self.out_dim = 2 * self.action_dim
        self.kernel_init_scale = 0.01 if kernel_init is None else float(kernel_init)
        self.kernel_blocks = (self.action_dim, self.action_dim)   # initialised as two independent orthogonal matrices

    def distribution(self, head_out: torch.Tensor) -> AffineTanhNormal:
        return AffineTanhNormal(head_out, self.minimum, self.maximum, self.min_scale)
