"""SyntheticContinuousEnv -- the continuous-control workload named by BASELINE.json configs[3] (ff_sac: "synthetic Box
obs_dim=17 act_dim=6"): Box observations ~ N(0,1)^obs_dim, actions in [-1, 1]^act_dim, Bernoulli termination /
truncation, auto-reset + episode metrics like the other environments (stoix/utils/make_env.py:29-61).  The observation
stream ignores the action (as in SyntheticBoxEnv); the REWARD does not: r = -mean((a - tanh(obs[:act_dim]))^2), a
contextual bandit whose optimum an actor can learn, so that learning curves are meaningful for the off-policy systems.

Elementwise torch code, like CartPoleEnv: environments are plugins behind the stoa interface and sit outside the
accelerated path; randomness is a counter-based hash of (seed, env, draw), so a run is a function of arch.seed."""
from __future__ import annotations

from typing import Any, Dict, Tuple

import torch

from .base import ArraySpace, Environment, StepType, TimeStep


class BoundedArraySpace(ArraySpace):
    def __init__(self, shape, minimum: float, maximum: float, dtype=torch.float32, device="cpu"):
        super().__init__(shape, dtype, device)
        self.minimum, self.maximum = float(minimum), float(maximum)


class SyntheticContinuousEnv(Environment):
    _M32 = 0xFFFFFFFF

    def __init__(self, obs_dim: int = 17, act_dim: int = 6, p_term: float = 1.0 / 200, p_trunc: float = 1.0 / 500, seed: int = 42, device="cuda"):
        self.obs_dim, self.act_dim = int(obs_dim), int(act_dim)
        self.p_term, self.p_trunc, self.seed = float(p_term), float(p_trunc), int(seed)
        self.device = torch.device(device)

    def observation_space(self) -> ArraySpace:
        return ArraySpace((self.obs_dim,), torch.float32, self.device)

    def action_space(self) -> BoundedArraySpace:
        return BoundedArraySpace((self.act_dim,), -1.0, 1.0, torch.float32, self.device)

    def _u01(self, n: int, cols: int, ctr: torch.Tensor, seed: int, tag: int) -> torch.Tensor:
        """(n, cols) uniforms in (0, 1): murmur3 finaliser over 32-bit lanes held in int64 (see CartPoleEnv._fresh)."""
        M = self._M32
        idx = torch.arange(n * cols, dtype=torch.int64, device=self.device).view(n, cols)
        mix = ((seed & M) * 0x27D4EB2F + ((seed >> 32) & M) * 0x165667B1 + tag * 0x9E3779B9) & M
        h = (idx * 0x9E3779B1 + ctr * 0x85EBCA77 + mix) & M
        h = h ^ (h >> 16)
        h = (h * 0x85EBCA6B) & M
        h = h ^ (h >> 13)
        h = (h * 0xC2B2AE35) & M
        h = h ^ (h >> 16)
        return ((h >> 9).to(torch.float32) + 0.5) * (1.0 / 8388608.0)

    def _normal(self, n: int, cols: int, ctr: torch.Tensor, seed: int, tag: int) -> torch.Tensor:
        u1, u2 = self._u01(n, cols, ctr, seed, tag), self._u01(n, cols, ctr, seed, tag + 1)
        return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(6.283185307179586 * u2)

    def reset(self, keys) -> Tuple[Dict[str, Any], TimeStep]:
        E, dev = len(keys), self.device
        seed = int(self.seed)
        ctr = torch.zeros((), dtype=torch.int64, device=dev)
        obs = self._normal(E, self.obs_dim, ctr, seed, 1)
        ctr += 1
        state = {"obs": obs, "ctr": ctr, "seed": seed, "run_return": torch.zeros(E, device=dev), "run_length": torch.zeros(E, dtype=torch.int32, device=dev)}
        ts = TimeStep(torch.full((E,), StepType.FIRST, dtype=torch.int8, device=dev), torch.zeros(E, device=dev), torch.ones(E, device=dev), obs.clone(),
                      {"next_obs": obs.clone(), "episode_metrics": {"episode_return": torch.zeros(E, device=dev),
                                                                    "episode_length": torch.zeros(E, dtype=torch.int32, device=dev),
                                                                    "is_terminal_step": torch.zeros(E, dtype=torch.bool, device=dev)}})
        return state, ts

    def step(self, state, action: torch.Tensor) -> Tuple[Dict[str, Any], TimeStep]:
        E, A = state["obs"].shape[0], self.act_dim
        obs, ctr, seed = state["obs"], state["ctr"] + 0, state["seed"]
        reward = -((action.float() - torch.tanh(obs[:, :A])) ** 2).mean(-1)
        nxt = self._normal(E, self.obs_dim, ctr, seed, 1)          # true successor (a fresh draw)
        rst = self._normal(E, self.obs_dim, ctr, seed, 3)          # observation after an auto-reset
        u = self._u01(E, 2, ctr, seed, 5)
        term = u[:, 0] < self.p_term
        trunc = (~term) & (u[:, 1] < self.p_trunc)
        last = term | trunc
        ret, ln = state["run_return"] + reward, state["run_length"] + 1
        new_obs = torch.where(last[:, None], rst, nxt)
        # in place on the state tensors (fixed addresses: the step can sit inside a captured CUDA graph, like CartPoleEnv)
        state["obs"].copy_(new_obs)
        state["ctr"].add_(1)
        state["run_return"].copy_(torch.where(last, torch.zeros_like(ret), ret))
        state["run_length"].copy_(torch.where(last, torch.zeros_like(ln), ln))
        step_type = torch.where(last, StepType.LAST, StepType.MID).to(torch.int8)
        ts = TimeStep(step_type, reward, 1.0 - term.float(), new_obs,
                      {"next_obs": nxt, "episode_metrics": {"episode_return": ret, "episode_length": ln, "is_terminal_step": last}})
        return state, ts
