"""CartPole-v1 on the GPU (BASELINE.json configs[0]: the reference's CPU-runnable plumbing case,
`env=gymnax/cartpole`).  Classic-control physics as in gymnax 0.0.9 CartPole-v1 (Euler integration,
tau 0.02, force 10, |x| > 2.4 or |theta| > 12 deg terminates, 500-step time limit as truncation)
with the auto-reset / episode-metrics wrapper semantics of stoix/utils/make_env.py:29-61.

This environment is elementwise torch code, not a hand-written kernel: environments are plugins
behind the stoa interface and sit outside the accelerated path."""
from __future__ import annotations

import math
from typing import Any, Dict, Tuple

import torch

from .base import ArraySpace, DiscreteSpace, Environment, StepOut, StepType, TimeStep, timestep_from_out


class CartPoleEnv(Environment):
    gravity, masscart, masspole, length, force_mag, tau = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    theta_threshold = 12 * 2 * math.pi / 360
    x_threshold = 2.4
    max_steps = 500

    def __init__(self, device="cuda", seed: int = 0):
        self.device = torch.device(device)
        # learner_setup assigns `arch.seed + 7919 * rank + ...` per rank / shard before reset() (the reference splits its
        # reset keys per device, batch and env, ff_ppo.py:492-501); reset states are a pure function of (seed, env, draw #)
        self.seed = int(seed)

    def observation_space(self) -> ArraySpace:
        return ArraySpace((4,), torch.float32, self.device)

    def action_space(self) -> DiscreteSpace:
        return DiscreteSpace(2)

    _M32 = 0xFFFFFFFF

    def _fresh(self, E: int, ctr: torch.Tensor, seed: int) -> torch.Tensor:
        """U(-0.05, 0.05) reset states from a counter-based hash (murmur3 finaliser on 32-bit lanes held in int64) of
        (seed, env, component, draw counter).  Stateless: no torch.Generator, so it is CUDA-graph safe (the counter is a
        device tensor advanced inside the graph) and a run is a function of arch.seed."""
        M = self._M32
        idx = torch.arange(E * 4, dtype=torch.int64, device=self.device).view(E, 4)
        mix = ((seed & M) * 0x27D4EB2F + ((seed >> 32) & M) * 0x165667B1 + 0x9E3779B9) & M
        h = (idx * 0x9E3779B1 + ctr * 0x85EBCA77 + mix) & M
        h = h ^ (h >> 16)
        h = (h * 0x85EBCA6B) & M
        h = h ^ (h >> 13)
        h = (h * 0xC2B2AE35) & M
        h = h ^ (h >> 16)
        u = (h >> 8).to(torch.float32) * (1.0 / 16777216.0)
        return (u - 0.5) * 0.1

    def reset(self, keys) -> Tuple[Dict[str, Any], TimeStep]:
        E, dev = len(keys), self.device
        seed = int(self.seed)
        ctr = torch.zeros((), dtype=torch.int64, device=dev)
        phys = self._fresh(E, ctr, seed)
        ctr += 1
        state = {"phys": phys, "reset_ctr": ctr, "seed": seed, "time": torch.zeros(E, dtype=torch.int32, device=dev),
                 "run_return": torch.zeros(E, device=dev), "run_length": torch.zeros(E, dtype=torch.int32, device=dev)}
        ts = TimeStep(torch.full((E,), StepType.FIRST, dtype=torch.int8, device=dev), torch.zeros(E, device=dev),
                      torch.ones(E, device=dev), phys.clone(),
                      {"next_obs": phys.clone(),
                       "episode_metrics": {"episode_return": torch.zeros(E, device=dev),
                                           "episode_length": torch.zeros(E, dtype=torch.int32, device=dev),
                                           "is_terminal_step": torch.zeros(E, dtype=torch.bool, device=dev)}})
        return state, ts

    def step(self, state, action) -> Tuple[Dict[str, Any], TimeStep]:
        E, dev = state["phys"].shape[0], self.device
        out = StepOut(torch.empty(E, 4, device=dev), torch.empty(E, 4, device=dev), torch.empty(E, device=dev),
                      torch.empty(E, dtype=torch.uint8, device=dev), torch.empty(E, dtype=torch.uint8, device=dev),
                      torch.empty(E, device=dev), torch.empty(E, dtype=torch.int32, device=dev),
                      torch.empty(E, dtype=torch.uint8, device=dev))
        self.step_into(state, action, out, 0)
        return state, timestep_from_out(out)

    def step_into(self, state, action: torch.Tensor, out: StepOut, t: int) -> None:
        x, x_dot, th, th_dot = state["phys"].unbind(-1)
        force = torch.where(action > 0, self.force_mag, -self.force_mag).to(torch.float32)
        cos, sin = torch.cos(th), torch.sin(th)
        total_mass = self.masscart + self.masspole
        pml = self.masspole * self.length
        temp = (force + pml * th_dot * th_dot * sin) / total_mass
        th_acc = (self.gravity * sin - cos * temp) / (self.length * (4.0 / 3.0 - self.masspole * cos * cos / total_mass))
        x_acc = temp - pml * th_acc * cos / total_mass
        nxt = torch.stack([x + self.tau * x_dot, x_dot + self.tau * x_acc, th + self.tau * th_dot, th_dot + self.tau * th_acc], -1)
        time = state["time"] + 1
        term = (nxt[:, 0].abs() > self.x_threshold) | (nxt[:, 2].abs() > self.theta_threshold)
        trunc = (~term) & (time >= self.max_steps)
        last = term | trunc
        ret = state["run_return"] + 1.0
        ln = state["run_length"] + 1
        out.next_obs.copy_(nxt)
        new_phys = torch.where(last[:, None], self._fresh(nxt.shape[0], state["reset_ctr"], state["seed"]), nxt)
        state["reset_ctr"] += 1
        out.obs.copy_(new_phys)
        out.reward.fill_(1.0)
        out.done.copy_(term)
        out.truncated.copy_(trunc)
        out.episode_return.copy_(ret)
        out.episode_length.copy_(ln)
        out.is_terminal_step.copy_(last)
        state["phys"].copy_(new_phys)
        state["time"].copy_(torch.where(last, torch.zeros_like(time), time))
        state["run_return"].copy_(torch.where(last, torch.zeros_like(ret), ret))
        state["run_length"].copy_(torch.where(last, torch.zeros_like(ln), ln))

    def advance(self, state, steps: int) -> None:
        return None
