"""CPU-side vectorised environments for the Sebulba architecture (actors step environments on host threads).

The reference wraps envpool / gymnasium vector envs into a stateful stoa-style API (stoix/wrappers/envpool.py:8-164,
stoix/utils/env_factory.py): `envs.reset(seed=...) -> TimeStep`, `envs.step(np_action) -> TimeStep` with NumPy leaves,
auto-reset inside `step`, `extras["metrics"] = {episode_return, episode_length, is_terminal_step}` (envpool.py:94-133),
`discount = 1 - terminated` except on truncation (envpool.py:135-150).  envpool / gymnasium are not installable here;
the two environments BASELINE names are provided in NumPy behind that same interface:

  SyntheticBoxCpuEnv   the synthetic Box workload (obs ~ N(0,1), reward ~ N(0,1), Bernoulli termination / truncation)
  CartPoleCpuEnv       CartPole-v1 physics (gymnasium semantics), vectorised

and `make_factory(config)` returns the EnvFactory the actor threads call with their env count (env_factory.py:22-58:
a lock hands every actor a distinct seed range).
"""
from __future__ import annotations

import math
import threading
from typing import Any, Dict, NamedTuple, Optional

import numpy as np


class StepType:
    FIRST, MID, TERMINATED, TRUNCATED = 0, 1, 2, 3


class TimeStep(NamedTuple):
    step_type: np.ndarray
    reward: np.ndarray
    discount: np.ndarray
    observation: np.ndarray
    extras: Dict[str, Any]

    def last(self) -> np.ndarray:
        return self.step_type >= StepType.TERMINATED


class _Space:
    def __init__(self, shape=(), num_values: Optional[int] = None):
        self.shape, self.num_values = tuple(shape), num_values


class _VecEnvBase:
    num_envs: int
    obs_shape: tuple
    num_actions: int

    def _metrics_reset(self) -> Dict[str, np.ndarray]:
        self._run_ret = np.zeros(self.num_envs, np.float64)
        self._run_len = np.zeros(self.num_envs, np.int64)
        self._ep_ret = np.zeros(self.num_envs, np.float64)
        self._ep_len = np.zeros(self.num_envs, np.int64)
        return {"episode_return": np.zeros(self.num_envs), "episode_length": np.zeros(self.num_envs, np.int64),
                "is_terminal_step": np.zeros(self.num_envs, bool)}

    def _metrics_step(self, reward: np.ndarray, ep_done: np.ndarray) -> Dict[str, np.ndarray]:
        """envpool.py:94-133 (no lives): publish the finished episode's totals on its last step, keep the previous ones otherwise."""
        new_ret, new_len = self._run_ret + reward, self._run_len + 1
        self._ep_ret = np.where(ep_done, new_ret, self._ep_ret)
        self._ep_len = np.where(ep_done, new_len, self._ep_len)
        self._run_ret = np.where(ep_done, 0.0, new_ret)
        self._run_len = np.where(ep_done, 0, new_len)
        return {"episode_return": self._ep_ret.copy(), "episode_length": self._ep_len.copy(), "is_terminal_step": ep_done.copy()}

    @staticmethod
    def _timestep(obs, reward, terminated, truncated, metrics) -> TimeStep:
        ep_done = terminated | truncated
        step_type = np.where(ep_done, StepType.TERMINATED, StepType.MID)
        step_type = np.where(truncated, StepType.TRUNCATED, step_type).astype(np.int8)
        discount = np.where(truncated, 1.0, 1.0 - terminated.astype(np.float32)).astype(np.float32)   # envpool.py:143-146
        return TimeStep(step_type, reward.astype(np.float32), discount, obs, {"metrics": metrics})

    def observation_space(self) -> _Space:
        return _Space(self.obs_shape)

    def action_space(self) -> _Space:
        return _Space((), self.num_actions)

    def close(self) -> None:
        pass


class SyntheticBoxCpuEnv(_VecEnvBase):
    def __init__(self, num_envs: int, obs_dim: int = 64, num_actions: int = 8, seed: int = 0, p_term: float = 0.005, p_trunc: float = 0.002):
        self.num_envs, self.obs_shape, self.num_actions = int(num_envs), (int(obs_dim),), int(num_actions)
        self.p_term, self.p_trunc = float(p_term), float(p_trunc)
        self._rng = np.random.default_rng(seed)

    def reset(self, *, seed=None, options=None) -> TimeStep:
        if seed is not None:
            self._rng = np.random.default_rng(seed)
        obs = self._rng.standard_normal((self.num_envs,) + self.obs_shape, dtype=np.float32)
        z = np.zeros(self.num_envs, bool)
        ts = self._timestep(obs, np.zeros(self.num_envs, np.float32), z, z, self._metrics_reset())
        return ts._replace(step_type=np.full(self.num_envs, StepType.FIRST, np.int8))

    def step(self, action) -> TimeStep:
        action = np.asarray(action)
        assert action.shape == (self.num_envs,), action.shape
        r = self._rng
        obs = r.standard_normal((self.num_envs,) + self.obs_shape, dtype=np.float32)   # auto-reset: a fresh draw either way
        reward = r.standard_normal(self.num_envs, dtype=np.float32)
        terminated = r.random(self.num_envs) < self.p_term
        truncated = (~terminated) & (r.random(self.num_envs) < self.p_trunc)
        return self._timestep(obs, reward, terminated, truncated, self._metrics_step(reward.astype(np.float64), terminated | truncated))


class CartPoleCpuEnv(_VecEnvBase):
    gravity, masscart, masspole, length, force_mag, tau = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    theta_threshold = 12 * 2 * math.pi / 360
    x_threshold, max_steps = 2.4, 500

    def __init__(self, num_envs: int, seed: int = 0):
        self.num_envs, self.obs_shape, self.num_actions = int(num_envs), (4,), 2
        self._rng = np.random.default_rng(seed)

    def _fresh(self, n: int) -> np.ndarray:
        return self._rng.uniform(-0.05, 0.05, (n, 4)).astype(np.float32)

    def reset(self, *, seed=None, options=None) -> TimeStep:
        if seed is not None:
            self._rng = np.random.default_rng(seed)
        self._phys = self._fresh(self.num_envs)
        self._time = np.zeros(self.num_envs, np.int64)
        z = np.zeros(self.num_envs, bool)
        ts = self._timestep(self._phys.copy(), np.zeros(self.num_envs, np.float32), z, z, self._metrics_reset())
        return ts._replace(step_type=np.full(self.num_envs, StepType.FIRST, np.int8))

    def step(self, action) -> TimeStep:
        action = np.asarray(action)
        x, x_dot, th, th_dot = self._phys.T.astype(np.float64)
        force = np.where(action > 0, self.force_mag, -self.force_mag)
        cos, sin = np.cos(th), np.sin(th)
        total_mass, pml = self.masscart + self.masspole, self.masspole * self.length
        temp = (force + pml * th_dot * th_dot * sin) / total_mass
        th_acc = (self.gravity * sin - cos * temp) / (self.length * (4.0 / 3.0 - self.masspole * cos * cos / total_mass))
        x_acc = temp - pml * th_acc * cos / total_mass
        nxt = np.stack([x + self.tau * x_dot, x_dot + self.tau * x_acc, th + self.tau * th_dot, th_dot + self.tau * th_acc], -1).astype(np.float32)
        self._time += 1
        terminated = (np.abs(nxt[:, 0]) > self.x_threshold) | (np.abs(nxt[:, 2]) > self.theta_threshold)
        truncated = (~terminated) & (self._time >= self.max_steps)
        ep_done = terminated | truncated
        reward = np.ones(self.num_envs, np.float32)
        if ep_done.any():
            nxt[ep_done] = self._fresh(int(ep_done.sum()))
            self._time[ep_done] = 0
        self._phys = nxt
        return self._timestep(nxt.copy(), reward, terminated, truncated, self._metrics_step(reward.astype(np.float64), ep_done))


class EnvFactory:
    """env_factory.py:22-58: thread-safe factory; every call draws a distinct seed range."""

    def __init__(self, make_fn, init_seed: int = 42, **kwargs: Any):
        self.make_fn, self.seed, self.kwargs = make_fn, int(init_seed), kwargs
        self.lock = threading.Lock()

    def __call__(self, num_envs: int) -> Any:
        with self.lock:
            seed = self.seed
            self.seed += num_envs
        return self.make_fn(num_envs=num_envs, seed=seed, **self.kwargs)


def make_factory(config) -> EnvFactory:
    """stoix/utils/make_env.py make_factory: the Sebulba env factory for config.env."""
    name = config.env.env_name
    seed = int(config.arch.seed)
    if name == "synthetic":
        return EnvFactory(SyntheticBoxCpuEnv, seed, **dict(config.env.kwargs))
    if name == "gymnax" and config.env.scenario.name == "CartPole-v1":
        return EnvFactory(CartPoleCpuEnv, seed)
    raise NotImplementedError(f"no CPU environment factory for env '{name}'")
