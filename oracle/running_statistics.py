"""CPU oracle for the observation-normalisation branch of ff_ppo  --  TEST INFRASTRUCTURE, NOT PRODUCT.

The "next" row 2 of SURVEY.md section 8(f): running mean / std of the raw observations, updated once per update step
from the whole rollout and reduced over the `device` and `batch` axes, used to normalise observations before both
networks (stoix/systems/ppo/anakin/ff_ppo.py:90-94, 113-115, 145-162; stoix/utils/running_statistics.py:123-135,
204-345, 348-363).  NumPy restatement of that arithmetic for the kernels of the next round; the product does not
implement the branch yet (`normalize_observations=True` raises).

**Parity unpinned**: the reference has no test for running_statistics.py and JAX is not installable here; the
restatement follows the source lines cited on each function and is checked for self-consistency against NumPy batch
statistics in tests/test_oracle_running_statistics.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np


@dataclass
class RunningStatistics:
    """running_statistics.py:94-100: count, mean, summed_variance, std (one leaf: the observation vector)."""

    count: float
    mean: np.ndarray
    summed_variance: np.ndarray
    std: np.ndarray


def initialize(feature_shape: Sequence[int], dtype=np.float64) -> RunningStatistics:
    """running_statistics.py:123-135: zero mean / variance, std ONE so that normalising with the initial state is the identity."""
    z = np.zeros(tuple(feature_shape), dtype)
    return RunningStatistics(0.0, z.copy(), z.copy(), np.ones(tuple(feature_shape), dtype))


def update(state: RunningStatistics, shards: Sequence[np.ndarray], std_min_value: float = 1e-6,
           std_max_value: float = 1e6) -> RunningStatistics:
    """One `update_statistics(..., pmap_axes=[...])` call (running_statistics.py:204-345).

    `shards`: the per-replica batches (leading dims = batch dims); the psum over the mapped axes (:62-70, 274-276,
    302, 308) is the sum over this list.  Batched Welford: the mean moves by sum(x - old_mean) / new_count, the summed
    variance grows by sum((x - old_mean) * (x - new_mean)); std = sqrt(clip(max(sv, 0) / count, min^2, max^2))
    clipped again to [min, max] (:329-340)."""
    nd = state.mean.ndim
    dt = state.mean.dtype
    flat = [np.asarray(s, dt).reshape((-1,) + state.mean.shape) if nd else np.asarray(s, dt).reshape(-1) for s in shards]
    count = state.count + float(sum(f.shape[0] for f in flat))                 # :270-278 (step_increment psum'd)
    mean_update = sum((f - state.mean).sum(axis=0) / count for f in flat)      # :301-302 (per replica, then psum)
    mean = state.mean + mean_update
    var_update = sum(((f - state.mean) * (f - mean)).sum(axis=0) for f in flat)  # :305-309
    summed_variance = state.summed_variance + var_update
    variance = np.clip(np.maximum(summed_variance, 0) / count, std_min_value**2, std_max_value**2)  # :333-337
    std = np.clip(np.sqrt(variance), std_min_value, std_max_value)            # :338-339
    return RunningStatistics(count, mean.astype(dt), summed_variance.astype(dt), std.astype(dt))


def normalize(x: np.ndarray, state: RunningStatistics, max_abs_value: Optional[float] = None) -> np.ndarray:
    """running_statistics.py:348-363: (x - mean) / std on inexact dtypes, optional symmetric clip."""
    x = np.asarray(x)
    if not np.issubdtype(x.dtype, np.inexact):
        return x
    y = (x - state.mean) / state.std
    if max_abs_value is not None:
        y = np.clip(y, -max_abs_value, max_abs_value)
    return y


def ppo_update_step_statistics(state: RunningStatistics, traj_obs_shards: Sequence[np.ndarray]):
    """The branch of `_update_step` (ff_ppo.py:145-162): the trajectory observations are normalised with the statistics
    from BEFORE the update (they are what the rollout's networks saw, :90-94), then the statistics absorb the RAW
    observations of every replica with std limits 5e-4 / 5e4.  Returns (normalised shards, new state)."""
    normalised = [normalize(o, state) for o in traj_obs_shards]
    return normalised, update(state, traj_obs_shards, std_min_value=5e-4, std_max_value=5e4)
