"""Host-side logic that mirrors the reference's Python faces, checked on CPU against the oracle's restatement of the
same reference lines: shape derivation (total_timestep_checker.py), the learning-rate schedule (utils/training.py),
the config layer (Hydra-style composition, `_target_` resolution) and the leading-dim helpers (utils/jax_utils.py)."""
import itertools

import numpy as np
import pytest
import torch

from oracle import ppo_oracle as O
from stoix_b200.config import _locate, compose, instantiate
from stoix_b200.utils import jax_utils
from stoix_b200.utils.total_timestep_checker import check_total_timesteps
from stoix_b200.utils.training import make_learning_rate, make_learning_rate_schedule


@pytest.mark.parametrize("devices,ubs", [(1, 1), (2, 1), (8, 1), (4, 2), (1, 4)])
def test_shape_derivation_matches_the_successive_floor_divisions(devices, ubs):
    """total_timestep_checker.py:57-61, 88-96, 107: num_envs, num_updates (floor divisions IN THAT ORDER), updates per eval."""
    for total_envs, T, steps, n_eval in itertools.product((64, 4096 * 8, 1024), (8, 128), (1e6, 41943040, 3e7), (1, 5, 20)):
        if total_envs % (devices * ubs):
            continue
        cfg = compose("default_ff_ppo", ["env=synthetic/box", f"arch.total_num_envs={total_envs}", f"arch.update_batch_size={ubs}",
                                         f"system.rollout_length={T}", f"arch.total_timesteps={steps}", f"arch.num_evaluation={n_eval}"])
        cfg.num_devices = devices
        want = O.derive_shapes(total_envs, devices, ubs, steps, T, n_eval)
        if want[1] < n_eval:  # fewer updates than evaluations: nothing to derive (the reference asserts on this configuration)
            continue
        cfg = check_total_timesteps(cfg, quiet=True)
        assert (cfg.arch.num_envs, cfg.arch.num_updates, cfg.arch.num_updates_per_eval) == want


def test_shape_derivation_from_num_updates_and_divisibility_error():
    cfg = compose("default_ff_ppo", ["env=synthetic/box", "arch.total_num_envs=512", "arch.total_timesteps=null", "arch.num_updates=7",
                                     "system.rollout_length=16", "arch.num_evaluation=1"])
    cfg.num_devices = 2
    cfg = check_total_timesteps(cfg, quiet=True)
    assert cfg.arch.num_envs == 256 and cfg.arch.total_timesteps == 2 * 7 * 16 * 256  # :72-85
    cfg.arch.total_num_envs = 511
    with pytest.raises(AssertionError, match="must be divisible"):
        check_total_timesteps(cfg, quiet=True)


def test_learning_rate_schedule_floor_division_and_constants():
    """training.py:24-26: lr(count) = init * (1 - (count // (epochs*minibatches)) / num_updates); constant when decay is off."""
    sched = make_learning_rate_schedule(3e-4, num_updates=10, num_epochs=4, num_minibatches=16)
    for count in (0, 1, 63, 64, 65, 127, 128, 639, 640):
        assert sched(count) == pytest.approx(O.linear_schedule(3e-4, count, 10, 4, 16), rel=0, abs=0)
    assert sched(63) == sched(0) and sched(64) < sched(63)  # steps down once per update, not per optimiser step
    assert (sched.init_lr, sched.num_updates, sched.steps_per_update) == (3e-4, 10, 64)  # what the device-side schedule reads
    cfg = compose("default_ff_ppo", ["env=synthetic/box", "system.decay_learning_rates=False"])
    assert make_learning_rate(1e-3, cfg, 4, 16) == 1e-3
    cfg = compose("default_ff_ppo", ["env=synthetic/box", "system.decay_learning_rates=True", "arch.num_updates=5"])
    assert make_learning_rate(1e-3, cfg, 4, 16)(64 * 5 - 1) == pytest.approx(1e-3 * (1 - 4 / 5))


def test_config_composition_overrides_and_target_resolution():
    cfg = compose("default_ff_ppo", ["env=gymnax/cartpole", "system.clip_eps=0.3", "network.actor_network.pre_torso.layer_sizes=[32,32]",
                                     "arch.precision=bf16", "arch.seed=7"])
    assert cfg.env.scenario.name == "CartPole-v1" and cfg.system.clip_eps == 0.3 and cfg.arch.seed == 7
    assert cfg.network.actor_network.pre_torso.layer_sizes == [32, 32]
    assert cfg.network.critic_network.pre_torso.layer_sizes == [256, 256]  # the other network keeps the default
    # the reference's `_target_` strings resolve to this package's classes (same constructor arguments)
    assert cfg.network.actor_network.pre_torso._target_ == "stoix.networks.torso.MLPTorso"
    torso = instantiate(cfg.network.actor_network.pre_torso)
    assert type(torso) is _locate("stoix.networks.torso.MLPTorso") and type(torso).__module__ == "stoix_b200.networks.torso"
    assert tuple(torso.layer_sizes) == (32, 32)
    with pytest.raises(FileNotFoundError):
        compose("default_ff_ppo", ["env=does/not_exist"])
    # system defaults of configs/system/ppo/ff_ppo.yaml:6-22
    d = compose("default_ff_ppo", ["env=synthetic/box"]).system
    assert (d.rollout_length, d.epochs, d.num_minibatches, d.gamma, d.gae_lambda, d.clip_eps, d.ent_coef, d.vf_coef, d.max_grad_norm) == \
        (128, 4, 16, 0.99, 0.95, 0.2, 0.01, 0.5, 0.5)


def test_merge_leading_dims_is_time_major_flat_index():
    """jax_utils.py:29-43: the flat batch index of (T, E) is t*E + e -- the index space of the shuffle permutation."""
    T, E, D = 3, 5, 2
    x = torch.arange(T * E * D).reshape(T, E, D)
    flat = jax_utils.merge_leading_dims(x, 2)
    assert flat.shape == (T * E, D)
    assert torch.equal(flat[2 * E + 4], x[2, 4])
    assert jax_utils.merge_leading_dims(torch.zeros(4), 2).shape == (4,)  # fewer dims than requested: unchanged
    tree = {"a": torch.ones(2, 3, 4), "b": (torch.zeros(2, 3), 5)}
    out = jax_utils.unreplicate_n_dims(tree, 2)
    assert out["a"].shape == (4,) and out["b"][0].ndim == 0 and out["b"][1] == 5
    assert jax_utils.unreplicate_batch_dim({"m": torch.ones(6, 2, 3)})["m"].shape == (6, 3)


def test_oracle_update_is_invariant_to_the_shard_axis_layout():
    """The learner flattens (T, E) time-major; the oracle must agree with an explicit loop over the flat index."""
    rng = np.random.default_rng(3)
    T, E = 6, 4
    r, d = rng.standard_normal((T, E)), (rng.random((T, E)) < 0.2)
    v, bv = rng.standard_normal((T, E)), rng.standard_normal((T, E))
    r_t, d_t, trunc = O.ppo_gae_inputs(r, d, np.zeros_like(d), 0.99, 1.0)
    adv, tgt = O.gae(r_t, d_t, 0.95, v_tm1=v, v_t=bv, truncation_t=trunc, time_major=True, standardize_advantages=False)
    for e in range(E):  # column by column: the scan runs over time only, environments are independent
        a1, t1 = O.gae(r_t[:, e : e + 1], d_t[:, e : e + 1], 0.95, v_tm1=v[:, e : e + 1], v_t=bv[:, e : e + 1],
                       truncation_t=trunc[:, e : e + 1], time_major=True, standardize_advantages=False)
        np.testing.assert_allclose(adv[:, e], a1[:, 0], rtol=1e-12)
        np.testing.assert_allclose(tgt[:, e], t1[:, 0], rtol=1e-12)
