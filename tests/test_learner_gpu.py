"""End-to-end parity of one full Anakin update step (rollout bookkeeping -> batched critic -> GAE ->
4 epochs x minibatches of fused grads + clip/Adam with LR schedule) against the oracle's restatement
of ff_ppo.py:61-341, with the actions and permutations the learner actually used injected into the
oracle.  Stated tolerance after a whole update (fp32 kernels vs fp64 oracle):
parameters rtol 1e-4 / atol 2e-6, advantages/targets rtol 1e-5 / atol 1e-5."""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as O

pytestmark = pytest.mark.gpu


def _cfg(overrides):
    from stoix_b200.config import compose

    base = ["env=synthetic/box", "arch.num_evaluation=1", "arch.absolute_metric=False", "arch.num_eval_episodes=8",
            "logger.use_console=False"]
    cfg = compose("default_ff_ppo", base + overrides)
    cfg.num_devices = 1
    cfg.rank = 0
    return cfg


def _np(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


def adv_raw(traj):
    r_t, d_t, tr = O.ppo_gae_inputs(traj.reward, traj.done, traj.truncated, 0.99, 1.0)
    return O.gae(r_t, d_t, 0.95, v_tm1=traj.value, v_t=traj.bootstrap_value, truncation_t=tr, time_major=True)[0]


def _tree_to_oracle(tree):
    spec = tree.spec
    return O.MLPParams.from_flat(_np(tree.flat), list(spec.sizes))


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("E,T,nmb,layers,obs_dim,A", [(8, 16, 4, [32, 32], 12, 5), (64, 8, 2, [256, 256], 64, 8)])
def test_update_step_matches_oracle(E, T, nmb, layers, obs_dim, A, graph):
    from stoix_b200 import ops, random as srandom
    from stoix_b200.systems.ppo.anakin import ff_ppo
    from stoix_b200.utils import make_env
    from stoix_b200.utils.total_timestep_checker import check_total_timesteps

    n_updates = 3
    cfg = _cfg([f"arch.total_num_envs={E}", f"system.rollout_length={T}", f"system.num_minibatches={nmb}",
                f"arch.total_timesteps={E * T * n_updates}", f"arch.cuda_graph={graph}",
                f"network.actor_network.pre_torso.layer_sizes={layers}", f"network.critic_network.pre_torso.layer_sizes={layers}",
                f"env.kwargs.obs_dim={obs_dim}", f"env.kwargs.num_actions={A}", "env.kwargs.p_term=0.05", "env.kwargs.p_trunc=0.05"])
    cfg = check_total_timesteps(cfg, quiet=True)
    assert cfg.arch.num_updates == n_updates and cfg.arch.num_updates_per_eval == n_updates
    env, _ = make_env.make(cfg)
    keys = srandom.split(srandom.PRNGKey(cfg.arch.seed), 4)
    learn, actor_net, state = ff_ppo.learner_setup(env, (keys[0], keys[2], keys[3]), cfg)
    # give the networks non-trivial biases/heads so every gradient path is exercised
    with torch.no_grad():
        g = torch.Generator(device="cuda").manual_seed(1)
        arena = state.params.actor_params.arena
        arena.add_(torch.randn(arena.shape, device="cuda", generator=g) * 0.05)
    actor = _tree_to_oracle(state.params.actor_params)
    critic = _tree_to_oracle(state.params.critic_params)
    a_st = O.AdamState(np.zeros(actor.flat().size), np.zeros(actor.flat().size))
    c_st = O.AdamState(np.zeros(critic.flat().size), np.zeros(critic.flat().size))
    h = O.PPOHyper(epochs=4, num_minibatches=nmb, num_updates=n_updates)

    for upd in range(n_updates):
        cfg.arch.num_updates_per_eval = 1  # step update by update so the trajectory can be read back
        out = learn(state)
        state = out.learner_state
        torch.cuda.synchronize()
        sh = learn.built["shards"][0]
        traj = O.Trajectory(obs=_np(sh.obs[:T]), action=sh.action.cpu().numpy(), reward=_np(sh.reward),
                            done=sh.done.cpu().numpy().astype(bool), truncated=sh.truncated.cpu().numpy().astype(bool),
                            next_obs=_np(sh.next_obs))
        O.evaluate_rollout(actor, critic, traj)
        np.testing.assert_allclose(_np(sh.log_prob), traj.log_prob, rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(_np(sh.value), traj.value, rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(_np(sh.bootstrap_value), traj.bootstrap_value, rtol=1e-4, atol=2e-6)
        perms = np.stack([
            ops.make_permutation(T * E, state.key[1], ep + 4 * upd, device="cuda").cpu().numpy() for ep in range(4)
        ])
        actor, critic, metrics, adv, tgt = O.ppo_update(actor, critic, a_st, c_st, traj, perms, h)
        np.testing.assert_allclose(_np(sh.targets), tgt, rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(_np(sh.advantages), adv_raw(traj), rtol=1e-4, atol=2e-5)  # raw: standardised on load
        for name in ("actor_loss", "entropy", "value_loss"):
            np.testing.assert_allclose(_np(out.train_metrics[name][0]), metrics[name], rtol=2e-4, atol=2e-6)
        np.testing.assert_allclose(_np(state.params.actor_params.flat), actor.flat(), rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(_np(state.params.critic_params.flat), critic.flat(), rtol=1e-4, atol=2e-6)
        # the carried observation is the env's latest one (ff_ppo.py:131-134)
        assert torch.equal(state.timestep[0].observation, learn.built["shards"][0].obs[T])
    counts = state.params.actor_params.arena_counts.cpu().tolist()
    assert counts == [n_updates * 4 * nmb] * 4
    assert out.episode_metrics["episode_return"].shape == (1, 1, T, E)


def test_update_batch_size_two_averages_shards():
    """U=2: two env shards, gradients averaged before one optimiser step (vmap 'batch' + pmean)."""
    from stoix_b200 import random as srandom
    from stoix_b200.systems.ppo.anakin import ff_ppo
    from stoix_b200.utils import make_env
    from stoix_b200.utils.total_timestep_checker import check_total_timesteps

    E, T = 8, 8
    cfg = _cfg([f"arch.total_num_envs={2 * E}", "arch.update_batch_size=2", f"system.rollout_length={T}", "system.num_minibatches=2",
                f"arch.total_timesteps={2 * E * T * 2}", "network.actor_network.pre_torso.layer_sizes=[32,32]",
                "network.critic_network.pre_torso.layer_sizes=[32,32]", "env.kwargs.obs_dim=8", "env.kwargs.num_actions=3"])
    cfg = check_total_timesteps(cfg, quiet=True)
    assert cfg.arch.num_envs == E
    env, _ = make_env.make(cfg)
    keys = srandom.split(srandom.PRNGKey(1), 4)
    learn, _, state = ff_ppo.learner_setup(env, (keys[0], keys[2], keys[3]), cfg)
    p0 = state.params.actor_params.arena.clone()
    out = learn(state)
    torch.cuda.synchronize()
    assert out.episode_metrics["episode_return"].shape == (2, 2, T, E)
    assert torch.isfinite(out.learner_state.params.actor_params.arena).all()
    assert not torch.equal(p0, out.learner_state.params.actor_params.arena)
    s0, s1 = learn.built["shards"]
    assert not torch.equal(s0.obs, s1.obs) and not torch.equal(s0.perms, s1.perms)


def test_cartpole_learns_with_the_fp32_path():
    """BASELINE configs[0] plumbing case scaled up enough to see learning: CartPole-v1 return rises well
    above the ~22 of a random policy (reference README: PPO solves CartPole, docs/images/ppo_compare.png)."""
    from stoix_b200.systems.ppo.anakin import ff_ppo

    cfg = _cfg(["env=gymnax/cartpole", "arch.total_num_envs=256", "system.rollout_length=32", "system.num_minibatches=4",
                "arch.total_timesteps=600000", "arch.num_evaluation=3", "arch.num_eval_episodes=32", "arch.evaluation_greedy=True"])
    perf = ff_ppo.run_experiment(cfg)
    assert perf > 100.0, f"greedy CartPole return {perf} did not improve"
