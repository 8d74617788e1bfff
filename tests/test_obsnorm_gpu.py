"""Observation normalisation on the GPU vs oracle/running_statistics.py (the restatement of
stoix/utils/running_statistics.py:123-135, 204-345, 348-363 and of the ff_ppo branch :90-94, 113-115, 145-162).

Tolerances: the kernels accumulate in double and keep the state in fp32, the oracle is fp64: mean / std rtol 1e-5,
summed_variance rtol 1e-5 (atol scaled by the data), normalised values rtol 1e-5 / atol 1e-6 (fp32 output) or one bf16
ulp (bf16 output)."""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as O
from oracle import running_statistics as RO

pytestmark = pytest.mark.gpu


def _t(x, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda:0")


@pytest.mark.parametrize("D,shapes", [(64, [(16, 256), (128, 64), (3, 5)]), (4, [(7, 33), (1, 1)]), (17, [(50, 20)]), (256, [(9, 40)])])
def test_update_and_normalize_match_oracle(D, shapes):
    """Several consecutive batched-Welford updates (different batch shapes, std limits of the ff_ppo call site), then
    normalize: public functions of stoix_b200.utils.running_statistics vs the oracle."""
    from stoix_b200.utils import running_statistics as rs

    rng = np.random.default_rng(D)
    scale, shift = rng.uniform(0.1, 5.0, D), rng.uniform(-3, 3, D)
    state = rs.initialize_statistics(torch.zeros(D, device="cuda:0"))
    ref = RO.initialize((D,))
    x0 = rng.standard_normal((4, D))
    np.testing.assert_array_equal(rs.normalize(_t(x0), state).cpu().numpy(), x0.astype(np.float32))  # identity before any update
    for shp in shapes:
        x = (rng.standard_normal(shp + (D,)) * scale + shift).astype(np.float32)
        new = rs.update_statistics(state, _t(x), std_min_value=5e-4, std_max_value=5e4)
        assert new is not state and int(state.count.item()) == ref.count  # functional: the old state is untouched
        state = new
        ref = RO.update(ref, [x.astype(np.float64)], std_min_value=5e-4, std_max_value=5e4)
        assert int(state.count.item()) == int(ref.count)
        np.testing.assert_allclose(state.mean.cpu().numpy(), ref.mean, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(state.summed_variance.cpu().numpy(), ref.summed_variance, rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(state.std.cpu().numpy(), ref.std, rtol=1e-5, atol=1e-7)
    y = (rng.standard_normal((33, D)) * scale + shift).astype(np.float32)
    want = RO.normalize(y.astype(np.float64), ref)
    np.testing.assert_allclose(rs.normalize(_t(y), state).cpu().numpy(), want, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(rs.normalize(_t(y), state, max_abs_value=0.5).cpu().numpy(), np.clip(want, -0.5, 0.5), rtol=2e-5, atol=2e-6)
    from stoix_b200 import ops

    yb = ops.obs_normalize(_t(y), state.mean, state.std, out_dtype=torch.bfloat16).float().cpu().numpy()
    np.testing.assert_allclose(yb, want, rtol=2 ** -8, atol=1e-6)
    np.testing.assert_allclose(rs.denormalize(rs.normalize(_t(y), state), state).cpu().numpy(), y, rtol=1e-4, atol=1e-4)
    assert rs.normalize(torch.arange(4, device="cuda:0"), state).dtype == torch.int64  # non-float leaves pass through


def test_weights_shards_and_degenerate_features():
    from stoix_b200.utils import running_statistics as rs

    rng = np.random.default_rng(5)
    D = 8
    x = rng.standard_normal((6, 10, D)).astype(np.float32)
    x[..., 3] = 2.5  # constant feature: variance 0 -> std clipped to std_min_value
    w = rng.integers(0, 3, (6, 10)).astype(np.float32)
    st = rs.initialize_statistics_from_data(torch.zeros(D, device="cuda:0"), _t(x), weights=_t(w))
    # weight k == the row repeated k times (running_statistics.py docstring of update_statistics)
    rep = np.repeat(x.reshape(-1, D), w.reshape(-1).astype(int), axis=0)
    ref = RO.update(RO.initialize((D,)), [rep.astype(np.float64)], std_min_value=5e-4, std_max_value=5e4)
    assert int(st.count.item()) == int(ref.count)
    np.testing.assert_allclose(st.mean.cpu().numpy(), ref.mean, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st.std.cpu().numpy(), ref.std, rtol=1e-5, atol=1e-7)
    assert abs(float(st.std[3]) - 5e-4) < 1e-9
    # the "batch" axis: a list of shards is one update over their union
    a, b = x[:2], x[2:]
    st2 = rs.update_statistics(rs.initialize_statistics(torch.zeros(D, device="cuda:0")), [_t(a), _t(b)], std_min_value=5e-4, std_max_value=5e4)
    ref2 = RO.update(RO.initialize((D,)), [a.astype(np.float64), b.astype(np.float64)], std_min_value=5e-4, std_max_value=5e4)
    np.testing.assert_allclose(st2.mean.cpu().numpy(), ref2.mean, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st2.summed_variance.cpu().numpy(), ref2.summed_variance, rtol=1e-5, atol=1e-4)
    with pytest.raises(ValueError):
        rs.update_statistics(st2, _t(x[..., :4]))


def test_full_size_statistics_properties():
    """BASELINE config 2 batch (128 x 4096 x 64 fp32 = 134 MB): mean / std against torch's own reductions, and the
    normalised batch has zero mean / unit variance per feature."""
    from stoix_b200.utils import running_statistics as rs

    g = torch.Generator(device="cuda:0").manual_seed(0)
    D = 64
    x = torch.randn(128, 4096, D, device="cuda:0", generator=g) * torch.linspace(0.5, 4.0, D, device="cuda:0") + torch.linspace(-2, 2, D, device="cuda:0")
    st = rs.initialize_statistics_from_data(torch.zeros(D, device="cuda:0"), x)
    assert int(st.count.item()) == 128 * 4096
    xd = x.double().view(-1, D)
    np.testing.assert_allclose(st.mean.cpu().numpy(), xd.mean(0).cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(st.std.cpu().numpy(), xd.std(0, unbiased=False).cpu().numpy(), rtol=1e-5)
    y = rs.normalize(x, st).double().view(-1, D)
    assert float(y.mean(0).abs().max()) < 1e-4 and float((y.std(0, unbiased=False) - 1).abs().max()) < 1e-4


@pytest.mark.parametrize("precision", ["f32", "bf16"])
def test_learner_with_observation_normalisation_matches_oracle(precision):
    """Two whole update steps with system.normalize_observations=True: warm-up statistics, normalised rollout inputs
    (pre-update statistics), statistics absorbing the raw trajectory, normalised minibatch observations -- against
    oracle.ppo_update on the oracle-normalised trajectory and oracle/running_statistics.ppo_update_step_statistics."""
    from stoix_b200 import ops, random as srandom
    from stoix_b200.config import compose
    from stoix_b200.systems.ppo.anakin import ff_ppo
    from stoix_b200.utils import make_env
    from stoix_b200.utils.total_timestep_checker import check_total_timesteps

    bf16 = precision == "bf16"
    E, T, nmb, n_upd = 128, 8, 2, 2
    cfg = compose("default_ff_ppo", ["env=synthetic/box", f"arch.total_num_envs={E}", f"system.rollout_length={T}", f"system.num_minibatches={nmb}",
                                     f"arch.total_timesteps={E * T * n_upd}", "arch.num_evaluation=1", f"arch.precision={precision}",
                                     "system.normalize_observations=True", "system.obs_norm_warmup_steps=4", "logger.use_console=False",
                                     "env.kwargs.p_term=0.05", "env.kwargs.p_trunc=0.05"])
    cfg.num_devices, cfg.rank = 1, 0
    cfg = check_total_timesteps(cfg, quiet=True)
    env, _ = make_env.make(cfg)
    keys = srandom.split(srandom.PRNGKey(7), 4)
    learn, _, state = ff_ppo.learner_setup(env, (keys[0], keys[2], keys[3]), cfg)
    rs0 = state.running_statistics
    assert int(rs0.count.item()) == (4 + 1) * E
    params, opt_states, key, env_state, timestep = state  # unpacking yields the ORIGINAL fields (running_statistics.py:444-530)
    f64 = lambda t: t.detach().float().cpu().numpy().astype(np.float64)
    ref = RO.RunningStatistics(float(rs0.count.item()), f64(rs0.mean), f64(rs0.summed_variance), f64(rs0.std))
    tree = lambda tr: O.MLPParams.from_flat(f64(tr.flat), list(tr.spec.sizes))
    actor, critic = tree(state.params.actor_params), tree(state.params.critic_params)
    n_a, n_c = actor.flat().size, critic.flat().size
    a_st, c_st = O.AdamState(np.zeros(n_a), np.zeros(n_a)), O.AdamState(np.zeros(n_c), np.zeros(n_c))
    h = O.PPOHyper(num_minibatches=nmb, num_updates=n_upd)
    cfg.arch.num_updates_per_eval = 1
    rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    for upd in range(n_upd):
        out = learn(state)
        state = out.learner_state
        torch.cuda.synchronize()
        sh = learn.built["shards"][0]
        raw, raw_next = f64(sh.obs_raw[:T]), f64(sh.next_obs_raw)
        (norm, norm_next), ref_new = RO.ppo_update_step_statistics(ref, [raw, raw_next])[0], RO.update(ref, [raw], std_min_value=5e-4, std_max_value=5e4)
        tol = dict(rtol=2 ** -8, atol=1e-6) if bf16 else dict(rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(f64(sh.obs[:T]), norm, **tol)        # what the networks saw: pre-update statistics
        np.testing.assert_allclose(f64(sh.next_obs), norm_next, **tol)
        rs_now = state.running_statistics
        assert int(rs_now.count.item()) == int(ref_new.count)
        np.testing.assert_allclose(f64(rs_now.mean), ref_new.mean, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(f64(rs_now.std), ref_new.std, rtol=1e-5, atol=1e-7)
        ref = ref_new
        # the update itself on the (kernel-)normalised observations
        traj = O.Trajectory(obs=f64(sh.obs[:T]), action=sh.action.cpu().numpy(), reward=f64(sh.reward), done=sh.done.cpu().numpy().astype(bool),
                            truncated=sh.truncated.cpu().numpy().astype(bool), next_obs=f64(sh.next_obs))
        O.evaluate_rollout(actor, critic, traj, bf16=bf16)
        if bf16:
            np.testing.assert_allclose(f64(sh.value), traj.value, rtol=2e-3, atol=2e-2)
            traj.value, traj.bootstrap_value, traj.log_prob = f64(sh.value), f64(sh.bootstrap_value), f64(sh.log_prob)
        else:
            np.testing.assert_allclose(f64(sh.value), traj.value, rtol=1e-4, atol=2e-6)
            np.testing.assert_allclose(f64(sh.log_prob), traj.log_prob, rtol=1e-4, atol=2e-6)
        perms = np.stack([ops.make_permutation(T * E, state.key[1], ep + 4 * upd, device="cuda").cpu().numpy() for ep in range(4)])
        actor, critic, metrics, _, tgt = O.ppo_update(actor, critic, a_st, c_st, traj, perms, h, bf16=bf16)
        np.testing.assert_allclose(f64(sh.targets), tgt, rtol=1e-4, atol=2e-5)
        a_tree, c_tree = state.params.actor_params, state.params.critic_params
        if bf16:
            _, coff, _ = ops.arena_offsets(a_tree.spec, c_tree.spec)
            mu, nu = f64(a_tree.arena_mu), f64(a_tree.arena_nu)
            assert max(rel(mu[:n_a], a_st.mu), rel(mu[coff:coff + n_c], c_st.mu)) < 0.3
            assert max(rel(nu[:n_a], a_st.nu), rel(nu[coff:coff + n_c], c_st.nu)) < 0.1
            actor, critic = tree(a_tree), tree(c_tree)
            a_st.mu, a_st.nu, c_st.mu, c_st.nu = mu[:n_a].copy(), nu[:n_a].copy(), mu[coff:coff + n_c].copy(), nu[coff:coff + n_c].copy()
        else:
            np.testing.assert_allclose(f64(a_tree.flat), actor.flat(), rtol=1e-4, atol=2e-6)
            np.testing.assert_allclose(f64(c_tree.flat), critic.flat(), rtol=1e-4, atol=2e-6)
        # the carried observation stays RAW (the next rollout normalises it with the NEW statistics)
        assert torch.equal(state.timestep[0].observation, sh.obs_raw[T])
