"""MLPTorso options of the reference (stoix/networks/torso.py:12-33: activation table networks/utils.py:9-24,
use_layer_norm -> Dense(no bias) + nn.LayerNorm) on the fp32 CUDA path, through the C ABI, vs the oracle:
forward (stx_mlp_forward), both PPO losses' gradients incl. LayerNorm scale / bias (stx_ppo_minibatch_grads), and a
whole learner update step with a LayerNorm + tanh network.  Tolerances: the fp32-path ones of tests/test_kernels_gpu.py."""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as O

pytestmark = pytest.mark.gpu

ACTS = ["relu", "tanh", "silu", "elu", "gelu", "sigmoid", "softplus", "identity"]


def _t(x, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype, device="cuda:0")


def _net(rng, sizes, act, ln, head_scale):
    p = O.init_mlp(rng, sizes, head_scale)
    n = len(sizes) - 1
    b = [rng.standard_normal(sizes[i + 1]) * 0.2 + (1.0 if ln and i < n - 1 else 0.0) for i in range(n)]
    lnb = [rng.standard_normal(sizes[i + 1]) * 0.2 if i < n - 1 else None for i in range(n)] if ln else None
    W = [w.copy() for w in p.W]
    W[-1] = rng.standard_normal(W[-1].shape) * 0.3
    return O.MLPParams(W, b, act, lnb)


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("ln", [False, True])
@pytest.mark.parametrize("act", ACTS)
def test_forward_and_ppo_grads_match_oracle(act, ln):
    from stoix_b200 import ops

    rng = np.random.default_rng(ACTS.index(act) * 2 + int(ln))
    D, A, B, mb, mb_off = 12, 5, 1536, 1024, 256
    sizes_a, sizes_c = [D, 48, 40, A], [D, 48, 40, 1]
    actor, critic = _net(rng, sizes_a, act, ln, 0.3), _net(rng, sizes_c, act, ln, 1.0)
    sa = ops.MlpSpec(tuple(sizes_a), activation=act, use_layer_norm=ln)
    sc = ops.MlpSpec(tuple(sizes_c), activation=act, use_layer_norm=ln)
    _, coff, total = ops.arena_offsets(sa, sc)
    assert sa.param_count == actor.flat().size and sc.param_count == critic.flat().size
    flat = np.zeros(total, np.float32)
    flat[: sa.param_count] = actor.flat()
    flat[coff:coff + sc.param_count] = critic.flat()
    arena = _t(flat)
    obs = rng.standard_normal((B, D)).astype(np.float32)
    # forward
    out = ops.mlp_forward(sa, arena[:coff], _t(obs))
    ref, _ = O.mlp_forward(actor, obs.astype(np.float64))
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=2e-4, atol=2e-5)
    # gradients of both losses on a shuffled minibatch
    act_idx = rng.integers(0, A, B).astype(np.int32)
    lp_old = (O.categorical_log_prob(ref, act_idx) + rng.standard_normal(B) * 0.3).astype(np.float32)
    v_old, tgt = rng.standard_normal(B).astype(np.float32), rng.standard_normal(B).astype(np.float32)
    adv = (rng.standard_normal(B) * 2 + 0.3).astype(np.float32)
    mean = adv.astype(np.float64).mean()
    rstd = 1.0 / np.sqrt((adv.astype(np.float64) ** 2).mean() - mean * mean + 1e-5)
    perm = rng.permutation(B).astype(np.int32)
    idx = perm[mb_off:mb_off + mb]
    adv_n = (adv.astype(np.float64) - mean) * rstd
    x = obs.astype(np.float64)[idx]
    lg, a_acts = O.mlp_forward(actor, x)
    _, dlg, a_info = O.actor_loss_and_dlogits(lg, act_idx[idx], lp_old[idx].astype(np.float64), adv_n[idx], 0.2, 0.01)
    ga = O.mlp_backward(actor, a_acts, dlg).flat()
    v, c_acts = O.mlp_forward(critic, x)
    _, dv, c_info = O.critic_loss_and_dvalue(v[:, 0], v_old[idx].astype(np.float64), tgt[idx].astype(np.float64), 0.2, 0.5)
    gc = O.mlp_backward(critic, c_acts, dv[:, None]).flat()
    batch = ops.PpoBatch(_t(obs), _t(act_idx, torch.int32), _t(lp_old), _t(v_old), _t(adv), _t(tgt), adv_stats=_t(np.array([mean, rstd])),
                         perm=_t(perm, torch.int32))
    grads, metrics = torch.zeros(total, device="cuda:0"), torch.zeros(8, device="cuda:0")
    ws = ops.ppo_workspace(sa, sc, mb, ops.STX_PREC_F32, "cuda:0")
    ops.ppo_minibatch_grads(sa, sc, arena, batch, mb_off, mb, 0.2, 0.01, 0.5, True, grads, metrics, ws)
    torch.cuda.synchronize()
    g = grads.cpu().numpy().astype(np.float64)
    ea, ec = _rel(g[: sa.param_count], ga), _rel(g[coff:coff + sc.param_count], gc)
    print(f"{act} ln={ln}: actor grad rel {ea:.2e}, critic grad rel {ec:.2e}")
    assert ea < 2e-4 and ec < 2e-4
    np.testing.assert_allclose(g[: sa.param_count], ga, rtol=5e-3, atol=2e-5 * np.abs(ga).max())
    np.testing.assert_allclose(metrics.cpu().numpy()[:3], [a_info["actor_loss"], a_info["entropy"], c_info["value_loss"]], rtol=2e-4, atol=1e-5)
    # a second call accumulates (overwrite=False): deterministic kernels -> exactly twice
    ops.ppo_minibatch_grads(sa, sc, arena, batch, mb_off, mb, 0.2, 0.01, 0.5, True, grads, metrics, ws)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(grads.cpu().numpy(), (2 * g).astype(np.float32))
    # the tensor-core path refuses what it does not implement
    if act != "relu" or ln:
        with pytest.raises(ops.StxError):
            ops.mlp_forward(ops.MlpSpec((64, 256, 256, 8), activation=act, use_layer_norm=ln), torch.zeros(200000, device="cuda:0"),
                            torch.zeros(128, 64, device="cuda:0", dtype=torch.bfloat16), precision=ops.STX_PREC_BF16,
                            params_bf16=torch.zeros(200000, device="cuda:0", dtype=torch.bfloat16))


def test_learner_update_with_layernorm_tanh_torso_matches_oracle():
    """A whole Anakin update step with network.*.pre_torso.{use_layer_norm=True, activation=tanh}: parameter tree with the
    flax names (Dense_i/kernel, LayerNorm_i/{scale,bias}), oracle parity of the post-update parameters."""
    from stoix_b200 import ops, random as srandom
    from stoix_b200.config import compose
    from stoix_b200.systems.ppo.anakin import ff_ppo
    from stoix_b200.utils import make_env
    from stoix_b200.utils.total_timestep_checker import check_total_timesteps

    E, T, nmb = 32, 8, 2
    tor = ["use_layer_norm=True", "activation=tanh", "layer_sizes=[32,32]"]
    cfg = compose("default_ff_ppo", ["env=synthetic/box", f"arch.total_num_envs={E}", f"system.rollout_length={T}", f"system.num_minibatches={nmb}",
                                     f"arch.total_timesteps={E * T * 2}", "arch.num_evaluation=1", "logger.use_console=False",
                                     "env.kwargs.obs_dim=12", "env.kwargs.num_actions=4"]
                  + [f"network.actor_network.pre_torso.{t}" for t in tor] + [f"network.critic_network.pre_torso.{t}" for t in tor])
    cfg.num_devices, cfg.rank = 1, 0
    cfg = check_total_timesteps(cfg, quiet=True)
    env, _ = make_env.make(cfg)
    keys = srandom.split(srandom.PRNGKey(2), 4)
    learn, _, state = ff_ppo.learner_setup(env, (keys[0], keys[2], keys[3]), cfg)
    a_tree = state.params.actor_params
    assert set(a_tree["params"]["torso"]) == {"Dense_0", "Dense_1", "LayerNorm_0", "LayerNorm_1"}
    assert set(a_tree["params"]["torso"]["Dense_0"]) == {"kernel"} and set(a_tree["params"]["torso"]["LayerNorm_0"]) == {"scale", "bias"}
    assert float(a_tree["params"]["torso"]["LayerNorm_1"]["scale"].min()) == 1.0
    with torch.no_grad():
        g = torch.Generator(device="cuda").manual_seed(1)
        a_tree.arena.add_(torch.randn(a_tree.arena.shape, device="cuda", generator=g) * 0.05)
    f64 = lambda t: t.detach().float().cpu().numpy().astype(np.float64)
    tree = lambda tr: O.MLPParams.from_flat(f64(tr.flat), list(tr.spec.sizes), tr.spec.activation, tr.spec.use_layer_norm)
    actor, critic = tree(state.params.actor_params), tree(state.params.critic_params)
    cfg.arch.num_updates_per_eval = 1
    out = learn(state)
    torch.cuda.synchronize()
    sh = learn.built["shards"][0]
    traj = O.Trajectory(obs=f64(sh.obs[:T]), action=sh.action.cpu().numpy(), reward=f64(sh.reward), done=sh.done.cpu().numpy().astype(bool),
                        truncated=sh.truncated.cpu().numpy().astype(bool), next_obs=f64(sh.next_obs))
    O.evaluate_rollout(actor, critic, traj)
    np.testing.assert_allclose(f64(sh.log_prob), traj.log_prob, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(f64(sh.value), traj.value, rtol=1e-4, atol=2e-6)
    perms = np.stack([ops.make_permutation(T * E, state.key[1], ep, device="cuda").cpu().numpy() for ep in range(4)])
    n_a, n_c = actor.flat().size, critic.flat().size
    h = O.PPOHyper(num_minibatches=nmb, num_updates=int(cfg.arch.num_updates))
    a2, c2, _, _, tgt = O.ppo_update(actor, critic, O.AdamState(np.zeros(n_a), np.zeros(n_a)), O.AdamState(np.zeros(n_c), np.zeros(n_c)), traj, perms, h)
    np.testing.assert_allclose(f64(sh.targets), tgt, rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(f64(out.learner_state.params.actor_params.flat), a2.flat(), rtol=2e-4, atol=4e-6)
    np.testing.assert_allclose(f64(out.learner_state.params.critic_params.flat), c2.flat(), rtol=2e-4, atol=4e-6)
