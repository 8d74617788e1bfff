"""ff_sac on the GPU, fp32, against oracle/sac_oracle.py (itself checked against torch autograd in
tests/test_oracle_sac.py):

* tanh-Normal head: sample / log_prob (clipped tails included) and the reparameterised backward, with injected noise;
* replay ring: add (wrap-around, more rows than slots) and the fused sample vs uniform_indices + gather_rows;
* one whole `_update_epoch` of the learner (three losses, gradients, clip + Adam over three segments, Polyak) with an injected
  batch and injected noise vs sac_update, twice (Adam counters advance);
* a short experiment through run_experiment (CUDA graph replays included): finite, parameters move, alpha adapts."""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as O
from oracle import sac_oracle as S

pytestmark = pytest.mark.gpu

f64 = lambda t: t.detach().float().cpu().numpy().astype(np.float64)
dev = lambda x, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(x), device="cuda").to(dt)


@pytest.mark.parametrize("scale", [1.0, 4.0])   # 4.0 pushes samples into the clipped tails
def test_tanh_normal_head_matches_oracle(scale):
    from stoix_b200 import ops

    rng = np.random.default_rng(0)
    M, A, lo, hi = 777, 6, -1.0, 1.0
    head = (rng.standard_normal((M, 2 * A)) * scale).astype(np.float32)
    eps = rng.standard_normal((M, A)).astype(np.float32)
    wide = torch.zeros(M, 17 + A, device="cuda")
    action, logp, eps_out = ops.tanh_normal_sample(dev(head), lo, hi, eps=dev(eps), action_out=wide[:, 17:])
    t = S.head_sample(head.astype(np.float64), eps.astype(np.float64), lo, hi)
    np.testing.assert_allclose(f64(wide[:, 17:]), t.action, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(f64(logp), t.log_prob, rtol=2e-4, atol=2e-4)
    assert float(wide[:, :17].abs().max()) == 0.0
    if scale > 1:   # tails are exercised
        assert (np.abs(t.action) > 1 - 1e-3).any()
    g_action = rng.standard_normal((M, A)).astype(np.float32)
    log_alpha = torch.tensor([0.3], device="cuda")
    d = ops.tanh_normal_backward(dev(head), dev(eps), lo, hi, log_alpha, 1.0 / M, dev(g_action))
    want = S.head_log_prob_grad(t, np.full(M, np.exp(0.3) / M), g_action.astype(np.float64), lo, hi)
    np.testing.assert_allclose(f64(d), want, rtol=2e-3, atol=2e-6)
    # Philox noise: reproducible per (seed, offset + counter), standard normal, returned for the backward
    ctr = torch.tensor([5], dtype=torch.int64, device="cuda")
    a1, l1, e1 = ops.tanh_normal_sample(dev(head), lo, hi, seed=9, offset=2, dev_counter=ctr)
    a2, l2, e2 = ops.tanh_normal_sample(dev(head), lo, hi, seed=9, offset=7)
    assert torch.equal(a1, a2) and torch.equal(e1, e2) and torch.equal(l1, l2)
    a3, _, e3 = ops.tanh_normal_sample(dev(head), lo, hi, seed=9, offset=8)
    assert not torch.equal(e1, e3)
    assert abs(float(e1.mean())) < 0.05 and abs(float(e1.std()) - 1) < 0.05
    t1 = S.head_sample(head.astype(np.float64), f64(e1), lo, hi)
    np.testing.assert_allclose(f64(a1), t1.action, rtol=1e-5, atol=2e-6)


def test_replay_ring_add_and_sample():
    from stoix_b200 import ops
    from stoix_b200.systems.sac.sac_types import Transition
    from stoix_b200.utils.replay import TransitionBuffer

    D, A, cap, B = 17, 6, 1000, 256
    buf = TransitionBuffer(cap, B, B, D, A, "cuda", seed=77)
    rng = np.random.default_rng(1)
    ref = {k: np.zeros((cap,) + s, np.float32) for k, s in (("obs", (D,)), ("action", (A,)), ("reward", ()), ("done", ()), ("next_obs", (D,)))}
    wp, filled = 0, 0
    for T, E in ((3, 100), (1, 64), (5, 128), (2, 640)):   # wraps; the last add has more rows (1280) than slots
        tr = {"obs": rng.standard_normal((T, E, D)), "action": rng.standard_normal((T, E, A)), "reward": rng.standard_normal((T, E)),
              "done": (rng.random((T, E)) < 0.3), "next_obs": rng.standard_normal((T, E, D))}
        tr = {k: v.astype(np.float32) for k, v in tr.items()}
        buf.add(Transition(dev(tr["obs"]), dev(tr["action"]), dev(tr["reward"]), dev(tr["done"], torch.uint8), dev(tr["next_obs"]), {}))
        n = T * E
        for k, v in tr.items():
            rows = v.reshape((n,) + v.shape[2:])
            for r in range(max(0, n - cap), n):
                ref[k][(wp + r) % cap] = rows[r]
        wp, filled = (wp + n) % cap, min(filled + n, cap)
        assert buf.ring.state.cpu().tolist() == [wp, filled]
        assert buf.can_sample() == (filled >= B)
        for k in ref:
            np.testing.assert_array_equal(getattr(buf.ring, k).float().cpu().numpy(), ref[k])
    ld = D + A
    xo, xn, xx = (torch.full((B, ld), 9.0, device="cuda") for _ in range(3))
    rew, done, idx = torch.zeros(B, device="cuda"), torch.zeros(B, dtype=torch.uint8, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda")
    for call in range(3):
        buf.sample_into(xo, rew, done, xq_new=xn, xq_next=xx, idx_out=idx)
        want = ops.uniform_indices(B, buf.ring.state[1:2], buf.seed, offset=call)
        assert torch.equal(idx, want) and int(idx.min()) >= 0 and int(idx.max()) < filled
        j = idx.cpu().numpy()
        np.testing.assert_array_equal(xo.cpu().numpy(), np.concatenate([ref["obs"][j], ref["action"][j]], 1))
        np.testing.assert_array_equal(xn[:, :D].cpu().numpy(), ref["obs"][j])
        np.testing.assert_array_equal(xx[:, :D].cpu().numpy(), ref["next_obs"][j])
        assert float((xn[:, D:] - 9).abs().max()) == 0 and float((xx[:, D:] - 9).abs().max()) == 0
        np.testing.assert_array_equal(rew.cpu().numpy(), ref["reward"][j])
        np.testing.assert_array_equal(done.cpu().numpy(), ref["done"][j].astype(np.uint8))
        g = torch.zeros(B, D, device="cuda")
        ops.gather_rows(buf.ring.obs, idx, g)
        assert torch.equal(g, xo[:, :D])
    assert int(buf.counter) == 3
    # roughly uniform over the valid part
    big = ops.uniform_indices(1 << 16, buf.ring.state[1:2], 3).cpu().numpy()
    hist = np.bincount(big * 8 // filled, minlength=8) / big.size
    assert np.abs(hist - 0.125).max() < 0.01


def _cfg(extra=()):
    from stoix_b200.config import compose

    return compose("default_ff_sac", ["arch.total_num_envs=64", "system.total_batch_size=128", "system.total_buffer_size=4096",
                                      "system.warmup_steps=4", "arch.total_timesteps=6400", "arch.num_evaluation=2", "arch.num_eval_episodes=16",
                                      "logger.use_console=False", "arch.max_eval_steps=100", "env.kwargs.p_term=0.05",
                                      "network.actor_network.pre_torso.layer_sizes=[64,64]",
                                      "network.q_network.pre_torso.layer_sizes=[64,64,64]"] + list(extra), config_dir="default/anakin")


def _trees(state):
    p = state.params
    mk = lambda tr: O.MLPParams.from_flat(f64(tr.flat), list(tr.spec.sizes), tr.spec.activation, tr.spec.use_layer_norm)
    return mk(p.actor_params), tuple(mk(t) for t in p.q_params.online), tuple(mk(t) for t in p.q_params.target)


@pytest.mark.parametrize("autotune", [True, False])
def test_update_epoch_matches_oracle(autotune):
    from stoix_b200 import random as srandom
    from stoix_b200.systems.sac import ff_sac
    from stoix_b200.utils import make_env as environments
    from stoix_b200.utils.total_timestep_checker import check_total_timesteps

    torch.cuda.set_device(0)
    cfg = _cfg([f"system.autotune={autotune}", "arch.cuda_graph=False"])
    cfg.num_devices, cfg.rank = 1, 0
    cfg = check_total_timesteps(cfg, quiet=True)
    env, _ = environments.make(cfg)
    keys = srandom.split(srandom.PRNGKey(3), 3)
    learn, actor_network, state = ff_sac.learner_setup(env, tuple(keys), cfg)
    with torch.no_grad():   # non-trivial biases, LayerNorm parameters and targets that differ from the online networks
        g = torch.Generator(device="cuda").manual_seed(1)
        a = state.params.actor_params
        a.arena.add_(torch.randn(a.arena.shape, device="cuda", generator=g) * 0.05)
        a.target_arena.add_(torch.randn(a.target_arena.shape, device="cuda", generator=g) * 0.05)
        if autotune:
            state.params.log_alpha.fill_(-0.7)
    learn.ensure_built(state)
    b = learn.built
    sh = b["shards"][0]
    lay, D, A, B = b["lay"], b["D"], b["A"], int(cfg.system.batch_size)
    assert (D, A, B) == (17, 6, 128)
    actor, q_on, q_tg = _trees(state)
    log_alpha = float(state.params.log_alpha)
    n_a, n_q = actor.flat().size, q_on[0].flat().size
    opt = {"actor": O.AdamState(np.zeros(n_a), np.zeros(n_a)), "q": O.AdamState(np.zeros(2 * n_q), np.zeros(2 * n_q)),
           "alpha": O.AdamState(np.zeros(1), np.zeros(1))}
    h = S.SACHyper(gamma=float(cfg.system.gamma), tau=float(cfg.system.tau), max_grad_norm=float(cfg.system.max_grad_norm),
                   actor_lr=float(cfg.system.actor_lr), q_lr=float(cfg.system.q_lr), alpha_lr=float(cfg.system.alpha_lr), autotune=autotune,
                   target_entropy=float(cfg.system.target_entropy), minimum=-1.0, maximum=1.0)
    assert h.target_entropy == -6.0
    rng = np.random.default_rng(5)
    for step in range(2):
        batch = {"obs": rng.standard_normal((B, D)), "action": np.tanh(rng.standard_normal((B, A))), "reward": rng.standard_normal(B),
                 "done": rng.random(B) < 0.2, "next_obs": rng.standard_normal((B, D))}
        batch = {k: (v.astype(np.float32).astype(np.float64) if k != "done" else v) for k, v in batch.items()}
        noise = {k: rng.standard_normal((B, A)).astype(np.float32) for k in ("actor", "q", "alpha")}
        sh.xq_old.copy_(dev(np.concatenate([batch["obs"], batch["action"]], 1)))
        sh.xq_new[:, :D].copy_(dev(batch["obs"]))
        sh.xq_next[:, :D].copy_(dev(batch["next_obs"]))
        sh.b_reward.copy_(dev(batch["reward"]))
        sh.b_done.copy_(dev(batch["done"], torch.uint8))
        learn.update_epoch(state, 0, noise={k: dev(v) for k, v in noise.items()}, sample=False)
        torch.cuda.synchronize()
        ag, qg, alg, info, aux = S.sac_losses_and_grads(actor, q_on, q_tg, log_alpha, batch, {k: v.astype(np.float64) for k, v in noise.items()}, h)
        grads = f64(b["grads"])
        rel = lambda x, y: float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-30))
        assert rel(grads[:n_a], ag.flat()) < 2e-4
        assert rel(grads[lay["q1"]: lay["q1"] + n_q], qg[0].flat()) < 2e-4
        assert rel(grads[lay["q2"]: lay["q2"] + n_q], qg[1].flat()) < 2e-4
        if autotune:
            np.testing.assert_allclose(grads[lay["alpha"]], alg, rtol=2e-4)
        m = f64(b["metrics"][0])
        for j, name in enumerate(ff_sac._METRIC_NAMES):
            want = info[name] if (autotune or name != "alpha_loss") else 0.0
            np.testing.assert_allclose(m[j], want, rtol=5e-4, atol=5e-5, err_msg=name)
        np.testing.assert_allclose(f64(sh.xq_new[:, D:]), aux["actor_sample"].action, rtol=1e-4, atol=1e-5)
        actor, q_on, q_tg, log_alpha, _ = S.sac_update(actor, q_on, q_tg, log_alpha, opt, batch, {k: v.astype(np.float64) for k, v in noise.items()}, h)
        a2, q2, t2 = _trees(state)
        np.testing.assert_allclose(a2.flat(), actor.flat(), rtol=2e-4, atol=3e-6)
        for k in range(2):
            np.testing.assert_allclose(q2[k].flat(), q_on[k].flat(), rtol=2e-4, atol=3e-6)
            np.testing.assert_allclose(t2[k].flat(), q_tg[k].flat(), rtol=2e-4, atol=3e-6)
        np.testing.assert_allclose(float(state.params.log_alpha), log_alpha, rtol=1e-4, atol=1e-6)
        # keep the two sides on one trajectory (fp32 vs fp64 drift is not what this test is about)
        actor, q_on, q_tg = a2, q2, t2
        log_alpha = float(state.params.log_alpha)
        a_tree = state.params.actor_params
        mu, nu = f64(a_tree.arena_mu), f64(a_tree.arena_nu)
        np.testing.assert_allclose(mu[:n_a], opt["actor"].mu, rtol=1e-3, atol=1e-8)
        opt["actor"].mu, opt["actor"].nu = mu[:n_a].copy(), nu[:n_a].copy()
        qsl = np.r_[lay["q1"]: lay["q1"] + n_q, lay["q2"]: lay["q2"] + n_q]
        opt["q"].mu, opt["q"].nu = mu[qsl].copy(), nu[qsl].copy()
        opt["alpha"].mu, opt["alpha"].nu = mu[lay["alpha"]: lay["alpha"] + 1].copy(), nu[lay["alpha"]: lay["alpha"] + 1].copy()
    counts = state.params.actor_params.arena_counts.cpu().tolist()
    assert counts[:4] == [2, 2, 2, 2] and counts[4:] == ([2, 2] if autotune else [0, 0])


@pytest.mark.parametrize("graph", [True, False])
def test_experiment_runs(graph):
    from stoix_b200.systems.sac import ff_sac

    torch.cuda.set_device(0)
    cfg = _cfg([f"arch.cuda_graph={graph}"])
    perf = ff_sac.run_experiment(cfg)
    assert np.isfinite(perf)


def test_learning_moves_towards_the_optimum():
    """The synthetic reward is -mean((a - tanh(obs[:A]))^2): a few hundred SAC updates must beat the initial policy."""
    from stoix_b200 import random as srandom
    from stoix_b200.systems.sac import ff_sac
    from stoix_b200.utils import make_env as environments
    from stoix_b200.utils.total_timestep_checker import check_total_timesteps

    torch.cuda.set_device(0)
    # fixed small temperature: with alpha = 1 the entropy term (|log_prob| ~ 6) swamps a reward of magnitude < 1
    cfg = _cfg(["arch.total_num_envs=256", "arch.total_timesteps=768000", "arch.num_evaluation=3", "system.actor_lr=1e-3", "system.q_lr=1e-3",
                "system.gamma=0.0", "system.autotune=False", "system.init_alpha=0.01"])
    cfg.num_devices, cfg.rank = 1, 0
    cfg = check_total_timesteps(cfg, quiet=True)
    env, _ = environments.make(cfg)
    learn, actor_network, state = ff_sac.learner_setup(env, tuple(srandom.split(srandom.PRNGKey(0), 3)), cfg)
    rewards = []
    rewards.append(float(env.step(state.env_state[0], actor_network.apply(state.params.actor_params, state.timestep[0].observation).sample(seed=1))[1]
                         .reward.mean()))
    for _ in range(3):
        out = learn(state)
        state = out.learner_state
        rewards.append(float(out.episode_metrics["episode_return"].sum() * 0 + learn.built["shards"][0].reward.mean()))
    assert np.isfinite(rewards).all() and rewards[-1] > rewards[0] + 0.1, rewards
    assert learn.built["graph"] is not None
    st = learn.built["shards"][0].buffer.ring.state.cpu().tolist()
    assert st[1] == min(4096, (4 + 3 * cfg.arch.num_updates_per_eval) * 256)
