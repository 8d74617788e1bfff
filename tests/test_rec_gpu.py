"""rec_ppo on the GPU, fp32, against oracle/rec_oracle.py (itself checked against torch autograd in
tests/test_oracle_rec.py):

* GRU sequence kernels (forward with resets, backward through time incl. d h0, weighted / accumulated parameter gradients);
* PPO loss heads on precomputed outputs with a row gather;
* RecurrentActor / RecurrentCritic faces (parameter tree with flax's names, multi-step apply);
* whole update steps of the learner (rollout with carried hidden states, GAE from the stored flags, epochs x minibatches over
  column subsets of the chunked batch, clip + Adam) vs the oracle replaying the same trajectory and permutations;
* a short experiment through run_experiment on CartPole."""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as O
from oracle import rec_oracle as R

pytestmark = pytest.mark.gpu

f64 = lambda t: t.detach().float().cpu().numpy().astype(np.float64)
dev = lambda x, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(x), device="cuda").to(dt)
rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("persistent", ["cluster", "1", "0"])
@pytest.mark.parametrize("T,E,H", [(7, 33, 20), (1, 64, 128), (16, 300, 128), (5, 1030, 6), (40, 130, 128)])
def test_gru_sequence_matches_oracle(T, E, H, persistent):
    """"cluster" (default): 2-CTA thread-block clusters, half of the hidden units / gate columns per CTA, h through distributed shared
    memory (sequences of >= 4 steps, < 1024 rows); "1": one CTA per 4 / 8 sequences with the whole W_h in its shared memory;
    "0": the per-step form (a GEMM + a gate kernel per step), the fallback for hidden sizes whose W_h does not fit."""
    import subprocess, sys, os
    if persistent != "cluster":   # the switches are read once per process: run these cases in a child
        code = (f"import os; os.environ['STX_GRU_CLUSTER']='0'; os.environ['STX_GRU_PERSISTENT']='{persistent}'; import sys; sys.path.insert(0, os.getcwd());"
                f"import tests.test_rec_gpu as m; m._gru_case({T}, {E}, {H}); print('child-ok')")
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert "child-ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
        return
    _gru_case(T, E, H)


def _gru_case(T, E, H):
    from stoix_b200 import ops

    rng = np.random.default_rng(T * 100 + E)
    gi, h0 = rng.standard_normal((T, E, 3 * H)).astype(np.float32), rng.standard_normal((E, H)).astype(np.float32)
    Wh, bhn = (rng.standard_normal((H, 3 * H)) * 0.2).astype(np.float32), (rng.standard_normal(H) * 0.1).astype(np.float32)
    reset = rng.random((T, E)) < 0.2
    d_h = rng.standard_normal((T, E, H)).astype(np.float32)
    ws = ops.gru_workspace(T, E, H, "cuda")
    h_seq = ops.gru_sequence_forward(dev(gi), dev(reset, torch.uint8), dev(h0), dev(Wh), dev(bhn), ws)
    hs_o, cache = R.gru_forward(gi.astype(np.float64), reset, h0.astype(np.float64), Wh.astype(np.float64), bhn.astype(np.float64))
    np.testing.assert_allclose(f64(h_seq), hs_o, rtol=2e-5, atol=2e-6)
    d_gi, d_wh, d_bhn, d_h0 = torch.zeros(T, E, 3 * H, device="cuda"), torch.full((H, 3 * H), 7.0, device="cuda"), torch.full((H,), 7.0, device="cuda"), torch.zeros(E, H, device="cuda")
    ops.gru_sequence_backward(dev(d_h), dev(reset, torch.uint8), dev(Wh), ws, d_gi, d_w_h=d_wh, d_b_hn=d_bhn, grad_weight=0.5, overwrite=False, d_h0=d_h0)
    dgi_o, dWh_o, dbhn_o, dh0_o = R.gru_backward(cache, reset, d_h.astype(np.float64), Wh.astype(np.float64))
    np.testing.assert_allclose(f64(d_gi), dgi_o, rtol=1e-4, atol=1e-5)
    assert rel(f64(d_wh), 7.0 + 0.5 * dWh_o) < 1e-5 and rel(f64(d_bhn), 7.0 + 0.5 * dbhn_o) < 1e-5
    np.testing.assert_allclose(f64(d_h0), dh0_o, rtol=1e-4, atol=1e-5)
    ops.gru_sequence_backward(dev(d_h), dev(reset, torch.uint8), dev(Wh), ws, d_gi, d_w_h=d_wh, d_b_hn=d_bhn)     # overwrite, weight 1, no d_h0
    assert rel(f64(d_wh), dWh_o) < 1e-5 and rel(f64(d_bhn), dbhn_o) < 1e-5


def test_ppo_head_grads_match_oracle():
    from stoix_b200 import ops

    rng = np.random.default_rng(3)
    N, mb, A = 500, 192, 5
    idx = rng.permutation(N)[:mb].astype(np.int32)
    logits, value = rng.standard_normal((mb, A)).astype(np.float32), rng.standard_normal(mb).astype(np.float32)
    action, logp_old = rng.integers(0, A, N).astype(np.int32), (-rng.random(N) - 0.3).astype(np.float32)
    v_old, adv, tgt = rng.standard_normal(N).astype(np.float32), rng.standard_normal(N).astype(np.float32), rng.standard_normal(N).astype(np.float32)
    stats = np.array([adv.mean(), 1.0 / np.sqrt(adv.var() + 1e-5)], np.float32)
    d_logits, d_value, metrics = torch.zeros(mb, A, device="cuda"), torch.zeros(mb, device="cuda"), torch.zeros(8, device="cuda")
    ops.ppo_head_grads(dev(logits), None, dev(idx, torch.int32), dev(action, torch.int32), dev(logp_old), dev(v_old), dev(adv), dev(tgt), dev(stats),
                       0.2, 0.01, 0.5, d_logits, None, metrics)
    ops.ppo_head_grads(None, dev(value), dev(idx, torch.int32), dev(action, torch.int32), dev(logp_old), dev(v_old), dev(adv), dev(tgt), dev(stats),
                       0.2, 0.01, 0.5, None, d_value, metrics)
    a_std = (adv[idx].astype(np.float64) - stats[0]) * stats[1]
    _, dlg, ai = O.actor_loss_and_dlogits(logits.astype(np.float64), action[idx], logp_old[idx].astype(np.float64), a_std, 0.2, 0.01)
    _, dv, ci = O.critic_loss_and_dvalue(value.astype(np.float64), v_old[idx].astype(np.float64), tgt[idx].astype(np.float64), 0.2, 0.5)
    np.testing.assert_allclose(f64(d_logits), dlg, rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(f64(d_value), dv, rtol=1e-4, atol=1e-8)
    m = f64(metrics)
    np.testing.assert_allclose(m[:3], [ai["actor_loss"], ai["entropy"], ci["value_loss"]], rtol=1e-4, atol=1e-6)


def _cfg(extra=()):
    from stoix_b200.config import compose
    from stoix_b200.utils.total_timestep_checker import check_total_timesteps

    c = compose("default_rec_ppo", ["env=synthetic/box", "env.kwargs.obs_dim=12", "env.kwargs.num_actions=5", "env.kwargs.p_term=0.06", "env.kwargs.p_trunc=0.04",
                                    "arch.total_num_envs=32", "system.rollout_length=16", "system.num_minibatches=4", "system.epochs=2",
                                    "arch.total_timesteps=1536", "arch.num_evaluation=1", "logger.use_console=False",
                                    "network.actor_network.pre_torso.layer_sizes=[24]", "network.actor_network.post_torso.layer_sizes=[20]",
                                    "network.critic_network.pre_torso.layer_sizes=[24]", "network.critic_network.post_torso.layer_sizes=[20]",
                                    "network.actor_network.rnn_layer.hidden_state_dim=16", "network.critic_network.rnn_layer.hidden_state_dim=16"]
                + list(extra), config_dir="default/anakin")
    c.num_devices, c.rank = 1, 0
    return check_total_timesteps(c, quiet=True)


def _setup(cfg):
    from stoix_b200 import random as srandom
    from stoix_b200.systems.ppo.anakin import rec_ppo
    from stoix_b200.utils import make_env

    torch.cuda.set_device(0)
    env, _ = make_env.make(cfg)
    learn, actor_network, state = rec_ppo.learner_setup(env, tuple(srandom.split(srandom.PRNGKey(5), 3)), cfg)
    with torch.no_grad():
        g = torch.Generator(device="cuda").manual_seed(2)
        a = state.params.actor_params
        a.arena.add_(torch.randn(a.arena.shape, device="cuda", generator=g) * 0.05)
    return rec_ppo, learn, actor_network, state


def _oracle_net(tree) -> R.RecNet:
    lay = tree.layout
    return R.RecNet.from_flat(f64(tree.flat), lay.spec_pre.sizes, lay.H, lay.spec_post.sizes, lay.spec_pre.activation)


def test_recurrent_faces_and_parameter_tree():
    cfg = _cfg()
    rec_ppo, learn, actor_network, state = _setup(cfg)
    tree = state.params.actor_params
    p = tree["params"]
    assert set(p) == {"pre_torso", "ScannedRNN_0", "post_torso", "action_head"}
    cell = p["ScannedRNN_0"]["GRUCell_0"]
    assert set(cell) == {"ir", "iz", "in", "hr", "hz", "hn"} and "bias" not in cell["hr"] and "bias" in cell["hn"]
    assert tuple(cell["ir"]["kernel"].shape) == (24, 16) and tuple(cell["hn"]["kernel"].shape) == (16, 16)
    net = _oracle_net(tree)
    np.testing.assert_array_equal(f64(cell["iz"]["kernel"]), net.pre.W[1][:, 16:32])
    np.testing.assert_array_equal(f64(cell["hn"]["bias"]), net.bhn)
    rng = np.random.default_rng(0)
    T, E = 3, 32
    obs, h0 = rng.standard_normal((T, E, 12)).astype(np.float32), rng.standard_normal((E, 16)).astype(np.float32)
    done = rng.random((T, E)) < 0.3
    h_new, pi = actor_network.apply(tree, dev(h0), (dev(obs), dev(done, torch.bool)))
    out_o, h_o, _ = R.rec_forward(net, h0.astype(np.float64), obs.astype(np.float64), done)
    np.testing.assert_allclose(f64(h_new), h_o, rtol=1e-4, atol=1e-5)
    a = pi.sample(seed=1)
    np.testing.assert_allclose(f64(pi.log_prob(a)), O.categorical_log_prob(out_o, a.cpu().numpy()), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("chunk,cell", [(None, "gru"), (8, "gru"), (8, "lstm")])
def test_update_steps_match_oracle(chunk, cell):
    from stoix_b200 import ops

    cfg = _cfg(([] if chunk is None else [f"system.recurrent_chunk_size={chunk}"])
               + [f"network.actor_network.rnn_layer.cell_type={cell}", f"network.critic_network.rnn_layer.cell_type={cell}"])
    S = 16 if cell == "gru" else 32     # carry width: lstm carries (c | h)
    rec_ppo, learn, actor_network, state = _setup(cfg)
    cfg.arch.num_updates_per_eval = 1     # one update per learn() call; arch.num_updates (3) still drives the LR schedule
    T, E, nmb, epochs = 16, 32, 4, 2
    ch = T if chunk is None else chunk
    nc = T // ch
    actor, critic = _oracle_net(state.params.actor_params), _oracle_net(state.params.critic_params)
    n_a, n_c = actor.flat().size, critic.flat().size
    a_st, c_st = O.AdamState(np.zeros(n_a), np.zeros(n_a)), O.AdamState(np.zeros(n_c), np.zeros(n_c))
    hyp = O.PPOHyper(ent_coef=float(cfg.system.ent_coef), actor_lr=float(cfg.system.actor_lr), critic_lr=float(cfg.system.critic_lr), epochs=epochs,
                     num_minibatches=nmb, num_updates=int(cfg.arch.num_updates))
    h_a, h_c = np.zeros((E, S)), np.zeros((E, S))
    for upd in range(2):
        out = learn(state)
        state = out.learner_state
        torch.cuda.synchronize()
        b = learn.built
        sh = b["shards"][0]
        # ---- replay the rollout with the kernels' actions: values / log-probs / hidden states must match (rec_ppo.py:69-143) ----
        obs, done, trunc = f64(sh.obs), sh.done.cpu().numpy().astype(bool), sh.trunc.cpu().numpy().astype(bool)
        if upd == 0:
            assert not done[0].any() and not trunc[0].any()
        else:   # row 0 of this rollout = row T of the previous one (observation and flags carried over)
            np.testing.assert_array_equal(obs[0], carry[0]), np.testing.assert_array_equal(done[0], carry[1]), np.testing.assert_array_equal(trunc[0], carry[2])
        action, reward = sh.action.cpu().numpy(), f64(sh.reward)
        val, lp, hs_a, hs_c = np.zeros((T, E)), np.zeros((T, E)), np.zeros((T, E, S)), np.zeros((T, E, S))
        for t in range(T):
            reset = (done[t] | trunc[t])[None]
            lg, h_a, _ = R.rec_forward(actor, h_a, obs[t][None], reset)
            v, h_c, _ = R.rec_forward(critic, h_c, obs[t][None], reset)
            lp[t], val[t] = O.categorical_log_prob(lg[0], action[t]), v[0, :, 0]
            hs_a[t], hs_c[t] = h_a, h_c
        v_last, _, _ = R.rec_forward(critic, h_c, obs[T][None], (done[T] | trunc[T])[None])
        last_val = np.where(done[T], 0.0, v_last[0, :, 0])
        assert (done[1:] | trunc[1:]).any(), "the test wants episode boundaries inside the rollout"
        np.testing.assert_allclose(f64(sh.value[:T]), val, rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(f64(sh.log_prob), lp, rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(f64(sh.h_actor), hs_a, rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(f64(sh.h_critic), hs_c, rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(f64(sh.value[T]), last_val, rtol=2e-4, atol=2e-5)
        traj = R.RecTrajectory(obs=obs[:T], done=done[:T], truncated=trunc[:T], action=action, value=f64(sh.value[:T]), reward=reward, log_prob=f64(sh.log_prob),
                               h_actor=f64(sh.h_actor), h_critic=f64(sh.h_critic), last_val=f64(sh.value[T]))
        adv, tgt = R.rec_gae(traj, hyp.gamma, hyp.gae_lambda, True)
        np.testing.assert_allclose(f64(sh.targets), tgt, rtol=1e-4, atol=2e-5)
        raw = tgt - traj.value
        np.testing.assert_allclose(f64(sh.advantages), raw, rtol=1e-4, atol=2e-5)      # stored raw, standardised on load
        # ---- epochs x minibatches with the learner's permutations ----
        for ep in range(epochs):
            perm = ops.make_permutation(E * nc, state.key[1], ep + epochs * upd, device="cuda").cpu().numpy()
            for i in range(nmb):
                cols = perm[i * (E * nc // nmb):(i + 1) * (E * nc // nmb)]
                ga, gc, _ = R.rec_minibatch_grads(actor, critic, traj, adv, tgt, cols, ch, hyp)
                k = (a_st.sched_count // (epochs * nmb))
                lr_scale = 1.0 - k / hyp.num_updates
                pa, _ = O.clip_adam_step(actor.flat(), ga.flat(), a_st, hyp.actor_lr * lr_scale, hyp.max_grad_norm)
                pc, _ = O.clip_adam_step(critic.flat(), gc.flat(), c_st, hyp.critic_lr * lr_scale, hyp.max_grad_norm)
                lay_a, lay_c = state.params.actor_params.layout, state.params.critic_params.layout
                actor = R.RecNet.from_flat(pa, lay_a.spec_pre.sizes, lay_a.H, lay_a.spec_post.sizes, lay_a.spec_pre.activation)
                critic = R.RecNet.from_flat(pc, lay_c.spec_pre.sizes, lay_c.H, lay_c.spec_post.sizes, lay_c.spec_pre.activation)
        np.testing.assert_allclose(f64(state.params.actor_params.flat), actor.flat(), rtol=2e-4, atol=5e-6)
        np.testing.assert_allclose(f64(state.params.critic_params.flat), critic.flat(), rtol=2e-4, atol=5e-6)
        # continue both sides from the kernels' state (parameters, moments, hidden states): the next update is checked on its own
        actor, critic = _oracle_net(state.params.actor_params), _oracle_net(state.params.critic_params)
        a_tree = state.params.actor_params
        mu, nu = f64(a_tree.arena_mu), f64(a_tree.arena_nu)
        coff = b["coff"]
        a_st.mu, a_st.nu, c_st.mu, c_st.nu = mu[:n_a].copy(), nu[:n_a].copy(), mu[coff:coff + n_c].copy(), nu[coff:coff + n_c].copy()
        h_a, h_c = f64(sh.h_a_cur), f64(sh.h_c_cur)
        carry = (obs[T].copy(), done[T].copy(), trunc[T].copy())
    assert state.params.actor_params.arena_counts.cpu().tolist() == [2 * epochs * nmb] * 4


def test_experiment_runs_on_cartpole():
    from stoix_b200.config import compose
    from stoix_b200.systems.ppo.anakin import rec_ppo

    torch.cuda.set_device(0)
    cfg = compose("default_rec_ppo", ["arch.total_num_envs=64", "system.rollout_length=16", "system.num_minibatches=4", "system.epochs=2",
                                      "arch.total_timesteps=8192", "arch.num_evaluation=2", "arch.num_eval_episodes=16", "arch.max_eval_steps=100",
                                      "logger.use_console=False"], config_dir="default/anakin")
    assert np.isfinite(rec_ppo.run_experiment(cfg))


@pytest.mark.parametrize("cluster", ["1", "0"])
@pytest.mark.parametrize("T,E,H", [(6, 40, 12), (1, 64, 128), (9, 130, 32), (40, 131, 128), (3, 2050, 128)])
def test_lstm_sequence_matches_oracle(T, E, H, cluster):
    """cluster = "1": the persistent form on 2-CTA thread-block clusters (half of the hidden units and of W_h per CTA, h exchanged
    through distributed shared memory every step; the default); "0": one GEMM + one gate kernel per step (fallback)."""
    import os, subprocess, sys
    if cluster == "0":   # the switch is read once per process: run this case in a child
        code = ("import os; os.environ['STX_LSTM_CLUSTER']='0'; import sys; sys.path.insert(0, os.getcwd());"
                f"import tests.test_rec_gpu as m; m._lstm_case({T}, {E}, {H}); print('child-ok')")
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert "child-ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
        return
    _lstm_case(T, E, H)


def _lstm_case(T, E, H):
    from stoix_b200 import ops

    rng = np.random.default_rng(T * 10 + H)
    gi, c0 = rng.standard_normal((T, E, 4 * H)).astype(np.float32), rng.standard_normal((E, 2 * H)).astype(np.float32)
    Wh = (rng.standard_normal((H, 4 * H)) * 0.2).astype(np.float32)
    reset = rng.random((T, E)) < 0.2
    d_h = rng.standard_normal((T, E, H)).astype(np.float32)
    ws = ops.lstm_workspace(T, E, H, "cuda")
    last = torch.zeros(E, 2 * H, device="cuda")
    h_seq = ops.lstm_sequence_forward(dev(gi), dev(reset, torch.uint8), dev(c0), dev(Wh), ws, carry_last=last)
    hs_o, last_o, cache = R.lstm_forward(gi.astype(np.float64), reset, c0.astype(np.float64), Wh.astype(np.float64))
    np.testing.assert_allclose(f64(h_seq), hs_o, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(f64(last), last_o, rtol=2e-5, atol=2e-6)
    d_gi, d_wh, d_c0 = torch.zeros(T, E, 4 * H, device="cuda"), torch.full((H, 4 * H), 3.0, device="cuda"), torch.zeros(E, 2 * H, device="cuda")
    ops.lstm_sequence_backward(dev(d_h), dev(reset, torch.uint8), dev(Wh), ws, d_gi, d_w_h=d_wh, grad_weight=0.25, overwrite=False, d_carry0=d_c0)
    dgi_o, dWh_o, dc0_o = R.lstm_backward(cache, reset, d_h.astype(np.float64), Wh.astype(np.float64))
    np.testing.assert_allclose(f64(d_gi), dgi_o, rtol=1e-4, atol=1e-5)
    assert rel(f64(d_wh), 3.0 + 0.25 * dWh_o) < 1e-5
    np.testing.assert_allclose(f64(d_c0), dc0_o, rtol=1e-4, atol=1e-5)
