"""oracle/rec_oracle.py (recurrent PPO restatement, parity unpinned by the reference) against torch.autograd on CPU: the GRU
recurrence with episode resets (flax GRUCell formula written in torch ops) and the whole recurrent actor / critic loss."""
import numpy as np
import torch

from oracle import ppo_oracle as O
from oracle import rec_oracle as R


def _net(rng, D, P, H, Q, A):
    pre = O.MLPParams([rng.standard_normal((D, P)) * 0.3, rng.standard_normal((P, 3 * H)) * 0.3],
                      [rng.standard_normal(P) * 0.1, rng.standard_normal(3 * H) * 0.1], "silu")
    post = O.MLPParams([rng.standard_normal((H, Q)) * 0.3, rng.standard_normal((Q, A)) * 0.3],
                       [rng.standard_normal(Q) * 0.1, rng.standard_normal(A) * 0.1], "silu")
    return R.RecNet(pre, rng.standard_normal((H, 3 * H)) * 0.3, rng.standard_normal(H) * 0.1, post)


def _torch_forward(tn, h0, obs, reset, H):
    silu = torch.nn.functional.silu
    T = obs.shape[0]
    x = silu(obs @ tn["W0"] + tn["b0"])
    gi = x @ tn["Wi"] + tn["bi"]
    h, out = h0, []
    for t in range(T):
        hp = torch.where(reset[t][:, None], torch.zeros_like(h), h)
        gh = hp @ tn["Wh"]
        r = torch.sigmoid(gi[t][:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[t][:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[t][:, 2 * H:] + r * (gh[:, 2 * H:] + tn["bhn"]))
        h = (1 - z) * n + z * hp
        out.append(h)
    hs = torch.stack(out)
    y = silu(hs @ tn["W1"] + tn["b1"])
    return y @ tn["W2"] + tn["b2"], hs


def _to_torch(net):
    t = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    return {"W0": t(net.pre.W[0]), "b0": t(net.pre.b[0]), "Wi": t(net.pre.W[1]), "bi": t(net.pre.b[1]), "Wh": t(net.Wh), "bhn": t(net.bhn),
            "W1": t(net.post.W[0]), "b1": t(net.post.b[0]), "W2": t(net.post.W[1]), "b2": t(net.post.b[1])}


def test_recurrent_network_gradients_match_autograd():
    rng = np.random.default_rng(0)
    T, E, D, P, H, Q, A = 9, 6, 5, 7, 4, 6, 3
    net = _net(rng, D, P, H, Q, A)
    obs, h0 = rng.standard_normal((T, E, D)), rng.standard_normal((E, H))
    reset = rng.random((T, E)) < 0.25
    reset[0, :2] = True
    d_out = rng.standard_normal((T, E, A))
    out, h_last, cache = R.rec_forward(net, h0, obs, reset)
    g = R.rec_backward(net, cache, d_out)
    tn = _to_torch(net)
    to, hs = _torch_forward(tn, torch.tensor(h0), torch.tensor(obs), torch.tensor(reset), H)
    np.testing.assert_allclose(out, to.detach().numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(h_last, hs[-1].detach().numpy(), rtol=1e-10, atol=1e-12)
    (to * torch.tensor(d_out)).sum().backward()
    for name, got in (("W0", g.pre.W[0]), ("b0", g.pre.b[0]), ("Wi", g.pre.W[1]), ("bi", g.pre.b[1]), ("Wh", g.Wh), ("bhn", g.bhn),
                      ("W1", g.post.W[0]), ("b1", g.post.b[0]), ("W2", g.post.W[1]), ("b2", g.post.b[1])):
        np.testing.assert_allclose(got, tn[name].grad.numpy(), rtol=1e-9, atol=1e-11, err_msg=name)
    # flat layout round trip
    net2 = R.RecNet.from_flat(net.flat(), (D, P, 3 * H), H, (H, Q, A))
    np.testing.assert_array_equal(net2.flat(), net.flat())


def test_gru_backward_initial_state_gradient_and_reset_blocks_it():
    rng = np.random.default_rng(1)
    T, E, H = 5, 4, 3
    gi, h0 = rng.standard_normal((T, E, 3 * H)), rng.standard_normal((E, H))
    Wh, bhn = rng.standard_normal((H, 3 * H)) * 0.4, rng.standard_normal(H) * 0.1
    reset = np.zeros((T, E), bool)
    reset[2, 1] = True          # env 1 forgets everything before t = 2
    reset[0, 3] = True          # env 3 never sees h0
    d_h = rng.standard_normal((T, E, H))
    hs, cache = R.gru_forward(gi, reset, h0, Wh, bhn)
    d_gi, dWh, dbhn, dh0 = R.gru_backward(cache, reset, d_h, Wh)
    th0 = torch.tensor(h0, requires_grad=True)
    tgi = torch.tensor(gi, requires_grad=True)
    tWh, tb = torch.tensor(Wh, requires_grad=True), torch.tensor(bhn, requires_grad=True)
    h, outs = th0, []
    for t in range(T):
        hp = torch.where(torch.tensor(reset[t])[:, None], torch.zeros_like(h), h)
        gh = hp @ tWh
        r = torch.sigmoid(tgi[t][:, :H] + gh[:, :H])
        z = torch.sigmoid(tgi[t][:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(tgi[t][:, 2 * H:] + r * (gh[:, 2 * H:] + tb))
        h = (1 - z) * n + z * hp
        outs.append(h)
    (torch.stack(outs) * torch.tensor(d_h)).sum().backward()
    np.testing.assert_allclose(hs, torch.stack(outs).detach().numpy(), rtol=1e-12)
    np.testing.assert_allclose(d_gi, tgi.grad.numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(dWh, tWh.grad.numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(dbhn, tb.grad.numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(dh0, th0.grad.numpy(), rtol=1e-9, atol=1e-12)
    assert np.all(dh0[3] == 0.0) and np.any(dh0[0] != 0.0)
    assert np.all(d_gi[:2, 1] != 0.0)   # before its reset env 1 still receives the gradient of d_h[:2]


def test_minibatch_gradients_match_autograd_with_reference_quirks():
    """Columns of the (chunk, E * num_chunks) reshape, initial state = stored state AFTER the first step, reset = done | truncated."""
    rng = np.random.default_rng(2)
    T, E, D, P, H, Q, A, chunk = 8, 5, 4, 6, 3, 5, 3, 4
    actor, critic = _net(rng, D, P, H, Q, A), _net(rng, D, P, H, Q, 1)
    traj = R.RecTrajectory(obs=rng.standard_normal((T, E, D)), done=rng.random((T, E)) < 0.2, truncated=rng.random((T, E)) < 0.1,
                           action=rng.integers(0, A, (T, E)), value=rng.standard_normal((T, E)), reward=rng.standard_normal((T, E)),
                           log_prob=-rng.random((T, E)) - 0.5, h_actor=rng.standard_normal((T, E, H)), h_critic=rng.standard_normal((T, E, H)),
                           last_val=rng.standard_normal(E))
    hyp = O.PPOHyper(ent_coef=0.01)
    adv, tgt = R.rec_gae(traj, hyp.gamma, hyp.gae_lambda, True)
    # GAE restatement: discount from the PREVIOUS done flag, values = [value, last_val]
    v = np.concatenate([traj.value, traj.last_val[None]], 0)
    disc = (1.0 - traj.done) * hyp.gamma
    acc, raw = np.zeros(E), np.zeros((T, E))
    for t in range(T - 1, -1, -1):
        acc = traj.reward[t] + disc[t] * v[t + 1] - v[t] + disc[t] * hyp.gae_lambda * acc
        raw[t] = acc
    np.testing.assert_allclose(tgt, raw + v[:-1], rtol=1e-10)
    np.testing.assert_allclose(adv, O.standardize(raw), rtol=1e-8, atol=1e-10)
    nc = T // chunk
    cols = rng.permutation(E * nc)[:6]
    ga, gc, info = R.rec_minibatch_grads(actor, critic, traj, adv, tgt, cols, chunk, hyp)
    r2 = lambda x: torch.tensor(x.reshape((chunk, E * nc) + x.shape[2:])[:, cols])
    ta, tc = _to_torch(actor), _to_torch(critic)
    reset = r2(traj.done | traj.truncated)
    logits, _ = _torch_forward(ta, r2(traj.h_actor)[0], r2(traj.obs), reset, H)
    lp = torch.log_softmax(logits, -1)
    logp = torch.gather(lp, -1, r2(traj.action)[..., None].long())[..., 0]
    ratio = torch.exp(logp - r2(traj.log_prob))
    a_ = r2(adv)
    loss_a = -torch.minimum(ratio * a_, torch.clamp(ratio, 0.8, 1.2) * a_).mean() - 0.01 * (-(lp.exp() * lp).sum(-1)).mean()
    loss_a.backward()
    val, _ = _torch_forward(tc, r2(traj.h_critic)[0], r2(traj.obs), reset, H)
    val = val[..., 0]
    vo, tg = r2(traj.value), r2(tgt)
    vclip = vo + torch.clamp(val - vo, -0.2, 0.2)
    loss_c = 0.5 * (0.5 * torch.maximum((val - tg) ** 2, (vclip - tg) ** 2).mean())
    loss_c.backward()
    for net, g, tn in ((actor, ga, ta), (critic, gc, tc)):
        for name, got in (("W0", g.pre.W[0]), ("bi", g.pre.b[1]), ("Wh", g.Wh), ("bhn", g.bhn), ("W1", g.post.W[0]), ("b2", g.post.b[1])):
            np.testing.assert_allclose(got, tn[name].grad.numpy(), rtol=1e-8, atol=1e-11, err_msg=name)
    np.testing.assert_allclose(info["value_loss"], float(loss_c) / 0.5, rtol=1e-10)


def test_lstm_recurrence_matches_autograd():
    """flax LSTMCell formula in torch ops, carry = (c | h) zeroed at resets, vs oracle lstm_forward / lstm_backward and the whole net."""
    rng = np.random.default_rng(4)
    T, E, D, P, H, Q, A = 7, 5, 4, 6, 3, 5, 2
    pre = O.MLPParams([rng.standard_normal((D, P)) * 0.3, rng.standard_normal((P, 4 * H)) * 0.3], [rng.standard_normal(P) * 0.1, rng.standard_normal(4 * H) * 0.1], "silu")
    post = O.MLPParams([rng.standard_normal((H, Q)) * 0.3, rng.standard_normal((Q, A)) * 0.3], [rng.standard_normal(Q) * 0.1, rng.standard_normal(A) * 0.1], "silu")
    net = R.RecNet(pre, rng.standard_normal((H, 4 * H)) * 0.3, np.zeros(0), post)
    assert net.is_lstm
    obs, carry0 = rng.standard_normal((T, E, D)), rng.standard_normal((E, 2 * H))
    reset = rng.random((T, E)) < 0.25
    d_out = rng.standard_normal((T, E, A))
    out, last, cache = R.rec_forward(net, carry0, obs, reset)
    g = R.rec_backward(net, cache, d_out)
    t = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    tn = {"W0": t(pre.W[0]), "b0": t(pre.b[0]), "Wi": t(pre.W[1]), "bi": t(pre.b[1]), "Wh": t(net.Wh), "W1": t(post.W[0]), "b1": t(post.b[0]),
          "W2": t(post.W[1]), "b2": t(post.b[1])}
    silu = torch.nn.functional.silu
    gi = silu(torch.tensor(obs) @ tn["W0"] + tn["b0"]) @ tn["Wi"] + tn["bi"]
    c, h = torch.tensor(carry0[:, :H]), torch.tensor(carry0[:, H:])
    hs = []
    for k in range(T):
        m = torch.tensor(reset[k])[:, None]
        c, h = torch.where(m, torch.zeros_like(c), c), torch.where(m, torch.zeros_like(h), h)
        z = gi[k] + h @ tn["Wh"]
        i_, f_, g_, o_ = torch.sigmoid(z[:, :H]), torch.sigmoid(z[:, H:2 * H]), torch.tanh(z[:, 2 * H:3 * H]), torch.sigmoid(z[:, 3 * H:])
        c = f_ * c + i_ * g_
        h = o_ * torch.tanh(c)
        hs.append(h)
    to = silu(torch.stack(hs) @ tn["W1"] + tn["b1"]) @ tn["W2"] + tn["b2"]
    np.testing.assert_allclose(out, to.detach().numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(last, torch.cat([c, h], 1).detach().numpy(), rtol=1e-10, atol=1e-12)
    (to * torch.tensor(d_out)).sum().backward()
    for name, got in (("W0", g.pre.W[0]), ("Wi", g.pre.W[1]), ("bi", g.pre.b[1]), ("Wh", g.Wh), ("W1", g.post.W[0]), ("b2", g.post.b[1])):
        np.testing.assert_allclose(got, tn[name].grad.numpy(), rtol=1e-9, atol=1e-11, err_msg=name)
    net2 = R.RecNet.from_flat(net.flat(), (D, P, 4 * H), H, (H, Q, A))
    np.testing.assert_array_equal(net2.flat(), net.flat())
    assert net2.is_lstm
