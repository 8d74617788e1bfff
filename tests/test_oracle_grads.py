"""The oracle's hand-written backward pass, losses and optimiser are not pinned by any reference test
('parity unpinned').  They are cross-checked here against torch.autograd / torch.optim on
CPU in float64, restating the reference's loss code (stoix/utils/loss.py:17-32,68-78;
ff_ppo.py:191-235) with torch ops."""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as O


def _torch_mlp(params, x):
    h = x
    n = len(params) // 2
    for i in range(n):
        h = h @ params[2 * i] + params[2 * i + 1]
        if i < n - 1:
            h = torch.relu(h)
    return h


def _setup(seed=0, m=64, D=12, H=(32, 24), A=5):
    rng = np.random.default_rng(seed)
    actor = O.init_mlp(rng, [D, *H, A], 0.01)
    critic = O.init_mlp(rng, [D, *H, 1], 1.0)
    for p in (actor, critic):  # non-zero biases and larger head so every term matters
        for i in range(len(p.W)):
            p.b[i] = rng.standard_normal(p.b[i].shape) * 0.1
        p.W[-1] = rng.standard_normal(p.W[-1].shape) * 0.5
    x = rng.standard_normal((m, D))
    act = rng.integers(0, A, m)
    logp_old = np.log(rng.uniform(0.05, 0.9, m))
    adv = rng.standard_normal(m)
    v_old = rng.standard_normal(m)
    tgt = rng.standard_normal(m)
    return actor, critic, x, act, logp_old, adv, v_old, tgt


def test_actor_grads_match_autograd():
    actor, _, x, act, logp_old, adv, _, _ = _setup()
    eps, ent = 0.2, 0.01
    logits, acts = O.mlp_forward(actor, x)
    total, dlogits, info = O.actor_loss_and_dlogits(logits, act, logp_old, adv, eps, ent)
    g = O.mlp_backward(actor, acts, dlogits)
    tp = [torch.tensor(a, requires_grad=True) for pair in zip(actor.W, actor.b) for a in pair]
    tl = _torch_mlp(tp, torch.tensor(x))
    lp = torch.log_softmax(tl, -1)
    logp = lp.gather(1, torch.tensor(act)[:, None])[:, 0]
    ratio = torch.exp(logp - torch.tensor(logp_old))
    tadv = torch.tensor(adv)
    loss = -torch.minimum(ratio * tadv, torch.clamp(ratio, 1 - eps, 1 + eps) * tadv).mean()
    entropy = -(lp.exp() * lp).sum(-1).mean()
    ttotal = loss - ent * entropy
    ttotal.backward()
    np.testing.assert_allclose(total, ttotal.item(), rtol=1e-12)
    np.testing.assert_allclose(info["actor_loss"], loss.item(), rtol=1e-12)
    np.testing.assert_allclose(info["entropy"], entropy.item(), rtol=1e-12)
    for i in range(len(actor.W)):
        np.testing.assert_allclose(g.W[i], tp[2 * i].grad.numpy(), rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(g.b[i], tp[2 * i + 1].grad.numpy(), rtol=1e-9, atol=1e-14)
    # both clip branches must be exercised by the data
    r = ratio.detach().numpy()
    assert (r > 1 + eps).any() and (r < 1 - eps).any() and ((r > 1 - eps) & (r < 1 + eps)).any()


def test_critic_grads_match_autograd():
    _, critic, x, _, _, _, v_old, tgt = _setup(1)
    eps, vf = 0.2, 0.5
    v, acts = O.mlp_forward(critic, x)
    total, dvalue, info = O.critic_loss_and_dvalue(v[:, 0], v_old, tgt, eps, vf)
    g = O.mlp_backward(critic, acts, dvalue[:, None])
    tp = [torch.tensor(a, requires_grad=True) for pair in zip(critic.W, critic.b) for a in pair]
    tv = _torch_mlp(tp, torch.tensor(x))[:, 0]
    tvo, ttg = torch.tensor(v_old), torch.tensor(tgt)
    vclip = tvo + (tv - tvo).clamp(-eps, eps)
    vl = 0.5 * torch.maximum((tv - ttg) ** 2, (vclip - ttg) ** 2).mean()
    (vf * vl).backward()
    np.testing.assert_allclose(info["value_loss"], vl.item(), rtol=1e-12)
    for i in range(len(critic.W)):
        np.testing.assert_allclose(g.W[i], tp[2 * i].grad.numpy(), rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(g.b[i], tp[2 * i + 1].grad.numpy(), rtol=1e-9, atol=1e-14)


def test_loss_functions_match_reference_formulas():
    rng = np.random.default_rng(3)
    lp, lpo, adv = rng.standard_normal(50) * 0.3, rng.standard_normal(50) * 0.3, rng.standard_normal(50)
    ratio = np.exp(lp - lpo)
    ref = -np.minimum(ratio * adv, np.clip(ratio, 0.8, 1.2) * adv).mean()
    np.testing.assert_allclose(O.ppo_clip_loss(lp, lpo, adv, 0.2), ref)
    v, vo, tg = rng.standard_normal(50), rng.standard_normal(50), rng.standard_normal(50)
    vc = vo + np.clip(v - vo, -0.2, 0.2)
    np.testing.assert_allclose(O.clipped_value_loss(v, vo, tg, 0.2), 0.5 * np.maximum((v - tg) ** 2, (vc - tg) ** 2).mean())


def test_adam_matches_torch_adam_without_clipping():
    rng = np.random.default_rng(4)
    p0 = rng.standard_normal(100)
    st = O.AdamState(np.zeros(100), np.zeros(100))
    tp = torch.tensor(p0.copy(), requires_grad=True)
    opt = torch.optim.Adam([tp], lr=3e-4, betas=(0.9, 0.999), eps=1e-5)
    p = p0.copy()
    for _ in range(5):
        g = rng.standard_normal(100) * 1e-3  # norm << max_grad_norm: clip inactive
        p, gn = O.clip_adam_step(p, g, st, 3e-4, max_grad_norm=0.5)
        assert gn < 0.5
        tp.grad = torch.tensor(g)
        opt.step()
    np.testing.assert_allclose(p, tp.detach().numpy(), rtol=1e-10, atol=1e-14)


def test_clip_by_global_norm_semantics():
    p = np.zeros(4)
    g = np.array([3.0, 4.0, 0.0, 0.0])  # norm 5 >= 0.5 -> scaled to norm 0.5
    st = O.AdamState(np.zeros(4), np.zeros(4))
    _, gn = O.clip_adam_step(p, g, st, 1.0, 0.5)
    assert gn == 5.0
    np.testing.assert_allclose(st.mu, 0.1 * g / 5.0 * 0.5)


def test_linear_schedule_floor_division():
    # utils/training.py:24-26 with epochs*minibatches = 64, num_updates = 10
    assert O.linear_schedule(3e-4, 0, 10, 4, 16) == 3e-4
    assert O.linear_schedule(3e-4, 63, 10, 4, 16) == 3e-4
    np.testing.assert_allclose(O.linear_schedule(3e-4, 64, 10, 4, 16), 3e-4 * 0.9)
    np.testing.assert_allclose(O.linear_schedule(3e-4, 639, 10, 4, 16), 3e-4 * 0.1)
    assert O.linear_schedule(3e-4, 1000, 10, 4, 16, decay=False) == 3e-4


def test_derive_shapes():
    # total_timestep_checker.py:57-61,88-96,104-131 with the default config values
    assert O.derive_shapes(1024, 1, 1, 1e7, 128, 20) == (1024, 76, 3)
    assert O.derive_shapes(4096 * 8, 8, 1, 4096 * 8 * 128 * 10, 128, 5) == (4096, 10, 2)


@pytest.mark.parametrize("use_layer_norm", [False, True])
@pytest.mark.parametrize("activation", ["relu", "tanh", "silu", "elu", "gelu", "sigmoid", "softplus", "identity"])
def test_torso_options_forward_backward_match_autograd(activation, use_layer_norm):
    """MLPTorso(activation, use_layer_norm) of stoix/networks/torso.py:24-33 (flax Dense / LayerNorm(eps 1e-6) /
    nn.<activation>, gelu in its default tanh approximation): the oracle's forward and hand-derived backward against
    torch.nn.functional + autograd in float64."""
    F = torch.nn.functional
    acts_t = {"relu": torch.relu, "tanh": torch.tanh, "silu": F.silu, "elu": F.elu, "gelu": lambda z: F.gelu(z, approximate="tanh"),
              "sigmoid": torch.sigmoid, "softplus": F.softplus, "identity": lambda z: z}
    rng = np.random.default_rng(len(activation) + int(use_layer_norm))
    sizes = [6, 16, 16, 3]
    W = [rng.standard_normal((sizes[i], sizes[i + 1])) * 0.5 for i in range(3)]
    b = [rng.standard_normal(sizes[i + 1]) * 0.3 + (1.0 if use_layer_norm and i < 2 else 0.0) for i in range(3)]
    lnb = [rng.standard_normal(sizes[i + 1]) * 0.3 if i < 2 else None for i in range(3)] if use_layer_norm else None
    p = O.MLPParams(W, b, activation, lnb)
    assert np.array_equal(O.MLPParams.from_flat(p.flat(), sizes, activation, use_layer_norm).flat(), p.flat())
    x, dout = rng.standard_normal((9, 6)), rng.standard_normal((9, 3))
    out, cache = O.mlp_forward(p, x)
    g = O.mlp_backward(p, cache, dout)
    Wt = [torch.tensor(w, requires_grad=True) for w in W]
    bt = [torch.tensor(v, requires_grad=True) for v in b]
    lt = [torch.tensor(v, requires_grad=True) if v is not None else None for v in (lnb or [None] * 3)]
    h = torch.tensor(x)
    for i in range(3):
        z = F.layer_norm(h @ Wt[i], (sizes[i + 1],), bt[i], lt[i], eps=1e-6) if (use_layer_norm and i < 2) else h @ Wt[i] + bt[i]
        h = acts_t[activation](z) if i < 2 else z
    (h * torch.tensor(dout)).sum().backward()
    np.testing.assert_allclose(out, h.detach().numpy(), rtol=1e-12, atol=1e-12)
    for i in range(3):
        np.testing.assert_allclose(g.W[i], Wt[i].grad.numpy(), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(g.b[i], bt[i].grad.numpy(), rtol=1e-10, atol=1e-12)
        if use_layer_norm and i < 2:
            np.testing.assert_allclose(g.ln_bias[i], lt[i].grad.numpy(), rtol=1e-10, atol=1e-12)
