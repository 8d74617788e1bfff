"""The hand-derived gradients of oracle/sac_oracle.py (ff_sac losses, tanh-Normal head with the clipped tails, Q input
gradient, LayerNorm + silu twin-Q networks) against torch.autograd in float64, and the clip+Adam / Polyak update."""
import numpy as np
import pytest
import torch

from oracle import ppo_oracle as O
from oracle import sac_oracle as S


def _mk(rng, sizes, act, ln, head_scale=0.3):
    n = len(sizes) - 1
    W = [rng.standard_normal((sizes[i], sizes[i + 1])) * (0.4 if i < n - 1 else head_scale) for i in range(n)]
    b = [rng.standard_normal(sizes[i + 1]) * 0.1 + (1.0 if ln and i < n - 1 else 0.0) for i in range(n)]
    lnb = [rng.standard_normal(sizes[i + 1]) * 0.1 if i < n - 1 else None for i in range(n)] if ln else None
    return O.MLPParams(W, b, act, lnb)


def _torch_mlp(p, x):
    F = torch.nn.functional
    n = len(p["W"])
    h = x
    for i in range(n):
        if p["ln"] is not None and i < n - 1:
            z = F.layer_norm(h @ p["W"][i], (p["W"][i].shape[1],), p["b"][i], p["ln"][i], eps=1e-6)
        else:
            z = h @ p["W"][i] + p["b"][i]
        h = F.silu(z) if i < n - 1 else z
    return h


def _tp(p):
    t = lambda v: torch.tensor(v, dtype=torch.float64, requires_grad=True)
    return {"W": [t(w) for w in p.W], "b": [t(v) for v in p.b], "ln": None if p.ln_bias is None else [None if v is None else t(v) for v in p.ln_bias]}


def _torch_head(ho, eps, lo_a, hi_a, min_scale=1e-3, epsilon=1e-3):
    A = ho.shape[-1] // 2
    loc, raw = ho[:, :A], ho[:, A:]
    sigma = torch.nn.functional.softplus(raw) + min_scale
    u = loc + sigma * eps
    s, sh = (hi_a - lo_a) / 2.0, (lo_a + hi_a) / 2.0
    action = sh + s * torch.tanh(u)
    lo, hi = lo_a + epsilon, hi_a - epsilon
    ev = torch.clamp(action, lo, hi)
    uc = torch.atanh(torch.clamp((ev - sh) / s, -1 + 1e-15, 1 - 1e-15))
    z = (uc - loc) / sigma
    inner = -0.5 * z * z - torch.log(sigma) - S.LOG_SQRT_2PI - 2.0 * (np.log(2.0) - uc - torch.nn.functional.softplus(-2.0 * uc)) - np.log(s)
    u_lo, u_hi = np.arctanh((lo - sh) / s), np.arctanh((hi - sh) / s)
    left = torch.special.log_ndtr((u_lo - loc) / sigma) - np.log(epsilon)
    right = torch.special.log_ndtr(-(u_hi - loc) / sigma) - np.log(epsilon)
    per = torch.where(ev <= lo, left, torch.where(ev >= hi, right, inner))
    return action, per.sum(-1)


@pytest.mark.parametrize("wide", [False, True])   # wide: large sigma so that both clipped tails occur
def test_sac_losses_gradients_match_autograd(wide):
    rng = np.random.default_rng(3 + int(wide))
    D, A, B = 5, 3, 64
    actor = _mk(rng, [D, 16, 16, 2 * A], "silu", False, head_scale=1.5 if wide else 0.3)
    qs = (_mk(rng, [D + A, 16, 16, 1], "silu", True), _mk(rng, [D + A, 16, 16, 1], "silu", True))
    qt = (_mk(rng, [D + A, 16, 16, 1], "silu", True), _mk(rng, [D + A, 16, 16, 1], "silu", True))
    batch = {"obs": rng.standard_normal((B, D)), "action": rng.uniform(-1, 1, (B, A)), "reward": rng.standard_normal(B),
             "done": rng.random(B) < 0.2, "next_obs": rng.standard_normal((B, D))}
    noise = {k: rng.standard_normal((B, A)) * (2.5 if wide else 1.0) for k in ("actor", "q", "alpha")}
    h = S.SACHyper(target_entropy=-float(A))
    log_alpha = -0.7
    ag, qg, alg, info, aux = S.sac_losses_and_grads(actor, qs, qt, log_alpha, batch, noise, h)
    if wide:
        assert (aux["actor_sample"].branch != 0).any() and (aux["actor_sample"].branch == 0).any()

    ta, tq, tt = _tp(actor), [_tp(q) for q in qs], [_tp(q) for q in qt]
    T = lambda v: torch.tensor(np.asarray(v, np.float64))
    la = torch.tensor(log_alpha, dtype=torch.float64, requires_grad=True)
    obs, act, rew, done, nobs = T(batch["obs"]), T(batch["action"]), T(batch["reward"]), T(batch["done"].astype(float)), T(batch["next_obs"])
    alpha = torch.exp(la)
    qf = lambda ps, o, a: torch.stack([_torch_mlp(p, torch.cat([o, a], -1))[:, 0] for p in ps], -1)
    # actor loss: gradients w.r.t. the actor only (q params / alpha are constants there)
    a_new, lp = _torch_head(_torch_mlp(ta, obs), T(noise["actor"]), h.minimum, h.maximum)
    actor_loss = (alpha.detach() * lp - qf(tq, obs, a_new).min(-1).values).mean()
    g_actor = torch.autograd.grad(actor_loss, ta["W"] + ta["b"])
    # q loss
    with torch.no_grad():
        a_n, lp_n = _torch_head(_torch_mlp(ta, nobs), T(noise["q"]), h.minimum, h.maximum)
        target = rew + (1 - done) * h.gamma * (qf(tt, nobs, a_n).min(-1).values - alpha * lp_n)
    q_loss = 0.5 * ((qf(tq, obs, act) - target[:, None]) ** 2).mean()
    q_leaves = [x for p in tq for x in (p["W"] + p["b"] + [v for v in p["ln"] if v is not None])]
    g_q = torch.autograd.grad(q_loss, q_leaves)
    # alpha loss
    with torch.no_grad():
        _, lp_a = _torch_head(_torch_mlp(ta, obs), T(noise["alpha"]), h.minimum, h.maximum)
    alpha_loss = (torch.exp(la) * (-lp_a - h.target_entropy)).mean()
    g_alpha = torch.autograd.grad(alpha_loss, la)[0]

    np.testing.assert_allclose(info["actor_loss"], actor_loss.item(), rtol=1e-10)
    np.testing.assert_allclose(info["q_loss"], q_loss.item(), rtol=1e-10)
    np.testing.assert_allclose(alg, g_alpha.item(), rtol=1e-10)
    n = len(actor.W)
    for i in range(n):
        np.testing.assert_allclose(ag.W[i], g_actor[i].numpy(), rtol=1e-7, atol=1e-11)
        np.testing.assert_allclose(ag.b[i], g_actor[n + i].numpy(), rtol=1e-7, atol=1e-11)
    k = 0
    for qi in range(2):
        for arr in qg[qi].W + qg[qi].b + [v for v in qg[qi].ln_bias if v is not None]:
            np.testing.assert_allclose(arr, g_q[k].numpy(), rtol=1e-7, atol=1e-12)
            k += 1


def test_sac_update_step_moves_everything_consistently():
    rng = np.random.default_rng(9)
    D, A, B = 4, 2, 32
    actor = _mk(rng, [D, 8, 8, 2 * A], "silu", False)
    qs = (_mk(rng, [D + A, 8, 8, 1], "silu", True), _mk(rng, [D + A, 8, 8, 1], "silu", True))
    qt = tuple(q.copy() for q in qs)
    batch = {"obs": rng.standard_normal((B, D)), "action": rng.uniform(-1, 1, (B, A)), "reward": rng.standard_normal(B),
             "done": rng.random(B) < 0.2, "next_obs": rng.standard_normal((B, D))}
    noise = {k: rng.standard_normal((B, A)) for k in ("actor", "q", "alpha")}
    h = S.SACHyper(target_entropy=-float(A), tau=0.1)
    z = lambda n: O.AdamState(np.zeros(n), np.zeros(n))
    opt = {"actor": z(actor.flat().size), "q": z(2 * qs[0].flat().size), "alpha": z(1)}
    a2, q2, t2, la2, info = S.sac_update(actor, qs, qt, 0.0, opt, batch, noise, h)
    # first Adam step: |delta| ~ lr for every entry with a non-negligible gradient
    assert np.abs(a2.flat() - actor.flat()).max() <= 3e-4 * 1.0001
    assert abs(la2) <= 3e-4 * 1.0001 and la2 != 0.0
    q_new, q_old, tgt = np.concatenate([q.flat() for q in q2]), np.concatenate([q.flat() for q in qs]), np.concatenate([q.flat() for q in t2])
    np.testing.assert_allclose(tgt, 0.1 * q_new + 0.9 * q_old, rtol=1e-12)   # target started equal to online
    assert opt["q"].count == 1 and opt["actor"].count == 1 and opt["alpha"].count == 1
    h2 = S.SACHyper(target_entropy=-float(A), autotune=False)
    _, _, _, la3, _ = S.sac_update(actor, qs, qt, 0.3, {k: z(v.mu.size) for k, v in opt.items()}, batch, noise, h2)
    assert la3 == 0.3
