"""CPU oracle for the ff_sac update step  --  TEST INFRASTRUCTURE, NOT PRODUCT.

NumPy (fp64) restatement of the arithmetic of stoix/systems/sac/ff_sac.py:149-321 (one `_update_epoch`: actor, Q and
alpha losses with their gradients, three clip+Adam updates, Polyak target update) on INJECTED batches and noise, plus the
distribution head it relies on:

* NormalAffineTanhDistributionHead                       stoix/networks/heads.py:44-65
* AffineTanhTransformedDistribution (log_prob with the   stoix/networks/distributions.py:19-79
  clipped tails: log cdf / log survival minus log eps)
* twin Q networks = MultiNetwork of two                  stoix/networks/base.py:104-121, inputs.py:26-33
  CompositeNetwork(EmbeddingActionInput, MLPTorso, ScalarCriticHead)

**Parity unpinned**: the reference has no test for SAC and JAX / tfp / distrax are not installable here.  The hand-derived
gradients below are checked against torch.autograd in tests/test_oracle_sac.py; the tfp pieces (Normal.log_prob / log_cdf /
log_survival_function, bijectors Tanh: forward_log_det_jacobian(x) = 2 (log 2 - x - softplus(-2x)), Scale, Shift) are
restated from their published definitions (tensorflow-probability 0.25.0)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
from scipy.special import erfc, log_ndtr

from oracle import ppo_oracle as O

LOG_SQRT_2PI = 0.5 * np.log(2.0 * np.pi)


def softplus(x):
    return np.logaddexp(x, 0.0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


@dataclass
class TanhNormal:
    """What one evaluation of the head leaves behind (per row, per action dim)."""

    loc: np.ndarray
    raw: np.ndarray      # output of the scale Dense, sigma = softplus(raw) + min_scale
    sigma: np.ndarray
    eps: np.ndarray      # the standard-normal noise of the sample
    u: np.ndarray        # pre-tanh sample loc + sigma * eps
    action: np.ndarray   # shift + scale * tanh(u)
    log_prob: np.ndarray  # (rows,)
    branch: np.ndarray   # 0 inside the clip band, -1 left tail, +1 right tail (per dim)


def head_sample(head_out: np.ndarray, eps: np.ndarray, minimum: float, maximum: float, min_scale: float = 1e-3,
                epsilon: float = 1e-3) -> TanhNormal:
    """heads.py:54-65 + distributions.py:66-75: sample with the given N(0,1) noise and its log_prob."""
    A = head_out.shape[-1] // 2
    loc, raw = head_out[..., :A], head_out[..., A:]
    sigma = softplus(raw) + min_scale
    u = loc + sigma * eps
    s, sh = (maximum - minimum) / 2.0, (minimum + maximum) / 2.0
    action = sh + s * np.tanh(u)
    lo, hi = minimum + epsilon, maximum - epsilon
    ev = np.clip(action, lo, hi)
    u_lo, u_hi = np.arctanh((lo - sh) / s), np.arctanh((hi - sh) / s)
    uc = np.arctanh(np.clip((ev - sh) / s, -1 + 1e-15, 1 - 1e-15))      # bijector.inverse(event)
    z = (uc - loc) / sigma
    inner = -0.5 * z * z - np.log(sigma) - LOG_SQRT_2PI - 2.0 * (np.log(2.0) - uc - softplus(-2.0 * uc)) - np.log(s)
    left = log_ndtr((u_lo - loc) / sigma) - np.log(epsilon)              # distribution.log_cdf(...) - log eps
    right = log_ndtr(-(u_hi - loc) / sigma) - np.log(epsilon)            # log_survival_function(...) - log eps
    branch = np.where(ev <= lo, -1, np.where(ev >= hi, 1, 0))
    per_dim = np.where(branch < 0, left, np.where(branch > 0, right, inner))
    return TanhNormal(loc, raw, sigma, eps, u, action, per_dim.sum(-1), branch)


def head_log_prob_grad(t: TanhNormal, g_logp: np.ndarray, g_action: np.ndarray, minimum: float, maximum: float,
                       epsilon: float = 1e-3) -> np.ndarray:
    """d(loss)/d(head output) for loss = sum_rows g_logp * log_prob + sum g_action * action with the reparameterised sample
    (u = loc + sigma * eps, eps held fixed) -- what jax.grad sees through `actor_policy.sample` + `.log_prob` (ff_sac.py:216-222).
    Inside the band: log_prob_j = -eps^2/2 - log sigma - c - fldj(u) - log s, d(-fldj)/du = 2 tanh(u); in a clipped tail the
    event is a constant and only log cdf / log sf of (loc, sigma) remain."""
    s = (maximum - minimum) / 2.0
    sh = (minimum + maximum) / 2.0
    lo, hi = minimum + epsilon, maximum - epsilon
    u_lo, u_hi = np.arctanh((lo - sh) / s), np.arctanh((hi - sh) / s)
    th = np.tanh(t.u)
    gl = g_logp[..., None]
    # action path (always: the Q network sees the unclipped action)
    d_u = g_action * s * (1.0 - th * th)
    d_loc = d_u.copy()
    d_sigma = d_u * t.eps
    # log-prob path
    band = t.branch == 0
    d_loc += np.where(band, gl * 2.0 * th, 0.0)
    d_sigma += np.where(band, gl * (2.0 * th * t.eps - 1.0 / t.sigma), 0.0)
    for sign, u_thr, mask in ((1.0, u_lo, t.branch < 0), (-1.0, u_hi, t.branch > 0)):
        z = sign * (u_thr - t.loc) / t.sigma                  # left: log Phi(z); right: log Phi(-z') with z' = (u_hi - loc)/sigma
        ratio = np.exp(-0.5 * z * z - LOG_SQRT_2PI - log_ndtr(z))   # phi(z) / Phi(z)
        dz_dloc, dz_dsigma = -sign / t.sigma, -z / t.sigma
        d_loc += np.where(mask, gl * ratio * dz_dloc, 0.0)
        d_sigma += np.where(mask, gl * ratio * dz_dsigma, 0.0)
    d_raw = d_sigma * sigmoid(t.raw)
    return np.concatenate([d_loc, d_raw], axis=-1)


def q_forward(qs: Tuple[O.MLPParams, O.MLPParams], obs: np.ndarray, action: np.ndarray):
    """MultiNetwork([Q, Q]) on concat(obs, action): (rows, 2) values + the caches."""
    x = np.concatenate([obs, action], axis=-1)
    outs, caches = [], []
    for q in qs:
        o, c = O.mlp_forward(q, x)
        outs.append(o[:, 0])
        caches.append(c)
    return np.stack(outs, -1), caches, x


def q_input_grad(q: O.MLPParams, cache, dout: np.ndarray) -> np.ndarray:
    """d(sum dout * q)/d(input) of one Q network (the action columns feed the actor loss)."""
    n = len(q.W)
    _, fp = O.ACTIVATIONS[q.activation]
    d = dout
    for i in range(n - 1, -1, -1):
        dh = d @ q.W[i].T
        if i == 0:
            return dh
        j = i - 1
        if q.has_ln(j):
            u, (mean, rstd) = cache.pre[j], cache.stats[j]
            uh = (u - mean) * rstd
            dz = dh * fp(uh * q.b[j] + q.ln_bias[j])
            dhat = dz * q.b[j]
            d = rstd * (dhat - dhat.mean(-1, keepdims=True) - uh * (dhat * uh).mean(-1, keepdims=True))
        else:
            d = dh * fp(cache.pre[j])
    raise AssertionError


@dataclass
class SACHyper:
    gamma: float = 0.99
    tau: float = 0.005
    max_grad_norm: float = 0.5
    actor_lr: float = 3e-4
    q_lr: float = 3e-4
    alpha_lr: float = 3e-4
    autotune: bool = True
    target_entropy: float = -6.0
    minimum: float = -1.0
    maximum: float = 1.0


def sac_losses_and_grads(actor: O.MLPParams, q_online, q_target, log_alpha: float, batch: Dict[str, np.ndarray],
                         noise: Dict[str, np.ndarray], h: SACHyper):
    """ff_sac.py:149-222, 235-283: the three loss functions and their gradients on one sampled batch.
    noise: {"actor", "q", "alpha"}: (rows, A) standard normals for the three `sample(seed=...)` calls."""
    obs, act, rew, done, nobs = batch["obs"], batch["action"], batch["reward"], batch["done"].astype(np.float64), batch["next_obs"]
    B = obs.shape[0]
    alpha = np.exp(log_alpha)
    # ---- actor loss (:207-226): mean(alpha * log_prob - min_k Q_k(obs, a)) ----
    ho, a_cache = O.mlp_forward(actor, obs)
    ta = head_sample(ho, noise["actor"], h.minimum, h.maximum)
    qa, q_caches, _ = q_forward(q_online, obs, ta.action)
    kmin = np.argmin(qa, axis=-1)
    min_q = qa[np.arange(B), kmin]
    actor_loss = alpha * ta.log_prob - min_q
    g_action = np.zeros_like(ta.action)
    D = obs.shape[1]
    for k in range(2):
        dq = np.where(kmin == k, -1.0 / B, 0.0)[:, None]
        g_action += q_input_grad(q_online[k], q_caches[k], dq)[:, D:]
    d_head = head_log_prob_grad(ta, np.full(B, alpha / B), g_action, h.minimum, h.maximum)
    actor_grads = O.mlp_backward(actor, a_cache, d_head)
    # ---- Q loss (:177-205) ----
    q_old, qo_caches, _ = q_forward(q_online, obs, act)
    hn, _ = O.mlp_forward(actor, nobs)
    tn = head_sample(hn, noise["q"], h.minimum, h.maximum)
    next_q, _, _ = q_forward(q_target, nobs, tn.action)
    next_v = next_q.min(-1) - alpha * tn.log_prob
    target_q = rew + (1.0 - done) * h.gamma * next_v
    q_error = q_old - target_q[:, None]
    q_loss = 0.5 * np.mean(q_error ** 2)
    q_grads = [O.mlp_backward(q_online[k], qo_caches[k], (q_error[:, k] / (2.0 * B))[:, None]) for k in range(2)]
    # ---- alpha loss (:157-175) ----
    tal = head_sample(ho, noise["alpha"], h.minimum, h.maximum)
    alpha_loss = np.mean(alpha * (-tal.log_prob - h.target_entropy))
    alpha_grad = alpha_loss  # d/d log_alpha of exp(log_alpha) * c = the loss itself
    info = {"actor_loss": actor_loss.mean(), "entropy": (-ta.log_prob).mean(), "q_loss": q_loss, "q_error": np.abs(q_error).mean(),
            "q1_pred": next_q[:, 0].mean(), "q2_pred": next_q[:, 1].mean(), "alpha_loss": alpha_loss, "alpha": alpha}
    return actor_grads, q_grads, alpha_grad, info, {"actor_sample": ta, "next_sample": tn, "q_old": q_old, "target_q": target_q}


def sac_update(actor, q_online, q_target, log_alpha, opt, batch, noise, h: SACHyper):
    """One `_update_epoch` (ff_sac.py:227-321) after sampling: returns the new (actor, q_online, q_target, log_alpha) and info.
    opt = {"actor", "q", "alpha"}: oracle AdamState objects (the Q optimiser covers BOTH Q networks: one global norm)."""
    ag, qg, alg, info, _ = sac_losses_and_grads(actor, q_online, q_target, log_alpha, batch, noise, h)
    mk = lambda p, flat: O.MLPParams.from_flat(flat, [p.W[0].shape[0]] + [w.shape[1] for w in p.W], p.activation, p.ln_bias is not None)
    new_log_alpha = log_alpha
    if h.autotune:
        na, _ = O.clip_adam_step(np.array([log_alpha], np.float64), np.array([alg], np.float64), opt["alpha"], h.alpha_lr, h.max_grad_norm)
        new_log_alpha = float(na[0])
    a_new, _ = O.clip_adam_step(actor.flat(), ag.flat(), opt["actor"], h.actor_lr, h.max_grad_norm)
    q_flat = np.concatenate([q.flat() for q in q_online])
    qg_flat = np.concatenate([g.flat() for g in qg])
    q_new, _ = O.clip_adam_step(q_flat, qg_flat, opt["q"], h.q_lr, h.max_grad_norm)
    n0 = q_online[0].flat().size
    q_new_nets = (mk(q_online[0], q_new[:n0]), mk(q_online[1], q_new[n0:]))
    # optax.incremental_update(new, old, tau) = tau * new + (1 - tau) * old
    t_flat = h.tau * q_new + (1.0 - h.tau) * np.concatenate([q.flat() for q in q_target])
    q_tgt_nets = (mk(q_target[0], t_flat[:n0]), mk(q_target[1], t_flat[n0:]))
    return mk(actor, a_new), q_new_nets, q_tgt_nets, new_log_alpha, info
