"""Multi-threaded CPU port of the reference's Anakin update step  --  TEST / BASELINE INFRASTRUCTURE.

The reference's own CPU path is JAX/XLA (`JAX_PLATFORMS=cpu`), which cannot be installed here
(SURVEY.md 8c), so `bench.py`'s cpu_baseline / `--impl reference` arm times this restatement instead:
the same arithmetic as oracle/ppo_oracle.py (which is pinned/cross-checked in tests/), written with
torch CPU ops + torch.autograd in float32 so it uses every host core the way XLA:CPU would
(kind = "port").  It follows stoix/systems/ppo/anakin/ff_ppo.py:81-341: T env steps with three MLP
applies each (:98-116), GAE reverse scan (multistep.py:119-130), epochs x minibatches of two
jax.grad-style losses (:191-247) and optax clip+adam (:264-273).  tests/test_torch_port.py checks it
against the NumPy oracle."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch


def mlp(params: List[torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    n = len(params) // 2
    h = x
    for i in range(n):
        h = torch.addmm(params[2 * i + 1], h, params[2 * i])
        if i < n - 1:
            h = torch.relu(h)
    return h


def init_params(sizes, head_scale, gen) -> List[torch.Tensor]:
    out = []
    for i in range(len(sizes) - 1):
        w = torch.empty(sizes[i], sizes[i + 1])
        torch.nn.init.orthogonal_(w, gain=(2.0 ** 0.5 if i < len(sizes) - 2 else head_scale), generator=gen)
        out += [w, torch.zeros(sizes[i + 1])]
    return out


class CpuAnakinPPO:
    def __init__(self, E=4096, T=128, D=64, A=8, hidden=(256, 256), epochs=4, num_minibatches=16, seed=42,
                 gamma=0.99, lam=0.95, clip_eps=0.2, ent_coef=0.01, vf_coef=0.5, max_grad_norm=0.5, lr=3e-4,
                 num_updates=1000, p_term=0.005, p_trunc=0.002):
        self.E, self.T, self.D, self.A = E, T, D, A
        self.epochs, self.nmb = epochs, num_minibatches
        self.gamma, self.lam, self.clip_eps, self.ent_coef, self.vf_coef = gamma, lam, clip_eps, ent_coef, vf_coef
        self.max_grad_norm, self.lr, self.num_updates = max_grad_norm, lr, num_updates
        self.p_term, self.p_trunc = p_term, p_trunc
        self.gen = torch.Generator().manual_seed(seed)
        self.actor = init_params([D, *hidden, A], 0.01, self.gen)
        self.critic = init_params([D, *hidden, 1], 1.0, self.gen)
        self.opt = {id(p): (torch.zeros_like(p), torch.zeros_like(p)) for p in self.actor + self.critic}
        self.count = 0
        self.obs = torch.randn(E, D, generator=self.gen)

    # ---- rollout (ff_ppo.py:81-140) with the synthetic Box env ---------------------------------
    @torch.no_grad()
    def rollout(self) -> Dict[str, torch.Tensor]:
        T, E, D = self.T, self.E, self.D
        tr = {k: torch.empty(T, E) for k in ("value", "reward", "bootstrap", "log_prob")}
        tr["obs"] = torch.empty(T, E, D)
        tr["action"] = torch.empty(T, E, dtype=torch.long)
        tr["done"] = torch.empty(T, E, dtype=torch.bool)
        tr["trunc"] = torch.empty(T, E, dtype=torch.bool)
        obs = self.obs
        for t in range(T):
            logits = mlp(self.actor, obs)
            value = mlp(self.critic, obs)[:, 0]
            gumbel = -torch.log(-torch.log(torch.rand(E, self.A, generator=self.gen).clamp_min(1e-20)))
            action = torch.argmax(logits + gumbel, dim=-1)
            logp = torch.log_softmax(logits, -1).gather(1, action[:, None])[:, 0]
            next_obs = torch.randn(E, D, generator=self.gen)
            reward = torch.randn(E, generator=self.gen)
            term = torch.rand(E, generator=self.gen) < self.p_term
            trunc = (~term) & (torch.rand(E, generator=self.gen) < self.p_trunc)
            bootstrap = mlp(self.critic, next_obs)[:, 0]
            tr["obs"][t], tr["action"][t], tr["value"][t], tr["reward"][t] = obs, action, value, reward
            tr["bootstrap"][t], tr["log_prob"][t], tr["done"][t], tr["trunc"][t] = bootstrap, logp, term, trunc
            last = term | trunc
            obs = torch.where(last[:, None], torch.randn(E, D, generator=self.gen), next_obs)
        self.obs = obs
        return tr

    # ---- GAE (multistep.py:116-139) -------------------------------------------------------------
    @staticmethod
    @torch.no_grad()
    def gae(reward, value, bootstrap, done, trunc, gamma, lam, standardize=True) -> Tuple[torch.Tensor, torch.Tensor]:
        disc = (1.0 - done.float()) * gamma
        delta = reward + disc * bootstrap - value
        nt = 1.0 - trunc.float()
        adv = torch.empty_like(delta)
        acc = torch.zeros(reward.shape[1])
        for t in range(reward.shape[0] - 1, -1, -1):
            acc = delta[t] + disc[t] * lam * acc * nt[t]
            adv[t] = acc
        targets = value + adv
        if standardize:
            mean = adv.mean()
            adv = (adv - mean) * torch.rsqrt((adv * adv).mean() - mean * mean + 1e-5)
        return adv, targets

    # ---- one optimiser (optax clip_by_global_norm + adam eps=1e-5, A.5) -------------------------
    @torch.no_grad()
    def _apply(self, params, grads, lr, c):
        gn = torch.sqrt(sum((g * g).sum() for g in grads))
        scale = 1.0 if gn < self.max_grad_norm else self.max_grad_norm / gn
        for p, g in zip(params, grads):
            mu, nu = self.opt[id(p)]
            g = g * scale
            mu.mul_(0.9).add_(g, alpha=0.1)
            nu.mul_(0.999).addcmul_(g, g, value=0.001)
            p.sub_(lr * (mu / (1 - 0.9 ** c)) / (torch.sqrt(nu / (1 - 0.999 ** c)) + 1e-5))

    def update(self, tr, perms=None) -> Dict[str, float]:
        T, E = self.T, self.E
        B = T * E
        mb = B // self.nmb
        adv, tgt = self.gae(tr["reward"], tr["value"], tr["bootstrap"], tr["done"], tr["trunc"], self.gamma, self.lam)
        f = lambda x: x.reshape((B,) + x.shape[2:])
        obs, act, lpo, vo, adv, tgt = f(tr["obs"]), f(tr["action"]), f(tr["log_prob"]), f(tr["value"]), f(adv), f(tgt)
        info = {}
        for ep in range(self.epochs):
            perm = torch.randperm(B, generator=self.gen) if perms is None else torch.as_tensor(perms[ep], dtype=torch.long)
            for i in range(self.nmb):
                idx = perm[i * mb:(i + 1) * mb]
                x = obs[idx]
                ap = [p.detach().requires_grad_(True) for p in self.actor]
                lp_all = torch.log_softmax(mlp(ap, x), -1)
                logp = lp_all.gather(1, act[idx][:, None])[:, 0]
                ratio = torch.exp(logp - lpo[idx])
                a = adv[idx]
                loss_actor = -torch.minimum(ratio * a, torch.clamp(ratio, 1 - self.clip_eps, 1 + self.clip_eps) * a).mean()
                entropy = -(lp_all.exp() * lp_all).sum(-1).mean()
                ag = torch.autograd.grad(loss_actor - self.ent_coef * entropy, ap)
                cp = [p.detach().requires_grad_(True) for p in self.critic]
                v = mlp(cp, x)[:, 0]
                vclip = vo[idx] + (v - vo[idx]).clamp(-self.clip_eps, self.clip_eps)
                vloss = 0.5 * torch.maximum((v - tgt[idx]) ** 2, (vclip - tgt[idx]) ** 2).mean()
                cg = torch.autograd.grad(self.vf_coef * vloss, cp)
                lr = self.lr * (1.0 - (self.count // (self.epochs * self.nmb)) / self.num_updates)
                self.count += 1
                self._apply(self.actor, ag, lr, self.count)
                self._apply(self.critic, cg, lr, self.count)
                info = {"actor_loss": float(loss_actor), "entropy": float(entropy), "value_loss": float(vloss)}
        return info

    def step(self) -> Dict[str, float]:
        return self.update(self.rollout())
