"""Anakin feed-forward SAC on B200 -- drop-in for stoix/systems/sac/ff_sac.py.

Same entry points and signatures (get_warmup_fn :40-92, get_learner_fn :95-344, learner_setup :347-513, run_experiment :516,
hydra_entry_point) and the same state / output contracts (sac_types.py), on a different machine:

  reference (JAX)                                      here (B200)
  ---------------------------------------------------  ------------------------------------------------------------------
  flashbax item buffer (add / sample, host-traced)      device ring buffer, write position in HBM (stx_replay_add / _sample):
                                                        ONE launch lays the sampled batch out as the three network inputs
  3 x jax.grad (actor, twin-Q, alpha) per epoch         train-mode MLP kernels (stx_mlp_forward_train / _backward: silu,
                                                        LayerNorm), tanh-Normal head kernels with reparameterised backward,
                                                        three small loss kernels; the actor loss reaches the policy through
                                                        d Q / d action of the twin-Q backward (no parameter gradients there)
  3 x optax.chain(clip, adam) + incremental_update      one fused clip+Adam launch over [actor | q1 q2 | log_alpha], one Polyak
  pmean(batch); pmean(device)                           mean over update-batch shards; one NCCL all-reduce of the gradient arena
  one XLA program per learn() call                      one CUDA graph per update step (rollout + add + epochs), replayed

Every random stream position (policy noise, replay indices) is a device counter, so graph replays draw fresh numbers.
"""
from __future__ import annotations

import copy
import gc
import os
import sys
import time
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from stoix_b200 import ops
from stoix_b200 import optim as optax
from stoix_b200 import random as srandom
from stoix_b200.base_types import AnakinExperimentOutput
from stoix_b200.config import DictConfig, compose, instantiate, to_container
from stoix_b200.envs.base import Environment, TimeStep
from stoix_b200.networks.base import FeedForwardActor as Actor
from stoix_b200.networks.base import FeedForwardQ, MultiNetwork, build_param_tree
from stoix_b200.systems.sac.sac_types import OffPolicyLearnerState, OnlineAndTarget, SACOptStates, SACParams, Transition
from stoix_b200.utils import make_env as environments
from stoix_b200.utils.logger import LogEvent, StoixLogger
from stoix_b200.utils.replay import TransitionBuffer
from stoix_b200.utils.total_timestep_checker import check_total_timesteps
from stoix_b200.utils.training import make_learning_rate

_METRIC_NAMES = ("actor_loss", "entropy", "q_loss", "q_error", "q1_pred", "q2_pred", "alpha_loss", "alpha")
_pad8 = lambda n: (int(n) + 7) // 8 * 8


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def sac_arena_layout(actor: ops.MlpSpec, q: ops.MlpSpec) -> Dict[str, int]:
    """[actor | q1 | q2 | log_alpha], every block padded to 8 floats; the Q optimiser segment is q1..q2 (both networks: ONE global
    norm, like optax over the MultiNetwork tree); the target arena mirrors the [q1 | q2] block."""
    off_q1 = _pad8(actor.param_count)
    off_q2 = off_q1 + _pad8(q.param_count)
    off_alpha = off_q2 + _pad8(q.param_count)
    return {"actor": 0, "q1": off_q1, "q2": off_q2, "alpha": off_alpha, "total": off_alpha + 8, "q_block": off_alpha - off_q1}


class _Shard:
    """Per (device, update-batch) shard: rollout trajectory (Transition fields, time-major), replay buffer, epoch workspaces."""

    def __init__(self, T: int, E: int, B: int, D: int, A: int, sa: ops.MlpSpec, sq: ops.MlpSpec, buffer: TransitionBuffer, device):
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=device)
        self.obs, self.next_obs, self.action = z(T, E, D), z(T, E, D), z(T, E, A)
        self.reward, self.done = z(T, E), z(T, E, dt=torch.uint8)
        self.episode_return, self.episode_length = z(T, E), z(T, E, dt=torch.int32)
        self.is_terminal_step = z(T, E, dt=torch.bool)
        self.cur_obs = z(E, D)                      # last_timestep.observation
        self.buffer = buffer
        # one sampled batch as the three network inputs (obs | action), leading dimension D + A
        self.xq_old, self.xq_new, self.xq_next = z(B, D + A), z(B, D + A), z(B, D + A)
        self.b_reward, self.b_done, self.b_idx = z(B), z(B, dt=torch.uint8), z(B, dt=torch.int32)
        self.ws_actor, self.ws_actor_next = ops.mlp_train_workspace(sa, B, device), ops.mlp_train_workspace(sa, B, device)
        self.ws_q = [ops.mlp_train_workspace(sq, B, device) for _ in range(2)]
        self.ws_qt = [ops.mlp_train_workspace(sq, B, device) for _ in range(2)]
        self.head, self.head_next, self.d_head = z(B, 2 * A), z(B, 2 * A), z(B, 2 * A)
        self.q = [z(B, 1) for _ in range(2)]        # Q_k(obs, .) -- first on the fresh action, then on the stored one
        self.nq = [z(B, 1) for _ in range(2)]       # target Q_k(next_obs, next action)
        self.dq = [z(B, 1) for _ in range(2)]
        self.d_in = [z(B, D + A) for _ in range(2)]

    def transition(self) -> Transition:
        return Transition(self.obs, self.action, self.reward, self.done, self.next_obs,
                          {"episode_return": self.episode_return, "episode_length": self.episode_length, "is_terminal_step": self.is_terminal_step})


def _owner(fn, kind, what):
    obj = getattr(fn, "__self__", None)
    if not isinstance(obj, kind):
        raise TypeError(f"{what} must be a bound method of a stoix_b200 {kind.__name__} (got {fn!r}); arbitrary callables cannot be "
                        "lowered onto the CUDA kernels")
    return obj


def _policy_step(env: Environment, actor_apply_fn, actor_params, env_state, obs: torch.Tensor, seed: int, offset: int, counter):
    """SELECT ACTION + STEP ENVIRONMENT (ff_sac.py:56-75 / 121-140)."""
    action = actor_apply_fn(actor_params, obs).sample(seed=seed, offset=offset, dev_counter=counter)
    env_state, timestep = env.step(env_state, action)
    return env_state, timestep, action


def get_warmup_fn(env: Environment, params: SACParams, actor_apply_fn: Callable, buffer_add_fn: Callable, config: DictConfig) -> Callable:
    """ff_sac.py:40-92: `warmup_steps` environment steps with the initial policy, added to the buffer (one shard per call)."""
    W = int(config.system.warmup_steps)

    def warmup(env_state, timestep: TimeStep, buffer_state, key: int):
        obs_l, act_l, rew_l, done_l, nobs_l = [], [], [], [], []
        for t in range(W):
            obs = timestep.observation.float()
            env_state, timestep, action = _policy_step(env, actor_apply_fn, params.actor_params, env_state, obs, key, t, None)
            obs_l.append(obs.clone()), act_l.append(action), rew_l.append(timestep.reward.float())
            done_l.append(timestep.last().to(torch.uint8)), nobs_l.append(timestep.extras["next_obs"].float())
        if W > 0:
            st = lambda xs: torch.stack(xs).contiguous()
            buffer_add_fn(buffer_state, Transition(st(obs_l), st(act_l), st(rew_l), st(done_l), st(nobs_l), {}))
        return env_state, timestep, key, buffer_state

    return warmup


def get_learner_fn(env: Environment, apply_fns: Tuple[Callable, Callable], update_fns: Tuple[Callable, Callable, Callable],
                   buffer_fns: Tuple[Callable, Callable], config: DictConfig):
    """ff_sac.py:95-344.  apply_fns = (actor_network.apply, double_q_network.apply), update_fns = the three optimisers' `.update`;
    they are read for the shapes / hyper-parameters they are bound to, the arithmetic runs in the kernels named in the header.
    buffer_fns = (TransitionBuffer.add, TransitionBuffer.sample_into) unbound, applied to learner_state.buffer_state[u]."""
    actor_apply_fn, q_apply_fn = apply_fns
    actor_net = _owner(actor_apply_fn, Actor, "apply_fns[0]")
    _owner(q_apply_fn, MultiNetwork, "apply_fns[1]")
    actor_opt, q_opt, alpha_opt = (_owner(f, optax.GradientTransformation, f"update_fns[{i}]") for i, f in enumerate(update_fns))
    buffer_add_fn, buffer_sample_fn = buffer_fns
    rank, world = _world()
    sysc, arch = config.system, config.arch
    T, E, U = int(sysc.rollout_length), int(arch.num_envs), int(arch.update_batch_size)
    epochs, B = int(sysc.epochs), int(sysc.batch_size)
    gamma, tau, autotune = float(sysc.gamma), float(sysc.tau), bool(sysc.autotune)
    target_entropy = float(sysc.target_entropy)
    lo, hi = float(sysc.action_minimum), float(sysc.action_maximum)
    min_scale = float(actor_net.action_head.min_scale)
    use_graph = bool(arch.get("cuda_graph", True))
    built: Dict[str, Any] = {}

    def _build(state: OffPolicyLearnerState) -> None:
        a_tree = state.params.actor_params
        sa, sq = a_tree.spec, state.params.q_params.online[0].spec
        dev = a_tree.flat.device
        lay = sac_arena_layout(sa, sq)
        D, A = sa.sizes[0], sa.sizes[-1] // 2
        segs = [(lay["actor"], sa.param_count, actor_opt.init_lr, actor_opt.max_grad_norm),
                (lay["q1"], lay["q_block"], q_opt.init_lr, q_opt.max_grad_norm)]
        if autotune:
            segs.append((lay["alpha"], 1, alpha_opt.init_lr, alpha_opt.max_grad_norm))
        plan = ops.AdamPlan(segs, dev, b1=actor_opt.adam.b1, b2=actor_opt.adam.b2, eps=actor_opt.adam.eps,
                            decay=bool(sysc.decay_learning_rates), steps_per_update=epochs, num_updates=int(arch.num_updates))
        plan.counts = a_tree.arena_counts[: 2 * len(segs)]   # optimiser counters live in the learner state
        shards = [_Shard(T, E, B, D, A, sa, sq, state.buffer_state[u], dev) for u in range(U)]
        built.update(sa=sa, sq=sq, dev=dev, lay=lay, D=D, A=A, plan=plan, shards=shards, arena=a_tree.arena, target=a_tree.target_arena,
                     grads=torch.zeros(lay["total"], dtype=torch.float32, device=dev),
                     metrics=torch.zeros(epochs, 8, dtype=torch.float32, device=dev),
                     act_ctr=torch.zeros(1, dtype=torch.int64, device=dev),     # policy-noise stream position (rollout)
                     eps_ctr=torch.zeros(1, dtype=torch.int64, device=dev),     # policy-noise stream position (losses)
                     graph=None, eager_done=False)

    def _rollout_phase(state: OffPolicyLearnerState) -> None:
        """_env_step x rollout_length (ff_sac.py:114-146) + buffer add (:149-150), per shard."""
        b = built
        act_seed = state.key[0]
        for u in range(U):
            sh = b["shards"][u]
            for t in range(T):
                sh.obs[t].copy_(sh.cur_obs)
                env_state, ts, action = _policy_step(env, actor_apply_fn, state.params.actor_params, state.env_state[u], sh.obs[t],
                                                     act_seed + 0x9E37 * u, t, b["act_ctr"])
                state.env_state[u] = env_state
                sh.action[t].copy_(action)
                sh.reward[t].copy_(ts.reward)
                sh.done[t].copy_(ts.last())
                sh.next_obs[t].copy_(ts.extras["next_obs"])
                info = ts.extras["episode_metrics"]
                sh.episode_return[t].copy_(info["episode_return"])
                sh.episode_length[t].copy_(info["episode_length"])
                sh.is_terminal_step[t].copy_(info["is_terminal_step"])
                sh.cur_obs.copy_(ts.observation)
            buffer_add_fn(sh.buffer, sh.transition())
        ops.counter_add(b["act_ctr"], T)

    def _epoch_grads(state: OffPolicyLearnerState, u: int, metrics: torch.Tensor, noise: Optional[Dict[str, torch.Tensor]] = None,
                     sample: bool = True, idx_in: Optional[torch.Tensor] = None) -> None:
        """The three losses of one `_update_epoch` on shard u's batch and their gradients, accumulated (weight 1/U) into the
        gradient arena (ff_sac.py:157-226, 229-283).  noise = {"actor", "q", "alpha"}: injected standard normals (tests)."""
        b, sh = built, built["shards"][u]
        sa, sq, lay, D, A = b["sa"], b["sq"], b["lay"], b["D"], b["A"]
        arena, target, grads = b["arena"], b["target"], b["grads"]
        p_actor = arena[lay["actor"]:]
        p_q = [arena[lay["q1"]:], arena[lay["q2"]:]]
        p_qt = [target[0:], target[lay["q2"] - lay["q1"]:]]
        log_alpha = arena[lay["alpha"]: lay["alpha"] + 1]
        w, first = 1.0 / U, u == 0
        eps_seed = state.key[1] + 0x51ED * u
        nz = noise or {}
        if sample:   # SAMPLE TRANSITIONS (:226-227): one launch writes (obs | action), (obs | .), (next_obs | .), reward, done
            buffer_sample_fn(sh.buffer, sh.xq_old, sh.b_reward, sh.b_done, xq_new=sh.xq_new, xq_next=sh.xq_next, idx_in=idx_in, idx_out=sh.b_idx)
        # ---- actor loss (:207-226): mean(alpha * log_prob(a) - min_k Q_k(obs, a)), a = reparameterised sample ----
        ops.mlp_forward_train(sa, p_actor, sh.xq_new, sh.ws_actor, out=sh.head)
        _, logp_new, eps1 = ops.tanh_normal_sample(sh.head, lo, hi, min_scale, eps=nz.get("actor"), seed=eps_seed, offset=0, dev_counter=b["eps_ctr"],
                                                   action_out=sh.xq_new[:, D:])
        for k in range(2):
            ops.mlp_forward_train(sq, p_q[k], sh.xq_new, sh.ws_q[k], out=sh.q[k])
        ops.sac_actor_seed(sh.q[0], sh.q[1], logp_new, log_alpha, sh.dq[0], sh.dq[1], metrics=metrics, weight=w)
        for k in range(2):   # input gradients only: the Q parameters are constants of the actor loss
            ops.mlp_backward(sq, p_q[k], sh.xq_new, sh.dq[k], sh.ws_q[k], net_grad=None, d_input=sh.d_in[k])
        sh.d_in[0].add_(sh.d_in[1])
        ops.tanh_normal_backward(sh.head, eps1, lo, hi, log_alpha, 1.0 / B, sh.d_in[0][:, D:], min_scale, out=sh.d_head)
        ops.mlp_backward(sa, p_actor, sh.xq_new, sh.d_head, sh.ws_actor, net_grad=grads[lay["actor"]:], grad_weight=w, overwrite=first)
        # ---- Q loss (:177-205): 0.5 mean((Q_k(obs, action) - target)^2), target from the TARGET networks on a fresh next action ----
        for k in range(2):
            ops.mlp_forward_train(sq, p_q[k], sh.xq_old, sh.ws_q[k], out=sh.q[k])
        ops.mlp_forward_train(sa, p_actor, sh.xq_next, sh.ws_actor_next, out=sh.head_next)
        _, logp_next, _ = ops.tanh_normal_sample(sh.head_next, lo, hi, min_scale, eps=nz.get("q"), seed=eps_seed, offset=1, dev_counter=b["eps_ctr"],
                                                 action_out=sh.xq_next[:, D:], want_eps=False)
        for k in range(2):
            ops.mlp_forward_train(sq, p_qt[k], sh.xq_next, sh.ws_qt[k], out=sh.nq[k])
        ops.sac_q_loss(sh.q[0], sh.q[1], sh.nq[0], sh.nq[1], logp_next, sh.b_reward, sh.b_done, log_alpha, gamma, sh.dq[0], sh.dq[1],
                       metrics=metrics, weight=w)
        for k, name in enumerate(("q1", "q2")):
            ops.mlp_backward(sq, p_q[k], sh.xq_old, sh.dq[k], sh.ws_q[k], net_grad=grads[lay[name]:], grad_weight=w, overwrite=first)
        # ---- alpha loss (:157-175): mean(alpha * (-log_prob(a') - target_entropy)), a' a second sample of pi(. | obs) ----
        _, logp_alpha, _ = ops.tanh_normal_sample(sh.head, lo, hi, min_scale, eps=nz.get("alpha"), seed=eps_seed, offset=2, dev_counter=b["eps_ctr"],
                                                  want_eps=False)
        ops.sac_alpha_grad(logp_alpha, log_alpha, target_entropy, autotune, grads[lay["alpha"]: lay["alpha"] + 1], grad_weight=w, overwrite=first,
                           metrics=metrics, weight=w)

    def _update_epoch(state: OffPolicyLearnerState, ep: int, noise=None, sample: bool = True, idx_in=None) -> None:
        """One `_update_epoch` (ff_sac.py:219-321): gradients of every shard, mean over shards and devices, the three optimisers in
        one launch (alpha first in the reference: its update uses the OLD alpha in all three losses, as here), Polyak."""
        b = built
        metrics = b["metrics"][ep]
        metrics.zero_()
        for u in range(U):
            _epoch_grads(state, u, metrics, noise=noise, sample=sample, idx_in=idx_in)
        ops.counter_add(b["eps_ctr"], 3)
        if world > 1:
            dist.all_reduce(b["grads"], op=dist.ReduceOp.SUM)
            dist.all_reduce(metrics, op=dist.ReduceOp.SUM)
            metrics.mul_(1.0 / world)
        a_tree = state.params.actor_params
        ops.clip_adam_step(b["plan"], b["arena"], b["grads"], a_tree.arena_mu, a_tree.arena_nu, grad_scale=1.0 / world)
        lay = b["lay"]
        ops.polyak_update(b["target"], b["arena"][lay["q1"]: lay["alpha"]], tau)   # optax.incremental_update (:296-299)

    def _update_step(state: OffPolicyLearnerState) -> None:
        _rollout_phase(state)
        for ep in range(epochs):
            _update_epoch(state, ep)

    def learner_fn(learner_state: OffPolicyLearnerState) -> AnakinExperimentOutput:
        """arch.num_updates_per_eval update steps (ff_sac.py:330-344)."""
        learner_fn.ensure_built(learner_state)
        b = built
        n_upd, dev = int(arch.num_updates_per_eval), b["dev"]
        ep_out = {"episode_return": torch.empty(n_upd, U, T, E, device=dev),
                  "episode_length": torch.empty(n_upd, U, T, E, dtype=torch.int32, device=dev),
                  "is_terminal_step": torch.empty(n_upd, U, T, E, dtype=torch.bool, device=dev)}
        train_out = torch.empty(n_upd, epochs, 8, device=dev)
        for k in range(n_upd):
            if use_graph and b["eager_done"] and b["graph"] is None:
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                gc.collect()   # no cyclic GC (which may destroy an older CUDAGraph) while the stream is capturing
                gc_was_enabled = gc.isenabled()
                gc.disable()
                added = [sh.buffer.added for sh in b["shards"]]   # capturing launches nothing: keep the host mirrors where they are
                try:
                    with torch.cuda.graph(g):
                        _update_step(learner_state)
                finally:
                    if gc_was_enabled:
                        gc.enable()
                for sh, n in zip(b["shards"], added):
                    sh.buffer.added = n
                b["graph"] = g
            if b["graph"] is not None:
                b["graph"].replay()
            else:
                _update_step(learner_state)
                b["eager_done"] = True
            for u in range(U):
                sh = b["shards"][u]
                sh.buffer.added += T * E if b["graph"] is not None else 0   # host mirror of the item count under replay
                ep_out["episode_return"][k, u].copy_(sh.episode_return)
                ep_out["episode_length"][k, u].copy_(sh.episode_length)
                ep_out["is_terminal_step"][k, u].copy_(sh.is_terminal_step)
            train_out[k].copy_(b["metrics"])
        new_ts = [learner_state.timestep[u]._replace(observation=b["shards"][u].cur_obs) for u in range(U)]
        learner_state = learner_state._replace(timestep=new_ts)
        train_metrics = {name: train_out[..., j] for j, name in enumerate(_METRIC_NAMES)}
        return AnakinExperimentOutput(learner_state=learner_state, episode_metrics=ep_out, train_metrics=train_metrics)

    def _ensure_built(st: OffPolicyLearnerState) -> None:
        if not built:
            _build(st)
            for u in range(U):
                built["shards"][u].cur_obs.copy_(st.timestep[u].observation)

    learner_fn.built = built
    learner_fn.ensure_built = _ensure_built
    learner_fn.update_step = _update_step
    learner_fn.update_epoch = _update_epoch
    learner_fn.rollout_phase = _rollout_phase
    return learner_fn


def learner_setup(env: Environment, keys: Tuple[int, int, int], config: DictConfig):
    """Initialise learner_fn, networks, optimisers, buffer, environment and states (ff_sac.py:347-513)."""
    rank, world = _world()
    device = torch.device("cuda", torch.cuda.current_device())
    n_devices = world

    action_space = env.action_space()
    action_dim = int(action_space.shape[-1])
    config.system.action_dim = action_dim
    config.system.action_minimum = float(action_space.minimum)
    config.system.action_maximum = float(action_space.maximum)

    key, actor_net_key, q_net_key = keys

    # Define actor_network, q_network and optimisers (:360-393).
    actor_torso = instantiate(config.network.actor_network.pre_torso)
    actor_action_head = instantiate(config.network.actor_network.action_head, action_dim=action_dim,
                                    minimum=config.system.action_minimum, maximum=config.system.action_maximum)
    actor_network = Actor(torso=actor_torso, action_head=actor_action_head)

    def create_q_network(cfg: DictConfig) -> FeedForwardQ:
        return FeedForwardQ(critic_head=instantiate(cfg.network.q_network.critic_head), torso=instantiate(cfg.network.q_network.pre_torso),
                            input_layer=instantiate(cfg.network.q_network.input_layer))

    double_q_network = MultiNetwork([create_q_network(config), create_q_network(config)])

    actor_lr = make_learning_rate(config.system.actor_lr, config, config.system.epochs)
    q_lr = make_learning_rate(config.system.q_lr, config, config.system.epochs)
    alpha_lr = make_learning_rate(config.system.alpha_lr, config, config.system.epochs)
    actor_optim = optax.chain(optax.clip_by_global_norm(config.system.max_grad_norm), optax.adam(actor_lr, eps=1e-5))
    q_optim = optax.chain(optax.clip_by_global_norm(config.system.max_grad_norm), optax.adam(q_lr, eps=1e-5))
    alpha_optim = optax.chain(optax.clip_by_global_norm(config.system.max_grad_norm), optax.adam(alpha_lr, eps=1e-5))

    init_x = env.observation_space().generate_value()[None, ...].to(device).float()
    init_a = torch.zeros(1, action_dim, device=device)
    init_xa = torch.cat([init_x, init_a], dim=-1)

    # One flat arena [actor | q1 | q2 | log_alpha]; the target Q networks mirror the [q1 | q2] block (:398-426).
    sa = actor_network.spec_for(init_x.shape[-1])
    sq = double_q_network.networks[0].spec_for(init_xa.shape[-1])
    lay = sac_arena_layout(sa, sq)
    arena = torch.zeros(lay["total"], dtype=torch.float32, device=device)
    actor_params = actor_network.init(actor_net_key, init_x, flat=arena[lay["actor"]:])
    online_q_params = double_q_network.init(q_net_key, init_xa, flats=[arena[lay["q1"]:], arena[lay["q2"]:]])
    target_arena = arena[lay["q1"]: lay["alpha"]].clone()
    target_q_params = [build_param_tree(sq, target_arena[off:], "critic_head") for off in (0, lay["q2"] - lay["q1"])]

    # Automatic entropy tuning (:405-412)
    target_entropy = -float(config.system.target_entropy_scale) * action_dim
    log_alpha = arena[lay["alpha"]: lay["alpha"] + 1]
    if not config.system.autotune:
        import math

        log_alpha.fill_(math.log(float(config.system.init_alpha)))
    config.system.target_entropy = target_entropy

    mu, nu = torch.zeros_like(arena), torch.zeros_like(arena)
    counts = torch.zeros(6, dtype=torch.int32, device=device)
    seg = lambda off, n, i: optax.OptState(counts[2 * i: 2 * i + 1], mu[off: off + n], nu[off: off + n], counts[2 * i + 1: 2 * i + 2])
    opt_states = SACOptStates(seg(lay["actor"], sa.param_count, 0), seg(lay["q1"], lay["q_block"], 1), seg(lay["alpha"], 1, 2))
    actor_params.arena, actor_params.arena_mu, actor_params.arena_nu, actor_params.arena_counts = arena, mu, nu, counts
    actor_params.target_arena = target_arena
    params = SACParams(actor_params, OnlineAndTarget(online_q_params, target_q_params), log_alpha)

    apply_fns = (actor_network.apply, double_q_network.apply)
    update_fns = (actor_optim.update, q_optim.update, alpha_optim.update)

    # Replay buffer (:429-456): one ring per (device, update-batch shard)
    U, E = int(config.arch.update_batch_size), int(config.arch.num_envs)
    assert int(config.system.total_buffer_size) % n_devices == 0, "The total buffer size should be divisible by the number of devices!"
    assert int(config.system.total_batch_size) % n_devices == 0, "The total batch size should be divisible by the number of devices!"
    config.system.buffer_size = int(config.system.total_buffer_size) // (n_devices * U)
    config.system.batch_size = int(config.system.total_batch_size) // (n_devices * U)
    shard_keys = srandom.split(key, U + 2)
    buffer_states = [TransitionBuffer(config.system.buffer_size, config.system.batch_size, config.system.batch_size, init_x.shape[-1], action_dim,
                                      device, seed=srandom.split(shard_keys[U], n_devices * U)[rank * U + u]) for u in range(U)]
    buffer_fns = (TransitionBuffer.add, TransitionBuffer.sample_into)

    learn = get_learner_fn(env, apply_fns, update_fns, buffer_fns, config)
    warmup = get_warmup_fn(env, params, actor_network.apply, TransitionBuffer.add, config)

    # Initialise environment states and timesteps across update-batch shards (:463-474).
    env_states: List[Any] = []
    timesteps: List[TimeStep] = []
    for u in range(U):
        if hasattr(env, "seed"):
            env.seed = (int(config.arch.seed) + 7919 * rank + 15485863 * u) & ((1 << 63) - 1)
        st, ts = env.reset(srandom.split(shard_keys[u], E))
        env_states.append(st)
        timesteps.append(ts)

    if config.logger.checkpointing.load_model:   # params only, as in the reference (:477-486)
        from stoix_b200.utils.checkpointing import Checkpointer

        loaded = Checkpointer(model_name=config.system.system_name, **to_container(config.logger.checkpointing.load_args))
        loaded.restore_params(arena)
        target_arena.copy_(arena[lay["q1"]: lay["alpha"]])

    # per-rank streams: (rollout policy noise, loss noise); warm-up noise on its own key (:489-496)
    step_key = srandom.split(shard_keys[U + 1], 3 * n_devices)
    m62 = (1 << 62) - 1
    act_seed, eps_seed, warm_seed = (step_key[3 * rank + i] & m62 for i in range(3))

    # Warm up the buffer (:507-509).
    for u in range(U):
        env_states[u], timesteps[u], _, _ = warmup(env_states[u], timesteps[u], buffer_states[u], warm_seed + 0xA5A5 * u)

    init_learner_state = OffPolicyLearnerState(params, opt_states, buffer_states, (act_seed, eps_seed), env_states, timesteps)
    return learn, actor_network, init_learner_state


def get_final_step_metrics(metrics: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], bool]:
    mask = metrics["is_terminal_step"]
    has_final = bool(mask.any().item())
    out = {k: (v[mask] if has_final else v.reshape(-1)[:0]) for k, v in metrics.items() if k != "is_terminal_step"}
    return out, has_final


def run_experiment(_config: DictConfig) -> float:
    """Runs experiment (ff_sac.py:516-675)."""
    from stoix_b200.evaluator import evaluator_setup, get_distribution_act_fn

    config = copy.deepcopy(_config)
    rank, world = _world()
    n_devices = world
    config.num_devices = n_devices
    config.rank = rank
    config = check_total_timesteps(config, quiet=rank != 0)
    assert config.arch.num_updates >= config.arch.num_evaluation, \
        "Number of updates per evaluation must be less than total number of updates."

    env, eval_env = environments.make(config=config)
    key, key_e, actor_net_key, q_net_key = srandom.split(srandom.PRNGKey(config.arch.seed), num=4)
    learn, actor_network, learner_state = learner_setup(env, (key, actor_net_key, q_net_key), config)

    evaluator, absolute_metric_evaluator = evaluator_setup(
        eval_env=eval_env, key_e=key_e, eval_act_fn=get_distribution_act_fn(config, actor_network.apply), config=config)

    steps_per_rollout = (n_devices * config.arch.num_updates_per_eval * config.system.rollout_length
                         * config.arch.update_batch_size * config.arch.num_envs)
    logger = StoixLogger(config)
    logger.log_config(to_container(config, resolve=True))
    save_checkpoint = config.logger.checkpointing.save_model
    if save_checkpoint:
        from stoix_b200.utils.checkpointing import Checkpointer

        checkpointer = Checkpointer(metadata=to_container(config), model_name=config.system.system_name,
                                    **to_container(config.logger.checkpointing.save_args))

    max_episode_return = -float("inf")
    best_params = learner_state.params.actor_params.flat.clone()
    evaluator_output = None
    eval_step = 0
    for eval_step in range(config.arch.num_evaluation):
        start_time = time.time()
        learner_output = learn(learner_state)
        torch.cuda.synchronize()
        elapsed_time = time.time() - start_time
        t = int(steps_per_rollout * (eval_step + 1))
        episode_metrics, ep_completed = get_final_step_metrics(learner_output.episode_metrics)
        episode_metrics["steps_per_second"] = torch.tensor(steps_per_rollout / elapsed_time)
        logger.log({"timestep": t}, t, eval_step, LogEvent.MISC)
        if ep_completed:
            logger.log(episode_metrics, t, eval_step, LogEvent.ACT)
        train_metrics = dict(learner_output.train_metrics)
        opt_steps_per_eval = config.arch.num_updates_per_eval * config.system.epochs
        train_metrics["steps_per_second"] = torch.tensor(opt_steps_per_eval / elapsed_time)
        logger.log(train_metrics, t, eval_step, LogEvent.TRAIN)

        start_time = time.time()
        trained_params = learner_output.learner_state.params.actor_params
        evaluator_output = evaluator(trained_params, srandom.split(key_e, eval_step + 2)[-1])
        torch.cuda.synchronize()
        elapsed_time = time.time() - start_time
        episode_return = float(evaluator_output["episode_return"].mean().item())
        steps_per_eval = int(evaluator_output["episode_length"].sum().item())
        evaluator_output["steps_per_second"] = torch.tensor(steps_per_eval / max(elapsed_time, 1e-9))
        logger.log(evaluator_output, t, eval_step, LogEvent.EVAL)

        if save_checkpoint and rank == 0:
            checkpointer.save(timestep=t, unreplicated_learner_state=learner_output.learner_state, episode_return=episode_return)
        if config.arch.absolute_metric and max_episode_return <= episode_return:
            best_params = trained_params.flat.clone()
            max_episode_return = episode_return
        learner_state = learner_output.learner_state

    if config.arch.absolute_metric:
        start_time = time.time()
        t = int(steps_per_rollout * (eval_step + 1))
        best_tree = build_param_tree(learner_state.params.actor_params.spec, best_params, "action_head")
        evaluator_output = absolute_metric_evaluator(best_tree, srandom.split(key_e, 1)[0])
        torch.cuda.synchronize()
        elapsed_time = time.time() - start_time
        steps_per_eval = int(evaluator_output["episode_length"].sum().item())
        evaluator_output["steps_per_second"] = torch.tensor(steps_per_eval / max(elapsed_time, 1e-9))
        logger.log(evaluator_output, t, eval_step, LogEvent.ABSOLUTE)

    logger.stop()
    return float(evaluator_output[config.env.eval_metric].float().mean().item())


def hydra_entry_point(cfg: Optional[DictConfig] = None, overrides: Optional[List[str]] = None) -> float:
    """Experiment entry point: `python -m stoix_b200.systems.sac.ff_sac arch.total_num_envs=1024 system.total_batch_size=256 ...`"""
    if cfg is None:
        cfg = compose("default_ff_sac", overrides if overrides is not None else sys.argv[1:], config_dir="default/anakin")
    if "RANK" in os.environ and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    t0 = time.time()
    eval_performance = run_experiment(cfg)
    if _world()[0] == 0:
        print(f"SAC experiment completed in {time.time() - t0:.2f} seconds.")
    return eval_performance


if __name__ == "__main__":
    hydra_entry_point()
