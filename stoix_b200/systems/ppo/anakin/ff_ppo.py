"""Anakin feed-forward PPO on B200 -- drop-in for stoix/systems/ppo/anakin/ff_ppo.py.

Same entry points and signatures (get_learner_fn :49-54, learner_setup :424-426, run_experiment :554,
hydra_entry_point :709-727) and the same state / output contracts (stoix_b200/base_types.py), but a
different machine underneath:

  reference (JAX)                                 here (B200)
  ----------------------------------------------  -------------------------------------------------
  pmap over devices                               one process per GPU (torchrun), rank = device index
  lax.scan over T env steps, 3 MLP applies/step   T x (actor MLP kernel + sample kernel + env kernel);
                                                  the two critic evaluations are batched over the
                                                  whole (T*E) rollout afterwards (same params, same
                                                  rows => same values, 2 launches instead of 2T)
  lax.scan reverse over T for GAE                 one chunked parallel-scan launch (K2)
  permutation + take (shuffled copy per epoch)    keyed bijection -> index vector; kernels gather
  jax.grad of two loss fns                        fused forward/loss/backward launches (K3)
  pmean(batch) ; pmean(device)                    mean over local shards ; NCCL all-reduce of one flat
                                                  gradient arena
  2 x optax.chain(clip, adam) over 12 leaves      one fused clip+Adam launch over the arena (K4)
  one XLA program per learn() call                one CUDA graph per update step, replayed

The per-update work is captured into a CUDA graph after the first (eager) update; every RNG stream
position lives in device memory so replays draw fresh numbers.
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
from stoix_b200.base_types import (
    ActorApply,
    ActorCriticOptStates,
    ActorCriticParams,
    AnakinExperimentOutput,
    CriticApply,
    LearnerFn,
    OnPolicyLearnerState,
)
from stoix_b200.config import DictConfig, compose, instantiate, to_container
from stoix_b200.envs.base import Environment, StepOut, TimeStep
from stoix_b200.networks.base import FeedForwardActor as Actor
from stoix_b200.networks.base import FeedForwardCritic as Critic
from stoix_b200.networks.base import build_param_tree
from stoix_b200.systems.ppo.ppo_types import PPOTransition
from stoix_b200.utils import make_env as environments
from stoix_b200.utils.logger import LogEvent, StoixLogger
from stoix_b200.utils.total_timestep_checker import check_total_timesteps
from stoix_b200.utils.training import make_learning_rate

_METRIC_NAMES = ("actor_loss", "entropy", "value_loss", "advantages", "pred_value", "target_value")


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class _Shard:
    """Trajectory buffers of one (device, update-batch) shard: PPOTransition fields, time-major."""

    def __init__(self, T: int, E: int, D: int, A: int, obs_dtype, device, normalize: bool = False):
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=device)
        self.obs = z(T + 1, E, D, dt=obs_dtype)      # row t = last_timestep.observation of step t
        self.next_obs = z(T, E, D, dt=obs_dtype)     # timestep.extras["next_obs"]   (ff_ppo.py:113)
        # normalize_observations: the environment writes RAW fp32 observations here (they feed the running statistics,
        # ff_ppo.py:145-162) and obs / next_obs above hold the normalised copies every network kernel reads (:90-94,113-115)
        self.obs_raw = z(T + 1, E, D) if normalize else None
        self.next_obs_raw = z(T, E, D) if normalize else None
        self.action = z(T, E, dt=torch.int32)
        self.log_prob = z(T, E)
        self.value = z(T, E)
        self.bootstrap_value = z(T, E)
        self.reward = z(T, E)
        self.done = z(T, E, dt=torch.uint8)
        self.truncated = z(T, E, dt=torch.uint8)
        self.episode_return = z(T, E)
        self.episode_length = z(T, E, dt=torch.int32)
        self.is_terminal_step = z(T, E, dt=torch.uint8)
        self.advantages = z(T, E)
        self.targets = z(T, E)
        self.logits = z(E, A)
        self.perms = None  # (epochs, T*E) int32 shuffles of this update, allocated by the learner
        self.adv_stats: Optional[torch.Tensor] = None

    def step_out(self, t: int) -> StepOut:
        obs, nxt = (self.obs, self.next_obs) if self.obs_raw is None else (self.obs_raw, self.next_obs_raw)
        return StepOut(obs[t + 1], nxt[t], self.reward[t], self.done[t], self.truncated[t],
                       self.episode_return[t], self.episode_length[t], self.is_terminal_step[t])

    @property
    def env_obs(self) -> torch.Tensor:
        """(T+1, E, D) observations as the environment produced them (raw when normalising)."""
        return self.obs if self.obs_raw is None else self.obs_raw

    def transition(self, T: int) -> PPOTransition:
        info = {"episode_return": self.episode_return, "episode_length": self.episode_length,
                "is_terminal_step": self.is_terminal_step.bool()}
        return PPOTransition(self.done.bool(), self.truncated.bool(), self.action, self.value, self.reward,
                             self.bootstrap_value, self.log_prob, self.obs[:T], info)


def get_learner_fn(
    env: Environment,
    apply_fns: Tuple[ActorApply, CriticApply],
    update_fns: Tuple[Callable, Callable],
    config: DictConfig,
) -> LearnerFn[OnPolicyLearnerState]:
    """Get the learner function (ff_ppo.py:49-372).

    apply_fns: (actor_network.apply, critic_network.apply) bound methods; update_fns:
    (actor_optim.update, critic_optim.update) handles from stoix_b200.optim -- both are used for the
    network shapes / optimiser hyper-parameters they carry; the arithmetic runs in the fused kernels.
    """
    actor_apply_fn, critic_apply_fn = apply_fns
    # The arithmetic of apply / update runs inside the fused kernels; what the learner needs from the four callables is the
    # description they are bound to (layer sizes, optimiser hyper-parameters).  Anything else cannot be honoured silently.
    def _owner(fn, kind, what):
        obj = getattr(fn, "__self__", None)
        if not isinstance(obj, kind):
            raise TypeError(f"get_learner_fn: {what} must be the bound `{'apply' if kind is not optax.GradientTransformation else 'update'}` of a "
                            f"stoix_b200 {kind.__name__} (got {fn!r}); arbitrary callables cannot be lowered onto the CUDA kernels")
        return obj

    actor_net, critic_net = _owner(actor_apply_fn, Actor, "apply_fns[0]"), _owner(critic_apply_fn, Critic, "apply_fns[1]")
    actor_opt = _owner(update_fns[0], optax.GradientTransformation, "update_fns[0]")
    critic_opt = _owner(update_fns[1], optax.GradientTransformation, "update_fns[1]")
    rank, world = _world()

    sysc, arch = config.system, config.arch
    T, E, U = int(sysc.rollout_length), int(arch.num_envs), int(arch.update_batch_size)
    epochs, nmb = int(sysc.epochs), int(sysc.num_minibatches)
    B = T * E
    assert B % nmb == 0, "rollout_length * num_envs must be divisible by num_minibatches"
    mb = B // nmb
    precision = ops.STX_PREC_BF16 if str(arch.get("precision", "f32")) == "bf16" else ops.STX_PREC_F32
    use_graph = bool(arch.get("cuda_graph", True))
    has_step_into = hasattr(env, "step_into")

    built: Dict[str, Any] = {}

    def _build(state: OnPolicyLearnerState) -> None:
        a_tree, c_tree = state.params.actor_params, state.params.critic_params
        sa, sc = a_tree.spec, c_tree.spec
        dev = a_tree.flat.device
        arena = state.params.actor_params.arena
        _, coff, total = ops.arena_offsets(sa, sc)
        D, A = sa.sizes[0], sa.sizes[-1]
        obs_dtype = torch.bfloat16 if precision == ops.STX_PREC_BF16 else torch.float32
        rs = getattr(state, "running_statistics", None)
        shards = [_Shard(T, E, D, A, obs_dtype, dev, normalize=rs is not None) for _ in range(U)]
        decay = bool(sysc.decay_learning_rates)
        plan = ops.AdamPlan(
            [(0, sa.param_count, actor_opt.init_lr, actor_opt.max_grad_norm),
             (coff, sc.param_count, critic_opt.init_lr, critic_opt.max_grad_norm)],
            dev, b1=actor_opt.adam.b1, b2=actor_opt.adam.b2, eps=actor_opt.adam.eps, decay=decay,
            steps_per_update=epochs * nmb, num_updates=int(arch.num_updates),
        )
        # optimiser counters live in the learner state (so a checkpointed state carries them)
        plan.counts = a_tree.arena_counts
        peers = None
        if world > 1 and bool(arch.get("fused_allreduce", True)) and (epochs * nmb) % 2 == 0 and world <= 8:
            try:  # gradient arenas in NVLink peer memory: the all-reduce is fused into the optimiser kernel
                peers = ops.PeerGradBuffers(total, dev)
            except Exception as e:  # symmetric memory unavailable (no P2P): NCCL all-reduce + K4
                if rank == 0:
                    print(f"[stoix_b200] fused all-reduce unavailable ({type(e).__name__}: {e}); using NCCL all-reduce")
                peers = None
        built.update(
            sa=sa, sc=sc, dev=dev, arena=arena, coff=coff, total=total, shards=shards, plan=plan,
            grads=torch.zeros(total, dtype=torch.float32, device=dev),
            metrics=torch.zeros(epochs, nmb, 8, dtype=torch.float32, device=dev),
            ws=ops.ppo_workspace(sa, sc, mb, precision, dev),
            arena_bf16=getattr(state.params.actor_params, "arena_bf16", None),
            peers_obj=peers,
            peers=None,
            roll_ctr=torch.zeros(1, dtype=torch.int64, device=dev),   # categorical call index
            perm_ctr=torch.zeros(1, dtype=torch.int64, device=dev),   # shuffle stream index
            side_stream=torch.cuda.Stream(device=dev),                # the shuffles run here, underneath the rollout
            critic_stream=torch.cuda.Stream(device=dev),              # chunked critic evaluation underneath the fused rollout
            rs=rs, rs_sums=torch.zeros(2 * D + 1, dtype=torch.float64, device=dev) if rs is not None else None,
            graph=None, eager_done=False,
        )

    def _env_step(state: OnPolicyLearnerState, u: int, t: int, seeds: Tuple[int, int]) -> None:
        """One vectorised env step of shard u (ff_ppo.py:81-135) writing the transition in place."""
        b = built
        sh: _Shard = b["shards"][u]
        a_tree = state.params.actor_params
        if b["rs"] is not None:  # normalise the observation with the PRE-update statistics (ff_ppo.py:90-94)
            ops.obs_normalize(sh.obs_raw[t], b["rs"].mean, b["rs"].std, out=sh.obs[t])
        # SELECT ACTION (ff_ppo.py:97-101): logits -> sample -> log_prob
        ops.mlp_forward(b["sa"], a_tree.flat, sh.obs[t], precision=precision, params_bf16=a_tree.flat_bf16, out=sh.logits, ws_key="learner")
        ops.categorical(sh.logits, None, seeds[0] + u, t, b["roll_ctr"], out=(sh.action[t], sh.log_prob[t]))
        # STEP ENVIRONMENT (ff_ppo.py:104-116)
        if has_step_into:
            env.step_into(state.env_state[u], sh.action[t], sh.step_out(t), t)
        else:
            new_state, ts = env.step(state.env_state[u], sh.action[t])
            state.env_state[u] = new_state
            out = sh.step_out(t)
            out.obs.copy_(ts.observation)
            out.next_obs.copy_(ts.extras["next_obs"])
            out.reward.copy_(ts.reward)
            out.done.copy_(ts.discount == 0.0)                           # ff_ppo.py:107
            out.truncated.copy_(ts.last() & (ts.discount != 0.0))        # ff_ppo.py:108
            em = ts.extras["episode_metrics"]
            out.episode_return.copy_(em["episode_return"])
            out.episode_length.copy_(em["episode_length"])
            out.is_terminal_step.copy_(em["is_terminal_step"])

    def _rollout_phase(state: OnPolicyLearnerState) -> None:
        """STEP ENVIRONMENT FOR ROLLOUT LENGTH (ff_ppo.py:138-140) + the batched critic evaluations."""
        b = built
        sa, sc = b["sa"], b["sc"]
        c_tree = state.params.critic_params
        D = sa.sizes[0]
        a_tree = state.params.actor_params
        fused = (precision == ops.STX_PREC_BF16 and bool(arch.get("fused_rollout", True)) and E % 128 == 0
                 and getattr(env, "fused_rollout_supported", False) and b["rs"] is None)
        # The two critic evaluations of _env_step (value = critic(obs_t), bootstrap_value = critic(next_obs_t), ff_ppo.py:99,
        # 113-116) depend only on the observations, so they are batched over (steps x envs) rows.  With the fused rollout
        # (E/128 of the 148 SMs busy) the scan is cut into chunks and the critic of chunk c runs on a second stream on the
        # idle SMs while the rollout kernel produces chunk c+1; only the last chunk's critic is left after the scan.
        chunks = max(int(arch.get("rollout_chunks", 8)), 1) if fused else 1
        while T % chunks != 0:  # largest divisor of T not above the request
            chunks -= 1
        Tc = T // chunks
        main = torch.cuda.current_stream()
        side = b["critic_stream"] if chunks > 1 else None

        def critic_rows(sh: _Shard, t_lo: int, t_hi: int) -> None:
            n = (t_hi - t_lo) * E
            ops.mlp_forward(sc, c_tree.flat, sh.obs[t_lo:t_hi].view(n, D), precision=precision, params_bf16=c_tree.flat_bf16,
                            out=sh.value[t_lo:t_hi].view(n, 1), ws_key="learner-critic")
            ops.mlp_forward(sc, c_tree.flat, sh.next_obs[t_lo:t_hi].view(n, D), precision=precision, params_bf16=c_tree.flat_bf16,
                            out=sh.bootstrap_value[t_lo:t_hi].view(n, 1), ws_key="learner-critic")

        for u in range(U):
            sh: _Shard = b["shards"][u]
            if fused:  # persistent launches (envs whose dynamics ignore the action), `chunks` of Tc steps each
                for c in range(chunks):
                    env.fused_rollout(state.env_state[u], sa, a_tree.flat, a_tree.flat_bf16, sh, T, state.key[0] + u, b["roll_ctr"],
                                      t0=c * Tc, steps=Tc)
                    if side is not None:
                        side.wait_stream(main)
                        with torch.cuda.stream(side):
                            # leave the rollout's SMs alone: a persistent forward CTA that has to wait for one of them
                            # would hold its share of the tiles back until the rollout chunk ends
                            ops.set_forward_cta_budget(ops.NUM_SMS - E // 128 if c + 1 < chunks else 0)
                            critic_rows(sh, c * Tc, (c + 1) * Tc)
                            ops.set_forward_cta_budget(0)
                if side is not None:
                    main.wait_stream(side)
                else:
                    critic_rows(sh, 0, T)
            else:
                for t in range(T):
                    _env_step(state, u, t, state.key)
                if b["rs"] is not None:  # bootstrap observations, same pre-update statistics (ff_ppo.py:113-115), one launch
                    ops.obs_normalize(sh.next_obs_raw, b["rs"].mean, b["rs"].std, out=sh.next_obs)
                critic_rows(sh, 0, T)

    def _stats_phase(state: OnPolicyLearnerState) -> None:
        """UPDATE RUNNING STATISTICS (ff_ppo.py:145-162) with the raw trajectory observations of every shard ("batch")
        and rank ("device"), std limits 5e-4 / 5e4.  Runs after the rollout: everything in this update that is
        normalised (rollout inputs, the minibatch observations of the epochs) used the statistics from before it."""
        b = built
        if b["rs"] is None:
            return
        from stoix_b200.utils import running_statistics as rstat

        rstat.update_statistics_(b["rs"], [sh.obs_raw[:T] for sh in b["shards"]], std_min_value=5e-4, std_max_value=5e4,
                                 pmap_axes=["device", "batch"], validate_shapes=False, sums=b["rs_sums"])

    def _gae_phase(state: OnPolicyLearnerState) -> None:
        """CALCULATE ADVANTAGE (ff_ppo.py:164-179)."""
        for u in range(U):
            sh: _Shard = built["shards"][u]
            _, _, sh.adv_stats = ops.gae_ppo(
                sh.reward, sh.value, sh.bootstrap_value, sh.done, sh.truncated, float(sysc.gamma),
                float(sysc.gae_lambda), float(sysc.reward_scale), 1 if sysc.standardize_advantages else 0,
                out=(sh.advantages, sh.targets),
            )

    def _shuffle_phase(state: OnPolicyLearnerState) -> None:
        """SHUFFLE MINIBATCHES (ff_ppo.py:294-307) for ALL epochs of this update: one keyed permutation of the flat index per
        (shard, epoch).  They depend on the key and the device-resident stream counter only, not on the rollout, so the
        update step issues them on a side stream underneath the rollout (which occupies E/128 of the 148 SMs)."""
        b = built
        for u in range(U):
            sh = b["shards"][u]
            if sh.perms is None:
                sh.perms = torch.zeros(epochs, T * E, dtype=torch.int32, device=b["dev"])
            for ep in range(epochs):
                ops.make_permutation(T * E, state.key[1] + u, ep, dev_counter=b["perm_ctr"], out=sh.perms[ep])

    def _update_phase(state: OnPolicyLearnerState, shuffled: bool = False) -> None:
        """UPDATE EPOCHS (ff_ppo.py:181-338)."""
        b = built
        if not shuffled:
            _shuffle_phase(state)
        sa, sc = b["sa"], b["sc"]
        a_tree = state.params.actor_params
        seeds = state.key
        D = sa.sizes[0]
        metrics = b["metrics"]
        metrics.zero_()
        grads = b["grads"]
        for ep in range(epochs):  # _update_epoch (ff_ppo.py:286-322)
            batches = []
            for u in range(U):
                sh = b["shards"][u]
                batches.append(ops.PpoBatch(sh.obs[:T].view(B, D), sh.action.view(B), sh.log_prob.view(B), sh.value.view(B),
                                            sh.advantages.view(B), sh.targets.view(B), sh.adv_stats, sh.perms[ep]))
            # single shard on a single device: the gradient reduction can hand sum(g^2) straight to the fused
            # optimiser (no separate norm pass / grid barrier); otherwise the all-reduce sits in between.
            prenorm = precision == ops.STX_PREC_BF16 and U == 1 and world == 1
            fused_update = bool(arch.get("fused_update", False)) or os.environ.get("STX_FUSED_UPDATE", "0") == "1"
            peers = b["peers_obj"]
            for i in range(nmb):  # _update_minibatch (ff_ppo.py:184-284)
                which = (ep * nmb + i) & 1
                if peers is not None:
                    grads = peers.bufs[which]  # ping-pong arenas in peer-mapped memory
                if prenorm and fused_update:
                    # gradients + both optimiser updates (ff_ppo.py:184-273) with the reduction and clip+Adam in one launch
                    ops.ppo_minibatch_update(sa, sc, b["arena"], batches[0], i * mb, mb, float(sysc.clip_eps), float(sysc.ent_coef),
                                             float(sysc.vf_coef), bool(sysc.standardize_advantages), grads, metrics[ep, i], b["ws"],
                                             b["plan"], a_tree.arena_mu, a_tree.arena_nu, b["arena_bf16"])
                    continue
                for u in range(U):  # vmap over "batch" + pmean("batch") (ff_ppo.py:253-256); first shard overwrites
                    ops.ppo_minibatch_grads(sa, sc, b["arena"], batches[u], i * mb, mb, float(sysc.clip_eps),
                                            float(sysc.ent_coef), float(sysc.vf_coef), bool(sysc.standardize_advantages),
                                            grads, metrics[ep, i], b["ws"], precision, 1.0 / U, b["arena_bf16"],
                                            overwrite=(u == 0), adam_scratch=b["plan"].scratch if prenorm else None)
                if peers is not None:
                    # pmean over "device" (ff_ppo.py:258-261) + both optimiser updates (:264-273) in ONE launch:
                    # one-shot all-reduce by direct NVLink peer loads fused into clip+Adam
                    ops.allreduce_clip_adam_step(b["plan"], peers, which, b["arena"], a_tree.arena_mu, a_tree.arena_nu,
                                                 params_bf16=b["arena_bf16"])
                    continue
                if world > 1:  # pmean over "device": summed by NCCL here, scaled in K4
                    dist.all_reduce(grads, op=dist.ReduceOp.SUM)
                # UPDATE ACTOR AND CRITIC PARAMS AND OPTIMISER STATE (ff_ppo.py:264-273), one launch
                ops.clip_adam_step(b["plan"], b["arena"], grads, a_tree.arena_mu, a_tree.arena_nu,
                                   grad_scale=1.0 / world, params_bf16=b["arena_bf16"], prenorm=prenorm)
        if world > 1:
            dist.all_reduce(metrics, op=dist.ReduceOp.SUM)
            metrics.mul_(1.0 / world)

    def _carry_in_phase(state: OnPolicyLearnerState) -> None:
        """The latest observation (row T of the previous rollout) is the first of this rollout
        (learner_state.timestep of ff_ppo.py:131-134); the finished trajectory stays readable until then."""
        for u in range(U):
            sh = built["shards"][u]
            sh.env_obs[0].copy_(sh.env_obs[T])

    def _advance_phase(state: OnPolicyLearnerState) -> None:
        """Advance the device-resident RNG stream positions (graph replays then draw fresh numbers)."""
        b = built
        for u in range(U):
            if hasattr(env, "advance"):
                env.advance(state.env_state[u], T)
        ops.counter_add(b["roll_ctr"], T)
        ops.counter_add(b["perm_ctr"], epochs)

    def _update_step(state: OnPolicyLearnerState) -> None:
        """A single update of the network (ff_ppo.py:61-341), in place on the learner state."""
        _carry_in_phase(state)
        main = torch.cuda.current_stream()
        side = built["side_stream"]
        side.wait_stream(main)  # fork (a graph branch when captured)
        with torch.cuda.stream(side):
            _shuffle_phase(state)
        _rollout_phase(state)
        _stats_phase(state)
        _gae_phase(state)
        main.wait_stream(side)  # join
        _update_phase(state, shuffled=True)
        _advance_phase(state)

    def learner_fn(learner_state: OnPolicyLearnerState) -> AnakinExperimentOutput[OnPolicyLearnerState]:
        """Run arch.num_updates_per_eval update steps (ff_ppo.py:343-370)."""
        learner_fn.ensure_built(learner_state)  # buffers + rollout seeded with last_timestep.observation
        b = built
        n_upd = int(arch.num_updates_per_eval)
        dev = b["dev"]
        ep_out = {
            "episode_return": torch.empty(n_upd, U, T, E, device=dev),
            "episode_length": torch.empty(n_upd, U, T, E, dtype=torch.int32, device=dev),
            "is_terminal_step": torch.empty(n_upd, U, T, E, dtype=torch.bool, device=dev),
        }
        train_out = torch.empty(n_upd, epochs, nmb, 8, device=dev)
        for k in range(n_upd):
            if use_graph and b["eager_done"] and b["graph"] is None:
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                # No cyclic GC while the stream is capturing: collecting an older learner (its CUDAGraph is destroyed
                # in the collector) is "not permitted when stream is capturing" and would invalidate this capture.
                gc.collect()
                gc_was_enabled = gc.isenabled()
                gc.disable()
                try:
                    with torch.cuda.graph(g):
                        _update_step(learner_state)
                finally:
                    if gc_was_enabled:
                        gc.enable()
                b["graph"] = g
            if b["graph"] is not None:
                b["graph"].replay()
            else:
                _update_step(learner_state)  # first update runs eagerly (module loading, NCCL warm-up)
                b["eager_done"] = True
            for u in range(U):
                sh = b["shards"][u]
                ep_out["episode_return"][k, u].copy_(sh.episode_return)
                ep_out["episode_length"][k, u].copy_(sh.episode_length)
                ep_out["is_terminal_step"][k, u].copy_(sh.is_terminal_step)
            train_out[k].copy_(b["metrics"])
        # the observation the next learn() call starts from
        new_ts = [learner_state.timestep[u]._replace(observation=b["shards"][u].env_obs[T]) for u in range(U)]
        learner_state = learner_state._replace(timestep=new_ts)
        train_metrics = {name: train_out[..., j] for j, name in enumerate(_METRIC_NAMES)}
        return AnakinExperimentOutput(learner_state=learner_state, episode_metrics=ep_out, train_metrics=train_metrics)

    learner_fn.built = built  # exposed for tests / bench (trajectory buffers, graph handle)
    learner_fn.update_step = _update_step
    learner_fn.phases = {"rollout": _rollout_phase, "stats": _stats_phase, "gae": _gae_phase, "update": _update_phase}

    def _ensure_built(st: OnPolicyLearnerState) -> None:
        if not built:
            _build(st)
            for u in range(U):  # last_timestep.observation seeds the carry slot (row T)
                built["shards"][u].env_obs[T].copy_(st.timestep[u].observation)

    learner_fn.ensure_built = _ensure_built
    return learner_fn


def learner_setup(
    env: Environment, keys: Tuple[int, int, int], config: DictConfig
) -> Tuple[LearnerFn[OnPolicyLearnerState], Actor, OnPolicyLearnerState]:
    """Initialise learner_fn, network, optimiser, environment and states (ff_ppo.py:424-551)."""
    rank, world = _world()
    device = torch.device("cuda", torch.cuda.current_device())
    n_devices = world

    # Get number/dimension of actions (ff_ppo.py:432-433).
    num_actions = int(env.action_space().num_values)
    config.system.action_dim = num_actions

    key, actor_net_key, critic_net_key = keys

    # Define network and optimiser (ff_ppo.py:439-463).
    actor_torso = instantiate(config.network.actor_network.pre_torso)
    actor_action_head = instantiate(config.network.actor_network.action_head, action_dim=num_actions)
    critic_torso = instantiate(config.network.critic_network.pre_torso)
    critic_head = instantiate(config.network.critic_network.critic_head)

    actor_network = Actor(torso=actor_torso, action_head=actor_action_head)
    critic_network = Critic(torso=critic_torso, critic_head=critic_head)
    precision = ops.STX_PREC_BF16 if str(config.arch.get("precision", "f32")) == "bf16" else ops.STX_PREC_F32
    actor_network.precision = critic_network.precision = precision

    actor_lr = make_learning_rate(config.system.actor_lr, config, config.system.epochs, config.system.num_minibatches)
    critic_lr = make_learning_rate(config.system.critic_lr, config, config.system.epochs, config.system.num_minibatches)
    actor_optim = optax.chain(optax.clip_by_global_norm(config.system.max_grad_norm), optax.adam(actor_lr, eps=1e-5))
    critic_optim = optax.chain(optax.clip_by_global_norm(config.system.max_grad_norm), optax.adam(critic_lr, eps=1e-5))

    # Initialise observation (ff_ppo.py:466-467).
    init_x = env.observation_space().generate_value()[None, ...].to(device)

    # One flat arena [actor | pad | critic | pad] holds both networks (K4, C1).
    sa, sc = actor_network.spec_for(init_x.shape[-1]), critic_network.spec_for(init_x.shape[-1])
    _, coff, total = ops.arena_offsets(sa, sc)
    arena = torch.zeros(total, dtype=torch.float32, device=device)
    actor_params = actor_network.init(actor_net_key, init_x, flat=arena[:coff])
    critic_params = critic_network.init(critic_net_key, init_x, flat=arena[coff:])
    arena_bf16 = None
    if precision == ops.STX_PREC_BF16:
        arena_bf16 = ops.cast_bf16(arena)
        actor_params.flat_bf16, critic_params.flat_bf16 = arena_bf16[:coff], arena_bf16[coff:]
    actor_params.arena, actor_params.arena_bf16 = arena, arena_bf16

    # Optimiser state: flat mu / nu arenas + device-resident counters (ScaleByAdamState, ScaleByScheduleState).
    mu, nu = torch.zeros_like(arena), torch.zeros_like(arena)
    counts = torch.zeros(4, dtype=torch.int32, device=device)
    a_state = optax.OptState(counts[0:1], mu[: sa.param_count], nu[: sa.param_count], counts[1:2])
    c_state = optax.OptState(counts[2:3], mu[coff : coff + sc.param_count], nu[coff : coff + sc.param_count], counts[3:4])
    # the flat arenas every fused kernel works on ride along on the actor tree
    actor_params.arena_mu, actor_params.arena_nu, actor_params.arena_counts = mu, nu, counts

    params = ActorCriticParams(actor_params, critic_params)
    opt_states = ActorCriticOptStates(a_state, c_state)

    apply_fns = (actor_network.apply, critic_network.apply)
    update_fns = (actor_optim.update, critic_optim.update)
    learn = get_learner_fn(env, apply_fns, update_fns, config)

    # Initialise environment states and timesteps across update-batch shards (ff_ppo.py:492-501).
    U, E = int(config.arch.update_batch_size), int(config.arch.num_envs)
    env_states: List[Any] = []
    timesteps: List[TimeStep] = []
    shard_keys = srandom.split(key, U + 1)
    for u in range(U):
        if hasattr(env, "seed"):
            env.seed = (int(config.arch.seed) + 7919 * rank + 15485863 * u) & ((1 << 63) - 1)
        st, ts = env.reset(srandom.split(shard_keys[u], E))
        env_states.append(st)
        timesteps.append(ts)

    # Load model from checkpoint if specified (ff_ppo.py:504-512) -- params only, as in the reference.
    if config.logger.checkpointing.load_model:
        from stoix_b200.utils.checkpointing import Checkpointer

        loaded = Checkpointer(model_name=config.system.system_name, **to_container(config.logger.checkpointing.load_args))
        loaded.restore_params(arena)
        if arena_bf16 is not None:
            ops.cast_bf16(arena, out=arena_bf16)

    # step keys: (sampling seed, shuffle seed) -- per-rank streams like the per-device keys at :515-518
    step_key = srandom.split(shard_keys[U], 2 * n_devices)
    rollout_seed = step_key[2 * rank] & ((1 << 62) - 1)
    shuffle_seed = step_key[2 * rank + 1] & ((1 << 62) - 1)

    init_learner_state = OnPolicyLearnerState(
        params=params, opt_states=opt_states, key=(rollout_seed, shuffle_seed), env_state=env_states, timestep=timesteps,
    )
    # If normalizing observations, initialize running statistics from warm-up rollouts (ff_ppo.py:539-549); this adds a
    # `running_statistics` field to the learner state.
    if config.system.normalize_observations:
        from stoix_b200.utils import running_statistics as rstat

        warmup_observations = _collect_obs_norm_rollouts(env, key, config)
        running_statistics = rstat.initialize_statistics_from_data(init_x[0], warmup_observations)
        init_learner_state = rstat.create_with_running_statistics(init_learner_state, running_statistics)
    return learn, actor_network, init_learner_state


def _collect_obs_norm_rollouts(env: Environment, key: int, config: DictConfig) -> torch.Tensor:
    """Collect observations for observation normalisation by taking uniformly random actions for
    system.obs_norm_warmup_steps steps (ff_ppo.py:375-421); not counted in the timestep budget.
    Returns (warmup_steps + 1, num_envs, *obs_shape) raw observations of this rank's env shard."""
    steps, E = int(config.system.obs_norm_warmup_steps), int(config.arch.num_envs)
    if _world()[0] == 0:
        print(f"Initializing observation normalization with {steps * int(config.arch.total_num_envs)} observations... "
              "Be aware, we do not count this in the timestep budget.")
    keys = srandom.split(key, E + 1)
    if hasattr(env, "seed"):
        env.seed = (int(config.arch.seed) + 104729 * (_world()[0] + 1)) & ((1 << 63) - 1)
    env_state, ts = env.reset(keys[:E])
    dev = ts.observation.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(keys[E]) & ((1 << 62) - 1))
    obs = [ts.observation.float().clone()]
    for _ in range(steps):
        action = torch.randint(0, int(config.system.action_dim), (E,), device=dev, generator=gen, dtype=torch.int32)
        env_state, ts = env.step(env_state, action)
        obs.append(ts.observation.float().clone())
    return torch.stack(obs, 0)


def get_final_step_metrics(metrics: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], bool]:
    """stoa.get_final_step_metrics as used at ff_ppo.py:624: keep entries of finished episodes."""
    mask = metrics["is_terminal_step"]
    has_final = bool(mask.any().item())
    out = {k: (v[mask] if has_final else v.reshape(-1)[:0]) for k, v in metrics.items() if k != "is_terminal_step"}
    return out, has_final


def run_experiment(_config: DictConfig) -> float:
    """Runs experiment (ff_ppo.py:554-706)."""
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

    key, key_e, actor_net_key, critic_net_key = srandom.split(srandom.PRNGKey(config.arch.seed), num=4)
    learn, actor_network, learner_state = learner_setup(env, (key, actor_net_key, critic_net_key), config)

    evaluator, absolute_metric_evaluator = evaluator_setup(
        eval_env=eval_env, key_e=key_e, eval_act_fn=get_distribution_act_fn(config, actor_network.apply), config=config,
    )

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
        torch.cuda.synchronize()  # jax.block_until_ready (ff_ppo.py:619)
        elapsed_time = time.time() - start_time
        t = int(steps_per_rollout * (eval_step + 1))
        episode_metrics, ep_completed = get_final_step_metrics(learner_output.episode_metrics)
        episode_metrics["steps_per_second"] = torch.tensor(steps_per_rollout / elapsed_time)

        logger.log({"timestep": t}, t, eval_step, LogEvent.MISC)
        if ep_completed:
            logger.log(episode_metrics, t, eval_step, LogEvent.ACT)
        train_metrics = dict(learner_output.train_metrics)
        opt_steps_per_eval = config.arch.num_updates_per_eval * (config.system.epochs * config.system.num_minibatches)
        train_metrics["steps_per_second"] = torch.tensor(opt_steps_per_eval / elapsed_time)
        logger.log(train_metrics, t, eval_step, LogEvent.TRAIN)

        start_time = time.time()
        trained_params = learner_output.learner_state.params.actor_params
        evaluator_output = evaluator(trained_params, srandom.split(key_e, eval_step + 2)[-1],
                                     running_statistics=getattr(learner_state, "running_statistics", None))
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
        if learner_state.params.actor_params.flat_bf16 is not None:
            best_tree.flat_bf16 = ops.cast_bf16(best_params)
        evaluator_output = absolute_metric_evaluator(best_tree, srandom.split(key_e, 1)[0],
                                                     running_statistics=getattr(learner_state, "running_statistics", None))
        torch.cuda.synchronize()
        elapsed_time = time.time() - start_time
        steps_per_eval = int(evaluator_output["episode_length"].sum().item())
        evaluator_output["steps_per_second"] = torch.tensor(steps_per_eval / max(elapsed_time, 1e-9))
        logger.log(evaluator_output, t, eval_step, LogEvent.ABSOLUTE)

    logger.stop()
    return float(evaluator_output[config.env.eval_metric].float().mean().item())


def hydra_entry_point(cfg: Optional[DictConfig] = None, overrides: Optional[List[str]] = None) -> float:
    """Experiment entry point (ff_ppo.py:709-727).  Without hydra-core the composition happens here:
    `python -m stoix_b200.systems.ppo.anakin.ff_ppo env=synthetic/box arch.total_num_envs=4096 ...`"""
    if cfg is None:
        cfg = compose("default_ff_ppo", overrides if overrides is not None else sys.argv[1:], config_dir="default/anakin")
    if "RANK" in os.environ and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    t0 = time.time()
    eval_performance = run_experiment(cfg)
    if _world()[0] == 0:
        print(f"PPO experiment completed in {time.time() - t0:.2f} seconds.")
    return eval_performance


if __name__ == "__main__":
    hydra_entry_point()
