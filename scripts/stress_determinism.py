"""Run the bf16 learner repeatedly from the same seed and report any run-to-run difference (race detector)."""
import sys
import torch
sys.path.insert(0, ".")
from stoix_b200 import random as srandom
from stoix_b200.config import compose
from stoix_b200.systems.ppo.anakin import ff_ppo
from stoix_b200.utils import make_env
from stoix_b200.utils.total_timestep_checker import check_total_timesteps

FIELDS = ("obs", "next_obs", "action", "log_prob", "reward", "done", "truncated", "value", "advantages")


def run(fused, E=256, T=12, updates=3):
    cfg = compose("default_ff_ppo", ["env=synthetic/box", f"arch.total_num_envs={E}", f"system.rollout_length={T}",
                                     "system.num_minibatches=2", f"arch.total_timesteps={E * T * updates}", "arch.num_evaluation=1",
                                     "arch.precision=bf16", f"arch.fused_rollout={fused}", "logger.use_console=False",
                                     "env.kwargs.p_term=0.05", "env.kwargs.p_trunc=0.05"])
    cfg.num_devices, cfg.rank = 1, 0
    cfg = check_total_timesteps(cfg, quiet=True)
    env, _ = make_env.make(cfg)
    keys = srandom.split(srandom.PRNGKey(5), 4)
    learn, _, state = ff_ppo.learner_setup(env, (keys[0], keys[2], keys[3]), cfg)
    cfg.arch.num_updates_per_eval = 1
    outs = []
    for _ in range(updates):
        out = learn(state)
        state = out.learner_state
        torch.cuda.synchronize()
        sh = learn.built["shards"][0]
        outs.append({k: getattr(sh, k).clone() for k in FIELDS} | {"params": state.params.actor_params.arena.clone()})
    return outs


n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ref = {True: run(True), False: run(False)}
bad = 0
for i in range(n):
    for fused in (True, False):
        got = run(fused)
        for u, (g, r) in enumerate(zip(got, ref[fused])):
            for k in g:
                if not torch.equal(g[k], r[k]):
                    bad += 1
                    nd = (g[k] != r[k]).sum().item()
                    print(f"iter {i} fused={fused} update {u} field {k}: {nd} of {g[k].numel()} entries differ", flush=True)
                    break
for u in range(len(ref[True])):
    for k in ref[True][u]:
        if not torch.equal(ref[True][u][k], ref[False][u][k]):
            print(f"fused vs per-step: update {u} field {k} differs")
            bad += 1
print(f"done: {n} iterations, {bad} mismatches")
