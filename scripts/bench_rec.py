"""rec_ppo update-step timing at the BASELINE configs[4] shape (synthetic Box obs_dim=32, rollout_len=256, num_envs=2048/GPU), GRU(128)
actor-critic of configs/network/rnn.yaml, fp32.  Prints one JSON line with the phase split.

    python scripts/bench_rec.py [--envs 2048] [--rollout 256] [--epochs 4] [--minibatches 16] [--updates 2]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=2048)
    ap.add_argument("--rollout", type=int, default=256)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--minibatches", type=int, default=16)
    ap.add_argument("--updates", type=int, default=2)
    ap.add_argument("--cell", default="gru")
    a = ap.parse_args()
    from stoix_b200 import _lib, random as srandom
    from stoix_b200.config import compose
    from stoix_b200.systems.ppo.anakin import rec_ppo
    from stoix_b200.utils import make_env
    from stoix_b200.utils.total_timestep_checker import check_total_timesteps

    torch.cuda.set_device(0)
    cfg = compose("default_rec_ppo", ["env=synthetic/box", "env.kwargs.obs_dim=32", "env.kwargs.num_actions=8", f"arch.total_num_envs={a.envs}",
                                      f"system.rollout_length={a.rollout}", f"system.epochs={a.epochs}", f"system.num_minibatches={a.minibatches}",
                                      f"arch.total_timesteps={a.envs * a.rollout * (a.updates + 2)}", "arch.num_evaluation=1", "logger.use_console=False",
                                      f"network.actor_network.rnn_layer.cell_type={a.cell}", f"network.critic_network.rnn_layer.cell_type={a.cell}"],
                  config_dir="default/anakin")
    cfg.num_devices, cfg.rank = 1, 0
    cfg = check_total_timesteps(cfg, quiet=True)
    env, _ = make_env.make(cfg)
    learn, _, state = rec_ppo.learner_setup(env, tuple(srandom.split(srandom.PRNGKey(0), 3)), cfg)
    cfg.arch.num_updates_per_eval = 1
    for _ in range(2):                          # warm-up: the first update runs eagerly, the second captures the CUDA graph
        state = learn(state).learner_state
    torch.cuda.synchronize()
    lib = _lib.load()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    phase = {"rollout": [], "gae": [], "update": []}
    l0 = lib.stx_launch_count()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(a.updates):
        state = learn(state).learner_state
    e1.record()
    torch.cuda.synchronize()
    launches = (lib.stx_launch_count() - l0) / a.updates
    ms = e0.elapsed_time(e1) / a.updates
    for name in phase:          # the phases alone, on the same state
        s, e = ev(), ev()
        s.record()
        learn.phases[name](state)
        e.record()
        torch.cuda.synchronize()
        phase[name] = s.elapsed_time(e)
    steps = a.envs * a.rollout
    print(json.dumps({"metric": f"env steps/sec rec_ppo Anakin (synthetic Box, {a.cell.upper()} actor-critic)", "value": steps / ms * 1e3, "unit": "env_steps/s",
                      "ms_per_update": ms, "n_gpus": 1, "phase_ms": phase, "stx_launches_per_update": launches,
                      "config": {"workload": f"rec_ppo, obs_dim=32, envs={a.envs}, rollout={a.rollout}, epochs={a.epochs}, minibatches={a.minibatches}, "
                                             f"pre MLP[128] silu -> {a.cell.upper()}(128) -> post MLP[128] silu, fp32", "dtype": "f32", "cuda_graph": bool(cfg.arch.get("cuda_graph", True)), "phase_ms_note": "phases timed eagerly, ms_per_update from graph replays"}}))


if __name__ == "__main__":
    main()
