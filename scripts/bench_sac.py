"""ff_sac update-step timing at the BASELINE configs[3] shape (synthetic Box obs_dim=17 act_dim=6, buffer 1e6, MLP 4x256 actor,
LayerNorm 4x256 twin-Q): env steps/s and SGD steps/s of the captured update step, plus the per-phase split measured eagerly.

    python scripts/bench_sac.py [--envs 1024] [--batch 256] [--updates 200] [--rollout 1] [--epochs 1]
Prints one JSON line.  fp32 SIMT kernels (parity first; no tensor-core path for this system yet)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--buffer", type=int, default=1_000_000)
    ap.add_argument("--updates", type=int, default=200)
    ap.add_argument("--rollout", type=int, default=1)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    from stoix_b200 import random as srandom
    from stoix_b200.config import compose
    from stoix_b200.systems.sac import ff_sac
    from stoix_b200.utils import make_env as environments
    from stoix_b200.utils.total_timestep_checker import check_total_timesteps

    torch.cuda.set_device(0)
    cfg = compose("default_ff_sac", [f"arch.total_num_envs={a.envs}", f"system.total_batch_size={a.batch}", f"system.total_buffer_size={a.buffer}",
                                     f"system.rollout_length={a.rollout}", f"system.epochs={a.epochs}", f"arch.cuda_graph={not a.no_graph}",
                                     f"arch.total_timesteps={a.envs * a.rollout * a.updates * 2}", "arch.num_evaluation=2",
                                     "logger.use_console=False"], config_dir="default/anakin")
    cfg.num_devices, cfg.rank = 1, 0
    cfg = check_total_timesteps(cfg, quiet=True)
    env, _ = environments.make(cfg)
    learn, _, state = ff_sac.learner_setup(env, tuple(srandom.split(srandom.PRNGKey(0), 3)), cfg)
    out = learn(state)                       # eager update + capture + replays (warm-up)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = learn(out.learner_state)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / cfg.arch.num_updates_per_eval
    # phase split, eager, on the same state
    b = learn.built
    ph = {}
    for name, fn in (("rollout+add", lambda: learn.rollout_phase(out.learner_state)), ("epoch", lambda: learn.update_epoch(out.learner_state, 0))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ph[name + "_eager_ms"] = e0.elapsed_time(e1) / 20
    steps = a.envs * a.rollout
    print(json.dumps({"metric": "env steps/sec ff_sac Anakin (synthetic continuous Box)", "value": steps / ms * 1e3, "unit": "env_steps/s",
                      "sgd_steps_per_s": a.epochs / ms * 1e3, "ms_per_update": ms, "n_gpus": 1,
                      "config": {"workload": f"ff_sac, obs_dim=17 act_dim=6, envs={a.envs}, rollout={a.rollout}, epochs={a.epochs}, batch={a.batch}, "
                                             f"buffer={a.buffer}, actor MLP[256x4] silu, twin-Q MLP[256x4] LayerNorm silu", "dtype": "f32",
                                 "cuda_graph": not a.no_graph}, **ph}))


if __name__ == "__main__":
    main()
