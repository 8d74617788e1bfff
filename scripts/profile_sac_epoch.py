"""Per-kernel time of one ff_sac update epoch (torch profiler, CUPTI): which launches own the step.
    python scripts/profile_sac_epoch.py [--envs 1024] [--batch 256]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    from stoix_b200 import random as srandom
    from stoix_b200.config import compose
    from stoix_b200.systems.sac import ff_sac
    from stoix_b200.utils import make_env as environments
    from stoix_b200.utils.total_timestep_checker import check_total_timesteps

    torch.cuda.set_device(0)
    cfg = compose("default_ff_sac", [f"arch.total_num_envs={a.envs}", f"system.total_batch_size={a.batch}", "system.total_buffer_size=100000",
                                     "arch.cuda_graph=False", f"arch.total_timesteps={a.envs * 8}", "arch.num_evaluation=2",
                                     "logger.use_console=False"], config_dir="default/anakin")
    cfg.num_devices, cfg.rank = 1, 0
    cfg = check_total_timesteps(cfg, quiet=True)
    env, _ = environments.make(cfg)
    learn, _, state = ff_sac.learner_setup(env, tuple(srandom.split(srandom.PRNGKey(0), 3)), cfg)
    out = learn(state)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            learn.rollout_phase(out.learner_state)
            learn.update_epoch(out.learner_state, 0)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))


if __name__ == "__main__":
    main()
