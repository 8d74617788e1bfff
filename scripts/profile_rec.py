"""Per-kernel GPU time of one rec_ppo update phase (torch profiler).   python scripts/profile_rec.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from stoix_b200 import random as srandom
from stoix_b200.config import compose
from stoix_b200.systems.ppo.anakin import rec_ppo
from stoix_b200.utils import make_env
from stoix_b200.utils.total_timestep_checker import check_total_timesteps

torch.cuda.set_device(0)
cfg = compose("default_rec_ppo", ["env=synthetic/box", "env.kwargs.obs_dim=32", "env.kwargs.num_actions=8", "arch.total_num_envs=2048", "system.rollout_length=256",
                                  "system.epochs=1", "system.num_minibatches=16", "arch.total_timesteps=2097152", "arch.num_evaluation=1",
                                  "logger.use_console=False"], config_dir="default/anakin")
cfg.num_devices, cfg.rank = 1, 0
cfg = check_total_timesteps(cfg, quiet=True)
env, _ = make_env.make(cfg)
learn, _, state = rec_ppo.learner_setup(env, tuple(srandom.split(srandom.PRNGKey(0), 3)), cfg)
cfg.arch.num_updates_per_eval = 1
state = learn(state).learner_state
torch.cuda.synchronize()
which = sys.argv[1] if len(sys.argv) > 1 else "update"
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    learn.phases[which](state)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=80))
