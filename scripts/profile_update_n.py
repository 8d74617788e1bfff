"""Per-kernel GPU time of the captured Anakin update step at N ranks (torch profiler / CUPTI on graph replays), rank 0 prints.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/profile_update_n.py
    python scripts/profile_update_n.py            (N = 1)
Answers: what does the optimiser step cost with the all-reduce inside (clip_adam_kernel<..., PEER>) next to the single-GPU one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from stoix_b200 import random as srandom
    from stoix_b200.config import compose
    from stoix_b200.systems.ppo.anakin import ff_ppo
    from stoix_b200.utils import make_env
    from stoix_b200.utils.total_timestep_checker import check_total_timesteps

    E, T = 4096, 128
    cfg = compose("default_ff_ppo", ["env=synthetic/box", "env.kwargs.obs_dim=64", "env.kwargs.num_actions=8", f"arch.total_num_envs={E * world}",
                                     f"system.rollout_length={T}", "system.epochs=4", "system.num_minibatches=16",
                                     f"arch.total_timesteps={E * world * T * 40}", "arch.num_evaluation=1", "arch.precision=bf16",
                                     "logger.use_console=False"] + sys.argv[1:])
    cfg.num_devices, cfg.rank = world, rank
    cfg = check_total_timesteps(cfg, quiet=True)
    env, _ = make_env.make(cfg)
    keys = srandom.split(srandom.PRNGKey(cfg.arch.seed), 4)
    learn, _, state = ff_ppo.learner_setup(env, (keys[0], keys[2], keys[3]), cfg)
    cfg.arch.num_updates_per_eval = 1
    for _ in range(5):
        state = learn(state).learner_state
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n = 5
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            state = learn(state).learner_state
        torch.cuda.synchronize()
    if rank == 0:
        print(f"# N={world}: {n} captured update steps, per-kernel totals on rank 0")
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
