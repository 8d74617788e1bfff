"""Minimal 2-rank diagnostic: eager all_reduce, captured all_reduce, then one learner update (progress lines)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
def log(*a):
    print(f"[r{rank} {time.time() % 1000:.1f}]", *a, flush=True)
log("init")
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
log("init done")
x = torch.ones(1 << 16, device="cuda")
dist.all_reduce(x); torch.cuda.synchronize(); log("eager all_reduce ok", float(x[0]))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    dist.all_reduce(x)
log("captured")
g.replay(); torch.cuda.synchronize(); log("replay ok", float(x[0]))
stage = sys.argv[1] if len(sys.argv) > 1 else "nccl"
if stage == "learner":
    from stoix_b200 import random as srandom
    from stoix_b200.config import compose
    from stoix_b200.systems.ppo.anakin import ff_ppo
    from stoix_b200.utils import make_env
    from stoix_b200.utils.total_timestep_checker import check_total_timesteps
    E, T = 256, 16
    cfg = compose("default_ff_ppo", ["env=synthetic/box", f"arch.total_num_envs={E * world}", f"system.rollout_length={T}",
                                     "system.num_minibatches=4", f"arch.total_timesteps={E * world * T * 4}", "arch.num_evaluation=1",
                                     f"arch.precision={sys.argv[2]}", "logger.use_console=False", f"arch.cuda_graph={sys.argv[3]}"])
    cfg.num_devices, cfg.rank = world, rank
    cfg = check_total_timesteps(cfg, quiet=True)
    env, _ = make_env.make(cfg)
    keys = srandom.split(srandom.PRNGKey(cfg.arch.seed), 4)
    learn, _, state = ff_ppo.learner_setup(env, (keys[0], keys[2], keys[3]), cfg)
    cfg.arch.num_updates_per_eval = 1
    log("setup done")
    for i in range(3):
        out = learn(state); state = out.learner_state
        torch.cuda.synchronize(); log("update", i, "done")
dist.barrier(); log("bye"); os._exit(0)
