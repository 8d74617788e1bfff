"""Sebulba ff_ppo throughput on config-3-shaped batches (BASELINE configs[2]: 256 envs per actor thread), MLP torso on the
synthetic Box CPU environment (the conv torso / envpool of config 3 are outside this build).

    python scripts/bench_sebulba.py [--precision bf16] [--actor-device 0] [--learner-device 0] [--threads 2] [--updates 24]

One process: `threads` actor threads on the actor GPU (each: 256 CPU envs -> pinned H2D -> forward + sampling kernels ->
pinned D2H), one learner thread on the learner GPU.  Prints ONE JSON line: end-to-end env-steps/s (wall clock around the
learner loop, host env stepping included), the actors' per-step inference latency, the learner step time, bytes moved."""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from stoix_b200 import random as srandom  # noqa: E402
from stoix_b200.config import compose  # noqa: E402
from stoix_b200.envs import cpu as cpu_envs  # noqa: E402
from stoix_b200.systems.ppo.sebulba import ff_ppo as seb  # noqa: E402
from stoix_b200.utils.logger import StoixLogger  # noqa: E402
from stoix_b200.utils.sebulba_utils import OnPolicyPipeline, ParameterServer, ThreadLifetime  # noqa: E402
from stoix_b200.utils.total_timestep_checker import check_total_timesteps  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--actor-device", type=int, default=0)
ap.add_argument("--learner-device", type=int, default=0)
ap.add_argument("--threads", type=int, default=2)
ap.add_argument("--envs-per-thread", type=int, default=256)
ap.add_argument("--rollout-length", type=int, default=128)
ap.add_argument("--updates", type=int, default=24)
a = ap.parse_args()

E = a.threads * a.envs_per_thread
cfg = compose("default_ff_ppo", [f"arch.total_num_envs={E}", f"system.rollout_length={a.rollout_length}", "system.num_minibatches=4",
                                 f"arch.total_timesteps={E * a.rollout_length * a.updates}", f"arch.actor.actor_per_device={a.threads}",
                                 f"arch.actor.device_ids=[{a.actor_device}]", f"arch.learner.device_ids=[{a.learner_device}]",
                                 "arch.num_evaluation=1", "logger.use_console=False", f"arch.precision={a.precision}"],
              config_dir="default/sebulba")
actor_dev, learner_dev = torch.device("cuda", a.actor_device), torch.device("cuda", a.learner_device)
cfg.num_learner_devices, cfg.num_actor_devices, cfg.arch.world_size = 1, 1, 1
cfg.arch.total_num_actor_threads = a.threads
cfg = check_total_timesteps(cfg, quiet=True)
factory = cpu_envs.make_factory(cfg)
keys = srandom.split(srandom.PRNGKey(cfg.arch.seed), 4)
torch.cuda.set_device(learner_dev)
learn_step, apply_fns, state = seb.learner_setup(factory, (keys[0], keys[2], keys[3]), [learner_dev], cfg)
logger = StoixLogger(cfg)
ps = ParameterServer(a.threads, [actor_dev], a.threads, queue_maxsize=1)
pipe = OnPolicyPipeline(a.threads, queue_maxsize=1)
ps.distribute_params(state.params)
np_rng = np.random.default_rng(0)
threads, lifetimes = [], []
tkeys = srandom.split(keys[0], a.threads)
for i in range(a.threads):
    lt = ThreadLifetime(f"Actor-{i}", i)
    th = seb.get_actor_thread(factory, actor_dev, ps, pipe, apply_fns, tkeys[i], cfg, np_rng.integers(2 ** 31 - 1, size=a.envs_per_thread).tolist(),
                              logger, [learner_dev], lt)
    th.start()
    threads.append(th)
    lifetimes.append(lt)
learner = seb.get_learner_thread(cfg, learn_step, state, ps, pipe, logger, None, keys[1])
t0 = time.perf_counter()
learner.start()
learner.join()
wall = time.perf_counter() - t0
seb.stop_all_actor_threads(lifetimes, ps, pipe, threads)
ls = learner.learner_rollout.stats
acts = [th.rollout_fn.stats for th in threads]
steps = int(cfg.arch.num_updates) * E * a.rollout_length
mean = lambda k: float(np.mean([s.get(k, 0.0) for s in acts]))
print(json.dumps({
    "metric": "env steps/sec Sebulba ff_ppo (MLP torso, synthetic Box CPU envs)", "value": steps / wall, "unit": "env_steps/s",
    "config": {"workload": f"ff_ppo Sebulba, {a.threads} actor threads x {a.envs_per_thread} envs, rollout_len={a.rollout_length}, obs_dim=64, "
                           f"MLP[256,256], actor cuda:{a.actor_device}, learner cuda:{a.learner_device}", "precision": a.precision},
    "updates": int(ls["updates"]), "wall_s": wall, "learner_step_ms": 1e3 * ls.get("learn_step_time", 0.0),
    "learner_wait_for_rollouts_ms": 1e3 * ls.get("rollout_queue_get_time", 0.0), "param_broadcast_ms": 1e3 * ls.get("params_queue_put_time", 0.0),
    "actor_inference_us_per_step": 1e6 * mean("inference_time"), "actor_env_step_us": 1e6 * mean("env_step_time"),
    "actor_rollout_ms": 1e3 * mean("single_actor_rollout_time"), "actor_prepare_data_ms": 1e3 * mean("prepare_data_time"),
    "h2d_bytes_per_env_step": acts[0]["h2d_bytes"] / max(acts[0]["local_step_count"], 1),
    "d2h_bytes_per_env_step": acts[0]["d2h_bytes"] / max(acts[0]["local_step_count"], 1),
    "note": "host-side numpy env stepping is inside the wall clock; the learner step is launch-bound Python + the Anakin K2/K3/K4 kernels"}))
