"""Two-rank data-parallel update (the reference's pmap 'device' axis -> one process per GPU + NCCL
all-reduce of the flat gradient arena, ff_ppo.py:258-261).  Needs >= 2 GPUs; run with
`gpurun --gpus 2 -- python -m pytest tests/test_distributed_gpu.py -m gpu`."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from stoix_b200 import random as srandom
from stoix_b200.config import compose
from stoix_b200.systems.ppo.anakin import ff_ppo
from stoix_b200.utils import make_env
from stoix_b200.utils.total_timestep_checker import check_total_timesteps
precision, fused = sys.argv[1], sys.argv[2]
E, T = 256, 16
cfg = compose("default_ff_ppo", ["env=synthetic/box", f"arch.total_num_envs={E * world}", f"system.rollout_length={T}",
                                 "system.num_minibatches=4", f"arch.total_timesteps={E * world * T * 4}", "arch.num_evaluation=1",
                                 f"arch.precision={precision}", f"arch.fused_allreduce={fused}", "logger.use_console=False"])
cfg.num_devices, cfg.rank = world, rank
cfg = check_total_timesteps(cfg, quiet=True)
env, _ = make_env.make(cfg)
keys = srandom.split(srandom.PRNGKey(cfg.arch.seed), 4)
learn, _, state = ff_ppo.learner_setup(env, (keys[0], keys[2], keys[3]), cfg)
cfg.arch.num_updates_per_eval = 1
p0 = state.params.actor_params.arena.clone()
for _ in range(3):   # eager, capture, replay
    out = learn(state); state = out.learner_state
torch.cuda.synchronize()
arena = state.params.actor_params.arena
gathered = [torch.empty_like(arena) for _ in range(world)]
dist.all_gather(gathered, arena)
obs_sum = learn.built["shards"][0].obs[:T].float().sum()
sums = [torch.empty_like(obs_sum) for _ in range(world)]
dist.all_gather(sums, obs_sum)
if rank == 0:
    same = all(torch.equal(gathered[0], g) for g in gathered)
    print(json.dumps({"params_identical_across_ranks": same, "changed": not torch.equal(p0, arena), "finite": bool(torch.isfinite(arena).all()),
                      "shards_differ": len({float(s) for s in sums}) == world, "fused_used": learn.built["peers_obj"] is not None,
                      "arena_sum": float(arena.double().sum()),
                      "value_loss": float(out.train_metrics["value_loss"].mean())}))
dist.barrier(); torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)  # NCCL teardown can hang at exit here
'''


@pytest.mark.parametrize("precision,fused", [("f32", "False"), ("bf16", "False"), ("bf16", "True"), ("f32", "True")])
def test_two_rank_update_keeps_replicas_identical(precision, fused, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29511", str(script), precision, fused]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=150, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    import json

    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["params_identical_across_ranks"] and res["changed"] and res["finite"] and res["shards_differ"], res
    assert res["fused_used"] == (fused == "True"), res
    print(res)
