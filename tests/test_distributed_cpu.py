"""The N>1 data path on CPU: two `gloo` ranks run the update step of the oracle with the gradient exchange the
learner uses -- all-reduce(SUM) of the flat gradient arenas per minibatch step, 1/world folded into the optimiser
(`pmean` over "device", stoix/systems/ppo/anakin/ff_ppo.py:258-261) -- plus the per-rank env sharding of
`total_timestep_checker.py:57-61`.  (The CUDA kernels themselves are covered by tests/test_distributed_gpu.py.)"""
import json
import os
import socket
import subprocess
import sys

import numpy as np

WORKER = r'''
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd())
from oracle import ppo_oracle as O
from stoix_b200.config import compose
from stoix_b200.utils.total_timestep_checker import check_total_timesteps

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['MASTER_PORT']}", rank=rank, world_size=world)
same_shards = sys.argv[1] == "same"

# host logic: every rank derives its own env shard from the global config
E, T, D, A = 4, 4, 6, 3
cfg = compose("default_ff_ppo", ["env=synthetic/box", f"arch.total_num_envs={E * world}", f"system.rollout_length={T}",
                                 "system.num_minibatches=2", "system.epochs=2", f"arch.total_timesteps={E * world * T * 3}",
                                 "arch.num_evaluation=1", "logger.use_console=False"])
cfg.num_devices, cfg.rank = world, rank
cfg = check_total_timesteps(cfg, quiet=True)
assert cfg.arch.num_envs == E and cfg.arch.num_updates == 3, (cfg.arch.num_envs, cfg.arch.num_updates)

def shard(seed):
    r = np.random.default_rng(seed)
    tr = O.Trajectory(obs=r.standard_normal((T, E, D)), action=r.integers(0, A, (T, E)), reward=r.standard_normal((T, E)),
                      done=r.random((T, E)) < 0.1, truncated=np.zeros((T, E), bool), next_obs=r.standard_normal((T, E, D)))
    return tr

init = np.random.default_rng(0)
actor, critic = O.init_mlp(init, [D, 16, 16, A], 0.01), O.init_mlp(init, [D, 16, 16, 1], 1.0)
h = O.PPOHyper(epochs=2, num_minibatches=2, num_updates=3, max_grad_norm=0.05)
perms = np.stack([np.random.default_rng(100 + ep).permutation(T * E) for ep in range(2)])

def run(traj, sync):
    a, c = O.MLPParams([w.copy() for w in actor.W], [b.copy() for b in actor.b]), O.MLPParams([w.copy() for w in critic.W], [b.copy() for b in critic.b])
    traj = O.evaluate_rollout(a, c, traj)
    na, nc = a.flat().size, c.flat().size
    sa, sc = O.AdamState(np.zeros(na), np.zeros(na)), O.AdamState(np.zeros(nc), np.zeros(nc))
    a2, c2, metrics, _, _ = O.ppo_update(a, c, sa, sc, traj, perms, h, grad_sync=sync)
    return np.concatenate([a2.flat(), c2.flat()]), metrics

calls = [0]
def grad_sync(a_g, c_g, info):
    calls[0] += 1
    arena = torch.from_numpy(np.concatenate([a_g, c_g]))        # ONE flat arena, like the learner's gradient arena
    dist.all_reduce(arena, op=dist.ReduceOp.SUM)                 # summed by the collective ...
    arena = arena.numpy() * (1.0 / world)                        # ... 1/world applied by the optimiser (K4 grad_scale)
    m = torch.tensor([info["actor_loss"], info["entropy"], info["value_loss"]], dtype=torch.float64)
    dist.all_reduce(m, op=dist.ReduceOp.SUM)
    m = m.numpy() / world
    return arena[: a_g.size], arena[a_g.size :], {**info, "actor_loss": m[0], "entropy": m[1], "value_loss": m[2]}

mine = shard(7 if same_shards else 7 + rank)
params, metrics = run(mine, grad_sync)
gathered = [torch.empty(params.size, dtype=torch.float64) for _ in range(world)]
dist.all_gather(gathered, torch.from_numpy(params))
solo, _ = run(shard(7 if same_shards else 7 + rank), None)
if rank == 0:
    print(json.dumps({"replicas_identical": all(torch.equal(gathered[0], g) for g in gathered),
                      "sync_calls": calls[0], "max_abs_vs_solo": float(np.abs(params - solo).max()),
                      "finite": bool(np.isfinite(params).all())}))
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(mode, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), mode], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                      text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (_, err) in zip(procs, outs):
        assert p.returncode == 0, err[-2000:]
    return json.loads(outs[0][0].strip().splitlines()[-1])


def test_two_gloo_ranks_same_shard_equals_single_process(tmp_path):
    """pmean of identical gradients is the gradient: (g + g) * 0.5 == g exactly, so two ranks holding the same shard
    must reproduce the single-process update bit for bit -- sum + 1/world is the exchange, nothing else."""
    r = _run("same", tmp_path)
    assert r["replicas_identical"] and r["finite"]
    assert r["sync_calls"] == 2 * 2  # epochs x minibatches: one exchange per optimiser step
    assert r["max_abs_vs_solo"] == 0.0


def test_two_gloo_ranks_different_shards_stay_identical(tmp_path):
    """Different env shards per rank (total_timestep_checker.py:57-61): the replicas stay identical and differ from what
    either rank would have learned alone."""
    r = _run("different", tmp_path)
    assert r["replicas_identical"] and r["finite"]
    assert r["max_abs_vs_solo"] > 1e-6
