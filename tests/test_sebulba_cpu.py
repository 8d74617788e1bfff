"""Host side of the Sebulba architecture on CPU: shape derivation (stoix/utils/total_timestep_checker.py:134-287), the
NumPy environments' stoa-style contract (stoix/wrappers/envpool.py:42-155), the actor -> learner pipeline and the thread
helpers (stoix/utils/sebulba_utils.py:20-98)."""
import threading

import numpy as np
import pytest

from stoix_b200.config import compose
from stoix_b200.envs import cpu as cpu_envs
from stoix_b200.utils.sebulba_utils import OnPolicyPipeline, ThreadLifetime, TimingTracker
from stoix_b200.utils.total_timestep_checker import check_total_timesteps


def _cfg(extra=()):
    c = compose("default_ff_ppo", ["arch.total_num_envs=512", "system.rollout_length=16", "arch.total_timesteps=400000",
                                   "arch.actor.actor_per_device=2", "arch.num_evaluation=4"] + list(extra), config_dir="default/sebulba")
    c.num_actor_devices, c.num_learner_devices, c.arch.world_size = 2, 1, 1
    return c


def test_sebulba_shape_derivation_matches_the_reference_formulas():
    c = check_total_timesteps(_cfg(), quiet=True)
    # total_timestep_checker.py:190-213
    assert c.arch.actor.num_envs_per_actor == 512 // 2 // 2
    assert c.arch.learner_parallel_env_consumption == 128 * 2 * 2
    assert c.arch.local_batch_size == 16 * 512 and c.arch.global_batch_size == 16 * 512
    # :235-263
    assert c.arch.num_updates == 400000 // 8192 and c.arch.num_updates_per_eval == (400000 // 8192) // 4
    c2 = _cfg(["arch.total_timesteps=~", "arch.num_updates=20"])
    c2 = check_total_timesteps(c2, quiet=True)
    assert c2.arch.total_timesteps == 20 * 8192
    with pytest.raises(AssertionError):   # :176-187
        bad = _cfg(["arch.total_num_envs=510"])
        check_total_timesteps(bad, quiet=True)
    with pytest.raises(AssertionError):   # :281-284: more evaluations than updates
        bad = _cfg(["arch.total_timesteps=16384"])
        check_total_timesteps(bad, quiet=True)


def test_synthetic_cpu_env_contract():
    env = cpu_envs.SyntheticBoxCpuEnv(64, obs_dim=8, num_actions=3, seed=1, p_term=0.2, p_trunc=0.1)
    ts = env.reset(seed=[0] * 64)
    assert ts.observation.shape == (64, 8) and ts.observation.dtype == np.float32 and (ts.step_type == cpu_envs.StepType.FIRST).all()
    assert not ts.last().any() and set(ts.extras["metrics"]) == {"episode_return", "episode_length", "is_terminal_step"}
    run_ret, run_len = np.zeros(64), np.zeros(64, np.int64)
    saw_term = saw_trunc = False
    for _ in range(60):
        ts = env.step(np.zeros(64, np.int32))
        done = ts.last()
        trunc = ts.step_type == cpu_envs.StepType.TRUNCATED
        term = ts.step_type == cpu_envs.StepType.TERMINATED
        assert (done == (trunc | term)).all()
        np.testing.assert_array_equal(ts.discount, np.where(term, 0.0, 1.0).astype(np.float32))   # envpool.py:143-146
        run_ret, run_len = run_ret + ts.reward, run_len + 1
        m = ts.extras["metrics"]
        np.testing.assert_array_equal(m["is_terminal_step"], done)
        np.testing.assert_allclose(m["episode_return"][done], run_ret[done], rtol=1e-6)          # envpool.py:118-133
        np.testing.assert_array_equal(m["episode_length"][done], run_len[done])
        run_ret[done], run_len[done] = 0.0, 0
        saw_term |= term.any()
        saw_trunc |= trunc.any()
    assert saw_term and saw_trunc
    with pytest.raises(AssertionError):
        env.step(np.zeros(3, np.int32))


def test_cartpole_cpu_env_terminates_and_resets():
    env = cpu_envs.CartPoleCpuEnv(32, seed=0)
    ts = env.reset(seed=list(range(32)))
    assert ts.observation.shape == (32, 4) and np.abs(ts.observation).max() <= 0.05
    lengths = []
    for _ in range(300):
        ts = env.step(np.ones(32, np.int32))      # always push right: the pole falls within a few dozen steps
        done = ts.last()
        if done.any():
            lengths.extend(ts.extras["metrics"]["episode_length"][done].tolist())
            assert np.abs(ts.observation[done]).max() <= 0.05   # auto-reset observation
            assert (ts.discount[done] == 0).all()
    assert lengths and 5 < np.mean(lengths) < 60
    assert (ts.reward == 1).all()


def test_env_factory_hands_out_distinct_seeds():
    c = _cfg()
    f = cpu_envs.make_factory(c)
    a, b = f(8), f(8)
    assert not np.array_equal(a.reset().observation, b.reset().observation)
    assert f.seed == int(c.arch.seed) + 16


def test_pipeline_collects_one_rollout_per_actor_in_actor_order():
    pipe = OnPolicyPipeline(total_num_actors=3, queue_maxsize=1)
    lifetimes = [ThreadLifetime(f"a{i}", i) for i in range(3)]

    def actor(i):
        for r in range(4):
            assert pipe.send_rollout(i, (r, i, f"payload-{i}-{r}"))   # blocks while the learner has not taken the previous one

    ths = [threading.Thread(target=actor, args=(i,)) for i in range(3)]
    for t in ths:
        t.start()
    for r in range(4):
        got = pipe.collect_rollouts(timeout=10)
        assert [g[1] for g in got] == [0, 1, 2] and all(g[0] == r for g in got)
    for t in ths:
        t.join(timeout=10)
    assert not pipe.send_rollout(0, "x", timeout=0.01) or pipe.rollout_queues[0].qsize() == 1
    pipe.clear_all_queues()
    assert all(q.empty() for q in pipe.rollout_queues)
    lifetimes[0].stop()
    assert lifetimes[0].should_stop() and not lifetimes[1].should_stop() and lifetimes[2].id == 2 and lifetimes[2].name == "a2"
    with pytest.raises(RuntimeError):
        pipe.collect_rollouts(timeout=0.01)


def test_timing_tracker_rolling_means():
    t = TimingTracker(maxlen=2)
    for _ in range(3):
        with t.time("x"):
            pass
    m = t.get_all_means()
    assert set(m) == {"x"} and m["x"] >= 0 and len(t._spans["x"]) == 2
