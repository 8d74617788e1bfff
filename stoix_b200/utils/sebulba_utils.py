"""Thread plumbing of the Sebulba architecture -- drop-in for stoix/utils/sebulba_utils.py.

Same classes and methods (ThreadLifetime :20-43, OnPolicyPipeline :46-98, ParameterServer :101-259,
AsyncEvaluatorBase :262-...): actors -> learner through one Queue(maxsize=1) per actor, learner -> actors through one
parameter queue per actor.  What changes underneath:

* "parameters" are the learner's flat fp32 arena (+ its bf16 shadow): `distribute_params` makes ONE device-to-device
  copy of the arena per actor device (cudaMemcpyPeerAsync over NVLink when the actor is another GPU, on a copy stream
  of the parameter server) instead of a `jax.device_put` of a parameter tree; every actor thread of that device then
  receives the same ParamSnapshot through its queue.  A snapshot carries the CUDA event of its copy, and
  `get_params` waits for it (the reference's `jax.block_until_ready`).
* a timing helper (TimingTracker of stoix/utils/timing_utils.py) is included because the actor / learner loops log it.
"""
from __future__ import annotations

import queue
import threading
import time
import warnings
from abc import ABC, abstractmethod
from collections import defaultdict, deque
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch


class ThreadLifetime:
    """Manages thread lifecycle with stop signaling (sebulba_utils.py:20-43)."""

    def __init__(self, thread_name: str, thread_id: int) -> None:
        self._stop = False
        self.thread_name = thread_name
        self.thread_id = thread_id

    @property
    def name(self) -> str:
        return self.thread_name

    @property
    def id(self) -> int:
        return self.thread_id

    def __repr__(self) -> str:
        return f"Thread(thread_name={self.thread_name}, thread_id={self.thread_id}, stop={self._stop})"

    def should_stop(self) -> bool:
        return self._stop

    def stop(self) -> None:
        self._stop = True


class OnPolicyPipeline:
    """Handles rollout communication for on-policy Sebulba systems (sebulba_utils.py:46-98)."""

    def __init__(self, total_num_actors: int, queue_maxsize: int = 1):
        self.num_actors = total_num_actors
        self.rollout_queues: List[queue.Queue] = [queue.Queue(maxsize=queue_maxsize) for _ in range(total_num_actors)]

    def send_rollout(self, actor_idx: int, rollout_data: Tuple[int, int, Any], timeout: Optional[float] = None) -> bool:
        try:
            if timeout is not None:
                self.rollout_queues[actor_idx].put(rollout_data, timeout=timeout)
            else:
                self.rollout_queues[actor_idx].put(rollout_data)
            return True
        except queue.Full:
            return False

    def collect_rollouts(self, timeout: Optional[float] = None) -> List[Tuple[int, int, Any]]:
        """Collect rollout data from ALL actors (one entry per actor, in actor order: the learner's batch layout)."""
        collected = []
        for actor_idx in range(self.num_actors):
            try:
                collected.append(self.rollout_queues[actor_idx].get(timeout=timeout) if timeout is not None
                                 else self.rollout_queues[actor_idx].get())
            except queue.Empty:
                raise RuntimeError(f"Failed to collect rollout from actor {actor_idx}")
        return collected

    def clear_all_queues(self) -> None:
        for q in self.rollout_queues:
            while not q.empty():
                try:
                    q.get_nowait()
                except queue.Empty:
                    break


@dataclass
class ParamSnapshot:
    """One version of the parameters on one actor device: flat fp32 arena, bf16 shadow (or None) and the event that
    marks the end of the device-to-device copy."""

    arena: torch.Tensor
    arena_bf16: Optional[torch.Tensor]
    ready: Optional[torch.cuda.Event]
    version: int


class ParameterServer:
    """Handles parameter distribution for Sebulba systems (sebulba_utils.py:101-259)."""

    def __init__(self, total_num_actors: int, actor_devices: Sequence[torch.device], actors_per_device: int, queue_maxsize: int = 1):
        self.num_actors = total_num_actors
        self.actor_devices = [torch.device(d) for d in actor_devices]
        self.actors_per_device = actors_per_device
        self.param_queues: List[queue.Queue] = [queue.Queue(maxsize=queue_maxsize) for _ in range(total_num_actors)]
        self._copy_streams: Dict[torch.device, torch.cuda.Stream] = {}
        self._version = 0

    def _prepare_device_params(self, params: Any, device: torch.device, block_until_ready: bool) -> Optional[ParamSnapshot]:
        """One copy of the flat arena(s) onto `device`; `params` is an ActorCriticParams whose actor tree carries the
        arena (`.arena`, `.arena_bf16`) -- see stoix_b200/networks/base.py."""
        try:
            src = params.actor_params
            arena, shadow = src.arena, getattr(src, "arena_bf16", None)
            stream = self._copy_streams.get(device)
            if stream is None:
                stream = self._copy_streams[device] = torch.cuda.Stream(device=device)
            # the copy must see the learner's finished update: order it behind the learner's current stream
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(arena.device))
            with torch.cuda.device(device), torch.cuda.stream(stream):
                stream.wait_event(done)
                a = torch.empty(arena.shape, dtype=arena.dtype, device=device)
                a.copy_(arena, non_blocking=True)
                s = None
                if shadow is not None:
                    s = torch.empty(shadow.shape, dtype=shadow.dtype, device=device)
                    s.copy_(shadow, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
            if block_until_ready:
                ev.synchronize()
            return ParamSnapshot(a, s, ev, self._version)
        except Exception as e:  # the reference warns and carries on
            warnings.warn(f"Failed to place parameters on device {device}: {e}", stacklevel=2)
            return None

    def distribute_params(self, params: Any, block: bool = True, timeout: Optional[float] = None,
                          block_params_until_ready: bool = False) -> None:
        """Distribute parameters to all actors with device placement."""
        self._version += 1
        actor_idx = 0
        for device in self.actor_devices:
            snap = self._prepare_device_params(params, device, block_params_until_ready)
            if snap is not None:
                for i in range(self.actors_per_device):
                    self._put_params_in_queue(actor_idx + i, snap, block, timeout)
            actor_idx += self.actors_per_device

    def _put_params_in_queue(self, actor_idx: int, params: Any, block: bool, timeout: Optional[float]) -> None:
        try:
            if block:
                if timeout is not None:
                    self.param_queues[actor_idx].put(params, timeout=timeout)
                else:
                    self.param_queues[actor_idx].put(params)
            else:
                self.param_queues[actor_idx].put_nowait(params)
        except (queue.Full, queue.Empty):
            warnings.warn(f"Failed to put parameters in queue {actor_idx}", stacklevel=2)

    def get_params(self, actor_idx: int, timeout: Optional[float] = None) -> Optional[ParamSnapshot]:
        """Get parameters for an actor (None = shutdown signal); the calling stream waits for the copy."""
        try:
            snap = self.param_queues[actor_idx].get(timeout=timeout) if timeout is not None else self.param_queues[actor_idx].get()
        except queue.Empty:
            return None
        if snap is not None and snap.ready is not None:
            torch.cuda.current_stream(snap.arena.device).wait_event(snap.ready)
        return snap

    def shutdown_actors(self) -> None:
        for q in self.param_queues:
            try:
                q.put_nowait(None)
            except queue.Full:
                pass  # the actor will eventually check its lifetime

    def clear_all_queues(self) -> None:
        for q in self.param_queues:
            while not q.empty():
                try:
                    q.get_nowait()
                except queue.Empty:
                    break


class AsyncEvaluatorBase(threading.Thread, ABC):
    """Evaluation on its own thread (sebulba_utils.py:262-...): the learner submits (state, key, step, t) and carries on."""

    def __init__(self, evaluator: Callable, logger: Any, config: Any, checkpointer: Any, save_checkpoint: bool, lifetime: ThreadLifetime):
        super().__init__(name="AsyncEvaluator")
        self.evaluator = evaluator
        self.logger = logger
        self.config = config
        self.checkpointer = checkpointer
        self.save_checkpoint = save_checkpoint
        self.lifetime = lifetime
        self.eval_queue: queue.Queue = queue.Queue()
        self.max_episode_return = -float("inf")
        self.best_params: Any = None
        self.eval_step = 0
        self.eval_metrics: List[Dict[str, Any]] = []
        self.num_evaluation = max(int(config.arch.num_evaluation), 1)
        self._done = threading.Event()

    def submit_evaluation(self, learner_state: Any, eval_key: Any, eval_step: int, global_step_count: int) -> None:
        self.eval_queue.put((learner_state, eval_key, eval_step, global_step_count))

    def _update_best_params(self, episode_return: float, params: Any) -> None:
        if self.config.arch.absolute_metric and self.max_episode_return <= episode_return:
            self.best_params = params.flat.clone()
            self.max_episode_return = episode_return

    def _update_evaluation_progress(self) -> None:
        self.eval_step += 1
        if self.eval_step >= self.num_evaluation:
            self._done.set()

    def add_eval_metrics(self, metrics: Dict[str, Any]) -> None:
        self.eval_metrics.append(metrics)

    def wait_for_all_evaluations(self, timeout: Optional[float] = None) -> bool:
        return self._done.wait(timeout)

    def get_final_episode_return(self) -> float:
        if not self.eval_metrics:
            return float("nan")
        return float(self.eval_metrics[-1]["episode_return"].float().mean().item())

    def shutdown(self) -> None:
        self.lifetime.stop()
        self.eval_queue.put(None)

    @abstractmethod
    def run(self) -> None:
        ...


class TimingTracker:
    """Rolling means of named wall-clock spans (stoix/utils/timing_utils.py: `with timer.time(name)`, get_all_means)."""

    def __init__(self, maxlen: int = 10):
        self._spans: Dict[str, deque] = defaultdict(lambda: deque(maxlen=maxlen))

    @contextmanager
    def time(self, name: str):
        t0 = time.perf_counter()
        try:
            yield
        finally:
            self._spans[name].append(time.perf_counter() - t0)

    def get_all_means(self) -> Dict[str, float]:
        return {k: (sum(v) / len(v) if v else 0.0) for k, v in self._spans.items()}
