"""Console/JSON subset of stoix/utils/logger.py (StoixLogger.log(metrics, t, t_eval, LogEvent),
:77-158): TRAIN metrics are mean-reduced, the others described by mean/std/min/max."""
from __future__ import annotations

import json
import time
from enum import Enum
from pathlib import Path
from typing import Any, Dict

import numpy as np
import torch


class LogEvent(Enum):
    ACT = "actor"
    TRAIN = "trainer"
    EVAL = "evaluator"
    ABSOLUTE = "absolute"
    MISC = "misc"


def describe(x: np.ndarray) -> Dict[str, float]:
    if x.size <= 1:
        return {"": float(x.reshape(-1)[0])} if x.size else {}
    return {"mean": float(np.mean(x)), "std": float(np.std(x)), "min": float(np.min(x)), "max": float(np.max(x))}


class StoixLogger:
    def __init__(self, config):
        self.console = bool(config.logger.get("use_console", True))
        self.json_path = None
        if config.logger.get("use_json", False):
            base = Path(config.logger.get("base_exp_path", "results"))
            base.mkdir(parents=True, exist_ok=True)
            self.json_path = base / f"metrics_{int(time.time())}.jsonl"
        self.rank = int(config.get("rank", 0))

    def log_config(self, config: Dict[str, Any]) -> None:
        if self.json_path is not None and self.rank == 0:
            self.json_path.with_suffix(".config.json").write_text(json.dumps(config, default=str))

    def log(self, metrics: Dict[str, Any], t: int, t_eval: int, event: LogEvent) -> None:
        if self.rank != 0:
            return
        flat: Dict[str, float] = {}
        for k, v in metrics.items():
            arr = v.detach().float().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v, dtype=np.float64)
            if event == LogEvent.TRAIN:
                flat[k] = float(np.mean(arr))
            else:
                for stat, val in describe(arr).items():
                    flat[f"{k}/{stat}" if stat else k] = val
        if self.console:
            keys = [k for k in flat if k.endswith("/mean") or "/" not in k]
            body = " | ".join(f"{k.replace('/mean', '')}: {flat[k]:.4g}" for k in keys)
            print(f"[{event.value.upper():9s}] t={t:,} eval={t_eval} | {body}", flush=True)
        if self.json_path is not None:
            with self.json_path.open("a") as f:
                f.write(json.dumps({"event": event.value, "t": t, "eval": t_eval, **flat}) + "\n")

    def stop(self) -> None:
        return None
