"""MLPTorso -- constructor-compatible with stoix/networks/torso.py:12-33.

A torso here is a *description* (layer sizes, activation); the arithmetic runs in the fused CUDA MLP
kernels, so the object carries no tensors of its own."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .utils import parse_activation_fn


class MLPTorso:
    def __init__(self, layer_sizes: Sequence[int], activation: str = "relu", use_layer_norm: bool = False,
                 kernel_init: Optional[float] = None, activate_final: bool = True):
        if not activate_final:
            raise NotImplementedError("MLPTorso(activate_final=False) is not built (every PPO network config activates the last torso layer)")
        self.layer_sizes = tuple(int(s) for s in layer_sizes)
        self.activation = parse_activation_fn(activation)
        self.use_layer_norm = bool(use_layer_norm)   # Dense(use_bias=False) -> LayerNorm -> activation (torso.py:26-32)
        self.activate_final = True
        # orthogonal(sqrt(2)) is the reference default (torso.py:18)
        self.kernel_init_scale = float(np.sqrt(2.0)) if kernel_init is None else float(kernel_init)
