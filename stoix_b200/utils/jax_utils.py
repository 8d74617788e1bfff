"""Shape helpers with the reference's names (stoix/utils/jax_utils.py:29-65), on torch tensors."""
from __future__ import annotations

from typing import Any

import torch


def tree_map(fn, tree: Any) -> Any:
    if isinstance(tree, dict):
        out = {k: tree_map(fn, v) for k, v in tree.items()}
        return out
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return type(tree)(*[tree_map(fn, v) for v in tree])
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(fn, v) for v in tree)
    return fn(tree) if isinstance(tree, torch.Tensor) else tree


def merge_leading_dims(x: torch.Tensor, num_dims: int) -> torch.Tensor:
    """jax_utils.py:29-43: flat index of (T, E) is t*E + e."""
    if x.ndim < num_dims:
        return x
    return x.reshape((-1,) + tuple(x.shape[num_dims:]))


def unreplicate_n_dims(x: Any, unreplicate_depth: int = 2) -> Any:
    return tree_map(lambda t: t[(0,) * unreplicate_depth], x)


def unreplicate_batch_dim(x: Any) -> Any:
    return tree_map(lambda t: t[:, 0, ...], x)
