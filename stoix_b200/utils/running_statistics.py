"""Running statistics of the observations -- drop-in for stoix/utils/running_statistics.py.

Same names, signatures and state contract (RunningStatisticsState :94-100; initialize_statistics :123-135;
initialize_statistics_from_data :137-161; update_statistics :204-345; normalize :348-363; denormalize :366-387;
add_field_to_state / create_with_running_statistics :444-560), for the case the PPO systems use: the nest is ONE
array (the observation).  The arithmetic runs in the CUDA kernels of csrc/stx_obsnorm.cu through the C ABI
(stx_running_stats_accumulate / _finalize, stx_obs_normalize); there is no CPU path.

Differences from the reference, by construction of this framework:
* `pmap_axes` names the reference's mapped axes; here "device" is the process group (one rank per GPU): when
  torch.distributed is initialised and pmap_axes is given, the accumulated sums are all-reduced once (the two psums
  of :302 and :308 collapse into that one exchange, see the kernel file).  The "batch" axis (update_batch_size
  shards) is expressed by passing the shards as a list of tensors.
* `count` is an int64 device tensor (the reference uses a Python float / int32 array).
"""
from __future__ import annotations

import types
from dataclasses import dataclass
from typing import Any, Dict, NamedTuple, Optional, Sequence, Type, Union

import torch
import torch.distributed as dist

from stoix_b200 import ops


@dataclass(frozen=True)
class NestedMeanStd:
    """running_statistics.py:86-91."""

    mean: torch.Tensor
    std: torch.Tensor


@dataclass(frozen=True)
class RunningStatisticsState(NestedMeanStd):
    """running_statistics.py:94-100 (field order of the reference's chex dataclass: mean, std, count, summed_variance)."""

    count: torch.Tensor
    summed_variance: torch.Tensor


def initialize_statistics(nest: torch.Tensor) -> RunningStatisticsState:
    """Zero mean / summed variance, std ONE: normalising with the initial state is the identity (:123-135)."""
    if not nest.is_cuda:
        raise ops.StxError("running statistics live on the GPU (no CPU path exists): pass a CUDA tensor as the nest")
    z = torch.zeros(nest.shape, dtype=torch.float32, device=nest.device)
    return RunningStatisticsState(mean=z.clone(), std=torch.ones_like(z), count=torch.zeros(1, dtype=torch.int64, device=nest.device),
                                  summed_variance=z.clone())


def _shards(batch: Union[torch.Tensor, Sequence[torch.Tensor]]) -> Sequence[torch.Tensor]:
    return list(batch) if isinstance(batch, (list, tuple)) else [batch]


def _validate_batch_shapes(batch: torch.Tensor, reference_sample: torch.Tensor) -> None:
    """running_statistics.py:164-186: trailing dims must equal the feature shape exactly (no silent broadcasting)."""
    nd = reference_sample.ndim
    if batch.ndim < nd or tuple(batch.shape[batch.ndim - nd:]) != tuple(reference_sample.shape):
        raise ValueError(f"batch of shape {tuple(batch.shape)} does not end with the feature shape {tuple(reference_sample.shape)}")


def update_statistics_(state: RunningStatisticsState, batch: Union[torch.Tensor, Sequence[torch.Tensor]], *,
                       weights: Optional[Union[torch.Tensor, Sequence[torch.Tensor]]] = None, std_min_value: float = 1e-6,
                       std_max_value: float = 1e6, pmap_axes: Optional[Union[str, Sequence[str]]] = None,
                       validate_shapes: bool = True, sums: Optional[torch.Tensor] = None) -> RunningStatisticsState:
    """In-place form of update_statistics (the learner calls this inside its CUDA graph): same arithmetic, the state
    tensors are overwritten.  `sums` is an optional preallocated float64[2D+1] work buffer."""
    shards = _shards(batch)
    wts = _shards(weights) if weights is not None else [None] * len(shards)
    D = state.mean.numel()
    total = None
    for x, w in zip(shards, wts):
        if validate_shapes:
            _validate_batch_shapes(x, state.mean)
            if w is not None and tuple(w.shape) != tuple(x.shape[: x.ndim - state.mean.ndim]):
                raise ValueError(f"{tuple(w.shape)} != {tuple(x.shape[: x.ndim - state.mean.ndim])}")
        x32 = x if x.dtype == torch.float32 else x.float()
        s = ops.running_stats_accumulate(x32.contiguous().view(-1, D), state.mean.view(-1), None if w is None else w.float().contiguous().view(-1),
                                         out=sums if (sums is not None and total is None) else None)
        total = s if total is None else total.add_(s)
    if pmap_axes is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)  # psum over "device" (:274-276, 302, 308)
    ops.running_stats_finalize(total, state.count, state.mean.view(-1), state.summed_variance.view(-1), state.std.view(-1),
                               std_min_value, std_max_value)
    return state


def update_statistics(state: RunningStatisticsState, batch: Union[torch.Tensor, Sequence[torch.Tensor]], *, config: Any = None,
                      weights: Optional[Union[torch.Tensor, Sequence[torch.Tensor]]] = None, std_min_value: float = 1e-6,
                      std_max_value: float = 1e6, pmap_axes: Optional[Union[str, Sequence[str]]] = None,
                      validate_shapes: bool = True) -> RunningStatisticsState:
    """Batched Welford update (running_statistics.py:204-345); returns a NEW state like the reference."""
    if config is not None:
        raise NotImplementedError("NestStatisticsConfig: the nest here is a single array (the observation)")
    new = RunningStatisticsState(mean=state.mean.clone(), std=state.std.clone(), count=state.count.clone(),
                                 summed_variance=state.summed_variance.clone())
    return update_statistics_(new, batch, weights=weights, std_min_value=std_min_value, std_max_value=std_max_value,
                              pmap_axes=pmap_axes, validate_shapes=validate_shapes)


def initialize_statistics_from_data(nest: torch.Tensor, data_sample: Union[torch.Tensor, Sequence[torch.Tensor]], *, config: Any = None,
                                    weights: Optional[torch.Tensor] = None, std_min_value: float = 5e-4, std_max_value: float = 5e4,
                                    pmap_axes: Optional[Union[str, Sequence[str]]] = None,
                                    validate_shapes: bool = True) -> RunningStatisticsState:
    """running_statistics.py:137-161."""
    return update_statistics(initialize_statistics(nest), data_sample, config=config, weights=weights, std_min_value=std_min_value,
                             std_max_value=std_max_value, pmap_axes=pmap_axes, validate_shapes=validate_shapes)


def normalize(batch: torch.Tensor, mean_std: NestedMeanStd, max_abs_value: Optional[float] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(batch - mean) / std on inexact dtypes, optional symmetric clip (running_statistics.py:348-363)."""
    if not batch.dtype.is_floating_point:
        return batch
    x32 = batch if batch.dtype == torch.float32 else batch.float()
    return ops.obs_normalize(x32.contiguous(), mean_std.mean.view(-1), mean_std.std.view(-1), out=out, max_abs_value=max_abs_value)


def denormalize(batch: torch.Tensor, mean_std: NestedMeanStd) -> torch.Tensor:
    """running_statistics.py:366-387 (cold path: plain tensor arithmetic)."""
    if not batch.dtype.is_floating_point:
        return batch
    return batch * mean_std.std + mean_std.mean


def add_field_to_state(base_class: Type[Any], extra_field_name: str, extra_field_type: Type[Any]) -> Type[Any]:
    """A NamedTuple class with base_class's fields plus one extra that is reachable as an attribute and through
    _replace but is NOT part of unpacking (`a, b, c = state` yields the original fields): running_statistics.py:444-530."""
    fields = tuple(getattr(base_class, "_fields", ()))
    annotations = dict(getattr(base_class, "__annotations__", {}))
    new_fields: Dict[str, Any] = {f: annotations.get(f, Any) for f in fields}
    new_fields[extra_field_name] = extra_field_type
    cls = types.new_class(f"Enhanced{base_class.__name__}", (NamedTuple,), {},
                          lambda ns: ns.update({"__annotations__": new_fields, "__module__": base_class.__module__,
                                                "__doc__": f"{base_class.__name__} with the extra field '{extra_field_name}'."}))
    all_fields = tuple(cls._fields)

    def custom_iter(self: Any) -> Any:
        return iter(getattr(self, f) for f in fields)

    def custom_replace(self: Any, **kwargs: Any) -> Any:
        values = {f: getattr(self, f) for f in all_fields}
        unknown = set(kwargs) - set(all_fields)
        if unknown:
            raise ValueError(f"Got unexpected field names: {sorted(unknown)}")
        values.update(kwargs)
        return cls(**values)

    cls.__iter__ = custom_iter  # type: ignore[assignment]
    cls._replace = custom_replace  # type: ignore[assignment]
    return cls


_ENHANCED: Dict[type, type] = {}


def create_with_running_statistics(state: Any, running_statistics: RunningStatisticsState) -> Any:
    """running_statistics.py:533-560: the learner state with a `running_statistics` attribute."""
    base = type(state)
    if base not in _ENHANCED:
        _ENHANCED[base] = add_field_to_state(base, "running_statistics", RunningStatisticsState)
    cls = _ENHANCED[base]
    return cls(**{f: getattr(state, f) for f in base._fields}, running_statistics=running_statistics)
