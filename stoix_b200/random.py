"""Minimal host-side PRNG-key plumbing (stand-in for jax.random.PRNGKey / split at
ff_ppo.py:570-572, 492-494, 515-518).  Keys are uint64 Python ints; device-side randomness is
Philox4x32-10 inside the kernels, keyed by these values (bit-parity with JAX's threefry streams is
out of scope, SURVEY.md A.7: parity tests inject actions and permutations)."""
from __future__ import annotations

from typing import List

_MASK = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _MASK
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
    return z ^ (z >> 31)


def PRNGKey(seed: int) -> int:
    return _splitmix64(int(seed) & _MASK)


def split(key: int, num: int = 2) -> List[int]:
    return [_splitmix64((int(key) + (i + 1) * 0xD1B54A32D192ED03) & _MASK) for i in range(num)]
